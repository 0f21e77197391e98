"""Layer objects with explicit forward / backward over the tcgen05 conv engine.

A layer owns its parameters (in the ParamStore arena) and exposes
  fwd(x, tag)        -> y            (activations live in the Workspace, keyed by scope+tag)
  wgrad(x, dy)                        (accumulates into the fp32 gradient arena)
  dgrad(dy, ...)     -> dx
There is no autograd tape: the callers (nets/resnet_v1.py, the meta-architecture) sequence the
backward pass explicitly, which is what lets one training step be a fixed list of kernel
launches (CUDA-graph capturable).  Mirrors slim.conv2d + slim.batch_norm(is_training=False) +
activation as configured by resnet_arg_scope (/root/reference/slim/nets/resnet_utils.py:203-256).
"""
import torch

from .. import ops_conv as oc
from .. import ops


class Concurrency(object):
    """Fork/join helper: weight-gradient GEMMs are off the backward critical path (they only feed the
    optimizer), so they are launched on side streams and overlap the dgrad chain.  Under CUDA-graph
    capture the stream waits become graph edges, i.e. the captured step keeps the concurrency."""
    enabled = True
    num_streams = int(__import__("os").environ.get("MTL_WGRAD_STREAMS", "4"))
    _streams = None
    _idx = 0

    @classmethod
    def fork(cls):
        cur = torch.cuda.current_stream()
        if not cls.enabled:
            return cur
        if cls._streams is None:
            cls._streams = [torch.cuda.Stream() for _ in range(cls.num_streams)]
        s = cls._streams[cls._idx % len(cls._streams)]
        cls._idx += 1
        s.wait_stream(cur)
        return s

    @classmethod
    def join(cls):
        if cls._streams is None:
            return
        cur = torch.cuda.current_stream()
        for s in cls._streams:
            cur.wait_stream(s)


def same_pad(in_size, k, stride, rate=1):
    """TensorFlow SAME padding: (out, pad_begin)."""
    out = -(-in_size // stride)
    ke = k + (k - 1) * (rate - 1)
    total = max((out - 1) * stride + ke - in_size, 0)
    return out, total // 2


class PooledTail(object):
    """What a forward-only box-classifier tail hands to the box predictor when its last conv fused the spatial mean
    (ops_conv.conv_fprop pool_out): the partial row sums instead of the [R, H, W, C] feature maps, which are never
    written (the refine pass of the MTL step: 1 280 ROIs x 49 x 2048 bf16 = 257 MB, written once and read once)."""

    def __init__(self, part, shape):
        self.part, self.shape = part, tuple(shape)


class Conv2d(object):
    """slim.conv2d (+ frozen batch norm folded, or bias) + activation.  `k` is an int or (rows, cols);
    `out_scale` multiplies the whole pre-activation output (Inception-ResNet `net += scale * up`): it is
    folded into the bf16 weights like a batch-norm scale and applied to the bias in the epilogue."""

    def __init__(self, store, scope, cin, cout, k=1, stride=1, rate=1, padding="SAME", bn=True, bias=False,
                 relu=True, l2=1e-4, trainable=True, init=("variance_scaling",), bn_eps=1e-5, weight_name="weights",
                 bn_scale=True, bias_l2=0.0, out_scale=None):
        self.scope = scope
        self.kh, self.kw = (k, k) if isinstance(k, int) else k
        self.k = self.kh
        self.cin, self.cout, self.stride, self.rate = cin, cout, stride, rate
        self.padding = padding          # "SAME" | "VALID" | "EXPLICIT" (resnet_utils.conv2d_same)
        self.relu = relu                # False | True (ReLU) | 2 (ReLU6)
        self.trainable = trainable
        self.out_scale = out_scale
        self.bn = store.add_bn(scope + "/BatchNorm", cout, bn_eps, bn_scale) if bn else None
        self.fold = self.bn
        if out_scale is not None:
            assert not bn
            self.fold = store.add_bn(None, cout, 0.0)           # virtual: constant per-channel scale
            self.fold.gamma = self.fold.gamma * float(out_scale)
        self.weight = store.add(scope + "/" + weight_name, (cout, self.kh, self.kw, cin), l2=l2, trainable=trainable,
                                init=init, fold=self.fold)
        self.bias = store.add(scope + "/biases", (cout,), l2=bias_l2, trainable=trainable) if bias else None

    # geometry -------------------------------------------------------------------------------
    def geom(self, H, W):
        s, r = self.stride, self.rate
        if self.padding == "SAME":
            P, ph = same_pad(H, self.kh, s, r)
            Q, pw = same_pad(W, self.kw, s, r)
        elif self.padding == "VALID":
            keh, kew = self.kh + (self.kh - 1) * (r - 1), self.kw + (self.kw - 1) * (r - 1)
            P, Q, ph, pw = (H - keh) // s + 1, (W - kew) // s + 1, 0, 0
        else:   # conv2d_same with stride > 1: pad (ke-1)//2 before, rest after, then VALID
            ke = self.kh + (self.kh - 1) * (r - 1)
            ph = pw = (ke - 1) // 2
            P, Q = (H + ke - 1 - ke) // s + 1, (W + ke - 1 - ke) // s + 1
        return P, Q, ph, pw

    def epilogue_bias(self):
        if self.bn is not None:
            return self.bn.bias
        return self.bias.w if self.bias is not None else None

    # kernels --------------------------------------------------------------------------------
    def fwd(self, x, out, res=None, relu=None, pool_out=None, pool_hw=0):
        N, H, W, C = x.shape
        P, Q, ph, pw = self.geom(H, W)
        assert tuple(out.shape) == (N, P, Q, self.cout), (out.shape, (N, P, Q, self.cout))
        return oc.conv_fprop(x, self.weight.wb, self.stride, (ph, pw), self.rate, (P, Q),
                             bias=self.epilogue_bias(), res=res, relu=int(self.relu if relu is None else relu),
                             out=out, bias_scale=self.out_scale if self.out_scale is not None else 1.0,
                             pool_out=pool_out, pool_hw=pool_hw)

    def wgrad(self, x, dy):
        if not self.trainable:
            return
        N, H, W, C = x.shape
        P, Q, ph, pw = self.geom(H, W)
        with torch.cuda.stream(Concurrency.fork()):
            oc.conv_wgrad(dy, x, self.weight.g, self.stride, (ph, pw), self.rate,
                          rowscale=self.fold.scale if self.fold is not None else None)
            if self.bias is not None:
                rows = dy.shape[0] * dy.shape[1] * dy.shape[2]
                ops.call("mtl_colsum", dy, 0, oc._pitch(dy), rows, self.cout,
                         self.out_scale if self.out_scale is not None else 1.0, self.bias.g)

    def dgrad(self, dy, x_shape, out, res=None, mask=None, mask_hi=0.0):
        N, H, W, C = x_shape
        P, Q, ph, pw = self.geom(H, W)
        return oc.conv_dgrad(dy, self.weight.wb, x_shape, self.stride, (ph, pw), self.rate, res=res, mask=mask,
                             out=out, mask_hi=mask_hi)


class DepthwiseConv3x3(object):
    """Depthwise stage of slim.separable_conv2d (depth_multiplier 1), SAME padding; weights [C,3,3]
    (TF depthwise_weights [3,3,C,1] transposed), optional folded batch norm + ReLU6."""

    def __init__(self, store, scope, channels, stride=1, bn=True, act=2, l2=0.0, trainable=True,
                 init=("truncated_normal", 0.09), bn_eps=1e-3):
        self.scope, self.C, self.stride, self.act, self.trainable = scope, channels, stride, act, trainable
        self.bn = store.add_bn(scope + "/BatchNorm", channels, bn_eps) if bn else None
        self.weight = store.add(scope + "/depthwise_weights", (channels, 3, 3), l2=l2, trainable=trainable,
                                init=init, fold=self.bn)

    def geom(self, H, W):
        P, ph = same_pad(H, 3, self.stride)
        Q, pw = same_pad(W, 3, self.stride)
        return P, Q, ph, pw

    def fwd(self, x, out):
        N, H, W, C = x.shape
        P, Q, ph, pw = self.geom(H, W)
        ops.call("mtl_dwconv3x3_fwd", x, self.weight.wb, self.bn.bias if self.bn is not None else None, N, H, W, C,
                 self.stride, ph, pw, P, Q, self.act, out)
        return out

    def wgrad(self, x, dy):
        if not self.trainable:
            return
        N, H, W, C = x.shape
        P, Q, ph, pw = self.geom(H, W)
        with torch.cuda.stream(Concurrency.fork()):
            ops.call("mtl_dwconv3x3_wgrad", dy, x, N, H, W, C, self.stride, ph, pw, P, Q,
                     self.bn.scale if self.bn is not None else None, self.weight.g)

    def dgrad(self, dy, x_shape, out, mask=None, mask_hi=0.0):
        N, H, W, C = x_shape
        P, Q, ph, pw = self.geom(H, W)
        ops.call("mtl_dwconv3x3_dgrad", dy, self.weight.wb, N, H, W, C, self.stride, ph, pw, P, Q, mask, mask_hi, out)
        return out


def max_pool(x, out, k, stride, padding="SAME"):
    N, H, W, C = x.shape
    if padding == "SAME":
        P, ph = same_pad(H, k, stride)
        Q, pw = same_pad(W, k, stride)
    else:
        P, Q, ph, pw = (H - k) // stride + 1, (W - k) // stride + 1, 0, 0
    assert tuple(out.shape) == (N, P, Q, C)
    ops.call("mtl_maxpool_fwd", x, N, H, W, C, k, stride, ph, pw, P, Q, out, out.stride(2) if Q > 1 else 0)
    return out


def max_pool_out_hw(H, W, k, stride, padding="SAME"):
    if padding == "SAME":
        return same_pad(H, k, stride)[0], same_pad(W, k, stride)[0]
    return (H - k) // stride + 1, (W - k) // stride + 1


def max_pool_bwd(x, dy, dx, k, stride, padding="SAME"):
    N, H, W, C = x.shape
    _, P, Q, _ = dy.shape
    ph = same_pad(H, k, stride)[1] if padding == "SAME" else 0
    pw = same_pad(W, k, stride)[1] if padding == "SAME" else 0
    ops.call("mtl_maxpool_bwd", x, dy, dy.stride(2) if Q > 1 else 0, N, H, W, C, k, stride, ph, pw, P, Q, dx)
    return dx
