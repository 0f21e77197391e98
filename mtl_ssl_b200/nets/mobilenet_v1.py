"""MobileNet-v1 (depth multiplier 1) for Faster R-CNN, forward and explicit backward.

Stage 1 = `mobilenet_v1_base(..., final_endpoint='Conv2d_11_pointwise')` (slim/nets/mobilenet_v1.py:
142-266): Conv2d_0 3x3/2 -> 11 x (depthwise 3x3 + BN + ReLU6, pointwise 1x1 + BN + ReLU6), stride 16,
512 channels.  Stage 2 = two fused slim.separable_conv2d layers (depthwise 3x3 WITHOUT normaliser or
activation, pointwise 1x1 + BN + ReLU6), the first with stride 2
(object_detection/models/faster_rcnn_mobilenet_v1_feature_extractor.py:148-184).  Batch norm runs in
inference mode (eps 1e-3) and is folded; only slim.conv2d weights are L2-regularised
(mobilenet_v1_arg_scope :376-413: separable_conv2d gets depthwise_regularizer = None).
Depthwise layers are HBM-bound CUDA-core kernels, pointwise layers run on the tcgen05 engine."""
import torch

from .layers import Conv2d, DepthwiseConv3x3, same_pad
from .. import ops

# (depth, stride) of DepthSepConv layers 1..13 (mobilenet_v1.py:124-139)
DEFS = [(64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2), (512, 1), (512, 1), (512, 1), (512, 1), (512, 1),
        (1024, 2), (1024, 1)]
INIT = ("truncated_normal", 0.09)


class MobilenetV1Trunk(object):
    def __init__(self, store, scope, l2, trainable=True):
        self.scope = scope
        # Conv2d_0: 3x3/2 on 3 channels, run as a GEMM over im2col rows of 64 = 27 + zero pad bf16; the
        # variable is stored in that packed [32, 64] layout (pad columns stay exactly zero: zero input,
        # zero gradient, zero L2)
        self.conv0 = Conv2d(store, scope + "/Conv2d_0", 64, 32, 1, 1, relu=2, l2=l2, trainable=trainable, init=INIT,
                            bn_eps=1e-3)
        self.conv0.weight.init = ("packed_conv", 27, 0.09)
        self.conv0.weight.tf_kind = ("packed_conv", 3, 3, 3)       # TF: MobilenetV1/Conv2d_0/weights [3,3,3,32]
        self.layers = []
        cin = 32
        for i, (depth, stride) in enumerate(DEFS[:11]):
            s = "%s/Conv2d_%d" % (scope, i + 1)
            dw = DepthwiseConv3x3(store, s + "_depthwise", cin, stride, bn=True, act=2, l2=0.0, trainable=trainable)
            pw = Conv2d(store, s + "_pointwise", cin, depth, 1, 1, relu=2, l2=l2, trainable=trainable, init=INIT,
                        bn_eps=1e-3)
            self.layers.append((dw, pw))
            cin = depth
        self.out_channels = cin

    def out_hw(self, H, W):
        h, w = same_pad(H, 3, 2)[0], same_pad(W, 3, 2)[0]
        for dw, _ in self.layers:
            h, w = dw.geom(h, w)[:2]
        return h, w

    def fwd(self, img, ws):
        B, H, W, _ = img.shape
        P, ph = same_pad(H, 3, 2)
        Q, pw_ = same_pad(W, 3, 2)
        rows = ws.get(self.scope + "/im2col", (B, P, Q, 64))
        # preprocess (fe:96-108): (2/255) x - 1, fused into the im2col pass
        ops.call("mtl_im2col_f32", img, B, H, W, 3, 3, 3, 2, ph, pw_, P, Q, [127.5, 127.5, 127.5], 2.0 / 255.0, rows,
                 64)
        x = self.conv0.fwd(rows, ws.get(self.scope + "/c0", (B, P, Q, 32)))
        self.saved = [rows, x]
        for i, (dw, pw) in enumerate(self.layers):
            N, h, w, c = x.shape
            p, q = dw.geom(h, w)[:2]
            d = dw.fwd(x, ws.get("%s/d%d" % (self.scope, i), (N, p, q, c)))
            x = pw.fwd(d, ws.get("%s/x%d" % (self.scope, i), (N, p, q, pw.cout)))
            self.saved += [d, x]
        return x

    def bwd(self, g, ws):
        """g: gradient w.r.t. the trunk output, already masked by 0 < out < 6."""
        sv = self.saved
        for i in range(len(self.layers) - 1, -1, -1):
            dw, pw = self.layers[i]
            xin, d = sv[1 + 2 * i], sv[2 + 2 * i]
            pw.wgrad(d, g)
            gd = pw.dgrad(g, d.shape, ws.get("%s/gd%d" % (self.scope, i), d.shape), mask=d, mask_hi=6.0)
            dw.wgrad(xin, gd)
            g = dw.dgrad(gd, xin.shape, ws.get("%s/gx%d" % (self.scope, i), xin.shape), mask=xin, mask_hi=6.0)
        self.conv0.wgrad(sv[0], g)


class MobilenetV1Tail(object):
    """Conv2d_12_pointwise (stride 2) and Conv2d_13_pointwise fused separable convs on ROI crops."""

    def __init__(self, store, scope, cin=512, trainable=True):
        self.scope = scope
        self.layers = []
        for name, stride in (("Conv2d_12_pointwise", 2), ("Conv2d_13_pointwise", 1)):
            s = scope + "/" + name
            dw = DepthwiseConv3x3(store, s, cin, stride, bn=False, act=0, l2=0.0, trainable=trainable)
            pw = Conv2d(store, s, cin, 1024, 1, 1, relu=2, l2=0.0, trainable=trainable, init=INIT, bn_eps=1e-3,
                        weight_name="pointwise_weights")
            self.layers.append((dw, pw))
            cin = 1024
        self.out_channels = 1024
        self.saved = {}

    def fwd(self, x, ws, tag, keep=True):
        sv = [x]
        for i, (dw, pw) in enumerate(self.layers):
            N, h, w, c = x.shape
            p, q = dw.geom(h, w)[:2]
            d = dw.fwd(x, ws.get("%s/%s/d%d" % (self.scope, tag, i), (N, p, q, c)))
            x = pw.fwd(d, ws.get("%s/%s/x%d" % (self.scope, tag, i), (N, p, q, pw.cout)))
            sv += [d, x]
        self.saved[tag] = sv
        return x

    def bwd(self, g, ws, tag, need_dx=True, dx_extra=None, pre_unit0=None):
        assert dx_extra is None
        sv = self.saved[tag]
        for i in (1, 0):
            dw, pw = self.layers[i]
            xin, d = sv[2 * i], sv[1 + 2 * i]
            pw.wgrad(d, g)
            gd = pw.dgrad(g, d.shape, ws.get("%s/%s/gd%d" % (self.scope, tag, i), d.shape))
            dw.wgrad(xin, gd)
            if i == 0 and not need_dx:
                return None
            g = dw.dgrad(gd, xin.shape, ws.get("%s/%s/gx%d" % (self.scope, tag, i), xin.shape),
                         mask=xin if i == 1 else None, mask_hi=6.0 if i == 1 else 0.0)
        return g
