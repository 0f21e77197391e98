"""ResNet-v1 (bottleneck) on the tcgen05 conv engine, forward and explicit backward.

Structure follows /root/reference/slim/nets/resnet_v1.py:69-130 (`bottleneck`: stride on the
3x3 conv2, shortcut = 1x1 conv when the depth changes else a 1x1 max-pool subsample,
resnet_utils.py:59-74), :133-237 (`resnet_v1` root block: conv2d_same 7x7/2 with frozen weights,
3x3/2 SAME max pool) and resnet_utils.py:126-200 (`stack_blocks_dense`: once the requested
output stride is reached, later strides become atrous rates).  Batch norm always runs in
inference mode (fe:139) and is folded into the conv weights / epilogue bias.
"""
import torch

from .layers import Conv2d, PooledTail, max_pool, max_pool_bwd, max_pool_out_hw
from .. import ops
from .. import ops_conv as oc

BLOCKS = {
    "resnet_v1_50": [("block1", 256, 64, 3, 2), ("block2", 512, 128, 4, 2), ("block3", 1024, 256, 6, 2),
                     ("block4", 2048, 512, 3, 1)],
    "resnet_v1_101": [("block1", 256, 64, 3, 2), ("block2", 512, 128, 4, 2), ("block3", 1024, 256, 23, 2),
                      ("block4", 2048, 512, 3, 1)],
    "resnet_v1_152": [("block1", 256, 64, 3, 2), ("block2", 512, 128, 8, 2), ("block3", 1024, 256, 36, 2),
                      ("block4", 2048, 512, 3, 1)],
}


class Bottleneck(object):
    def __init__(self, store, scope, cin, depth, depth_bottleneck, stride, rate=1, l2=1e-4, trainable=True):
        self.scope = scope
        self.cin, self.depth, self.stride = cin, depth, stride
        s = scope + "/bottleneck_v1"
        kw = dict(l2=l2, trainable=trainable)
        self.shortcut = None
        if depth != cin:
            self.shortcut = Conv2d(store, s + "/shortcut", cin, depth, 1, stride, relu=False, **kw)
        self.conv1 = Conv2d(store, s + "/conv1", cin, depth_bottleneck, 1, 1, **kw)
        self.conv2 = Conv2d(store, s + "/conv2", depth_bottleneck, depth_bottleneck, 3, stride, rate,
                            padding="SAME" if stride == 1 else "EXPLICIT", **kw)
        self.conv3 = Conv2d(store, s + "/conv3", depth_bottleneck, depth, 1, 1, relu=False, **kw)
        self.trainable = trainable

    def out_hw(self, H, W):
        P, Q, _, _ = self.conv2.geom(H, W)
        return P, Q

    def fwd(self, x, ws, tag, keep=True, parity=0, pool=False):
        """keep=False: forward-only pass, intermediates share scratch buffers across units.  pool (forward-only): the
        unit's output is only ever averaged over the ROI grid -- conv3 writes partial row sums instead (PooledTail)."""
        N, H, W, C = x.shape
        P, Q = self.out_hw(H, W)
        key = (self.scope if keep else "scratch") + "/" + tag
        db = self.conv1.cout
        if self.shortcut is not None:
            sc = self.shortcut.fwd(x, ws.get(key + "/sc", (N, P, Q, self.depth)))
        elif self.stride == 1:
            sc = x
        else:
            sc = max_pool(x, ws.get(key + "/sc", (N, P, Q, self.depth)), 1, self.stride)
        r1 = self.conv1.fwd(x, ws.get(key + "/r1", (N, H, W, db)))
        r2 = self.conv2.fwd(r1, ws.get(key + "/r2", (N, P, Q, db)))
        okey = (self.scope + "/" + tag + "/out") if keep else ("scratch/" + tag + "/out%d" % parity)
        if pool and P * Q >= 32 and self.depth % 64 == 0:      # (what the kernel's pooled output supports)
            assert not keep
            part = ws.get("scratch/%s/pool_part" % tag, (oc.pool_partial_rows(N * P * Q), self.depth), torch.float32)
            self.conv3.fwd(r2, ws.get(okey, (N, P, Q, self.depth)), res=sc, relu=True, pool_out=part, pool_hw=P * Q)
            return PooledTail(part, (N, P, Q, self.depth))
        out = self.conv3.fwd(r2, ws.get(okey, (N, P, Q, self.depth)), res=sc, relu=True)
        if keep:
            self.saved = getattr(self, "saved", {})
            self.saved[tag] = (x, r1, r2, out)
        return out

    def bwd(self, g, ws, tag, need_dx=True, mask_x=True, dx_extra=None):
        """g: gradient w.r.t. the pre-ReLU sum (already masked by out > 0).  Returns the gradient
        w.r.t. x, masked by x > 0 when x is itself a ReLU output (mask_x)."""
        x, r1, r2, out = self.saved[tag]
        key = self.scope + "/" + tag
        self.conv3.wgrad(r2, g)
        g2 = self.conv3.dgrad(g, r2.shape, ws.get(key + "/g2", r2.shape), mask=r2)
        self.conv2.wgrad(r1, g2)
        g1 = self.conv2.dgrad(g2, r1.shape, ws.get(key + "/g1", r1.shape), mask=r1)
        self.conv1.wgrad(x, g1)
        if self.shortcut is not None:
            self.shortcut.wgrad(x, g)
        if not need_dx:
            return None
        if self.shortcut is not None:
            res = self.shortcut.dgrad(g, x.shape, ws.get(key + "/dsc", x.shape), res=dx_extra)
        elif self.stride == 1:
            assert dx_extra is None
            res = g
        else:
            assert dx_extra is None
            res = max_pool_bwd(x, g, ws.get(key + "/dsc", x.shape), 1, self.stride)
        return self.conv1.dgrad(g1, x.shape, ws.get(key + "/dx", x.shape), res=res, mask=x if mask_x else None)


class Stem(object):
    """conv1 7x7/2 (explicit pad 3 + VALID, frozen: resnet_v1.py:216-221) + 3x3/2 SAME max pool.
    The 3-channel input is expanded by mtl_im2col_f32 (mean subtraction fused) into rows of
    160 = 147 + pad bf16, then conv1 is a plain GEMM on the tensor cores."""

    def __init__(self, store, scope, l2=1e-4, means=(123.68, 116.779, 103.939), scale=1.0):
        self.scope = scope
        self.means, self.scale = means, scale
        self.ld = 160
        self.bn = store.add_bn(scope + "/conv1/BatchNorm", 64, 1e-5)
        # stored as [64, 7, 7, 3] like every other conv (TF HWIO transposed); the GEMM operand
        # [64, 160] is packed in the bf16 arena by pack()
        self.weight = store.add(scope + "/conv1/weights", (64, 7, 7, 3), l2=l2, trainable=False,
                                init=("variance_scaling",), fold=self.bn)
        self.packed = None
        store.post_load_hooks.append(self.invalidate)

    def invalidate(self):
        self.packed = None

    def pack(self, device):
        w = torch.zeros(64, self.ld, dtype=torch.bfloat16, device=device)
        w[:, :147] = self.weight.wb.reshape(64, 147)
        self.packed = w.view(64, 1, 1, self.ld)

    def out_hw(self, H, W):
        P, Q = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        return max_pool_out_hw(P, Q, 3, 2)

    def fwd(self, img, ws, tag=""):
        B, H, W, C = img.shape
        P, Q = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
        if self.packed is None:
            self.pack(img.device)
        rows = ws.get(self.scope + tag + "/im2col", (B, P, Q, self.ld))
        ops.call("mtl_im2col_f32", img, B, H, W, 3, 7, 7, 2, 3, 3, P, Q, list(self.means), self.scale, rows,
                 self.ld)
        c1 = ws.get(self.scope + tag + "/conv1", (B, P, Q, 64))
        oc.conv_fprop(rows, self.packed, bias=self.bn.bias, relu=True, out=c1)
        P2, Q2 = max_pool_out_hw(P, Q, 3, 2)
        return max_pool(c1, ws.get(self.scope + tag + "/pool1", (B, P2, Q2, 64)), 3, 2)


class ResNetV1(object):
    """Stage-1 trunk up to `last_block` with output_stride 16 (fe:92-146)."""

    def __init__(self, store, scope, arch="resnet_v1_101", l2=1e-4, n_freeze_blocks=1, last_block="block3",
                 output_stride=16, means=(123.68, 116.779, 103.939)):
        self.scope = scope
        self.stem = Stem(store, scope, l2, means)
        self.units = []
        cin = 64
        current_stride, rate = 4, 1
        for bi, (bname, depth, db, n, bstride) in enumerate(BLOCKS[arch]):
            trainable = bi >= n_freeze_blocks
            for u in range(n):
                stride = bstride if u == n - 1 else 1
                if current_stride == output_stride:
                    unit = Bottleneck(store, "%s/%s/unit_%d" % (scope, bname, u + 1), cin, depth, db, 1, rate,
                                      l2, trainable)
                    rate *= stride
                else:
                    unit = Bottleneck(store, "%s/%s/unit_%d" % (scope, bname, u + 1), cin, depth, db, stride, 1,
                                      l2, trainable)
                    current_stride *= stride
                unit.block = bname
                self.units.append(unit)
                cin = depth
            if bname == last_block:
                break
        self.out_channels = cin

    def num_frozen_units(self):
        return next((i for i, u in enumerate(self.units) if u.trainable), len(self.units))

    def fwd_prefix(self, img, ws, tag="s1"):
        """Stem + the frozen leading units (freeze_layer, fe:117-121).  Their output depends on the image only,
        never on a weight update, so the trainer may compute it for the NEXT batch while the current step is
        still in its backward pass (`tag` selects a separate set of activation buffers for that)."""
        x = self.stem.fwd(img, ws, "" if tag == "s1" else "/" + tag)
        for u in self.units[:self.num_frozen_units()]:
            x = u.fwd(x, ws, tag)
        return x

    def fwd(self, img, ws, prefix=None):
        x = self.fwd_prefix(img, ws) if prefix is None else prefix
        for u in self.units[self.num_frozen_units():]:
            x = u.fwd(x, ws, "s1")
        return x

    def bwd(self, g, ws, every=0, checkpoint=None, part=None):
        """g: gradient w.r.t. the trunk output, already masked by (output > 0).  `checkpoint(j)` is called after every
        `every` units (deferred weight-gradient work is handed to a side stream in chunks while the chain goes on).
        part="hi" stops after the unit `split_unit()` (its input gradient is kept), part="lo" resumes there: with
        several replicas the gradient bucket of the units already done is exchanged while the rest still computes."""
        first_trainable = next(i for i, u in enumerate(self.units) if u.trainable)
        last = len(self.units) - 1
        split = self.split_unit()
        hi = last if part in (None, "hi") else split - 1
        lo = first_trainable if part in (None, "lo") else split
        if part == "lo":
            g = self._g_split
        for i in range(hi, lo - 1, -1):
            g = self.units[i].bwd(g, ws, "s1", need_dx=i > first_trainable)
            n = last - i + 1
            if every and checkpoint is not None and n % every == 0 and i > first_trainable:
                checkpoint(n // every)
        if part == "hi":
            self._g_split = g
        return None

    def split_unit(self):
        """Index of the first unit of the "hi" half of the backward pass: the middle of the trainable units."""
        first_trainable = next(i for i, u in enumerate(self.units) if u.trainable)
        return first_trainable + (len(self.units) - first_trainable + 1) // 2


class Block4(object):
    """Second-stage per-ROI tail: block4 = 3 bottlenecks 1024->2048, stride 1 (fe:148-185)."""

    def __init__(self, store, scope, l2=1e-4, cin=1024, trainable=True):
        self.scope = scope
        self.units = []
        for u in range(3):
            self.units.append(Bottleneck(store, "%s/block4/unit_%d" % (scope, u + 1), cin, 2048, 512, 1, 1, l2,
                                         trainable))
            cin = 2048
        self.out_channels = 2048

    def fwd(self, x, ws, tag, keep=True, pool=False):
        for i, u in enumerate(self.units):
            x = u.fwd(x, ws, tag, keep, i % 2, pool=pool and not keep and i == len(self.units) - 1)
        return x

    def bwd(self, g, ws, tag, need_dx=True, dx_extra=None, pre_unit0=None):
        for i in (2, 1):
            g = self.units[i].bwd(g, ws, tag)
        if pre_unit0 is not None:
            pre_unit0()
        return self.units[0].bwd(g, ws, tag, need_dx=need_dx, mask_x=False, dx_extra=dx_extra)
