"""Inception-ResNet-v2 for Faster R-CNN on the tcgen05 conv engine, forward and explicit backward.

Structure follows /root/reference/slim/nets/inception_resnet_v2.py: block35 :33-52, block17 :55-72,
block8 :75-92, inception_resnet_v2_base :94-268 (stage 1 = everything up to 'PreAuxLogits' with
align_feature_maps=True, i.e. SAME padding everywhere, output stride 16, 1088 channels) and the
second-stage tail of object_detection/models/faster_rcnn_inception_resnet_v2_feature_extractor.py
(:112-170: Mixed_7a with VALID stride-2 convs, 9 x block8(0.2), block8 without activation,
Conv2d_7b_1x1 -> 1536).  inception_resnet_v2_arg_scope (:334-360): conv + batch norm WITHOUT gamma
(eps 1e-3, inference mode -> folded) + ReLU, L2 on weights and biases, xavier initialisation; the
block `up` convolutions have biases and no normaliser and are scaled by the block's residual scale.

Branch outputs are written by the conv epilogues straight into channel slices of the concat buffer,
and gradients are read back from slices of the concat gradient (no copy kernels).  A tiny node graph
(Seq / Branches / ResUnit) sequences forward and backward explicitly."""
import torch

from .layers import Conv2d, same_pad, max_pool, max_pool_bwd, max_pool_out_hw
from .. import ops

EPS = 1e-3
INIT = ("xavier",)


class ConvNode(object):
    def __init__(self, store, scope, cin, cout, k=1, stride=1, padding="SAME", l2=0.0, trainable=True, mask_x=True):
        self.conv = Conv2d(store, scope, cin, cout, k, stride, 1, padding, bn=True, relu=True, l2=l2,
                           trainable=trainable, init=INIT, bn_eps=EPS, bn_scale=False)
        self.cout = cout
        self.mask_x = mask_x
        self.saved = {}

    def out_shape(self, shp):
        N, H, W, _ = shp
        P, Q, _, _ = self.conv.geom(H, W)
        return (N, P, Q, self.cout)

    def fwd(self, x, ws, key, out=None):
        if out is None:
            out = ws.get(self.conv.scope + "/" + key, self.out_shape(x.shape))
        self.saved[key] = x
        return self.conv.fwd(x, out)

    def bwd(self, dy, ws, key, dx_res=None, need_dx=True):
        x = self.saved[key]
        self.conv.wgrad(x, dy)
        if not need_dx:
            return None
        return self.conv.dgrad(dy, x.shape, ws.get(self.conv.scope + "/" + key + "/dx", x.shape), res=dx_res,
                               mask=x if self.mask_x else None)


class PoolNode(object):
    """3x3 max pool (stride 2, SAME or VALID) or the 3x3/1 SAME average pool of Mixed_5b."""

    def __init__(self, scope, kind, stride=2, padding="SAME"):
        self.scope, self.kind, self.stride, self.padding = scope, kind, stride, padding
        self.saved = {}

    def out_shape(self, shp):
        N, H, W, C = shp
        if self.kind == "avg":
            return shp
        P, Q = max_pool_out_hw(H, W, 3, self.stride, self.padding)
        return (N, P, Q, C)

    def fwd(self, x, ws, key, out=None):
        if out is None:
            out = ws.get(self.scope + "/" + key, self.out_shape(x.shape))
        self.saved[key] = x
        if self.kind == "avg":
            N, H, W, C = x.shape
            assert out.is_contiguous()
            ops.call("mtl_avgpool3x3_same", x, N, H, W, C, 0, out)
            return out
        return max_pool(x, out, 3, self.stride, self.padding)

    def bwd(self, dy, ws, key, dx_res=None, need_dx=True):
        assert dx_res is None, "pool nodes must be the first gradient producer of their fan-in"
        x = self.saved[key]
        dx = ws.get(self.scope + "/" + key + "/dx", x.shape)
        if self.kind == "avg":
            N, H, W, C = x.shape
            assert dy.is_contiguous()
            ops.call("mtl_avgpool3x3_same", dy, N, H, W, C, 1, dx)
            return dx
        return max_pool_bwd(x, dy, dx, 3, self.stride, self.padding)


class Seq(object):
    def __init__(self, nodes):
        self.nodes = nodes

    @property
    def cout(self):
        return None

    def out_shape(self, shp):
        for n in self.nodes:
            shp = n.out_shape(shp)
        return shp

    def fwd(self, x, ws, key, out=None):
        for i, n in enumerate(self.nodes):
            x = n.fwd(x, ws, key, out if i == len(self.nodes) - 1 else None)
        return x

    def bwd(self, dy, ws, key, dx_res=None, need_dx=True):
        for i in range(len(self.nodes) - 1, -1, -1):
            first = i == 0
            dy = self.nodes[i].bwd(dy, ws, key, dx_res if first else None, need_dx or not first)
        return dy

    def starts_with_pool(self):
        return isinstance(self.nodes[0], PoolNode)


class Branches(object):
    """Parallel branches on the same input, concatenated along channels (tf.concat axis 3)."""

    def __init__(self, scope, branches):
        self.scope, self.branches = scope, branches
        self.saved = {}

    def out_shape(self, shp):
        shapes = [b.out_shape(shp) for b in self.branches]
        return shapes[0][:3] + (sum(s[3] for s in shapes),)

    def fwd(self, x, ws, key, out=None):
        shp = self.out_shape(x.shape)
        if out is None:
            out = ws.get(self.scope + "/concat/" + key, shp)
        c0 = 0
        slices = []
        for b in self.branches:
            c = b.out_shape(x.shape)[3]
            b.fwd(x, ws, key, out[..., c0:c0 + c])
            slices.append((c0, c))
            c0 += c
        self.saved[key] = slices
        return out

    def bwd(self, dcat, ws, key, dx_res=None, need_dx=True):
        slices = self.saved[key]
        order = sorted(range(len(self.branches)), key=lambda i: 0 if self.branches[i].starts_with_pool() else 1)
        acc = dx_res
        for j, i in enumerate(order):
            c0, c = slices[i]
            b = self.branches[i]
            if b.starts_with_pool() and len(b.nodes) == 1:
                assert acc is None or acc is dx_res
                d = b.bwd(dcat[..., c0:c0 + c], ws, key, None, need_dx)
                if acc is not None:          # identity-path gradient + pooled branch: merge through the next dgrad
                    raise NotImplementedError("pool-only branch combined with an incoming residual gradient")
                acc = d
            elif b.starts_with_pool():
                d = b.bwd(dcat[..., c0:c0 + c], ws, key, None, need_dx)
                assert acc is None
                acc = d
            else:
                acc = b.bwd(dcat[..., c0:c0 + c], ws, key, acc, need_dx)
        return acc


class ResUnit(object):
    """block35 / block17 / block8: net = act(net + scale * up(concat(branches)))."""

    def __init__(self, store, scope, cin, branches, cat_channels, scale, act=True, l2=0.0, trainable=True):
        self.scope = scope
        self.branches = Branches(scope, branches)
        self.up = Conv2d(store, scope + "/Conv2d_1x1", cat_channels, cin, 1, 1, 1, "SAME", bn=False, bias=True,
                         relu=act, l2=l2, bias_l2=l2, trainable=trainable, init=INIT, out_scale=scale)
        self.saved = {}

    def out_shape(self, shp):
        return shp

    def fwd(self, x, ws, key, out=None):
        mixed = self.branches.fwd(x, ws, key)
        if out is None:
            out = ws.get(self.scope + "/out/" + key, x.shape)
        self.saved[key] = (x, mixed)
        return self.up.fwd(mixed, out, res=x)

    def bwd(self, g, ws, key, dx_res=None, need_dx=True):
        """g: gradient w.r.t. the unit output, already masked by the consumer (out > 0 when activated)."""
        assert dx_res is None
        x, mixed = self.saved[key]
        self.up.wgrad(mixed, g)
        dmixed = self.up.dgrad(g, mixed.shape, ws.get(self.scope + "/dmixed/" + key, mixed.shape), mask=mixed)
        return self.branches.bwd(dmixed, ws, key, dx_res=g, need_dx=need_dx)

    def starts_with_pool(self):
        return False


def _block35(store, scope, l2, t):
    c = lambda s, ci, co, k=1: ConvNode(store, scope + "/" + s, ci, co, k, l2=l2, trainable=t)
    br = [Seq([c("Branch_0/Conv2d_1x1", 320, 32)]),
          Seq([c("Branch_1/Conv2d_0a_1x1", 320, 32), c("Branch_1/Conv2d_0b_3x3", 32, 32, 3)]),
          Seq([c("Branch_2/Conv2d_0a_1x1", 320, 32), c("Branch_2/Conv2d_0b_3x3", 32, 48, 3),
               c("Branch_2/Conv2d_0c_3x3", 48, 64, 3)])]
    return ResUnit(store, scope, 320, br, 128, 0.17, True, l2, t)


def _block17(store, scope, l2, t):
    c = lambda s, ci, co, k=1: ConvNode(store, scope + "/" + s, ci, co, k, l2=l2, trainable=t)
    br = [Seq([c("Branch_0/Conv2d_1x1", 1088, 192)]),
          Seq([c("Branch_1/Conv2d_0a_1x1", 1088, 128), c("Branch_1/Conv2d_0b_1x7", 128, 160, (1, 7)),
               c("Branch_1/Conv2d_0c_7x1", 160, 192, (7, 1))])]
    return ResUnit(store, scope, 1088, br, 384, 0.10, True, l2, t)


def _block8(store, scope, l2, t, scale=0.20, act=True):
    c = lambda s, ci, co, k=1: ConvNode(store, scope + "/" + s, ci, co, k, l2=l2, trainable=t)
    br = [Seq([c("Branch_0/Conv2d_1x1", 2080, 192)]),
          Seq([c("Branch_1/Conv2d_0a_1x1", 2080, 192), c("Branch_1/Conv2d_0b_1x3", 192, 224, (1, 3)),
               c("Branch_1/Conv2d_0c_3x1", 224, 256, (3, 1))])]
    return ResUnit(store, scope, 2080, br, 448, scale, act, l2, t)


class InceptionResnetV2Trunk(object):
    def __init__(self, store, scope, l2, trainable=True):
        self.scope = scope
        t = trainable
        c = lambda s, ci, co, k=1, st=1, **kw: ConvNode(store, scope + "/" + s, ci, co, k, st, "SAME", l2, t, **kw)
        # Conv2d_1a_3x3 (3x3/2 on RGB): GEMM over im2col rows of 64 = 27 + zero pad (packed variable layout)
        self.conv1a = Conv2d(store, scope + "/Conv2d_1a_3x3", 64, 32, 1, 1, relu=True, l2=l2, trainable=t,
                             init=("packed_conv", 27, 0.1), bn_eps=EPS, bn_scale=False)
        self.conv1a.weight.tf_kind = ("packed_conv", 3, 3, 3)      # TF: Conv2d_1a_3x3/weights [3,3,3,32]
        nodes = [c("Conv2d_2a_3x3", 32, 32, 3), c("Conv2d_2b_3x3", 32, 64, 3),
                 PoolNode(scope + "/MaxPool_3a_3x3", "max", 2, "SAME"),
                 c("Conv2d_3b_1x1", 64, 80), c("Conv2d_4a_3x3", 80, 192, 3),
                 PoolNode(scope + "/MaxPool_5a_3x3", "max", 2, "SAME")]
        m5 = scope + "/Mixed_5b"
        nodes.append(Branches(m5, [
            Seq([c("Mixed_5b/Branch_0/Conv2d_1x1", 192, 96)]),
            Seq([c("Mixed_5b/Branch_1/Conv2d_0a_1x1", 192, 48), c("Mixed_5b/Branch_1/Conv2d_0b_5x5", 48, 64, 5)]),
            Seq([c("Mixed_5b/Branch_2/Conv2d_0a_1x1", 192, 64), c("Mixed_5b/Branch_2/Conv2d_0b_3x3", 64, 96, 3),
                 c("Mixed_5b/Branch_2/Conv2d_0c_3x3", 96, 96, 3)]),
            Seq([PoolNode(m5 + "/Branch_3/AvgPool_0a_3x3", "avg"), c("Mixed_5b/Branch_3/Conv2d_0b_1x1", 192, 64)])]))
        for i in range(10):
            nodes.append(_block35(store, "%s/Repeat/block35_%d" % (scope, i + 1), l2, t))
        m6 = scope + "/Mixed_6a"
        nodes.append(Branches(m6, [
            Seq([c("Mixed_6a/Branch_0/Conv2d_1a_3x3", 320, 384, 3, 2)]),
            Seq([c("Mixed_6a/Branch_1/Conv2d_0a_1x1", 320, 256), c("Mixed_6a/Branch_1/Conv2d_0b_3x3", 256, 256, 3),
                 c("Mixed_6a/Branch_1/Conv2d_1a_3x3", 256, 384, 3, 2)]),
            Seq([PoolNode(m6 + "/Branch_2/MaxPool_1a_3x3", "max", 2, "SAME")])]))
        for i in range(20):
            nodes.append(_block17(store, "%s/Repeat_1/block17_%d" % (scope, i + 1), l2, t))
        self.body = Seq(nodes)
        self.out_channels = 1088

    def out_hw(self, H, W):
        shp = (1, same_pad(H, 3, 2)[0], same_pad(W, 3, 2)[0], 32)
        return self.body.out_shape(shp)[1:3]

    def fwd(self, img, ws):
        B, H, W, _ = img.shape
        P, ph = same_pad(H, 3, 2)
        Q, pw = same_pad(W, 3, 2)
        rows = ws.get(self.scope + "/im2col", (B, P, Q, 64))
        ops.call("mtl_im2col_f32", img, B, H, W, 3, 3, 3, 2, ph, pw, P, Q, [127.5, 127.5, 127.5], 2.0 / 255.0, rows, 64)
        x = self.conv1a.fwd(rows, ws.get(self.scope + "/c1a", (B, P, Q, 32)))
        self.saved = (rows, x)
        return self.body.fwd(x, ws, "s1")

    def bwd(self, g, ws):
        g = self.body.bwd(g, ws, "s1", None, True)
        self.conv1a.wgrad(self.saved[0], g)


class InceptionResnetV2Tail(object):
    def __init__(self, store, scope, l2, trainable=True):
        self.scope = scope
        t = trainable
        c = lambda s, ci, co, k=1, st=1, pad="SAME", **kw: ConvNode(store, scope + "/" + s, ci, co, k, st, pad, l2, t,
                                                                     **kw)
        m7 = scope + "/Mixed_7a"
        nodes = [Branches(m7, [
            Seq([c("Mixed_7a/Branch_0/Conv2d_0a_1x1", 1088, 256, mask_x=False),
                 c("Mixed_7a/Branch_0/Conv2d_1a_3x3", 256, 384, 3, 2, "VALID")]),
            Seq([c("Mixed_7a/Branch_1/Conv2d_0a_1x1", 1088, 256, mask_x=False),
                 c("Mixed_7a/Branch_1/Conv2d_1a_3x3", 256, 288, 3, 2, "VALID")]),
            Seq([c("Mixed_7a/Branch_2/Conv2d_0a_1x1", 1088, 256, mask_x=False),
                 c("Mixed_7a/Branch_2/Conv2d_0b_3x3", 256, 288, 3),
                 c("Mixed_7a/Branch_2/Conv2d_1a_3x3", 288, 320, 3, 2, "VALID")]),
            Seq([PoolNode(m7 + "/Branch_3/MaxPool_1a_3x3", "max", 2, "VALID")])])]
        for i in range(9):
            nodes.append(_block8(store, "%s/Repeat_2/block8_%d" % (scope, i + 1), l2, t))
        nodes.append(_block8(store, scope + "/Block8", l2, t, scale=1.0, act=False))
        nodes.append(c("Conv2d_7b_1x1", 2080, 1536, mask_x=False))
        self.body = Seq(nodes)
        self.out_channels = 1536

    def fwd(self, x, ws, tag, keep=True):
        return self.body.fwd(x, ws, tag)

    def bwd(self, g, ws, tag, need_dx=True, dx_extra=None, pre_unit0=None):
        nodes = self.body.nodes
        for i in range(len(nodes) - 1, 0, -1):
            g = nodes[i].bwd(g, ws, tag, None, True)
        if pre_unit0 is not None:
            pre_unit0()
        if dx_extra is not None:
            raise NotImplementedError
        return nodes[0].bwd(g, ws, tag, None, need_dx)
