"""Box predictors on the device, mirroring /root/reference/object_detection/core/box_predictor.py:
`ConvolutionalBoxPredictor` (:613-755, the RPN 1x1 heads) and `MaskRCNNBoxPredictor` (:339-611,
spatial average + FC heads; the fork's `predict_class`).  Each predictor owns one fused GEMM
operand per scope: the box and class weight matrices are stored back to back in the parameter
arena (runtime.ParamStore.add_group), so one tcgen05 GEMM produces both outputs while the
optimizer still clips BoxEncodingPredictor/weights and ClassPredictor/weights separately.
"""
import torch

from .. import ops
from .. import ops_conv as oc
from ..nets.layers import Concurrency, PooledTail
from .standard_fields import (BOX_ENCODINGS, CLASS_PREDICTIONS, CLASS_PREDICTIONS_WITH_BACKGROUND)


# MaskRCNNBoxPredictor: fused pool + FC forward and fused backward (csrc/head.cu); MTL_NO_FUSED_HEAD=1 restores the
# separate avgpool / tcgen05 GEMM / cast / colsum / avgpool_bwd launches (kept for A/B measurements)
FUSED_HEAD = not __import__("os").environ.get("MTL_NO_FUSED_HEAD")
FUSED_HEAD_MAX_OUT = 128      # wider heads (K = 90: 456 columns) re-read too many weight bytes per ROI: tensor-core GEMM


def _round8(n):
    return (n + 7) // 8 * 8


class _FusedHead(object):
    """[sum(outs) (+pad to 8), cin] weight group + bias group, fprop / wgrad / dgrad via the conv engine."""

    def __init__(self, store, scope, names_outs, cin, hp, trainable, fc=False):
        """fc=True: the reference builds these outputs with slim.fully_connected (variables [cin, n], bp:482-496)."""
        self.outs = [n for _, n in names_outs]
        self.n_out = sum(self.outs)
        self.n_pad = _round8(self.n_out)
        self.cin = cin
        self.trainable = trainable
        wspecs, bspecs = [], []
        for name, n in names_outs:
            wspecs.append(dict(name="%s/%s/weights" % (scope, name), shape=(n, 1, 1, cin), l2=hp.l2_weight,
                               trainable=trainable, init=hp.init, tf_kind="fc" if fc else None))
            bspecs.append(dict(name="%s/%s/biases" % (scope, name), shape=(n,), l2=0.0, trainable=trainable))
        if self.n_pad > self.n_out:
            wspecs.append(dict(name="%s/_pad/weights" % scope, shape=(self.n_pad - self.n_out, 1, 1, cin),
                               trainable=False))
            bspecs.append(dict(name="%s/_pad/biases" % scope, shape=(self.n_pad - self.n_out,), trainable=False))
        self.wgroup = store.add_group(wspecs)
        self.bgroup = store.add_group(bspecs)
        self.store = store

    def w_bf16(self):
        return self.store.group_view(self.wgroup, self.n_pad, self.cin, "wb").view(self.n_pad, 1, 1, self.cin)

    def w_grad(self):
        return self.store.group_view(self.wgroup, self.n_pad, self.cin, "g").view(self.n_pad, 1, 1, self.cin)

    def bias(self):
        return self.store.group_view(self.bgroup, 1, self.n_pad, "w").view(self.n_pad)

    def bias_grad(self):
        return self.store.group_view(self.bgroup, 1, self.n_pad, "g").view(self.n_pad)

    def fwd(self, x, out):
        """x bf16 [N,H,W,cin] -> out fp32 [N,H,W,n_pad]"""
        return oc.conv_fprop(x, self.w_bf16(), bias=self.bias(), out=out)

    def bwd(self, x, dy_bf16, dx=None, dx_mask=None):
        if self.trainable:
            with torch.cuda.stream(Concurrency.fork()):
                oc.conv_wgrad(dy_bf16, x, self.w_grad())
                rows = dy_bf16.numel() // self.n_pad
                ops.call("mtl_colsum", dy_bf16, 0, self.n_pad, rows, self.n_pad, 1.0, self.bias_grad())
        if dx is not None:
            oc.conv_dgrad(dy_bf16, self.w_bf16(), x.shape, mask=dx_mask, out=dx)
        return dx


class BoxPredictor(object):
    def __init__(self, is_training, num_classes):
        self._is_training = is_training
        self._num_classes = num_classes

    @property
    def num_classes(self):
        return self._num_classes


class ConvolutionalBoxPredictor(BoxPredictor):
    """RPN predictor: 1x1 conv -> A*4 box codes and 1x1 conv -> A*(num_classes+1) logits, no
    activation, with biases (bp:682-755).  Output buffer: fp32 [B,H,W, A*4 + A*2 (+pad)]."""

    def __init__(self, is_training, num_classes, conv_hyperparams, min_depth=0, max_depth=0,
                 num_layers_before_predictor=0, use_dropout=False, dropout_keep_prob=1.0, kernel_size=1,
                 box_code_size=4, apply_sigmoid_to_scores=False):
        super(ConvolutionalBoxPredictor, self).__init__(is_training, num_classes)
        if num_layers_before_predictor or use_dropout or apply_sigmoid_to_scores or kernel_size != 1:
            raise ValueError("B200 path: ConvolutionalBoxPredictor supports the RPN configuration only "
                             "(kernel 1, no extra layers / dropout / sigmoid)")
        self._hp = conv_hyperparams
        self._box_code_size = box_code_size
        self._heads = {}

    def create_variables(self, store, scope, in_channels, num_predictions_per_location):
        a = num_predictions_per_location
        self._heads[scope] = (_FusedHead(store, scope, [("BoxEncodingPredictor", a * self._box_code_size),
                                                         ("ClassPredictor", a * (self._num_classes + 1))],
                                         in_channels, self._hp, self._is_training), a)

    def layout(self, scope):
        head, a = self._heads[scope]
        return dict(ld=head.n_pad, box_col0=0, cls_col0=a * self._box_code_size, A=a)

    def predict(self, image_features, num_predictions_per_location, scope, out=None):
        head, a = self._heads[scope]
        assert a == num_predictions_per_location
        B, H, W, _ = image_features.shape
        if out is None:
            out = torch.empty(B, H, W, head.n_pad, dtype=torch.float32, device=image_features.device)
        head.fwd(image_features, out)
        nb = a * self._box_code_size
        k1 = self._num_classes + 1
        return {"_raw": out,
                BOX_ENCODINGS: lambda: out[..., :nb].reshape(B, H * W * a, 1, self._box_code_size),
                CLASS_PREDICTIONS_WITH_BACKGROUND: lambda: out[..., nb:nb + a * k1].reshape(B, H * W * a, k1)}

    def backward(self, scope, image_features, d_out_bf16, dx, dx_mask):
        head, _ = self._heads[scope]
        return head.bwd(image_features, d_out_bf16, dx, dx_mask)


class MaskRCNNBoxPredictor(BoxPredictor):
    """Spatial average over the ROI grid, then FC -> K*4 box codes and FC -> K+1 logits
    (bp:430-528); `predict_class` (bp:530-611) emits only `num_classes` logits (the aux predictors
    are built with num_classes = K+1, model_builder.py:288-303)."""

    def __init__(self, is_training, num_classes, fc_hyperparams, use_dropout=False, dropout_keep_prob=1.0,
                 box_code_size=4, conv_hyperparams=None, predict_instance_masks=False,
                 mask_prediction_conv_depth=256, predict_keypoints=False, spatial_average=False,
                 box_initializer=None):
        super(MaskRCNNBoxPredictor, self).__init__(is_training, num_classes)
        if predict_instance_masks or predict_keypoints:
            raise ValueError("Mask / keypoint prediction is not supported on the B200 path")
        if use_dropout:
            raise ValueError("use_dropout is not supported on the B200 path (all shipped configs disable it)")
        if not spatial_average:
            raise ValueError("B200 path requires spatial_average: true (as in every shipped config)")
        self._hp = fc_hyperparams
        self._box_code_size = box_code_size
        self._heads = {}
        self._saved = {}
        self._feature_mask_hi = 0.0     # 6.0 when the ROI features come out of a ReLU6 (MobileNet tail)

    def create_variables(self, store, scope, in_channels, class_only=False):
        if class_only:
            outs = [("ClassPredictor", self._num_classes)]
        else:
            outs = [("BoxEncodingPredictor", self._num_classes * self._box_code_size),
                    ("ClassPredictor", self._num_classes + 1)]
        self._heads[scope] = _FusedHead(store, scope, outs, in_channels, self._hp, self._is_training, fc=True)

    def accepts_pooled_tail(self, scope):
        """The fused head kernel serves this head (narrow enough): a forward-only tail may skip writing its maps."""
        return FUSED_HEAD and self._heads[scope].n_pad <= FUSED_HEAD_MAX_OUT

    def layout(self, scope):
        head = self._heads[scope]
        if len(head.outs) == 1:
            return dict(ld=head.n_pad, cls_col0=0, num=head.outs[0])
        return dict(ld=head.n_pad, box_col0=0, cls_col0=head.outs[0], num=head.outs[1])

    def _run(self, image_features, scope, ws, tag):
        head = self._heads[scope]
        R, H, W, C = image_features.shape
        pooled = ws.get("%s/%s/pooled" % (scope, tag), (R, 1, 1, C))
        out = ws.get("%s/%s/head_out" % (scope, tag), (R, 1, 1, head.n_pad), torch.float32)
        if isinstance(image_features, PooledTail):
            # forward-only tail whose last conv already summed the ROI grid (never stored its output)
            ops.call("mtl_head_fwd_pooled", image_features.part, R, H * W, C, head.w_bf16(), head.bias(), head.n_pad,
                     pooled, out, head.n_pad)
            self._saved[(scope, tag)] = None
            return out.view(R, head.n_pad)
        if FUSED_HEAD and head.n_pad <= FUSED_HEAD_MAX_OUT:
            # spatial average + both FC layers in one kernel per ROI batch (csrc/head.cu)
            ops.call("mtl_head_fwd", image_features, R, H * W, C, head.w_bf16(), head.bias(), head.n_pad, pooled, out,
                     head.n_pad)
        else:
            ops.call("mtl_avgpool_fwd", image_features, R, H * W, C, pooled)
            head.fwd(pooled, out)
        self._saved[(scope, tag)] = (image_features, pooled)
        return out.view(R, head.n_pad)

    def predict(self, image_features, num_predictions_per_location, scope, ws=None, tag="main", **params):
        if num_predictions_per_location != 1:
            raise ValueError("Currently FullyConnectedBoxPredictor only supports predicting a single box per "
                             "class per location.")
        out = self._run(image_features, scope, ws, tag)
        R = out.shape[0]
        nb = self._num_classes * self._box_code_size
        k1 = self._num_classes + 1
        return {"_raw": out,
                BOX_ENCODINGS: lambda: out[:, :nb].reshape(R, 1, self._num_classes, self._box_code_size),
                CLASS_PREDICTIONS_WITH_BACKGROUND: lambda: out[:, nb:nb + k1].reshape(R, 1, k1)}

    def predict_class(self, image_features, scope, ws=None, tag="main", activation_fn=None, with_background=False):
        if activation_fn is not None:
            raise ValueError("predict_class: only activation_fn=None is used by the reference and supported")
        out = self._run(image_features, scope, ws, tag)
        R = out.shape[0]
        n = self._heads[scope].outs[0]
        key = CLASS_PREDICTIONS_WITH_BACKGROUND if with_background else CLASS_PREDICTIONS
        return {"_raw": out, key: lambda: out[:, :n].reshape(R, 1, n)}

    def backward(self, scope, tag, d_out, ws, need_dx=True):
        """d_out fp32 [R, ld] -> gradient w.r.t. image_features (bf16, masked by features > 0)."""
        head = self._heads[scope]
        feats, pooled = self._saved[(scope, tag)]
        R, H, W, C = feats.shape
        dyb = ws.get("%s/%s/d_head_bf16" % (scope, tag), (R, 1, 1, head.n_pad))
        if FUSED_HEAD and head.n_pad <= FUSED_HEAD_MAX_OUT and d_out.stride(-1) == 1 and d_out.shape[-1] == head.n_pad:
            # logit gradient -> bf16 operand of the weight-gradient GEMM, bias gradient, pooled-feature gradient and
            # its broadcast over the ROI grid under the ReLU mask: one kernel; the weight gradient stays a GEMM
            g = ws.get("%s/%s/d_feat" % (scope, tag), feats.shape) if need_dx else None
            ops.call("mtl_head_bwd", d_out, d_out.stride(0), head.n_pad, head.w_bf16(), feats, self._feature_mask_hi,
                     R, H * W, C, dyb, head.bias_grad() if head.trainable else None, g)
            if head.trainable:
                with torch.cuda.stream(Concurrency.fork()):
                    oc.conv_wgrad(dyb, pooled, head.w_grad())
            return g
        ops.call("mtl_cast_f32_bf16", d_out, d_out.numel(), 1.0, dyb)
        dpool = ws.get("%s/%s/d_pooled" % (scope, tag), (R, 1, 1, C)) if need_dx else None
        head.bwd(pooled, dyb, dpool)
        if not need_dx:
            return None
        g = ws.get("%s/%s/d_feat" % (scope, tag), feats.shape)
        ops.call("mtl_avgpool_bwd", dpool, 0, C, feats, self._feature_mask_hi, R, H * W, C, g)
        return g


class RfcnBoxPredictor(BoxPredictor):
    """R-FCN predictor (/root/reference/object_detection/core/box_predictor.py:131-337): 1x1
    `reduce_depth` conv (+bias, activation of the conv hyperparams), 1x1 `refined_locations`
    (bins*K*4) and `class_predictions` (bins*(K+1)) maps without activation, then position-sensitive
    ROI pooling with global average (utils/ops.py:462-609).  The two map convs are one fused GEMM;
    pooling reads the bf16 maps with one block per ROI (mtl_psroi_fwd / mtl_psroi_bwd)."""

    def __init__(self, is_training, num_classes, conv_hyperparams, num_spatial_bins, depth, crop_size,
                 box_code_size=4):
        super(RfcnBoxPredictor, self).__init__(is_training, num_classes)
        self._hp = conv_hyperparams
        self._num_spatial_bins = list(num_spatial_bins)
        self._depth = depth
        self._crop_size = list(crop_size)
        self._box_code_size = box_code_size
        self._vars = {}
        self._saved = {}
        self._feature_mask_hi = 0.0

    def create_variables(self, store, scope, in_channels, class_only=False):
        from ..nets.layers import Conv2d
        hp = self._hp
        if hp.activation == "RELU_6":
            raise ValueError("RELU_6 in the R-FCN predictor is not supported on the B200 path")
        reduce = Conv2d(store, scope + "/reduce_depth", in_channels, self._depth, 1, 1, bn=False, bias=True,
                        relu=(hp.activation != "NONE"), l2=hp.l2_weight, trainable=self._is_training, init=hp.init)
        bins = self._num_spatial_bins[0] * self._num_spatial_bins[1]
        if class_only:
            groups = [("class_predictions", self._num_classes)]
        else:
            groups = [("refined_locations", self._num_classes * self._box_code_size),
                      ("class_predictions", self._num_classes + 1)]
        head = _FusedHead(store, scope, [(n, bins * d) for n, d in groups], self._depth, hp, self._is_training)
        self._vars[scope] = (reduce, head, groups)

    def layout(self, scope):
        _, _, groups = self._vars[scope]
        ld = _round8(sum(d for _, d in groups))
        if len(groups) == 1:
            return dict(ld=ld, cls_col0=0, num=groups[0][1])
        return dict(ld=ld, box_col0=0, cls_col0=groups[0][1], num=groups[1][1])

    def _pool(self, scope, tag, src_tag, boxes, box_ind, ws):
        reduce, head, groups = self._vars[scope]
        feats, red, pmap = self._saved[(scope, src_tag)][:3]
        B, H, W, _ = feats.shape
        R = boxes.shape[0]
        ld = self.layout(scope)["ld"]
        out = ws.get("%s/%s/head_out" % (scope, tag), (R, ld), torch.float32)
        bins = self._num_spatial_bins[0] * self._num_spatial_bins[1]
        c0 = ocol = 0
        for _, d in groups:
            ops.call("mtl_psroi_fwd", pmap, B, H, W, head.n_pad, c0, d, self._num_spatial_bins[0],
                     self._num_spatial_bins[1], self._crop_size[0], self._crop_size[1], boxes, box_ind, R, out, ld, ocol)
            c0 += bins * d
            ocol += d
        self._saved[(scope, tag)] = (feats, red, pmap, boxes, box_ind)
        return out

    def _run(self, image_features, boxes, box_ind, scope, ws, tag):
        reduce, head, groups = self._vars[scope]
        B, H, W, C = image_features.shape
        red = reduce.fwd(image_features, ws.get("%s/%s/reduce" % (scope, tag), (B, H, W, self._depth)))
        pmap = ws.get("%s/%s/ps_maps" % (scope, tag), (B, H, W, head.n_pad))
        head.fwd(red, pmap)
        self._saved[(scope, tag)] = (image_features, red, pmap, boxes, box_ind)
        return self._pool(scope, tag, tag, boxes, box_ind, ws)

    def predict(self, image_features, num_predictions_per_location, scope, proposal_boxes=None, box_ind=None,
                ws=None, tag="main", **params):
        if num_predictions_per_location != 1:
            raise ValueError("Currently RfcnBoxPredictor only supports predicting a single box per class per "
                             "location.")
        out = self._run(image_features, proposal_boxes, box_ind, scope, ws, tag)
        R = out.shape[0]
        nb, k1 = self._num_classes * self._box_code_size, self._num_classes + 1
        return {"_raw": out,
                BOX_ENCODINGS: lambda: out[:, :nb].reshape(R, 1, self._num_classes, self._box_code_size),
                CLASS_PREDICTIONS_WITH_BACKGROUND: lambda: out[:, nb:nb + k1].reshape(R, 1, k1)}

    def predict_class(self, image_features, scope, proposal_boxes=None, box_ind=None, ws=None, tag="main",
                      activation_fn=None, with_background=False, reuse_maps_of=None):
        """reuse_maps_of: tag of an earlier call on the SAME features and scope whose position-sensitive
        maps are pooled again with new boxes (the reference recomputes block4 + maps for the refine
        windows, rfcn_meta_arch.py:312-381; the values are identical)."""
        if reuse_maps_of is not None:
            out = self._pool(scope, tag, reuse_maps_of, proposal_boxes, box_ind, ws)
        else:
            out = self._run(image_features, proposal_boxes, box_ind, scope, ws, tag)
        R, n = out.shape[0], self._num_classes
        key = CLASS_PREDICTIONS_WITH_BACKGROUND if with_background else CLASS_PREDICTIONS
        return {"_raw": out, key: lambda: out[:, :n].reshape(R, 1, n)}

    def backward(self, scope, tag, d_out, ws, need_dx=True):
        """d_out fp32 [R, ld] -> gradient w.r.t. image_features (bf16, masked by features > 0)."""
        reduce, head, groups = self._vars[scope]
        feats, red, pmap, boxes, box_ind = self._saved[(scope, tag)]
        B, H, W, C = feats.shape
        R = boxes.shape[0]
        ld = self.layout(scope)["ld"]
        dmap = ws.get("%s/%s/d_ps_maps_f32" % (scope, tag), pmap.shape, torch.float32, zero=True)
        bins = self._num_spatial_bins[0] * self._num_spatial_bins[1]
        c0 = ocol = 0
        for _, d in groups:
            ops.call("mtl_psroi_bwd", d_out, ld, ocol, B, H, W, head.n_pad, c0, d, self._num_spatial_bins[0],
                     self._num_spatial_bins[1], self._crop_size[0], self._crop_size[1], boxes, box_ind, R, dmap)
            c0 += bins * d
            ocol += d
        dmb = ws.get("%s/%s/d_ps_maps" % (scope, tag), pmap.shape)
        ops.call("mtl_cast_f32_bf16", dmap, dmap.numel(), 1.0, dmb)
        dred = ws.get("%s/%s/d_reduce" % (scope, tag), red.shape)
        head.bwd(red, dmb, dred, red if reduce.relu else None)
        reduce.wgrad(feats, dred)
        if not need_dx:
            return None
        return reduce.dgrad(dred, feats.shape, ws.get("%s/%s/d_feat" % (scope, tag), feats.shape), mask=feats)
