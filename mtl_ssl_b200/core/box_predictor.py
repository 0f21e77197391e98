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
from ..nets.layers import Concurrency
from .standard_fields import (BOX_ENCODINGS, CLASS_PREDICTIONS, CLASS_PREDICTIONS_WITH_BACKGROUND)


def _round8(n):
    return (n + 7) // 8 * 8


class _FusedHead(object):
    """[sum(outs) (+pad to 8), cin] weight group + bias group, fprop / wgrad / dgrad via the conv engine."""

    def __init__(self, store, scope, names_outs, cin, hp, trainable):
        self.outs = [n for _, n in names_outs]
        self.n_out = sum(self.outs)
        self.n_pad = _round8(self.n_out)
        self.cin = cin
        self.trainable = trainable
        wspecs, bspecs = [], []
        for name, n in names_outs:
            wspecs.append(dict(name="%s/%s/weights" % (scope, name), shape=(n, 1, 1, cin), l2=hp.l2_weight,
                               trainable=trainable, init=hp.init))
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
        self._heads[scope] = _FusedHead(store, scope, outs, in_channels, self._hp, self._is_training)

    def layout(self, scope):
        head = self._heads[scope]
        if len(head.outs) == 1:
            return dict(ld=head.n_pad, cls_col0=0, num=head.outs[0])
        return dict(ld=head.n_pad, box_col0=0, cls_col0=head.outs[0], num=head.outs[1])

    def _run(self, image_features, scope, ws, tag):
        head = self._heads[scope]
        R, H, W, C = image_features.shape
        pooled = ws.get("%s/%s/pooled" % (scope, tag), (R, 1, 1, C))
        ops.call("mtl_avgpool_fwd", image_features, R, H * W, C, pooled)
        out = ws.get("%s/%s/head_out" % (scope, tag), (R, 1, 1, head.n_pad), torch.float32)
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
        ops.call("mtl_cast_f32_bf16", d_out, d_out.numel(), 1.0, dyb)
        dpool = ws.get("%s/%s/d_pooled" % (scope, tag), (R, 1, 1, C)) if need_dx else None
        head.bwd(pooled, dyb, dpool)
        if not need_dx:
            return None
        g = ws.get("%s/%s/d_feat" % (scope, tag), feats.shape)
        ops.call("mtl_avgpool_bwd", dpool, 0, C, feats, self._feature_mask_hi, R, H * W, C, g)
        return g
