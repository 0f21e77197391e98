"""Foreground-mask ("edgemask") predictor: one k x k conv with tanh, `num_classes * channels`
outputs, no batch norm (/root/reference/object_detection/core/mask_predictor.py:90-119; variable
scope 'BoxEncodingPredictor' as in the reference, mp:116-117).  Two output channels are far too
narrow for a tensor-core tile, so the head is a fused SIMT warp-per-pixel kernel."""
import torch

from .. import ops
from .standard_fields import MASK_PREDICTIONS


class MaskPredictor(object):
    def __init__(self, is_training, num_classes, conv_hyperparams, kernel_size=1, channels=1):
        if kernel_size != 1 or num_classes * channels != 2:
            raise ValueError("B200 path: edgemask predictor supports kernel_size 1 with 2 outputs "
                             "(every shipped config)")
        self._is_training = is_training
        self._num_classes = num_classes
        self._hp = conv_hyperparams
        self._vars = {}

    @property
    def num_classes(self):
        return self._num_classes

    def create_variables(self, store, scope, in_channels):
        w = store.add(scope + "/BoxEncodingPredictor/weights", (2, 1, 1, in_channels), l2=self._hp.l2_weight,
                      trainable=self._is_training, init=self._hp.init)
        b = store.add(scope + "/BoxEncodingPredictor/biases", (2,), trainable=self._is_training)
        self._vars[scope] = (w, b)

    def predict(self, image_features, scope, out=None):
        w, b = self._vars[scope]
        B, H, W, C = image_features.shape
        if out is None:
            out = torch.empty(B, H, W, 2, dtype=torch.float32, device=image_features.device)
        ops.call("mtl_edgemask_fwd", image_features, B * H * W, C, w.w, b.w, out)
        return {MASK_PREDICTIONS: out}

    def backward(self, scope, image_features, act, d_act, dfeat):
        w, b = self._vars[scope]
        B, H, W, C = image_features.shape
        ops.call("mtl_edgemask_bwd", image_features, B * H * W, C, w.w, act, d_act, dfeat, w.g, b.g)
