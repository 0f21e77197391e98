"""Faster R-CNN MobileNet-v1 feature extractor
(/root/reference/object_detection/models/faster_rcnn_mobilenet_v1_feature_extractor.py:53-184)."""
from .faster_rcnn_resnet_v1_feature_extractor import FasterRCNNFeatureExtractor
from ..nets import mobilenet_v1


class FasterRCNNMobilenetV1FeatureExtractor(FasterRCNNFeatureExtractor):
    def __init__(self, is_training, first_stage_features_stride, reuse_weights=None, batch_norm_trainable=False,
                 weight_decay=0.0, depth_multiplier=1.0, min_depth=16, skip_last_stride=False,
                 conv_depth_ratio_in_percentage=100, freeze_layer=""):
        if first_stage_features_stride != 8 and first_stage_features_stride != 16:
            raise ValueError("`first_stage_features_stride` must be 8 or 16.")
        if first_stage_features_stride != 16 or depth_multiplier != 1.0 or skip_last_stride:
            raise ValueError("B200 path: MobileNet-v1 with stride 16, depth multiplier 1, no skip_last_stride only")
        super(FasterRCNNMobilenetV1FeatureExtractor, self).__init__(
            is_training, first_stage_features_stride, reuse_weights, weight_decay, freeze_layer,
            batch_norm_trainable)
        self._architecture = "MobilenetV1"
        self._trunks, self._tails = {}, {}
        self.feature_depth = 512
        self.classifier_depth = 1024
        self.feature_mask_hi = 6.0           # activations are ReLU6
        self.supports_dx_extra = False

    def preprocess(self, resized_inputs):
        """fe:96-108 maps pixels to [-1, 1]; fused into the first layer's im2col kernel here."""
        return resized_inputs

    def create_proposal_variables(self, store, scope):
        self._trunks[scope] = mobilenet_v1.MobilenetV1Trunk(store, scope + "/MobilenetV1", self._weight_decay,
                                                            self._is_training)

    def create_box_classifier_variables(self, store, scope, trainable=None):
        t = self._is_training if trainable is None else trainable
        self._tails[scope] = mobilenet_v1.MobilenetV1Tail(store, scope + "/MobilenetV1", self.feature_depth, t)

    def create_dead_variables(self, store, scope):
        """mobilenet_v1_base stops at Conv2d_11_pointwise: no unused variables exist (unlike ResNet)."""

    def feature_map_shape(self, H, W):
        return next(iter(self._trunks.values())).out_hw(H, W)

    def extract_proposal_features(self, preprocessed_inputs, scope, ws):
        if preprocessed_inputs.shape[1] < 33 or preprocessed_inputs.shape[2] < 33:
            raise ValueError("image size must at least be 33 in both height and width.")
        return self._trunks[scope].fwd(preprocessed_inputs, ws)

    def backward_proposal_features(self, scope, grad, ws):
        self._trunks[scope].bwd(grad, ws)

    def extract_box_classifier_features(self, proposal_feature_maps, scope, ws, tag="main", keep=True):
        return self._tails[scope].fwd(proposal_feature_maps, ws, tag, keep)

    def backward_box_classifier_features(self, scope, grad, ws, tag="main", need_dx=True, dx_extra=None,
                                         pre_unit0=None):
        return self._tails[scope].bwd(grad, ws, tag, need_dx, dx_extra, pre_unit0)
