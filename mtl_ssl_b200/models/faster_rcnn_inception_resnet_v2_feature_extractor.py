"""Faster R-CNN Inception-ResNet-v2 feature extractor
(/root/reference/object_detection/models/faster_rcnn_inception_resnet_v2_feature_extractor.py:31-198)."""
from .faster_rcnn_resnet_v1_feature_extractor import FasterRCNNFeatureExtractor
from ..nets import inception_resnet_v2


class FasterRCNNInceptionResnetV2FeatureExtractor(FasterRCNNFeatureExtractor):
    def __init__(self, is_training, first_stage_features_stride, reuse_weights=None, batch_norm_trainable=False,
                 weight_decay=0.0, base_features="block3", freeze_layer=""):
        if first_stage_features_stride != 8 and first_stage_features_stride != 16:
            raise ValueError("`first_stage_features_stride` must be 8 or 16.")
        if first_stage_features_stride != 16:
            raise ValueError("B200 path: first_stage_features_stride 16 only")
        super(FasterRCNNInceptionResnetV2FeatureExtractor, self).__init__(
            is_training, first_stage_features_stride, reuse_weights, weight_decay, freeze_layer,
            batch_norm_trainable)
        self._architecture = "InceptionResnetV2"
        self._trunks, self._tails = {}, {}
        self.feature_depth = 1088
        self.classifier_depth = 1536
        self.feature_mask_hi = 0.0
        self.supports_dx_extra = False

    def preprocess(self, resized_inputs):
        """fe:60-74 maps pixels to [-1, 1]; fused into the first layer's im2col kernel here."""
        return resized_inputs

    def create_proposal_variables(self, store, scope):
        self._trunks[scope] = inception_resnet_v2.InceptionResnetV2Trunk(store, scope + "/InceptionResnetV2",
                                                                         self._weight_decay, self._is_training)

    def create_box_classifier_variables(self, store, scope, trainable=None):
        t = self._is_training if trainable is None else trainable
        self._tails[scope] = inception_resnet_v2.InceptionResnetV2Tail(store, scope + "/InceptionResnetV2",
                                                                       self._weight_decay, t)

    def create_dead_variables(self, store, scope):
        """inception_resnet_v2_base stops at PreAuxLogits in stage 1: no unused variables."""

    def feature_map_shape(self, H, W):
        return next(iter(self._trunks.values())).out_hw(H, W)

    def extract_proposal_features(self, preprocessed_inputs, scope, ws):
        if preprocessed_inputs.dim() != 4:
            raise ValueError("`preprocessed_inputs` must be 4 dimensional, got a tensor of shape %s"
                             % (tuple(preprocessed_inputs.shape),))
        return self._trunks[scope].fwd(preprocessed_inputs, ws)

    def backward_proposal_features(self, scope, grad, ws):
        self._trunks[scope].bwd(grad, ws)

    def extract_box_classifier_features(self, proposal_feature_maps, scope, ws, tag="main", keep=True):
        return self._tails[scope].fwd(proposal_feature_maps, ws, tag, keep)

    def backward_box_classifier_features(self, scope, grad, ws, tag="main", need_dx=True, dx_extra=None,
                                         pre_unit0=None):
        return self._tails[scope].bwd(grad, ws, tag, need_dx, dx_extra, pre_unit0)
