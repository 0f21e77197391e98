"""Faster R-CNN ResNet-v1 feature extractors on the tcgen05 conv engine, mirroring
/root/reference/object_detection/models/faster_rcnn_resnet_v1_feature_extractor.py:36-276 and the
abstract base at meta_architectures/faster_rcnn_meta_arch.py:95-205."""
from ..nets import resnet_v1


class FasterRCNNFeatureExtractor(object):
    def __init__(self, is_training, first_stage_features_stride, reuse_weights=None, weight_decay=0.0,
                 freeze_layer="", batch_norm_trainable=False):
        self._is_training = is_training
        self._first_stage_features_stride = first_stage_features_stride
        self._reuse_weights = reuse_weights
        self._weight_decay = weight_decay
        self._freeze_layer = freeze_layer
        self._batch_norm_trainable = batch_norm_trainable
        if batch_norm_trainable:
            raise ValueError("batch_norm_trainable is not supported: batch norm is folded into the conv "
                             "weights (frozen statistics, as in every shipped config)")


class FasterRCNNResnetV1FeatureExtractor(FasterRCNNFeatureExtractor):
    def __init__(self, architecture, resnet_model, is_training, first_stage_features_stride, reuse_weights=None,
                 weight_decay=0.0, freeze_layer="", batch_norm_trainable=False):
        if first_stage_features_stride != 8 and first_stage_features_stride != 16:
            raise ValueError("`first_stage_features_stride` must be 8 or 16.")
        if first_stage_features_stride != 16:
            raise ValueError("B200 path: first_stage_features_stride 16 only (no shipped config uses 8)")
        self._architecture = architecture
        self._base_features = "block3"
        super(FasterRCNNResnetV1FeatureExtractor, self).__init__(
            is_training, first_stage_features_stride, reuse_weights, weight_decay, freeze_layer,
            batch_norm_trainable)
        self._trunks = {}
        self._tails = {}
        self.means = (123.68, 116.779, 103.939)
        self.feature_depth = 1024
        self.classifier_depth = 2048
        self.feature_mask_hi = 0.0           # activations are plain ReLU
        self.supports_dx_extra = True
        self.deferred_wgrad_safe = True      # every gradient tensor of the trunk's backward pass has its own buffer

    def preprocess(self, resized_inputs):
        """fe:74-90 subtracts the ImageNet channel means.  On the B200 path the subtraction is
        fused into the stem's im2col kernel (mtl_im2col_f32), so this is the identity."""
        return resized_inputs

    # variables ------------------------------------------------------------------------------
    def create_proposal_variables(self, store, scope):
        n_freeze = int(self._freeze_layer[-1]) if self._freeze_layer else 0     # fe:117-121
        n_freeze = n_freeze if self._is_training else 4
        self._trunks[scope] = resnet_v1.ResNetV1(store, scope + "/" + self._architecture, self._architecture,
                                                 self._weight_decay, n_freeze, self._base_features, 16, self.means)

    def create_box_classifier_variables(self, store, scope, trainable=None):
        t = self._is_training if trainable is None else trainable
        self._tails[scope] = resnet_v1.Block4(store, scope + "/" + self._architecture, self._weight_decay,
                                              self.feature_depth, t)

    def create_dead_variables(self, store, scope):
        """The reference also instantiates block4 inside the stage-1 network (output unused, fe:145-146):
        the variables exist, are regularised and decay (trap T4)."""
        self.create_box_classifier_variables(store, scope + "/_dead")

    def feature_map_shape(self, H, W):
        h, w = self._trunk_any().stem.out_hw(H, W)
        for u in self._trunk_any().units:
            h, w = u.out_hw(h, w)
        return h, w

    def box_classifier_shape(self, h, w):
        return h, w

    def _trunk_any(self):
        return next(iter(self._trunks.values()))

    # forward / backward ---------------------------------------------------------------------
    def extract_frozen_prefix(self, preprocessed_inputs, scope, ws, tag="s1"):
        """Output of conv1 + the frozen blocks (None if nothing is frozen): feeds extract_proposal_features(prefix=)."""
        trunk = self._trunks[scope]
        if trunk.num_frozen_units() == 0:
            return None
        return trunk.fwd_prefix(preprocessed_inputs, ws, tag)

    def extract_proposal_features(self, preprocessed_inputs, scope, ws, prefix=None):
        if preprocessed_inputs.dim() != 4:
            raise ValueError("`preprocessed_inputs` must be 4 dimensional, got a tensor of shape %s"
                             % (tuple(preprocessed_inputs.shape),))
        if preprocessed_inputs.shape[1] < 33 or preprocessed_inputs.shape[2] < 33:
            raise ValueError("image size must at least be 33 in both height and width.")
        return self._trunks[scope].fwd(preprocessed_inputs, ws, prefix)

    def backward_proposal_features(self, scope, grad, ws, every=0, checkpoint=None, part=None):
        self._trunks[scope].bwd(grad, ws, every, checkpoint, part)

    def trunk_split_variable(self, scope):
        """Name prefix of the first variable of the "hi" half of the trunk's backward pass (see ResNetV1.bwd)."""
        t = self._trunks[scope]
        return t.units[t.split_unit()].scope + "/"

    supports_pooled_tail = True     # forward-only tails can hand the box predictor partial row sums (layers.PooledTail)

    def extract_box_classifier_features(self, proposal_feature_maps, scope, ws, tag="main", keep=True, pool=False):
        return self._tails[scope].fwd(proposal_feature_maps, ws, tag, keep, pool)

    def backward_box_classifier_features(self, scope, grad, ws, tag="main", need_dx=True, dx_extra=None,
                                         pre_unit0=None):
        return self._tails[scope].bwd(grad, ws, tag, need_dx, dx_extra, pre_unit0)


class FasterRCNNResnet50FeatureExtractor(FasterRCNNResnetV1FeatureExtractor):
    def __init__(self, is_training, first_stage_features_stride, reuse_weights=None, weight_decay=0.0,
                 freeze_layer="", batch_norm_trainable=False):
        super(FasterRCNNResnet50FeatureExtractor, self).__init__(
            "resnet_v1_50", None, is_training, first_stage_features_stride, reuse_weights, weight_decay,
            freeze_layer, batch_norm_trainable)


class FasterRCNNResnet101FeatureExtractor(FasterRCNNResnetV1FeatureExtractor):
    def __init__(self, is_training, first_stage_features_stride, reuse_weights=None, weight_decay=0.0,
                 freeze_layer="", batch_norm_trainable=False):
        super(FasterRCNNResnet101FeatureExtractor, self).__init__(
            "resnet_v1_101", None, is_training, first_stage_features_stride, reuse_weights, weight_decay,
            freeze_layer, batch_norm_trainable)


class FasterRCNNResnet152FeatureExtractor(FasterRCNNResnetV1FeatureExtractor):
    def __init__(self, is_training, first_stage_features_stride, reuse_weights=None, weight_decay=0.0,
                 freeze_layer="", batch_norm_trainable=False):
        super(FasterRCNNResnet152FeatureExtractor, self).__init__(
            "resnet_v1_152", None, is_training, first_stage_features_stride, reuse_weights, weight_decay,
            freeze_layer, batch_norm_trainable)
