"""Builds a DetectionModel from a parsed pipeline config, argument-for-argument like
/root/reference/object_detection/builders/model_builder.py:68-380 (`build`,
`_build_faster_rcnn_feature_extractor`, `_build_faster_rcnn_model`) and the sub-builders it calls
(anchor_generator_builder, box_predictor_builder, mask_predictor_builder, hyperparams_builder,
image_resizer_builder)."""
import torch

from ..anchor_generators.grid_anchor_generator import GridAnchorGenerator
from ..core import box_predictor
from ..core.hyperparams import Hyperparams
from ..core.mask_predictor import MaskPredictor
from ..core import preprocessor
from ..meta_architectures import faster_rcnn_meta_arch, rfcn_meta_arch
from ..models import faster_rcnn_inception_resnet_v2_feature_extractor as frcnn_inc_res
from ..models import faster_rcnn_mobilenet_v1_feature_extractor as frcnn_mobilenet_v1
from ..models import faster_rcnn_resnet_v1_feature_extractor as frcnn_resnet_v1

# model_builder.py:50-66
FASTER_RCNN_FEATURE_EXTRACTOR_CLASS_MAP = {
    "faster_rcnn_resnet50": frcnn_resnet_v1.FasterRCNNResnet50FeatureExtractor,
    "faster_rcnn_resnet101": frcnn_resnet_v1.FasterRCNNResnet101FeatureExtractor,
    "faster_rcnn_resnet152": frcnn_resnet_v1.FasterRCNNResnet152FeatureExtractor,
    "frcnn_mobilenet_v1": frcnn_mobilenet_v1.FasterRCNNMobilenetV1FeatureExtractor,
    # model_builder.py:64-65: the fork registers Inception-ResNet-v2 under both names
    "faster_rcnn_inception_v2": frcnn_inc_res.FasterRCNNInceptionResnetV2FeatureExtractor,
    "faster_rcnn_inception_resnet_v2": frcnn_inc_res.FasterRCNNInceptionResnetV2FeatureExtractor,
}


def build(model_config, is_training, device="cuda", seed=0):
    """model_builder.py:68-95."""
    meta_architecture = model_config.WhichOneof("model")
    if meta_architecture == "faster_rcnn":
        return _build_faster_rcnn_model(model_config.faster_rcnn, is_training, model_config.mtl, device, seed)
    if meta_architecture == "ssd":
        raise ValueError("SSD meta-architecture is outside the B200 hot path (DESIGN.md, out of scope)")
    raise ValueError("Unknown meta architecture: {}".format(meta_architecture))


def _build_faster_rcnn_feature_extractor(feature_extractor_config, is_training, reuse_weights=None, **kwargs):
    feature_type = feature_extractor_config.type
    first_stage_features_stride = feature_extractor_config.first_stage_features_stride
    if feature_type not in FASTER_RCNN_FEATURE_EXTRACTOR_CLASS_MAP:
        raise ValueError("Unknown Faster R-CNN feature_extractor: {}".format(feature_type))
    cls = FASTER_RCNN_FEATURE_EXTRACTOR_CLASS_MAP[feature_type]
    return cls(is_training, first_stage_features_stride, reuse_weights, **kwargs)


def build_image_resizer(cfg):
    """builders/image_resizer_builder.py: keep_aspect_ratio_resizer | fixed_shape_resizer."""
    which = cfg.WhichOneof("image_resizer_oneof")
    if which == "keep_aspect_ratio_resizer":
        r = cfg.keep_aspect_ratio_resizer
        if not r.min_dimension <= r.max_dimension:
            raise ValueError("min_dimension > max_dimension")
        fn = lambda img: preprocessor.resize_to_range(img, r.min_dimension, r.max_dimension)
        fn.static_size = lambda h, w: tuple(preprocessor._compute_new_static_size(h, w, r.min_dimension, r.max_dimension))
        return fn
    if which == "fixed_shape_resizer":
        r = cfg.fixed_shape_resizer
        fn = lambda img: preprocessor.resize_image(img, r.height, r.width)
        fn.static_size = lambda h, w: (int(r.height), int(r.width))
        return fn
    raise ValueError("Invalid image resizer option.")


def build_anchor_generator(cfg):
    """builders/anchor_generator_builder.py:24-55 (grid_anchor_generator branch)."""
    which = cfg.WhichOneof("anchor_generator_oneof")
    if which != "grid_anchor_generator":
        raise ValueError("B200 path supports grid_anchor_generator only (got %s)" % which)
    g = cfg.grid_anchor_generator
    # the fork's paired height/width scales and left-top alignment (grid_anchor_generator.py:39-42, proto fields 9-12)
    # are used by no shipped config and have no device kernel: refused rather than ignored
    if g.use_hw_scales or g.align_lefttop:
        raise ValueError("grid_anchor_generator.use_hw_scales / align_lefttop are not supported on the B200 path")
    return GridAnchorGenerator(scales=[float(s) for s in g.scales],
                               aspect_ratios=[float(a) for a in g.aspect_ratios],
                               base_anchor_size=[g.height, g.width],
                               anchor_stride=[g.height_stride, g.width_stride],
                               anchor_offset=[g.height_offset, g.width_offset])


def build_box_predictor(cfg, is_training, num_classes):
    """builders/box_predictor_builder.py:24-125."""
    which = cfg.WhichOneof("box_predictor_oneof")
    if which == "mask_rcnn_box_predictor":
        m = cfg.mask_rcnn_box_predictor
        return box_predictor.MaskRCNNBoxPredictor(
            is_training=is_training, num_classes=num_classes, fc_hyperparams=Hyperparams.from_proto(m.fc_hyperparams),
            use_dropout=m.use_dropout, dropout_keep_prob=m.dropout_keep_probability, box_code_size=m.box_code_size,
            predict_instance_masks=m.predict_instance_masks, spatial_average=m.spatial_average)
    if which == "rfcn_box_predictor":
        r = cfg.rfcn_box_predictor
        return box_predictor.RfcnBoxPredictor(
            is_training=is_training, num_classes=num_classes,
            conv_hyperparams=Hyperparams.from_proto(r.conv_hyperparams),
            crop_size=[r.crop_height, r.crop_width],
            num_spatial_bins=[r.num_spatial_bins_height, r.num_spatial_bins_width], depth=r.depth,
            box_code_size=r.box_code_size)
    raise ValueError("Unknown box predictor: {}".format(which))


def _build_faster_rcnn_model(frcnn_config, is_training, mtl=None, device="cuda", seed=0):
    """model_builder.py:213-380."""
    num_classes = frcnn_config.num_classes
    image_resizer_fn = build_image_resizer(frcnn_config.image_resizer)
    fe_kwargs = {"freeze_layer": frcnn_config.feature_extractor.freeze_layer,
                 "batch_norm_trainable": frcnn_config.feature_extractor.batch_norm_trainable}
    if frcnn_config.feature_extractor.HasField("weight_decay"):
        fe_kwargs["weight_decay"] = frcnn_config.feature_extractor.weight_decay
    feature_extractor = _build_faster_rcnn_feature_extractor(
        frcnn_config.feature_extractor, is_training and frcnn_config.feature_extractor.trainable, **fe_kwargs)
    second_stage_box_predictor = build_box_predictor(
        frcnn_config.second_stage_box_predictor,
        is_training and frcnn_config.second_stage_box_predictor.trainable, num_classes)
    if mtl.window:
        window_box_predictor = build_box_predictor(mtl.window_box_predictor,
                                                   is_training and mtl.window_box_predictor.trainable, num_classes + 1)
    else:
        window_box_predictor = second_stage_box_predictor
    if mtl.closeness:
        closeness_box_predictor = build_box_predictor(
            mtl.closeness_box_predictor, is_training and mtl.closeness_box_predictor.trainable, num_classes + 1)
    else:
        closeness_box_predictor = second_stage_box_predictor
    edgemask_predictor = None
    if mtl.edgemask:
        e = mtl.edgemask_predictor
        edgemask_predictor = MaskPredictor(is_training and e.trainable, 2, Hyperparams.from_proto(e.conv_hyperparams),
                                           kernel_size=e.kernel_size, channels=1)
    mtl_refiner_arg_scope = Hyperparams.from_proto(mtl.refiner_fc_hyperparams) if mtl.refine else None
    if frcnn_config.HasField("hard_example_miner"):
        raise ValueError("hard_example_miner is not supported on the B200 training path")
    pp = frcnn_config.second_stage_post_processing
    common_kwargs = dict(
        is_training=is_training, num_classes=num_classes, image_resizer_fn=image_resizer_fn,
        feature_extractor=feature_extractor, first_stage_only=frcnn_config.first_stage_only,
        first_stage_anchor_generator=build_anchor_generator(frcnn_config.first_stage_anchor_generator),
        first_stage_clip_window=frcnn_config.first_stage_clip_window,
        first_stage_atrous_rate=frcnn_config.first_stage_atrous_rate,
        first_stage_box_predictor_trainable=frcnn_config.first_stage_box_predictor_trainable,
        first_stage_box_predictor_arg_scope=Hyperparams.from_proto(
            frcnn_config.first_stage_box_predictor_conv_hyperparams),
        first_stage_box_predictor_kernel_size=frcnn_config.first_stage_box_predictor_kernel_size,
        first_stage_box_predictor_depth=frcnn_config.first_stage_box_predictor_depth,
        first_stage_minibatch_size=frcnn_config.first_stage_minibatch_size,
        first_stage_positive_balance_fraction=frcnn_config.first_stage_positive_balance_fraction,
        first_stage_nms_score_threshold=frcnn_config.first_stage_nms_score_threshold,
        first_stage_nms_iou_threshold=frcnn_config.first_stage_nms_iou_threshold,
        first_stage_max_proposals=frcnn_config.first_stage_max_proposals,
        first_stage_localization_loss_weight=frcnn_config.first_stage_localization_loss_weight,
        first_stage_objectness_loss_weight=frcnn_config.first_stage_objectness_loss_weight,
        second_stage_batch_size=frcnn_config.second_stage_batch_size,
        second_stage_balance_fraction=frcnn_config.second_stage_balance_fraction,
        second_stage_non_max_suppression_fn=pp.batch_non_max_suppression,
        second_stage_score_conversion_fn=pp.score_converter,
        second_stage_localization_loss_weight=frcnn_config.second_stage_localization_loss_weight,
        second_stage_classification_loss_weight=frcnn_config.second_stage_classification_loss_weight,
        hard_example_miner=None, mtl=mtl, mtl_refiner_arg_scope=mtl_refiner_arg_scope,
        window_box_predictor=window_box_predictor, closeness_box_predictor=closeness_box_predictor,
        edgemask_predictor=edgemask_predictor, device=device, seed=seed)
    if isinstance(second_stage_box_predictor, box_predictor.RfcnBoxPredictor):
        return rfcn_meta_arch.RFCNMetaArch(second_stage_rfcn_box_predictor=second_stage_box_predictor,
                                           **common_kwargs)
    return faster_rcnn_meta_arch.FasterRCNNMetaArch(
        initial_crop_size=frcnn_config.initial_crop_size, maxpool_kernel_size=frcnn_config.maxpool_kernel_size,
        maxpool_stride=frcnn_config.maxpool_stride,
        second_stage_mask_rcnn_box_predictor=second_stage_box_predictor, **common_kwargs)
