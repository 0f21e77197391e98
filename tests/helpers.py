"""Shared test helpers: config fixtures, oracle configuration derived from a parsed config."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CONFIG_DIR = os.path.join(HERE, "golden", "configs")


def load_config(name="model12.config", replace=()):
    from mtl_ssl_b200.protos import text_format
    text = open(os.path.join(CONFIG_DIR, name)).read()
    for a, b in replace:
        assert a in text, a
        text = text.replace(a, b)
    return text_format.Merge(text, text_format.Message("TrainEvalPipelineConfig"))


def oracle_config(cfg):
    """The hyper-parameters oracle/model.py needs, read from the parsed pipeline config."""
    fr, mtl = cfg.model.faster_rcnn, cfg.model.mtl
    g = fr.first_stage_anchor_generator.grid_anchor_generator
    arch = {"faster_rcnn_resnet50": "resnet_v1_50", "faster_rcnn_resnet101": "resnet_v1_101",
            "faster_rcnn_resnet152": "resnet_v1_152", "frcnn_mobilenet_v1": "MobilenetV1",
            "faster_rcnn_inception_v2": "InceptionResnetV2"}[fr.feature_extractor.type]
    rfcn = None
    if fr.second_stage_box_predictor.WhichOneof("box_predictor_oneof") == "rfcn_box_predictor":
        r = fr.second_stage_box_predictor.rfcn_box_predictor
        rfcn = dict(bins=(r.num_spatial_bins_height, r.num_spatial_bins_width), crop=(r.crop_height, r.crop_width),
                    depth=r.depth)
    return dict(
        rfcn=rfcn, architecture=arch, num_classes=fr.num_classes, scales=list(g.scales), aspect_ratios=list(g.aspect_ratios),
        first_stage_max_proposals=fr.first_stage_max_proposals, second_stage_batch_size=fr.second_stage_batch_size,
        second_stage_balance_fraction=fr.second_stage_balance_fraction,
        first_stage_minibatch_size=fr.first_stage_minibatch_size,
        first_stage_positive_balance_fraction=fr.first_stage_positive_balance_fraction,
        nms_score_threshold=fr.first_stage_nms_score_threshold, nms_iou_threshold=fr.first_stage_nms_iou_threshold,
        initial_crop_size=fr.initial_crop_size, maxpool_kernel_size=fr.maxpool_kernel_size,
        first_stage_localization_loss_weight=fr.first_stage_localization_loss_weight,
        first_stage_objectness_loss_weight=fr.first_stage_objectness_loss_weight,
        second_stage_localization_loss_weight=fr.second_stage_localization_loss_weight,
        second_stage_classification_loss_weight=fr.second_stage_classification_loss_weight,
        mtl=dict(window=mtl.window, closeness=mtl.closeness, edgemask=mtl.edgemask, refine=mtl.refine,
                 refine_residue=mtl.refine_residue, stop_gradient_for_aux_tasks=mtl.stop_gradient_for_aux_tasks,
                 window_class_loss_weight=mtl.window_class_loss_weight,
                 closeness_loss_weight=mtl.closeness_loss_weight, edgemask_loss_weight=mtl.edgemask_loss_weight,
                 refined_classification_loss_weight=mtl.refined_classification_loss_weight))


def randomize_bn(sd, seed=0):
    """Non-trivial frozen batch-norm statistics so that the fold (scale, bias) is exercised, chosen
    so that activations stay O(1) through the residual stack (stem variance ~ pixel variance x
    fan-in x He variance; small conv3 gammas keep the residual sum from growing)."""
    rng = np.random.default_rng(seed)
    import torch
    u = lambda lo, hi, n: torch.from_numpy(rng.uniform(lo, hi, n).astype(np.float32))
    for k in list(sd):
        n = sd[k].numel()
        stem = ("/block" not in k and "/conv1/BatchNorm/" in k)
        if "/MobilenetV1/" in k:                # keep ReLU6 activations in range through 13 layers
            if k.endswith("/BatchNorm/gamma"):
                sd[k] = u(0.8, 1.1, n)
            elif k.endswith("/BatchNorm/beta"):
                sd[k] = u(0.0, 0.3, n)
            elif k.endswith("/BatchNorm/moving_mean"):
                sd[k] = u(-0.1, 0.1, n)
            elif k.endswith("/BatchNorm/moving_variance"):
                # normalise each layer to ~unit output variance: var = fan_in * stddev_init^2 * E[x^2]
                base = k[:-len("/BatchNorm/moving_variance")]
                wkey = [c for c in (base + "/weights", base + "/depthwise_weights", base + "/pointwise_weights")
                        if c in sd and "depthwise" not in c or (c in sd and base.endswith("_depthwise"))][0]
                w = sd[wkey]
                fan_in = 27 if base.endswith("Conv2d_0") else w[0].numel()
                sd[k] = u(0.8, 1.2, n) * fan_in * 0.0081 * 0.6
            continue
        if k.endswith("/BatchNorm/gamma"):
            if "/conv3/" in k:
                sd[k] = u(0.15, 0.3, n)
            elif "/shortcut/" in k:
                sd[k] = u(0.5, 0.8, n)
            else:
                sd[k] = u(0.7, 1.2, n)
        elif k.endswith("/BatchNorm/beta"):
            sd[k] = u(-0.2, 0.2, n)
        elif k.endswith("/BatchNorm/moving_mean"):
            sd[k] = u(-0.2, 0.2, n)
        elif k.endswith("/BatchNorm/moving_variance"):
            sd[k] = u(0.7, 1.3, n) * (14000.0 if stem else 1.0)
    return sd
