"""Generates tests/golden/stage2_loss_reference.npz by RUNNING the loss methods of the reference's meta-architecture
(/root/reference/object_detection/meta_architectures/faster_rcnn_meta_arch.py: `_loss_box_classifier` :1670-1793 incl.
the closeness loss, `_loss_refined_classifier` :1795-1837, `_loss_window_class` :1839-1858, `_loss_edgemask` :1860-1881)
as unbound functions on a stand-in `self` that carries exactly the attributes `__init__` (:388-429) would have set,
with the NumPy TensorFlow stand-in of tf_numpy_shim.py underneath.  Pins the graph code around the loss objects: which
tensors are fed, the normalisers (T15), the padding indicator, the class-selected box code (T12), the closeness column drop
and its two normalisers (T11), the edge-mask label construction.  TF kernels are formula restatements in the shim:
softmax CE (-sum labels * log_softmax) and `tf.image.resize_images` bilinear with align_corners=False (source index =
destination index * in/out, the TF-1 legacy mapping).
Run from the repo root:  python tests/golden/make_stage2_loss_golden.py"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tf_numpy_shim as shim      # noqa: E402

tf = shim.install()
t = shim.t


def _log_softmax(x):
    m = x.max(-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))


def _resize_bilinear_legacy(img, size, **k):
    """tf.image.resize_images(..., method=BILINEAR, align_corners=False) of TF 1.x: in = out * (in_size / out_size)."""
    img = np.asarray(img, np.float32)
    oh, ow = [int(v) for v in np.asarray(size).reshape(-1)]
    B, H, W, C = img.shape
    ys, xs = np.arange(oh) * (H / oh), np.arange(ow) * (W / ow)
    y0, x0 = np.floor(ys).astype(int), np.floor(xs).astype(int)
    y1, x1 = np.minimum(y0 + 1, H - 1), np.minimum(x0 + 1, W - 1)
    fy, fx = (ys - y0).astype(np.float32)[None, :, None, None], (xs - x0).astype(np.float32)[None, None, :, None]
    top = img[:, y0][:, :, x0] * (1 - fx) + img[:, y0][:, :, x1] * fx
    bot = img[:, y1][:, :, x0] * (1 - fx) + img[:, y1][:, :, x1] * fx
    return (top * (1 - fy) + bot * fy).astype(np.float32)


ce = lambda labels, logits, **k: t(-(np.asarray(labels) * _log_softmax(np.asarray(logits))).sum(-1))
tf.nn.softmax_cross_entropy_with_logits = ce
tf.nn.softmax_cross_entropy_with_logits_v2 = ce
tf.image.resize_images = lambda images, size, **k: t(_resize_bilinear_legacy(images, size))
gu = types.ModuleType("global_utils"); cu = types.ModuleType("global_utils.custom_utils")
cu.log = types.SimpleNamespace(**{k: (lambda *a, **kw: None) for k in ("info", "infov", "warn", "warning", "error")})
sys.modules["global_utils"], sys.modules["global_utils.custom_utils"] = gu, cu
for name in ("object_detection.matchers.bipartite_matcher", "object_detection.box_coders.mean_stddev_box_coder"):
    sys.modules[name] = types.ModuleType(name)
for name in ("object_detection.core.box_predictor", "object_detection.core.mask_predictor"):
    m = types.ModuleType(name)
    m.BOX_ENCODINGS, m.CLASS_PREDICTIONS_WITH_BACKGROUND = "box_encodings", "class_predictions_with_background"
    m.MASK_PREDICTIONS, m.CLASS_PREDICTIONS = "mask_predictions", "class_predictions"
    sys.modules[name] = m
for mname in ("object_detection.core.standard_fields", "object_detection.utils.shape_utils",
              "object_detection.utils.static_shape", "object_detection.core.box_list", "object_detection.core.box_list_ops",
              "object_detection.core.box_coder", "object_detection.box_coders.faster_rcnn_box_coder",
              "object_detection.core.matcher", "object_detection.matchers.argmax_matcher",
              "object_detection.core.region_similarity_calculator", "object_detection.core.target_assigner",
              "object_detection.utils.ops", "object_detection.core.losses", "object_detection.core.minibatch_sampler",
              "object_detection.core.balanced_positive_negative_sampler", "object_detection.core.anchor_generator",
              "object_detection.anchor_generators.grid_anchor_generator", "object_detection.core.model",
              "object_detection.core.post_processing", "object_detection.meta_architectures.faster_rcnn_meta_arch"):
    shim.load_reference_module(mname)
M = sys.modules
fm = M["object_detection.meta_architectures.faster_rcnn_meta_arch"]
losses, ta = M["object_detection.core.losses"], M["object_detection.core.target_assigner"]
box_list, fields = M["object_detection.core.box_list"], M["object_detection.core.standard_fields"]
Arch = fm.FasterRCNNMetaArch


def make_self(K, P):
    s = types.SimpleNamespace()
    s.max_num_proposals = P
    s._mtl = types.SimpleNamespace(window_class_loss_weight=1.0, closeness_loss_weight=0.3, edgemask_loss_weight=1.0,
                                   refined_classification_loss_weight=1.0)
    s._detector_target_assigner = ta.create_target_assigner(
        "FasterRCNN", "detection", unmatched_cls_target=tf.constant([1] + K * [0], dtype=tf.float32, shape=[1, K + 1]))
    s._second_stage_localization_loss = losses.WeightedSmoothL1LocalizationLoss(anchorwise_output=True)
    s._second_stage_classification_loss = losses.WeightedSoftmaxClassificationLoss(anchorwise_output=True)
    s._second_stage_loc_loss_weight, s._second_stage_cls_loss_weight = 2.0, 1.0
    s._hard_example_miner = None
    s._window_class_loss = losses.WeightedSoftmaxClassificationLoss_v2(anchorwise_output=True)
    s._closeness_loss = losses.WeightedSoftmaxClassificationLoss_v2(anchorwise_output=True)
    s._edgemask_loss = losses.WeightedSoftmaxClassificationLoss_v2(anchorwise_output=True)
    s._padded_batched_proposals_indicator = lambda n, m: Arch._padded_batched_proposals_indicator(s, n, m)
    return s


def boxes(rng, n, hi=200):
    y0, x0 = rng.uniform(0, hi * 0.7, n), rng.uniform(0, hi * 0.7, n)
    return np.stack([y0, x0, y0 + rng.uniform(8, hi * 0.5, n), x0 + rng.uniform(8, hi * 0.5, n)], 1).astype(np.float32)


def main():
    rng = np.random.default_rng(23)
    out = {}
    K, P, B = 5, 24, 2
    K1 = K + 1
    s = make_self(K, P)
    for case in range(3):
        nprop = np.array([P, int(rng.integers(1, P))] if case else [P, P], np.int32)
        gts, clss, closes, props = [], [], [], np.zeros((B, P, 4), np.float32)
        gt_lists, cls_list = [], []
        for b in range(B):
            G = int(rng.integers(1, 5))
            gt = boxes(rng, G)
            pr = np.concatenate([gt[rng.integers(0, G, P // 2)] + rng.normal(0, 5, (P // 2, 4)).astype(np.float32),
                                 boxes(rng, P - P // 2)]).astype(np.float32)
            pr[nprop[b]:] = 0                                                  # zero padding past num_proposals
            props[b] = pr
            cls = np.eye(K1, dtype=np.float32)[rng.integers(1, K1, G)]
            close = rng.random((G, K1)).astype(np.float32)
            close[:, 0] = 0
            close *= rng.random((G, K1)) < 0.6
            bl = box_list.BoxList(t(gt))
            bl.add_field(fields.BoxListFields.closeness, t(close))
            bl.add_field("ignore", t(np.zeros((G,), bool)))
            gt_lists.append(bl); cls_list.append(t(cls))
            gts.append(gt); clss.append(cls); closes.append(close)
        enc = rng.normal(0, 1, (B * P, K, 4)).astype(np.float32)
        logits = rng.normal(0, 2, (B * P, K1)).astype(np.float32)
        refined = rng.normal(0, 2, (B * P, K1)).astype(np.float32)
        close_pred = rng.normal(0, 2, (B * P, K1)).astype(np.float32)
        d1 = Arch._loss_box_classifier(s, t(enc), t(logits), t(props), t(nprop), gt_lists, cls_list,
                                       closeness_predictions=t(close_pred))
        d2 = Arch._loss_refined_classifier(s, t(enc), t(refined), t(props), t(nprop), gt_lists, cls_list)
        NW, EH = 16, 64
        win_pred = rng.normal(0, 2, (B * NW, K1)).astype(np.float32)
        win_lab = rng.random((B, NW, K1)).astype(np.float32)
        win_lab /= win_lab.sum(-1, keepdims=True)
        d3 = Arch._loss_window_class(s, t(win_pred), [t(win_lab[b]) for b in range(B)])
        em_pred = np.tanh(rng.normal(0, 1, (B, 14, 20, 2))).astype(np.float32)
        fg = (rng.random((B, EH, EH)) < 0.3).astype(np.float32)
        wt = rng.random((B, EH, EH)).astype(np.float32) * 2
        d4 = Arch._loss_edgemask(s, t(em_pred), [t(np.stack([fg[b], wt[b]])) for b in range(B)])
        p = "case%d/" % case
        out[p + "nprop"], out[p + "props"], out[p + "enc"], out[p + "logits"] = nprop, props, enc, logits
        out[p + "refined"], out[p + "close_pred"] = refined, close_pred
        for b in range(B):
            out[p + "gt%d" % b], out[p + "cls%d" % b], out[p + "close%d" % b] = gts[b], clss[b], closes[b]
        out[p + "win_pred"], out[p + "win_lab"], out[p + "em_pred"] = win_pred, win_lab, em_pred
        out[p + "em_gt"] = np.stack([fg, wt], 1)
        for d in (d1, d2, d3, d4):
            for k, v in d.items():
                out[p + "loss/" + k] = np.asarray(v, np.float64)
    out["meta"] = np.array([K, P, B, 3])
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "stage2_loss_reference.npz"), **out)
    print("wrote stage2_loss_reference.npz", {k: float(v) for k, v in out.items() if k.startswith("case1/loss/")})


if __name__ == "__main__":
    main()
