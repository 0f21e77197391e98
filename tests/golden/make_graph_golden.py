"""Generates tests/golden/graph_reference.npz by RUNNING graph-building methods of the reference's meta-architecture
(/root/reference/object_detection/meta_architectures/faster_rcnn_meta_arch.py) as unbound functions on a stand-in `self`,
with the NumPy TensorFlow stand-in of tf_numpy_shim.py underneath:

  rpn/      `_remove_invalid_anchors_and_predictions` :930-976, `_postprocess_rpn` :1055-1132 in training mode (decode,
            objectness softmax, `batch_multiclass_non_max_suppression`, `_format_groundtruth_data` :1218-1266,
            `_unpad_proposals_and_sample_box_classifier_batch` :1134-1216, `_sample_box_classifier_minibatch`
            :1268-1302, normalisation) and `_loss_rpn` :1591-1668 (T1, T15)
  refine/   `predict_with_mtl_results` :764-846: the five expanded windows per proposal, their flattening order, the
            [5, P, K+1] -> [P, 5(K+1)] transposition, the global closeness mean, the concatenation order and the residue
  crop/     `_compute_second_stage_input_feature_maps` :1304-1348: the box -> image index map and the max-pool arguments

TF kernels are formula restatements in the shim: `tf.image.non_max_suppression` is the reference's own NumPy NMS
(utils/np_box_list_ops.py:185-257, same suppress-iff-IoU>thr rule), `tf.nn.top_k` a stable descending sort, softmax /
softmax CE the textbook formulas.  `tf.random_shuffle` is replaced by a KEYED permutation (candidates ordered by ascending
key) so that the sampler's count logic can be compared with oracle/assign.py, which takes the same keys as an input.
Run from the repo root:  python tests/golden/make_graph_golden.py"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tf_numpy_shim as shim      # noqa: E402

tf = shim.install()
t = shim.t
F = np.float32


def _softmax(x, **k):
    x = np.asarray(x, F)
    e = np.exp(x - x.max(-1, keepdims=True))
    return t((e / e.sum(-1, keepdims=True)).astype(F))


def _log_softmax(x):
    m = x.max(-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))


ce = lambda labels, logits, **k: t(-(np.asarray(labels) * _log_softmax(np.asarray(logits))).sum(-1))
tf.nn.softmax = _softmax
tf.nn.softmax_cross_entropy_with_logits = ce
tf.nn.softmax_cross_entropy_with_logits_v2 = ce


def _top_k(x, k=1, sorted=True):
    x = np.asarray(x)
    order = np.argsort(-x, kind="stable")[:int(k)]
    return t(x[order]), t(order.astype(np.int32))


tf.nn.top_k = _top_k
KEYS = []                                   # queue of per-call key arrays for the keyed shuffle


def _keyed_shuffle(x, seed=None):
    x = np.asarray(x)
    keys = KEYS.pop(0)
    flat = x.reshape(-1)
    return t(x[np.argsort(np.asarray(keys, F)[flat], kind="stable")])


tf.random_shuffle = _keyed_shuffle
RECORD = {}


def _map_fn(fn, elems, dtype=None, parallel_iterations=None, back_prop=True, **k):
    if isinstance(elems, (list, tuple)):
        outs = [fn([t(np.asarray(e)[i]) for e in elems]) for i in range(len(elems[0]))]
    else:
        outs = [fn(t(np.asarray(elems)[i])) for i in range(len(elems))]
    if isinstance(outs[0], (list, tuple)):
        return [t(np.stack([np.asarray(o[j]) for o in outs])) for j in range(len(outs[0]))]
    return t(np.stack([np.asarray(o) for o in outs]))


tf.map_fn = _map_fn
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
              "object_detection.core.post_processing", "object_detection.utils.np_box_list",
              "object_detection.utils.np_box_ops", "object_detection.utils.np_box_list_ops",
              "object_detection.meta_architectures.faster_rcnn_meta_arch"):
    shim.load_reference_module(mname)
M = sys.modules
fm = M["object_detection.meta_architectures.faster_rcnn_meta_arch"]
losses, ta = M["object_detection.core.losses"], M["object_detection.core.target_assigner"]
box_list, fields = M["object_detection.core.box_list"], M["object_detection.core.standard_fields"]
sampler = M["object_detection.core.balanced_positive_negative_sampler"]
np_box_list, np_box_list_ops = M["object_detection.utils.np_box_list"], M["object_detection.utils.np_box_list_ops"]
coder = M["object_detection.box_coders.faster_rcnn_box_coder"]
Arch = fm.FasterRCNNMetaArch


def _nms(boxes, scores, max_output_size, iou_threshold=0.5, **k):
    """tf.image.non_max_suppression through the reference's NumPy NMS (np_box_list_ops.py:185-257)."""
    boxes, scores = np.asarray(boxes, F), np.asarray(scores, F)
    if len(boxes) == 0:
        return t(np.zeros((0,), np.int32))
    bl = np_box_list.BoxList(boxes)
    bl.add_field("scores", scores)
    bl.add_field("index", np.arange(len(boxes), dtype=np.int32))
    out = np_box_list_ops.non_max_suppression(bl, int(max_output_size), float(iou_threshold), score_threshold=-np.inf)
    return t(out.get_field("index").astype(np.int32))


tf.image.non_max_suppression = _nms


def make_self(K, P, M_, minibatch):
    s = types.SimpleNamespace()
    s._is_training, s._hard_example_miner, s._parallel_iterations = True, None, 1
    s._num_classes = K
    s.max_num_proposals = P
    s._second_stage_batch_size, s._first_stage_max_proposals = P, M_
    s._first_stage_nms_score_threshold, s._first_stage_nms_iou_threshold = 0.0, 0.7
    s._first_stage_minibatch_size = minibatch
    s._box_coder = coder.FasterRcnnBoxCoder(scale_factors=[10.0, 10.0, 5.0, 5.0])
    s._proposal_target_assigner = ta.create_target_assigner("FasterRCNN", "proposal")
    s._detector_target_assigner = ta.create_target_assigner(
        "FasterRCNN", "detection", unmatched_cls_target=tf.constant([1] + K * [0], dtype=tf.float32, shape=[1, K + 1]))
    s._first_stage_sampler = sampler.BalancedPositiveNegativeSampler(positive_fraction=0.5)
    s._second_stage_sampler = sampler.BalancedPositiveNegativeSampler(positive_fraction=0.25)
    s._first_stage_localization_loss = losses.WeightedSmoothL1LocalizationLoss(anchorwise_output=True, sigma=3.0)
    s._first_stage_objectness_loss = losses.WeightedSoftmaxClassificationLoss(anchorwise_output=True)
    s._first_stage_loc_loss_weight, s._first_stage_obj_loss_weight = 2.0, 1.0
    for name in ("_batch_decode_boxes", "_format_groundtruth_data", "_unpad_proposals_and_sample_box_classifier_batch",
                 "_sample_box_classifier_minibatch", "_flatten_first_two_dimensions"):
        setattr(s, name, (lambda f: lambda *a, **k: f(s, *a, **k))(getattr(Arch, name)))
    return s


def rand_boxes(rng, n, H, W):
    y0, x0 = rng.uniform(0, H * 0.7, n), rng.uniform(0, W * 0.7, n)
    return np.stack([y0, x0, np.minimum(y0 + rng.uniform(8, H * 0.6, n), H), np.minimum(x0 + rng.uniform(8, W * 0.6, n), W)],
                    1).astype(F)


def rpn_cases(out):
    rng = np.random.default_rng(41)
    K, P, M_, MB = 4, 16, 40, 32
    H, W, Hf, Wf = 160, 224, 10, 14
    scales, ars = (0.25, 0.5, 1.0), (0.5, 1.0, 2.0)
    # input anchors only (the generator itself is pinned by the reference's KATs in test_oracle_kats.py): a 16-px grid of
    # 64-px base boxes, some of which cross the image border
    ys, xs = np.meshgrid(np.arange(Hf) * 16.0, np.arange(Wf) * 16.0, indexing="ij")
    hw = np.array([[64 * sc / np.sqrt(ar), 64 * sc * np.sqrt(ar)] for ar in ars for sc in scales])
    cy, cx = ys.reshape(-1, 1), xs.reshape(-1, 1)
    anchors_all = box_list.BoxList(t(np.stack([cy - hw[:, 0] / 2, cx - hw[:, 1] / 2, cy + hw[:, 0] / 2, cx + hw[:, 1] / 2],
                                              -1).reshape(-1, 4).astype(F)))
    s = make_self(K, P, M_, MB)
    image_shape = t(np.array([2, H, W, 3], np.int32))
    clip_window = t(np.array([0, 0, H, W], F))
    B = 2
    for case in range(3):
        N = len(np.asarray(anchors_all.get()))
        enc = rng.normal(0, 1.0, (B, N, 4)).astype(F)
        logits = rng.normal(0, 2.0, (B, N, 2)).astype(F)
        if case == 1:
            # image 1: every anchor regresses onto one of five far-apart boxes, so that NMS leaves five proposals
            # (a proposal list shorter than second_stage_batch_size: zero padding + num_proposals < P)
            tg = np.array([[8, 8, 60, 70], [90, 20, 150, 80], [10, 120, 70, 200], [95, 130, 155, 215], [40, 85, 110, 125]], F)
            a = np.asarray(anchors_all.get(), F)
            ha, wa, ya, xa = a[:, 2] - a[:, 0], a[:, 3] - a[:, 1], (a[:, 0] + a[:, 2]) / 2, (a[:, 1] + a[:, 3]) / 2
            g = tg[np.arange(N) % 5]
            hg, wg, yg, xg = g[:, 2] - g[:, 0], g[:, 3] - g[:, 1], (g[:, 0] + g[:, 2]) / 2, (g[:, 1] + g[:, 3]) / 2
            enc[1] = np.stack([10 * (yg - ya) / ha, 10 * (xg - xa) / wa, 5 * np.log(hg / ha), 5 * np.log(wg / wa)], 1)
        enc_k, logit_k, anchors_k = Arch._remove_invalid_anchors_and_predictions(s, t(enc), t(logits), anchors_all,
                                                                                  clip_window)
        anchors_np = np.asarray(anchors_k.get(), F)
        Nk = len(anchors_np)
        gts, clss = [], []
        for b in range(B):
            G = int(rng.integers(1, 4)) if case < 2 else 1
            gt = rand_boxes(rng, G, H, W)
            if case == 2 and b == 1:
                gt = np.array([[0.0, 0.0, 9.0, 9.0]], F)            # no proposal reaches IoU 0.5: negatives only
            gts.append(gt)
            clss.append(np.eye(K, dtype=F)[rng.integers(0, K, len(gt))])
        s.groundtruth_lists = lambda field: {
            fields.BoxListFields.boxes: [t(g / np.array([H, W, H, W], F)) for g in gts],
            fields.BoxListFields.classes: [t(c) for c in clss],
            fields.BoxListFields.closeness: [None] * B, fields.BoxListFields.ignore: [None] * B}[field]
        keys2 = rng.random((B, M_)).astype(F)
        for b in range(B):
            KEYS.extend([keys2[b], keys2[b]])
        pb, ps, npz = Arch._postprocess_rpn(s, t(np.asarray(enc_k)), t(np.asarray(logit_k)), t(anchors_np), image_shape)
        assert not KEYS
        keys1 = rng.random((B, Nk)).astype(F)
        for b in range(B):
            KEYS.extend([keys1[b], keys1[b]])
        gt_lists, _ = Arch._format_groundtruth_data(s, image_shape)
        ld = Arch._loss_rpn(s, t(np.asarray(enc_k)), t(np.asarray(logit_k)), t(anchors_np), gt_lists,
                            [None] * B)
        assert not KEYS
        p = "rpn%d/" % case
        out[p + "enc"], out[p + "logits"], out[p + "anchors_all"] = enc, logits, np.asarray(anchors_all.get(), F)
        out[p + "anchors"], out[p + "enc_kept"], out[p + "logits_kept"] = anchors_np, np.asarray(enc_k), np.asarray(logit_k)
        for b in range(B):
            out[p + "gt%d" % b], out[p + "cls%d" % b] = gts[b], clss[b]
            out[p + "gt_abs%d" % b] = np.asarray(gt_lists[b].get(), F)
        out[p + "keys1"], out[p + "keys2"] = keys1, keys2
        out[p + "prop_norm"], out[p + "prop_scores"], out[p + "nprop"] = np.asarray(pb, F), np.asarray(ps, F), np.asarray(npz)
        for k, v in ld.items():
            out[p + "loss/" + k] = np.asarray(v, np.float64)
    out["rpn_meta"] = np.array([K, P, M_, MB, H, W, Hf, Wf, 3])
    out["rpn_scales"], out["rpn_ars"] = np.array(scales, F), np.array(ars, F)


def refine_case(out):
    """predict_with_mtl_results with a recording stand-in for `predict_with_window` (logits = a fixed linear function of the
    window box, so every window stays identifiable) and `slim.fully_connected` = x @ W + b."""
    rng = np.random.default_rng(43)
    K1, P = 5, 12
    wmat = rng.normal(0, 1, (4, K1)).astype(F)
    fcw = rng.normal(0, 0.3, (K1 + 5 * K1 + K1, K1)).astype(F)
    fcb = rng.normal(0, 0.1, (K1,)).astype(F)
    s = types.SimpleNamespace()
    s._mtl = types.SimpleNamespace(stop_gradient_for_prediction_org=True, window=True, closeness=True,
                                   global_closeness=True, refine_num_fc_layers=0, refine_dropout_rate=1.0,
                                   refine_residue=True)
    s._is_training, s.num_classes, s.mtl_refiner_scope, s._mtl_refiner_arg_scope = True, K1 - 1, "MTLClassRefiner", None

    def predict_with_window(d, window_boxes_normalized=None):
        wb = np.asarray(window_boxes_normalized, F)
        RECORD["refine_windows"] = wb.copy()
        d["window_class_predictions"] = t(wb.reshape(-1, 4) @ wmat)
        return d

    s.predict_with_window = predict_with_window
    fm.slim = types.SimpleNamespace(
        arg_scope=lambda *a, **k: shim._Ctx(),
        fully_connected=lambda net, n, scope=None, activation_fn=None: t(np.asarray(net, F) @ fcw + fcb),
        dropout=lambda net, *a, **k: net)
    props = np.sort(rng.random((1, P, 2, 2)).astype(F), axis=2).transpose(0, 1, 3, 2).reshape(1, P, 4)[..., [0, 2, 1, 3]]
    props = np.stack([props[..., 0], props[..., 2], props[..., 1], props[..., 3]], -1)
    y0, y1 = np.minimum(props[..., 0], props[..., 2]), np.maximum(props[..., 0], props[..., 2])
    x0, x1 = np.minimum(props[..., 1], props[..., 3]), np.maximum(props[..., 1], props[..., 3])
    props = np.stack([y0, x0, y1, x1], -1).astype(F)
    cls = rng.normal(0, 2, (P, K1)).astype(F)
    close = rng.normal(0, 2, (P, K1)).astype(F)
    d = dict(class_predictions_with_background=t(cls), proposal_boxes_normalized=t(props),
             rpn_features_to_crop=t(np.zeros((1, 2, 2, 1), F)), closeness_predictions=t(close))
    d = Arch.predict_with_mtl_results(s, d)
    out["refine/props"], out["refine/cls"], out["refine/close"] = props, cls, close
    out["refine/wmat"], out["refine/fcw"], out["refine/fcb"] = wmat, fcw, fcb
    out["refine/windows"] = RECORD["refine_windows"]
    out["refine/expand_window_class_predictions"] = np.asarray(d["expand_window_class_predictions"], F)
    out["refine/refined"] = np.asarray(d["mtl_refined_class_predictions_with_background"], F)


def crop_case(out):
    """_compute_second_stage_input_feature_maps with recording stand-ins for crop_and_resize / max_pool2d."""
    s = types.SimpleNamespace(_initial_crop_size=14, _maxpool_kernel_size=2, _maxpool_stride=2)
    s._flatten_first_two_dimensions = lambda x: Arch._flatten_first_two_dimensions(s, x)

    def crop_and_resize(image, boxes, box_ind, crop_size, **k):
        RECORD["crop"] = (np.asarray(boxes, F).copy(), np.asarray(box_ind).copy(), tuple(int(v) for v in crop_size))
        return t(np.zeros((len(np.asarray(boxes)),) + tuple(int(v) for v in crop_size) + (1,), F))

    tf.image.crop_and_resize = crop_and_resize
    fm.slim = types.SimpleNamespace(max_pool2d=lambda x, k, stride=None, **kw: (RECORD.__setitem__("pool", (k, stride)), x)[1])
    rng = np.random.default_rng(47)
    boxes = rng.random((3, 5, 4)).astype(F)
    Arch._compute_second_stage_input_feature_maps(s, t(np.zeros((3, 4, 4, 1), F)), t(boxes))
    out["crop/boxes"], out["crop/flat_boxes"], out["crop/box_ind"] = boxes, RECORD["crop"][0], RECORD["crop"][1]
    out["crop/size_pool"] = np.array(list(RECORD["crop"][2]) + [int(np.ravel(RECORD["pool"][0])[0]), int(np.ravel(RECORD["pool"][1])[0])])
    Arch._compute_second_stage_input_feature_maps(s, t(np.zeros((1, 4, 4, 1), F)), t(boxes.reshape(1, 15, 4)))
    out["crop/box_ind_rank3_batch1"] = RECORD["crop"][1]


def postprocess_case(out):
    """`postprocess` (second stage, fmA:1038-1053) -> `_postprocess_box_classifier` :1387-1469 with the configs'
    second-stage NMS (post_processing_builder: score_threshold 0.0 / 0.3, IoU 0.6, per class = total) and the SOFTMAX
    score converter, num_proposals shorter than max_num_proposals for one image."""
    import functools
    pp = M["object_detection.core.post_processing"]
    rng = np.random.default_rng(53)
    K, P, B, H, W = 4, 24, 2, 160, 224
    for case, (thr, max_det) in enumerate(((0.0, 20), (0.3, 100))):
        s = types.SimpleNamespace(num_classes=K, max_num_proposals=P, _first_stage_only=False, _parallel_iterations=1,
                                  _mtl=types.SimpleNamespace(refine=False))
        s._box_coder = coder.FasterRcnnBoxCoder(scale_factors=[10.0, 10.0, 5.0, 5.0])
        s._second_stage_score_conversion_fn = _softmax
        s._second_stage_nms_fn = functools.partial(pp.batch_multiclass_non_max_suppression, score_thresh=thr,
                                                   iou_thresh=0.6, max_size_per_class=max_det, max_total_size=max_det)
        s._batch_decode_boxes = lambda e, a: Arch._batch_decode_boxes(s, e, a)
        s._postprocess_box_classifier = lambda *a, **k: Arch._postprocess_box_classifier(s, *a, **k)
        props = np.stack([rand_boxes(rng, P, H, W) for _ in range(B)])
        nprop = np.array([P, 17], np.int32)
        props[1, 17:] = 0
        enc = rng.normal(0, 1.5, (B * P, K, 4)).astype(F)
        logits = rng.normal(0, 2.5, (B * P, K + 1)).astype(F)
        d = Arch.postprocess(s, dict(image_shape=t(np.array([B, H, W, 3], np.int32)), refined_box_encodings=t(enc),
                                     class_predictions_with_background=t(logits), proposal_boxes=t(props),
                                     num_proposals=t(nprop)))
        p = "post%d/" % case
        out[p + "props"], out[p + "nprop"], out[p + "enc"], out[p + "logits"] = props, nprop, enc, logits
        out[p + "params"] = np.array([thr, 0.6, max_det, H, W], np.float64)
        for k in ("detection_boxes", "detection_scores", "detection_classes", "num_detections"):
            out[p + k] = np.asarray(d[k], F)


def _crop_and_resize(image, boxes, box_ind, crop_size, extrapolation_value=0.0, **k):
    """tf.image.crop_and_resize (bilinear) restated: y = y1 (H-1) + i (y2 - y1)(H-1)/(ch-1) (centre when ch == 1),
    samples outside [0, H-1] give the extrapolation value."""
    image, boxes, box_ind = np.asarray(image, F), np.asarray(boxes, F), np.asarray(box_ind).astype(int)
    ch, cw = [int(v) for v in crop_size]
    _, H, W, C = image.shape
    out_ = np.full((len(boxes), ch, cw, C), 0.0 if extrapolation_value is None else extrapolation_value, F)   # None -> op default 0
    for n, (y1, x1, y2, x2) in enumerate(boxes):
        for i in range(ch):
            y = y1 * (H - 1) + i * (y2 - y1) * (H - 1) / (ch - 1) if ch > 1 else 0.5 * (y1 + y2) * (H - 1)
            if y < 0 or y > H - 1:
                continue
            t0, t1, ly = int(np.floor(y)), int(np.ceil(y)), y - np.floor(y)
            for j in range(cw):
                x = x1 * (W - 1) + j * (x2 - x1) * (W - 1) / (cw - 1) if cw > 1 else 0.5 * (x1 + x2) * (W - 1)
                if x < 0 or x > W - 1:
                    continue
                l0, l1, lx = int(np.floor(x)), int(np.ceil(x)), x - np.floor(x)
                im = image[box_ind[n]]
                top = im[t0, l0] + (im[t0, l1] - im[t0, l0]) * F(lx)
                bot = im[t1, l0] + (im[t1, l1] - im[t1, l0]) * F(lx)
                out_[n, i, j] = top + (bot - top) * F(ly)
    return t(out_)


def psroi_case(out):
    """utils/ops.py:462-609 `position_sensitive_crop_regions` (global_pool=True, the R-FCN configs' 3x3 bins of an
    18x18 crop): bin geometry, channel-group order, bin / spatial averaging -- with the crop_and_resize above."""
    ops_mod = M["object_detection.utils.ops"]
    tf.image.crop_and_resize = _crop_and_resize
    tf.add_n = lambda xs, **k: t(sum(np.asarray(x, F) for x in xs))
    rng = np.random.default_rng(59)
    D, R = 5, 7
    fmap = rng.normal(0, 1, (2, 9, 11, 9 * D)).astype(F)
    y0, x0 = rng.uniform(0, 0.6, R), rng.uniform(0, 0.6, R)
    boxes = np.stack([y0, x0, y0 + rng.uniform(0.1, 0.4, R), x0 + rng.uniform(0.1, 0.4, R)], 1).astype(F)
    boxes[0] = [-0.1, 0.2, 0.5, 1.1]                                  # partly outside: extrapolation 0
    box_ind = rng.integers(0, 2, R).astype(np.int32)
    got = ops_mod.position_sensitive_crop_regions(t(fmap), t(boxes), t(box_ind), [18, 18], [3, 3], global_pool=True)
    out["psroi/fmap"], out["psroi/boxes"], out["psroi/box_ind"] = fmap, boxes, box_ind
    out["psroi/out"] = np.asarray(got, F)


def main():
    out = {}
    psroi_case(out)
    rpn_cases(out)
    postprocess_case(out)
    refine_case(out)
    crop_case(out)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "graph_reference.npz"), **out)
    print("wrote graph_reference.npz", {k: np.asarray(v).tolist() for k, v in out.items() if "loss/" in k or "nprop" in k})


if __name__ == "__main__":
    main()
