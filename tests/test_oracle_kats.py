"""Pins the CPU oracle (oracle/) to the reference's own known-answer tests and to outputs of the
reference's importable NumPy code (tests/golden/np_reference.npz, made by tests/golden/make_golden.py).

Every test names the reference test it restates (paths under /root/reference/object_detection/).
CPU only; runs in seconds."""
import math
import os

import numpy as np
import torch

from oracle import assign as OA
from oracle import boxes as OB
from oracle import nn as ON
from oracle import postprocess as OP

F = np.float32
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "np_reference.npz"))


# ---- anchor_generators/grid_anchor_generator_test.py:25-73
def test_construct_single_anchor():
    exp = [[-121, -35, 135, 29], [-249, -67, 263, 61], [-505, -131, 519, 125], [-57, -67, 71, 61],
           [-121, -131, 135, 125], [-249, -259, 263, 253], [-25, -131, 39, 125], [-57, -259, 71, 253],
           [-121, -515, 135, 509]]
    got = OB.grid_anchors(1, 1, [0.5, 1.0, 2.0], [0.25, 1.0, 4.0], anchor_offset=(7, -3))
    np.testing.assert_allclose(got, exp, rtol=1e-6, atol=1e-4)


def test_construct_anchor_grid():
    exp = [[-2.5, -2.5, 2.5, 2.5], [-5., -5., 5., 5.], [-10., -10., 10., 10.], [-2.5, 16.5, 2.5, 21.5],
           [-5., 14., 5, 24], [-10., 9., 10, 29], [16.5, -2.5, 21.5, 2.5], [14., -5., 24, 5], [9., -10., 29, 10],
           [16.5, 16.5, 21.5, 21.5], [14., 14., 24, 24], [9., 9., 29, 29]]
    got = OB.grid_anchors(2, 2, [0.5, 1.0, 2.0], [1.0], (10, 10), (19, 19), (0, 0))
    np.testing.assert_allclose(got, exp, rtol=1e-6, atol=1e-5)


# ---- box_coders/faster_rcnn_box_coder_test.py:26-92 (scale factors passed explicitly there)
BOXES = [[10.0, 10.0, 20.0, 15.0], [0.2, 0.1, 0.5, 0.4]]
ANCHORS = [[15.0, 12.0, 30.0, 18.0], [0.1, 0.0, 0.7, 0.9]]


def test_box_coder_encode():
    exp = [[-0.5, -0.416666, -0.405465, -0.182321], [-0.083333, -0.222222, -0.693147, -1.098612]]
    np.testing.assert_allclose(OB.box_encode(BOXES, ANCHORS, None), exp, rtol=1e-5, atol=1e-6)


def test_box_coder_encode_with_scaling():
    exp = [[-1., -1.25, -1.62186, -0.911608], [-0.166667, -0.666667, -2.772588, -5.493062]]
    np.testing.assert_allclose(OB.box_encode(BOXES, ANCHORS, [2, 3, 4, 5]), exp, rtol=1e-5, atol=1e-6)


def test_box_coder_decode():
    codes = [[-0.5, -0.416666, -0.405465, -0.182321], [-0.083333, -0.222222, -0.693147, -1.098612]]
    np.testing.assert_allclose(OB.box_decode(codes, ANCHORS, None), BOXES, rtol=1e-5, atol=1e-5)


def test_box_coder_decode_with_scaling():
    codes = [[-1., -1.25, -1.62186, -0.911608], [-0.166667, -0.666667, -2.772588, -5.493062]]
    np.testing.assert_allclose(OB.box_decode(codes, ANCHORS, [2, 3, 4, 5]), BOXES, rtol=1e-5, atol=1e-5)


def test_box_coder_very_small_width():
    exp = [[-0.833333, 0., -21.128731, 0.510826]]
    got = OB.box_encode([[10.0, 10.0, 10.0000001, 20.0]], [[15.0, 12.0, 30.0, 18.0]], None)
    np.testing.assert_allclose(got, exp, rtol=1e-5, atol=1e-5)


# ---- core/region_similarity_calculator_test.py:25-36
def test_iou_similarity():
    c1 = [[4.0, 3.0, 7.0, 5.0], [5.0, 6.0, 10.0, 7.0]]
    c2 = [[3.0, 4.0, 6.0, 8.0], [14.0, 14.0, 15.0, 15.0], [0.0, 0.0, 20.0, 20.0]]
    exp = [[2.0 / 16.0, 0, 6.0 / 400.0], [1.0 / 16.0, 0.0, 5.0 / 400.0]]
    np.testing.assert_allclose(OB.iou(c1, c2), exp, rtol=1e-6)


# ---- reference NumPy code run in the build container (utils/np_box_ops.py:25-97)
def test_geometry_matches_reference_numpy():
    a, b = GOLD["iou_a"], GOLD["iou_b"]
    np.testing.assert_allclose(OB.area(b), GOLD["area"], rtol=1e-6)
    np.testing.assert_allclose(OB.intersection(a, b), GOLD["intersection"], rtol=1e-6, atol=1e-3)
    np.testing.assert_allclose(OB.iou(a, b), GOLD["iou"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(OB.ioa(a, b), GOLD["ioa"], rtol=1e-5, atol=1e-7)


def test_clip_and_prune_match_reference_numpy():
    """utils/np_box_list_ops.py:467-522 clip_to_window, :524-552 prune_outside_window."""
    b, w = GOLD["iou_b"], GOLD["window"]
    got, _ = OB.clip_to_window(b, w)
    np.testing.assert_array_equal(got, GOLD["clip_to_window"])
    pb, idx = OB.prune_outside_window(b, w)
    np.testing.assert_array_equal(idx, GOLD["prune_outside_window_idx"])
    np.testing.assert_array_equal(pb, GOLD["prune_outside_window_boxes"])


def test_nms_matches_reference_numpy():
    """utils/np_box_list_ops.py:185-257 non_max_suppression (suppress iff IoU > threshold)."""
    b, s = GOLD["nms_boxes_in"], GOLD["nms_scores_in"]
    for thr, mx in ((0.5, 50), (0.7, 300)):
        for fn in (OP.tf_non_max_suppression, OP.nms_vectorized):
            sel = fn(b, s, mx, thr)
            np.testing.assert_array_equal(b[sel], GOLD["nms_%g_%d_boxes" % (thr, mx)])
            np.testing.assert_array_equal(s[sel], GOLD["nms_%g_%d_scores" % (thr, mx)])


# ---- matchers/argmax_matcher_test.py:26-192
SIM = np.array([[1., 1, 1, 3, 1], [2, -1, 2, 0, 4], [3, 0, -1, 0, 0]])
SIM2 = np.array([[1, 1, 1, 3, 1], [-1, 0, -2, -2, -1], [3, 0, -1, 2, 0]], F)


def _cols(m):
    return np.nonzero(m >= 0)[0], m[m >= 0], np.nonzero(m == -1)[0]


def test_matcher_default_thresholds():
    m = OA.argmax_match(SIM, None)
    np.testing.assert_array_equal(m, [2, 0, 1, 0, 1])


def test_matcher_empty_rows():
    m = OA.argmax_match(np.zeros((0, 5), F), None)
    np.testing.assert_array_equal(m, [-1] * 5)


def test_matcher_matched_threshold():
    mc, mr, un = _cols(OA.argmax_match(SIM, 3.0))
    np.testing.assert_array_equal(mc, [0, 3, 4]); np.testing.assert_array_equal(mr, [2, 0, 1])
    np.testing.assert_array_equal(un, [1, 2])


def test_matcher_matched_and_unmatched_threshold():
    mc, mr, un = _cols(OA.argmax_match(SIM, 3.0, 2.0))
    np.testing.assert_array_equal(mc, [0, 3, 4]); np.testing.assert_array_equal(mr, [2, 0, 1])
    np.testing.assert_array_equal(un, [1])


def test_matcher_negatives_lower_than_unmatched_false():
    mc, mr, un = _cols(OA.argmax_match(SIM, 3.0, 2.0, negatives_lower_than_unmatched=False))
    np.testing.assert_array_equal(mc, [0, 3, 4]); np.testing.assert_array_equal(mr, [2, 0, 1])
    np.testing.assert_array_equal(un, [2])


def test_matcher_unmatched_row_without_and_with_force_match():
    mc, mr, un = _cols(OA.argmax_match(SIM2, 3.0, 2.0))
    np.testing.assert_array_equal(mc, [0, 3]); np.testing.assert_array_equal(mr, [2, 0])
    np.testing.assert_array_equal(un, [1, 2, 4])
    mc, mr, un = _cols(OA.argmax_match(SIM2, 3.0, 2.0, force_match_for_each_row=True))
    np.testing.assert_array_equal(mc, [0, 1, 3]); np.testing.assert_array_equal(mr, [2, 1, 0])
    np.testing.assert_array_equal(un, [2, 4])


# ---- core/target_assigner_test.py:412-466 (empty groundtruth) + fmA detector semantics
def test_assign_empty_groundtruth():
    anchors = [[0.0, 0.0, 0.5, 0.5], [0.5, 0.5, 1.0, 0.8], [0, 0.5, .5, 1.0], [.75, 0, 1.0, .25]]
    t = OA.assign_targets(anchors, np.zeros((0, 4), F), np.zeros((0, 4), F), [0, 0, 0, 0], 0.5)
    np.testing.assert_array_equal(t["cls_targets"], np.zeros((4, 4)))
    np.testing.assert_array_equal(t["cls_weights"], [1, 1, 1, 1])
    np.testing.assert_array_equal(t["reg_targets"], np.zeros((4, 4)))
    np.testing.assert_array_equal(t["reg_weights"], [0, 0, 0, 0])


def test_assign_detection_targets_and_weights():
    props = [[0, 0, 10, 10], [0, 0, 10, 9], [50, 50, 60, 60], [0, 0, 3, 3]]
    gt = [[0, 0, 10, 10], [50, 50, 61, 61]]
    cls = [[0, 0, 1], [0, 1, 0]]
    t = OA.assign_detection(props, gt, cls)
    np.testing.assert_array_equal(t["match"], [0, 0, 1, -1])
    np.testing.assert_array_equal(t["cls_targets"], [[0, 0, 1], [0, 0, 1], [0, 1, 0], [1, 0, 0]])
    np.testing.assert_array_equal(t["cls_weights"], [1, 1, 1, 1])
    np.testing.assert_array_equal(t["reg_weights"], [1, 1, 1, 0])
    np.testing.assert_allclose(t["reg_targets"][0], [0, 0, 0, 0], atol=1e-6)


def test_rpn_assign_ignore_band_and_force_match():
    anchors = [[0, 0, 10, 10], [0, 0, 10, 5], [0, 0, 10, 4.5], [20, 20, 30, 30]]
    gt = [[0, 0, 10, 10], [20, 20, 26, 26]]            # second GT has best IoU 0.36 with anchor 3
    t = OA.assign_proposal(anchors, gt)
    np.testing.assert_array_equal(t["match"], [0, -2, -2, 1])       # 0.5 / 0.45 ignored, forced match
    np.testing.assert_array_equal(t["cls_weights"], [1, 0, 0, 1])
    np.testing.assert_array_equal(t["cls_targets"][:, 0], [1, 0, 0, 1])


# ---- core/balanced_positive_negative_sampler_test.py:26-63 (counts; indices are key-defined here)
def test_balanced_sampler_counts():
    rng = np.random.default_rng(0)
    labels = np.array([True] * 100 + [False] * 200)
    keys = rng.random(300).astype(F)
    s = OA.balanced_subsample(np.ones(300, bool), 64, labels, 0.5, keys)
    assert s.sum() == 64 and s[labels].sum() == 32
    ind = np.array([True] * 90 + [False] * 210)                    # only 10 negatives available
    ind[290:] = True
    s = OA.balanced_subsample(ind, 64, labels, 0.5, keys)
    assert s[labels].sum() == 32 and s[~labels].sum() == 10 and not s[~ind].any()
    few = np.zeros(300, bool); few[:5] = True                      # 5 positives -> 59 negatives
    s = OA.balanced_subsample(np.ones(300, bool), 64, few, 0.5, keys)
    assert s[few].sum() == 5 and s[~few].sum() == 59


# ---- core/losses_test.py:98-118, :228-283
def test_smooth_l1_loss():
    pred = torch.tensor([[[2.5, 0, .4, 0], [0, 0, 0, 0], [0, 2.5, 0, .4]],
                         [[3.5, 0, 0, 0], [0, .4, 0, .9], [0, 0, 1.5, 0]]])
    w = torch.tensor([[2., 1, 1], [0, 3, 0]])
    np.testing.assert_allclose(ON.smooth_l1(pred, torch.zeros_like(pred), w).sum().item(), 7.695, rtol=1e-6)


def test_softmax_loss_and_anchorwise():
    pred = torch.tensor([[[-100., 100, -100], [100, -100, -100], [0, 0, -100], [-100, -100, 100]],
                         [[-100, 0, 0], [-100, 100, -100], [-100, 100, -100], [100, -100, -100]]])
    tgt = torch.tensor([[[0., 1, 0], [1, 0, 0], [1, 0, 0], [0, 0, 1]], [[0, 0, 1], [0, 1, 0], [0, 1, 0], [1, 0, 0]]])
    w = torch.tensor([[1, 1, .5, 1], [1., 1, 1, 0]])
    aw = ON.softmax_ce(pred, tgt, w)
    np.testing.assert_allclose(aw.sum().item(), -1.5 * math.log(.5), rtol=1e-6)
    exp = [[0, 0, -0.5 * math.log(.5), 0], [-math.log(.5), 0, 0, 0]]
    np.testing.assert_allclose(aw.numpy(), exp, atol=1e-6)


# ---- core/post_processing_test.py:43-76, :301-347
def test_multiclass_nms_select_with_shared_boxes():
    boxes = np.array([[[0, 0, 1, 1]], [[0, 0.1, 1, 1.1]], [[0, -0.1, 1, 0.9]], [[0, 10, 1, 11]], [[0, 10.1, 1, 11.1]],
                      [[0, 100, 1, 101]], [[0, 1000, 1, 1002]], [[0, 1000, 1, 1002.1]]], F)
    scores = np.array([[.9, 0.01], [.75, 0.05], [.6, 0.01], [.95, 0], [.5, 0.01], [.3, 0.01], [.01, .85], [.01, .5]], F)
    b, s, c = OP.multiclass_non_max_suppression(boxes, scores, 0.1, .5, 4)
    np.testing.assert_allclose(b, [[0, 10, 1, 11], [0, 0, 1, 1], [0, 1000, 1, 1002], [0, 100, 1, 101]])
    np.testing.assert_allclose(s, [.95, .9, .85, .3])
    np.testing.assert_array_equal(c, [0, 0, 1, 0])


def test_multiclass_nms_with_clip_window():
    boxes = np.array([[[0, 0, 10, 10]], [[1, 1, 11, 11]]], F)
    scores = np.array([[.9], [.75]], F)
    b, s, c = OP.multiclass_non_max_suppression(boxes, scores, 0.0, 0.5, 100, clip_window=[5, 4, 8, 7])
    np.testing.assert_allclose(b, [[5, 4, 8, 7]]); np.testing.assert_allclose(s, [.9])
    b, s, c = OP.multiclass_non_max_suppression(boxes, scores, 0.0, 0.5, 100, clip_window=[5, 4, 8, 7],
                                                change_coordinate_frame=True)
    np.testing.assert_allclose(b, [[0, 0, 1, 1]])


# ---- meta_architectures/faster_rcnn_meta_arch_test_lib.py:603-987 closed-form loss values
def test_second_stage_loc_loss_closed_form():
    """One positive proposal whose code is off by 5*ln(0.8)-ish in the reference test yields
    (-5 ln 0.8 - 0.5)/3; restated on the loss primitive: |d| = -5 ln 0.8 > 1 -> |d| - 0.5."""
    d = -5 * math.log(0.8)
    got = ON.smooth_l1(torch.tensor([[d, 0, 0, 0.]]), torch.zeros(1, 4), torch.ones(1), 1.0).item() / 3
    np.testing.assert_allclose(got, (-5 * math.log(.8) - 0.5) / 3, rtol=1e-6)
    # the fork's first stage uses sigma = 3 (fmA:391-392): |d| - 0.5/9
    got = ON.smooth_l1(torch.tensor([[d, 0, 0, 0.]]), torch.zeros(1, 4), torch.ones(1), 3.0).item() / 3
    np.testing.assert_allclose(got, (-5 * math.log(.8) - 0.5 / 9) / 3, rtol=1e-6)


# ---- TF kernel restatements: self-consistency against torch primitives
def test_crop_and_resize_identity_and_extrapolation():
    img = torch.arange(2 * 5 * 7 * 3, dtype=torch.float32).reshape(2, 5, 7, 3)
    full = ON.crop_and_resize(img, torch.tensor([[0., 0, 1, 1]]), torch.tensor([1]), (5, 7))
    torch.testing.assert_close(full[0], img[1])
    out = ON.crop_and_resize(img, torch.tensor([[-1., -1, -0.5, -0.5]]), torch.tensor([0]), (2, 2))
    assert not out.any()


def test_resize_bilinear_matches_formula():
    x = torch.rand(1, 3, 4, 2)
    y = ON.resize_bilinear(x, (6, 8))
    # src = dst * in/out (no half-pixel offset): output (0,0) equals input (0,0); row 2 reads source row 1
    torch.testing.assert_close(y[0, 0, 0], x[0, 0, 0])
    torch.testing.assert_close(y[0, 2, 0], x[0, 1, 0])
    torch.testing.assert_close(y[0, 1, 0], 0.5 * (x[0, 0, 0] + x[0, 1, 0]))


def test_same_padding_arithmetic():
    assert ON.same_pad(300, 3, 2) == (150, 0, 1)
    assert ON.same_pad(75, 3, 1) == (75, 1, 1)
    assert ON.same_pad(7, 1, 2) == (4, 0, 0)


# ---- utils/ops_test.py:711-781 position_sensitive_crop_regions(global_pool=True)
def _psroi(fmap, boxes, box_ind, D, bins, crop):
    from oracle.model import Oracle
    o = Oracle({}, {"architecture": "resnet_v1_50"}, bf16=False)
    return o.psroi(fmap, boxes, np.asarray(box_ind, np.int64), D, bins, crop)


def test_position_sensitive_constant_channels():
    img = torch.tensor([float(c) for c in range(1, 7)] * 6).reshape(1, 3, 2, 6)   # channel c holds c+1
    rng = np.random.default_rng(0)
    boxes = rng.random((2, 4)).astype(F)
    for mult in (1, 2):
        out = _psroi(img, boxes, [0, 0], 1, (3, 2), (3 * mult, 2 * mult))
        np.testing.assert_allclose(out.numpy(), [[3.5], [3.5]], rtol=1e-6)


def test_position_sensitive_equal_channels_and_single_bin():
    rng = np.random.default_rng(1)
    image = torch.arange(1, 10, dtype=torch.float32).reshape(1, 3, 3, 1)
    boxes = np.sort(rng.random((3, 2, 2)).astype(F), axis=1).reshape(3, 4)[:, [0, 2, 1, 3]]
    boxes = np.ascontiguousarray(boxes[:, [0, 1, 2, 3]])
    want = ON.crop_and_resize(image, torch.from_numpy(boxes), torch.zeros(3, dtype=torch.long), (2, 2)).mean((1, 2))
    # position-sensitive pooling of identical channels over a 2x2 bin grid with 1x1 crops per bin samples each
    # bin's centre; with a single bin it IS crop_and_resize + mean (ops_test.py:759-781)
    img2 = torch.rand(2, 3, 3, 4)
    b6 = rng.random((6, 4)).astype(F)
    bi = np.array([0, 0, 0, 1, 1, 1])
    got = _psroi(img2, b6, bi, 4, (1, 1), (2, 2))
    ref = ON.crop_and_resize(img2, torch.from_numpy(b6), torch.from_numpy(bi), (2, 2)).mean((1, 2))
    torch.testing.assert_close(got, ref)
    assert want.shape == (3, 1)


def test_second_stage_postprocess_hand_case():
    """fmA:1387-1469 + post_processing.py:25-164 on a case small enough to do by hand: zero encodings (decoded boxes ==
    proposals), IDENTITY scores, threshold 0.25, IoU 0.5.  Class 0: A .9 keeps, B .8 suppressed by A (IoU 100/105),
    C .3 keeps; class 1: only C .7 passes the threshold.  Merged by score: A/c0, C/c1, C/c0; boxes in the window frame."""
    from oracle import postprocess as PP
    props = np.array([[[0, 0, 10, 10], [0, 0, 10, 10.5], [20, 20, 30, 30]]], np.float32)
    enc = np.zeros((3, 2, 4), np.float32)
    logits = np.array([[0, .9, .1], [0, .8, .2], [0, .3, .7]], np.float32)
    b, s, c, n = PP.second_stage_postprocess(enc, logits, props, [3], (40, 40), 0.25, 0.5, 10, 5, score_mode="identity")
    assert n.tolist() == [3.0]
    np.testing.assert_allclose(s[0], [.9, .7, .3, 0, 0])
    assert c[0].tolist() == [0, 1, 0, 0, 0]
    np.testing.assert_allclose(b[0, :3], np.array([[0, 0, 10, 10], [20, 20, 30, 30], [20, 20, 30, 30]], np.float32) / 40)
    assert not b[0, 3:].any()
    # num_valid_boxes: only the first proposal is valid -> one detection
    b, s, c, n = PP.second_stage_postprocess(enc, logits, props, [1], (40, 40), 0.25, 0.5, 10, 5, score_mode="identity")
    assert n.tolist() == [1.0] and s[0, 0] == np.float32(.9)


def test_aux_labels_against_reference_record_writer():
    """N1 pinned: window soft labels, closeness labels and the 64x64 edge mask of data/aux_labels.py against the outputs
    of the reference's record writer (create_pascal_tf_record.py:121-421) RUN in this container under recording stubs
    for TensorFlow (tests/golden/make_aux_golden.py -> aux_reference.npz).  The reference samples its windows from
    Python's global RNG; the labels are recomputed here for exactly those windows."""
    import os
    from mtl_ssl_b200.data import aux_labels as A
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "aux_reference.npz"))
    K = int(g["num_classes"])
    rows = lambda arr: np.array([[float(t) for t in str(s).split()] for s in arr], np.float64)
    n_windows = 0
    for c in range(int(g["num_cases"])):
        p = "case%d/" % c
        H, W = [float(v) for v in g[p + "hw"]]
        boxes, classes = g[p + "boxes"], g[p + "classes"]
        # closeness: one row per object, K+1 columns, rounded to 3 decimals by get_string_label
        want = rows(g[p + "closeness"])
        got = A.closeness_labels(boxes, classes, H, W, K)
        assert got.shape == want.shape == (len(boxes), K + 1)
        np.testing.assert_allclose(got, want, atol=1e-6)
        # edge mask: [2, 64, 64] = (foreground, weight plane)
        eh, ew = [int(v) for v in g[p + "edgemask_hw"]]
        want_em = g[p + "edgemask"].reshape(-1, eh, ew)
        got_em = A.edgemask(boxes, H, W)
        assert got_em.shape == want_em.shape == (2, 64, 64)
        np.testing.assert_array_equal(got_em[0], want_em[0])
        np.testing.assert_allclose(got_em[1], want_em[1], rtol=1e-6)
        # window soft labels for the reference's own windows
        wl = rows(g[p + "window_labels"])
        assert wl.shape == (64, K + 1)
        for i in range(64):
            window = [g[p + "window_ymin"][i] * H, g[p + "window_xmin"][i] * W, g[p + "window_ymax"][i] * H,
                      g[p + "window_xmax"][i] * W]
            lab, bg = A.window_label(boxes, classes, window, K)
            # the window travelled through float32 and a divide / multiply by the image size: allow one rounding step
            np.testing.assert_allclose(lab, wl[i], atol=1.001e-3)
            assert abs(lab.sum() - 1.0) < 5e-3 and bg < 1.0
            n_windows += 1
    assert n_windows == 6 * 64


def test_losses_against_reference_loss_classes_run_on_a_numpy_tf_shim():
    """core/losses.py of the reference, RUN here on a NumPy stand-in for its ~15 TensorFlow ops
    (tests/golden/make_loss_golden.py): pins `Loss.__call__`'s rank-mismatch flattening (the window-class call of
    fmA:1839-1858 passes rank-2 logits with rank-3 soft labels), the weight / reduction plumbing of the softmax losses and
    the smooth-L1 loss with the fork's sigma -- against the oracle functions oracle/model.py builds its losses from."""
    import os
    import torch
    from oracle import nn as ON
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "loss_reference.npz"))
    T = torch.from_numpy
    # window-class soft-label CE: the oracle flattens [B,N,K+1] labels against [B*N,K+1] logits row by row
    B, N, K1 = g["win/labels"].shape
    got = ON.softmax_ce(T(g["win/logits"]), T(g["win/labels"]).reshape(B * N, K1)).numpy()
    assert g["win/loss"].shape == (B * N,)                         # one value per window, batch-major order
    np.testing.assert_allclose(got, g["win/loss"], rtol=2e-5, atol=2e-6)
    # one-hot CE with per-anchor weights, anchorwise and summed, v1 == v2 for constant labels
    ce = ON.softmax_ce(T(g["cls/logits"]), T(g["cls/labels"]), T(g["cls/weights"])).numpy()
    np.testing.assert_allclose(ce, g["cls/anchorwise"], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(ce.sum(), g["cls/scalar"], rtol=1e-5)
    np.testing.assert_allclose(g["cls/v2_scalar"], g["cls/scalar"], rtol=1e-6)
    # smooth L1, sigma 1 (second stage) and 3 (RPN, trap T1)
    for sigma in (1.0, 3.0):
        l = ON.smooth_l1(T(g["loc/pred"]), T(g["loc/target"]), T(g["loc/weights"]), sigma=sigma).numpy()
        np.testing.assert_allclose(l, g["loc/sigma%d" % sigma], rtol=2e-5, atol=1e-6)
    l3 = ON.smooth_l1(T(g["loc/pred"]), T(g["loc/target"]), T(g["loc/weights"]), sigma=3.0).numpy()
    np.testing.assert_allclose(l3.sum(), g["loc/scalar_sigma3"], rtol=1e-5)
    assert not np.allclose(g["loc/sigma1"], g["loc/sigma3"])
    # ignore_nan_targets: a NaN target contributes zero (it is replaced by the prediction)
    tn = g["loc/target_nan"]
    filled = np.where(np.isnan(tn), g["loc/pred"], tn)
    l = ON.smooth_l1(T(g["loc/pred"]), T(filled), T(g["loc/weights"]), sigma=1.0).numpy()
    np.testing.assert_allclose(l, g["loc/ignore_nan"], rtol=2e-5, atol=1e-6)


def test_target_assigner_against_reference_code_run_on_the_tf_shim():
    """core/target_assigner.py `TargetAssigner.assign` of the reference (IouSimilarity, ArgMaxMatcher incl. force-match,
    FasterRcnnBoxCoder, the dead crowd / ignore branch T6, and the fork's `extension=True` closeness targets
    `_create_mtl_targets`), EXECUTED here on tests/golden/tf_numpy_shim.py (make_assign_golden.py ->
    assign_reference.npz), against oracle/assign.py: matches and class targets bit-exact, box encodings to fp32
    rounding (log / divide)."""
    import os
    from oracle import assign as OA
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "assign_reference.npz"))
    saw_forced = False
    for c in range(int(g["num_cases"])):
        p = "det%d/" % c
        got = OA.assign_detection(g[p + "props"], g[p + "gt"], g[p + "cls"], g[p + "closeness"])
        assert np.array_equal(got["match"], g[p + "match"])
        assert np.array_equal(got["cls_targets"], g[p + "cls_targets"])
        assert np.array_equal(got["cls_weights"], g[p + "cls_weights"])
        assert np.array_equal(got["reg_weights"], g[p + "reg_weights"])
        np.testing.assert_allclose(got["reg_targets"], g[p + "reg_targets"], rtol=2e-5, atol=2e-6)
        assert np.array_equal(got["closeness_targets"], g[p + "closeness_targets"])          # gathered, not computed
        assert got["closeness_targets"][got["match"] < 0].sum() == 0 and (got["match"] >= 0).any()
        r = "rpn%d/" % c
        got = OA.assign_proposal(g[p + "props"], g[p + "gt"])
        assert np.array_equal(got["match"], g[r + "match"])
        assert np.array_equal(got["cls_targets"].reshape(-1), g[r + "cls_targets"].reshape(-1))
        assert np.array_equal(got["cls_weights"], g[r + "cls_weights"])
        assert np.array_equal(got["reg_weights"], g[r + "reg_weights"])
        np.testing.assert_allclose(got["reg_targets"], g[r + "reg_targets"], rtol=2e-5, atol=2e-6)
        # 0.7 / 0.3 thresholds: ignored anchors (-2) carry zero class weight, and every ground-truth box owns an anchor
        assert (got["cls_weights"][got["match"] == -2] == 0).all()
        assert set(range(len(g[p + "gt"]))) <= set(got["match"][got["match"] >= 0].tolist())
        iou_best = None
        saw_forced = saw_forced or bool(((got["match"] >= 0) & (g[r + "cls_weights"] > 0)).any())
    assert saw_forced


def test_second_stage_losses_against_reference_meta_arch_methods_run_on_the_tf_shim():
    """fmA `_loss_box_classifier` (:1670-1793, with the closeness loss), `_loss_refined_classifier` (:1795-1837),
    `_loss_window_class` (:1839-1858) and `_loss_edgemask` (:1860-1881) of the reference, EXECUTED here as unbound
    methods on the NumPy TF shim (tests/golden/make_stage2_loss_golden.py -> stage2_loss_reference.npz), against
    `Oracle.loss_second_stage`: the normalisers (T15), the padding indicator for short proposal lists, the class-selected
    box code (T12), the closeness column drop and its two normalisers (T11), the edge-mask label construction.
    Tolerance 2e-5 relative (fp32 log-softmax / bilinear resize sums)."""
    import os
    import torch
    from oracle.model import Oracle
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stage2_loss_reference.npz"))
    K, P, B, ncases = [int(v) for v in g["meta"]]
    cfg = dict(architecture="resnet_v1_101", num_classes=K, second_stage_batch_size=P,
               second_stage_localization_loss_weight=2.0, second_stage_classification_loss_weight=1.0,
               mtl=dict(window=True, closeness=True, edgemask=True, refine=True, window_class_loss_weight=1.0,
                        closeness_loss_weight=0.3, edgemask_loss_weight=1.0, refined_classification_loss_weight=1.0))
    o = Oracle({}, cfg, bf16=False)
    T = torch.from_numpy
    short = False
    for c in range(ncases):
        p = "case%d/" % c
        nprop = g[p + "nprop"]
        short = short or bool((nprop < P).any())
        out = dict(gts=[(g[p + "gt%d" % b], g[p + "cls%d" % b], g[p + "close%d" % b]) for b in range(B)],
                   prop_abs=g[p + "props"], nprop=nprop, refined_box_encodings=T(g[p + "enc"]),
                   class_predictions_with_background=T(g[p + "logits"]), closeness_predictions=T(g[p + "close_pred"]),
                   mtl_refined_class_predictions_with_background=T(g[p + "refined"]),
                   window_class_predictions=T(g[p + "win_pred"]), edgemask_predictions=T(g[p + "em_pred"]))
        examples = [dict(window_classes=g[p + "win_lab"][b], groundtruth_edgemask=g[p + "em_gt"][b]) for b in range(B)]
        got = o.loss_second_stage(out, examples)
        want = {k[len(p) + 5:]: float(g[k]) for k in g.files if k.startswith(p + "loss/")}
        assert set(got) == set(want) and len(want) == 6
        for k, v in want.items():
            np.testing.assert_allclose(float(got[k]), v, rtol=2e-5, err_msg=p + k)
    assert short                                                    # a padded proposal list was exercised


def _graph_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graph_reference.npz"))


def test_rpn_graph_code_against_reference_meta_arch_methods_run_on_the_tf_shim():
    """fmA `_remove_invalid_anchors_and_predictions` (:930-976), `_postprocess_rpn` in training mode (:1055-1132, through
    `batch_multiclass_non_max_suppression`, `_format_groundtruth_data`, `_unpad_proposals_and_sample_box_classifier_batch`,
    `_sample_box_classifier_minibatch`) and `_loss_rpn` (:1591-1668) of the reference, EXECUTED on the NumPy TF shim with a
    keyed shuffle in place of `tf.random_shuffle` (tests/golden/make_graph_golden.py), against the oracle methods the
    device path is compared with: kept anchor indices / proposal counts bit-exact, proposal boxes and scores to fp32
    rounding of exp / divide (1e-5), the two RPN losses 2e-5 relative."""
    import torch
    from oracle import boxes as OB
    from oracle.model import Oracle
    g = _graph_golden()
    K, P, M, MB, H, W, Hf, Wf, ncases = [int(v) for v in g["rpn_meta"]]
    cfg = dict(architecture="resnet_v1_101", num_classes=K, second_stage_batch_size=P, first_stage_max_proposals=M,
               nms_score_threshold=0.0, nms_iou_threshold=0.7, second_stage_balance_fraction=0.25,
               first_stage_minibatch_size=MB, first_stage_positive_balance_fraction=0.5,
               first_stage_localization_loss_weight=2.0, first_stage_objectness_loss_weight=1.0, mtl={})
    o = Oracle({}, cfg, bf16=False)
    short = False
    for c in range(ncases):
        p = "rpn%d/" % c
        anchors, keep = OB.prune_outside_window(g[p + "anchors_all"], (0, 0, H, W))
        assert 0 < len(keep) < len(g[p + "anchors_all"])
        assert np.array_equal(anchors, g[p + "anchors"])
        assert np.array_equal(g[p + "enc"][:, keep], g[p + "enc_kept"])
        assert np.array_equal(g[p + "logits"][:, keep], g[p + "logits_kept"])
        B = g[p + "enc"].shape[0]
        gts = []
        for b in range(B):
            gt_abs = OB.to_absolute_coordinates(g[p + "gt%d" % b] / np.array([H, W, H, W], np.float32), H, W)
            np.testing.assert_allclose(gt_abs, g[p + "gt_abs%d" % b], rtol=0, atol=0)       # _format_groundtruth_data
            cls_bg = np.concatenate([np.zeros((len(gt_abs), 1), np.float32), g[p + "cls%d" % b]], 1)
            gts.append((gt_abs, cls_bg, None))
            boxes, scores, cnt, _ = o.training_proposals(g[p + "enc_kept"][b], g[p + "logits_kept"][b], anchors, gt_abs,
                                                         cls_bg, g[p + "keys2"][b], H, W)
            assert cnt == int(g[p + "nprop"][b])
            short = short or cnt < P
            np.testing.assert_allclose(boxes, g[p + "prop_norm"][b], rtol=1e-5, atol=1e-6)
            np.testing.assert_allclose(scores, g[p + "prop_scores"][b], rtol=1e-5, atol=1e-7)
            assert (boxes[cnt:] == 0).all() and (np.diff(scores[:cnt]) <= 0).all()           # score order kept
        out = dict(anchors=anchors, gts=gts, rpn_box=torch.from_numpy(g[p + "enc_kept"]),
                   rpn_cls=torch.from_numpy(g[p + "logits_kept"]))
        got = o.loss_first_stage(out, g[p + "keys1"])
        for k in ("first_stage_localization_loss", "first_stage_objectness_loss"):
            np.testing.assert_allclose(float(got[k]), float(g[p + "loss/" + k]), rtol=2e-5, err_msg=p + k)
    assert short


def test_refiner_input_assembly_against_reference_predict_with_mtl_results_run_on_the_tf_shim():
    """fmA `predict_with_mtl_results` (:764-846) EXECUTED on the shim with a recording `predict_with_window`
    (logits = fixed linear map of the window box) and `slim.fully_connected` = x @ W + b: the five expanded windows per
    proposal and their flattening order, the [5, P, K+1] -> [P, 5(K+1)] transposition, the global closeness mean, the
    concatenation order and the residue -- against `Oracle.expanded_windows` / `Oracle.refine_logits` (batch 1, T9)."""
    import torch
    from oracle.model import Oracle
    g = _graph_golden()
    props, cls, close = g["refine/props"], g["refine/cls"], g["refine/close"]
    P, K1 = cls.shape
    exp = Oracle.expanded_windows(props)                                      # [5, 1, P, 4]
    assert g["refine/windows"].shape == (1, 5 * P, 4)
    np.testing.assert_allclose(exp.reshape(1, 5 * P, 4), g["refine/windows"], rtol=0, atol=1e-7)
    assert np.array_equal(exp[0, 0], props[0])                                # i = 0 is the proposal itself
    np.testing.assert_allclose(exp[4], np.broadcast_to(np.float32([0, 0, 1, 1]), exp[4].shape), atol=1e-6)
    o = Oracle({"MTLClassRefiner/fc1/weights": torch.from_numpy(g["refine/fcw"].T.copy()),
                "MTLClassRefiner/fc1/biases": torch.from_numpy(g["refine/fcb"])},
               dict(architecture="resnet_v1_101", mtl=dict(refine_residue=True)), bf16=False)
    ew = torch.from_numpy(exp.reshape(-1, 4) @ g["refine/wmat"])
    ref, net = o.refine_logits(torch.from_numpy(cls), ew, torch.from_numpy(close))
    np.testing.assert_allclose(net[:, K1:6 * K1].reshape(P, 5, K1).numpy(), g["refine/expand_window_class_predictions"],
                               rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ref.numpy(), g["refine/refined"], rtol=1e-5, atol=1e-5)


def test_roi_crop_box_index_map_against_reference_run_on_the_tf_shim():
    """fmA `_compute_second_stage_input_feature_maps` (:1304-1348) with recording stand-ins for crop_and_resize and
    max_pool2d: rank-3 proposals [B, P, 4] are flattened batch-major with box_ind = image index; a [1, n, 4] list reads
    image 0 throughout (T9); crop size and pool window come from the config."""
    g = _graph_golden()
    boxes = g["crop/boxes"]
    B, P, _ = boxes.shape
    assert np.array_equal(g["crop/flat_boxes"], boxes.reshape(-1, 4))
    assert np.array_equal(g["crop/box_ind"], np.repeat(np.arange(B), P))      # what Oracle.forward builds (`bi`)
    assert np.array_equal(g["crop/size_pool"], [14, 14, 2, 2])
    assert (g["crop/box_ind_rank3_batch1"] == 0).all() and len(g["crop/box_ind_rank3_batch1"]) == B * P


def test_oracle_full_step_runs_on_cpu_and_is_deterministic():
    """The whole restated step (forward, the eight losses, L2, autograd backward) on a 224x320 image with the
    model12 graph: guards the oracle's own plumbing on the CPU-only driver run (the values are what the -m gpu
    parity tests compare the device path with)."""
    import torch
    from helpers import load_config, oracle_config
    from mtl_ssl_b200.data import synthetic
    from oracle.model import Oracle
    import oracle.refparams as refparams
    H, W, B = 224, 320, 1
    cfg = load_config("model12.config", (("min_dimension: 600", "min_dimension: 224"),
                                         ("max_dimension: 1024", "max_dimension: 320"),
                                         ("first_stage_max_proposals: 300", "first_stage_max_proposals: 100"),
                                         ("second_stage_batch_size: 256", "second_stage_batch_size: 16")))
    ocfg = oracle_config(cfg)
    params, l2 = refparams.build_params(ocfg, seed=0)
    examples = synthetic.make_batch(1, B, H, W, ocfg["num_classes"], max_boxes=4, num_windows=8)
    keys = synthetic.make_sampler_keys(2, B, refparams.num_kept_anchors(ocfg, H, W), ocfg["first_stage_max_proposals"])
    img = torch.from_numpy(np.stack([e["image"] for e in examples]))
    totals = []
    for _ in range(2):
        orc = Oracle(params, ocfg, bf16=True)
        orc.require_grad([k for k in params if refparams.is_trainable(k)])
        out = orc.forward(img, examples, keys, H, W)
        losses = orc.loss(out, examples, keys, H, W)
        assert len(losses) == 8 and all(np.isfinite(float(v.detach())) for v in losses.values())
        totals.append(float((sum(losses.values()) + orc.regularization_loss(l2)).detach()))
    assert totals[0] == totals[1]
    (sum(losses.values()) + orc.regularization_loss(l2)).backward()
    g = orc.p["SecondStageBoxPredictor/ClassPredictor/weights"].grad
    assert g is not None and np.isfinite(g.numpy()).all() and float(g.abs().sum()) > 0
    assert int(out["nprop"][0]) == 16 and out["prop_norm"].shape == (1, 16, 4)


def test_second_stage_postprocess_against_reference_meta_arch_run_on_the_tf_shim():
    """fmA `postprocess` -> `_postprocess_box_classifier` (:1387-1469) -> `batch_multiclass_non_max_suppression`
    (post_processing.py:167-312: clip to the image, change_coordinate_frame, num_valid_boxes = num_proposals, per-class
    NMS, merge, top max_total, zero padding) EXECUTED on the shim, against oracle/postprocess.py
    `second_stage_postprocess`: detection counts and classes exact, boxes / scores to fp32 rounding of exp / softmax."""
    from oracle import postprocess as OP
    g = _graph_golden()
    for c in range(2):
        p = "post%d/" % c
        thr, iou, max_det, H, W = [float(v) for v in g[p + "params"]]
        b, s, cl, n = OP.second_stage_postprocess(g[p + "enc"], g[p + "logits"], g[p + "props"], g[p + "nprop"],
                                                  (int(H), int(W)), thr, iou, int(max_det), int(max_det))
        assert np.array_equal(n, g[p + "num_detections"]) and (n > 0).all()
        assert np.array_equal(cl, g[p + "detection_classes"])
        np.testing.assert_allclose(s, g[p + "detection_scores"], rtol=1e-5, atol=1e-7)
        np.testing.assert_allclose(b, g[p + "detection_boxes"], rtol=1e-5, atol=2e-6)
        assert b.max() <= 1.0 + 1e-6 and b.min() >= 0.0                 # normalised to the clip window
    assert (g["post0/num_detections"] == 20).all()                      # the top-20 cut was exercised
    assert (g["post1/num_detections"] < 100).any()                      # ... and the zero padding


def test_coco_examples_against_reference_mscoco_record_writer():
    """N1, COCO twin: data/mscoco.py against outputs of create_mscoco_tf_record.py `dict_to_tf_example` (:87-477) RUN here
    under recording stubs (tests/golden/make_coco_aux_golden.py): ground-truth boxes clamped into the image, auxiliary
    labels from the RAW boxes of all annotations, indexed by the raw (sparse) category id."""
    import os
    from mtl_ssl_b200.data import aux_labels as A, mscoco
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "coco_aux_reference.npz"))
    class_indices = [int(v) for v in g["class_indices"]]
    kmax = max(class_indices)
    cats = {c: {"id": c, "name": "c%d" % c} for c in class_indices}
    label_map = {"c%d" % c: c for c in class_indices}
    rows = lambda arr: np.array([[float(t) for t in str(s).split()] for s in arr], np.float64)
    clamped = False
    for c in range(int(g["num_cases"])):
        p = "case%d/" % c
        H, W = [int(v) for v in g[p + "hw"]]
        anns = [dict(bbox=[float(v) for v in bb], category_id=int(ci), iscrowd=int(cr), image_id=0, id=i)
                for i, (bb, ci, cr) in enumerate(zip(g[p + "bbox"], g[p + "category_id"], g[p + "iscrowd"]))]
        ex = mscoco.annotations_to_example(anns, np.zeros((H, W, 3), np.uint8), cats, label_map, kmax,
                                           np.random.default_rng(0), 64, kmax)
        want_gt = np.stack([g[p + "gt_" + k] for k in ("ymin", "xmin", "ymax", "xmax")], 1)
        np.testing.assert_allclose(ex["groundtruth_boxes"], want_gt, rtol=1e-6, atol=1e-7)
        clamped = clamped or bool((g[p + "bbox"][:, :2] < 0).any())
        assert ex["groundtruth_boxes"].min() >= 0 and ex["groundtruth_boxes"].max() <= 1
        assert np.array_equal(ex["groundtruth_labels"], g[p + "gt_label"])
        assert np.array_equal(ex["groundtruth_is_crowd"].astype(int), g[p + "gt_is_crowd"])
        want = rows(g[p + "closeness"])
        assert ex["groundtruth_closeness"].shape == want.shape == (len(want_gt), kmax + 1)
        np.testing.assert_allclose(ex["groundtruth_closeness"], want, atol=1e-6)
        eh, ew = [int(v) for v in g[p + "edgemask_hw"]]
        want_em = g[p + "edgemask"].reshape(-1, eh, ew)
        np.testing.assert_array_equal(ex["groundtruth_edgemask"][0], want_em[0])
        np.testing.assert_allclose(ex["groundtruth_edgemask"][1], want_em[1], rtol=1e-6)
        wl = rows(g[p + "window_labels"])
        assert wl.shape == (64, kmax + 1) and ex["window_classes"].shape == (64, kmax + 1)
        raw = mscoco._raw_boxes(anns)
        for i in range(64):
            window = [g[p + "window_ymin"][i] * H, g[p + "window_xmin"][i] * W, g[p + "window_ymax"][i] * H,
                      g[p + "window_xmax"][i] * W]
            lab, bg = A.window_label(raw, g[p + "category_id"], window, kmax)
            np.testing.assert_allclose(lab, wl[i], atol=1.001e-3)
        assert mscoco.get_image_id("/x/COCO_val2014_%012d.jpg" % int(str(g[p + "source_id"]))) == int(str(g[p + "source_id"]))
    assert clamped


def test_position_sensitive_crop_against_reference_ops_run_on_the_tf_shim():
    """utils/ops.py:462-609 `position_sensitive_crop_regions` (global_pool=True, 3x3 bins of an 18x18 crop as in the
    R-FCN configs) EXECUTED on the shim over random features (the bilinear crop inside is a NumPy restatement written
    independently of the oracle's): bin geometry, channel-group order, the two averaging steps and the zero extrapolation
    outside the map -- against `Oracle.psroi`; this also cross-checks oracle/nn.py `crop_and_resize` with a second
    implementation.  fp32 summation order differs: 2e-6 absolute."""
    import torch
    from oracle.model import Oracle
    g = _graph_golden()
    o = Oracle({}, dict(architecture="resnet_v1_101", mtl={}), bf16=False)
    D = g["psroi/fmap"].shape[-1] // 9
    got = o.psroi(torch.from_numpy(g["psroi/fmap"]), g["psroi/boxes"], g["psroi/box_ind"].astype(np.int64), D, (3, 3),
                  (18, 18)).numpy()
    want = g["psroi/out"].reshape(len(g["psroi/boxes"]), D)
    assert np.abs(want).max() > 0.05
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-6)


def test_oracle_inference_graph_with_refiner_runs_on_cpu():
    """`Oracle.forward(inference=True, inference_mtl=True)`: the evaluator's graph (evaluator.py:145-152): clipped anchors,
    unsampled proposals, closeness head and class refiner over the 5 expanded windows of every inference proposal; the
    refined logits are what `postprocess` scores (fmA:1040-1043)."""
    import torch
    from helpers import load_config, oracle_config
    from mtl_ssl_b200.data import synthetic
    from oracle import postprocess as OP
    from oracle.model import Oracle
    import oracle.refparams as refparams
    H, W = 224, 320
    cfg = load_config("model12.config", (("min_dimension: 600", "min_dimension: 224"),
                                         ("max_dimension: 1024", "max_dimension: 320"),
                                         ("first_stage_max_proposals: 300", "first_stage_max_proposals: 24"),
                                         ("second_stage_batch_size: 256", "second_stage_batch_size: 16")))
    ocfg = oracle_config(cfg)
    params, _ = refparams.build_params(ocfg, seed=0)
    examples = synthetic.make_batch(1, 1, H, W, ocfg["num_classes"], max_boxes=4, num_windows=8)
    img = torch.from_numpy(np.stack([e["image"] for e in examples]))
    orc = Oracle(params, ocfg, bf16=True)
    with torch.no_grad():
        plain = orc.forward(img, None, None, H, W, inference=True)
        out = orc.forward(img, None, None, H, W, inference=True, inference_mtl=True)
    M, K1 = 24, ocfg["num_classes"] + 1
    assert "mtl_refined_class_predictions_with_background" not in plain
    ref = out["mtl_refined_class_predictions_with_background"]
    assert ref.shape == (M, K1) and out["closeness_predictions"].shape == (M, K1)
    assert out["refine_in"].shape == (M, K1 + 5 * K1 + K1) and "window_class_predictions" not in out
    assert torch.equal(out["class_predictions_with_background"], plain["class_predictions_with_background"])
    # residue: refined = FC(refiner input) + original logits, so the two differ by the FC output only
    delta = (ref - out["class_predictions_with_background"]).abs().max()
    assert 0 < float(delta) < 10
    n = int(out["nprop"][0])
    b, s, c, nd = OP.second_stage_postprocess(out["refined_box_encodings"].numpy(), ref.numpy(), out["prop_abs"],
                                              out["nprop"], (H, W), 0.0, 0.6, 100, 100)
    assert 0 < nd[0] <= 100 and n > 0


def test_nms_and_iou_against_torchvision_ops():
    """A third, independent implementation of the greedy-NMS rule the TF 1.7 kernel follows (suppress iff IoU > thr,
    candidates by descending score): torchvision.ops.nms / box_iou on the CPU.  oracle/postprocess.py already agrees
    with the reference's own NumPy NMS (np_reference.npz); this guards the restatement from a second side."""
    import pytest
    tv = pytest.importorskip("torchvision.ops")
    import torch
    from oracle import boxes as OB, postprocess as OP
    rng = np.random.default_rng(77)
    for n, thr in ((400, 0.7), (1500, 0.5), (300, 0.3)):
        y0, x0 = rng.uniform(0, 500, n), rng.uniform(0, 800, n)
        b = np.stack([y0, x0, y0 + rng.uniform(5, 250, n), x0 + rng.uniform(5, 300, n)], 1).astype(np.float32)
        s = rng.random(n).astype(np.float32)
        xyxy = torch.from_numpy(b[:, [1, 0, 3, 2]].copy())
        want = tv.nms(xyxy, torch.from_numpy(s), thr).numpy()
        got = OP.nms_vectorized(b, s, n, thr)
        iou_tv = tv.box_iou(xyxy, xyxy).numpy()
        # (a pair whose IoU sits within float rounding of the threshold could legitimately flip: none in these draws)
        assert np.array_equal(got, want)
        assert np.array_equal(OP.tf_non_max_suppression(b[:200], s[:200], 50, thr),
                              tv.nms(xyxy[:200], torch.from_numpy(s[:200]), thr).numpy()[:50])
        np.testing.assert_allclose(OB.iou(b[:64], b[64:160]), iou_tv[:64, 64:160], rtol=1e-5, atol=1e-7)
