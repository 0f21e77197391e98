"""CPU tests of the VOC-style detection metrics: the reference's own unit-test vectors
(utils/metrics_test.py:24-75, utils/object_detection_evaluation_test.py:26-130) and golden outputs of the reference
evaluator RUN in this container on seeded random detections (tests/golden/make_eval_golden.py -> eval_reference.npz:
subsets / difficult boxes, tied scores, invalid boxes, per-class NMS)."""
import os

import numpy as np
import pytest

from mtl_ssl_b200.utils import detection_evaluation as E

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_reference.npz")


def test_metrics_reference_vectors():
    np.testing.assert_allclose(E.compute_cor_loc(np.array([100, 1, 5, 1, 1]), np.array([10, 0, 1, 0, 0])),
                               [0.1, 0, 0.2, 0, 0])
    got = E.compute_cor_loc(np.array([100, 0, 0, 1, 1]), np.array([10, 0, 1, 0, 0]))
    np.testing.assert_allclose(got, [0.1, np.nan, np.nan, 0, 0], equal_nan=True)
    scores = np.array([0.4, 0.3, 0.6, 0.2, 0.7, 0.1])
    labels = np.array([0, 1, 1, 0, 0, 1], bool)
    tp = np.array([0, 1, 1, 2, 2, 3], float)
    p, r = E.compute_precision_recall(scores, labels, 10)
    np.testing.assert_allclose(p, tp / np.arange(1, 7))
    np.testing.assert_allclose(r, tp / 10)
    precision = np.array([0.8, 0.76, 0.9, 0.65, 0.7, 0.5, 0.55, 0])
    recall = np.array([0.3, 0.3, 0.4, 0.4, 0.45, 0.45, 0.5, 0.5])
    want = np.sum(np.array([0.3, 0, 0.1, 0, 0.05, 0, 0.05, 0]) * np.array([0.9, 0.9, 0.9, 0.7, 0.7, 0.55, 0.55, 0]))
    assert abs(E.compute_average_precision(precision, recall) - want) < 1e-12
    p, r = E.compute_precision_recall(scores, np.zeros(6, bool), 0)
    assert p is None and r is None and np.isnan(E.compute_average_precision(p, r))
    assert E.compute_average_precision(np.array([]), np.array([])) == 0.0
    with pytest.raises(ValueError):
        E.compute_precision_recall(scores, labels, 2)                 # more true positives than ground truth
    with pytest.raises(ValueError):
        E.compute_average_precision(np.array([0.5, 0.5]), np.array([0.4, 0.3]))


def test_object_detection_evaluation_reference_case():
    ev = E.ObjectDetectionEvaluation(3)
    ev.add_single_ground_truth_image_info("img1", np.array([[0, 0, 1, 1], [0, 0, 2, 2], [0, 0, 3, 3]], float),
                                          np.array([0, 2, 0]))
    ev.add_single_ground_truth_image_info("img2", np.array([[10, 10, 11, 11], [500, 500, 510, 510], [10, 10, 12, 12]],
                                                           float), np.array([0, 0, 2]),
                                          np.array(["default", "", "default"]))
    ev.add_single_ground_truth_image_info("img3", np.array([[0, 0, 1, 1]], float), np.array([1]))
    ev.add_single_detected_image_info("img2", np.array([[10, 10, 11, 11], [100, 100, 120, 120], [100, 100, 220, 220]],
                                                       float), np.array([0.7, 0.8, 0.9]), np.array([0, 0, 2]))
    assert ev.num_gt_instances_per_class["default"].tolist() == [3, 1, 2]
    assert ev.num_gt_imgs_per_class.tolist() == [2, 1, 2]
    assert ev.groundtruth_subset["default"]["img2"].tolist() == [True, False, True]
    np.testing.assert_allclose(ev.scores_per_class["default"][0][0], [0.8, 0.7])
    assert ev.tp_fp_labels_per_class["default"][0][0].tolist() == [False, True]
    np.testing.assert_allclose(ev.scores_per_class["default"][2][0], [0.9])
    assert ev.tp_fp_labels_per_class["default"][2][0].tolist() == [False]
    assert ev.num_images_correctly_detected_per_class.tolist() == [0, 0, 0]
    ap, mean_ap, prec, rec, corloc, mean_corloc = ev.evaluate()
    np.testing.assert_allclose(ap["default"], [1.0 / 6.0, 0, 0])
    assert abs(mean_ap["default"] - 1.0 / 18) < 1e-12 and mean_corloc == 0.0
    np.testing.assert_allclose(prec["default"][0], [0, 0.5])
    np.testing.assert_allclose(rec["default"][0], [0, 1.0 / 3.0])
    assert len(prec["default"][1]) == 0
    np.testing.assert_allclose(prec["default"][2], [0])
    # adding the same image twice is ignored; length mismatch is an error
    ev.add_single_detected_image_info("img2", np.zeros((1, 4)), np.zeros(1), np.zeros(1, int))
    assert len(ev.scores_per_class["default"][0]) == 1
    with pytest.raises(ValueError):
        ev.add_single_detected_image_info("img9", np.zeros((2, 4)), np.zeros(1), np.zeros(2, int))
    with pytest.raises(ValueError):
        ev.add_single_ground_truth_image_info("img9", np.zeros((1, 4)), np.zeros(1, int), ["nosuchsubset"])
    with pytest.raises(ValueError):
        E.PerImageEvaluation(3, nms_type="soft-cubic")


def test_against_reference_evaluator_outputs():
    g = np.load(GOLD)
    ncase = len([k for k in g.files if k.endswith("/meta")])
    assert ncase == 7                                  # 4 x standard NMS, soft-linear, 2 x soft-gaussian
    for ci in range(ncase):
        seed, C, N, nms_max = [int(v) for v in g["case%d/meta" % ci]]
        subsets = tuple(str(s) for s in g["case%d/subsets" % ci])
        ev = E.ObjectDetectionEvaluation(C, matching_iou_threshold=0.5, nms_type=str(g["case%d/nms_type" % ci]),
                                         nms_iou_threshold=float(g["case%d/nms_iou" % ci]), nms_max_output_boxes=nms_max,
                                         soft_nms_sigma=float(g["case%d/sigma" % ci]), subset_names=subsets)
        for i in range(N):
            p = "case%d/img%d/" % (ci, i)
            sub = [str(s) for s in g[p + "sub"]]
            sub_arg = sub if (i % 2 or subsets != ("default",)) else None          # as in make_eval_golden.py
            ev.add_single_ground_truth_image_info("img%d" % i, g[p + "gb"], g[p + "gc"], sub_arg)
            ev.add_single_detected_image_info("img%d" % i, g[p + "db"], g[p + "ds"], g[p + "dc"])
        ap, mean_ap, _, _, corloc, mean_corloc = ev.evaluate()
        for s in subsets:
            np.testing.assert_allclose(ap[s], g["case%d/ap/%s" % (ci, s)], rtol=1e-12, atol=0, equal_nan=True)
            np.testing.assert_allclose(mean_ap[s], g["case%d/map/%s" % (ci, s)], rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(corloc, g["case%d/corloc" % ci], rtol=1e-12, equal_nan=True)
        np.testing.assert_allclose(mean_corloc, g["case%d/mean_corloc" % ci], rtol=1e-12, equal_nan=True)
    for k in range(4):
        p, r = E.compute_precision_recall(g["curve%d/scores" % k], g["curve%d/labels" % k], int(g["curve%d/num_gt" % k]))
        np.testing.assert_allclose(p, g["curve%d/precision" % k], rtol=1e-13)
        np.testing.assert_allclose(r, g["curve%d/recall" % k], rtol=1e-13)
        np.testing.assert_allclose(E.compute_average_precision(p, r), g["curve%d/ap" % k], rtol=1e-12)


def test_mtl_metrics_against_reference_outputs():
    """utils/mtl_util.py:20-87 (window mAP, closeness arg-max agreement) on the vectors the reference produced here;
    the edge-mask accuracy (skimage resize, absent: parity unpinned) by construction."""
    from mtl_ssl_b200.utils import mtl_metrics as M
    g = np.load(GOLD)
    n = int(g["mtl/n"])
    res = {k: [] for k in ("groundtruth_boxes", "detection_boxes", "window_classes_gt", "window_classes_dt",
                           "closeness_gt", "closeness_dt")}
    for i in range(n):
        for k in res:
            v = g["mtl/img%d/%s" % (i, k)]
            res[k].append([str(s) for s in v] if v.dtype.kind == "U" else v)
    m = M.get_mtl_metrics(res)
    np.testing.assert_allclose(m["mtl/window_map"], g["mtl/window_map"], rtol=1e-12)
    np.testing.assert_allclose(m["mtl/closeness_diff"], g["mtl/closeness_diff"], rtol=1e-12)
    assert 0.0 < m["mtl/window_map"] <= 1.0 and "mtl/edgemask_ap" not in m
    # numeric label rows (what data/tfrecord.py decodes to) give the same numbers as the record's text rows
    res2 = dict(res)
    res2["window_classes_gt"] = [[M._label_row(r) for r in rows] for rows in res["window_classes_gt"]]
    res2["closeness_gt"] = [[M._label_row(r) for r in rows] for rows in res["closeness_gt"]]
    m2 = M.get_mtl_metrics(res2)
    assert m2 == m
    # edge mask: a prediction whose foreground logit wins exactly on the ground-truth foreground scores 1.0
    fg = np.zeros((64, 64), np.float32)
    fg[16:48, 8:40] = 1
    gt = np.stack([fg, np.ones_like(fg)])
    logits = np.stack([-(fg * 2 - 1), fg * 2 - 1], -1)[None] * 3.0            # [1, 64, 64, 2] at the mask resolution
    perfect = M.get_mtl_metrics({"edgemask_gt": [gt], "edgemask_dt": [logits], "groundtruth_boxes": [],
                                 "detection_boxes": []})
    assert perfect["mtl/edgemask_ap"] == 1.0
    coarse = logits[:, ::2, ::2]                                              # half resolution: only the border can differ
    acc = M.get_mtl_metrics({"edgemask_gt": [gt], "edgemask_dt": [coarse], "groundtruth_boxes": [],
                             "detection_boxes": []})["mtl/edgemask_ap"]
    assert 0.95 < acc <= 1.0
    inverted = M.get_mtl_metrics({"edgemask_gt": [gt], "edgemask_dt": [-logits], "groundtruth_boxes": [],
                                  "detection_boxes": []})["mtl/edgemask_ap"]
    assert inverted == 0.0


def test_eval_util_pascal_metric_assembly():
    """eval_util.py:233-372: label offset, `difficult` -> not in any subset, metric names."""
    from mtl_ssl_b200 import eval_util
    cats = [{"id": 1, "name": "a"}, {"id": 2, "name": "b"}]
    gb = np.array([[0, 0, 10, 10], [20, 20, 40, 40], [50, 50, 60, 60]], float)
    gc = np.array([1, 2, 1])
    lists = {"detection_boxes": [gb[:2]], "detection_scores": [np.array([.9, .8])], "detection_classes": [gc[:2]],
             "image_id": ["7"], "groundtruth_boxes": [gb], "groundtruth_classes": [gc],
             "difficult": [np.array([0, 0, 1])]}
    m = eval_util.evaluate_detection_results_pascal_voc(lists, cats, corloc_summary=True)
    assert m["Subset default    mAP@0.5IOU"] == 1.0 and m["Subset default    mAP@0.5IOU/b"] == 1.0
    assert m["CorLoc/CorLoc@0.5IOU"] == 1.0 and m["PerformanceByCategory/CorLoc@0.5IOU/a"] == 1.0
    lists["difficult"] = [np.array([0, 0, 0])]                     # now the third box counts: recall of class a is 1/2
    m = eval_util.evaluate_detection_results_pascal_voc(lists, cats)
    assert m["Subset default    mAP@0.5IOU/a"] == 0.5 and abs(m["Subset default    mAP@0.5IOU"] - 0.75) < 1e-12
    lists["groundtruth_subset"] = [np.array(["easy", "easy|hard", "hard"])]
    m = eval_util.evaluate_detection_results_pascal_voc(lists, cats)
    assert m["Subset easy       mAP@0.5IOU/a"] == 1.0 and m["Subset hard       mAP@0.5IOU/b"] == 1.0
    # in subset 'hard' box 0 is difficult: its detection is dropped, the remaining class-a box is missed
    assert m["Subset hard       mAP@0.5IOU/a"] == 0.0
    with pytest.raises(ValueError):
        eval_util.evaluate_detection_results_pascal_voc({"detection_boxes": []}, cats)


def test_eval_util_against_reference_function_outputs():
    """eval_util.evaluate_detection_results_pascal_voc of the reference, RUN here under stubs for matplotlib / TensorFlow
    (tests/golden/make_eval_golden.py): identical metric names and values on random detections with `difficult` flags and
    two subsets."""
    from mtl_ssl_b200 import eval_util
    g = np.load(GOLD)
    n = int(g["voc/n"])
    keys = ("detection_boxes", "detection_scores", "detection_classes", "image_id", "groundtruth_boxes",
            "groundtruth_classes", "difficult", "groundtruth_subset")
    lists = {k: [g["voc/%s/%d" % (k, i)] for i in range(n)] for k in keys}
    lists["image_id"] = [str(v) for v in lists["image_id"]]
    lists["groundtruth_subset"] = [np.array([str(s) for s in v]) for v in lists["groundtruth_subset"]]
    cats = [{"id": i + 1, "name": "cat%d" % (i + 1)} for i in range(4)]
    m = eval_util.evaluate_detection_results_pascal_voc(lists, cats, corloc_summary=True)
    names = [str(s) for s in g["voc/metric_names"]]
    assert sorted(m) == names
    np.testing.assert_allclose([m[k] for k in names], g["voc/metric_values"], rtol=1e-12, equal_nan=True)
    assert any(k.startswith("Subset big") for k in names) and any(k.startswith("Subset all") for k in names)


# ----------------------------------------------------------------------------- MS-COCO metrics (pycocotools restated)
def _xywh_iou_box(gt, iou):
    """A box [x, y, w, h] sharing gt's top-left corner and height whose IoU with gt is `iou` (narrower box)."""
    return [gt[0], gt[1], gt[2] * iou, gt[3]]


def test_coco_bbox_eval_hand_computed_cases():
    """COCOeval('bbox') restated in utils/coco_evaluation.py (parity unpinned: pycocotools is absent), on cases small
    enough to compute by hand:
    A  one image, two ground-truth boxes; detections: 0.9 -> exact hit, 0.8 -> miss, 0.7 -> IoU 0.62 with the second box.
       IoU thresholds .5/.55/.6 see tp,fp,tp: precision envelope [1, 2/3, 2/3] at recalls [.5, .5, 1] -> the 101-point
       mean is (51 + 50 * 2/3) / 101; the seven higher thresholds see tp,fp,fp -> 51 / 101.
    B  a crowd region absorbs any number of detections without penalty and does not count as ground truth.
    C  area ranges: a 20x20 ground truth counts only for 'small'; a detection matched to an out-of-range ground truth
       is ignored, an unmatched detection outside the range as well.
    D  max detections: with the hit ranked second, AR_max1 = 0 and AR_max10 = 1."""
    from mtl_ssl_b200.utils.coco_evaluation import CocoBBoxEval, bbox_iou
    g1, g2 = [10.0, 10.0, 100.0, 100.0], [200.0, 50.0, 120.0, 100.0]
    np.testing.assert_allclose(bbox_iou([_xywh_iou_box(g2, 0.62)], [g2], [0]), [[0.62]], rtol=1e-12)
    gts = [dict(id=1, image_id=1, category_id=1, bbox=g1, area=100.0 * 100.0, iscrowd=0),
           dict(id=2, image_id=1, category_id=1, bbox=g2, area=120.0 * 100.0, iscrowd=0)]
    dts = [dict(image_id=1, category_id=1, bbox=g1, score=0.9),
           dict(image_id=1, category_id=1, bbox=[400.0, 300.0, 100.0, 100.0], score=0.8),
           dict(image_id=1, category_id=1, bbox=_xywh_iou_box(g2, 0.62), score=0.7)]
    s = CocoBBoxEval(gts, dts).run()
    hi, lo = (51 + 50 * 2.0 / 3.0) / 101, 51.0 / 101
    np.testing.assert_allclose(s[1], hi, rtol=1e-9)                      # AP50
    np.testing.assert_allclose(s[2], lo, rtol=1e-9)                      # AP75
    np.testing.assert_allclose(s[0], (3 * hi + 7 * lo) / 10, rtol=1e-9)  # AP@[.5:.95]
    assert s[3] == -1 and s[4] == -1                                     # no small / medium ground truth
    np.testing.assert_allclose(s[5], s[0], rtol=1e-12)                   # everything is 'large'
    np.testing.assert_allclose(s[6], 0.5, rtol=1e-12)                    # AR_max1: the best detection finds one of two
    np.testing.assert_allclose(s[8], (3 * 1.0 + 7 * 0.5) / 10, rtol=1e-12)
    # perfect detections
    s = CocoBBoxEval(gts, [dict(image_id=1, category_id=1, bbox=g["bbox"], score=0.5 + 0.1 * i)
                           for i, g in enumerate(gts)]).run()
    np.testing.assert_allclose(s[[0, 1, 2, 5, 7, 8]], 1.0, rtol=1e-12)
    # B: crowd
    crowd = dict(id=3, image_id=1, category_id=1, bbox=[0.0, 300.0, 300.0, 200.0], area=60000.0, iscrowd=1)
    inside = [dict(image_id=1, category_id=1, bbox=[10.0 + 40 * i, 320.0, 30.0, 60.0], score=0.95 - 0.01 * i)
              for i in range(4)]                                         # four detections inside the crowd region
    s = CocoBBoxEval(gts + [crowd], inside + [dict(image_id=1, category_id=1, bbox=g["bbox"], score=0.5)
                                               for g in gts]).run()
    np.testing.assert_allclose(s[[0, 1, 2]], 1.0, rtol=1e-12)            # ranked above the hits, yet no false positive
    s_fp = CocoBBoxEval(gts, inside + [dict(image_id=1, category_id=1, bbox=g["bbox"], score=0.5) for g in gts]).run()
    assert s_fp[1] < 0.4                                                 # without the crowd annotation they are
    # C: area ranges
    small = dict(id=4, image_id=2, category_id=1, bbox=[5.0, 5.0, 20.0, 20.0], area=400.0, iscrowd=0)
    dsmall = dict(image_id=2, category_id=1, bbox=[5.0, 5.0, 20.0, 20.0], score=0.6)
    stray = dict(image_id=2, category_id=1, bbox=[100.0, 100.0, 200.0, 200.0], score=0.9)   # large, unmatched
    s = CocoBBoxEval([small], [dsmall, stray]).run()
    np.testing.assert_allclose(s[3], 1.0, rtol=1e-12)                    # small: the stray detection is out of range
    np.testing.assert_allclose(s[1], 0.5, rtol=1e-9)                     # all areas: fp ranked first -> precision 1/2
    assert s[4] == -1 and s[5] == -1 and s[9] == 1.0
    # D: max detections
    s = CocoBBoxEval([small], [dsmall, dict(stray, bbox=[100.0, 100.0, 10.0, 10.0])]).run()
    assert s[6] == 0.0 and s[7] == 1.0 and s[8] == 1.0
    # per-category restriction and categories without ground truth
    other = dict(id=5, image_id=1, category_id=2, bbox=g1, area=1e4, iscrowd=0)
    s2 = CocoBBoxEval(gts + [other], dts, cat_ids=[2]).run()
    assert s2[0] == 0.0 and s2[8] == 0.0
    s12 = CocoBBoxEval(gts + [other], dts).run()
    np.testing.assert_allclose(s12[1], (hi + 0.0) / 2, rtol=1e-9)        # mean over the two categories


def test_coco_evaluation_wrapper_and_eval_util(tmp_path):
    """`CocoEvaluation` (object_detection_evaluation.py:294-427) + `evaluate_detection_results_coco`
    (eval_util.py:393-548): [ymin,xmin,ymax,xmax] -> [x,y,w,h], category = class + 1, at most 100 rows per image in
    descending score, ground truth from the annotation file, metric names 'COCO_Eval/<All|name>/<metric>'."""
    import json
    from mtl_ssl_b200 import eval_util
    from mtl_ssl_b200.data.mscoco import CocoIndex
    from mtl_ssl_b200.utils.coco_evaluation import CocoEvaluation
    ann = {"images": [{"id": 1, "file_name": "1.jpg", "height": 480, "width": 640},
                      {"id": 2, "file_name": "2.jpg", "height": 480, "width": 640}],
           "categories": [{"id": 1, "name": "a"}, {"id": 2, "name": "b"}],
           "annotations": [{"id": 1, "image_id": 1, "category_id": 1, "bbox": [10.0, 20.0, 100.0, 50.0], "area": 5000.0, "iscrowd": 0},
                           {"id": 2, "image_id": 2, "category_id": 2, "bbox": [30.0, 40.0, 200.0, 150.0], "area": 30000.0, "iscrowd": 0}]}
    path = str(tmp_path / "instances_eval.json")
    open(path, "w").write(json.dumps(ann))
    ev = CocoEvaluation(2)
    ev.add_single_detected_image_info(1, np.array([[20.0, 10.0, 70.0, 110.0]]), np.array([0.9]), np.array([0]))
    np.testing.assert_allclose(ev.detection_result, [[1, 10.0, 20.0, 100.0, 50.0, 0.9, 1]])
    many = np.tile(np.array([[40.0, 30.0, 190.0, 230.0]]), (150, 1)) + np.arange(150)[:, None] * 1e-3
    ev.add_single_detected_image_info(2, many, np.linspace(0.1, 0.8, 150), np.ones(150, int))
    assert len(ev.detection_result) == 101 and ev.detection_result[1, 5] == 0.8       # best 100, descending
    assert (np.diff(ev.detection_result[1:, 5]) <= 0).all() and (ev.detection_result[1:, 6] == 2).all()
    m = ev.evaluate([0, 1, 2], path)
    assert sorted(m) == [0, 1, 2] and all(len(v) == 12 for v in m.values())
    np.testing.assert_allclose([m[0][1], m[1][1]], 1.0, rtol=1e-12)
    assert ev.evaluate([0], str(tmp_path / "missing.json")) is None
    with pytest.raises(ValueError):
        ev.add_single_detected_image_info(3, np.zeros((2, 4)), np.zeros(1), np.zeros(2))
    lists = dict(image_id=["1", "2"], groundtruth_boxes=[np.array([[20.0, 10.0, 70.0, 110.0]]), np.array([[40.0, 30.0, 190.0, 230.0]])],
                 groundtruth_classes=[np.array([1]), np.array([2])],
                 detection_boxes=[np.array([[20.0, 10.0, 70.0, 110.0]]), np.array([[40.0, 30.0, 190.0, 230.0], [0.0, 0.0, 50.0, 50.0]])],
                 detection_scores=[np.array([0.9]), np.array([0.6, 0.7])], detection_classes=[np.array([1]), np.array([2, 2])])
    cats = [{"id": 1, "name": "a"}, {"id": 2, "name": "b"}]
    got = eval_util.evaluate_detection_results_coco(lists, cats, label_id_offset=1, eval_ann_filename=CocoIndex(path))
    assert sorted(got) == ["COCO_Eval/All/AP", "COCO_Eval/a/AP", "COCO_Eval/b/AP"]
    np.testing.assert_allclose(got["COCO_Eval/a/AP"], 1.0)
    np.testing.assert_allclose(got["COCO_Eval/b/AP"], 0.5, rtol=1e-9)     # false positive ranked above the hit
    np.testing.assert_allclose(got["COCO_Eval/All/AP"], 0.75, rtol=1e-9)
    assert eval_util.evaluate_detection_results_coco(lists, cats, 1, eval_ann_filename=str(tmp_path / "none.json")) == {}


def test_detection_metrics_follow_the_pipeline_eval_config(tmp_path):
    """evaluator.py:312-323: the metric function and its options come from `eval_config` -- model22.config (COCO) asks for
    'coco_metrics', all twelve metric indices, all-categories only; model12.config (VOC) for the PASCAL metrics."""
    import json
    from helpers import load_config
    from mtl_ssl_b200 import evaluator
    lists = dict(image_id=["1"], groundtruth_boxes=[np.array([[20.0, 10.0, 70.0, 110.0]])], groundtruth_classes=[np.array([1])],
                 detection_boxes=[np.array([[20.0, 10.0, 70.0, 110.0]])], detection_scores=[np.array([0.9])],
                 detection_classes=[np.array([1])])
    cats = [{"id": 1, "name": "a"}, {"id": 2, "name": "b"}]
    ann = {"images": [{"id": 1, "file_name": "1.jpg", "height": 480, "width": 640}],
           "categories": [{"id": 1, "name": "a"}, {"id": 2, "name": "b"}],
           "annotations": [{"id": 1, "image_id": 1, "category_id": 1, "bbox": [10.0, 20.0, 100.0, 50.0], "area": 5000.0, "iscrowd": 0}]}
    path = str(tmp_path / "instances.json")
    open(path, "w").write(json.dumps(ann))
    coco_cfg = load_config("model22.config").eval_config
    assert coco_cfg.metrics_set == "coco_metrics" and list(coco_cfg.coco_eval_options.eval_metric_index) == list(range(12))
    m = evaluator.detection_metrics(lists, cats, coco_cfg, eval_ann_filename=path)
    assert sorted(m) == sorted("COCO_Eval/All/" + n for n in
                               ("AP", "AP_IoU50", "AP_IoU75", "AP_small", "AP_medium", "AP_large", "AR_max1", "AR_max10",
                                "AR_max100", "AR_small", "AR_medium", "AR_large"))
    # tp / (tp + fp + eps), as pycocotools: "1" is 1 - 2e-16
    assert abs(m["COCO_Eval/All/AP"] - 1.0) < 1e-12 and abs(m["COCO_Eval/All/AP_medium"] - 1.0) < 1e-12
    assert m["COCO_Eval/All/AP_small"] == -1.0
    voc_cfg = load_config("model12.config").eval_config
    m = evaluator.detection_metrics(lists, cats, voc_cfg)
    assert m["Subset default    mAP@0.5IOU/a"] == 1.0 and "CorLoc/CorLoc@0.5IOU" in m
    with pytest.raises(ValueError):
        evaluator.detection_metrics(lists, cats, metrics_set="open_images_metrics")


def test_evaluate_loop_plumbing_with_a_stubbed_inference(monkeypatch, tmp_path):
    """evaluator.evaluate around a stubbed `run_inference` (the device part is covered by the -m gpu tests): per-image
    results are collected into lists, image ids come from 'source_id' for the COCO metrics, the PASCAL and the COCO metric
    sets are both reachable, and the switch for the refiner is passed through."""
    import json
    from mtl_ssl_b200 import evaluator
    from mtl_ssl_b200.data.mscoco import CocoIndex
    seen = []

    def fake(model, ex, use_refiner=False):
        seen.append(use_refiner)
        b = np.asarray(ex["boxes"], np.float32)
        return dict(detection_boxes=b, detection_scores=np.linspace(0.9, 0.5, len(b)).astype(np.float32),
                    detection_classes=np.asarray(ex["classes"]), groundtruth_boxes=b,
                    groundtruth_classes=np.asarray(ex["classes"]))

    monkeypatch.setattr(evaluator, "run_inference", fake)
    examples = [dict(source_id="11", boxes=[[20.0, 10.0, 70.0, 110.0]], classes=[1]),
                dict(source_id="12", boxes=[[40.0, 30.0, 190.0, 230.0], [5.0, 5.0, 50.0, 60.0]], classes=[2, 1])]
    cats = [{"id": 1, "name": "a"}, {"id": 2, "name": "b"}]
    m = evaluator.evaluate(None, examples, cats)
    assert m["Subset default    mAP@0.5IOU"] == 1.0 and m["CorLoc/CorLoc@0.5IOU"] == 1.0 and seen == [False, False]
    anns, k = [], 0
    for ex in examples:
        for (y0, x0, y1, x1), c in zip(ex["boxes"], ex["classes"]):
            k += 1
            anns.append(dict(id=k, image_id=int(ex["source_id"]), category_id=c, bbox=[x0, y0, x1 - x0, y1 - y0],
                             area=(x1 - x0) * (y1 - y0), iscrowd=0))
    coco = CocoIndex({"images": [{"id": 11, "file_name": "a.jpg"}, {"id": 12, "file_name": "b.jpg"}],
                      "categories": cats, "annotations": anns})
    m = evaluator.evaluate(None, examples, cats, metrics_set="coco_metrics", eval_ann_filename=coco, use_refiner=True)
    assert abs(m["COCO_Eval/All/AP"] - 1.0) < 1e-12 and abs(m["COCO_Eval/b/AP"] - 1.0) < 1e-12 and seen[-2:] == [True, True]


def test_submission_format_writers(tmp_path):
    """eval_util.py:884-930 (`submission_format_output`): COCO result JSON ([x, y, w, h], raw category ids) and PASCAL
    comp4 files (xmin ymin xmax ymax, image id without extension)."""
    import json
    from mtl_ssl_b200 import eval_util
    lists = dict(image_id=["139", "285"], detection_boxes=[np.array([[20.0, 10.0, 70.5, 110.25]]), np.array([[1.0, 2.0, 3.0, 5.0], [0.0, 0.0, 9.0, 9.0]])],
                 detection_scores=[np.array([0.98765]), np.array([0.5, 0.25])], detection_classes=[np.array([3]), np.array([90, 1])])
    (path,) = eval_util.save_detection_results_for_submission(lists, [], str(tmp_path), "coco_metrics")
    rows = json.load(open(path))
    assert rows[0] == {"image_id": 139, "category_id": 3, "bbox": [10.0, 20.0, 100.2, 50.5], "score": 0.988}
    assert [r["category_id"] for r in rows] == [3, 90, 1] and rows[1]["bbox"] == [2.0, 1.0, 3.0, 2.0]
    cats = [{"id": 1, "name": "aeroplane"}, {"id": 3, "name": "bird"}, {"id": 90, "name": "x"}]
    lists["image_id"] = ["000139.jpg", "000285.png"]
    files = eval_util.save_detection_results_for_submission(lists, cats, str(tmp_path), "pascal_voc_metrics")
    assert sorted(os.path.basename(f) for f in files) == ["comp4_det_test_aeroplane.txt", "comp4_det_test_bird.txt", "comp4_det_test_x.txt"]
    bird = open(os.path.join(str(tmp_path), "detection_results", "comp4_det_test_bird.txt")).read()
    assert bird == "000139 0.987650 10.000000 20.000000 110.250000 70.500000\n"
    with pytest.raises(ValueError):
        eval_util.save_detection_results_for_submission(lists, cats, str(tmp_path), "other")


def test_main_metric_and_best_checkpoint_bookkeeping(tmp_path):
    """eval_util.py:932-990 (`save_best_ckpt`): the ranking metric per metrics_set / main_subset, the best/ directory with
    summary.json, replaced only by a higher (non-NaN) value."""
    import json
    from mtl_ssl_b200 import eval_util
    m = {"Subset all        mAP@0.5IOU": 0.61, "Subset all        mAP@0.5IOU/cat": 0.7, "Subset big        mAP@0.5IOU": 0.8,
         "CorLoc/CorLoc@0.5IOU": 0.9}
    assert eval_util.main_metric(m) == ("Subset all        mAP@0.5IOU", 0.61)
    assert eval_util.main_metric(m, main_subset="big")[1] == 0.8
    assert eval_util.main_metric({"COCO_Eval/All/AP": 0.33, "COCO_Eval/All/AP_IoU50": 0.5}, "coco_metrics") == ("COCO_Eval/All/AP", 0.33)
    with pytest.raises(KeyError):
        eval_util.main_metric({"Subset default    mAP@0.5IOU": 0.5})
    saved = []
    ck = str(tmp_path / "model.ckpt-100")
    assert eval_util.save_best_checkpoint(m, ck, saved.append) is True
    assert saved == [str(tmp_path / "best" / "model.ckpt")]
    s = json.load(open(str(tmp_path / "best" / "summary.json")))
    assert s["mAP"] == 0.61 and s["checkpoint_file"] == ck and s["CorLoc/CorLoc@0.5IOU"] == 0.9
    assert eval_util.save_best_checkpoint(dict(m, **{"Subset all        mAP@0.5IOU": 0.5}), ck, saved.append) is False
    assert eval_util.save_best_checkpoint(dict(m, **{"Subset all        mAP@0.5IOU": float("nan")}), ck, saved.append) is False
    assert eval_util.save_best_checkpoint(dict(m, **{"Subset all        mAP@0.5IOU": 0.7}), ck, saved.append) is True
    assert len(saved) == 2 and json.load(open(str(tmp_path / "best" / "summary.json")))["mAP"] == 0.7
