"""Generates tests/golden/eval_reference.npz by RUNNING the reference's NumPy evaluator
(/root/reference/object_detection/utils/{object_detection_evaluation,per_image_evaluation,metrics}.py) in this
container on seeded random detections.  The reference is Python-2 / NumPy-1.x code: the aliases it needs (np.bool,
np.float, np.NAN, xrange) and a stub for the pycocotools import are injected here, nothing in the reference is modified.
Run from the repo root:  python tests/golden/make_eval_golden.py"""
import builtins
import os
import sys
import types

import numpy as np

np.bool, np.float, np.NAN = bool, float, np.nan
builtins.xrange = range
for name in ("pycocotools", "pycocotools.coco", "pycocotools.cocoeval"):
    sys.modules[name] = types.ModuleType(name)
sys.modules["pycocotools.coco"].COCO = object
sys.modules["pycocotools.cocoeval"].COCOeval = object
sys.path.insert(0, "/root/reference")
from object_detection.utils import object_detection_evaluation as ref_ode      # noqa: E402
from object_detection.utils import metrics as ref_metrics                      # noqa: E402


def scenario(seed, num_class, num_images, subsets, nms_iou, nms_max, nms_type="standard", sigma=0.5):
    rng = np.random.default_rng(seed)
    ev = ref_ode.ObjectDetectionEvaluation(num_class, matching_iou_threshold=0.5, nms_type=nms_type,
                                           nms_iou_threshold=nms_iou, nms_max_output_boxes=nms_max,
                                           soft_nms_sigma=sigma, subset_names=subsets)
    data = []
    for i in range(num_images):
        g = int(rng.integers(0, 6))
        y0, x0 = rng.uniform(0, 60, g), rng.uniform(0, 60, g)
        gb = np.stack([y0, x0, y0 + rng.uniform(5, 40, g), x0 + rng.uniform(5, 40, g)], 1)
        gc = rng.integers(0, num_class, g)
        sub = ["|".join(s for s in subsets if rng.random() < 0.7) for _ in range(g)]
        n = int(rng.integers(0, 12))
        # detections: jittered copies of ground truth + random boxes, a few degenerate, some tied scores
        src = rng.integers(0, max(g, 1), n)
        jit = rng.normal(0, 4.0, (n, 4))
        db = (gb[src] + jit) if g else np.abs(rng.normal(30, 15, (n, 4)))
        rnd = rng.random(n) < 0.3
        ry, rx = rng.uniform(0, 60, n), rng.uniform(0, 60, n)
        db[rnd] = np.stack([ry, rx, ry + rng.uniform(5, 40, n), rx + rng.uniform(5, 40, n)], 1)[rnd]
        if n:
            db[rng.random(n) < 0.05, 2] = 0.0                       # ymax < ymin: invalid
        ds = np.round(rng.random(n), 1)                              # ties on purpose
        dc = np.where(rng.random(n) < 0.8, gc[src] if g else rng.integers(0, num_class, n), rng.integers(0, num_class, n))
        data.append((gb, gc, sub, db, ds, dc))
        ev.add_single_ground_truth_image_info("img%d" % i, gb, gc, sub if i % 2 else None if subsets == ("default",) else sub)
        ev.add_single_detected_image_info("img%d" % i, db, ds, dc)
    ap, mean_ap, prec, rec, corloc, mean_corloc = ev.evaluate()
    return data, {s: ap[s] for s in subsets}, mean_ap, corloc, mean_corloc


def main():
    out = {}
    cases = [(0, 3, 12, ("default",), 1.0, 10000, "standard", 0.5), (1, 4, 20, ("default", "small"), 1.0, 10000, "standard", 0.5),
             (2, 2, 15, ("default",), 0.6, 5, "standard", 0.5), (3, 5, 30, ("a", "b", "c"), 0.3, 10000, "standard", 0.5),
             (4, 3, 25, ("default",), 0.3, 10000, "soft-linear", 0.5), (5, 3, 25, ("default",), 0.5, 10000, "soft-gaussian", 0.3),
             (6, 2, 20, ("default",), 0.4, 4, "soft-gaussian", 0.5)]
    for ci, (seed, C, N, subsets, nms_iou, nms_max, nms_type, sigma) in enumerate(cases):
        data, ap, mean_ap, corloc, mean_corloc = scenario(seed, C, N, subsets, nms_iou, nms_max, nms_type, sigma)
        out["case%d/nms_type" % ci] = np.array(nms_type)
        out["case%d/sigma" % ci] = np.array(sigma)
        out["case%d/meta" % ci] = np.array([seed, C, N, nms_max], np.int64)
        out["case%d/nms_iou" % ci] = np.array(nms_iou)
        out["case%d/subsets" % ci] = np.array(subsets)
        for i, (gb, gc, sub, db, ds, dc) in enumerate(data):
            p = "case%d/img%d/" % (ci, i)
            out[p + "gb"], out[p + "gc"], out[p + "sub"] = gb, gc, np.array(sub, dtype="U32")
            out[p + "db"], out[p + "ds"], out[p + "dc"] = db, ds, dc
        for s in subsets:
            out["case%d/ap/%s" % (ci, s)] = ap[s]
            out["case%d/map/%s" % (ci, s)] = np.array(mean_ap[s])
        out["case%d/corloc" % ci] = corloc
        out["case%d/mean_corloc" % ci] = np.array(mean_corloc)
    # curve-level vectors
    rng = np.random.default_rng(9)
    for k in range(4):
        n = int(rng.integers(1, 40))
        sc = np.round(rng.random(n), 2)
        lb = rng.random(n) < 0.5
        num_gt = int(lb.sum() + rng.integers(0, 5))
        p, r = ref_metrics.compute_precision_recall(sc, lb, num_gt)
        out["curve%d/scores" % k], out["curve%d/labels" % k], out["curve%d/num_gt" % k] = sc, lb, np.array(num_gt)
        out["curve%d/precision" % k], out["curve%d/recall" % k] = p, r
        out["curve%d/ap" % k] = np.array(ref_metrics.compute_average_precision(p, r))
    # auxiliary-task metrics (utils/mtl_util.py): skimage is absent here and only used by the edge-mask metric, which
    # is therefore not exported
    sk = types.ModuleType("skimage"); skt = types.ModuleType("skimage.transform")
    skt.resize = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("skimage.resize is stubbed"))
    sys.modules["skimage"], sys.modules["skimage.transform"] = sk, skt
    from object_detection.utils import mtl_util as ref_mtl
    rng = np.random.default_rng(21)
    K1, nimg = 6, 5
    res = {"groundtruth_boxes": [], "detection_boxes": [], "window_classes_gt": [], "window_classes_dt": [],
           "closeness_gt": [], "closeness_dt": []}
    for i in range(nimg):
        g, d, nw = int(rng.integers(1, 5)), int(rng.integers(2, 8)), int(rng.integers(2, 6))
        y0, x0 = rng.uniform(0, 0.6, g), rng.uniform(0, 0.6, g)
        gb = np.stack([y0, x0, y0 + rng.uniform(0.1, 0.4, g), x0 + rng.uniform(0.1, 0.4, g)], 1)
        y0, x0 = rng.uniform(0, 0.6, d), rng.uniform(0, 0.6, d)
        db = np.stack([y0, x0, y0 + rng.uniform(0.1, 0.4, d), x0 + rng.uniform(0.1, 0.4, d)], 1)
        wl = np.round(rng.random((nw, K1)) * (rng.random((nw, K1)) < 0.5), 3)
        wl[:, 0] = np.maximum(wl[:, 0], 0.1)                                  # at least one positive label per window
        cl = np.round(rng.random((g, K1)) * (rng.random((g, K1)) < 0.4), 3)
        res["groundtruth_boxes"].append(gb); res["detection_boxes"].append(db)
        res["window_classes_gt"].append([" ".join(str(v) for v in row) for row in wl])
        res["window_classes_dt"].append(rng.normal(0, 2, (nw, K1)))
        res["closeness_gt"].append([" ".join(str(v) for v in row) for row in cl])
        res["closeness_dt"].append(rng.normal(0, 2, (d, K1)))
        for k in ("groundtruth_boxes", "detection_boxes", "window_classes_dt", "closeness_dt"):
            out["mtl/img%d/%s" % (i, k)] = res[k][-1]
        out["mtl/img%d/window_classes_gt" % i] = np.array(res["window_classes_gt"][-1], dtype="U128")
        out["mtl/img%d/closeness_gt" % i] = np.array(res["closeness_gt"][-1], dtype="U128")
    m = ref_mtl.get_mtl_metrics(res)
    out["mtl/n"] = np.array(nimg)
    out["mtl/window_map"], out["mtl/closeness_diff"] = np.array(m["mtl/window_map"]), np.array(m["mtl/closeness_diff"])
    # metric assembly of eval_util.evaluate_detection_results_pascal_voc (label offset, `difficult`, subsets, names);
    # matplotlib / TensorFlow / the label-map protobuf are absent: stubbed, the PR-curve plot is switched off
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.backends", "matplotlib.backends.backend_agg",
                 "matplotlib.figure", "tensorflow", "global_utils", "global_utils.custom_utils",
                 "object_detection.utils.label_map_util", "object_detection.utils.visualization_utils"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib.backends.backend_agg"].FigureCanvasAgg = object
    sys.modules["matplotlib.figure"].Figure = object
    sys.modules["tensorflow"].contrib = types.SimpleNamespace(slim=None)
    quiet = types.SimpleNamespace(**{k: (lambda *a, **kw: None) for k in ("info", "infov", "warn", "warning", "error")})
    sys.modules["global_utils.custom_utils"].log = quiet
    sys.modules["object_detection.utils.label_map_util"].create_category_index = \
        lambda categories: {c["id"]: c for c in categories}
    from object_detection import eval_util as ref_eval
    ref_eval.visualize_pr_curve = lambda *a, **k: None
    rng = np.random.default_rng(33)
    C, nimg = 4, 14
    cats = [{"id": i + 1, "name": "cat%d" % (i + 1)} for i in range(C)]
    lists = {k: [] for k in ("detection_boxes", "detection_scores", "detection_classes", "image_id", "groundtruth_boxes",
                             "groundtruth_classes", "difficult", "groundtruth_subset")}
    for i in range(nimg):
        g_, n_ = int(rng.integers(1, 6)), int(rng.integers(1, 10))
        y0, x0 = rng.uniform(0, 60, g_), rng.uniform(0, 60, g_)
        gb = np.stack([y0, x0, y0 + rng.uniform(5, 40, g_), x0 + rng.uniform(5, 40, g_)], 1)
        src = rng.integers(0, g_, n_)
        db = gb[src] + rng.normal(0, 4.0, (n_, 4))
        gc = rng.integers(1, C + 1, g_)
        lists["groundtruth_boxes"].append(gb); lists["groundtruth_classes"].append(gc)
        lists["difficult"].append((rng.random(g_) < 0.25).astype(np.int64))
        lists["groundtruth_subset"].append(np.array(["|".join(s for s in ("all", "big") if s == "all" or rng.random() < 0.5)
                                                     for _ in range(g_)]))
        lists["detection_boxes"].append(db); lists["detection_scores"].append(np.round(rng.random(n_), 2))
        lists["detection_classes"].append(np.where(rng.random(n_) < 0.8, gc[src], rng.integers(1, C + 1, n_)))
        lists["image_id"].append(str(1000 + i))
    m = ref_eval.evaluate_detection_results_pascal_voc(lists, cats, corloc_summary=True)
    out["voc/n"] = np.array(nimg)
    for k in lists:
        for i, v in enumerate(lists[k]):
            out["voc/%s/%d" % (k, i)] = np.asarray(v)
    out["voc/metric_names"] = np.array(sorted(m), dtype="U128")
    out["voc/metric_values"] = np.array([m[k] for k in sorted(m)], np.float64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "eval_reference.npz"), **out)
    print("wrote eval_reference.npz with %d arrays" % len(out))


if __name__ == "__main__":
    main()
