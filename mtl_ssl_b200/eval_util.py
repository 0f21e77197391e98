"""PASCAL-VOC metric assembly over a list of per-image results: the part of
/root/reference/object_detection/eval_util.py:233-372 (`evaluate_detection_results_pascal_voc`) that is arithmetic
(subset / difficult handling, label offset, metric names); plotting and summaries are not built."""
import copy

import numpy as np

from .utils.detection_evaluation import ObjectDetectionEvaluation


def evaluate_detection_results_pascal_voc(result_lists, categories, label_id_offset=1, iou_thres=0.5,
                                          corloc_summary=False, nms_type="standard", nms_thres=1.0,
                                          soft_nms_sigma=0.5):
    """result_lists: per-image lists 'detection_boxes' [N,4], 'detection_scores' [N], 'detection_classes' [N],
    'image_id', 'groundtruth_boxes' [M,4], 'groundtruth_classes' [M] (+ optional 'difficult' [M] bool and
    'groundtruth_subset' [M] str).  categories: [{'id': int, 'name': str}].  Returns {metric name: value} with the
    reference's names ('Subset default    mAP@0.5IOU', '.../<category>', 'CorLoc/CorLoc@0.5IOU', ...)."""
    need = ["detection_boxes", "detection_scores", "detection_classes", "image_id", "groundtruth_boxes",
            "groundtruth_classes"]
    if not set(need).issubset(result_lists):
        raise ValueError("result_lists does not have expected key set.")
    n = len(result_lists[need[0]])
    if any(len(result_lists[k]) != n for k in need):
        raise ValueError("Inconsistent list sizes in result_lists")
    subset_lists = result_lists.get("groundtruth_subset") or []
    names = {s for per_image in subset_lists for joined in per_image for s in str(joined).split("|")}
    names.discard("")
    subset_names = tuple(sorted(names)) or ("default",)
    cats = copy.deepcopy(categories)
    for c in cats:
        c["id"] -= label_id_offset
    num_classes = max(c["id"] for c in cats) + 1
    ids = result_lists["image_id"]
    image_ids = [int(i) for i in ids] if all(str(i).isdigit() for i in ids) else list(range(n))
    ev = ObjectDetectionEvaluation(num_classes, matching_iou_threshold=iou_thres, nms_type=nms_type,
                                   nms_iou_threshold=nms_thres, soft_nms_sigma=soft_nms_sigma, subset_names=subset_names)
    difficult = result_lists.get("difficult") or None
    for k, image_id in enumerate(image_ids):
        gb = np.asarray(result_lists["groundtruth_boxes"][k], float).reshape(-1, 4)
        subset = None
        if subset_lists and len(subset_lists[k]) == len(gb):
            subset = np.asarray([str(s) for s in subset_lists[k]])
        if difficult is not None and np.size(difficult[k]):
            hard = np.asarray(difficult[k]).astype(bool)
            subset = np.where(hard, "", "default" if subset is None else subset)
        ev.add_single_ground_truth_image_info(image_id, gb, np.asarray(result_lists["groundtruth_classes"][k], int)
                                              - label_id_offset, subset)
        ev.add_single_detected_image_info(image_id, result_lists["detection_boxes"][k],
                                          result_lists["detection_scores"][k],
                                          np.asarray(result_lists["detection_classes"][k], int) - label_id_offset)
    ap, mean_ap, _, _, corloc, mean_corloc = ev.evaluate()
    by_id = {c["id"]: c["name"] for c in cats}
    metrics = {"Subset {:10} mAP@{}IOU".format(s, iou_thres): mean_ap[s] for s in mean_ap}
    for s in ap:
        for idx in range(ap[s].size):
            if idx in by_id:
                metrics["Subset {:10} mAP@{}IOU/{}".format(s, iou_thres, by_id[idx])] = ap[s][idx]
    if corloc_summary:
        metrics["CorLoc/CorLoc@{}IOU".format(iou_thres)] = mean_corloc
        for idx in range(corloc.size):
            if idx in by_id:
                metrics["PerformanceByCategory/CorLoc@{}IOU/{}".format(iou_thres, by_id[idx])] = corloc[idx]
    return metrics


COCO_METRIC_NAMES = ("AP", "AP_IoU50", "AP_IoU75", "AP_small", "AP_medium", "AP_large",
                     "AR_max1", "AR_max10", "AR_max100", "AR_small", "AR_medium", "AR_large")


def evaluate_detection_results_coco(result_lists, categories, label_id_offset=1, iou_thres=0.5, corloc_summary=False,
                                    nms_type="standard", nms_thres=1.0, soft_nms_sigma=0.5, eval_config=None,
                                    eval_ann_filename=None):
    """eval_util.py:393-548 of the reference: MS-COCO metrics through `CocoEvaluation` (pycocotools restated in
    utils/coco_evaluation.py).  `eval_config.coco_eval_options` selects the metric indices (0..11, default [0] = AP),
    all-categories only (eval_class_type 0) or per category as well (1), and the annotation file the ground truth is
    read from (`eval_ann_filename` overrides it; it may also be a data.mscoco.CocoIndex).
    Returns {'COCO_Eval/<All | category name>/<metric name>': value}; {} when the annotation file is missing."""
    from .utils.coco_evaluation import CocoEvaluation
    need = ["detection_boxes", "detection_scores", "detection_classes", "image_id", "groundtruth_boxes",
            "groundtruth_classes"]
    if not set(need).issubset(result_lists):
        raise ValueError("result_lists does not have expected key set.")
    n = len(result_lists[need[0]])
    if any(len(result_lists[k]) != n for k in need):
        raise ValueError("Inconsistent list sizes in result_lists")
    cats = copy.deepcopy(categories)
    for c in cats:
        c["id"] -= label_id_offset
    num_classes = max(c["id"] for c in cats) + 1
    ids = result_lists["image_id"]
    image_ids = [int(i) for i in ids] if all(str(i).isdigit() for i in ids) else list(range(n))
    ev = CocoEvaluation(num_classes, matching_iou_threshold=iou_thres, nms_type=nms_type, nms_iou_threshold=nms_thres,
                        soft_nms_sigma=soft_nms_sigma)
    for k, image_id in enumerate(image_ids):
        ev.add_single_ground_truth_image_info(image_id, result_lists["groundtruth_boxes"][k],
                                              np.asarray(result_lists["groundtruth_classes"][k], int) - label_id_offset)
        ev.add_single_detected_image_info(image_id, result_lists["detection_boxes"][k],
                                          result_lists["detection_scores"][k],
                                          np.asarray(result_lists["detection_classes"][k], int) - label_id_offset)
    metric_index, class_type, ann = [0], 1, "../data/mscoco/annotations/instances_eval2014.json"
    opts = getattr(eval_config, "coco_eval_options", None) if eval_config is not None else None
    if opts is not None:
        metric_index = list(opts.eval_metric_index) or [0]
        class_type = opts.eval_class_type
        ann = opts.eval_ann_filename or ann
    if eval_ann_filename is not None:
        ann = eval_ann_filename
    if min(metric_index) < 0 or max(metric_index) > 11:
        raise ValueError("eval_metric_index")
    if class_type < 0 or class_type > 1:
        raise ValueError("eval_class_type")
    cat_index = [0] if class_type == 0 else list(range(num_classes + 1))
    coco_metrics = ev.evaluate(cat_index, ann)
    metrics = {}
    if coco_metrics is None:
        return metrics
    by_id = {c["id"]: c["name"] for c in cats}
    for cat_id in range(num_classes + 1):
        if cat_id == 0:
            name = "All"
        elif cat_id - 1 in by_id:
            name = by_id[cat_id - 1]
        else:
            continue
        if cat_id not in cat_index or cat_id not in coco_metrics:
            continue
        for mi, mname in enumerate(COCO_METRIC_NAMES):
            if mi in metric_index:
                metrics["COCO_Eval/%s/%s" % (name, mname)] = coco_metrics[cat_id][mi]
    return metrics


def save_detection_results_for_submission(result_lists, categories, summary_dir, metrics_set):
    """eval_util.py:884-930 of the reference (`eval_config.submission_format_output`): the evaluation servers' formats.
    'coco_metrics'       -> <summary_dir>/detection_results/detection_results.json, one
                            {"image_id", "category_id", "bbox": [x, y, w, h] (1 decimal), "score" (3 decimals)} per detection
    'pascal_voc_metrics' -> <summary_dir>/detection_results/comp4_det_test_<class>.txt, lines
                            '<image> <score> <xmin> <ymin> <xmax> <ymax>' (image ids without .jpg / .png)
    Boxes are the absolute [ymin, xmin, ymax, xmax] of evaluator.run_inference.  Returns the list of files written."""
    import os
    out_dir = os.path.join(summary_dir, "detection_results")
    os.makedirs(out_dir, exist_ok=True)
    ids = result_lists["image_id"]
    dets = list(zip(ids, result_lists["detection_boxes"], result_lists["detection_scores"],
                    result_lists["detection_classes"]))
    if metrics_set == "coco_metrics":
        path = os.path.join(out_dir, "detection_results.json")
        rows = []
        for image_name, boxes, scores, classes in dets:
            for (t, l, b, r), score, class_id in zip(boxes, scores, classes):
                rows.append('{"image_id":%s,"category_id":%d,"bbox":[%.1f,%.1f,%.1f,%.1f],"score":%.3f}'
                            % (image_name, class_id, l, t, r - l, b - t, score))
        with open(path, "w") as f:
            f.write("[" + ",".join(rows) + "]")
        return [path]
    if metrics_set == "pascal_voc_metrics":
        files = {c["id"]: open(os.path.join(out_dir, "comp4_det_test_" + c["name"] + ".txt"), "w") for c in categories}
        try:
            for image_name, boxes, scores, classes in dets:
                image_name = str(image_name).replace(".jpg", "").replace(".png", "")
                for (t, l, b, r), score, class_id in zip(boxes, scores, classes):
                    files[int(class_id)].write("%s %f %f %f %f %f\n" % (image_name, score, l, t, r, b))
        finally:
            for f in files.values():
                f.close()
        return [f.name for f in files.values()]
    raise ValueError("Metric not found: {}".format(metrics_set))


def main_metric(metrics, metrics_set="pascal_voc_metrics", main_subset=""):
    """The scalar `save_best_ckpt` ranks checkpoints by (eval_util.py:932-962): for PASCAL the mAP of
    `eval_config.main_subset` (default: the subset named 'all'), for COCO 'COCO_Eval/All/AP'.  -> (key, value)."""
    if metrics_set == "coco_metrics":
        return "COCO_Eval/All/AP", float(metrics["COCO_Eval/All/AP"])
    if metrics_set != "pascal_voc_metrics":
        raise ValueError("Metric not found: {}".format(metrics_set))
    for key in metrics:
        if "/" in key:
            continue
        if (not main_subset and "Subset all" in key) or (main_subset and main_subset in key):
            return key, float(metrics[key])
    raise KeyError("no mAP entry for subset %r among %s" % (main_subset or "all", [k for k in metrics if "/" not in k]))


def save_best_checkpoint(metrics, checkpoint_file, save_fn, metrics_set="pascal_voc_metrics", main_subset=""):
    """`save_best_ckpt` (eval_util.py:932-990): keep `<dir of checkpoint_file>/best/model.ckpt` + summary.json for the
    checkpoint with the highest main metric so far (a NaN metric never replaces a stored one).  `save_fn(prefix)`
    writes the checkpoint (e.g. functools.partial(checkpoint_io.save_tf_checkpoint, model)).  Returns True if saved."""
    import json
    import math
    import os
    _, value = main_metric(metrics, metrics_set, main_subset)
    best = os.path.join(os.path.dirname(checkpoint_file), "best")
    summary_path = os.path.join(best, "summary.json")
    if os.path.exists(summary_path):
        with open(summary_path) as f:
            old = json.load(f)
        if "mAP" in old and (math.isnan(value) or value < old["mAP"]):
            return False
    os.makedirs(best, exist_ok=True)
    out = dict({k: float(v) for k, v in metrics.items()}, checkpoint_file=checkpoint_file, mAP=value)
    with open(summary_path, "w") as f:
        json.dump(out, f, indent=2, sort_keys=True)
    save_fn(os.path.join(best, "model.ckpt"))
    return True
