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
