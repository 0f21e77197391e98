"""PASCAL-VOC style detection metrics (host side like the reference): average precision per class,
mAP, CorLoc, with the fork's "subset" bookkeeping.

Restates, in NumPy and in its own structure, the behaviour of
  /root/reference/object_detection/utils/metrics.py:22-130            precision / recall, AP (VOC devkit area), CorLoc
  /root/reference/object_detection/utils/per_image_evaluation.py:28-281  per-image true / false positive labelling
  /root/reference/object_detection/utils/object_detection_evaluation.py:42-292  accumulation over a dataset
keeping the class and method names a caller of the reference uses (`ObjectDetectionEvaluation
.add_single_ground_truth_image_info / .add_single_detected_image_info / .evaluate`).  Semantics worth spelling out:
  * detections are ranked with `np.argsort(scores)[::-1]` (ties: LATER index first), per image before matching and
    over the whole dataset before the precision / recall curve;
  * a detection matches the ground-truth box of its class with the largest IoU if IoU >= threshold; the first
    detection to claim a box is the true positive, later ones are false positives; detections matched to a
    "difficult" box (one that is not in the evaluated subset) are dropped from the statistics altogether;
  * difficult boxes do not count as ground-truth instances for recall but do count for CorLoc;
  * boxes with ymin >= ymax or xmin >= xmax are discarded; per-class NMS before matching is the reference's
    `np_box_list_ops.non_max_suppression` (IoU threshold 1.0 by default = "keep the best `max_output` boxes").
nms_type 'soft-linear' / 'soft-gaussian' (the fork's addition) decay the scores of overlapping boxes instead."""
import logging

import numpy as np


# ----------------------------------------------------------------------------- curves
def compute_precision_recall(scores, labels, num_gt):
    """metrics.py:22-73.  scores float [N], labels bool [N] (true positive flags), num_gt: instances of the class.
    Returns (precision [N], recall [N]) in ranking order, or (None, None) when num_gt == 0."""
    scores, labels = np.asarray(scores), np.asarray(labels)
    if labels.ndim != 1 or labels.dtype != np.bool_:
        raise ValueError("labels must be single dimension bool numpy array")
    if scores.ndim != 1:
        raise ValueError("scores must be single dimension numpy array")
    if num_gt < labels.sum():
        raise ValueError("Number of true positives must be smaller than num_gt.")
    if len(scores) != len(labels):
        raise ValueError("scores and labels must be of the same size.")
    if num_gt == 0:
        return None, None
    rank = np.argsort(scores)[::-1]
    tp = np.cumsum(labels[rank].astype(int))
    fp = np.cumsum(1 - labels[rank].astype(int))
    return tp.astype(float) / (tp + fp), tp.astype(float) / num_gt


def compute_average_precision(precision, recall):
    """metrics.py:76-130: area under the monotone envelope of the precision / recall curve (VOC devkit, 2010+)."""
    if precision is None:
        if recall is not None:
            raise ValueError("If precision is None, recall must also be None")
        return np.nan
    precision, recall = np.asarray(precision, float), np.asarray(recall, float)
    if len(precision) != len(recall):
        raise ValueError("precision and recall must be of the same size.")
    if precision.size == 0:
        return 0.0
    if precision.min() < 0 or precision.max() > 1:
        raise ValueError("Precision must be in the range of [0, 1].")
    if recall.min() < 0 or recall.max() > 1:
        raise ValueError("recall must be in the range of [0, 1].")
    if np.any(np.diff(recall) < 0):
        raise ValueError("recall must be a non-decreasing array")
    r = np.concatenate([[0.0], recall, [1.0]])
    p = np.concatenate([[0.0], precision, [0.0]])
    p = np.maximum.accumulate(p[::-1])[::-1]              # envelope: precision never rises with recall
    step = np.nonzero(r[1:] != r[:-1])[0] + 1
    return float(np.sum((r[step] - r[step - 1]) * p[step]))


def compute_cor_loc(num_gt_imgs_per_class, num_images_correctly_detected_per_class):
    """metrics.py:133-153: fraction of the images containing a class in which it was localised; NaN without images."""
    n = np.asarray(num_gt_imgs_per_class, float)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(n == 0, np.nan, np.asarray(num_images_correctly_detected_per_class, float) / n)


# ----------------------------------------------------------------------------- geometry
def _iou_matrix(a, b):
    a, b = np.asarray(a, float).reshape(-1, 4), np.asarray(b, float).reshape(-1, 4)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    ih = np.maximum(np.minimum(a[:, None, 2], b[None, :, 2]) - np.maximum(a[:, None, 0], b[None, :, 0]), 0.0)
    iw = np.maximum(np.minimum(a[:, None, 3], b[None, :, 3]) - np.maximum(a[:, None, 1], b[None, :, 1]), 0.0)
    inter = ih * iw
    return inter / (area_a[:, None] + area_b[None, :] - inter)


def _standard_nms(boxes, scores, max_output, iou_threshold, score_threshold=-10.0):
    """np_box_list_ops.non_max_suppression:185-257 -> indices (into boxes) of the kept detections, ranked."""
    keep = np.nonzero(scores > score_threshold)[0]
    if keep.size == 0:
        return keep
    ranked = keep[np.argsort(scores[keep])[::-1]]
    if iou_threshold == 1.0:
        return ranked[:max_output]
    alive = np.ones(len(ranked), bool)
    chosen = []
    for i in range(len(ranked)):
        if len(chosen) >= max_output:
            break
        if not alive[i]:
            continue
        chosen.append(ranked[i])
        alive[i] = False
        rest = np.nonzero(alive)[0]
        if rest.size == 0:
            break
        iou = _iou_matrix(boxes[ranked[i]][None], boxes[ranked[rest]])[0]
        alive[rest] = iou <= iou_threshold
    return np.asarray(chosen, int)


def _soft_nms(boxes, scores, max_output, iou_threshold, kind, sigma, score_threshold=-10.0):
    """np_box_list_ops.soft_non_max_suppression:258-363 (Bodla et al. 2017) -> (kept indices ranked, their DECAYED
    scores).  kind 2: linear (score *= 1 - IoU where IoU >= threshold), kind 3: gaussian (score *= exp(-IoU^2 / sigma)).
    The best remaining box is taken repeatedly (first index among equal scores, after the initial descending sort) and
    decays every box still in the pool; finally boxes with score > max(0, score_threshold) are re-ranked."""
    keep = np.nonzero(scores > score_threshold)[0]
    if keep.size == 0:
        return keep, scores[keep]
    ranked = keep[np.argsort(scores[keep])[::-1]]
    if iou_threshold == 1.0:
        ranked = ranked[:max_output]
        return ranked, scores[ranked]
    b = boxes[ranked]
    sc = np.array(scores[ranked], dtype=scores.dtype)
    pool = np.ones(len(ranked), bool)
    taken = 0
    for _ in range(len(ranked)):
        if taken >= max_output:
            break
        cand = np.where(pool, sc, -np.inf)
        j = int(np.argmax(cand))                       # first maximum, like the reference's strict `>` scan
        if not (pool[j] and sc[j] > score_threshold):
            break
        taken += 1
        pool[j] = False
        rest = np.nonzero(pool)[0]
        if rest.size == 0:
            break
        iou = _iou_matrix(b[j][None], b[rest])[0]
        if kind == 2:
            w = 1.0 - np.where(iou < iou_threshold, 0.0, iou)
        else:
            w = np.exp(-np.square(iou) / sigma)
        sc[rest] = sc[rest] * w
    alive = np.nonzero(sc > max(0.0, score_threshold))[0]
    order = alive[np.argsort(sc[alive])[::-1]][:max_output]
    return ranked[order], sc[order]


# ----------------------------------------------------------------------------- one image
class PerImageEvaluation(object):
    def __init__(self, num_groundtruth_classes, matching_iou_threshold=0.5, nms_type="standard", nms_iou_threshold=1.0,
                 nms_max_output_boxes=100, soft_nms_sigma=0.5):
        if nms_type not in ("standard", "soft-linear", "soft-gaussian"):
            raise ValueError("Cannot identify NMS type.")
        self.nms_type, self.soft_nms_sigma = nms_type, soft_nms_sigma
        if nms_iou_threshold < 0.0 or nms_iou_threshold > 1.0:
            raise ValueError("IOU threshold must be in [0, 1]")
        self.num_groundtruth_classes = num_groundtruth_classes
        self.matching_iou_threshold = matching_iou_threshold
        self.nms_iou_threshold = nms_iou_threshold
        self.nms_max_output_boxes = nms_max_output_boxes

    def compute_object_detection_metrics(self, detected_boxes, detected_scores, detected_class_labels,
                                         groundtruth_boxes, groundtruth_class_labels, groundtruth_is_difficult_lists):
        """-> (scores per class, tp/fp labels per class, is_class_correctly_detected_in_image [C] int)."""
        db = np.asarray(detected_boxes, float).reshape(-1, 4)
        ds, dc = np.asarray(detected_scores, float), np.asarray(detected_class_labels)
        ok = (db[:, 0] < db[:, 2]) & (db[:, 1] < db[:, 3])          # _remove_invalid_boxes
        db, ds, dc = db[ok], ds[ok], dc[ok]
        gb = np.asarray(groundtruth_boxes, float).reshape(-1, 4)
        gc = np.asarray(groundtruth_class_labels)
        hard = np.asarray(groundtruth_is_difficult_lists, bool)
        scores, labels = [], []
        corloc = np.zeros(self.num_groundtruth_classes, int)
        for c in range(self.num_groundtruth_classes):
            d_sel, g_sel = dc == c, gc == c
            s, l = self._label_class(db[d_sel], ds[d_sel], gb[g_sel], hard[g_sel])
            scores.append(s)
            labels.append(l)
            if d_sel.any() and g_sel.any():                         # CorLoc: the best-scoring detection hits any box
                best = db[d_sel][np.argmax(ds[d_sel])]
                corloc[c] = int(_iou_matrix(best[None], gb[g_sel]).max() >= self.matching_iou_threshold)
        return scores, labels, corloc

    def _label_class(self, boxes, scores, gt_boxes, gt_hard):
        if boxes.size == 0:
            return np.array([], float), np.array([], bool)
        if self.nms_type == "standard":
            keep = _standard_nms(boxes, scores, self.nms_max_output_boxes, self.nms_iou_threshold)
            boxes, scores = boxes[keep], scores[keep]
        else:
            keep, scores = _soft_nms(boxes, scores, self.nms_max_output_boxes, self.nms_iou_threshold,
                                     2 if self.nms_type == "soft-linear" else 3, self.soft_nms_sigma)
            boxes = boxes[keep]
        if gt_boxes.size == 0:
            return scores, np.zeros(len(scores), bool)
        iou = _iou_matrix(boxes, gt_boxes)
        target = iou.argmax(1)
        hit = iou[np.arange(len(boxes)), target] >= self.matching_iou_threshold
        tp = np.zeros(len(boxes), bool)
        on_hard = hit & gt_hard[target]
        taken = np.zeros(len(gt_boxes), bool)
        for i in np.nonzero(hit & ~on_hard)[0]:                      # ranking order: first claim wins
            if not taken[target[i]]:
                tp[i] = taken[target[i]] = True
        return scores[~on_hard], tp[~on_hard]


# ----------------------------------------------------------------------------- a dataset
class ObjectDetectionEvaluation(object):
    def __init__(self, num_groundtruth_classes, matching_iou_threshold=0.5, nms_type="standard", nms_iou_threshold=1.0,
                 nms_max_output_boxes=10000, soft_nms_sigma=0.5, subset_names=("default",)):
        self.per_image_eval = PerImageEvaluation(num_groundtruth_classes, matching_iou_threshold, nms_type,
                                                 nms_iou_threshold, nms_max_output_boxes, soft_nms_sigma)
        self.num_class = num_groundtruth_classes
        self.subset_names = tuple(subset_names)
        self.clear_groundtruths()
        self.clear_detections()

    def clear_groundtruths(self):
        self.groundtruth_boxes, self.groundtruth_class_labels = {}, {}
        self.groundtruth_subset = {s: {} for s in self.subset_names}
        self.num_gt_instances_per_class = {s: np.zeros(self.num_class, int) for s in self.subset_names}
        self.num_gt_imgs_per_class = np.zeros(self.num_class, int)

    def clear_detections(self):
        self.detection_keys = set()
        self.scores_per_class = {s: [[] for _ in range(self.num_class)] for s in self.subset_names}
        self.tp_fp_labels_per_class = {s: [[] for _ in range(self.num_class)] for s in self.subset_names}
        self.num_images_correctly_detected_per_class = np.zeros(self.num_class)
        self.average_precision_per_class = {s: np.full(self.num_class, np.nan) for s in self.subset_names}
        self.precisions_per_class = {s: [] for s in self.subset_names}
        self.recalls_per_class = {s: [] for s in self.subset_names}
        self.corloc_per_class = np.ones(self.num_class, float)

    def add_single_ground_truth_image_info(self, image_key, groundtruth_boxes, groundtruth_class_labels,
                                           groundtruth_subset=None):
        """groundtruth_subset: one string per box, subset names joined by '|' ('' = in no subset, i.e. difficult
        everywhere); None = every box in 'default'."""
        if image_key in self.groundtruth_boxes:
            logging.warning("image %s has already been added to the ground truth database.", image_key)
            return
        boxes = np.asarray(groundtruth_boxes, float).reshape(-1, 4)
        labels = np.asarray(groundtruth_class_labels, int)
        self.groundtruth_boxes[image_key], self.groundtruth_class_labels[image_key] = boxes, labels
        if groundtruth_subset is None:
            groundtruth_subset = ["default"] * len(boxes)
        member = {s: np.zeros(len(boxes), bool) for s in self.subset_names}
        for i, names in enumerate(groundtruth_subset):
            for name in names.split("|"):
                if name == "":
                    continue
                if name not in member:
                    raise ValueError("%s is not found in subset_names" % name)
                member[name][i] = True
        for s in self.subset_names:
            self.groundtruth_subset[s][image_key] = member[s]
            self.num_gt_instances_per_class[s] += np.bincount(labels[member[s]], minlength=self.num_class)[:self.num_class]
        self.num_gt_imgs_per_class[np.unique(labels[(labels >= 0) & (labels < self.num_class)])] += 1

    def add_single_detected_image_info(self, image_key, detected_boxes, detected_scores, detected_class_labels):
        if len(detected_boxes) != len(detected_scores) or len(detected_boxes) != len(detected_class_labels):
            raise ValueError("detected_boxes, detected_scores and detected_class_labels should all have same lengths. "
                             "Got[%d, %d, %d]" % (len(detected_boxes), len(detected_scores), len(detected_class_labels)))
        if image_key in self.detection_keys:
            logging.warning("image %s has already been added to the detection result database", image_key)
            return
        self.detection_keys.add(image_key)
        known = image_key in self.groundtruth_boxes
        gb = self.groundtruth_boxes[image_key] if known else np.empty((0, 4), float)
        gc = self.groundtruth_class_labels[image_key] if known else np.array([], int)
        corloc = 0
        for s in self.subset_names:
            member = self.groundtruth_subset[s][image_key] if known else np.array([], bool)
            scores, labels, corloc = self.per_image_eval.compute_object_detection_metrics(
                detected_boxes, detected_scores, detected_class_labels, gb, gc, ~member)
            for c in range(self.num_class):
                self.scores_per_class[s][c].append(scores[c])
                self.tp_fp_labels_per_class[s][c].append(labels[c])
        self.num_images_correctly_detected_per_class += corloc

    def evaluate(self):
        """-> (average_precision_per_class {subset: [C]}, mean_ap {subset: float}, precisions_per_class,
        recalls_per_class, corloc_per_class [C], mean_corloc)."""
        mean_ap = {}
        for s in self.subset_names:
            for c in range(self.num_class):
                n = self.num_gt_instances_per_class[s][c]
                if n == 0:
                    continue
                sc = np.concatenate(self.scores_per_class[s][c]) if self.scores_per_class[s][c] else np.array([], float)
                lb = (np.concatenate(self.tp_fp_labels_per_class[s][c]) if self.tp_fp_labels_per_class[s][c]
                      else np.array([], bool))
                p, r = compute_precision_recall(sc, lb.astype(bool), n)
                self.precisions_per_class[s].append(p)
                self.recalls_per_class[s].append(r)
                self.average_precision_per_class[s][c] = compute_average_precision(p, r)
            with np.errstate(invalid="ignore"):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore", RuntimeWarning)
                    mean_ap[s] = np.nanmean(self.average_precision_per_class[s])
        self.corloc_per_class = compute_cor_loc(self.num_gt_imgs_per_class, self.num_images_correctly_detected_per_class)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            mean_corloc = np.nanmean(self.corloc_per_class)
        return (self.average_precision_per_class, mean_ap, self.precisions_per_class, self.recalls_per_class,
                self.corloc_per_class, mean_corloc)
