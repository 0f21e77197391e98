"""Oracle (test infrastructure): RPN post-processing — decode, softmax, clip, greedy NMS.

Paths relative to /root/reference/object_detection/.  `tf.image.non_max_suppression` is a
TensorFlow 1.7.0 kernel (core/kernels/non_max_suppression_op.cc) that is not vendored in
the reference; its published algorithm is restated in `tf_non_max_suppression`.
"""
import numpy as np

from . import boxes as B

F = np.float32


def tf_nms_iou(boxes, i, j):
    """TF 1.7 non_max_suppression_op.cc `ComputeIOU` (float32, corner order normalised)."""
    bi, bj = boxes[i], boxes[j]
    ymin_i, xmin_i = min(bi[0], bi[2]), min(bi[1], bi[3])
    ymax_i, xmax_i = max(bi[0], bi[2]), max(bi[1], bi[3])
    ymin_j, xmin_j = min(bj[0], bj[2]), min(bj[1], bj[3])
    ymax_j, xmax_j = max(bj[0], bj[2]), max(bj[1], bj[3])
    area_i = F(F(ymax_i - ymin_i) * F(xmax_i - xmin_i))
    area_j = F(F(ymax_j - ymin_j) * F(xmax_j - xmin_j))
    if area_i <= 0 or area_j <= 0:
        return F(0)
    iy0, ix0 = max(ymin_i, ymin_j), max(xmin_i, xmin_j)
    iy1, ix1 = min(ymax_i, ymax_j), min(xmax_i, xmax_j)
    inter = F(max(F(iy1 - iy0), F(0)) * max(F(ix1 - ix0), F(0)))
    return F(inter / F(F(area_i + area_j) - inter))


def tf_non_max_suppression(boxes, scores, max_output_size, iou_threshold):
    """Greedy NMS as in TF 1.7: visit boxes by descending score; keep a box iff its IoU with
    every already-kept box is <= iou_threshold; stop at max_output_size.  TF sorts with
    std::sort (tie order unspecified); the restatement fixes ties to lower-index-first.
    Returns selected indices (int32) into `boxes`."""
    boxes = np.asarray(boxes, F).reshape(-1, 4)
    scores = np.asarray(scores, F)
    order = np.argsort(-scores, kind="stable")
    selected = []
    for idx in order:
        if len(selected) >= max_output_size:
            break
        keep = True
        for s in reversed(selected):
            if tf_nms_iou(boxes, idx, s) > F(iou_threshold):
                keep = False
                break
        if keep:
            selected.append(int(idx))
    return np.asarray(selected, np.int32)


def nms_vectorized(boxes, scores, max_output_size, iou_threshold):
    """Same result as tf_non_max_suppression, vectorised over the kept set (for large N)."""
    boxes = np.asarray(boxes, F).reshape(-1, 4)
    scores = np.asarray(scores, F)
    ymin = np.minimum(boxes[:, 0], boxes[:, 2]); ymax = np.maximum(boxes[:, 0], boxes[:, 2])
    xmin = np.minimum(boxes[:, 1], boxes[:, 3]); xmax = np.maximum(boxes[:, 1], boxes[:, 3])
    areas = ((ymax - ymin) * (xmax - xmin)).astype(F)
    order = np.argsort(-scores, kind="stable")
    sel = []
    for idx in order:
        if len(sel) >= max_output_size:
            break
        if sel:
            s = np.asarray(sel)
            ih = np.maximum((np.minimum(ymax[idx], ymax[s]) - np.maximum(ymin[idx], ymin[s])).astype(F), F(0))
            iw = np.maximum((np.minimum(xmax[idx], xmax[s]) - np.maximum(xmin[idx], xmin[s])).astype(F), F(0))
            inter = (ih * iw).astype(F)
            with np.errstate(divide="ignore", invalid="ignore"):
                iou = (inter / ((areas[idx] + areas[s]).astype(F) - inter).astype(F)).astype(F)
            iou = np.where((areas[idx] <= 0) | (areas[s] <= 0), F(0), iou)
            if np.any(iou > F(iou_threshold)):
                continue
        sel.append(int(idx))
    return np.asarray(sel, np.int32)


def softmax(x, axis=-1):
    x = np.asarray(x, F)
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m).astype(F)
    return (e / e.sum(axis=axis, keepdims=True)).astype(F)


def rpn_postprocess_single(box_encodings, objectness_logits, anchors, image_hw, score_thresh=0.0,
                           iou_thresh=0.7, max_proposals=300, decoded=None, scores=None):
    """meta_architectures/faster_rcnn_meta_arch.py:1055-1115 for one image, with the single-class
    body of core/post_processing.py:25-164: decode -> objectness softmax[:,1] ->
    filter score > thresh (blo:652-687) -> clip_to_window + drop zero area (blo:102-137) ->
    NMS (pp:144-149) -> sort by score (already sorted) -> zero-pad to max_proposals (pp:281-296).
    Returns (boxes [max,4] abs px, scores [max], num valid)."""
    anchors = np.asarray(anchors, F)
    if decoded is None:
        decoded = B.box_decode(box_encodings, anchors)
    if scores is None:
        scores = softmax(objectness_logits)[:, 1]
    keep = np.nonzero(scores > F(score_thresh))[0]
    boxes_f, scores_f = decoded[keep], scores[keep]
    window = (0.0, 0.0, float(image_hw[0]), float(image_hw[1]))
    boxes_c, kidx = B.clip_to_window(boxes_f, window)
    scores_c = scores_f[kidx]
    sel = nms_vectorized(boxes_c, scores_c, min(max_proposals, len(boxes_c)), iou_thresh)
    out_b = np.zeros([max_proposals, 4], F)
    out_s = np.zeros([max_proposals], F)
    out_b[:len(sel)] = boxes_c[sel]
    out_s[:len(sel)] = scores_c[sel]
    return out_b, out_s, len(sel)


def multiclass_non_max_suppression(boxes, scores, score_thresh, iou_thresh, max_size_per_class,
                                   max_total_size=0, clip_window=None, change_coordinate_frame=False):
    """core/post_processing.py:25-164.  boxes [N,q,4] (q = 1 or num_classes), scores [N,C].
    Returns (boxes [k,4], scores [k], classes [k]) sorted by descending score."""
    boxes = np.asarray(boxes, F)
    scores = np.asarray(scores, F)
    n, q = boxes.shape[0], boxes.shape[1]
    num_classes = scores.shape[1]
    out_b, out_s, out_c = [], [], []
    for c in range(num_classes):
        b = boxes[:, c if q > 1 else 0]
        s = scores[:, c]
        keep = np.nonzero(s > F(score_thresh))[0]                    # filter_greater_than (blo:652-687)
        b, s = b[keep], s[keep]
        if clip_window is not None:
            b, kidx = B.clip_to_window(b, clip_window)
            s = s[kidx]
            if change_coordinate_frame:
                wy0, wx0, wy1, wx1 = [F(v) for v in clip_window]
                hh, ww = wy1 - wy0, wx1 - wx0
                b = np.stack([(b[:, 0] - wy0) / hh, (b[:, 1] - wx0) / ww, (b[:, 2] - wy0) / hh,
                              (b[:, 3] - wx0) / ww], 1).astype(F)
        sel = nms_vectorized(b, s, min(max_size_per_class, len(b)), iou_thresh)
        out_b.append(b[sel]); out_s.append(s[sel]); out_c.append(np.full(len(sel), c, F))
    ob = np.concatenate(out_b) if out_b else np.zeros((0, 4), F)
    os_ = np.concatenate(out_s) if out_s else np.zeros((0,), F)
    oc = np.concatenate(out_c) if out_c else np.zeros((0,), F)
    order = np.argsort(-os_, kind="stable")                          # sort_by_field (blo:554-594)
    if max_total_size:
        order = order[:max_total_size]
    return ob[order], os_[order], oc[order]


def second_stage_postprocess(refined_box_encodings, class_logits, proposal_boxes, num_proposals, image_hw,
                             score_thresh, iou_thresh, max_per_class, max_total, score_mode="softmax",
                             decoded=None, scores=None):
    """meta_architectures/faster_rcnn_meta_arch.py:1387-1469 (`_postprocess_box_classifier`) +
    core/post_processing.py:167-312 (`batch_multiclass_non_max_suppression` with clip_window = image,
    change_coordinate_frame=True, num_valid_boxes = num_proposals): per image decode the [P,K,4] refined encodings
    against the proposals (`_batch_decode_boxes`, fmA:1471-1497), convert the logits, drop the background column,
    run the per-class NMS and zero-pad to `max_total`.
    `decoded` [B,P,K,4] / `scores` [B,P,K]: use these instead of recomputing them (index-level device parity)."""
    enc = np.asarray(refined_box_encodings, F)
    logits = np.asarray(class_logits, F)
    props = np.asarray(proposal_boxes, F)
    Bn, P = props.shape[0], props.shape[1]
    K = enc.shape[1]
    enc = enc.reshape(Bn, P, K, 4)
    logits = logits.reshape(Bn, P, K + 1)
    if decoded is None:
        decoded = np.zeros((Bn, P, K, 4), F)
        for b in range(Bn):
            for k in range(K):
                decoded[b, :, k] = B.box_decode(enc[b, :, k], props[b])
    if scores is None:
        if score_mode == "softmax":
            conv = softmax(logits, -1)
        elif score_mode == "sigmoid":
            conv = (F(1) / (F(1) + np.exp(-logits))).astype(F)
        else:
            conv = logits
        scores = conv[:, :, 1:]
    H, W = image_hw
    out_b = np.zeros((Bn, max_total, 4), F)
    out_s = np.zeros((Bn, max_total), F)
    out_c = np.zeros((Bn, max_total), F)
    out_n = np.zeros((Bn,), F)
    for b in range(Bn):
        n = int(num_proposals[b])
        bb, ss, cc = multiclass_non_max_suppression(decoded[b, :n], scores[b, :n], score_thresh, iou_thresh,
                                                    max_per_class, max_total, clip_window=(0, 0, H, W),
                                                    change_coordinate_frame=True)
        m = len(ss)
        out_b[b, :m], out_s[b, :m], out_c[b, :m], out_n[b] = bb, ss, cc, m
    return out_b, out_s, out_c, out_n
