"""Evaluation metrics of the three auxiliary tasks (host NumPy like the reference):
/root/reference/object_detection/utils/mtl_util.py:20-108 `get_mtl_metrics`, same input / output keys.

  mtl/window_map      mean over images of the mean over windows of the AP of ranking the K+1 labels of a window by
                      the softmax of its predicted logits against "ground-truth soft label > 0" (mtl_util.py:42-57)
  mtl/closeness_diff  (an accuracy despite its name) per ground-truth box with a non-zero closeness label: does the
                      arg-max over the K object classes of the sigmoid prediction -- taken from the DETECTION with
                      the largest intersection with that box -- agree with the label's arg-max (:60-87)
  mtl/edgemask_ap     pixel accuracy of "foreground logit > background logit" after resizing the prediction to the
                      ground-truth mask size (:89-106)

Parity: window_map / closeness_diff are pinned to outputs of the reference function run in this container
(tests/golden/make_eval_golden.py).  The reference resizes the edge-mask prediction with `skimage.transform.resize`
(absent here): restated as bilinear interpolation with half-pixel centres and edge clamping; only the SIGN of
(foreground - background) enters the metric, which is what both interpolators agree on away from exact ties --
parity unpinned for that metric."""
import numpy as np

from .detection_evaluation import compute_average_precision, compute_precision_recall


def _softmax(x):
    e = np.exp(x - np.max(x, axis=-1, keepdims=True))
    return e / np.sum(e, axis=-1, keepdims=True)


def _label_row(v):
    """A label row as the decoder hands it over: the record's text ('0.0 0.5 ... ') or already numbers."""
    if isinstance(v, (bytes, str)):
        v = v.decode("utf-8") if isinstance(v, bytes) else v
        return np.asarray([float(t) for t in v.split(" ") if t != ""], np.float32)
    return np.asarray(v, np.float32)


def _bilinear_resize(img, out_h, out_w):
    """[h, w, c] -> [out_h, out_w, c], half-pixel centres, clamped at the borders."""
    h, w = img.shape[:2]
    ys = np.clip((np.arange(out_h) + 0.5) * h / out_h - 0.5, 0, h - 1)
    xs = np.clip((np.arange(out_w) + 0.5) * w / out_w - 0.5, 0, w - 1)
    y0, x0 = np.floor(ys).astype(int), np.floor(xs).astype(int)
    y1, x1 = np.minimum(y0 + 1, h - 1), np.minimum(x0 + 1, w - 1)
    fy, fx = (ys - y0)[:, None, None], (xs - x0)[None, :, None]
    top = img[y0][:, x0] * (1 - fx) + img[y0][:, x1] * fx
    bot = img[y1][:, x0] * (1 - fx) + img[y1][:, x1] * fx
    return top * (1 - fy) + bot * fy


def _pairwise_intersection(a, b):
    a, b = np.asarray(a, float).reshape(-1, 4), np.asarray(b, float).reshape(-1, 4)
    ih = np.maximum(np.minimum(a[:, None, 2], b[None, :, 2]) - np.maximum(a[:, None, 0], b[None, :, 0]), 0.0)
    iw = np.maximum(np.minimum(a[:, None, 3], b[None, :, 3]) - np.maximum(a[:, None, 1], b[None, :, 1]), 0.0)
    return ih * iw


def get_mtl_metrics(result_lists):
    """result_lists: per-image lists under 'groundtruth_boxes', 'detection_boxes' and, per task, 'window_classes_gt' /
    'window_classes_dt' ([Nw] label rows / [Nw, K+1] logits), 'closeness_gt' / 'closeness_dt' ([G] label rows /
    [num_detections, K+1] logits), 'edgemask_gt' / 'edgemask_dt' ([2, h, w] / [1?, Hf, Wf, 2] logits)."""
    out = {}
    if "window_classes_gt" in result_lists:
        per_image = []
        for gts, dts in zip(result_lists["window_classes_gt"], result_lists["window_classes_dt"]):
            aps = []
            for gt, dt in zip(gts, dts):
                positive = _label_row(gt) > 0
                p, r = compute_precision_recall(_softmax(np.asarray(dt, float)), positive, int(positive.sum()))
                aps.append(compute_average_precision(p, r))
            per_image.append(float(np.mean(aps)))
        out["mtl/window_map"] = float(np.mean(per_image))
    if "closeness_gt" in result_lists:
        per_image = []
        for gts, gboxes, dboxes, dts in zip(result_lists["closeness_gt"], result_lists["groundtruth_boxes"],
                                            result_lists["detection_boxes"], result_lists["closeness_dt"]):
            nearest = np.argmax(_pairwise_intersection(gboxes, dboxes), axis=1)
            hits = []
            for gt, j in zip(gts, nearest):
                label = _label_row(gt)
                if not np.any(label != 0):
                    continue
                pred = 1.0 / (1.0 + np.exp(-np.asarray(dts[j], float)))
                hits.append(float(np.argmax(pred[1:]) == np.argmax(label[1:])))
            if hits:
                per_image.append(float(np.mean(hits)))
        out["mtl/closeness_diff"] = float(np.mean(per_image)) if per_image else 0.0
    if "edgemask_gt" in result_lists:
        acc = []
        for gt, dt in zip(result_lists["edgemask_gt"], result_lists["edgemask_dt"]):
            fg = np.asarray(gt)[0]
            logits = np.asarray(dt, np.float32)
            logits = logits[0] if logits.ndim == 4 else logits
            up = _bilinear_resize(logits, fg.shape[0], fg.shape[1]).astype(np.float32)
            acc.append(np.mean((up[:, :, 0] < up[:, :, 1]).astype(np.float32) == fg))
        out["mtl/edgemask_ap"] = float(np.mean(acc)) if acc else 0.0
    return out
