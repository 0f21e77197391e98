"""Self-supervised auxiliary labels "recycled" from the bounding boxes (host side, NumPy).

Restates the label semantics of
/root/reference/object_detection/create_records/create_pascal_tf_record.py:
  :140-161, :182-289  multi-object soft labels of 64 random windows: per class the area covered by
                      the class's boxes inside the window (inclusion-exclusion over the clipped,
                      window-normalised boxes == area of their union), label = sqrt(area), background
                      = sqrt(max(0, 1 - union of all boxes)), normalised by the sum, rounded to 3
                      decimals by the text encoding (:121-125);
  :325-358            closeness: per other-class object 1 - centre distance / image diagonal, max per
                      class, background 1 when nothing else is present, normalised by the sum;
  :375-421            64x64 foreground mask + per-box 1/area weight plane normalised to mean 1.
The reference writes these into TFRecords offline; here they are produced on the fly (synthetic batches,
data/pascal_voc.py, data/mscoco.py).  No reference TEST covers this code; it is pinned to outputs of the
reference's two record writers run under recording stubs (tests/golden/make_aux_golden.py, make_coco_aux_golden.py).
"""
import math

import numpy as np


def union_area(boxes):
    """Area of the union of axis-aligned boxes [N,4] (ymin,xmin,ymax,xmax) by coordinate compression;
    equals the inclusion-exclusion sum of get_rect_area_total (crp:140-161)."""
    boxes = np.asarray(boxes, np.float64).reshape(-1, 4)
    boxes = boxes[(boxes[:, 2] > boxes[:, 0]) & (boxes[:, 3] > boxes[:, 1])]
    if len(boxes) == 0:
        return 0.0
    ys = np.unique(np.concatenate([boxes[:, 0], boxes[:, 2]]))
    xs = np.unique(np.concatenate([boxes[:, 1], boxes[:, 3]]))
    cy = (ys[:-1] + ys[1:]) / 2
    cx = (xs[:-1] + xs[1:]) / 2
    inside = ((cy[None, :, None] > boxes[:, 0, None, None]) & (cy[None, :, None] < boxes[:, 2, None, None]) &
              (cx[None, None, :] > boxes[:, 1, None, None]) & (cx[None, None, :] < boxes[:, 3, None, None])).any(0)
    return float((inside * np.diff(ys)[:, None] * np.diff(xs)[None, :]).sum())


def _clip_and_normalise(boxes, window):
    """np_box_list_ops.clip_to_window + change_coordinate_frame (utils/np_box_list_ops.py:467, :639)."""
    b = np.asarray(boxes, np.float64).reshape(-1, 4).copy()
    b[:, 0] = np.clip(b[:, 0], window[0], window[2]); b[:, 2] = np.clip(b[:, 2], window[0], window[2])
    b[:, 1] = np.clip(b[:, 1], window[1], window[3]); b[:, 3] = np.clip(b[:, 3], window[1], window[3])
    b = b[(b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) > 0]
    wh, ww = window[2] - window[0], window[3] - window[1]
    return np.stack([(b[:, 0] - window[0]) / wh, (b[:, 1] - window[1]) / ww, (b[:, 2] - window[0]) / wh,
                     (b[:, 3] - window[1]) / ww], 1) if len(b) else b


def _round3(v):
    return np.round(np.asarray(v, np.float64), 3).astype(np.float32)      # get_string_label round trip


def window_label(boxes_abs, classes, window, num_classes):
    """get_multi_label (crp:197-222): [K+1] soft label of one window (label_option 1, normalize 1)."""
    lab = np.zeros(num_classes + 1, np.float64)
    nb = _clip_and_normalise(boxes_abs, window)
    lab[0] = math.sqrt(max(0.0, 1.0 - union_area(nb)))
    bg = lab[0]
    for c in np.unique(classes):
        a = union_area(_clip_and_normalise(np.asarray(boxes_abs)[np.asarray(classes) == c], window))
        lab[int(c)] = math.sqrt(max(a, 0.0))
    return _round3(lab / lab.sum()), bg


def random_windows(boxes_abs, classes, height, width, num_classes, rng, num_windows=64, min_obj_size=32.0):
    """crp:225-260 (`random_multi_object`): windows that see at least one object."""
    out_b, out_l = [], []
    tries = 0
    while len(out_b) < num_windows:
        tries += 1
        bh = rng.random() * (height - min_obj_size) + min_obj_size
        bw = rng.random() * (width - min_obj_size) + min_obj_size
        cy, cx = rng.random() * height, rng.random() * width
        ymin, xmin = max(0.0, cy - bh / 2), max(0.0, cx - bw / 2)
        ymax, xmax = min(height, cy + bh / 2), min(width, cx + bw / 2)
        if xmax - xmin < min_obj_size:
            if xmin == 0.0:
                xmax = min_obj_size
            elif xmax == width:
                xmin = width - min_obj_size
        if ymax - ymin < min_obj_size:
            if ymin == 0.0:
                ymax = min_obj_size
            elif ymax == height:
                ymin = height - min_obj_size
        window = [ymin, xmin, ymax, xmax]
        lab, bg = window_label(boxes_abs, classes, window, num_classes)
        if len(boxes_abs) and bg == 1.0 and tries < 100000:
            continue
        out_b.append([ymin / height, xmin / width, ymax / height, xmax / width])
        out_l.append(lab)
    return np.asarray(out_b, np.float32), np.asarray(out_l, np.float32)


def closeness_labels(boxes_abs, classes, height, width, num_classes):
    """get_closeness (crp:325-358) for every object -> [G, K+1]."""
    boxes_abs = np.asarray(boxes_abs, np.float64).reshape(-1, 4)
    G = len(boxes_abs)
    out = np.zeros((G, num_classes + 1), np.float64)
    diag = math.sqrt(width * width + height * height)
    cy = (boxes_abs[:, 0] + boxes_abs[:, 2]) / 2
    cx = (boxes_abs[:, 1] + boxes_abs[:, 3]) / 2
    for i in range(G):
        if G == 1:
            out[i, 0] = 1
            continue
        for j in range(G):
            if i == j or classes[i] == classes[j]:
                continue
            d = math.sqrt((cx[i] - cx[j]) ** 2 + (cy[i] - cy[j]) ** 2) / diag
            out[i, int(classes[j])] = max(out[i, int(classes[j])], 1.0 - d)
        if out[i, 1:].sum() == 0:
            out[i, 0] = 1
        out[i] /= out[i].sum()
    return _round3(out)


def edgemask(boxes_abs, height, width, mask_size=64):
    """create_edgemask (crp:375-421) -> float32 [2, 64, 64] (foreground mask, weight plane)."""
    m = np.zeros([mask_size, mask_size], np.float32)
    w = np.ones([mask_size, mask_size], np.float32) / mask_size / mask_size
    for b in np.asarray(boxes_abs, np.float64).reshape(-1, 4):
        ymin = int(b[0] / height * mask_size)
        xmin = int(b[1] / width * mask_size)
        ymax = min(mask_size - 1, int(b[2] / height * mask_size + 0.99))
        xmax = min(mask_size - 1, int(b[3] / width * mask_size + 0.99))
        bw, bh = xmax - xmin + 1, ymax - ymin + 1
        if bw == 0:
            if xmin + xmax > mask_size:
                xmin -= 1
            else:
                xmax += 1
            bw = 1
        if bh == 0:
            if ymin + ymax > mask_size:
                ymin -= 1
            else:
                ymax += 1
            bh = 1
        m[ymin:ymax + 1, xmin:xmax + 1] = 1.0
        wt = np.ones([bh, bw], np.float32) / bw / bh
        w[ymin:ymax + 1, xmin:xmax + 1] = np.maximum(wt, w[ymin:ymax + 1, xmin:xmax + 1])
    w /= np.mean(w)
    return np.array([m, w])
