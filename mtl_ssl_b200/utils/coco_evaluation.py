"""MS-COCO bounding-box metrics (AP@[.5:.95], AP50, AP75, AP by size, AR by max detections / size) without pycocotools.

The reference computes them with `pycocotools.cocoeval.COCOeval` (third-party, pycocotools 2.0 in its requirements; NOT
under /root/reference and not installed here) from utils/object_detection_evaluation.py:294-427 (`CocoEvaluation`) and
eval_util.py:393-548 (`evaluate_detection_results_coco`).  `CocoBBoxEval` restates COCOeval's published algorithm for
iouType 'bbox' (evaluate -> accumulate -> summarize); `CocoEvaluation` mirrors the reference wrapper (per-class NMS of the
detections, [x, y, w, h] conversion, the 100 best per image, ground truth read from the annotation file).
Parity UNPINNED against pycocotools itself (absent): tests/test_detection_evaluation.py holds hand-computed cases of
every branch (greedy matching at ten IoU thresholds, crowd regions, area ranges, max-detection cuts, the 101-point
interpolation)."""
import logging
import os

import numpy as np

from . import detection_evaluation as DE

METRIC_NAMES = ("AP", "AP_IoU50", "AP_IoU75", "AP_small", "AP_medium", "AP_large",
                "AR_max1", "AR_max10", "AR_max100", "AR_small", "AR_medium", "AR_large")


def bbox_iou(dt, gt, iscrowd):
    """maskApi bbIou: boxes [x, y, w, h]; for a crowd ground truth the union is the detection's own area."""
    dt = np.asarray(dt, np.float64).reshape(-1, 4)
    gt = np.asarray(gt, np.float64).reshape(-1, 4)
    out = np.zeros((len(dt), len(gt)), np.float64)
    if not len(dt) or not len(gt):
        return out
    da, ga = dt[:, 2] * dt[:, 3], gt[:, 2] * gt[:, 3]
    w = np.minimum(dt[:, None, 0] + dt[:, None, 2], gt[None, :, 0] + gt[None, :, 2]) - np.maximum(dt[:, None, 0], gt[None, :, 0])
    h = np.minimum(dt[:, None, 1] + dt[:, None, 3], gt[None, :, 1] + gt[None, :, 3]) - np.maximum(dt[:, None, 1], gt[None, :, 1])
    inter = np.where((w > 0) & (h > 0), w * h, 0.0)
    union = np.where(np.asarray(iscrowd, bool)[None, :], da[:, None], da[:, None] + ga[None, :] - inter)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(inter > 0, inter / union, 0.0)
    return out


class CocoBBoxEval(object):
    """COCOeval(cocoGt, cocoDt, 'bbox').  gts: dicts with image_id, category_id, bbox, area, iscrowd, id;
    dts: dicts with image_id, category_id, bbox, score (area = w*h and ids are assigned like COCO.loadRes)."""

    def __init__(self, gts, dts, img_ids=None, cat_ids=None):
        self.iou_thrs = np.linspace(.5, 0.95, int(np.round((0.95 - .5) / .05)) + 1, endpoint=True)
        self.rec_thrs = np.linspace(.0, 1.00, int(np.round((1.00 - .0) / .01)) + 1, endpoint=True)
        self.max_dets = [1, 10, 100]
        self.area_rng = [[0 ** 2, 1e5 ** 2], [0 ** 2, 32 ** 2], [32 ** 2, 96 ** 2], [96 ** 2, 1e5 ** 2]]
        self.gts, self.dts = {}, {}
        for g in gts:
            g = dict(g)
            g["ignore"] = int(bool(g.get("iscrowd", 0)))                       # COCOeval._prepare for 'bbox'
            g.setdefault("area", g["bbox"][2] * g["bbox"][3])
            self.gts.setdefault((g["image_id"], g["category_id"]), []).append(g)
        for i, d in enumerate(dts):
            d = dict(d)
            d["area"] = d["bbox"][2] * d["bbox"][3]                            # COCO.loadRes
            d["id"] = i + 1
            self.dts.setdefault((d["image_id"], d["category_id"]), []).append(d)
        self.img_ids = sorted(set(img_ids if img_ids is not None else [k[0] for k in list(self.gts) + list(self.dts)]))
        self.cat_ids = sorted(set(np.atleast_1d(cat_ids).tolist() if cat_ids is not None
                                  else [k[1] for k in list(self.gts) + list(self.dts)]))
        self.eval_imgs, self.eval, self.stats = None, None, None

    # ------------------------------------------------------------------ per image
    def _evaluate_img(self, img, cat, a_rng, max_det):
        gt, dt = self.gts.get((img, cat), []), self.dts.get((img, cat), [])
        if not gt and not dt:
            return None
        g_ig = np.array([g["ignore"] or g["area"] < a_rng[0] or g["area"] > a_rng[1] for g in gt], bool)
        gtind = np.argsort(g_ig, kind="mergesort")                             # evaluated ground truth first
        gt = [gt[i] for i in gtind]
        g_ig = g_ig[gtind].astype(int)
        dtind = np.argsort([-d["score"] for d in dt], kind="mergesort")
        dt_all = [dt[i] for i in dtind[:self.max_dets[-1]]]                      # computeIoU cut
        dt = dt_all[:max_det]
        crowd = np.array([int(g.get("iscrowd", 0)) for g in gt], int)
        ious = bbox_iou([d["bbox"] for d in dt], [g["bbox"] for g in gt], crowd)
        T, G, D = len(self.iou_thrs), len(gt), len(dt)
        gtm, dtm, dt_ig = np.zeros((T, G)), np.zeros((T, D)), np.zeros((T, D))
        if G and D:
            for ti, t in enumerate(self.iou_thrs):
                for di in range(D):
                    iou, m = min(t, 1 - 1e-10), -1
                    for gi in range(G):
                        if gtm[ti, gi] > 0 and not crowd[gi]:
                            continue
                        if m > -1 and g_ig[m] == 0 and g_ig[gi] == 1:           # ignored ground truth comes last
                            break
                        if ious[di, gi] < iou:
                            continue
                        iou, m = ious[di, gi], gi
                    if m == -1:
                        continue
                    dt_ig[ti, di], dtm[ti, di], gtm[ti, m] = g_ig[m], gt[m]["id"], dt[di]["id"]
        out_of_range = np.array([d["area"] < a_rng[0] or d["area"] > a_rng[1] for d in dt], bool).reshape(1, D)
        dt_ig = np.logical_or(dt_ig, np.logical_and(dtm == 0, np.repeat(out_of_range, T, 0)))
        return dict(dt_matches=dtm, dt_scores=np.array([d["score"] for d in dt], np.float64), gt_ignore=g_ig,
                    dt_ignore=dt_ig)

    def evaluate(self):
        self.eval_imgs = {}
        for k in self.cat_ids:
            for ai, a in enumerate(self.area_rng):
                for img in self.img_ids:
                    self.eval_imgs[(k, ai, img)] = self._evaluate_img(img, k, a, self.max_dets[-1])
        return self

    # ------------------------------------------------------------------ dataset
    def accumulate(self):
        T, R, K, A, M = len(self.iou_thrs), len(self.rec_thrs), len(self.cat_ids), len(self.area_rng), len(self.max_dets)
        precision, recall = -np.ones((T, R, K, A, M)), -np.ones((T, K, A, M))
        eps = np.spacing(1)
        for ki, k in enumerate(self.cat_ids):
            for ai in range(A):
                E = [self.eval_imgs[(k, ai, img)] for img in self.img_ids]
                E = [e for e in E if e is not None]
                if not E:
                    continue
                for mi, max_det in enumerate(self.max_dets):
                    scores = np.concatenate([e["dt_scores"][:max_det] for e in E])
                    inds = np.argsort(-scores, kind="mergesort")
                    dtm = np.concatenate([e["dt_matches"][:, :max_det] for e in E], axis=1)[:, inds]
                    dt_ig = np.concatenate([e["dt_ignore"][:, :max_det] for e in E], axis=1)[:, inds]
                    g_ig = np.concatenate([e["gt_ignore"] for e in E])
                    npig = np.count_nonzero(g_ig == 0)
                    if npig == 0:
                        continue
                    tps = np.logical_and(dtm, np.logical_not(dt_ig))
                    fps = np.logical_and(np.logical_not(dtm), np.logical_not(dt_ig))
                    tp_sum, fp_sum = np.cumsum(tps, axis=1).astype(float), np.cumsum(fps, axis=1).astype(float)
                    for ti in range(T):
                        tp, fp = tp_sum[ti], fp_sum[ti]
                        nd = len(tp)
                        rc = tp / npig
                        pr = tp / (fp + tp + eps)
                        recall[ti, ki, ai, mi] = rc[-1] if nd else 0
                        pr = pr.tolist()
                        for i in range(nd - 1, 0, -1):                          # precision envelope
                            if pr[i] > pr[i - 1]:
                                pr[i - 1] = pr[i]
                        q = np.zeros((R,))
                        for ri, pi in enumerate(np.searchsorted(rc, self.rec_thrs, side="left")):
                            if pi < nd:
                                q[ri] = pr[pi]
                        precision[ti, :, ki, ai, mi] = q
        self.eval = dict(precision=precision, recall=recall)
        return self

    def _summarize(self, ap, iou_thr=None, area=0, max_det=100):
        mi = self.max_dets.index(max_det)
        s = self.eval["precision"][:, :, :, area, mi] if ap else self.eval["recall"][:, :, area, mi]
        if iou_thr is not None:
            s = s[np.where(np.isclose(self.iou_thrs, iou_thr))[0]]
        return float(np.mean(s[s > -1])) if (s > -1).any() else -1.0

    def summarize(self):
        S = self._summarize
        self.stats = np.array([S(1), S(1, .5), S(1, .75), S(1, area=1), S(1, area=2), S(1, area=3),
                               S(0, max_det=1), S(0, max_det=10), S(0, max_det=100), S(0, area=1), S(0, area=2),
                               S(0, area=3)])
        return self.stats

    def run(self):
        self.evaluate()
        self.accumulate()
        return self.summarize()


class CocoEvaluation(DE.ObjectDetectionEvaluation):
    """utils/object_detection_evaluation.py:294-427.  Detections are collected as COCO result rows
    [image_id, x, y, w, h, score, category_id]; the ground truth comes from an annotation file at `evaluate`."""

    def __init__(self, num_groundtruth_classes, matching_iou_threshold=0.5, nms_type="standard", nms_iou_threshold=1.0,
                 nms_max_output_boxes=256, soft_nms_sigma=0.5):
        super(CocoEvaluation, self).__init__(num_groundtruth_classes, matching_iou_threshold, nms_type,
                                             nms_iou_threshold, nms_max_output_boxes, soft_nms_sigma)
        self.detection_result = np.zeros((0, 7), np.float64)
        self.max_detections_per_image = 100

    def add_single_detected_image_info(self, image_key, detected_boxes, detected_scores, detected_class_labels):
        if len(detected_boxes) != len(detected_scores) or len(detected_boxes) != len(detected_class_labels):
            raise ValueError("detected_boxes, detected_scores and detected_class_labels should all have same lengths. "
                             "Got[%d, %d, %d]" % (len(detected_boxes), len(detected_scores), len(detected_class_labels)))
        db = np.asarray(detected_boxes, float).reshape(-1, 4)
        ds, dc = np.asarray(detected_scores, float), np.asarray(detected_class_labels)
        ok = (db[:, 0] < db[:, 2]) & (db[:, 1] < db[:, 3])                      # _remove_invalid_boxes
        db, ds, dc = db[ok], ds[ok], dc[ok]
        pe = self.per_image_eval
        rows = []
        for c in range(self.num_class):
            b, s = db[dc == c], ds[dc == c]
            if not len(s):
                continue
            if pe.nms_type == "standard":
                keep = DE._standard_nms(b, s, pe.nms_max_output_boxes, pe.nms_iou_threshold)
                b, s = b[keep], s[keep]
            else:
                keep, s = DE._soft_nms(b, s, pe.nms_max_output_boxes, pe.nms_iou_threshold,
                                       2 if pe.nms_type == "soft-linear" else 3, pe.soft_nms_sigma)
                b = b[keep]
            for box, score in zip(b, s):                                       # [ymin,xmin,ymax,xmax] -> [x,y,w,h]
                rows.append([image_key, box[1], box[0], box[3] - box[1], box[2] - box[0], score, c + 1])
        # `reversed(sorted(range(n), key=score))`: descending score, the LATER row first among equal scores
        order = list(reversed(sorted(range(len(rows)), key=lambda k: rows[k][5])))[:self.max_detections_per_image]
        if order:
            self.detection_result = np.vstack([self.detection_result, np.asarray([rows[i] for i in order], np.float64)])

    def evaluate(self, eval_cat_index, eval_ann_filename):
        """-> {0: 12 stats over all categories, cat_id: 12 stats ...} for the categories in eval_cat_index, or None
        when the annotation file does not exist (the reference's behaviour)."""
        from ..data.mscoco import CocoIndex
        if isinstance(eval_ann_filename, CocoIndex):
            coco = eval_ann_filename
        elif not os.path.exists(eval_ann_filename):
            logging.warning("%s does not exists: create tf record for val", eval_ann_filename)
            return None
        else:
            coco = CocoIndex(eval_ann_filename)
        gts = list(coco.anns.values())
        dts = [dict(image_id=int(r[0]), bbox=[float(v) for v in r[1:5]], score=float(r[5]), category_id=int(r[6]))
               for r in self.detection_result]
        img_ids = sorted(coco.imgs)
        metrics = {0: CocoBBoxEval(gts, dts, img_ids, sorted(coco.cats)).run()}
        for cat_id in coco.cats:
            if cat_id in eval_cat_index:
                metrics[cat_id] = CocoBBoxEval(gts, dts, img_ids, [cat_id]).run()
        return metrics
