"""Evaluation loop on a B200: inference-mode model -> `predict` -> `postprocess` -> PASCAL-VOC and MTL metrics.
Host side of /root/reference/object_detection/evaluator.py:100-230 (`_extract_prediction_tensors`: which tensors are
collected per image) and eval_util.py; every array leaves the device once per image.

Collected per image, as in the reference: detections in ABSOLUTE image coordinates (eval_util.py
`result_dict_for_single_example` scales the normalised boxes by the image size), ground truth likewise, and for the
auxiliary tasks `window_class_predictions` on the ground-truth windows, `closeness_predictions` (one row per PROPOSAL --
the reference indexes these rows with a detection index, mtl_util.py:66-73 with evaluator.py:224; reproduced as is,
trap T17) and `edgemask_predictions`."""
import numpy as np
import torch

from . import eval_util
from .utils import mtl_metrics


def run_inference(model, example, use_refiner=False):
    """One example dict (data/synthetic.py / data/tfrecord.py format) through the inference-mode model.
    Returns the per-image result dict (NumPy).
    use_refiner: run `predict_with_mtl_results` on the inference proposals when the config enables the class refiner, so
    that `postprocess` scores the refined logits -- what the reference's evaluator does (evaluator.py:151-152,
    fmA:1040-1043).  Off by default until the path has been run on a GPU."""
    if model._is_training:
        raise ValueError("evaluation needs a model built with is_training=False")
    dev = model.device
    image = torch.from_numpy(np.ascontiguousarray(example["image"], np.float32))[None].to(dev)
    H, W = image.shape[1], image.shape[2]
    pd = model.predict(model.preprocess(image))
    mtl = model._mtl
    if mtl is not None and mtl.window and example.get("window_boxes") is not None:
        wb = torch.from_numpy(np.asarray(example["window_boxes"], np.float32))[None].to(dev)
        pd = model.predict_with_window(pd, window_boxes_normalized=wb, _keep=False)
    if mtl is not None and mtl.edgemask:
        pd = model.predict_edgemask(pd)
    if use_refiner and mtl is not None and mtl.refine:
        pd = model.predict_with_mtl_results(pd)
    det = model.postprocess(pd)
    torch.cuda.synchronize()          # the auxiliary heads ran on side streams ("lanes"): wait for all of them
    n = int(det["num_detections"][0].item())
    scale = np.array([H, W, H, W], np.float32)
    K = model.num_classes
    out = {
        "detection_boxes": det["detection_boxes"][0, :n].cpu().numpy() * scale,
        "detection_scores": det["detection_scores"][0, :n].cpu().numpy(),
        "detection_classes": det["detection_classes"][0, :n].cpu().numpy().astype(np.int64) + 1,     # label_id_offset
        "groundtruth_boxes": np.asarray(example["groundtruth_boxes"], np.float32).reshape(-1, 4) * scale,
        "groundtruth_classes": np.asarray(example["groundtruth_classes"]).reshape(-1, K).argmax(1) + 1,
    }
    for k in ("groundtruth_difficult", "groundtruth_subset"):
        if k in example:
            out["difficult" if k == "groundtruth_difficult" else k] = np.asarray(example[k])
    if "window_class_predictions" in pd and example.get("window_classes") is not None:
        out["window_classes_gt"] = list(np.asarray(example["window_classes"], np.float32))
        out["window_classes_dt"] = pd["window_class_predictions"].float().cpu().numpy().reshape(len(out["window_classes_gt"]), -1)
    if "closeness_predictions" in pd and example.get("groundtruth_closeness") is not None:
        out["closeness_gt"] = list(np.asarray(example["groundtruth_closeness"], np.float32))
        cd = pd["closeness_predictions"].float().cpu().numpy()
        if n > len(cd):      # more detections than proposals (small configs): rows past P read as zero logits, where the
            cd = np.pad(cd, ((0, n - len(cd)), (0, 0)))          # reference (max 300 of each) would index out of range
        out["closeness_dt"] = cd
    if "edgemask_predictions" in pd and example.get("groundtruth_edgemask") is not None:
        out["edgemask_gt"] = np.asarray(example["groundtruth_edgemask"], np.float32)
        out["edgemask_dt"] = pd["edgemask_predictions"].float().cpu().numpy()
    return out


EVAL_METRICS_FN_DICT = {                      # evaluator.py:40-43 of the reference
    "pascal_voc_metrics": eval_util.evaluate_detection_results_pascal_voc,
    "coco_metrics": eval_util.evaluate_detection_results_coco,
}


def detection_metrics(lists, categories, eval_config=None, metrics_set=None, iou_thres=None, corloc_summary=True,
                      eval_ann_filename=None):
    """The metric call of the reference's evaluator (evaluator.py:312-323): `eval_config` (eval.proto) supplies
    metrics_set, iou_threshold, nms_type, nms_threshold, soft_nms_sigma and coco_eval_options; explicit arguments win."""
    g = lambda name, default: getattr(eval_config, name, default) if eval_config is not None else default
    metrics_set = metrics_set or g("metrics_set", "pascal_voc_metrics")
    if metrics_set not in EVAL_METRICS_FN_DICT:
        raise ValueError("Metric not found: {}".format(metrics_set))
    kw = dict(iou_thres=g("iou_threshold", 0.5) if iou_thres is None else iou_thres, nms_type=g("nms_type", "standard"),
              nms_thres=g("nms_threshold", 1.0), soft_nms_sigma=g("soft_nms_sigma", 0.5))
    if metrics_set == "coco_metrics":
        return eval_util.evaluate_detection_results_coco(lists, categories, eval_config=eval_config,
                                                         eval_ann_filename=eval_ann_filename, **kw)
    return eval_util.evaluate_detection_results_pascal_voc(lists, categories, corloc_summary=corloc_summary, **kw)


def evaluate(model, examples, categories, iou_thres=None, corloc_summary=True, metrics_set=None, eval_config=None,
             eval_ann_filename=None, use_refiner=False):
    """Runs `examples` through the model and returns {metric name: value}: the detection metrics of eval_util
    (`metrics_set` as in eval.proto: 'pascal_voc_metrics' or 'coco_metrics', taken from `eval_config` when not given; for
    COCO the examples carry the image id in 'source_id' and the ground truth is read from the annotation file) plus
    'mtl/window_map', 'mtl/closeness_diff', 'mtl/edgemask_ap' for the auxiliary heads the config enables."""
    coco = (metrics_set or getattr(eval_config, "metrics_set", "pascal_voc_metrics")) == "coco_metrics"
    lists = {}
    for i, ex in enumerate(examples):
        r = run_inference(model, ex, use_refiner)
        r["image_id"] = str(ex.get("source_id", i)) if coco else str(i)
        for k, v in r.items():
            lists.setdefault(k, []).append(v)
    metrics = detection_metrics(lists, categories, eval_config, metrics_set, iou_thres, corloc_summary, eval_ann_filename)
    has_dets = all(len(d) for d in lists["detection_boxes"]) and all(len(g) for g in lists["groundtruth_boxes"])
    aux = {k: v for k, v in lists.items() if k.split("_")[0] in ("window", "edgemask") or
           (k.startswith("closeness") and has_dets)}
    aux["groundtruth_boxes"], aux["detection_boxes"] = lists["groundtruth_boxes"], lists["detection_boxes"]
    metrics.update(mtl_metrics.get_mtl_metrics(aux))
    return metrics
