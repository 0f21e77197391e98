"""Generates tests/golden/assign_reference.npz by RUNNING the reference's target assigner
(/root/reference/object_detection/core/target_assigner.py `TargetAssigner.assign`, with the fork's `extension=True`
closeness targets `_create_mtl_targets`, its dead crowd / ignore branch, ArgMaxMatcher incl. force-match, IouSimilarity,
FasterRcnnBoxCoder) on the NumPy TensorFlow stand-in of tests/golden/tf_numpy_shim.py.
Run from the repo root:  python tests/golden/make_assign_golden.py"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tf_numpy_shim as shim      # noqa: E402

tf = shim.install()
# modules the assigner imports but the Faster R-CNN path never calls
for name in ("object_detection.matchers.bipartite_matcher", "object_detection.box_coders.mean_stddev_box_coder"):
    sys.modules[name] = types.ModuleType(name)
L = shim.load_reference_module
for m in ("object_detection.core.standard_fields", "object_detection.utils.shape_utils", "object_detection.core.box_list",
          "object_detection.core.box_list_ops", "object_detection.core.box_coder",
          "object_detection.box_coders.faster_rcnn_box_coder", "object_detection.core.matcher",
          "object_detection.matchers.argmax_matcher", "object_detection.core.region_similarity_calculator"):
    L(m)
ta = L("object_detection.core.target_assigner")
box_list = sys.modules["object_detection.core.box_list"]
fields = sys.modules["object_detection.core.standard_fields"]


def boxes(rng, n, lo=0, hi=200):
    y0, x0 = rng.uniform(lo, hi * 0.7, n), rng.uniform(lo, hi * 0.7, n)
    return np.stack([y0, x0, y0 + rng.uniform(8, hi * 0.5, n), x0 + rng.uniform(8, hi * 0.5, n)], 1).astype(np.float32)


def main():
    rng = np.random.default_rng(11)
    out = {}
    K1 = 6
    for case in range(4):
        G, N = int(rng.integers(1, 6)), int(rng.integers(20, 60))
        gt = boxes(rng, G)
        props = np.concatenate([gt[rng.integers(0, G, N // 2)] + rng.normal(0, 6, (N // 2, 4)).astype(np.float32),
                                boxes(rng, N - N // 2)]).astype(np.float32)
        cls = np.eye(K1, dtype=np.float32)[rng.integers(1, K1, G)]                       # one-hot with background column
        closeness = rng.random((G, K1)).astype(np.float32)
        # --- detection assigner (fmA:1670-1700): IoU 0.5 / 0.5, unmatched target = one-hot background
        assigner = ta.create_target_assigner("FasterRCNN", "detection")
        gt_list = box_list.BoxList(shim.t(gt))
        gt_list.add_field(fields.BoxListFields.closeness, shim.t(closeness))
        gt_list.add_field("ignore", shim.t(np.zeros((G,), bool)))
        unmatched = np.zeros((K1,), np.float32)
        unmatched[0] = 1
        assigner._unmatched_cls_target = shim.t(unmatched)
        r = assigner.assign(box_list.BoxList(shim.t(props)), gt_list, shim.t(cls), extension=True)
        cls_t, cls_w, reg_t, reg_w, match, close_t = r
        p = "det%d/" % case
        out[p + "gt"], out[p + "props"], out[p + "cls"], out[p + "closeness"] = gt, props, cls, closeness
        out[p + "cls_targets"], out[p + "cls_weights"] = np.asarray(cls_t), np.asarray(cls_w)
        out[p + "reg_targets"], out[p + "reg_weights"] = np.asarray(reg_t), np.asarray(reg_w)
        out[p + "match"], out[p + "closeness_targets"] = np.asarray(match.match_results), np.asarray(close_t)
        # --- proposal (RPN) assigner: 0.7 / 0.3 with force-match
        assigner = ta.create_target_assigner("FasterRCNN", "proposal")
        gt_list = box_list.BoxList(shim.t(gt))
        gt_list.add_field("ignore", shim.t(np.zeros((G,), bool)))
        r = assigner.assign(box_list.BoxList(shim.t(props)), gt_list)
        p = "rpn%d/" % case
        out[p + "cls_targets"], out[p + "cls_weights"] = np.asarray(r[0]), np.asarray(r[1])
        out[p + "reg_targets"], out[p + "reg_weights"] = np.asarray(r[2]), np.asarray(r[3])
        out[p + "match"] = np.asarray(r[4].match_results)
    out["num_cases"] = np.array(4)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "assign_reference.npz"), **out)
    print("wrote assign_reference.npz")


if __name__ == "__main__":
    main()
