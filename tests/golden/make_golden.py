"""Generates tests/golden/np_reference.npz by running the reference's OWN importable NumPy code
(/root/reference/object_detection/utils/np_box_ops.py, np_box_list_ops.py, np_box_list.py) on
seeded inputs.  Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    # import the three pure-NumPy modules without triggering object_detection/__init__ side effects
    pkg = types.ModuleType("object_detection"); pkg.__path__ = [os.path.join(REF, "object_detection")]
    utils = types.ModuleType("object_detection.utils"); utils.__path__ = [os.path.join(REF, "object_detection", "utils")]
    sys.modules["object_detection"] = pkg
    sys.modules["object_detection.utils"] = utils
    mods = {}
    for name in ("np_box_ops", "np_box_list", "np_box_list_ops"):
        spec = importlib.util.spec_from_file_location("object_detection.utils." + name,
                                                      os.path.join(REF, "object_detection", "utils", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules["object_detection.utils." + name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    mods["np_box_list_ops"].xrange = range          # Python-2 builtin used at np_box_list_ops.py:232
    return mods


def rand_boxes(rng, n, h, w):
    cy = rng.uniform(0, h, n); cx = rng.uniform(0, w, n)
    bh = np.exp(rng.uniform(np.log(4.0), np.log(h * 0.8), n)); bw = np.exp(rng.uniform(np.log(4.0), np.log(w * 0.8), n))
    b = np.stack([cy - bh / 2, cx - bw / 2, cy + bh / 2, cx + bw / 2], 1)
    return b.astype(np.float32)


def main():
    m = _load()
    ops, bl, blo = m["np_box_ops"], m["np_box_list"], m["np_box_list_ops"]
    rng = np.random.default_rng(20261017)
    out = {}
    a, b = rand_boxes(rng, 40, 300, 400), rand_boxes(rng, 300, 300, 400)
    out["iou_a"], out["iou_b"] = a, b
    out["iou"] = ops.iou(a, b).astype(np.float32)
    out["ioa"] = ops.ioa(a, b).astype(np.float32)
    out["intersection"] = ops.intersection(a, b).astype(np.float32)
    out["area"] = ops.area(b).astype(np.float32)
    # greedy NMS (np_box_list_ops.non_max_suppression :185-257), clustered boxes, unique scores
    nb = rand_boxes(rng, 400, 300, 400)
    nb[200:] = nb[:200] + rng.uniform(-5, 5, (200, 4)).astype(np.float32)
    bad = (nb[:, 2] <= nb[:, 0]) | (nb[:, 3] <= nb[:, 1])
    nb[bad] = [10, 10, 50, 50]
    sc = rng.permutation(400).astype(np.float32) / 400.0 + 0.001
    boxlist = bl.BoxList(nb.astype(np.float32))
    boxlist.add_field("scores", sc)
    for thr, mx in ((0.5, 50), (0.7, 300)):
        res = blo.non_max_suppression(boxlist, max_output_size=mx, iou_threshold=thr, score_threshold=0.0)
        out["nms_%g_%d_boxes" % (thr, mx)] = res.get().astype(np.float32)
        out["nms_%g_%d_scores" % (thr, mx)] = res.get_field("scores").astype(np.float32)
    out["nms_boxes_in"], out["nms_scores_in"] = nb.astype(np.float32), sc
    # clip / prune
    window = np.array([20.0, 30.0, 250.0, 350.0], np.float32)
    cl = blo.clip_to_window(bl.BoxList(b.copy()), window)
    out["window"] = window
    out["clip_to_window"] = cl.get().astype(np.float32)
    pr = blo.prune_outside_window(bl.BoxList(b.copy()), window)
    out["prune_outside_window_boxes"] = pr[0].get().astype(np.float32)
    out["prune_outside_window_idx"] = np.asarray(pr[1]).reshape(-1).astype(np.int64)
    np.savez_compressed(os.path.join(HERE, "np_reference.npz"), **out)
    print("wrote np_reference.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
