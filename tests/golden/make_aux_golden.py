"""Generates tests/golden/aux_reference.npz by RUNNING the reference's record writer
(/root/reference/object_detection/create_records/create_pascal_tf_record.py `dict_to_tf_example`: window sampling +
soft labels, closeness labels, edge mask) in this container.  TensorFlow, lxml and the protoc-generated label-map module
do not exist here: they are replaced by recording stubs (tf.train.Feature & co. only carry values, tf.app.flags only
holds defaults); the arithmetic that produces the auxiliary labels is the reference's own NumPy / PIL code, unmodified.
Run from the repo root:  python tests/golden/make_aux_golden.py"""
import builtins
import os
import random
import sys
import tempfile
import types

import numpy as np

np.bool, np.float, np.NAN = bool, float, np.nan
builtins.xrange = range


class _Rec(object):
    def __init__(self, **kw):
        self.__dict__.update(kw)


class _Flags(object):
    pass


FLAGS = _Flags()
flags = types.SimpleNamespace(FLAGS=FLAGS)
for kind in ("string", "boolean", "integer", "float"):
    setattr(flags, "DEFINE_" + kind, lambda name, default, doc="", **kw: setattr(FLAGS, name, default))
tf = types.ModuleType("tensorflow")
tf.app = types.SimpleNamespace(flags=flags, run=lambda *a, **k: None)
tf.gfile = types.SimpleNamespace(GFile=open)
tf.train = types.SimpleNamespace(Example=_Rec, Features=_Rec, Feature=_Rec, BytesList=_Rec, FloatList=_Rec, Int64List=_Rec)
tf.python_io = types.SimpleNamespace(TFRecordWriter=None)
sys.modules["tensorflow"] = tf
lx = types.ModuleType("lxml"); lx.etree = types.ModuleType("lxml.etree")
sys.modules["lxml"], sys.modules["lxml.etree"] = lx, lx.etree
sys.modules["object_detection.utils.label_map_util"] = types.ModuleType("object_detection.utils.label_map_util")
gu = types.ModuleType("global_utils"); cu = types.ModuleType("global_utils.custom_utils")      # logging helper (colorlog absent)
cu.log = types.SimpleNamespace(info=lambda *a, **k: None, infov=lambda *a, **k: None, warn=lambda *a, **k: None,
                               warning=lambda *a, **k: None, error=lambda *a, **k: None)
gu.custom_utils = cu
sys.modules["global_utils"], sys.modules["global_utils.custom_utils"] = gu, cu
sys.path.insert(0, "/root/reference")
from object_detection.create_records import create_pascal_tf_record as ref      # noqa: E402


def values(feature):
    for k in ("bytes_list", "float_list", "int64_list"):
        if hasattr(feature, k):
            return list(getattr(feature, k).value)
    return []


def main():
    from PIL import Image
    out = {}
    tmp = tempfile.mkdtemp()
    os.makedirs(os.path.join(tmp, "JPEGImages"))
    rng = np.random.default_rng(5)
    K = 4
    label_map = {"c%d" % i: i for i in range(1, K + 1)}
    cases = 0
    for case in range(6):
        H, W = int(rng.integers(120, 400)), int(rng.integers(120, 500))
        name = "im%d.jpg" % case
        Image.fromarray(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), "RGB").save(os.path.join(tmp, "JPEGImages", name))
        objs = []
        for _ in range(int(rng.integers(1, 6))):
            y0, x0 = int(rng.integers(0, H - 40)), int(rng.integers(0, W - 40))
            y1, x1 = int(rng.integers(y0 + 20, H + 1)), int(rng.integers(x0 + 20, W + 1))
            objs.append({"name": "c%d" % int(rng.integers(1, K + 1)), "difficult": str(int(rng.random() < 0.2)),
                         "truncated": "0", "pose": "Unspecified",
                         "bndbox": {"xmin": str(x0), "ymin": str(y0), "xmax": str(x1), "ymax": str(y1)}})
        data = {"folder": "", "filename": name, "size": {"width": str(W), "height": str(H)}, "object": objs}
        random.seed(100 + case)
        np.random.seed(100 + case)
        ex = ref.dict_to_tf_example(data, tmp, label_map, list(range(1, K + 1)))
        f = ex.features.feature
        p = "case%d/" % case
        out[p + "hw"] = np.array([H, W])
        out[p + "boxes"] = np.array([[float(o["bndbox"][k]) for k in ("ymin", "xmin", "ymax", "xmax")] for o in objs])
        out[p + "classes"] = np.array([label_map[o["name"]] for o in objs])
        out[p + "difficult"] = np.array([int(o["difficult"]) for o in objs])
        for k in ("ymin", "xmin", "ymax", "xmax"):
            out[p + "window_" + k] = np.array(values(f["image/window/bbox/" + k]), np.float64)
        out[p + "window_labels"] = np.array([v.decode() if isinstance(v, bytes) else v
                                             for v in values(f["image/window/labels/text"])], dtype="U256")
        out[p + "closeness"] = np.array([v.decode() if isinstance(v, bytes) else v
                                         for v in values(f["image/object/closeness/text"])], dtype="U256")
        out[p + "edgemask"] = np.array(values(f["image/edgemask/masks"]), np.float32)
        out[p + "edgemask_hw"] = np.array([values(f["image/edgemask/height"])[0], values(f["image/edgemask/width"])[0]])
        cases += 1
    out["num_cases"] = np.array(cases)
    out["num_classes"] = np.array(K)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "aux_reference.npz"), **out)
    print("wrote aux_reference.npz:", cases, "cases")


if __name__ == "__main__":
    main()
