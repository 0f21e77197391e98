"""Generates tests/golden/coco_aux_reference.npz by RUNNING the reference's MS-COCO record writer
(/root/reference/object_detection/create_records/create_mscoco_tf_record.py `dict_to_tf_example` :87-477: clipped
ground-truth boxes, window sampling + soft labels, closeness labels, edge mask -- the auxiliary labels are computed from
the RAW annotation boxes and indexed by the raw category id) in this container.  TensorFlow, lxml, pycocotools and the
protoc-generated label-map module do not exist here: they are replaced by recording stubs (tf.train.Feature & co. only
carry values; `coco` is a three-method stand-in over plain lists); the arithmetic is the reference's own NumPy / PIL code.
Run from the repo root:  python tests/golden/make_coco_aux_golden.py"""
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
pc = types.ModuleType("pycocotools"); pc.coco = types.ModuleType("pycocotools.coco"); pc.coco.COCO = object
sys.modules["pycocotools"], sys.modules["pycocotools.coco"] = pc, pc.coco
sys.modules["object_detection.utils.label_map_util"] = types.ModuleType("object_detection.utils.label_map_util")
gu = types.ModuleType("global_utils"); cu = types.ModuleType("global_utils.custom_utils")
cu.log = types.SimpleNamespace(info=lambda *a, **k: None, infov=lambda *a, **k: None, warn=lambda *a, **k: None,
                               warning=lambda *a, **k: None, error=lambda *a, **k: None)
gu.custom_utils = cu
sys.modules["global_utils"], sys.modules["global_utils.custom_utils"] = gu, cu
sys.path.insert(0, "/root/reference")
from object_detection.create_records import create_mscoco_tf_record as ref      # noqa: E402


class Coco(object):
    """The three calls the writer makes on pycocotools.COCO."""

    def __init__(self, anns, cats):
        self.anns, self.cats = anns, cats

    def getAnnIds(self, imgIds=None):
        return [a["id"] for a in self.anns if a["image_id"] == imgIds]

    def loadAnns(self, ids):
        return [a for a in self.anns if a["id"] in ids]

    def loadCats(self, cid):
        return [self.cats[cid]]


class _LabelMap(dict):
    """label_map_dict[name.encode('utf8')]: a str key under Python 2, bytes under Python 3."""

    def __getitem__(self, k):
        return dict.__getitem__(self, k.decode() if isinstance(k, bytes) else k)


def values(feature):
    for k in ("bytes_list", "float_list", "int64_list"):
        if hasattr(feature, k):
            return list(getattr(feature, k).value)
    return []


def main():
    from PIL import Image
    out = {}
    tmp = tempfile.mkdtemp()
    rng = np.random.default_rng(9)
    cat_ids = [1, 2, 3, 5, 7]                         # COCO category ids have gaps: the labels get max(id) + 1 columns
    cats = {c: {"id": c, "name": "c%d" % c} for c in cat_ids}
    label_map = _LabelMap({"c%d" % c: c for c in cat_ids})
    class_indices = sorted(label_map.values())
    index_map = {v: k for k, v in label_map.items()}
    cases = 0
    for case in range(6):
        H, W = int(rng.integers(120, 400)), int(rng.integers(120, 500))
        img_id = 1000 + case
        name = os.path.join(tmp, "COCO_val2014_%012d.jpg" % img_id)
        Image.fromarray(rng.integers(0, 256, (H, W, 3), dtype=np.uint8), "RGB").save(name)
        anns = []
        for j in range(int(rng.integers(1, 6))):
            x0, y0 = float(rng.uniform(-10, W - 40)), float(rng.uniform(-10, H - 40))     # some boxes cross the border
            w, h = float(rng.uniform(20, W - max(x0, 0) + 15)), float(rng.uniform(20, H - max(y0, 0) + 15))
            anns.append({"id": case * 10 + j, "image_id": img_id, "bbox": [round(x0, 2), round(y0, 2), round(w, 2), round(h, 2)],
                         "category_id": int(rng.choice(cat_ids)), "iscrowd": int(rng.random() < 0.2)})
        random.seed(200 + case)
        np.random.seed(200 + case)
        ex = ref.dict_to_tf_example(label_map, name, Coco(anns, cats), class_indices, index_map, "val")
        f = ex.features.feature
        p = "case%d/" % case
        out[p + "hw"] = np.array([H, W])
        out[p + "bbox"] = np.array([a["bbox"] for a in anns], np.float64)
        out[p + "category_id"] = np.array([a["category_id"] for a in anns])
        out[p + "iscrowd"] = np.array([a["iscrowd"] for a in anns])
        for k in ("ymin", "xmin", "ymax", "xmax"):
            out[p + "gt_" + k] = np.array(values(f["image/object/bbox/" + k]), np.float64)
            out[p + "window_" + k] = np.array(values(f["image/window/bbox/" + k]), np.float64)
        out[p + "gt_label"] = np.array(values(f["image/object/class/label"]))
        out[p + "gt_is_crowd"] = np.array(values(f["image/object/is_crowd"]))
        out[p + "source_id"] = np.array(values(f["image/source_id"])[0].decode())
        out[p + "window_labels"] = np.array([v.decode() if isinstance(v, bytes) else v
                                             for v in values(f["image/window/labels/text"])], dtype="U512")
        out[p + "closeness"] = np.array([v.decode() if isinstance(v, bytes) else v
                                         for v in values(f["image/object/closeness/text"])], dtype="U512")
        out[p + "edgemask"] = np.array(values(f["image/edgemask/masks"]), np.float32)
        out[p + "edgemask_hw"] = np.array([values(f["image/edgemask/height"])[0], values(f["image/edgemask/width"])[0]])
        cases += 1
    out["num_cases"] = np.array(cases)
    out["class_indices"] = np.array(class_indices)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "coco_aux_reference.npz"), **out)
    print("wrote coco_aux_reference.npz:", cases, "cases")


if __name__ == "__main__":
    main()
