"""PASCAL-VOC annotations -> training examples / tf.Example records with the fork's auxiliary labels, without the
offline TFRecord step (SURVEY 8f N1 + N2).

Restates the data flow of /root/reference/object_detection/create_records/create_pascal_tf_record.py:68-499
(`dict_to_tf_example`: boxes normalised by the image size, class ids from the label map, difficult / truncated / pose,
the single area subset 'all', window boxes + soft labels, closeness labels, 64x64 edge mask) on top of
data/aux_labels.py, and of utils/dataset_util.py `recursive_parse_xml_to_dict` (every tag a key, 'object' a list).
The standard keys are pinned to create_pascal_tf_record_test.py:39-113; the auxiliary labels (aux_labels.py) to outputs
of the reference writer run under recording stubs (tests/golden/make_aux_golden.py); its window sampling draws from the
global RNG, here from an explicit generator."""
import hashlib
import io
import os
import xml.etree.ElementTree as ET

import numpy as np

from . import aux_labels, tfrecord


def recursive_parse_xml_to_dict(node):
    """dataset_util.recursive_parse_xml_to_dict: leaf -> {tag: text}; children merged into one dict, every
    <object> collected in a list."""
    children = list(node)
    if not children:
        return {node.tag: node.text}
    out = {}
    for child in children:
        sub = recursive_parse_xml_to_dict(child)
        if child.tag != "object":
            out[child.tag] = sub[child.tag]
        else:
            out.setdefault("object", []).append(sub["object"])
    return {node.tag: out}


def parse_annotation(xml_text_or_path):
    text = open(xml_text_or_path).read() if os.path.exists(str(xml_text_or_path)) else xml_text_or_path
    return recursive_parse_xml_to_dict(ET.fromstring(text))["annotation"]


def _objects(data, ignore_difficult_instances):
    objs = []
    for o in data.get("object", []):
        if ignore_difficult_instances and bool(int(o["difficult"])):
            continue
        objs.append(o)
    return objs


def annotation_to_example(data, image, label_map_dict, num_classes, rng, num_windows=64,
                          ignore_difficult_instances=False):
    """One annotation dict + decoded image (uint8 / float [H,W,3]) -> example dict (data/synthetic.py format) with
    window / closeness / edge-mask labels computed on the fly."""
    image = np.asarray(image)
    H, W = image.shape[:2]
    if "size" in data:
        W, H = int(data["size"]["width"]), int(data["size"]["height"])
    objs = _objects(data, ignore_difficult_instances)
    boxes = np.array([[float(o["bndbox"][k]) for k in ("ymin", "xmin", "ymax", "xmax")] for o in objs],
                     np.float64).reshape(-1, 4)
    classes = np.array([label_map_dict[o["name"]] for o in objs], np.int64)
    onehot = np.zeros((len(objs), num_classes), np.float32)
    onehot[np.arange(len(objs)), classes - 1] = 1.0
    norm = (boxes / np.array([H, W, H, W], np.float64)).astype(np.float32)
    wb, wl = aux_labels.random_windows(boxes, classes, float(H), float(W), num_classes, rng, num_windows)
    return dict(image=image.astype(np.float32), groundtruth_boxes=norm, groundtruth_classes=onehot,
                groundtruth_difficult=np.array([bool(int(o["difficult"])) for o in objs], bool),
                groundtruth_closeness=aux_labels.closeness_labels(boxes, classes, H, W, num_classes),
                window_boxes=wb, window_classes=wl, groundtruth_edgemask=aux_labels.edgemask(boxes, float(H), float(W)),
                filename=data.get("filename", ""))


def dict_to_tf_example(data, dataset_directory, label_map_dict, num_classes=None, rng=None,
                       ignore_difficult_instances=False, image_subdirectory="JPEGImages", num_windows=64):
    """create_pascal_tf_record.py:68-499 -> serialized tf.Example bytes with the record writer's key set."""
    img_path = os.path.join(data.get("folder") or "", image_subdirectory, data["filename"])
    encoded = open(os.path.join(dataset_directory, img_path), "rb").read()
    from PIL import Image
    img = Image.open(io.BytesIO(encoded))
    if img.format != "JPEG":
        raise ValueError("Image format not JPEG")
    K = num_classes if num_classes is not None else max(label_map_dict.values())
    ex = annotation_to_example(data, np.asarray(img.convert("RGB")), label_map_dict, K,
                               rng if rng is not None else np.random.default_rng(0), num_windows,
                               ignore_difficult_instances)
    objs = _objects(data, ignore_difficult_instances)
    H, W = (int(data["size"]["height"]), int(data["size"]["width"])) if "size" in data else (img.height, img.width)
    name = data["filename"].encode("utf8")
    b = ex["groundtruth_boxes"]
    f = {
        "image/height": np.asarray([H], np.int64), "image/width": np.asarray([W], np.int64),
        "image/filename": [name], "image/source_id": [name],
        "image/key/sha256": [hashlib.sha256(encoded).hexdigest().encode("utf8")],
        "image/encoded": [encoded], "image/format": [b"jpeg"],
        "image/object/bbox/xmin": b[:, 1], "image/object/bbox/xmax": b[:, 3],
        "image/object/bbox/ymin": b[:, 0], "image/object/bbox/ymax": b[:, 2],
        "image/object/class/text": [o["name"].encode("utf8") for o in objs],
        "image/object/class/label": np.asarray([label_map_dict[o["name"]] for o in objs], np.int64),
        "image/object/difficult": np.asarray([int(bool(int(o["difficult"]))) for o in objs], np.int64),
        "image/object/truncated": np.asarray([int(o["truncated"]) for o in objs], np.int64),
        "image/object/view": [(o.get("pose") or "").encode("utf8") for o in objs],
        "image/object/subset": [b"all"] * len(objs),                 # the writer's only subset: area in [0, inf)
        "image/object/label_type": [b""] * len(objs),
        "image/object/closeness/text": [tfrecord._fmt_row(r).encode() for r in ex["groundtruth_closeness"]],
        "image/window/labels/text": [tfrecord._fmt_row(r).encode() for r in ex["window_classes"]],
        "image/edgemask/masks": np.asarray(ex["groundtruth_edgemask"], np.float32).reshape(-1),
        "image/edgemask/height": np.asarray([ex["groundtruth_edgemask"].shape[1]], np.int64),
        "image/edgemask/width": np.asarray([ex["groundtruth_edgemask"].shape[2]], np.int64),
    }
    wb = np.asarray(ex["window_boxes"], np.float32).reshape(-1, 4)
    for i, k in enumerate(("ymin", "xmin", "ymax", "xmax")):
        f["image/window/bbox/" + k] = wb[:, i]
    return tfrecord.serialize_example(f)


class VocDataset(object):
    """Examples straight from a VOCdevkit year directory (Annotations/, JPEGImages/, ImageSets/Main/<set>.txt)."""

    def __init__(self, year_dir, image_set, label_map_dict, num_classes, seed=0, num_windows=64,
                 ignore_difficult_instances=False):
        self.dir, self.label_map, self.K = year_dir, label_map_dict, num_classes
        self.ids = [l.split()[0] for l in open(os.path.join(year_dir, "ImageSets", "Main", image_set + ".txt")) if l.strip()]
        self.seed, self.num_windows, self.ignore_difficult = seed, num_windows, ignore_difficult_instances

    def __len__(self):
        return len(self.ids)

    def __iter__(self):
        from PIL import Image
        rng = np.random.default_rng(self.seed)
        for i in self.ids:
            data = parse_annotation(os.path.join(self.dir, "Annotations", i + ".xml"))
            img = np.asarray(Image.open(os.path.join(self.dir, "JPEGImages", data["filename"])).convert("RGB"))
            yield annotation_to_example(data, img, self.label_map, self.K, rng, self.num_windows, self.ignore_difficult)
