"""MS-COCO instance annotations -> training examples / tf.Example records with the fork's auxiliary labels, without the
offline TFRecord step and without pycocotools (the COCO twin of data/pascal_voc.py).

Restates the data flow of /root/reference/object_detection/create_records/create_mscoco_tf_record.py:
  :73-85    `boundary_check`: the ground-truth box [x, y, w, h] is clamped into the image; boxes that end up empty are
            dropped from the GROUND TRUTH only;
  :87-477   `dict_to_tf_example`: ground-truth boxes normalised by the image size, labels from the label map, `is_crowd`;
            the auxiliary labels (windows + soft labels :176-284, closeness :306-340, edge mask :358-393) are computed from
            ALL annotations of the image with their RAW (unclamped) boxes and are indexed by the raw `category_id`, so
            they have max(class id) + 1 columns (91 for COCO's sparse 1..90 ids);
  :492-494  `getImageId`: the integer between the last '_' / '/' and the extension of the file name.
Pinned to outputs of that writer RUN here under recording stubs (tests/golden/make_coco_aux_golden.py)."""
import hashlib
import io
import json
import os
import re

import numpy as np

from . import aux_labels, tfrecord


def boundary_check(bbox, width, height):
    l, t, w, h = bbox[:4]
    l = max(0, min(width, l))
    t = max(0, min(height, t))
    w = max(0, min(width - l, w))
    h = max(0, min(height - t, h))
    return l, t, w, h


def get_image_id(filename):
    return int(re.split(r"_|/|\.", filename)[-2])


class CocoIndex(object):
    """The part of pycocotools.COCO the record writer uses: images, annotations per image, categories."""

    def __init__(self, annotation_file_or_dict):
        d = annotation_file_or_dict
        if not isinstance(d, dict):
            with open(d) as f:
                d = json.load(f)
        self.imgs = {im["id"]: im for im in d.get("images", [])}
        self.cats = {c["id"]: c for c in d.get("categories", [])}
        self.anns = {a["id"]: a for a in d.get("annotations", [])}
        self._by_image = {}
        for a in d.get("annotations", []):
            self._by_image.setdefault(a["image_id"], []).append(a["id"])

    def getAnnIds(self, imgIds):
        return list(self._by_image.get(imgIds, []))

    def loadAnns(self, ids):
        return [self.anns[i] for i in ids]

    def loadCats(self, ids):
        return [self.cats[i] for i in (ids if isinstance(ids, (list, tuple)) else [ids])]

    def label_map_dict(self):
        return {c["name"]: c["id"] for c in self.cats.values()}


def _raw_boxes(anns):
    """`get_box_coord` (:300-304): (ymin, xmin, ymax, xmax) from the UNCLAMPED [x, y, w, h]."""
    return np.array([[a["bbox"][1], a["bbox"][0], a["bbox"][1] + a["bbox"][3], a["bbox"][0] + a["bbox"][2]] for a in anns],
                    np.float64).reshape(-1, 4)


def annotations_to_example(anns, image, cats, label_map_dict, num_classes, rng, num_windows=64, max_class_index=None):
    """Annotations of one image + decoded image -> example dict (data/synthetic.py format).
    `num_classes`: width of the one-hot ground-truth classes (label-map ids are 1-based);
    `max_class_index`: max(class_indices) of the writer = columns - 1 of the auxiliary labels (default: num_classes)."""
    image = np.asarray(image)
    H, W = image.shape[:2]
    kmax = num_classes if max_class_index is None else max_class_index
    gt, labels, crowd, kept = [], [], [], []
    for i, a in enumerate(anns):
        l, t, w, h = boundary_check(a["bbox"], W, H)
        if not (W > 0 and H > 0 and w > 0 and h > 0):
            continue
        gt.append([float(t) / H, float(l) / W, float(h + t) / H, float(w + l) / W])
        labels.append(label_map_dict[cats[a["category_id"]]["name"]])
        crowd.append(int(a.get("iscrowd", 0)))
        kept.append(i)
    raw = _raw_boxes(anns)
    cat_ids = np.array([a["category_id"] for a in anns], np.int64)
    onehot = np.zeros((len(gt), num_classes), np.float32)
    if len(gt):
        onehot[np.arange(len(gt)), np.asarray(labels) - 1] = 1.0
    closeness = aux_labels.closeness_labels(raw, cat_ids, H, W, kmax)[kept] if len(anns) else np.zeros((0, kmax + 1), np.float32)
    if len(anns):
        wb, wl = aux_labels.random_windows(raw, cat_ids, float(H), float(W), kmax, rng, num_windows)
    else:                                                   # `create_multi_object` returns [] without annotations
        wb, wl = np.zeros((0, 4), np.float32), np.zeros((0, kmax + 1), np.float32)
    return dict(image=image.astype(np.float32), groundtruth_boxes=np.asarray(gt, np.float32).reshape(-1, 4),
                groundtruth_classes=onehot, groundtruth_is_crowd=np.asarray(crowd, bool),
                groundtruth_labels=np.asarray(labels, np.int64), groundtruth_closeness=closeness,
                window_boxes=wb, window_classes=wl, groundtruth_edgemask=aux_labels.edgemask(raw, float(H), float(W)))


def dict_to_tf_example(label_map_dict, image_name, coco, class_indices, rng=None, num_windows=64):
    """create_mscoco_tf_record.py:87-477 -> serialized tf.Example bytes with the record writer's key set."""
    from PIL import Image
    img_id = get_image_id(image_name)
    anns = coco.loadAnns(coco.getAnnIds(imgIds=img_id))
    encoded = open(image_name, "rb").read()
    img = Image.open(io.BytesIO(encoded))
    if img.format != "JPEG":
        raise ValueError("Image format not JPEG")
    cats = {a["category_id"]: coco.loadCats(a["category_id"])[0] for a in anns}
    ex = annotations_to_example(anns, np.asarray(img.convert("RGB")), cats, label_map_dict, max(class_indices),
                                rng if rng is not None else np.random.default_rng(0), num_windows, max(class_indices))
    b = ex["groundtruth_boxes"]
    names = [k for k in (cats[a["category_id"]]["name"] for a in anns)]
    kept_names = [n for a, n in zip(anns, names) if min(boundary_check(a["bbox"], img.width, img.height)[2:]) > 0]
    f = {
        "image/height": np.asarray([img.height], np.int64), "image/width": np.asarray([img.width], np.int64),
        "image/filename": [image_name.encode("utf8")], "image/source_id": [str(img_id).encode("utf8")],
        "image/key/sha256": [hashlib.sha256(encoded).hexdigest().encode("utf8")],
        "image/encoded": [encoded], "image/format": [b"jpg"],
        "image/object/bbox/xmin": b[:, 1], "image/object/bbox/xmax": b[:, 3],
        "image/object/bbox/ymin": b[:, 0], "image/object/bbox/ymax": b[:, 2],
        "image/object/class/text": [n.encode("utf8") for n in kept_names],
        "image/object/class/label": ex["groundtruth_labels"],
        "image/object/is_crowd": ex["groundtruth_is_crowd"].astype(np.int64),
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


class CocoDataset(object):
    """Examples straight from an annotation file + image directory (e.g. annotations/instances_train2017.json,
    images/train2017/)."""

    def __init__(self, annotation_file, image_dir, num_classes=90, seed=0, num_windows=64):
        self.coco = annotation_file if isinstance(annotation_file, CocoIndex) else CocoIndex(annotation_file)
        self.image_dir, self.K, self.seed, self.num_windows = image_dir, num_classes, seed, num_windows
        self.label_map = self.coco.label_map_dict()
        self.ids = sorted(self.coco.imgs)

    def __len__(self):
        return len(self.ids)

    def __iter__(self):
        from PIL import Image
        rng = np.random.default_rng(self.seed)
        for i in self.ids:
            anns = self.coco.loadAnns(self.coco.getAnnIds(imgIds=i))
            img = np.asarray(Image.open(os.path.join(self.image_dir, self.coco.imgs[i]["file_name"])).convert("RGB"))
            ex = annotations_to_example(anns, img, self.coco.cats, self.label_map, self.K, rng, self.num_windows)
            ex["source_id"], ex["filename"] = str(i), self.coco.imgs[i]["file_name"]       # image id for the COCO metrics
            yield ex
