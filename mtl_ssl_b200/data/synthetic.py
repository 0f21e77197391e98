"""Seeded synthetic detection batches in the shape of the reference's input queue
(/root/reference/object_detection/trainer.py:100-156 `_get_inputs`): image, normalised boxes,
one-hot classes, closeness labels, 64 windows + soft labels, 64x64 edgemask.  There is no network
and no dataset in this environment: images are uniform noise, boxes are log-uniform."""
import numpy as np

from . import aux_labels


def make_example(rng, height, width, num_classes, max_boxes=8, num_windows=64):
    g = int(rng.integers(1, max_boxes + 1))
    cy, cx = rng.uniform(0, height, g), rng.uniform(0, width, g)
    bh = np.exp(rng.uniform(np.log(32.0), np.log(0.8 * height), g))
    bw = np.exp(rng.uniform(np.log(32.0), np.log(0.8 * width), g))
    boxes = np.stack([cy - bh / 2, cx - bw / 2, cy + bh / 2, cx + bw / 2], 1)
    boxes[:, [0, 2]] = np.clip(boxes[:, [0, 2]], 0, height)
    boxes[:, [1, 3]] = np.clip(boxes[:, [1, 3]], 0, width)
    keep = ((boxes[:, 2] - boxes[:, 0]) >= 8) & ((boxes[:, 3] - boxes[:, 1]) >= 8)
    boxes = boxes[keep] if keep.any() else np.array([[0.25 * height, 0.25 * width, 0.75 * height, 0.75 * width]])
    g = len(boxes)
    classes = rng.integers(1, num_classes + 1, g)
    image = rng.integers(0, 256, (height, width, 3)).astype(np.float32)
    norm = (boxes / np.array([height, width, height, width])).astype(np.float32)
    onehot = np.zeros((g, num_classes), np.float32)
    onehot[np.arange(g), classes - 1] = 1
    wb, wl = aux_labels.random_windows(boxes, classes, float(height), float(width), num_classes, rng, num_windows)
    return dict(image=image, groundtruth_boxes=norm, groundtruth_classes=onehot,
                groundtruth_closeness=aux_labels.closeness_labels(boxes, classes, height, width, num_classes),
                window_boxes=wb, window_classes=wl,
                groundtruth_edgemask=aux_labels.edgemask(boxes, float(height), float(width)))


def make_batch(seed, batch_size, height, width, num_classes, max_boxes=8, num_windows=64):
    rng = np.random.default_rng(seed)
    return [make_example(rng, height, width, num_classes, max_boxes, num_windows) for _ in range(batch_size)]


def make_sampler_keys(seed, batch_size, num_anchors, num_proposals):
    rng = np.random.default_rng(seed)
    return (rng.random((batch_size, num_anchors)).astype(np.float32),
            rng.random((batch_size, num_proposals)).astype(np.float32))
