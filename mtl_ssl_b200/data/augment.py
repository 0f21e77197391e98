"""Training-time augmentation and batching of the input side (second slice), host NumPy like the
reference's graph ops: `random_horizontal_flip` of /root/reference/object_detection/core/preprocessor.py:239-345
including the fork's window boxes and edge mask, and the fixed-size batching of core/batcher.py.

Parity trap T16 (reproduced, switchable): the reference flips the edge mask with `tf.image.flip_left_right` applied
to the rank-3 tensor [2, h, w] (preprocessor.py:338-340).  That op reads a rank-3 input as [height, width, channels],
so it reverses axis 1 -- the mask's ROWS -- while the image and all boxes are mirrored left-right.  With
`reference_edgemask_axis=True` (default) this module does the same; False mirrors the mask columns."""
import numpy as np


def flip_boxes(boxes):
    """preprocessor.py `flip_boxes`: [ymin, xmin, ymax, xmax] normalised -> mirrored around x = 0.5."""
    b = np.asarray(boxes, np.float32).reshape(-1, 4)
    return np.stack([b[:, 0], np.float32(1.0) - b[:, 3], b[:, 2], np.float32(1.0) - b[:, 1]], 1).astype(np.float32)


def horizontal_flip(example, reference_edgemask_axis=True):
    """The `do_a_flip` branch of random_horizontal_flip on one example dict (data/synthetic.py format)."""
    out = dict(example)
    out["image"] = np.ascontiguousarray(np.asarray(example["image"])[:, ::-1, :])
    out["groundtruth_boxes"] = flip_boxes(example["groundtruth_boxes"])
    if example.get("window_boxes") is not None:
        out["window_boxes"] = flip_boxes(example["window_boxes"])
    if example.get("groundtruth_edgemask") is not None:
        em = np.asarray(example["groundtruth_edgemask"])
        out["groundtruth_edgemask"] = np.ascontiguousarray(em[:, ::-1, :] if reference_edgemask_axis else em[:, :, ::-1])
    return out


def random_horizontal_flip(example, rng, reference_edgemask_axis=True):
    """Flip with probability 0.5, never when there are no boxes (preprocessor.py:298-300: `size(boxes) > 0 and
    uniform() > 0.5`).  `rng`: numpy Generator (the reference draws from the graph-level TF seed)."""
    u = rng.random()
    if np.asarray(example["groundtruth_boxes"]).size > 0 and u > 0.5:
        return horizontal_flip(example, reference_edgemask_axis)
    return example


def batches(examples, batch_size, drop_remainder=True):
    """core/batcher.py BatchQueue semantics for static shapes: consecutive groups of `batch_size` examples."""
    group = []
    for e in examples:
        group.append(e)
        if len(group) == batch_size:
            yield group
            group = []
    if group and not drop_remainder:
        yield group


def input_examples(train_input_reader, num_classes, options=(), seed=0, epochs=None, start_epoch=0):
    """Examples of the pipeline config's `train_input_reader { tf_record_input_reader { input_path: ... } }`, shuffled
    per file when `shuffle` is set, augmented by the `data_augmentation_options` named in `options`
    (only random_horizontal_flip is built), repeated `epochs` times (None = forever)."""
    from .tfrecord import TfRecordDataset
    import glob
    ip = train_input_reader.tf_record_input_reader.input_path      # `optional string` (input_reader.proto:53): a path / glob
    paths = []
    for pat in ([ip] if isinstance(ip, str) else list(ip)):
        paths.extend(sorted(glob.glob(pat)) or [pat])
    unknown = [o for o in options if o != "random_horizontal_flip"]
    if unknown:
        raise NotImplementedError("data augmentation options not built: %s" % unknown)
    # `start_epoch`: a resumed run passes global_step * batch // records-per-epoch so that record order and flips go on
    # from where they were instead of replaying epoch 0 (the augmentation generator is seeded by (seed, start_epoch))
    rng = np.random.default_rng([seed, start_epoch])
    epoch = start_epoch
    while epochs is None or epoch < start_epoch + epochs:
        ds = TfRecordDataset(paths, num_classes, shuffle=bool(train_input_reader.shuffle),      # proto default: true
                             seed=seed + epoch)
        for e in ds:
            for o in options:
                e = random_horizontal_flip(e, rng)
            yield e
        epoch += 1
