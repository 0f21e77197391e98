"""Checkpoint interop of a built model with TensorFlow checkpoints: variable-name mapping of
/root/reference/object_detection/meta_architectures/faster_rcnn_meta_arch.py:1947-2013 (`restore_map`), extended to the
batch-norm statistics slim stores next to every conv (gamma / beta / moving_mean / moving_variance), on top of the
tensor-bundle reader of utils/tf_checkpoint.py.

  from_detection_checkpoint=True   names are the model's own TF variable names (the dead stage-1 block4 copy lives under
                                   the first-stage scope, as in the reference graph)
  from_detection_checkpoint=False  ImageNet classification checkpoint: the stage scope is stripped, so the three block4
                                   copies (second stage, closeness, window) and the dead one all initialise from the
                                   same `resnet_v1_101/block4/...` keys (traps T5, T14; trainer.py:341-348)"""
import numpy as np

from . import tf_checkpoint


def _tf_name(name):
    """Model variable name -> the name the reference's DETECTION graph gives that variable.
    * the dead stage-1 copy of block4 lives directly under the first-stage scope there;
    * Inception-ResNet-v2 (trap T18): the second-stage tails call `slim.repeat(net, 9, block8)` inside a fresh
      `InceptionResnetV2` variable scope (incres fe:133-170), so TF names that scope `Repeat`, while the same layers are
      `Repeat_2` in the classification network -- which is why the reference overrides
      `restore_from_classification_checkpoint_fn` (incres fe:173-248).  This framework keeps the classification name
      internally (the ImageNet map then needs no special case) and translates for detection checkpoints here."""
    name = name.replace("/_dead/", "/")
    if "/InceptionResnetV2/Repeat_2/" in name and not name.startswith("FirstStageFeatureExtractor/"):
        name = name.replace("/InceptionResnetV2/Repeat_2/", "/InceptionResnetV2/Repeat/")
    return name


def variable_name_map(model, from_detection_checkpoint=True):
    """{checkpoint variable name: [state-dict names of the model]} for weights, biases and batch-norm statistics."""
    st = model.param_store
    names = [p.name for p in st.params if "/_pad/" not in p.name]
    for b in st.bns:
        if b.scope is None:
            continue
        keys = (["gamma"] if b.has_gamma else []) + ["beta", "moving_mean", "moving_variance"]
        names += [b.scope + "/" + k for k in keys]
    out = {}
    if from_detection_checkpoint:
        for n in names:
            out.setdefault(_tf_name(n), []).append(n)
        return out
    arch = "/" + model._feature_extractor._architecture + "/"
    scopes = [model.first_stage_feature_extractor_scope + "/_dead/", model.first_stage_feature_extractor_scope + "/",
              model.second_stage_feature_extractor_scope + "/", model.closeness_box_predictor_scope + "/",
              model.window_box_predictor_scope + "/"]
    for n in names:
        for sc in scopes:
            if n.startswith(sc) and arch in "/" + n:
                out.setdefault(n[len(sc):], []).append(n)
                break
    return out


def load_tf_checkpoint(model, prefix, from_detection_checkpoint=True):
    """Initialise `model` from the TF checkpoint `prefix` (.index / .data-*).  Returns (number of model tensors set,
    checkpoint names the map expected but the file lacks)."""
    reader = tf_checkpoint.CheckpointReader(prefix)
    name_map = variable_name_map(model, from_detection_checkpoint)
    st = model.param_store
    shapes = {p.name: p.shape for p in st.params}
    kinds = {p.name: p.tf_kind for p in st.params if getattr(p, "tf_kind", None) is not None}
    sd, missing = tf_checkpoint.state_dict_from_checkpoint(reader, name_map, shapes, kinds)
    import torch
    st.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=False)
    return len(sd), missing


def save_tf_checkpoint(model, prefix, extra=None):
    """Write every variable of `model` (TF names, TF layouts) as a V2 checkpoint a `tf.train.Saver` can restore."""
    sd = model.param_store.state_dict()
    kinds = {p.name: getattr(p, "tf_kind", None) for p in model.param_store.params}
    tensors = {}
    for k, v in sd.items():
        if "/_pad/" in k:
            continue
        tensors[_tf_name(k)] = tf_checkpoint.native_to_tf(k, v.cpu().numpy().astype(np.float32), kinds.get(k))
    for k, v in (extra or {}).items():
        tensors[k] = np.asarray(v)
    tf_checkpoint.write_checkpoint(prefix, tensors)
    return sorted(tensors)


# ----------------------------------------------------------------------------- training state (resume)
def save_training_checkpoint(trainer, prefix):
    """Everything `tf.train.Saver` would write for the reference's training graph (trainer.py:431-447): the model
    variables, the MomentumOptimizer slot of every trainable variable under TF's slot name `<variable>/Momentum`, and
    `global_step` (int64) -- so that training resumes with the same update sequence."""
    model = trainer.model
    extra = {}
    for p in model.param_store.params:
        if "/_pad/" in p.name or not p.trainable:
            continue
        extra[_tf_name(p.name) + "/Momentum"] = tf_checkpoint.native_to_tf(
            p.name, p.m.detach().cpu().numpy().astype(np.float32), getattr(p, "tf_kind", None))
    extra["global_step"] = np.asarray(int(trainer.global_step), np.int64)
    return save_tf_checkpoint(model, prefix, extra)


def restore_training_checkpoint(trainer, prefix):
    """Inverse of save_training_checkpoint; a checkpoint without optimizer slots (e.g. written by `save_tf_checkpoint`
    or by a different optimizer) restores the variables and leaves the momenta at zero, as a fresh Saver restore of a
    fine-tune checkpoint would.  Returns (tensors set, momentum slots set, global_step)."""
    import torch
    model = trainer.model
    n, _ = load_tf_checkpoint(model, prefix, from_detection_checkpoint=True)
    reader = tf_checkpoint.CheckpointReader(prefix)
    slots = 0
    for p in model.param_store.params:
        key = _tf_name(p.name) + "/Momentum"
        if "/_pad/" in p.name or not p.trainable or not reader.has_tensor(key):
            continue
        v = tf_checkpoint.tf_to_native(p.name, reader.get_tensor(key), getattr(p, "tf_kind", None), p.shape).astype(np.float32)
        if tuple(v.shape) != tuple(p.shape):
            raise ValueError("%s: checkpoint shape %s, model shape %s" % (key, v.shape, tuple(p.shape)))
        p.m.copy_(torch.from_numpy(np.ascontiguousarray(v)).to(p.m.device))
        slots += 1
    if reader.has_tensor("global_step"):
        trainer.global_step = int(reader.get_tensor("global_step"))
    return n, slots, trainer.global_step


def latest_checkpoint(directory, basename="model.ckpt"):
    """`tf.train.latest_checkpoint` by file name: the `<basename>-<step>` prefix with the largest step, or None."""
    import os
    import re
    best, best_step = None, -1
    for f in os.listdir(directory):
        m = re.match(re.escape(basename) + r"-(\d+)\.index$", f)
        if m and int(m.group(1)) > best_step:
            best, best_step = os.path.join(directory, f[:-len(".index")]), int(m.group(1))
    return best
