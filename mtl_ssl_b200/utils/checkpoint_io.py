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
    return name.replace("/_dead/", "/")


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
    sd, missing = tf_checkpoint.state_dict_from_checkpoint(reader, name_map, shapes)
    import torch
    st.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd.items()}, strict=False)
    return len(sd), missing


def save_tf_checkpoint(model, prefix, extra=None):
    """Write every variable of `model` (TF names, TF layouts) as a V2 checkpoint a `tf.train.Saver` can restore."""
    sd = model.param_store.state_dict()
    tensors = {}
    for k, v in sd.items():
        if "/_pad/" in k:
            continue
        tensors[_tf_name(k)] = tf_checkpoint.native_to_tf(k, v.cpu().numpy().astype(np.float32))
    for k, v in (extra or {}).items():
        tensors[k] = np.asarray(v)
    tf_checkpoint.write_checkpoint(prefix, tensors)
    return sorted(tensors)
