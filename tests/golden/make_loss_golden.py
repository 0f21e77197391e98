"""Generates tests/golden/loss_reference.npz by RUNNING the reference's own loss classes
(/root/reference/object_detection/core/losses.py: `Loss.__call__` with its rank-mismatch flattening,
WeightedSmoothL1LocalizationLoss with the fork's sigma, WeightedSoftmaxClassificationLoss / _v2,
WeightedSigmoidClassificationLoss) on a NumPy stand-in for the handful of TensorFlow ops they use.

What this pins: the reference's CONTROL FLOW and tensor plumbing (reshapes, weights, reductions, which rows are
flattened).  What it cannot pin: the TensorFlow kernels themselves -- `tf.nn.softmax_cross_entropy_with_logits[_v2]` and
`tf.nn.sigmoid_cross_entropy_with_logits` are restated here from their documented formulas
(-sum(labels * log_softmax(logits)); max(x,0) - x*z + log(1+exp(-|x|))).
Run from the repo root:  python tests/golden/make_loss_golden.py"""
import os
import sys
import types

import numpy as np


class TT(np.ndarray):
    """ndarray with the two TensorShape accessors losses.py touches."""

    def get_shape(self):
        shape = self.shape
        return types.SimpleNamespace(as_list=lambda: list(shape))


def t(a, dtype=None):
    return np.asarray(a, dtype=dtype).view(TT)


def _wrap(fn):
    return lambda *a, **k: t(fn(*a, **k))


class _Scope(object):
    def __enter__(self):
        return "scope"

    def __exit__(self, *a):
        return False


def _log_softmax(x):
    m = x.max(-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))


tf = types.ModuleType("tensorflow")
tf.float32 = np.float32
tf.name_scope = lambda *a, **k: _Scope()
tf.contrib = types.SimpleNamespace(slim=None)
tf.reduce_sum = _wrap(lambda x, axis=None, **k: np.sum(x, axis=axis))
tf.reduce_mean = _wrap(lambda x, axis=None, **k: np.mean(x, axis=axis))
tf.reshape = _wrap(lambda x, shape: np.reshape(x, [int(s) for s in np.asarray(shape).reshape(-1)]))
tf.shape = lambda x: np.asarray(np.shape(x))
tf.expand_dims = _wrap(lambda x, axis: np.expand_dims(x, axis))
tf.where = _wrap(lambda c, a, b: np.where(c, a, b))
tf.is_nan = _wrap(np.isnan)
tf.abs = _wrap(np.abs)
tf.square = _wrap(np.square)
tf.less = _wrap(np.less)
tf.ones = _wrap(lambda shape, dtype=np.float32: np.ones([int(s) for s in np.asarray(shape).reshape(-1)], dtype))
tf.to_float = _wrap(lambda x: np.asarray(x, np.float32))
tf.sigmoid = _wrap(lambda x: 1.0 / (1.0 + np.exp(-x)))
tf.nn = types.SimpleNamespace(
    softmax_cross_entropy_with_logits=_wrap(lambda labels, logits: -(labels * _log_softmax(logits)).sum(-1)),
    softmax_cross_entropy_with_logits_v2=_wrap(lambda labels, logits: -(labels * _log_softmax(logits)).sum(-1)),
    sigmoid_cross_entropy_with_logits=_wrap(lambda labels, logits: np.maximum(logits, 0) - logits * labels +
                                            np.log1p(np.exp(-np.abs(logits)))))
sys.modules["tensorflow"] = tf
for name in ("object_detection.core.box_list", "object_detection.core.box_list_ops", "object_detection.utils.ops"):
    sys.modules[name] = types.ModuleType(name)
sys.path.insert(0, "/root/reference")
from object_detection.core import losses as ref      # noqa: E402


def main():
    rng = np.random.default_rng(17)
    out = {}
    B, N, K1 = 2, 64, 21
    # --- the window-class call of fmA:1839-1858: rank-2 logits [B*N, K+1] against rank-3 soft labels [B, N, K+1]
    logits = rng.normal(0, 2, (B * N, K1)).astype(np.float32)
    soft = rng.random((B, N, K1)).astype(np.float32)
    soft /= soft.sum(-1, keepdims=True)
    v2 = ref.WeightedSoftmaxClassificationLoss_v2(anchorwise_output=True)
    got = v2(t(logits), t(soft))
    out["win/logits"], out["win/labels"], out["win/loss"] = logits, soft, np.asarray(got)
    # --- same ranks: no flattening; weights applied per anchor; scalar reduction when not anchorwise
    logits3 = rng.normal(0, 2, (B, N, K1)).astype(np.float32)
    onehot = np.eye(K1, dtype=np.float32)[rng.integers(0, K1, (B, N))]
    w = (rng.random((B, N)) < 0.7).astype(np.float32)
    out["cls/logits"], out["cls/labels"], out["cls/weights"] = logits3, onehot, w
    out["cls/anchorwise"] = np.asarray(ref.WeightedSoftmaxClassificationLoss(True)(t(logits3), t(onehot), weights=t(w)))
    out["cls/scalar"] = np.asarray(ref.WeightedSoftmaxClassificationLoss(False)(t(logits3), t(onehot), weights=t(w)))
    out["cls/v2_scalar"] = np.asarray(ref.WeightedSoftmaxClassificationLoss_v2(False)(t(logits3), t(onehot), weights=t(w)))
    # --- smooth L1 with the fork's sigma (3 for the RPN, 1 for the second stage: fmA:392, :412-413)
    pred = rng.normal(0, 1, (B, N, 4)).astype(np.float32)
    tgt = rng.normal(0, 1, (B, N, 4)).astype(np.float32)
    out["loc/pred"], out["loc/target"], out["loc/weights"] = pred, tgt, w
    for sigma in (1.0, 3.0):
        out["loc/sigma%d" % sigma] = np.asarray(
            ref.WeightedSmoothL1LocalizationLoss(anchorwise_output=True, sigma=sigma)(t(pred), t(tgt), weights=t(w)))
    out["loc/scalar_sigma3"] = np.asarray(ref.WeightedSmoothL1LocalizationLoss(False, 3.0)(t(pred), t(tgt), weights=t(w)))
    # --- nan targets replaced by the prediction (ignore_nan_targets)
    tgt_nan = tgt.copy()
    tgt_nan[0, :5, 1] = np.nan
    out["loc/target_nan"] = tgt_nan
    out["loc/ignore_nan"] = np.asarray(
        ref.WeightedSmoothL1LocalizationLoss(True, 1.0)(t(pred), t(tgt_nan), ignore_nan_targets=True, weights=t(w)))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "loss_reference.npz"), **out)
    print("wrote loss_reference.npz:", {k: v.shape for k, v in out.items() if "/loss" in k or "anchorwise" in k})


if __name__ == "__main__":
    main()
