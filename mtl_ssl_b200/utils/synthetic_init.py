"""Deterministic, well-conditioned frozen batch-norm statistics for random-initialised models
(benchmarks / smoke runs; there is no network for the ImageNet checkpoint the reference fine-tunes
from, trainer.py:311-356).  With He-initialised convs and identity batch norm the residual stack
doubles its variance every unit and overflows; these statistics keep activations O(1):
  stem     moving_variance ~ Var(pixel - mean) * He gain  (pixels are uniform 0..255)
  conv3    gamma 0.25 (residual branch)          shortcut  gamma 0.7
"""
import torch


def apply_host(store):
    """The statistics only, on the host-side BatchNorm records (no device needed)."""
    for b in store.bns:
        if b.scope is None:             # virtual per-channel scale (Inception-ResNet residual scaling): not a batch norm
            continue
        c = b.channels
        stem = "/block" not in b.scope and b.scope.endswith("/conv1/BatchNorm")
        if "/conv3/" in b.scope:
            g = 0.25
        elif "/shortcut/" in b.scope:
            g = 0.7
        else:
            g = 1.0
        b.gamma = torch.full((c,), g)
        b.beta = torch.zeros(c)
        b.mean = torch.zeros(c)
        b.var = torch.full((c,), 14000.0 if stem else 1.0)


def apply(store):
    apply_host(store)
    store._upload_bn_inplace()
    store.fold()
    for h in store.post_load_hooks:
        h()
