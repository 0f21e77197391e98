"""Oracle (TEST INFRASTRUCTURE): random-initialised parameter dictionary for oracle/model.py.

The variable table (names, shapes, initialisers, L2 weights) is read from the host-side model
description built with device=None (no CUDA involved); values are drawn here on the CPU."""
import torch


def _table(ocfg):
    import os, sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "..", "tests"))
    from helpers import load_config
    from mtl_ssl_b200.builders import model_builder
    cfg = load_config("model12.config", (("second_stage_batch_size: 256",
                                          "second_stage_batch_size: %d" % ocfg["second_stage_batch_size"]),))
    return model_builder.build(cfg.model, True, device=None)


def build_params(ocfg, seed=0):
    from mtl_ssl_b200.runtime import _init_tensor
    model = _table(ocfg)
    st = model.param_store
    gen = torch.Generator().manual_seed(seed)
    params, l2 = {}, {}
    for p in st.params:
        if "/_pad/" in p.name or "/_dead/" in p.name:
            continue
        params[p.name] = _init_tensor(p, gen).float()
        l2[p.name] = p.l2
    for b in st.bns:
        if "/_dead/" in b.scope:
            continue
        c = b.channels
        stem = "/block" not in b.scope and b.scope.endswith("/conv1/BatchNorm")
        g = 0.25 if "/conv3/" in b.scope else (0.7 if "/shortcut/" in b.scope else 1.0)
        params[b.scope + "/gamma"] = torch.full((c,), g)
        params[b.scope + "/beta"] = torch.zeros(c)
        params[b.scope + "/moving_mean"] = torch.zeros(c)
        params[b.scope + "/moving_variance"] = torch.full((c,), 14000.0 if stem else 1.0)
    build_params.trainable = {p.name for p in st.params if p.trainable}
    return params, l2


def is_trainable(name):
    return name in getattr(build_params, "trainable", ())


def num_kept_anchors(ocfg, H, W):
    from . import boxes as OB
    hf = wf = None
    h, w = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1        # conv1 7x7/2 explicit pad
    for _ in range(3):                                        # pool1, block1, block2 strides
        h, w = -(-h // 2), -(-w // 2)
    a = OB.grid_anchors(h, w, ocfg["scales"], ocfg["aspect_ratios"])
    return len(OB.prune_outside_window(a, (0, 0, H, W))[1])
