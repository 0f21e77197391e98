"""Developer script (GPU): runs one step on device and oracle and prints per-tensor differences."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from test_gpu_train_step import _setup, SMALL, MOBILE
from helpers import oracle_config
from oracle.model import Oracle
from oracle import nn as ON

name = sys.argv[1] if len(sys.argv) > 1 else "model12.config"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
H, W = 224, 320
cfg, model, sd, examples, keys, tr = _setup(name, MOBILE if name[5] in "56" else SMALL, H, W, B)
arrays = tr.host_arrays(examples, keys)
image = tr._bind(arrays)
pd = tr._forward_backward(image)
torch.cuda.synchronize()
orc = Oracle({k: v for k, v in sd.items() if "/_pad/" not in k}, oracle_config(cfg), bf16=True)
orc.require_grad([p.name for p in model.param_store.params if p.trainable and "/_pad/" not in p.name and "/_dead/" not in p.name])
prop_in = (pd["rpn_box_encodings"].cpu().numpy(), pd["rpn_objectness_predictions_with_background"].cpu().numpy())
out = orc.forward(torch.from_numpy(arrays["image"]), examples, keys, H, W, proposal_inputs=prop_in)
want = orc.loss(out, examples, keys, H, W)


def cmp(tag, a, b):
    a = a.detach().float().cpu(); b = b.detach().float().cpu()
    d = (a - b).abs()
    print("%-28s shape %-22s max|a| %.4g  max diff %.4g  mean diff %.4g" % (tag, tuple(a.shape), a.abs().max(), d.max(), d.mean()))


cmp("feat", pd["rpn_features_to_crop"], out["feat"])
K = cfg.model.faster_rcnn.num_classes
cmp("rpn_box", pd["rpn_box_encodings"], out["rpn_box"])
cmp("rpn_cls", pd["rpn_objectness_predictions_with_background"], out["rpn_cls"])
cmp("prop_abs", pd["proposal_boxes"], torch.from_numpy(out["prop_abs"]))
cmp("refined_box_encodings", pd["refined_box_encodings"], out["refined_box_encodings"])
cmp("class_predictions", pd["class_predictions_with_background"], out["class_predictions_with_background"])
if "closeness_predictions" in out:
    cmp("closeness", pd["closeness_predictions"], out["closeness_predictions"])
if "window_class_predictions" in out:
    cmp("window", pd["window_class_predictions"], out["window_class_predictions"])
if "edgemask_predictions" in out:
    cmp("edgemask", pd["edgemask_predictions"], out["edgemask_predictions"])
if "refine_in" in out:
    cmp("refine_in", pd["_refine_in"], out["refine_in"])
    cmp("refined", pd["mtl_refined_class_predictions_with_background"], out["mtl_refined_class_predictions_with_background"])
from mtl_ssl_b200.meta_architectures.faster_rcnn_meta_arch import LOSS_KEYS
got = dict(zip(LOSS_KEYS, model.workspace.bufs["loss/values"].cpu().tolist()))
for k, v in want.items():
    print("%-36s gpu %.6f oracle %.6f diff %.2e" % (k, got[k], float(v), got[k] - float(v)))
sum(want.values()).backward()
st = model.param_store
rows = []
for p in st.params:
    if not p.trainable or "/_pad/" in p.name or "/_dead/" in p.name:
        continue
    g = p.g.float().cpu().reshape(-1)
    w = orc.p[p.name].grad
    w = torch.zeros_like(g) if w is None else w.reshape(-1)
    ng, nw = g.norm().item(), w.norm().item()
    cos = float((g @ w) / max(ng * nw, 1e-30))
    rows.append((cos, p.name, ng, nw))
rows.sort()
for r in rows[:25]:
    print("cos %.4f  %-90s |g| %.4g |w| %.4g" % r)
print("params compared", len(rows), "min cos", rows[0][0], "median", rows[len(rows) // 2][0])
