"""Developer tool (GPU): per-shape table of the tcgen05 conv launches of one full-size training step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load_config
from mtl_ssl_b200 import ops_conv
from mtl_ssl_b200.builders import model_builder
from mtl_ssl_b200.data import synthetic
from mtl_ssl_b200.trainer import Trainer
from mtl_ssl_b200.utils import synthetic_init

H, W, B = 600, 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = load_config("model12.config")
model = model_builder.build(cfg.model, True, device="cuda", seed=0)
synthetic_init.apply(model.param_store)
nk = model.num_kept_anchors((B, H, W, 3))
tr = Trainer(model, cfg.train_config, H, W, B, gmax=16, use_cuda_graph=False)
ex = synthetic.make_batch(1, B, H, W, 20)
arrays = tr.host_arrays(ex, synthetic.make_sampler_keys(2, B, nk, 300))
for _ in range(2):
    tr.step(arrays)
image = tr._bind(arrays)
from mtl_ssl_b200.nets.layers import Concurrency
Concurrency.enabled = False
tr._forward_backward(image)
torch.cuda.synchronize()
ops_conv.PROFILE = []
torch.cuda._sleep(int(60e-3 * 1.9e9))
tr._forward_backward(image)
torch.cuda.synchronize()
prof, ops_conv.PROFILE = ops_conv.PROFILE, None
tab = {}
for m, f, a, b, g in prof:
    k = (("fprop", "dgrad", "wgrad")[m],) + g
    d = tab.setdefault(k, [0, 0.0, 0.0])
    d[0] += 1; d[1] += a.elapsed_time(b); d[2] += f
rows = sorted(tab.items(), key=lambda kv: -kv[1][1])
tot = sum(v[1] for v in tab.values())
print("total conv ms %.3f  flops %.3f T" % (tot, sum(v[2] for v in tab.values()) / 1e12))
print("%-6s %5s %4s %4s %5s %5s %2s %2s | %4s %8s %8s %6s" % ("mode", "N", "H", "W", "C", "K", "R", "s", "cnt", "ms", "TF/s", "share"))
for k, v in rows[:60]:
    print("%-6s %5d %4d %4d %5d %5d %2d %2d | %4d %8.3f %8.1f %5.1f%%" % (k[0], k[1], k[2], k[3], k[4], k[5], k[6], k[7], v[0], v[1], v[2] / v[1] / 1e9, 100 * v[1] / tot))
