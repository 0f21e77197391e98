"""Developer tool (GPU): per-phase device time of one training step (single stream, no host gaps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load_config
from mtl_ssl_b200.builders import model_builder
from mtl_ssl_b200.data import synthetic
from mtl_ssl_b200.trainer import Trainer
from mtl_ssl_b200.utils import synthetic_init
from mtl_ssl_b200.nets.layers import Concurrency

H, W, B = 600, 1000, int(sys.argv[1]) if len(sys.argv) > 1 else 1
conc = len(sys.argv) > 2 and sys.argv[2] == "conc"
cfg = load_config("model12.config")
model = model_builder.build(cfg.model, True, device="cuda", seed=0)
synthetic_init.apply(model.param_store)
nk = model.num_kept_anchors((B, H, W, 3))
tr = Trainer(model, cfg.train_config, H, W, B, gmax=16, use_cuda_graph=False)
ex = synthetic.make_batch(1, B, H, W, 20)
arrays = tr.host_arrays(ex, synthetic.make_sampler_keys(2, B, nk, 300))
Concurrency.enabled = conc
for _ in range(2):
    tr.step(arrays)
image = tr._bind(arrays)
m = model
marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
torch.cuda.synchronize()
torch.cuda._sleep(int(80e-3 * 1.9e9))
mark("start")
fe = m._feature_extractor
# marks inside predict(): after the trunk, after the RPN head, after the proposal chain
_epf, _ppr, _rpnf = fe.extract_proposal_features, m._postprocess_rpn, m._rpn_conv.fwd
def epf(*a, **k):
    r = _epf(*a, **k); mark("  trunk forward (stem + block1-3)"); return r
def ppr(*a, **k):
    mark("  rpn conv + heads"); r = _ppr(*a, **k); mark("  proposals (decode, sort, nms, sample)"); return r
fe.extract_proposal_features, m._postprocess_rpn = epf, ppr
# replicate predict() with marks
pd = m.predict(m.preprocess(image)); mark("  crops + main tail + closeness tail launch")
fe.extract_proposal_features, m._postprocess_rpn = _epf, _ppr
pd = m.predict_with_window(pd); mark("predict_with_window")
pd = m.predict_edgemask(pd); mark("predict_edgemask")
pd = m.predict_with_mtl_results(pd); mark("predict_with_mtl_results (refine 1280 ROIs)")
m.loss(pd); mark("loss")
m.backward(pd, part="heads"); mark("backward heads")
m.backward(pd, part="trunk"); mark("backward rpn+trunk")
tr._optimize(); mark("optimizer")
torch.cuda.synchronize()
tot = marks[0][1].elapsed_time(marks[-1][1])
for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
    print("%-60s %8.3f ms" % (n1, e0.elapsed_time(e1)))
print("%-60s %8.3f ms  (concurrency %s)" % ("total", tot, conc))
