"""Developer tool (GPU): bisect of tests/test_gpu_zz_shape_buckets.py -- prints, per step, every loss of the plain
single-shape trainers and of ShapeBucketTrainer under several schedules, so the diverging step / loss / feature shows."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import test_gpu_train_step as T  # noqa: E402
from helpers import load_config, randomize_bn  # noqa: E402
from mtl_ssl_b200.builders import model_builder  # noqa: E402
from mtl_ssl_b200.data import synthetic  # noqa: E402
from mtl_ssl_b200.nets.layers import Concurrency  # noqa: E402
from mtl_ssl_b200.shape_buckets import ShapeBucketTrainer  # noqa: E402
from mtl_ssl_b200.trainer import Trainer  # noqa: E402

cfg = load_config("model12.config", T.SMALL)
K, M = cfg.model.faster_rcnn.num_classes, cfg.model.faster_rcnn.first_stage_max_proposals
shapes = [(224, 320), (224, 288), (224, 320), (224, 288)]
kw = dict(gmax=8, learning_rate=1e-5)


RAW = "--raw-init" in sys.argv      # the ill-conditioned start (identity batch norm) of the round-1 test


def build():
    model = model_builder.build(cfg.model, True, device="cuda", seed=0)
    if not RAW:
        model.param_store.load_state_dict(randomize_bn(model.param_store.state_dict(), 0))
    return model


def batch(model, i, hw):
    ex = synthetic.make_batch(60 + i, 1, hw[0], hw[1], K, max_boxes=4, num_windows=16)
    ky = synthetic.make_sampler_keys(70 + i, 1, model.num_kept_anchors((1, hw[0], hw[1], 3)), M)
    return ex, ky


def plain(conc=True, overlap=True, prefix=True):
    Concurrency.enabled = conc
    want, state = [], None
    for i, hw in enumerate(shapes):
        model = build()
        st = model.param_store
        if state is not None:
            st.w.copy_(state[0]); st.m.copy_(state[1])
            st.fold()
        tr = Trainer(model, None, hw[0], hw[1], 1, use_cuda_graph=False, **kw)
        tr.overlap_optimizer, tr.pipeline_prefix = overlap, prefix
        tr.global_step = i
        ex, ky = batch(model, i, hw)
        want.append(tr.step(tr.host_arrays(ex, ky)))
        torch.cuda.synchronize()
        state = (st.w.clone(), st.m.clone())
    Concurrency.enabled = True
    return want


def buckets(graph=False, conc=True, overlap=True, prefix=True, pipelined=True, sync_each=False):
    Concurrency.enabled = conc
    model = build()

    class Tr(Trainer):
        def __init__(self, *a, **k):
            Trainer.__init__(self, *a, **k)
            self.overlap_optimizer, self.pipeline_prefix = overlap, prefix

    bt = ShapeBucketTrainer(model, None, batch_size=1, max_buckets=4, use_cuda_graph=graph, trainer_cls=Tr, **kw)
    got = []
    for i, hw in enumerate(shapes):
        ex, ky = batch(model, i, hw)
        if pipelined:
            r = bt.step_pipelined(bt.host_arrays(ex, ky))
            if r is not None:
                got.append(r)
        else:
            got.append(bt.step(bt.host_arrays(ex, ky)))
        if sync_each:
            torch.cuda.synchronize()
    if pipelined:
        got.append(bt.flush())
    torch.cuda.synchronize()
    Concurrency.enabled = True
    return got


def show(name, want, got):
    print("==== %s" % name)
    for i, (a, b) in enumerate(zip(want, got)):
        worst = max(a, key=lambda k: abs(a[k] - b[k]) / max(1.0, abs(a[k])))
        bad = [k for k in a if abs(a[k] - b[k]) > 2e-3 * max(1.0, abs(a[k]))]
        print("step %d shape %s worst %s want %.6f got %.6f bad=%s" % (i, shapes[i], worst, a[worst], b[worst], bad))
        if bad:
            for k in a:
                print("      %-40s %.6f %.6f" % (k, a[k], b[k]))
    sys.stdout.flush()


want = plain()
show("plain vs plain (run-to-run noise)", want, plain())
show("plain no-concurrency/no-overlap/no-prefix vs plain", want, plain(False, False, False))
show("buckets eager pipelined", want, buckets())
show("buckets eager pipelined, sync after each call", want, buckets(sync_each=True))
show("buckets eager step()", want, buckets(pipelined=False))
show("buckets eager no-concurrency", want, buckets(conc=False))
show("buckets eager no-overlap-optimizer", want, buckets(overlap=False))
show("buckets eager no-prefix-pipelining", want, buckets(prefix=False))
show("buckets eager step() no-conc no-overlap no-prefix", want, buckets(conc=False, overlap=False, prefix=False,
                                                                         pipelined=False))
show("buckets graph pipelined", want, buckets(graph=True))
show("buckets graph step()", want, buckets(graph=True, pipelined=False))
