"""Developer tool (2 GPUs, torchrun): one data-parallel step (bucketed all-reduce, head-bucket optimizer under the
trunk backward, CUDA graphs) must leave every rank with the weights a single replica gets from the summed gradient
of both shards (model_deploy.py:221-225, :296: task-loss gradients averaged, L2 once).
usage: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
from helpers import load_config
from mtl_ssl_b200.builders import model_builder
from mtl_ssl_b200.data import synthetic
from mtl_ssl_b200.trainer import Trainer
from mtl_ssl_b200.utils import synthetic_init

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
H, W, B = 600, 1000, 1      # model12.config resizes to min 600 / max 1024: use the native size
cfg = load_config("model12.config")


def build():
    m = model_builder.build(cfg.model, True, device="cuda:%d" % local, seed=0)
    synthetic_init.apply(m.param_store)
    return m


def batch(r, tr, nk):
    ex = synthetic.make_batch(500 + r, B, H, W, 20, max_boxes=6, num_windows=64)
    return tr.host_arrays(ex, synthetic.make_sampler_keys(600 + r, B, nk, 300))


for graph in (False, True):
    model = build()
    nk = model.num_kept_anchors((B, H, W, 3))
    tr = Trainer(model, cfg.train_config, H, W, B, gmax=16, use_cuda_graph=graph, world_size=world)
    w_init = model.param_store.w.clone()
    tr.step(batch(rank, tr, nk))
    if graph:       # the first graphed step also ran the warm-up passes: compare the second one
        w_init = model.param_store.w.clone(); m_init = model.param_store.m.clone()
        tr.step(batch(rank, tr, nk))
    torch.cuda.synchronize()
    got_w = model.param_store.w.clone()
    # every rank must hold identical weights
    ref = got_w.clone()
    dist.broadcast(ref, 0)
    assert torch.equal(ref, got_w), "ranks diverged"
    if rank == 0:
        ref_model = build()
        st = ref_model.param_store
        if graph:
            st.w.copy_(w_init); st.m.copy_(m_init); st.fold()
        rt = Trainer(ref_model, cfg.train_config, H, W, B, gmax=16, use_cuda_graph=False, world_size=1)
        rt.overlap_optimizer = False
        rt._hyper_host[0] = float(rt.lr_fn(tr.global_step - 1)); rt._hyper_host[1] = rt.momentum; rt._hyper_host[2] = rt.clip_norm
        st.hyper.copy_(rt._hyper_host)
        for r in range(world):
            rt._forward_backward(rt._bind(batch(r, rt, nk)))
        st.stats_and_reg_loss(1.0 / world)
        st.apply(1.0 / world)
        torch.cuda.synchronize()
        dw_ref = st.w - w_init
        dw_got = got_w - w_init
        err = float((dw_got - dw_ref).norm() / dw_ref.norm().clamp_min(1e-30))
        print("graph=%s  |dw| %.4e  relative error of the update vs single-replica reference: %.3e"
              % (graph, float(dw_ref.norm()), err), flush=True)
        assert err < 2e-3, err
    dist.barrier()
# the pipelined API (deferred head update: its weight gradients, the all-reduce of that bucket and its optimizer pass run
# under the NEXT step's trunk forward; the trunk bucket travels in two pieces): replicas must stay bit-identical, and the
# losses of every step must equal those of the synchronous API on the same batches
for graph in (False, True):
    seqs = []
    for mode in ("sync", "pipe"):
        model = build()
        nk = model.num_kept_anchors((B, H, W, 3))
        tr = Trainer(model, cfg.train_config, H, W, B, gmax=16, use_cuda_graph=graph, world_size=world)
        # a vanishing learning rate: with random-init weights thousands of anchors have nearly equal scores, and the
        # last-bit run-to-run noise of the fp32 atomics, scaled by a real learning rate, flips a proposal now and then
        # (observed: second_stage_localization_loss 0.0 vs 0.0075 at one of four steps) -- in ANY two runs, whatever
        # the API.  What is compared here is the plumbing: same batches, same order of updates, same optimizer state.
        tr.lr_fn = lambda step: 1e-8
        out = []
        for i in range(4):
            b = batch(10 * rank + i, tr, nk)
            r = tr.step(b) if mode == "sync" else tr.step_pipelined(b)
            if r is not None:
                out.append(r)
        if mode == "pipe":
            out.append(tr.flush())
        torch.cuda.synchronize()
        assert len(out) == 4 and not tr._heads_pending
        got_w = model.param_store.w.clone()
        ref = got_w.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, got_w), "ranks diverged (%s, graph=%s)" % (mode, graph)
        got_m = model.param_store.m.clone()
        ref = got_m.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, got_m), "momenta diverged (%s, graph=%s)" % (mode, graph)
        seqs.append((out, got_w, got_m))
    for a, b in zip(seqs[0][0], seqs[1][0]):
        for k in a:
            assert abs(a[k] - b[k]) <= 1e-4 * max(1.0, abs(a[k])), (k, a[k], b[k])
    rel = float((seqs[0][1] - seqs[1][1]).norm() / seqs[0][1].norm())
    relm = float((seqs[0][2] - seqs[1][2]).norm() / seqs[0][2].norm())
    if rank == 0:
        print("graph=%s  pipelined vs synchronous API over 4 steps: losses equal, weights differ by %.2e, momenta "
              "(= the accumulated clipped gradients of all four steps) by %.2e (relative)" % (graph, rel, relm), flush=True)
    assert rel < 1e-6 and relm < 2e-3
    dist.barrier()
if rank == 0:
    print("dp_check ok")
dist.destroy_process_group()
