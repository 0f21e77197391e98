#!/usr/bin/env python
"""Benchmark of the hot path: images/sec of one full Faster R-CNN ResNet-101 (+3 aux heads + refiner)
training step (forward, 8 losses, explicit backward, gradient all-reduce, clip + momentum update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): tests/golden/configs/model12.config unchanged (ResNet-101, K=20,
12 anchors/location, 300 proposals, 256+256+64 trained ROIs + 1280 forward-only refine ROIs),
synthetic 600x1000 images, per-GPU batch = the config's train_config.batch_size (1), random-init
weights, bf16 tensor-core convolutions with fp32 accumulation and fp32 master weights.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA-graph replays, inputs in
HBM); `e2e` = the same step through the public API (Trainer.step) with pinned-host inputs copied
H2D and the loss vector read back D2H every step.  `roofline` is measured live: every tcgen05 conv
launch of one eager step is bracketed by CUDA events; achieved = algorithmic FLOPs / kernel time.
`--impl reference` times the CPU restatement of the reference path (oracle/, torch fp32 on all host
threads): the reference itself is Python-2/TF-1.7 graph code and cannot run in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIG = os.path.join(ROOT, "tests", "golden", "configs", "model12.config")
WORKLOAD = ("Faster R-CNN ResNet-101 + 3 aux heads + refiner, model12.config, 600x1000, "
            "train step (fwd+bwd+allreduce+clip+momentum)")
H, W = 600, 1000


def algorithmic_flops_per_image():
    """2*MACs of every conv/FC of one train step at 600x1000 (fwd 2.90 T + bwd 2.02 T)."""
    return 4.92e12


def dominant_conv_group(records, peak):
    """records: (mode, algorithmic flops, milliseconds, (N,H,W,C,K,R,stride,P,Q)) per tcgen05 launch of one step.
    Groups launches of identical mode + geometry and returns the roofline entry of the group with the largest total
    time (the step's dominant kernel): achieved = its algorithmic FLOPs / its summed launch durations."""
    groups = {}
    for mode, flops, ms, geom in records:
        g = groups.setdefault((int(mode), tuple(int(v) for v in geom)), [0.0, 0.0, 0])
        g[0] += float(flops); g[1] += float(ms); g[2] += 1
    if not groups:
        return None
    (mode, geom), (flops, ms, n) = max(groups.items(), key=lambda kv: kv[1][1])
    if ms <= 0:
        return None
    N, H_, W_, C, K, R, stride, P, Q = geom
    achieved = flops / (ms * 1e-3) / 1e12
    total_ms = sum(g[1] for g in groups.values())
    name = "tc_gemm_kernel %s %dx%d/%d C%d->K%d on %dx%dx%d" % (("fprop", "dgrad", "wgrad")[mode], R, R, stride, C, K, N,
                                                                H_, W_)
    out = {"kernel": name, "launches": n, "avg_us": ms * 1e3 / n, "share_of_conv_time": ms / total_ms,
           "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
           "traffic": None}
    try:        # DRAM bytes per launch of this kernel from the committed ncu --set full capture, when there is one
        t = json.load(open(os.path.join(ROOT, "profiles", "r1_kernel_traffic.json"))).get(name)
        if t:
            out["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"]
            out["traffic_source"] = "profiles/r1_kernel_traffic.json (%s)" % t.get("capture", "ncu --set full")
    except Exception:
        pass
    return out


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def oracle_sample(threads, rois=16, refine=True, full_image=True):
    """One bounded CPU sample of the reference path (oracle/model.py, fp32): the COMPLETE stage-1
    network + RPN + proposal path at 600x1000, but `rois` ROIs per second-stage head instead of
    256 (and rois//4 windows).  Returns (seconds_stage1_part, seconds_per_roi_equivalent, detail)."""
    import numpy as np
    import torch
    from helpers import load_config, oracle_config
    from mtl_ssl_b200.data import synthetic
    from oracle.model import Oracle
    import oracle.refparams as refparams
    torch.set_num_threads(threads)
    cfg = load_config("model12.config", (("second_stage_batch_size: 256", "second_stage_batch_size: %d" % rois),))
    ocfg = oracle_config(cfg)
    params, l2 = refparams.build_params(ocfg, seed=0)
    nwin = max(rois // 4, 1)
    examples = synthetic.make_batch(1, 1, H, W, ocfg["num_classes"], max_boxes=8, num_windows=nwin)
    orc = Oracle(params, ocfg, bf16=False)
    orc.require_grad([k for k in params if refparams.is_trainable(k)])
    img = torch.from_numpy(np.stack([e["image"] for e in examples]))
    nk = refparams.num_kept_anchors(ocfg, H, W)
    keys = synthetic.make_sampler_keys(2, 1, nk, ocfg["first_stage_max_proposals"])
    t0 = time.perf_counter()
    out = orc.forward(img, examples, keys, H, W)
    losses = orc.loss(out, examples, keys, H, W)
    total = sum(losses.values()) + orc.regularization_loss(l2)
    total.backward()
    dt = time.perf_counter() - t0
    # ROI-equivalents processed: main + closeness (fwd+bwd), windows (fwd+bwd), 5x refine (fwd only ~ 1/3)
    return dt, float(total.detach()), dict(rois=rois, windows=nwin)


_CPU_WARM = [False]


def cpu_reference_throughput(threads, budget_s=25.0):
    """images/sec of the CPU restatement: one discarded warm-up pass (8 ROIs per head), then ONE complete
    step of the real workload (256 ROIs per head, 64 windows, 1280 refine windows) timed end to end."""
    if not _CPU_WARM[0]:
        oracle_sample(threads, 8)
        _CPU_WARM[0] = True
    t_full, loss, _ = oracle_sample(threads, 256)
    return 1.0 / t_full, dict(full_step_s=t_full, loss=loss)


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    detail = None
    for i in range(args.warmup + args.steps):
        v, detail = cpu_reference_throughput(threads)
        if i >= args.warmup:
            vals.append(v)
    value = sum(vals) / len(vals)
    sample = ("oracle/model.py fp32 torch-CPU restatement of the reference step (TF1/py2 reference cannot run "
              "here); sample = one complete 600x1000 training step (forward, 8 losses + L2, autograd backward) after one "
              "discarded warm-up pass: %s" % json.dumps(detail))
    line = {"impl": "reference", "metric": "images/sec", "value": value, "unit": "images/sec", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 / value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": 1, "per_gpu_batch": 1, "parallelism": "cpu",
                       "note": "fp32 CPU restatement of the same step (oracle/model.py), one image per step"},
            "cpu_baseline": {"value": value, "unit": "images/sec", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from helpers import load_config
    from mtl_ssl_b200 import ops, ops_conv
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    from mtl_ssl_b200.trainer import Trainer
    from mtl_ssl_b200.utils import synthetic_init

    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    cfg = load_config("model12.config")
    B = args.batch_per_gpu or cfg.train_config.batch_size
    model = model_builder.build(cfg.model, True, device=dev, seed=0)      # identical weights on every rank
    synthetic_init.apply(model.param_store)
    K = cfg.model.faster_rcnn.num_classes
    nk = model.num_kept_anchors((B, H, W, 3))
    tr = Trainer(model, cfg.train_config, H, W, B, gmax=16, use_cuda_graph=not args.no_graph, world_size=world)
    pool = []
    for i in range(4):
        ex = synthetic.make_batch(1234 + rank * 100 + i, B, H, W, K, max_boxes=8, num_windows=64)
        keys = synthetic.make_sampler_keys(99 + rank * 100 + i, B, nk, cfg.model.faster_rcnn.first_stage_max_proposals)
        pool.append(tr.host_arrays(ex, keys))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also captures the CUDA graphs)
    l0 = ops.launch_count()
    for i in range(max(args.warmup, 3)):
        losses = tr.step(pool[i % len(pool)])
    sync()
    # kernels per step, counted by the library itself during one eager replay of the step body
    c0 = ops.launch_count()
    tr._prefix(tr.inputs.dev["image"])                      # frozen conv1 + block1 (pipelined one step ahead)
    tr._forward_backward(tr.inputs.dev["image"])
    if world > 1:
        tr._backward_trunk()
    model.param_store.g.zero_()
    c1 = ops.launch_count()
    # + the optimizer launches not in the count above: stats, fixed-order norm reduction, L2 reduction and apply for
    # the trunk range, and (several replicas: its own graph) stats + reduction + apply of the head bucket
    launches_per_step = (c1 - c0) + (4 if world == 1 else 7)
    sync()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    # ---- timed region 1: device-resident (inputs already in HBM), K steps
    st = model.param_store
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for i in range(args.steps):
        tr.run_resident_step()          # = the step body with the frozen prefix of the next step computed underneath
    e1.record()
    sync()
    ms_dev = e0.elapsed_time(e1)
    # ---- timed region 2: end to end through the public API (H2D of every batch + D2H of every loss vector).
    # step_pipelined() stages batch i+1 (host copy into pinned memory + H2D on a copy stream) while step i
    # computes and hands back the losses of the previous call; flush() inside the region collects the last one.
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    t_wall0 = time.perf_counter()
    e2.record()
    got = 0
    for i in range(args.steps):
        r = tr.step_pipelined(pool[i % len(pool)])
        if r is not None:
            losses, got = r, got + 1
    r = tr.flush()
    if r is not None:
        losses, got = r, got + 1
    e3.record()
    sync()
    assert got == args.steps, (got, args.steps)
    wall_e2e = time.perf_counter() - t_wall0
    ms_e2e = max(e2.elapsed_time(e3), wall_e2e * 1000.0)     # host-side work counts end to end
    # the fully synchronous variant (Trainer.step: copy, compute, read back, one after the other), for reference
    sync()
    t_wall0 = time.perf_counter()
    for i in range(args.steps):
        losses = tr.step(pool[i % len(pool)], read_losses=True)
    sync()
    ms_e2e_sync = (time.perf_counter() - t_wall0) * 1000.0
    clk = clocks.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_dev, ms_e2e, ms_e2e_sync], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e, ms_e2e_sync = t.tolist()
    if rank != 0:
        return
    # ---- roofline of the dominant kernel family, measured live: one eager step with per-launch events.
    # Two precautions make the event pairs bracket ONLY kernel time: (a) the multi-stream overlap is switched
    # off for this pass (overlapping kernels share SMs and would inflate each other's duration), (b) a spin
    # kernel keeps the GPU busy while Python enqueues the step, so no event pair contains host launch latency.
    from mtl_ssl_b200.nets.layers import Concurrency
    Concurrency.enabled = False
    overlap, tr.overlap_optimizer = tr.overlap_optimizer, False      # keep the optimizer out of the conv timings
    tr._prefix(tr.inputs.dev["image"])
    tr._forward_backward(tr.inputs.dev["image"])            # re-warm the single-stream path
    if world > 1:
        tr._backward_trunk()
    model.param_store.g.zero_()
    torch.cuda.synchronize()
    ops_conv.PROFILE = []
    torch.cuda._sleep(int(60e-3 * 1.9e9))                   # ~60 ms head start for the host
    tr._prefix(tr.inputs.dev["image"])
    tr._forward_backward(tr.inputs.dev["image"])
    if world > 1:
        tr._backward_trunk()
    torch.cuda.synchronize()
    prof, ops_conv.PROFILE = ops_conv.PROFILE, None
    Concurrency.enabled = True
    tr.overlap_optimizer = overlap
    model.param_store.g.zero_()
    conv_ms = sum(p_[2].elapsed_time(p_[3]) for p_ in prof)
    conv_flops = sum(p_[1] for p_ in prof)
    by_mode = {}
    for m, f, a, b, _g in prof:
        d = by_mode.setdefault(("fprop", "dgrad", "wgrad")[m], [0.0, 0.0, 0])
        d[0] += f; d[1] += a.elapsed_time(b); d[2] += 1
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained")
    peak_src = "measured sustained (MEASURED_PEAKS.json)"
    if not peak:
        peak, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained)"
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12
    try:
        dominant = dominant_conv_group([(m, f, a.elapsed_time(b), g_) for m, f, a, b, g_ in prof], peak)
    except Exception as e:                                   # never lose the bench line over the breakdown
        dominant = {"error": repr(e)}
    images = B * world * args.steps
    value = images / (ms_dev * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)
    h2d = tr.inputs.nbytes
    step_ms = ms_dev / args.steps
    line = {
        "metric": "images/sec", "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "global_batch": B * world, "per_gpu_batch": B, "parallelism": "dp%d" % world,
                   "l2_flush": "none needed: one step touches %.1f GB of activations+weights (>> 126 MB L2)"
                               % ((model.workspace.nbytes() + st.total * 14) / 1e9),
                   "cuda_graph": bool(tr.use_graph)},
        "e2e": {"value": e2e_value, "unit": "images/sec", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 36,
                "ms_per_step": ms_e2e / args.steps, "api": "Trainer.step_pipelined (batch i+1 staged while step i runs)",
                "synchronous_step_value": images / (ms_e2e_sync * 1e-3)},
        "gpu_launches": launches_per_step * args.steps * 2,
        "gpu_launches_per_step": launches_per_step,
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "tc_gemm_kernel (all %d conv/FC launches of one step)" % len(prof),
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": None, "dominant": dominant,
                     "conv_ms_per_step": conv_ms, "conv_share_of_step": conv_ms / step_ms,
                     "by_mode": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12, "ms": v[1], "launches": v[2]}
                                 for k, v in by_mode.items()},
                     "step_tflops": algorithmic_flops_per_image() * B / (step_ms * 1e-3) / 1e12},
        "losses": losses,
    }
    if world == 1 and not args.skip_cpu:
        threads = os.cpu_count() or 1
        v, detail = cpu_reference_throughput(threads)
        line["cpu_baseline"] = {"value": v, "unit": "images/sec", "cores": threads, "kind": "port",
                                "sample": "oracle/model.py (fp32 torch-CPU restatement; the TF1/py2 reference cannot "
                                          "run here): one complete 600x1000 training step (forward, losses, "
                                          "autograd backward) after a discarded warm-up pass: %s" % json.dumps(detail)}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--batch-per-gpu", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
