#!/usr/bin/env python
"""Benchmark of the hot path: images/sec of one full Faster R-CNN / R-FCN multi-task training step (forward, 8 losses,
explicit backward, gradient all-reduce, clip + momentum update).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c1..c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workloads = BASELINE.json `configs` (the shipped config files unchanged, synthetic inputs, random-init weights):
    c1  model51.config  Faster R-CNN MobileNet-v1 baseline (no aux heads), 2 x 300x300 per replica (-> 600x600)
    c2  model12.config  Faster R-CNN ResNet-101 + 3 aux heads + refiner, 600x1000, 1 image per replica   [DEFAULT]
    c3  model22.config  the same on COCO shapes: K=90, crop 14 + 2x2 max pool, 800x1333 (-> 600x1000), 2 per replica
    c4  model42.config  R-FCN ResNet-101 + aux heads (position-sensitive ROI pooling), 800x1333, 1 per replica
    c5  model62.config  Faster R-CNN Inception-ResNet-v2 + 3 aux heads, 800x1333, 4 per replica
bf16 tensor-core convolutions with fp32 accumulation and fp32 master weights.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA-graph replays, inputs in HBM; the frozen
conv1 + block1 prefix of the next step is computed underneath the current one -- on the same resident image, nothing is
skipped); `e2e` = the same step through the public API (Trainer.step_pipelined) with pinned-host inputs copied H2D and
the loss vector read back D2H every step.  `roofline` is measured live: every tcgen05 conv launch of one eager step is
bracketed by CUDA events; achieved = algorithmic FLOPs / kernel time.  `loss_parity` (1 GPU): the device's losses on the
first batch, from the initial weights, against the CPU oracle on the same weights and inputs.
`--impl reference` times the CPU restatement of the reference path (oracle/, torch fp32 on all host threads) on the
same weights and inputs: the reference itself is Python-2 / TF-1.7 graph code and cannot run in this image.
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

STEP = "train step (fwd+bwd+allreduce+clip+momentum)"
CONFIGS = {
    "c1": dict(file="model51.config", H=300, W=300, batch=2, bn="randomize",
               workload="Faster R-CNN MobileNet-v1 baseline (no aux heads), model51.config, 2 x 300x300 -> 600x600, " + STEP),
    "c2": dict(file="model12.config", H=600, W=1000, batch=1, bn="synthetic",
               workload="Faster R-CNN ResNet-101 + 3 aux heads + refiner, model12.config, 600x1000, " + STEP),
    "c3": dict(file="model22.config", H=800, W=1333, batch=2, bn="synthetic",
               workload="Faster R-CNN ResNet-101 + 3 aux heads + refiner, model22.config (K=90, crop 14 + maxpool), "
                        "COCO-shape 800x1333 -> 600x1000, 2 images per GPU, " + STEP),
    "c4": dict(file="model42.config", H=800, W=1333, batch=1, bn="synthetic",
               workload="R-FCN ResNet-101 + aux heads (PS-ROI), model42.config, COCO-shape 800x1333 -> 600x1000, " + STEP),
    "c5": dict(file="model62.config", H=800, W=1333, batch=4, bn="randomize",
               workload="Faster R-CNN Inception-ResNet-v2 + 3 aux heads, model62.config, COCO-shape 800x1333 -> 600x1000, "
                        "4 images per GPU, " + STEP),
}
# 2*MACs of every conv / FC of one train step per image (fwd + bwd), where the survey states it
SURVEY_FLOPS_PER_IMAGE = {"c2": 4.92e12}


def algorithmic_flops_per_image():
    """2*MACs of every conv/FC of one train step of the default workload (c2) at 600x1000
    (fwd 2.90 T + bwd 2.02 T)."""
    return SURVEY_FLOPS_PER_IMAGE["c2"]


def workload_config(key, B, world):
    """The `config` object of the JSON line: identical in both arms (ours / reference) for one workload."""
    c = CONFIGS[key]
    return {"workload": c["workload"], "bench_config": key, "config_file": c["file"], "input_hw": [c["H"], c["W"]],
            "per_gpu_batch": B, "global_batch": B * world, "parallelism": "dp%d" % world,
            "l2_flush": "none needed: one step touches several GB of activations + weights (>> 126 MB L2)"}


def initial_state(key, seed=0):
    """(parsed config, host state dict {TF variable name: fp32 tensor}): the weights BOTH arms start from.  Built on the
    host exactly as ParamStore.finalize draws them; frozen batch-norm statistics chosen so that random-init activations
    stay O(1) (utils/synthetic_init.py for the ResNets, tests/helpers.randomize_bn for MobileNet / Inception-ResNet)."""
    from helpers import load_config, randomize_bn
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.utils import synthetic_init
    c = CONFIGS[key]
    cfg = load_config(c["file"])
    table = model_builder.build(cfg.model, True, device=None, seed=seed)
    if c["bn"] == "synthetic":
        synthetic_init.apply_host(table.param_store)
    sd = table.param_store.host_state_dict(seed)
    if c["bn"] == "randomize":
        sd = randomize_bn(sd, seed)
    return cfg, sd, table


def first_batch(key, cfg, nk, B, rank=0, i=0):
    """Seeded synthetic batch + sampler keys number `i` of `rank` (the same function feeds both arms).
    nk: anchors kept by the window pruning = length of the first-stage sampler keys."""
    from mtl_ssl_b200.data import synthetic
    c = CONFIGS[key]
    fr = cfg.model.faster_rcnn
    ex = synthetic.make_batch(1234 + rank * 100 + i, B, c["H"], c["W"], fr.num_classes, max_boxes=8, num_windows=64)
    keys = synthetic.make_sampler_keys(99 + rank * 100 + i, B, nk, fr.first_stage_max_proposals)
    return ex, keys


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
    if R == 0:      # a grouped launch (ops_conv.ConvGroup): geometry slot 0 holds the number of problems
        name = "tc_gemm_kernel %s, grouped launch of %d problems (persistent grid capped for overlapped side work)" % (
            ("fprop", "dgrad", "wgrad")[mode], N)
    else:
        name = "tc_gemm_kernel %s %dx%d/%d C%d->K%d on %dx%dx%d" % (("fprop", "dgrad", "wgrad")[mode], R, R, stride, C, K,
                                                                    N, H_, W_)
    out = {"kernel": name, "launches": n, "avg_us": ms * 1e3 / n, "share_of_conv_time": ms / total_ms,
           "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
           "traffic": None}
    for fn in ("r2_kernel_traffic.json", "r1_kernel_traffic.json"):
        try:        # DRAM bytes per launch of this kernel from the committed ncu --set full capture, when there is one
            t = json.load(open(os.path.join(ROOT, "profiles", fn))).get(name)
            if t:
                out["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"]
                out["traffic_source"] = "profiles/%s (%s)" % (fn, t.get("capture", "ncu --set full"))
                break
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
class CpuReference(object):
    """The CPU restatement of the reference step (oracle/model.py, fp32, torch-CPU on all host threads) on the
    workload's own weights and first batch."""

    def __init__(self, key, threads):
        import torch
        from helpers import oracle_config
        torch.set_num_threads(threads)
        self.key, self.threads = key, threads
        self.cfg, self.sd, self.table = initial_state(key)
        self.B = CONFIGS[key]["batch"]
        self.ocfg = oracle_config(self.cfg)
        st = self.table.param_store
        # (the dead stage-1 block4 copy takes no part in the forward pass but its L2 terms are in the total loss, trap T4)
        self.params = {k: v for k, v in self.sd.items() if "/_pad/" not in k}
        self.l2 = {p.name: p.l2 for p in st.params if p.name in self.params}
        self.trainable = [p.name for p in st.params if p.trainable and p.name in self.params]
        resizer = getattr(self.table, "_image_resizer_fn", None)
        c = CONFIGS[key]
        self.Hr, self.Wr = (c["H"], c["W"]) if resizer is None else tuple(int(v) for v in resizer.static_size(c["H"], c["W"]))
        # anchors inside the resized image (fmA:930-976), counted with the oracle's own anchor code (no device here)
        from oracle import boxes as OB
        hf, wf = self.table._feature_extractor.feature_map_shape(self.Hr, self.Wr)
        anchors = OB.grid_anchors(hf, wf, self.ocfg["scales"], self.ocfg["aspect_ratios"], (256, 256), (16, 16), (0, 0))
        self.nk = len(OB.prune_outside_window(anchors, (0, 0, self.Hr, self.Wr))[1])
        self.examples, self.keys = first_batch(key, self.cfg, self.nk, self.B)

    def images(self):
        import numpy as np
        import torch
        from oracle import nn as ON
        c = CONFIGS[self.key]
        img = torch.from_numpy(np.stack([e["image"] for e in self.examples]).astype(np.float32))
        if (self.Hr, self.Wr) != (c["H"], c["W"]):
            img = ON.resize_bilinear(img, (self.Hr, self.Wr))          # model.preprocess (fmA:479-505)
        return img

    def step(self):
        """One complete training step (forward, 8 losses + L2, autograd backward).  -> (seconds, losses dict)."""
        from oracle.model import Oracle
        orc = Oracle(self.params, self.ocfg, bf16=False)
        orc.require_grad(self.trainable)
        t0 = time.perf_counter()
        img = self.images()
        out = orc.forward(img, self.examples, self.keys, self.Hr, self.Wr)
        losses = orc.loss(out, self.examples, self.keys, self.Hr, self.Wr)
        reg = orc.regularization_loss(self.l2)
        total = sum(losses.values()) + reg
        total.backward()
        dt = time.perf_counter() - t0
        ld = {k: float(v.detach()) for k, v in losses.items()}
        ld["regularization_loss"] = float(reg.detach()) if hasattr(reg, "detach") else float(reg)   # (0.0: no L2 term)
        ld["total_loss"] = float(total.detach())
        return dt, ld

    def mirror_losses(self, proposal_inputs):
        """Forward + losses with the device's bf16 rounding points mirrored, on the device's RPN outputs."""
        import torch
        from oracle.model import Oracle
        orc = Oracle(self.params, self.ocfg, bf16=True)
        with torch.no_grad():
            out = orc.forward(self.images(), self.examples, self.keys, self.Hr, self.Wr, proposal_inputs=proposal_inputs)
            losses = orc.loss(out, self.examples, self.keys, self.Hr, self.Wr)
        return {k: float(v) for k, v in losses.items()}, out

    def sample_text(self, detail):
        c = CONFIGS[self.key]
        return ("oracle/model.py fp32 torch-CPU restatement of the reference step (the TF1/py2 reference cannot run here) "
                "on the workload's own initial weights and first batch; sample = one complete %dx%d training step of %d "
                "image(s) (forward, 8 losses + L2, autograd backward): %s" % (c["H"], c["W"], self.B, json.dumps(detail)))


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    ref = CpuReference(args.config, threads)
    ref.step()                                       # discarded warm-up pass (thread pools, allocator)
    times, losses = [], None
    for i in range(args.warmup + args.steps):
        dt, losses = ref.step()
        if i >= args.warmup:
            times.append(dt)
    B = ref.B
    value = B * len(times) / sum(times)
    line = {"impl": "reference", "metric": "images/sec", "value": value, "unit": "images/sec", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * sum(times) / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args.config, B, world),
            "cpu_baseline": {"value": value, "unit": "images/sec", "cores": threads, "kind": "port",
                             "sample": ref.sample_text({"step_s": sum(times) / len(times)})},
            "e2e": {"value": value, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "losses": losses}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from mtl_ssl_b200 import ops, ops_conv
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.meta_architectures.faster_rcnn_meta_arch import LOSS_KEYS
    from mtl_ssl_b200.trainer import Trainer

    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    args.config = getattr(args, "config", "c2")
    c = CONFIGS[args.config]
    H, W = c["H"], c["W"]
    cfg, sd0, table = initial_state(args.config)
    B = args.batch_per_gpu or c["batch"]
    model = model_builder.build(cfg.model, True, device=dev, seed=0)      # identical weights on every rank
    model.param_store.load_state_dict(sd0)
    tr = Trainer(model, cfg.train_config, H, W, B, gmax=16, use_cuda_graph=not args.no_graph, world_size=world)
    pool = []
    nk = model.num_kept_anchors((B, H, W, 3))
    for i in range(4):
        ex, keys = first_batch(args.config, cfg, nk, B, rank, i)
        pool.append(tr.host_arrays(ex, keys))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (also captures the CUDA graphs)
    for i in range(max(args.warmup, 3)):
        losses = tr.step(pool[i % len(pool)])
    sync()
    # kernels per step, counted by the library itself during one eager replay of the step body
    st = model.param_store
    snap = (st.w.clone(), st.m.clone(), st.wb.clone())
    c0 = ops.launch_count()
    tr.finish()
    tr.eager_pass()                                         # prefix + forward + backward + both optimizer buckets
    c1 = ops.launch_count()
    launches_per_step = c1 - c0
    torch.cuda.synchronize()
    st.w.copy_(snap[0]); st.m.copy_(snap[1]); st.wb.copy_(snap[2])      # this pass skipped the gradient exchange: undo it
    st.g.zero_()
    del snap
    sync()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    # ---- timed region 1: device-resident (inputs already in HBM), K steps
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
    tr.finish()
    Concurrency.enabled = False                             # every launch of the pass on one stream
    snap = (st.w.clone(), st.m.clone(), st.wb.clone())
    tr.eager_pass()                                         # re-warm the single-stream path
    torch.cuda.synchronize()
    ops_conv.PROFILE = []
    torch.cuda._sleep(int(60e-3 * 1.9e9))                   # ~60 ms head start for the host
    tr.eager_pass()
    torch.cuda.synchronize()
    prof, ops_conv.PROFILE = ops_conv.PROFILE, None
    Concurrency.enabled = True
    st.w.copy_(snap[0]); st.m.copy_(snap[1]); st.wb.copy_(snap[2])
    model.param_store.g.zero_()
    del snap
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
    try:        # per-geometry table of the serialised pass (developer aid; summarised under profiles/)
        groups = {}
        for m, f, a, b, g_ in prof:
            k = "%s N%d %dx%d C%d->K%d %dx%d/%d" % (("fprop", "dgrad", "wgrad")[m], g_[0], g_[1], g_[2], g_[3], g_[4], g_[5],
                                                   g_[5], g_[6])
            d = groups.setdefault(k, [0, 0.0, 0.0])
            d[0] += 1; d[1] += a.elapsed_time(b) * 1e3; d[2] += f
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "layers_%s.json" % args.config), "w") as fh:
            json.dump({k: {"launches": v[0], "us": v[1], "tflops": v[2] / v[1] / 1e6} for k, v in groups.items()}, fh,
                      indent=1)
    except Exception:
        pass
    images = B * world * args.steps
    value = images / (ms_dev * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)
    h2d = tr.inputs.nbytes + tr._hyper_host.numel() * tr._hyper_host.element_size()
    d2h = tr._loss_host.numel() * tr._loss_host.element_size()
    step_ms = ms_dev / args.steps
    # algorithmic FLOPs of one step: SURVEY 8(d)'s figure where it states one, else the sum over the step's conv / FC
    # launches (2*N*P*Q*K*R*S*C each -- the same count the survey's figure is made of; tests/test_zz_dryrun_host_paths)
    survey = SURVEY_FLOPS_PER_IMAGE.get(args.config)
    step_flops = survey * B if survey else conv_flops
    cfg_line = workload_config(args.config, B, world)
    line = {
        "metric": "images/sec", "value": value, "unit": "images/sec", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": cfg_line,
        "schedule": {"cuda_graph": bool(tr.use_graph), "timed_region_ms": ms_dev,
                     "resident_step": "CUDA-graph replays on the batch resident in HBM; the frozen conv1 + block1 prefix "
                                      "of the NEXT step runs on a side stream under the current one (same resident image; "
                                      "every step still executes its own prefix, nothing is cached or skipped)",
                     "touched_gb_per_step": (model.workspace.nbytes() + st.total * 14) / 1e9},
        "e2e": {"value": e2e_value, "unit": "images/sec", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps, "api": "Trainer.step_pipelined (batch i+1 staged while step i runs)",
                "synchronous_step_value": images / (ms_e2e_sync * 1e-3)},
        "gpu_launches": launches_per_step * args.steps * 2,
        "gpu_launches_per_step": launches_per_step,
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "tc_gemm_kernel (all %d conv/FC launches of one step)" % len(prof),
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                     "peak_source": peak_src, "traffic": None, "dominant": dominant,
                     "conv_ms_per_step": conv_ms, "conv_share_of_step": conv_ms / step_ms,
                     "note": "achieved = per-launch rate of a serialised single-stream pass; step_tflops = the timed "
                             "multi-stream step's algorithmic rate (step_frac = step_tflops / peak)",
                     "by_mode": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12, "ms": v[1], "launches": v[2]}
                                 for k, v in by_mode.items()},
                     "step_flops": step_flops, "step_flops_source": "SURVEY 8(d)" if survey else "sum over launches",
                     "launch_flops": conv_flops,
                     "step_tflops": step_flops / (step_ms * 1e-3) / 1e12,
                     "step_frac": step_flops / (step_ms * 1e-3) / 1e12 / peak},
        "losses": losses,
    }
    if world == 1 and not args.skip_cpu:
        # ---- loss parity + CPU baseline: both arms start from the same weights (sd0) and use the same first batch
        threads = os.cpu_count() or 1
        ref = CpuReference(args.config, threads)
        tr.finish()
        st.load_state_dict(sd0)
        st.m.zero_(); st.g.zero_()
        tr.overlap_optimizer = False
        image = tr._bind(pool[0])
        tr._prefix(image)
        pd = tr._forward_backward(image)
        torch.cuda.synchronize()
        dev_l = dict(zip(LOSS_KEYS, model.workspace.bufs["loss/values"].cpu().tolist()))
        prop_in = (pd["rpn_box_encodings"].cpu().numpy(),
                   pd["rpn_objectness_predictions_with_background"].cpu().numpy())
        want, out = ref.mirror_losses(prop_in)
        dev_total = sum(dev_l[k] for k in want)
        orc_total = sum(want.values())
        ref.step()                                   # discarded warm-up pass
        dt, fp32 = ref.step()
        fp32_task = fp32["total_loss"] - fp32["regularization_loss"]
        line["loss_parity"] = {
            "device_total": dev_total, "oracle_total": orc_total, "abs_diff": abs(dev_total - orc_total), "tol": 1e-3,
            "tol_kind": "relative to max(1, oracle_total)", "rel_diff": abs(dev_total - orc_total) / max(1.0, abs(orc_total)),
            "ok": bool(abs(dev_total - orc_total) <= 1e-3 * max(1.0, abs(orc_total))
                       and np.array_equal(pd["num_proposals"].cpu().numpy(), out["nprop"])),
            "oracle": "oracle/model.py with the device's bf16 rounding points mirrored, on the device's RPN outputs "
                      "(proposal counts identical: %s); sum of the task losses of fmA:1514-1589, same initial weights and "
                      "batch on both sides" % bool(np.array_equal(pd["num_proposals"].cpu().numpy(), out["nprop"])),
            "max_loss_abs_diff": max(abs(dev_l[k] - v) for k, v in want.items()),
            "oracle_fp32_total": fp32_task, "fp32_abs_diff": abs(dev_total - fp32_task),
            "fp32_note": "plain fp32 oracle with its OWN proposals (the reference's arithmetic; the cpu_baseline step)",
            "device_losses": {k: dev_l[k] for k in want}, "oracle_losses": want}
        v = ref.B / dt
        line["cpu_baseline"] = {"value": v, "unit": "images/sec", "cores": threads, "kind": "port",
                                "sample": ref.sample_text({"step_s": dt, "total_loss": fp32["total_loss"]})}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)     # default: 100 (0.7 s timed at c2); CPU reference arm: 20
    ap.add_argument("--warmup", type=int, default=None)    # default: 5; CPU reference arm: 3
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch-per-gpu", type=int, default=0)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 20 if args.impl == "reference" else 100
    if args.warmup is None:
        args.warmup = 3 if args.impl == "reference" else 5
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
