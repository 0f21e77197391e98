"""Developer tool (GPU): a per-launch timeline of one training step across all streams (nsys is not in the image).

Every kernel launch of the step (tcgen05 convs through ops_conv, everything else through ops.call) is bracketed by CUDA
events on its own stream; the whole step is enqueued behind a spin kernel so the host never starves the GPU, exactly as a
CUDA-graph replay would issue it.  Output: gpurun_out/<tag>_timeline.csv (label, stream, start_us, end_us, flops) and a
coarse occupancy summary (how much of the step has 1, 2, 3+ kernels in flight; idle gaps).  The events sit between
consecutive launches of a stream, which disables programmatic dependent launch overlap there: chains of short kernels read
~1 us per launch longer than in the captured step.

    python tools/timeline.py [--config c2] [--tag r2] [--batch N]
    python tools/timeline.py --analyze gpurun_out/r2_timeline.csv        (no GPU needed)
"""
import argparse
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def analyze(path, top=25):
    rows = [(r["label"], r["stream"], float(r["start_us"]), float(r["end_us"]), float(r["flops"]))
            for r in csv.DictReader(open(path))]
    rows.sort(key=lambda r: r[2])
    t0 = min(r[2] for r in rows)
    t1 = max(r[3] for r in rows)
    print("%d launches over %d streams, span %.1f us" % (len(rows), len({r[1] for r in rows}), t1 - t0))
    # kernels in flight over time
    ev = []
    for r in rows:
        ev.append((r[2], 1)); ev.append((r[3], -1))
    ev.sort()
    depth, last, hist = 0, t0, collections.Counter()
    for t, d in ev:
        hist[min(depth, 4)] += t - last
        depth += d
        last = t
    for k in sorted(hist):
        print("  %s kernels in flight: %8.1f us (%4.1f %%)" % ("4+" if k == 4 else k, hist[k], 100 * hist[k] / (t1 - t0)))
    # per stream busy time
    by_stream = collections.defaultdict(float)
    for r in rows:
        by_stream[r[1]] += r[3] - r[2]
    for s, v in sorted(by_stream.items(), key=lambda kv: -kv[1]):
        print("  stream %-16s busy %8.1f us" % (s, v))
    # phases: windows of 250 us with the dominant labels
    print("timeline (250 us windows: busy stream-time, TFLOP/s, top labels)")
    w = 250.0
    n = int((t1 - t0) / w) + 1
    for i in range(n):
        a, b = t0 + i * w, t0 + (i + 1) * w
        busy, fl, lab = 0.0, 0.0, collections.Counter()
        for r in rows:
            o = min(r[3], b) - max(r[2], a)
            if o > 0:
                busy += o
                fl += r[4] * o / max(r[3] - r[2], 1e-9)
                lab[r[0].split(" ")[0] + ("" if not r[0].startswith("conv") else " " + " ".join(r[0].split(" ")[1:3]))] += o
        print("  %6.0f us  busy %6.0f  %6.0f TF/s  %s" % (a - t0, busy, fl / (w * 1e-6) / 1e12,
                                                            ", ".join("%s %.0f" % kv for kv in lab.most_common(3))))
    print("longest launches")
    for r in sorted(rows, key=lambda r: -(r[3] - r[2]))[:top]:
        print("  %8.1f us @%8.1f  %-14s %s" % (r[3] - r[2], r[2] - t0, r[1], r[0]))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    ap.add_argument("--tag", default="r2")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--analyze", default=None)
    args = ap.parse_args()
    if args.analyze:
        return analyze(args.analyze)
    import torch
    import bench
    from mtl_ssl_b200 import ops, ops_conv
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.trainer import Trainer
    c = bench.CONFIGS[args.config]
    cfg, sd0, _ = bench.initial_state(args.config)
    B = args.batch or c["batch"]
    model = model_builder.build(cfg.model, True, device="cuda", seed=0)
    model.param_store.load_state_dict(sd0)
    tr = Trainer(model, cfg.train_config, c["H"], c["W"], B, gmax=16, use_cuda_graph=False)
    nk = model.num_kept_anchors((B, c["H"], c["W"], 3))
    ex, keys = bench.first_batch(args.config, cfg, nk, B)
    arrays = tr.host_arrays(ex, keys)
    for _ in range(3):
        tr.step(arrays)
    torch.cuda.synchronize()
    base = torch.cuda.Event(enable_timing=True)
    tl = []
    ops.TIMELINE = tl
    ops_conv.TIMELINE = tl
    torch.cuda._sleep(int(150e-3 * 1.9e9))            # head start for the host: the step is queued before it runs
    base.record()
    tr._run_step_body(defer=True)
    tr._run_step_body(defer=True)           # (the second one carries the first one's deferred head update)
    torch.cuda.synchronize()
    ops.TIMELINE = None
    ops_conv.TIMELINE = None
    out = os.path.join(ROOT, "gpurun_out", "%s_timeline.csv" % args.tag)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    names = {}
    with open(out, "w", newline="") as f:
        wr = csv.writer(f)
        wr.writerow(["label", "stream", "start_us", "end_us", "flops"])
        for label, stream, e0, e1, fl in tl:
            sid = names.setdefault(stream, "s%d" % len(names))
            wr.writerow([label, sid, "%.2f" % (base.elapsed_time(e0) * 1e3), "%.2f" % (base.elapsed_time(e1) * 1e3),
                         "%.0f" % fl])
    analyze(out)


if __name__ == "__main__":
    main()
