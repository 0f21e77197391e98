"""Developer tool (GPU): device time of every stage graph of the captured training step replayed ALONE (nothing beside
it), against the whole step -- how much of the stages' summed time the schedule hides.

    python tools/stage_times.py [--config c2]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2")
    args = ap.parse_args()
    import torch
    import bench
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.trainer import Trainer
    c = bench.CONFIGS[args.config]
    cfg, sd0, _ = bench.initial_state(args.config)
    B = c["batch"]
    model = model_builder.build(cfg.model, True, device="cuda", seed=0)
    model.param_store.load_state_dict(sd0)
    tr = Trainer(model, cfg.train_config, c["H"], c["W"], B, gmax=16, use_cuda_graph=True)
    nk = model.num_kept_anchors((B, c["H"], c["W"], 3))
    ex, keys = bench.first_batch(args.config, cfg, nk, B)
    arrays = tr.host_arrays(ex, keys)
    for _ in range(4):
        tr.step(arrays)
    tr.finish()
    torch.cuda.synchronize()

    def timed(fn, n=10):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3

    stages = [("prefix (frozen conv1 + block1)", tr.graph_prefix), ("A  trunk forward + RPN + proposals", tr.graph_fa),
              ("B  second stage forward, losses, second-stage dgrad", tr.graph_fb),
              ("T  RPN + trunk backward (with its chunked weight gradients)", tr.graph_ft),
              ("D  trunk clip + momentum", tr.graph_opt),
              ("C  deferred second-stage weight gradients (%d SMs)" % tr._deferred_ctas, tr.graph_hw),
              ("   second-stage clip + momentum", tr.graph_opt_heads)]
    total = 0.0
    for name, g in stages:
        if g is None:
            continue
        us = timed(g.replay)
        total += us
        print("%-62s %8.1f us" % (name, us))
    print("%-62s %8.1f us" % ("sum of the stages replayed alone", total))
    step = timed(lambda: tr.run_resident_step(), 20)
    print("%-62s %8.1f us" % ("the step (run_resident_step)", step))


if __name__ == "__main__":
    main()
