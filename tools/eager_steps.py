"""Developer tool (GPU): a few eager (no CUDA graph) training steps of one bench workload, with the CUDA profiler range
open only around the last ones -- the command ncu wraps for the launch list and for the full captures of the
bandwidth-bound kernels (`--profile-from-start off`).

    ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file L.csv \
        python tools/eager_steps.py --config c2 --steps 1
    ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'crop_resize|nms_kernel|...' \
        -o gpurun_out/r2_bw_c2 python tools/eager_steps.py --config c2 --steps 1
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
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()
    import torch
    import bench
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
    for _ in range(args.warmup):
        tr.step(arrays)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for _ in range(args.steps):
        losses = tr.step(arrays)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("total_loss %.5f" % losses["total_loss"])


if __name__ == "__main__":
    main()
