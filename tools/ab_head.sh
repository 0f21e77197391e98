#!/bin/bash
# Developer tool (GPU): A/B of the fused head kernels (csrc/head.cu) against the separate pool + GEMM launches, same box,
# alternating order; then the ncu launch list of one eager step in the default configuration.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_roi_loss_kernels.py -q -x -k "fused_head" 2>&1 | tail -2
for i in 1 2; do
  MTL_NO_FUSED_HEAD=1 python bench.py --steps 40 --warmup 5 --skip-cpu 2>/dev/null | tail -1 > gpurun_out/ab_head_sep_$i.json
  python bench.py --steps 40 --warmup 5 --skip-cpu 2>/dev/null | tail -1 > gpurun_out/ab_head_fused_$i.json
done
python - <<'PY'
import json
for k in ("sep_1", "fused_1", "sep_2", "fused_2"):
    l = json.load(open("gpurun_out/ab_head_%s.json" % k))
    print(k, l["ms_per_step"], l["value"], l["e2e"]["value"], l["gpu_launches"])
PY
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/r2_launches_step_fused.csv python tools/eager_steps.py --config c2 --steps 1 > gpurun_out/ncu_fused.log 2>&1
echo "ncu rc=$?"
