#!/bin/bash
# One gpurun call at the end of a round: the GPU suite, smoke(), the default bench line and the reference arm.
# `final_check.sh all` adds the bench lines of the other BASELINE configs (1 GPU).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench_c2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/final_bench_c2.log > gpurun_out/r2_bench_c2_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/final_bench_ref.log | cut -c1-300
CFGS="c2"
if [ "$1" = "all" ]; then
  CFGS="c1 c2 c3 c4 c5"
  for c in c1 c3 c4 c5; do
    python bench.py --config $c > gpurun_out/final_bench_$c.log 2>&1; echo "$c rc=$?"; tail -1 gpurun_out/final_bench_$c.log > gpurun_out/r2_bench_${c}_1gpu.json
  done
fi
CFGS="$CFGS" python - <<'PY'
import json, os
for c in os.environ["CFGS"].split():
    try:
        l = json.load(open("gpurun_out/r2_bench_%s_1gpu.json" % c))
        lp = l.get("loss_parity") or {}
        print(c, round(l["value"], 1), round(l["ms_per_step"], 3), "e2e", round(l["e2e"]["value"], 1), "parity", lp.get("ok"), lp.get("rel_diff"),
              "tflops", round(l["roofline"]["step_tflops"], 1), "launches", l["gpu_launches"], l["clocks"])
    except Exception as e:
        print(c, "failed", e)
PY
