#!/bin/bash
# One gpurun call at the end of a round: the GPU suite, smoke(), and the bench lines of every BASELINE config (1 GPU).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench_c2.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/final_bench_c2.log > gpurun_out/r2_bench_c2_1gpu.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/final_bench_ref.log | cut -c1-400
for c in c1 c3 c4 c5; do
  python bench.py --config $c > gpurun_out/final_bench_$c.log 2>&1; echo "$c rc=$?"; tail -1 gpurun_out/final_bench_$c.log > gpurun_out/r2_bench_${c}_1gpu.json
done
python - <<'PY'
import json
for c in ("c1", "c2", "c3", "c4", "c5"):
    try:
        l = json.load(open("gpurun_out/r2_bench_%s_1gpu.json" % c))
        print(c, round(l["value"], 1), round(l["ms_per_step"], 3), round(l["e2e"]["value"], 1), l.get("loss_parity"), round(l["roofline"]["step_tflops"], 1))
    except Exception as e:
        print(c, "failed", e)
PY
