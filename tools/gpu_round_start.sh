#!/usr/bin/env bash
# One gpurun call for the start of a round: the device-side checks written after the previous round's GPU budget was
# spent come first (they have never run on a device), then the full parity suite, the bench line and the ncu launch list.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_round_start.sh'
# Everything lands in gpurun_out/ (merged back by gpurun); nothing here changes GPU clocks.
set -u
mkdir -p gpurun_out
python -m pytest tests/test_gpu_zz_inference_refiner.py tests/test_gpu_zz_raw_size_inputs.py \
    tests/test_gpu_zz_shape_buckets.py -m gpu -q > gpurun_out/s0_new_paths.log 2>&1
echo "new paths rc=$?" | tee -a gpurun_out/s0_new_paths.log
python -m pytest tests -m gpu -x -q > gpurun_out/s0_pytest.log 2>&1
echo "suite rc=$?" | tee -a gpurun_out/s0_pytest.log
python bench.py --steps 20 --warmup 3 > gpurun_out/s0_bench.log 2>&1
tail -c 600 gpurun_out/s0_bench.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 1300 -c 1500 --csv --log-file gpurun_out/s0_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-graph --skip-cpu > gpurun_out/s0_ncu_bench.log 2>&1
echo "ncu rc=$?"
