#!/bin/bash
# Developer experiment (GPU x N): variants of the data-parallel step on one box.  usage: dp_ab.sh N
N=${1:-2}
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    tools/dp_ab.py --gpus $N --steps 20 --warmup 4 --skip-cpu > gpurun_out/dp_ab_$name.log 2>&1
  tail -1 gpurun_out/dp_ab_$name.log | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.readline()); print('$name', round(l['ms_per_step'],3), round(l['value'],1), round(l['e2e']['value'],1))
except Exception as e: print('$name', 'failed', e)"
}
run default X=1
run early72 MTL_HW_EARLY=1 MTL_DEFERRED_CTAS=72
run nosplit MTL_NO_SPLIT_TRUNK=1
for c in c5 c1; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
    bench.py --gpus $N --steps 20 --warmup 4 --skip-cpu --config $c > gpurun_out/dp_bench_$c.log 2>&1
  tail -1 gpurun_out/dp_bench_$c.log > gpurun_out/r2_bench_${c}_${N}gpu.json
  python -c "
import json
try:
    l=json.load(open('gpurun_out/r2_bench_${c}_${N}gpu.json')); print('$c', round(l['ms_per_step'],3), round(l['value'],1))
except Exception as e: print('$c failed', e)"
done
tail -1 gpurun_out/dp_ab_default.log > gpurun_out/r2_bench_c2_${N}gpu.json
python bench.py --steps 20 --warmup 4 --skip-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('single', round(l['ms_per_step'],3), round(l['value'],1))"
