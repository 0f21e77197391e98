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
run chunk40 MTL_WGRAD_CHUNK_CTAS=40
run chunk40_def40 MTL_WGRAD_CHUNK_CTAS=40 MTL_DEFERRED_CTAS=40
run def40 MTL_DEFERRED_CTAS=40
run default2 X=2
python bench.py --steps 20 --warmup 4 --skip-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.readline()); print('single', round(l['ms_per_step'],3), round(l['value'],1))"
