#!/bin/bash
# Developer experiment (GPU): full ncu capture of the second-stage conv3 GEMM (1280 ROIs, K loop of 8 steps), stored output
# and pooled output; details page as text into gpurun_out/.
mkdir -p gpurun_out
RES=1 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -f -o /tmp/conv3_after \
  python tools/bench_conv.py fprop 1280 7 7 512 2048 1 1 2 > gpurun_out/ncu_conv3_after.log 2>&1
ncu -i /tmp/conv3_after.ncu-rep --page details > gpurun_out/r2_ncu_conv3_res1_C512_after.txt 2>/dev/null
ncu -i /tmp/conv3_after.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); d=dict(zip(rows[0],rows[2]))
for k in ('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_uniform.sum'):
    print(k, d.get(k))
for k,v in d.items():
    if 'tensor' in k and 'pct' in k: print(k, v)
"
