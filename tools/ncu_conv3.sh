#!/bin/bash
# Developer experiment (GPU): full ncu capture of the drain-bound second-stage conv3 GEMM (1280 ROIs, K loop of 2 and 8 steps).
mkdir -p gpurun_out
i=0
for cfg in "1 128" "1 512" "0 128"; do
  set -- $cfg
  RES=$1 ncu --set full --clock-control none --import-source on -k regex:tc_gemm -s 3 -c 1 -f -o /tmp/conv3_$i \
    python tools/bench_conv.py fprop 1280 7 7 $2 2048 1 1 2 > gpurun_out/ncu_conv3_$i.log 2>&1
  ncu -i /tmp/conv3_$i.ncu-rep --page raw --csv > gpurun_out/r2_conv3_res$1_C$2.csv 2>/dev/null
  ncu -i /tmp/conv3_$i.ncu-rep --page details > gpurun_out/r2_conv3_res$1_C$2.txt 2>/dev/null
  i=$((i+1))
done
ls -la gpurun_out | tail -8
