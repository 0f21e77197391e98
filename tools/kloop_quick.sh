#!/bin/bash
# Developer experiment (GPU): K-loop slope / intercept of the trunk-sized GEMMs (M = 2394)
for bn in 64 256; do
  for C in 256 1024 4096; do BN=$bn RES=0 python tools/bench_conv.py fprop 1 38 63 $C 256 1; done
done
BN=64 RES=0 python tools/bench_conv.py fprop 1 38 63 256 256 3
RES=1 python tools/bench_conv.py dgrad 1 38 63 256 256 3
python tools/bench_conv.py wgrad 1 38 63 256 256 3
python tools/bench_conv.py wgrad 1 38 63 1024 256 1
RES=0 python tools/bench_conv.py fprop 256 7 7 512 512 3
RES=1 python tools/bench_conv.py dgrad 256 7 7 512 512 3
python tools/bench_conv.py wgrad 256 7 7 512 512 3
RES=1 python tools/bench_conv.py fprop 1280 7 7 512 2048 1
RES=0 python tools/bench_conv.py fprop 1280 7 7 512 512 3
