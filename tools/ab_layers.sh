#!/bin/bash
# Developer experiment (GPU): A/B of an environment knob over the step's representative layer shapes.  usage: ab_layers.sh VAR v1 v2 ...
var=$1; shift
for v in "$@"; do
  echo "== $var=$v"
  export $var=$v
  RES=0 python tools/bench_conv.py fprop 1 38 63 1024 256 1
  RES=0 python tools/bench_conv.py fprop 1 38 63 256 256 3
  RES=1 python tools/bench_conv.py dgrad 1 38 63 256 256 3
  python tools/bench_conv.py wgrad 1 38 63 256 256 3
  RES=0 python tools/bench_conv.py fprop 256 7 7 512 512 3
  RES=1 python tools/bench_conv.py dgrad 256 7 7 512 512 3
  python tools/bench_conv.py wgrad 256 7 7 512 512 3
  RES=1 python tools/bench_conv.py fprop 256 7 7 512 2048 1
  RES=1 python tools/bench_conv.py dgrad 256 7 7 2048 512 1
  python tools/bench_conv.py wgrad 256 7 7 512 2048 1
  python tools/bench_conv.py wgrad 256 7 7 2048 512 1
  RES=1 python tools/bench_conv.py fprop 1280 7 7 512 2048 1
  RES=0 python tools/bench_conv.py fprop 1280 7 7 512 512 3
  RES=0 python tools/bench_conv.py fprop 1280 7 7 2048 512 1
done
