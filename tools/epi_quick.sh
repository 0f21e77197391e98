#!/bin/bash
# Developer experiment (GPU): streamlined drain on / off on the drain-bound shapes + the conv engine tests.
for v in 0 1; do
  echo "== MTL_NO_FAST_EPI=$v"
  if [ $v = 1 ]; then export MTL_NO_FAST_EPI=1; fi
  python - <<'PY'
import sys; sys.path.insert(0, "tools"); sys.argv = ["x"]
import sweep_conv as s
c = [(0, 1, 0, 1)]
for n in (1280, 256):
    for res in (1, 0):
        s.sweep("fprop", n, 7, 7, 512, 2048, 1, res, c)
s.sweep("fprop", 1280, 7, 7, 128, 2048, 1, 0, c)
s.sweep("fprop", 1280, 7, 7, 512, 512, 3, 0, c)
s.sweep("fprop", 1280, 7, 7, 2048, 512, 1, 0, c)
s.sweep("dgrad", 256, 7, 7, 2048, 512, 1, 3, c)
s.sweep("dgrad", 256, 7, 7, 512, 2048, 1, 2, c)
s.sweep("dgrad", 256, 7, 7, 512, 512, 3, 2, c)
s.sweep("fprop", 1, 38, 63, 256, 1024, 1, 1, c)
s.sweep("fprop", 1, 38, 63, 1024, 256, 1, 0, c)
s.sweep("fprop", 1, 38, 63, 256, 256, 3, 0, c)
s.sweep("dgrad", 1, 38, 63, 1024, 256, 1, 3, c)
PY
  python -m pytest tests/test_gpu_conv_engine.py -x -q 2>&1 | tail -3
done
