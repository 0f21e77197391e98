#!/bin/bash
# Developer experiment (GPU): forward-only conv3 with / without the fused spatial mean (1280 ROIs), + tests.
python - <<'PY'
import sys; sys.path.insert(0, "tools"); sys.argv = ["x"]
import torch
import sweep_conv as s
from mtl_ssl_b200 import ops_conv as oc
N, H, C, K = 1280, 7, 512, 2048
x = torch.randn(N, H, H, C, device="cuda").bfloat16()
w = (torch.randn(K, 1, 1, C, device="cuda") * 0.05).bfloat16()
res = torch.randn(N, H, H, K, device="cuda").bfloat16()
bias = torch.randn(K, device="cuda")
y = torch.empty(N, H, H, K, device="cuda", dtype=torch.bfloat16)
part = torch.empty(oc.pool_partial_rows(N * H * H), K, device="cuda")
print("stored  %.1f us" % s.timed(lambda: oc.conv_fprop(x, w, bias=bias, res=res, relu=True, out=y)))
print("pooled  %.1f us" % s.timed(lambda: oc.conv_fprop(x, w, bias=bias, res=res, relu=True, out=y, pool_out=part, pool_hw=49)))
PY
python -m pytest tests/test_gpu_conv_engine.py tests/test_gpu_full_size_parity.py tests/test_gpu_train_step.py tests/test_gpu_zz_inference_refiner.py -x -q 2>&1 | tail -4
