"""Developer tool (GPU): time / profile one conv shape.  usage: bench_conv.py MODE N H W C K R [stride] [iters]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mtl_ssl_b200 import ops_conv as oc

mode, N, H, W, C, K, R = sys.argv[1], *[int(v) for v in sys.argv[2:8]]
stride = int(sys.argv[8]) if len(sys.argv) > 8 else 1
iters = int(sys.argv[9]) if len(sys.argv) > 9 else 20
res_on = os.environ.get("RES", "1") == "1"
BNF = int(os.environ.get("BN", "0"))
pad = (R - 1) // 2
P, Q = oc.out_size(H, R, stride, pad, pad), oc.out_size(W, R, stride, pad, pad)
x = torch.randn(N, H, W, C, device="cuda").bfloat16()
w = (torch.randn(K, R, R, C, device="cuda") * 0.05).bfloat16()
dy = torch.randn(N, P, Q, K, device="cuda").bfloat16()
res = torch.randn(N, P, Q, K, device="cuda").bfloat16()
bias = torch.randn(K, device="cuda")
mask = torch.randn(N, H, W, C, device="cuda").bfloat16()
dw = torch.zeros(K, R, R, C, device="cuda")
scale = torch.rand(K, device="cuda")
y = torch.empty(N, P, Q, K, device="cuda", dtype=torch.bfloat16)
dx = torch.empty(N, H, W, C, device="cuda", dtype=torch.bfloat16)


def run():
    if mode == "fprop":
        oc.conv_fprop(x, w, stride, (pad, pad), 1, (P, Q), bias=bias, res=res if res_on else None, relu=True, out=y, force_bn=BNF)
    elif mode == "dgrad":
        oc.conv_dgrad(dy, w, (N, H, W, C), stride, (pad, pad), 1, mask=mask if res_on else None, out=dx, force_bn=BNF)
    else:
        oc.conv_wgrad(dy, x, dw, stride, (pad, pad), 1, rowscale=scale, force_bn=BNF, force_splits=int(os.environ.get('SPLITS','0')))


for _ in range(3):
    run()
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(iters):
        run()
g.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
fl = 2.0 * N * P * Q * K * R * R * C
print("%s N%d %dx%d C%d K%d k%d s%d: %.1f us  %.1f TFLOP/s" % (mode, N, H, W, C, K, R, stride, ms * 1e3, fl / ms / 1e9))
