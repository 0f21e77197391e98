"""Standalone GPU check of the tcgen05 conv engine against torch (fp32 math on the same
bf16-rounded operands).  Run on a B200: python tests/gpu_check_conv.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mtl_ssl_b200 import ops_conv as oc

torch.manual_seed(0)
dev = "cuda"


def ref_fprop(x, w, stride, pad, dil, out_hw):
    xf = x.float().permute(0, 3, 1, 2)
    wf = w.float().permute(0, 3, 1, 2)
    P, Q = out_hw
    # explicit begin pad, generous end pad then crop
    xp = F.pad(xf, (pad[1], pad[1] + stride * 2, pad[0], pad[0] + stride * 2))
    y = F.conv2d(xp, wf, stride=stride, dilation=dil)[:, :, :P, :Q]
    return y.permute(0, 2, 3, 1).contiguous()


def check(name, got, want, tol):
    err = (got.float() - want).abs().max().item()
    scale = want.abs().max().item() + 1e-6
    ok = err <= tol * scale
    print("%-60s max_err %.4g (scale %.4g) %s" % (name, err, scale, "OK" if ok else "FAIL"), flush=True)
    return ok


def run_case(N, H, W, C, K, R, stride, pad, dil=1, bn=0):
    S = R
    P = oc.out_size(H, R, stride, pad, pad, dil)
    Q = oc.out_size(W, S, stride, pad, pad, dil)
    x = torch.randn(N, H, W, C, device=dev).bfloat16()
    w = (torch.randn(K, R, S, C, device=dev) / (R * S * C) ** 0.5).bfloat16()
    bias = torch.randn(K, device=dev)
    res = torch.randn(N, P, Q, K, device=dev).bfloat16()
    ok = True
    tag = "N%d %dx%d C%d K%d k%d s%d p%d bn%d" % (N, H, W, C, K, R, stride, pad, bn)
    # fprop
    y = oc.conv_fprop(x, w, stride, (pad, pad), dil, (P, Q), bias=bias, res=res, relu=True, force_bn=bn)
    yr = torch.relu(ref_fprop(x, w, stride, (pad, pad), dil, (P, Q)) + bias + res.float())
    ok &= check("fprop " + tag, y, yr, 1e-2)
    y32 = oc.conv_fprop(x, w, stride, (pad, pad), dil, (P, Q), out_dtype=torch.float32, force_bn=bn)
    ok &= check("fprop32 " + tag, y32, ref_fprop(x, w, stride, (pad, pad), dil, (P, Q)), 2e-3)
    # dgrad / wgrad via autograd of the reference
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    yref = ref_fprop(xf, wf, stride, (pad, pad), dil, (P, Q))
    dy = torch.randn(N, P, Q, K, device=dev).bfloat16()
    yref.backward(dy.float())
    mask = torch.randn(N, H, W, C, device=dev).bfloat16()
    res2 = torch.randn(N, H, W, C, device=dev).bfloat16()
    dx = oc.conv_dgrad(dy, w, (N, H, W, C), stride, (pad, pad), dil, res=res2, mask=mask, force_bn=bn)
    dxr = torch.where(mask.float() > 0, xf.grad + res2.float(), torch.zeros_like(xf.grad))
    ok &= check("dgrad " + tag, dx, dxr, 1e-2)
    if C % 64 == 0:
        dw = torch.zeros(K, R, S, C, device=dev)
        rs = torch.rand(K, device=dev) + 0.5
        oc.conv_wgrad(dy, x, dw, stride, (pad, pad), dil, rowscale=rs, alpha=0.5)
        dwr = 0.5 * rs[:, None, None, None] * wf.grad
        ok &= check("wgrad " + tag, dw, dwr, 2e-3)
    return ok


cases = [
    # N, H, W, C, K, R, stride, pad
    (1, 16, 16, 64, 64, 1, 1, 0),
    (1, 38, 63, 1024, 256, 1, 1, 0),
    (1, 38, 63, 256, 1024, 1, 1, 0),
    (4, 7, 7, 512, 512, 3, 1, 1),
    (1, 38, 63, 256, 256, 3, 1, 1),
    (1, 75, 125, 128, 128, 3, 2, 1),
    (1, 38, 63, 1024, 512, 3, 1, 1),
    (1, 38, 63, 512, 48, 1, 1, 0),
    (1, 38, 63, 512, 24, 1, 1, 0),
    (64, 7, 7, 1024, 2048, 1, 1, 0),
    (64, 7, 7, 2048, 512, 1, 1, 0),
    (3, 9, 11, 192, 72, 3, 1, 1),
    (2, 17, 17, 320, 384, 3, 2, 0),
]
allok = True
for c in cases:
    try:
        allok &= run_case(*c)
    except Exception as e:
        print("EXC", c, repr(e), flush=True)
        allok = False
        break
# forced tile widths
for bn in (64, 128, 256):
    try:
        allok &= run_case(8, 7, 7, 512, 512, 3, 1, 1, 1, bn)
        allok &= run_case(8, 7, 7, 1024, 512, 1, 1, 0, 1, bn)
    except Exception as e:
        print("EXC bn", bn, repr(e), flush=True)
        allok = False
        break
torch.cuda.synchronize()

# quick timing of the dominant shapes (block4 on 256 ROIs)
def bench(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

if allok:
    for (N, C, K, R) in [(256, 1024, 2048, 1), (256, 2048, 512, 1), (256, 512, 512, 3), (256, 512, 2048, 1),
                         (1280, 2048, 512, 1), (1280, 512, 512, 3), (1280, 512, 2048, 1)]:
        pad = R // 2
        x = torch.randn(N, 7, 7, C, device=dev).bfloat16()
        w = torch.randn(K, R, R, C, device=dev).bfloat16()
        dy = torch.randn(N, 7, 7, K, device=dev).bfloat16()
        dw = torch.zeros(K, R, R, C, device=dev)
        fl = 2.0 * N * 49 * C * K * R * R
        t = bench(lambda: oc.conv_fprop(x, w, 1, (pad, pad)))
        print("fprop N%d C%d K%d k%d: %.3f ms  %.1f TFLOP/s" % (N, C, K, R, t, fl / t / 1e9), flush=True)
        t = bench(lambda: oc.conv_dgrad(dy, w, (N, 7, 7, C), 1, (pad, pad)))
        print("dgrad N%d C%d K%d k%d: %.3f ms  %.1f TFLOP/s" % (N, C, K, R, t, fl / t / 1e9), flush=True)
        t = bench(lambda: oc.conv_wgrad(dy, x, dw, 1, (pad, pad)))
        print("wgrad N%d C%d K%d k%d: %.3f ms  %.1f TFLOP/s" % (N, C, K, R, t, fl / t / 1e9), flush=True)
print("ALL OK" if allok else "FAILED")
sys.exit(0 if allok else 1)
