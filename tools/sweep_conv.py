"""Developer tool (GPU): sweep tile width / split-K / pipeline depth of the conv engine on the layer shapes of
the ResNet-101 step.  usage: sweep_conv.py [quick]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mtl_ssl_b200 import ops_conv as oc
from mtl_ssl_b200._lib import MtlError

ITERS = 20


def timed(fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(ITERS):
            fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / ITERS)
    return best * 1e3


def sweep(mode, N, H, W, C, K, R, res, cfgs):
    pad = (R - 1) // 2
    x = torch.randn(N, H, W, C, device="cuda").bfloat16()
    w = (torch.randn(K, R, R, C, device="cuda") * 0.05).bfloat16()
    dy = torch.randn(N, H, W, K, device="cuda").bfloat16()
    r_y = torch.randn(N, H, W, K, device="cuda").bfloat16()
    r_x = torch.randn(N, H, W, C, device="cuda").bfloat16()
    bias = torch.randn(K, device="cuda")
    y = torch.empty(N, H, W, K, device="cuda", dtype=torch.bfloat16)
    dx = torch.empty(N, H, W, C, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * N * H * W * K * R * R * C
    out = []
    for cfg in cfgs:
        bn, sp, st = cfg[:3]
        cl = cfg[3] if len(cfg) > 3 else 0
        def run():
            if mode == "fprop":
                oc.conv_fprop(x, w, 1, (pad, pad), 1, (H, W), bias=bias, res=r_y if res else None, relu=True, out=y,
                              force_bn=bn, force_splits=sp, force_stages=st, force_cluster=cl)
            else:
                oc.conv_dgrad(dy, w, (N, H, W, C), 1, (pad, pad), 1, res=r_x if res & 1 else None,
                              mask=r_x if res else None, out=dx, force_bn=bn, force_splits=sp, force_stages=st, force_cluster=cl)
        try:
            us = timed(run)
            out.append("bn%d/s%d/st%d/cl%d %.1fus %.0fTF" % (bn, sp, st, cl, us, fl / us / 1e6))
        except MtlError as e:
            out.append("bn%d/s%d/st%d ERR" % (bn, sp, st))
    print("%-5s N%-4d %dx%d C%-4d K%-4d k%d res%d | %s" % (mode, N, H, W, C, K, R, res, "  ".join(out)), flush=True)


def pairs():
    c = [(0, 1, 0, 1), (0, 1, 0, 2), (128, 1, 0, 1), (128, 1, 0, 2)]
    for n in (1280, 256, 64):
        sweep("fprop", n, 7, 7, 512, 512, 3, 0, c)
        sweep("fprop", n, 7, 7, 512, 2048, 1, 1, c)
        sweep("fprop", n, 7, 7, 2048, 512, 1, 0, c)
        sweep("fprop", n, 7, 7, 1024, 2048, 1, 0, c)
    sweep("dgrad", 256, 7, 7, 512, 512, 3, 2, c)
    sweep("dgrad", 256, 7, 7, 2048, 512, 1, 3, c)
    sweep("dgrad", 256, 7, 7, 512, 2048, 1, 2, c)
    sweep("dgrad", 256, 7, 7, 1024, 2048, 1, 2, c)
    sweep("fprop", 1, 75, 125, 128, 512, 1, 1, c)
    sweep("fprop", 1, 150, 250, 64, 256, 1, 1, c)


def kloop():
    """Per-K-iteration cost of the operand pipeline at the trunk's row count (M = 2394): time against the K depth for each
    tile width (slope = cost of one 64-deep K step, intercept = launch + fill + drain), and against the stage count."""
    T = (1, 38, 63)
    for bn in (64, 128, 256):
        for C in (256, 512, 1024, 2048, 4096):
            sweep("fprop", *T, C, 256, 1, 0, [(bn, 1, 0)])
    for bn in (64, 128, 256):
        sweep("fprop", *T, 1024, 256, 1, 0, [(bn, 1, st) for st in ((3, 4, 6, 8) if bn < 256 else (3, 4))])
        sweep("fprop", *T, 256, 256, 3, 0, [(bn, 1, st) for st in ((3, 4, 6, 8) if bn < 256 else (3, 4))])
    # wide output, short K (block3 conv3): drain-bound side
    for bn in (64, 128, 256):
        sweep("fprop", *T, 256, 1024, 1, 1, [(bn, 1, 0)])


def conv3():
    """Second-stage conv3 (K loop of 8 steps, 2048 outputs, residual + ReLU): what a 128 x 256 tile costs beyond its
    2.2 us of MMAs -- residual on / off, tile width, CTA pairs, and the K depth (slope = K step, intercept = drain)."""
    c = [(0, 1, 0, 1), (0, 1, 0, 2), (128, 1, 0, 1), (128, 1, 0, 2), (256, 1, 3, 1), (256, 1, 2, 1)]
    for n in (1280, 256):
        for res in (1, 0):
            sweep("fprop", n, 7, 7, 512, 2048, 1, res, c)
    for C in (128, 256, 512, 1024, 2048):
        sweep("fprop", 1280, 7, 7, C, 2048, 1, 1, [(0, 1, 0, 1)])
        sweep("fprop", 1280, 7, 7, C, 2048, 1, 0, [(0, 1, 0, 1)])


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "conv3":
        return conv3()
    if len(sys.argv) > 1 and sys.argv[1] == "pairs":
        return pairs()
    if len(sys.argv) > 1 and sys.argv[1] == "kloop":
        return kloop()
    T = (1, 38, 63)
    # trunk block3 (M = 2394): default vs split-K
    trunk_cfg = [(0, 1, 0), (64, 1, 0), (64, 2, 0), (128, 2, 0), (128, 3, 0), (128, 4, 0), (256, 2, 0), (256, 4, 0), (256, 7, 0)]
    sweep("fprop", *T, 256, 256, 3, 0, trunk_cfg)
    sweep("dgrad", *T, 256, 256, 3, 2, trunk_cfg)
    sweep("fprop", *T, 1024, 256, 1, 0, trunk_cfg)
    sweep("dgrad", *T, 256, 1024, 1, 2, trunk_cfg)          # dx [M,256] from dy [M,1024]
    c3 = [(0, 1, 0), (64, 1, 0), (128, 1, 0), (128, 2, 0), (256, 1, 0), (256, 2, 0), (256, 4, 0)]
    sweep("fprop", *T, 256, 1024, 1, 1, c3)
    sweep("dgrad", *T, 1024, 256, 1, 3, c3)                 # dx [M,1024] from dy [M,256], res + mask
    sweep("fprop", *T, 1024, 512, 3, 0, [(0, 1, 0), (128, 2, 0), (128, 3, 0), (256, 3, 0), (256, 7, 0)])
    b2 = (1, 75, 125)
    sweep("fprop", *b2, 128, 128, 3, 0, [(0, 1, 0), (128, 1, 0), (128, 2, 0)])
    sweep("fprop", *b2, 128, 512, 1, 1, [(0, 1, 0), (128, 1, 0), (256, 1, 0)])
    sweep("fprop", *b2, 512, 128, 1, 0, [(0, 1, 0), (128, 1, 0), (128, 2, 0)])
    # ROI tiles: residual ring depth vs pipeline depth; wave quantisation at 256 ROIs
    for n in (1280, 256, 64):
        sweep("fprop", n, 7, 7, 512, 2048, 1, 1, [(0, 1, 0), (256, 1, 4), (256, 1, 3), (128, 1, 0), (128, 1, 4)])
    sweep("dgrad", 256, 7, 7, 2048, 512, 1, 3, [(0, 1, 0), (256, 1, 3), (128, 1, 0), (128, 1, 4)])
    sweep("dgrad", 256, 7, 7, 1024, 512, 1, 2, [(0, 1, 0), (128, 1, 0), (128, 1, 4)])
    q = [(0, 1, 0), (128, 1, 0), (256, 2, 0), (256, 3, 0), (128, 2, 0)]
    sweep("fprop", 256, 7, 7, 512, 512, 3, 0, q)
    sweep("dgrad", 256, 7, 7, 512, 512, 3, 2, q + [(256, 1, 4), (256, 3, 4)])
    sweep("fprop", 256, 7, 7, 2048, 512, 1, 0, q)
    sweep("dgrad", 256, 7, 7, 512, 2048, 1, 2, q)
    sweep("fprop", 256, 7, 7, 1024, 512, 1, 0, q)
    q64 = [(0, 1, 0), (128, 1, 0), (128, 2, 0), (256, 2, 0), (256, 3, 0), (256, 6, 0)]
    sweep("fprop", 64, 7, 7, 512, 512, 3, 0, q64)
    sweep("dgrad", 64, 7, 7, 512, 512, 3, 2, q64)
    sweep("fprop", 64, 7, 7, 2048, 512, 1, 0, q64)
    sweep("fprop", 64, 7, 7, 1024, 2048, 1, 0, [(0, 1, 0), (128, 1, 0), (256, 1, 0), (256, 2, 0)])


if __name__ == "__main__":
    main()
