"""GPU numerics of the tcgen05 conv engine (mtl_conv_tc) against a plain PyTorch fp32 reference of the same
op on the same bf16-rounded operands: fprop / dgrad / wgrad, 1x1, square and rectangular filters, strides,
odd channel counts, channel-slice inputs/outputs (concatenated branches), fused epilogues.
Tolerances: bf16 outputs 1e-2 relative to the tensor scale (half a bf16 ulp is 2^-9); fp32 outputs 2e-3."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def ref_conv(x, w, stride, pad, P, Q):
    xf = x.float().permute(0, 3, 1, 2)
    wf = w.float().permute(0, 3, 1, 2)
    xp = F.pad(xf, (pad[1], pad[1] + stride * 8, pad[0], pad[0] + stride * 8))
    return F.conv2d(xp, wf, stride=stride)[:, :, :P, :Q].permute(0, 2, 3, 1).contiguous()


def close(got, want, tol):
    err = (got.float() - want).abs().max().item()
    scale = want.abs().max().item() + 1e-6
    assert err <= tol * scale, (err, scale)


CASES = [
    # N, H, W, C, K, R, S, stride, pad_h, pad_w
    (64, 7, 7, 1024, 512, 1, 1, 1, 0, 0),
    (8, 7, 7, 512, 512, 3, 3, 1, 1, 1),
    (3, 9, 11, 192, 72, 3, 3, 1, 1, 1),
    (2, 17, 17, 320, 384, 3, 3, 2, 0, 0),
    (2, 19, 23, 64, 256, 1, 1, 2, 0, 0),          # strided 1x1 (shortcut)
    (2, 19, 23, 128, 128, 3, 3, 2, 1, 1),         # conv2d_same stride 2
    (2, 17, 17, 128, 160, 1, 7, 1, 0, 3),         # block17 1x7
    (2, 17, 17, 160, 192, 7, 1, 1, 3, 0),         # block17 7x1 (C = 2.5 x 64)
    (4, 8, 8, 192, 224, 1, 3, 1, 0, 1),           # block8 1x3
    (4, 8, 8, 224, 256, 3, 1, 1, 1, 0),           # block8 3x1 (C = 3.5 x 64)
    (1, 35, 35, 48, 64, 5, 5, 1, 2, 2),           # Mixed_5b 5x5 on 48 channels
    (1, 35, 35, 32, 48, 3, 3, 1, 1, 1),           # block35 3x3 on 32 channels
    (1, 35, 35, 80, 192, 3, 3, 1, 1, 1),          # Conv2d_4a (C = 80)
    (2, 8, 8, 2080, 192, 1, 1, 1, 0, 0),          # block8 1x1 on 2080 channels
    (2, 17, 17, 288, 320, 3, 3, 2, 0, 0),         # Mixed_7a strided 3x3, C = 4.5 x 64 (gather path)
]


@pytest.mark.parametrize("case", CASES)
def test_fprop_dgrad_wgrad(case):
    from mtl_ssl_b200 import ops_conv as oc
    N, H, W, C, K, R, S, stride, ph, pw = case
    torch.manual_seed(hash(case) % 1000)
    P = (H + 2 * ph - R) // stride + 1
    Q = (W + 2 * pw - S) // stride + 1
    dev = "cuda"
    x = torch.randn(N, H, W, C, device=dev).bfloat16()
    w = (torch.randn(K, R, S, C, device=dev) / (R * S * C) ** 0.5).bfloat16()
    bias = torch.randn(K, device=dev)
    res = torch.randn(N, P, Q, K, device=dev).bfloat16()
    y = oc.conv_fprop(x, w, stride, (ph, pw), 1, (P, Q), bias=bias, res=res, relu=True, bias_scale=0.5)
    want = torch.relu(ref_conv(x, w, stride, (ph, pw), P, Q) + 0.5 * bias + res.float())
    close(y, want, 1e-2)
    y32 = oc.conv_fprop(x, w, stride, (ph, pw), 1, (P, Q), out_dtype=torch.float32)
    close(y32, ref_conv(x, w, stride, (ph, pw), P, Q), 2e-3)
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    yr = ref_conv(xf, wf, stride, (ph, pw), P, Q)
    dy = torch.randn(N, P, Q, K, device=dev).bfloat16()
    yr.backward(dy.float())
    mask = torch.randn(N, H, W, C, device=dev).bfloat16()
    res2 = torch.randn(N, H, W, C, device=dev).bfloat16()
    dx = oc.conv_dgrad(dy, w, (N, H, W, C), stride, (ph, pw), 1, res=res2, mask=mask)
    close(dx, torch.where(mask.float() > 0, xf.grad + res2.float(), torch.zeros_like(xf.grad)), 1e-2)
    dw = torch.zeros(K, R, S, C, device=dev)
    rs = torch.rand(K, device=dev) + 0.5
    oc.conv_wgrad(dy, x, dw, stride, (ph, pw), 1, rowscale=rs, alpha=0.5)
    oc.conv_wgrad(dy, x, dw, stride, (ph, pw), 1, rowscale=rs, alpha=0.5)       # accumulates
    close(dw, wf.grad * rs[:, None, None, None], 3e-3)


SPLIT_CASES = [
    # N, H, W, C, K, R, stride, bn, splits, stages
    (1, 38, 63, 256, 256, 3, 1, 256, 7, 0),       # block3 conv2: 19 tiles x 7 splits
    (1, 38, 63, 256, 256, 3, 1, 128, 3, 0),
    (1, 38, 63, 1024, 256, 1, 1, 64, 2, 0),       # block3 conv1
    (1, 38, 63, 256, 1024, 1, 1, 256, 2, 0),      # block3 conv3 (residual), 4 K iterations
    (256, 7, 7, 256, 512, 3, 1, 256, 3, 0),       # 196 tiles x 3 splits > 148 CTAs: several tiles per CTA
    (256, 7, 7, 256, 1024, 1, 1, 256, 1, 0),      # no split, residual ring crossing tile boundaries
    (256, 7, 7, 256, 1024, 1, 1, 128, 1, 4),      # explicit pipeline depth
    (256, 7, 7, 128, 200, 1, 1, 64, 2, 0),        # N not a multiple of 32, ragged last column chunk
    (2, 19, 23, 128, 128, 3, 2, 64, 3, 0),        # gather-warp path (stride 2) with split K
]


@pytest.mark.parametrize("case", SPLIT_CASES)
def test_split_k_and_residual_ring(case):
    """FPROP / DGRAD with forced split-K (fp32 workspace + last-arriver epilogue) and the deep residual /
    mask TMA ring; every call runs twice to prove the workspace is left zeroed."""
    from mtl_ssl_b200 import ops_conv as oc
    N, H, W, C, K, R, stride, bn, splits, stages = case
    torch.manual_seed(hash(case) % 1000)
    pad = (R - 1) // 2
    P = (H + 2 * pad - R) // stride + 1
    Q = (W + 2 * pad - R) // stride + 1
    dev = "cuda"
    x = torch.randn(N, H, W, C, device=dev).bfloat16()
    w = (torch.randn(K, R, R, C, device=dev) / (R * R * C) ** 0.5).bfloat16()
    bias = torch.randn(K, device=dev)
    res = torch.randn(N, P, Q, K, device=dev).bfloat16()
    want = torch.relu(ref_conv(x, w, stride, (pad, pad), P, Q) + bias + res.float())
    for _ in range(2):
        y = oc.conv_fprop(x, w, stride, (pad, pad), 1, (P, Q), bias=bias, res=res, relu=True,
                          force_bn=bn, force_splits=splits, force_stages=stages)
        close(y, want, 1e-2)
    y0 = oc.conv_fprop(x, w, stride, (pad, pad), 1, (P, Q), force_bn=bn, force_splits=splits, force_stages=stages)
    close(y0, ref_conv(x, w, stride, (pad, pad), P, Q), 1e-2)
    xf = x.float().requires_grad_(True)
    yr = ref_conv(xf, w.float(), stride, (pad, pad), P, Q)
    dy = torch.randn(N, P, Q, K, device=dev).bfloat16()
    yr.backward(dy.float())
    mask = torch.randn(N, H, W, C, device=dev).bfloat16()
    res2 = torch.randn(N, H, W, C, device=dev).bfloat16()
    want_dx = torch.where(mask.float() > 0, xf.grad + res2.float(), torch.zeros_like(xf.grad))
    for _ in range(2):
        dx = oc.conv_dgrad(dy, w, (N, H, W, C), stride, (pad, pad), 1, res=res2, mask=mask,
                           force_bn=bn, force_splits=splits, force_stages=stages)
        close(dx, want_dx, 1e-2)
    dx1 = oc.conv_dgrad(dy, w, (N, H, W, C), stride, (pad, pad), 1, mask=mask, force_bn=bn, force_splits=splits)
    close(dx1, torch.where(mask.float() > 0, xf.grad, torch.zeros_like(xf.grad)), 1e-2)
    assert splits == 1 or oc._ws_by_stream, "forced split-K did not get a workspace"
    for buf in oc._ws_by_stream.values():
        assert not buf.any(), "split-K workspace not left zeroed"


PAIR_CASES = [
    # N, H, W, C, K, R, bn   (stride 1; M = N*H*W rows)
    (256, 7, 7, 256, 512, 1, 256),      # 98 m-tiles -> 49 pairs, residual + mask rings
    (255, 7, 7, 128, 512, 3, 256),      # 98 m-tiles with a ragged last tile, TMA im2col operand
    (101, 7, 7, 192, 384, 1, 128),      # 39 m-tiles: odd -> one padding CTA; C = 3 x 64, N = 3 x 128
    (65, 7, 7, 64, 200, 3, 128),        # N not a multiple of 32
    (3, 38, 63, 128, 256, 3, 256),      # trunk-shaped rows, image borders inside the pair
]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_cta_pairs_with_multicast_weights(case):
    """FPROP / DGRAD on two-CTA clusters (each CTA multicasts half of the weight tile) against the fp32
    reference and, bit for bit, against the single-CTA schedule (same tiles, same accumulation order)."""
    from mtl_ssl_b200 import ops_conv as oc
    N, H, W, C, K, R, bn = case
    torch.manual_seed(hash(case) % 1000)
    pad = (R - 1) // 2
    dev = "cuda"
    x = torch.randn(N, H, W, C, device=dev).bfloat16()
    w = (torch.randn(K, R, R, C, device=dev) / (R * R * C) ** 0.5).bfloat16()
    bias = torch.randn(K, device=dev)
    res = torch.randn(N, H, W, K, device=dev).bfloat16()
    want = torch.relu(ref_conv(x, w, 1, (pad, pad), H, W) + bias + res.float())
    y2 = oc.conv_fprop(x, w, 1, (pad, pad), 1, (H, W), bias=bias, res=res, relu=True, force_bn=bn, force_cluster=2)
    y1 = oc.conv_fprop(x, w, 1, (pad, pad), 1, (H, W), bias=bias, res=res, relu=True, force_bn=bn, force_cluster=1)
    close(y2, want, 1e-2)
    assert torch.equal(y1, y2)
    xf = x.float().requires_grad_(True)
    yr = ref_conv(xf, w.float(), 1, (pad, pad), H, W)
    dy = torch.randn(N, H, W, K, device=dev).bfloat16()
    yr.backward(dy.float())
    mask = torch.randn(N, H, W, C, device=dev).bfloat16()
    res2 = torch.randn(N, H, W, C, device=dev).bfloat16()
    want_dx = torch.where(mask.float() > 0, xf.grad + res2.float(), torch.zeros_like(xf.grad))
    if C >= 128:
        bnd = 256 if C >= 256 else 128
        dx2 = oc.conv_dgrad(dy, w, (N, H, W, C), 1, (pad, pad), 1, res=res2, mask=mask, force_bn=bnd, force_cluster=2)
        dx1 = oc.conv_dgrad(dy, w, (N, H, W, C), 1, (pad, pad), 1, res=res2, mask=mask, force_bn=bnd, force_cluster=1)
        close(dx2, want_dx, 1e-2)
        assert torch.equal(dx1, dx2)


@pytest.mark.parametrize("R,S", [(1, 1), (3, 3), (1, 7)])
def test_channel_slices(R, S):
    """Branch outputs written straight into a concat buffer; gradients read from / masked by slices of it."""
    from mtl_ssl_b200 import ops_conv as oc
    torch.manual_seed(R * 10 + S)
    N, H, W, C, K, Ct, c0 = 2, 17, 17, 128, 192, 448, 64
    ph, pw = (R - 1) // 2, (S - 1) // 2
    dev = "cuda"
    x = torch.randn(N, H, W, C, device=dev).bfloat16()
    w = (torch.randn(K, R, S, C, device=dev) / (R * S * C) ** 0.5).bfloat16()
    cat = torch.full((N, H, W, Ct), 7.0, device=dev).bfloat16()
    sl = cat[..., c0:c0 + K]
    oc.conv_fprop(x, w, 1, (ph, pw), 1, (H, W), relu=True, out=sl)
    want = torch.relu(ref_conv(x, w, 1, (ph, pw), H, W))
    close(sl, want, 1e-2)
    assert (cat[..., :c0] == 7).all() and (cat[..., c0 + K:] == 7).all()
    dcat = torch.randn(N, H, W, Ct, device=dev).bfloat16()
    dsl = dcat[..., c0:c0 + K]
    xf = x.float().requires_grad_(True)
    wf = w.float().requires_grad_(True)
    ref_conv(xf, wf, 1, (ph, pw), H, W).backward(dsl.float())
    xcat = torch.randn(N, H, W, Ct, device=dev).bfloat16()                 # mask living in another concat buffer
    msl = xcat[..., 32:32 + C]
    dx = oc.conv_dgrad(dsl, w, (N, H, W, C), 1, (ph, pw), 1, mask=msl)
    close(dx, torch.where(msl.float() > 0, xf.grad, torch.zeros_like(xf.grad)), 1e-2)
    dw = torch.zeros(K, R, S, C, device=dev)
    oc.conv_wgrad(dsl, x, dw, 1, (ph, pw), 1)
    close(dw, wf.grad, 3e-3)


def test_relu6_and_mask_hi():
    from mtl_ssl_b200 import ops_conv as oc
    torch.manual_seed(5)
    N, H, W, C, K = 2, 9, 9, 64, 128
    x = (torch.randn(N, H, W, C, device="cuda") * 3).bfloat16()
    w = (torch.randn(K, 1, 1, C, device="cuda") * 0.3).bfloat16()
    y = oc.conv_fprop(x, w, relu=2)
    close(y, torch.clamp(ref_conv(x, w, 1, (0, 0), H, W), 0, 6), 1e-2)
    dy = torch.randn(N, H, W, K, device="cuda").bfloat16()
    mask = (torch.rand(N, H, W, C, device="cuda") * 8 - 1).bfloat16()
    dx = oc.conv_dgrad(dy, w, (N, H, W, C), mask=mask, mask_hi=6.0)
    full = oc.conv_dgrad(dy, w, (N, H, W, C))
    alive = (mask.float() > 0) & (mask.float() < 6)
    assert torch.equal(dx, torch.where(alive, full, torch.zeros_like(full)))


def test_avgpool3x3_same_fwd_bwd():
    from mtl_ssl_b200 import ops
    torch.manual_seed(6)
    N, H, W, C = 2, 7, 9, 32
    x = torch.randn(N, H, W, C, device="cuda").bfloat16()
    y = torch.empty_like(x)
    ops.call("mtl_avgpool3x3_same", x, N, H, W, C, 0, y)
    # reference on the CPU in NCHW-contiguous layout (torch's CUDA channels-last backward of
    # count_include_pad=False pooling returned wrong indices on this build)
    xr = x.float().cpu().permute(0, 3, 1, 2).contiguous().requires_grad_(True)
    want = F.avg_pool2d(xr, 3, 1, 1, count_include_pad=False)
    close(y.cpu(), want.detach().permute(0, 2, 3, 1), 1e-2)
    dy = torch.randn(N, H, W, C, device="cuda").bfloat16()
    want.backward(dy.float().cpu().permute(0, 3, 1, 2).contiguous())
    dx = torch.empty_like(x)
    ops.call("mtl_avgpool3x3_same", dy, N, H, W, C, 1, dx)
    close(dx.cpu(), xr.grad.permute(0, 2, 3, 1), 1e-2)


def test_grouped_wgrad_launches_match_individual_problems():
    """ops_conv.WgradCollector / ConvGroup: independent weight-gradient problems of one kernel instance share a persistent
    launch (the CTAs walk the concatenated tile space, K split per problem).  A bottleneck unit's three GEMMs at the
    trunk's size and at ROI-tile size, a 3x3 with C = 128 (another tile width -> its own group) and a strided 3x3 (gather
    path -> its own group); every result against the fp32 reference, twice (the second flush reuses the planned groups
    and accumulates)."""
    from mtl_ssl_b200 import ops_conv as oc
    dev = "cuda"
    torch.manual_seed(7)
    probs = [  # N, H, W, C, K, R, stride, pad
        (1, 38, 63, 1024, 256, 1, 1, 0), (1, 38, 63, 256, 256, 3, 1, 1), (1, 38, 63, 256, 1024, 1, 1, 0),
        (64, 7, 7, 512, 512, 3, 1, 1), (64, 7, 7, 512, 2048, 1, 1, 0), (64, 7, 7, 1024, 512, 1, 1, 0),
        (1, 75, 125, 128, 128, 3, 1, 1), (1, 75, 125, 128, 512, 1, 1, 0),
        (2, 19, 23, 128, 128, 3, 2, 1), (3, 9, 11, 192, 72, 3, 1, 1),
    ]
    data = []
    for (N, H, W, C, K, R, stride, pad) in probs:
        P = (H + 2 * pad - R) // stride + 1
        Q = (W + 2 * pad - R) // stride + 1
        x = torch.randn(N, H, W, C, device=dev).bfloat16()
        dy = (torch.randn(N, P, Q, K, device=dev) / (N * P * Q) ** 0.5).bfloat16()
        rs = torch.rand(K, device=dev) + 0.5
        xf = x.float().requires_grad_(True)
        wf = torch.zeros(K, R, R, C, device=dev, requires_grad=True)
        ref_conv(xf, wf, stride, (pad, pad), P, Q).backward(dy.float())
        data.append((x, dy, rs, torch.zeros(K, R, R, C, device=dev), wf.grad * rs[:, None, None, None], stride, pad))
    col = oc.WgradCollector(target_k_iters=24)
    for rep in (1, 2):
        with col:
            for x, dy, rs, dw, _want, stride, pad in data:
                oc.conv_wgrad(dy, x, dw, stride, (pad, pad), 1, rowscale=rs)
            assert not any(d[3].any() for d in data) or rep == 2         # deferred: nothing has run yet
        col.flush()
        torch.cuda.synchronize()
        for x, dy, rs, dw, want, stride, pad in data:
            close(dw, want * rep, 3e-3)
    sig, groups = col.groups[0]
    assert len(groups) >= 3 and sum(g.n for g in groups) == len(probs)
    assert max(g.n for g in groups) >= 5


def test_fprop_pooled_output_and_pooled_head():
    """conv3 of a forward-only tail (1x1, residual, ReLU) with the spatial mean fused (mtl_conv_args.pool_out): the
    partial row sums equal those of the stored output of the same launch without pooling, and mtl_head_fwd_pooled on them
    equals mtl_head_fwd on the stored maps (core/box_predictor.py:470-500 reduce_mean + fully connected)."""
    from mtl_ssl_b200 import ops
    from mtl_ssl_b200 import ops_conv as oc
    torch.manual_seed(5)
    for R, HW_side, C, K in ((37, 7, 512, 2048), (5, 6, 256, 1024), (64, 7, 128, 256)):
        x = torch.randn(R, HW_side, HW_side, C, device="cuda").bfloat16()
        w = (torch.randn(K, 1, 1, C, device="cuda") * 0.05).bfloat16()
        res = torch.randn(R, HW_side, HW_side, K, device="cuda").bfloat16()
        bias = torch.randn(K, device="cuda")
        y = torch.empty(R, HW_side, HW_side, K, device="cuda", dtype=torch.bfloat16)
        oc.conv_fprop(x, w, bias=bias, res=res, relu=True, out=y)
        hw = HW_side * HW_side
        rows = R * hw
        part = torch.full((oc.pool_partial_rows(rows), K), float("nan"), device="cuda")
        scratch = torch.zeros_like(y)
        oc.conv_fprop(x, w, bias=bias, res=res, relu=True, out=scratch, pool_out=part, pool_hw=hw,
                      force_bn=0 if K == 2048 else 128)
        torch.cuda.synchronize()
        assert not bool(scratch.any())                     # the maps are not written
        yf = y.float().reshape(rows, K)
        want = torch.zeros(oc.pool_partial_rows(rows), K, device="cuda")
        for g in range((rows + 31) // 32):
            lo, hi = 32 * g, min(32 * g + 32, rows)
            b = min((lo // hw + 1) * hw, hi)
            want[2 * g] = yf[lo:b].sum(0)
            if b < hi:
                want[2 * g + 1] = yf[b:hi].sum(0)
        assert torch.isfinite(part).all()
        assert torch.allclose(part, want, rtol=1e-5, atol=1e-4), (part - want).abs().max()
        n = 24
        hwt = (torch.randn(n, K, device="cuda") * 0.02).bfloat16()
        hb = torch.randn(n, device="cuda")
        p0 = torch.empty(R, K, device="cuda", dtype=torch.bfloat16); o0 = torch.empty(R, n, device="cuda")
        p1 = torch.empty_like(p0); o1 = torch.empty_like(o0)
        ops.call("mtl_head_fwd", y, R, hw, K, hwt, hb, n, p0, o0, n)
        ops.call("mtl_head_fwd_pooled", part, R, hw, K, hwt, hb, n, p1, o1, n)
        torch.cuda.synchronize()
        assert (p0.float() - p1.float()).abs().max() <= 0.01 * p0.float().abs().max()      # one bf16 ulp at most
        assert torch.allclose(o0, o1, rtol=2e-3, atol=2e-3), (o0 - o1).abs().max()
    try:                                  # windows shorter than a row group are refused
        oc.conv_fprop(x, w, bias=bias, res=res, relu=True, out=scratch, pool_out=part, pool_hw=16)
        refused = False
    except Exception:
        refused = True
    assert refused
