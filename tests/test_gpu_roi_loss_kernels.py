"""GPU parity: ROI / pooling / loss / optimizer kernels against the torch-CPU fp32 oracle
(oracle/nn.py), through the C ABI.  Float tolerances are stated per test."""
import numpy as np
import pytest
import torch

from oracle import nn as ON
from oracle import boxes as OB

pytestmark = pytest.mark.gpu
F = np.float32


def bf(t):
    return t.to(torch.bfloat16)


def rand_norm_boxes(g, R):
    y = torch.sort(torch.rand(R, 2, generator=g), 1)[0]
    x = torch.sort(torch.rand(R, 2, generator=g), 1)[0]
    b = torch.stack([y[:, 0], x[:, 0], y[:, 1], x[:, 1]], 1)
    b[0] = torch.tensor([0.0, 0.0, 1.0, 1.0])
    if R > 2:
        b[1] = torch.tensor([-0.2, 0.3, 0.5, 1.3])      # partially outside -> extrapolation 0
        b[2] = torch.tensor([0.4, 0.4, 0.4, 0.4])       # degenerate
    return b


@pytest.mark.parametrize("crop", [7, 14, 1])
def test_crop_and_resize_fwd_bwd(crop):
    from mtl_ssl_b200 import ops
    g = torch.Generator().manual_seed(crop)
    B, H, W, C, R = 2, 13, 17, 64, 24
    feat = bf(torch.randn(B, H, W, C, generator=g))
    boxes = rand_norm_boxes(g, R)
    bi = torch.randint(0, B, (R,), generator=g).int()
    out = torch.empty(R, crop, crop, C, dtype=torch.bfloat16, device="cuda")
    ops.call("mtl_crop_and_resize_fwd", feat.cuda(), B, H, W, C, boxes.cuda(), bi.cuda(), R, crop, crop, out)
    fr = feat.float().requires_grad_(True)
    want = ON.crop_and_resize(fr, boxes, bi, (crop, crop))
    # output is rounded to bf16: half an ulp of bf16 = 2^-9 relative
    torch.testing.assert_close(out.float().cpu(), want.detach(), rtol=4e-3, atol=4e-3)
    dcrop = bf(torch.randn(R, crop, crop, C, generator=g))
    want.backward(dcrop.float())
    dfeat = torch.zeros(B, H, W, C, device="cuda")
    ops.call("mtl_crop_and_resize_bwd", dcrop.cuda(), B, H, W, C, boxes.cuda(), bi.cuda(), R, crop, crop, dfeat)
    torch.testing.assert_close(dfeat.cpu(), fr.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("k,s,pad", [(3, 2, "SAME"), (2, 2, "VALID"), (1, 2, "SAME"), (3, 1, "SAME")])
def test_maxpool_fwd_bwd(k, s, pad):
    from mtl_ssl_b200.nets import layers as L
    g = torch.Generator().manual_seed(k * 10 + s)
    N, H, W, C = 2, 15, 18, 16
    x = bf(torch.randint(-3, 4, (N, H, W, C), generator=g).float())     # many ties
    P, Q = L.max_pool_out_hw(H, W, k, s, pad)
    y = torch.empty(N, P, Q, C, dtype=torch.bfloat16, device="cuda")
    L.max_pool(x.cuda(), y, k, s, pad)
    xr = x.float().requires_grad_(True)
    want = ON.max_pool_tf(xr, k, s, pad)
    assert torch.equal(y.float().cpu(), want.detach())
    dy = bf(torch.randint(-4, 5, (N, P, Q, C), generator=g).float())
    want.backward(dy.float())
    dx = torch.empty(N, H, W, C, dtype=torch.bfloat16, device="cuda")
    L.max_pool_bwd(x.cuda(), dy.cuda(), dx, k, s, pad)
    # torch routes the gradient to the first maximum in scan order, like TF/cuDNN; integers -> exact
    assert torch.equal(dx.float().cpu(), xr.grad)


def test_avgpool_fwd_bwd():
    from mtl_ssl_b200 import ops
    g = torch.Generator().manual_seed(0)
    R, HW, C = 10, 49, 64
    x = bf(torch.randn(R, HW, C, generator=g))
    y = torch.empty(R, C, dtype=torch.bfloat16, device="cuda")
    ops.call("mtl_avgpool_fwd", x.cuda(), R, HW, C, y)
    torch.testing.assert_close(y.float().cpu(), x.float().mean(1), rtol=4e-3, atol=4e-3)
    dy = torch.randn(R, 72, generator=g)
    dx = torch.empty(R, HW, C, dtype=torch.bfloat16, device="cuda")
    ops.call("mtl_avgpool_bwd", dy.cuda(), 1, 72, x.cuda(), 0.0, R, HW, C, dx)
    want = (dy[:, None, :C] / HW) * (x.float() > 0)
    torch.testing.assert_close(dx.float().cpu(), want, rtol=4e-3, atol=1e-4)


def test_im2col_stem():
    from mtl_ssl_b200 import ops
    g = torch.Generator().manual_seed(0)
    B, H, W = 2, 21, 30
    img = torch.rand(B, H, W, 3, generator=g) * 255
    P, Q = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    out = torch.empty(B, P, Q, 160, dtype=torch.bfloat16, device="cuda")
    means = [123.68, 116.779, 103.939]
    ops.call("mtl_im2col_f32", img.cuda(), B, H, W, 3, 7, 7, 2, 3, 3, P, Q, means, 1.0, out, 160)
    x = (img - torch.tensor(means)).permute(0, 3, 1, 2)
    xp = torch.nn.functional.pad(x, (3, 3, 3, 3))
    cols = torch.nn.functional.unfold(xp, 7, stride=2)              # [B, 3*49 (c,r,s), P*Q]
    cols = cols.view(B, 3, 49, P, Q).permute(0, 3, 4, 2, 1).reshape(B, P, Q, 147)
    assert torch.equal(out[..., :147].cpu(), bf(cols))
    assert not out[..., 147:].any()


def test_rpn_loss_value_and_grad():
    from mtl_ssl_b200 import ops
    from oracle import assign as OA
    rng = np.random.default_rng(0)
    B, Hf, Wf, A, Gmax = 2, 12, 16, 12, 4
    anchors_all = OB.grid_anchors(Hf, Wf, [0.25, 0.5, 1.0, 2.0], [0.5, 1.0, 2.0])
    Hi, Wi = Hf * 16.0, Wf * 16.0
    kept, kidx = OB.prune_outside_window(anchors_all, (0, 0, Hi, Wi))
    kidx = kidx.astype(np.int32)
    Nk, ld, HW = len(kidx), A * 6, Hf * Wf
    rpn_out = (rng.standard_normal((B, HW, ld))).astype(F)
    gt = np.zeros((B, Gmax, 4), F); ng = np.array([2, 3], np.int32)
    for b in range(B):
        sel = rng.choice(Nk, ng[b], replace=False)
        gt[b, :ng[b]] = kept[sel] + rng.uniform(-3, 3, (ng[b], 4)).astype(F)
    keys = rng.random((B, Nk)).astype(F)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    match = torch.empty(B, Nk, dtype=torch.int32, device="cuda")
    rb = torch.zeros(B, Gmax, dtype=torch.int64, device="cuda")
    ops.call("mtl_iou_match", d(gt), d(ng), Gmax, d(kept), 0, None, B, Nk, 0.7, 0.3, 1, 1, match, None, rb)
    sampled = torch.empty(B, Nk, dtype=torch.uint8, device="cuda")
    counts = torch.zeros(B, 4, dtype=torch.int32, device="cuda")
    ops.call("mtl_balanced_sample", match, d(keys), B, Nk, 64, 0.5, sampled, counts)
    losses = torch.zeros(2, device="cuda")
    dout = torch.empty(B, HW, ld, dtype=torch.bfloat16, device="cuda")
    ops.call("mtl_rpn_loss", d(rpn_out), ld, 0, A * 4, A, HW, d(kidx), d(kept), Nk, d(gt), Gmax, match, sampled,
             counts, B, 2.0, 1.0, 3.0, losses, dout)
    # oracle (fmA:1591-1668)
    z = torch.from_numpy(rpn_out).requires_grad_(True)
    tot_loc = tot_obj = 0
    for b in range(B):
        t = OA.assign_proposal(kept, gt[b, :ng[b]])
        s = OA.balanced_subsample(t["cls_weights"] > 0, 64, t["cls_targets"][:, 0] > 0, 0.5, keys[b])
        assert np.array_equal(s, sampled[b].cpu().numpy().astype(bool))
        sf = torch.from_numpy(s.astype(F))
        enc = z[b][:, :A * 4].reshape(-1, 4)[torch.from_numpy(kidx).long()]
        logit = z[b][:, A * 4:].reshape(-1, 2)[torch.from_numpy(kidx).long()]
        loc = ON.smooth_l1(enc, torch.from_numpy(t["reg_targets"]), sf * torch.from_numpy(t["reg_weights"]), 3.0)
        onehot = torch.nn.functional.one_hot(torch.from_numpy(t["cls_targets"][:, 0]).long(), 2).float()
        obj = ON.softmax_ce(logit, onehot, sf)
        tot_loc = tot_loc + loc.sum() / sf.sum()
        tot_obj = tot_obj + obj.sum() / sf.sum()
    l_loc, l_obj = 2.0 * tot_loc / B, 1.0 * tot_obj / B
    (l_loc + l_obj).backward()
    np.testing.assert_allclose(losses.cpu().numpy(), [l_loc.item(), l_obj.item()], rtol=2e-5)
    # gradient is emitted in bf16 for the tensor-core backward: 2^-8 relative
    torch.testing.assert_close(dout.float().cpu(), z.grad, rtol=8e-3, atol=1e-7)


def test_box_classifier_and_softmax_ce_losses():
    from mtl_ssl_b200 import ops
    g = torch.Generator().manual_seed(1)
    B, P, K = 2, 32, 20
    K1, ld = K + 1, 104
    head = torch.randn(B * P, ld, generator=g)
    cls_t = torch.randint(0, K1, (B * P,), generator=g).int()
    reg_t = torch.randn(B * P, 4, generator=g)
    reg_w = (cls_t > 0).float()
    cls_w = torch.ones(B * P)
    nprop = torch.tensor([P, 20], dtype=torch.int32)
    losses = torch.zeros(2, device="cuda")
    dh = torch.full((B * P, ld), 7.0, device="cuda")
    ops.call("mtl_box_classifier_loss", head.cuda(), ld, 0, 4 * K, K, cls_t.cuda(), reg_t.cuda(), reg_w.cuda(),
             cls_w.cuda(), nprop.cuda(), B, P, 2.0, 1.0, losses, dh, ld)
    z = head.clone().requires_grad_(True)
    pad = (torch.arange(P)[None, :] < nprop[:, None]).reshape(-1).float()
    norm = (nprop.clamp(min=1).float() * B)[:, None].expand(B, P).reshape(-1)
    enc = torch.cat([torch.zeros(B * P, 1, 4), z[:, :4 * K].reshape(B * P, K, 4)], 1)
    sel = enc[torch.arange(B * P), cls_t.long()]
    loc = (ON.smooth_l1(sel, reg_t, reg_w, 1.0) / norm * pad).sum() * 2.0
    onehot = torch.nn.functional.one_hot(cls_t.long(), K1).float()
    cls = (ON.softmax_ce(z[:, 4 * K:4 * K + K1], onehot, cls_w) / norm * pad).sum() * 1.0
    (loc + cls).backward()
    np.testing.assert_allclose(losses.cpu().numpy(), [loc.item(), cls.item()], rtol=2e-5)
    got = dh.cpu()
    torch.testing.assert_close(got[:, :4 * K + K1], z.grad[:, :4 * K + K1], rtol=1e-4, atol=1e-7)
    assert (got[:, 4 * K + K1:] == 7.0).all()       # pad columns untouched
    # soft-label CE with column offsets, row weights, accumulate (closeness: fmA:1771-1789)
    logits = torch.randn(B * P, 24, generator=g)
    tgt = torch.rand(B * P, K1, generator=g)
    w = torch.rand(B * P, generator=g)
    loss = torch.zeros(1, device="cuda")
    dl = torch.ones(B * P, 24, device="cuda")
    ops.call("mtl_softmax_ce", logits.cuda(), 24, 1, K, tgt.cuda(), K1, 1, None, w.cuda(), None, P, 0, B * P, 0.3,
             loss, dl, 24, 1, 1)
    zl = logits.clone().requires_grad_(True)
    want = (ON.softmax_ce(zl[:, 1:K1], tgt[:, 1:], w)).sum() * 0.3
    want.backward()
    np.testing.assert_allclose(loss.item(), want.item(), rtol=2e-5)
    torch.testing.assert_close(dl.cpu()[:, 1:K1] - 1.0, zl.grad[:, 1:K1], rtol=1e-4, atol=1e-6)
    # hard targets with padding mask and per-image normaliser (refined loss: fmA:1795-1837)
    loss2 = torch.zeros(1, device="cuda")
    dl2 = torch.empty(B * P, 24, device="cuda")
    ops.call("mtl_softmax_ce", logits.cuda(), 24, 0, K1, None, 0, 0, cls_t.cuda(), cls_w.cuda(), nprop.cuda(), P, 1,
             B * P, 1.0 / B, loss2, dl2, 24, 0, 0)
    zl = logits.clone().requires_grad_(True)
    want2 = (ON.softmax_ce(zl[:, :K1], onehot, cls_w) / norm * pad).sum()
    want2.backward()
    np.testing.assert_allclose(loss2.item(), want2.item(), rtol=2e-5)
    torch.testing.assert_close(dl2.cpu()[:, :K1], zl.grad[:, :K1], rtol=1e-4, atol=1e-7)


def test_edgemask_head():
    from mtl_ssl_b200 import ops
    g = torch.Generator().manual_seed(2)
    B, H, W, C = 2, 9, 13, 128
    x = bf(torch.randn(B, H, W, C, generator=g))
    w = torch.randn(2, C, generator=g) * 0.1
    bias = torch.randn(2, generator=g) * 0.1
    gt = torch.rand(B, 2, 64, 64, generator=g)
    gt[:, 0] = (gt[:, 0] > 0.5).float()
    act = torch.empty(B * H * W, 2, device="cuda")
    ops.call("mtl_edgemask_fwd", x.cuda(), B * H * W, C, w.cuda(), bias.cuda(), act)
    xr = x.float().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    a = torch.tanh(xr.reshape(-1, C) @ wr.t() + br).reshape(B, H, W, 2)
    torch.testing.assert_close(act.cpu().reshape(B, H, W, 2), a.detach(), rtol=1e-4, atol=1e-5)
    loss = torch.zeros(1, device="cuda")
    dact = torch.empty(B, H, W, 2, device="cuda")
    ops.call("mtl_edgemask_loss", act, B, H, W, gt.cuda(), 64, 64, 1.0, loss, dact)
    pr = ON.resize_bilinear(a, (64, 64))
    tg = torch.stack([1 - gt[:, 0], gt[:, 0]], -1)
    want = (ON.softmax_ce(pr, tg) * gt[:, 1]).mean()
    want.backward()
    np.testing.assert_allclose(loss.item(), want.item(), rtol=2e-5)
    dfeat = torch.zeros(B, H, W, C, device="cuda")
    dw = torch.zeros(2, C, device="cuda"); db = torch.zeros(2, device="cuda")
    ops.call("mtl_edgemask_bwd", x.cuda(), B * H * W, C, w.cuda(), act, dact, dfeat, dw, db)
    torch.testing.assert_close(dfeat.cpu(), xr.grad, rtol=1e-3, atol=1e-7)
    torch.testing.assert_close(dw.cpu(), wr.grad, rtol=1e-3, atol=1e-6)
    torch.testing.assert_close(db.cpu(), br.grad, rtol=1e-3, atol=1e-6)


def test_refiner_fc_concat_colsum():
    from mtl_ssl_b200 import ops
    g = torch.Generator().manual_seed(3)
    M, K1, E = 16, 21, 5
    org = torch.randn(M, 104, generator=g); win = torch.randn(E * M, 24, generator=g)
    close = torch.randn(M, 24, generator=g)
    Kf = K1 * (E + 2)
    cat = torch.empty(M, Kf, device="cuda")
    ops.call("mtl_refine_concat", org.cuda(), 104, 80, win.cuda(), 24, 0, E, close.cuda(), 24, 0, M, K1, cat, Kf)
    w_ = win[:, :K1].reshape(E, M, K1).permute(1, 0, 2).reshape(M, E * K1)
    want = torch.cat([org[:, 80:80 + K1], w_, close[:, :K1].mean(0, keepdim=True).expand(M, K1)], 1)
    torch.testing.assert_close(cat.cpu(), want, rtol=1e-6, atol=1e-6)
    w = torch.randn(K1, Kf, generator=g) * 0.1; b = torch.randn(K1, generator=g)
    y = torch.empty(M, K1, device="cuda")
    ops.call("mtl_fc_fwd", cat, Kf, w.cuda(), b.cuda(), org.cuda()[:, 80:], 104, M, K1, Kf, y, K1)
    wy = want @ w.t() + b + org[:, 80:80 + K1]
    torch.testing.assert_close(y.cpu(), wy, rtol=1e-5, atol=1e-5)
    dy = torch.randn(M, K1, generator=g)
    dw = torch.zeros(K1, Kf, device="cuda"); db = torch.zeros(K1, device="cuda"); dx = torch.empty(M, Kf, device="cuda")
    ops.call("mtl_fc_bwd", cat, Kf, w.cuda(), dy.cuda(), K1, M, K1, Kf, dw, db, dx, Kf)
    torch.testing.assert_close(dw.cpu(), dy.t() @ want, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(db.cpu(), dy.sum(0), rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(dx.cpu(), dy @ w, rtol=1e-5, atol=1e-5)
    big = bf(torch.randn(1000, 72, generator=g))
    cs = torch.zeros(72, device="cuda")
    ops.call("mtl_colsum", big.cuda(), 0, 72, 1000, 72, 1.0, cs)
    torch.testing.assert_close(cs.cpu(), big.float().sum(0), rtol=1e-4, atol=1e-4)


def test_optimizer_matches_reference_update():
    """per-tensor clip_by_norm + momentum + L2 (slim/learning.py:282-301, optimizer_builder.py:49-53)."""
    from mtl_ssl_b200.runtime import ParamStore
    store = ParamStore()
    bn = store.add_bn("a/BatchNorm", 8, 1e-5)
    bn.gamma = torch.rand(8) + 0.5; bn.var = torch.rand(8) + 0.5
    a = store.add("a/weights", (8, 3, 3, 16), l2=1e-2, init=("normal", 1.0), fold=bn)
    grp = store.add_group([dict(name="h/box", shape=(16, 64), l2=1e-3, init=("normal", 1.0)),
                           dict(name="h/cls", shape=(5, 64), l2=1e-3, init=("normal", 1.0))])
    fz = store.add("frozen/weights", (4, 100), l2=1e-4, trainable=False, init=("normal", 1.0))
    bias = store.add("a/biases", (7,), init=("normal", 1.0))
    store.finalize("cuda", seed=5)
    ps = [a, grp[0], grp[1], fz, bias]
    w0 = [p.w.cpu().clone() for p in ps]
    gen = torch.Generator().manual_seed(9)
    grads = [torch.randn(p.shape, generator=gen) * s for p, s in zip(ps, [30.0, 0.1, 50.0, 1.0, 0.5])]
    mom = [torch.zeros_like(w) for w in w0]
    lr, mu, clip = 0.1, 0.9, 10.0
    store.set_hyper(lr, mu, clip)
    w_ref = [w.clone() for w in w0]
    for step in range(2):
        for p, gr in zip(ps, grads):
            p.g.copy_(gr.cuda())
        reg = store.stats_and_reg_loss(0.5)
        store.apply(0.5)
        want_reg = sum(p.l2 * 0.5 * (w ** 2).sum() for p, w in zip(ps, w_ref))
        np.testing.assert_allclose(reg.item(), want_reg.item(), rtol=1e-5)
        for i, p in enumerate(ps):
            if not p.trainable:
                continue
            gt_ = grads[i] * 0.5 + p.l2 * w_ref[i]
            gt_ = gt_ * clip / max(gt_.norm().item(), clip)
            mom[i] = mu * mom[i] + gt_
            w_ref[i] = w_ref[i] - lr * mom[i]
        for p, w in zip(ps, w_ref):
            torch.testing.assert_close(p.w.cpu(), w, rtol=1e-5, atol=1e-6)
            assert not p.trainable or not p.g.any()
    scale = (bn.gamma / torch.sqrt(bn.var + 1e-5))
    torch.testing.assert_close(a.wb.float().cpu(), (w_ref[0] * scale[:, None, None, None]).to(torch.bfloat16).float())
    torch.testing.assert_close(fz.wb.float().cpu(), w0[3].to(torch.bfloat16).float())
    hv = store.group_view(grp, 21, 64)
    torch.testing.assert_close(hv.float().cpu(), torch.cat([w_ref[1], w_ref[2]]).to(torch.bfloat16).float())


def test_update_pass_leaves_the_new_weight_norms():
    """apply_range(refresh_norms=True) == apply_range + stats_range: same weights, and the squared norms of the UPDATED
    weights (frozen tensors: unchanged weights) in the statistics the next step's regularisation loss reads."""
    from mtl_ssl_b200.runtime import ParamStore

    def make():
        store = ParamStore()
        a = store.add("a/weights", (8, 3, 3, 16), l2=1e-2, init=("normal", 1.0))
        b = store.add("b/weights", (33, 7), l2=1e-3, init=("normal", 1.0))             # scalar path (odd sizes)
        fz = store.add("frozen/weights", (4, 100), l2=1e-4, trainable=False, init=("normal", 1.0))
        c = store.add("c/weights", (300, 1, 1, 2048), l2=1e-4, init=("normal", 1.0))    # several chunks
        store.finalize("cuda", seed=5)
        store.set_hyper(0.1, 0.9, 10.0)
        gen = torch.Generator().manual_seed(3)
        for p in (a, b, fz, c):
            p.g.copy_(torch.randn(p.shape, generator=gen).cuda())
        return store, (a, b, fz, c)

    s1, p1 = make()
    s2, p2 = make()
    T = s1.num_tensors
    s1.stats_range(0, T, 0.5); s1.apply_range(0, T, 0.5); s1.stats_range(0, T, 0.5)
    s2.stats_range(0, T, 0.5); s2.apply_range(0, T, 0.5, refresh_norms=True)
    torch.cuda.synchronize()
    for x, y in zip(p1, p2):
        assert torch.equal(x.w, y.w) and torch.equal(x.m, y.m)
    n1 = s1.stats.view(-1, 2)[:, 0].cpu(); n2 = s2.stats.view(-1, 2)[:, 0].cpu()
    torch.testing.assert_close(n2, n1, rtol=1e-6, atol=0)
    want = torch.stack([(p.w.double() ** 2).sum().float().cpu() for p in p2])
    torch.testing.assert_close(n2[:len(p2)], want, rtol=1e-5, atol=0)
    torch.testing.assert_close(s1.reg_loss_from_stats().cpu(), s2.reg_loss_from_stats().cpu(), rtol=1e-6, atol=0)


@pytest.mark.parametrize("stride", [1, 2])
def test_depthwise_conv3x3_fwd_dgrad_wgrad(stride):
    """slim.separable_conv2d depthwise stage (mobilenet_v1.py:230-238) against torch grouped conv."""
    from mtl_ssl_b200 import ops
    g = torch.Generator().manual_seed(stride)
    N, H, W, C = 2, 9, 11, 32
    x = bf(torch.randn(N, H, W, C, generator=g) * 2)
    w = bf(torch.randn(C, 3, 3, generator=g) * 0.3)
    bias = torch.randn(C, generator=g) * 0.1
    P, pt, pb = ON.same_pad(H, 3, stride)
    Q, pl, pr = ON.same_pad(W, 3, stride)
    y = torch.empty(N, P, Q, C, dtype=torch.bfloat16, device="cuda")
    ops.call("mtl_dwconv3x3_fwd", x.cuda(), w.cuda(), bias.cuda(), N, H, W, C, stride, pt, pl, P, Q, 2, y)
    xr = x.float().requires_grad_(True)
    wr = w.float().requires_grad_(True)
    xp = torch.nn.functional.pad(xr.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    pre = torch.nn.functional.conv2d(xp, wr[:, None], stride=stride, groups=C).permute(0, 2, 3, 1) + bias
    want = torch.clamp(torch.relu(pre), max=6.0)
    torch.testing.assert_close(y.float().cpu(), want.detach(), rtol=8e-3, atol=8e-3)
    dy = bf(torch.randn(N, P, Q, C, generator=g))
    pre.backward(dy.float())                      # gradient w.r.t. the pre-activation (callers mask dy themselves)
    mask = bf(torch.rand(N, H, W, C, generator=g) * 8 - 1)       # ReLU6-style mask: alive iff 0 < m < 6
    dx = torch.empty(N, H, W, C, dtype=torch.bfloat16, device="cuda")
    ops.call("mtl_dwconv3x3_dgrad", dy.cuda(), w.cuda(), N, H, W, C, stride, pt, pl, P, Q, mask.cuda(), 6.0, dx)
    alive = (mask.float() > 0) & (mask.float() < 6)
    torch.testing.assert_close(dx.float().cpu(), xr.grad * alive, rtol=8e-3, atol=8e-3)
    scale = torch.rand(C, generator=g) + 0.5
    dw = torch.zeros(C, 3, 3, device="cuda")
    ops.call("mtl_dwconv3x3_wgrad", dy.cuda(), x.cuda(), N, H, W, C, stride, pt, pl, P, Q, scale.cuda(), dw)
    torch.testing.assert_close(dw.cpu(), wr.grad * scale[:, None, None], rtol=1e-3, atol=1e-3)


def test_psroi_fwd_bwd_matches_oracle():
    """utils/ops.py:462-609 (global_pool=True) through mtl_psroi_fwd / mtl_psroi_bwd, with channel offset."""
    from mtl_ssl_b200 import ops
    from oracle.model import Oracle
    g = torch.Generator().manual_seed(4)
    B, H, W, D, nb = 2, 11, 13, 5, 9
    Ct, c0 = 64, 8                                   # PS channels live at [c0, c0 + 9*D) of a wider map
    fmap = bf(torch.randn(B, H, W, Ct, generator=g))
    R = 12
    boxes = rand_norm_boxes(g, R)
    bi = torch.randint(0, B, (R,), generator=g).int()
    out = torch.zeros(R, 16, device="cuda")
    ops.call("mtl_psroi_fwd", fmap.cuda(), B, H, W, Ct, c0, D, 3, 3, 18, 18, boxes.cuda(), bi.cuda(), R, out, 16, 3)
    fr = fmap.float().requires_grad_(True)
    o = Oracle({}, {"architecture": "resnet_v1_50"}, bf16=False)
    want = o.psroi(fr[..., c0:c0 + nb * D], boxes.numpy(), bi.numpy().astype(np.int64), D, (3, 3), (18, 18))
    torch.testing.assert_close(out[:, 3:3 + D].cpu(), want.detach(), rtol=1e-4, atol=1e-5)
    assert not out[:, :3].any() and not out[:, 3 + D:].any()
    dout = torch.randn(R, 16, generator=g)
    want.backward(dout[:, 3:3 + D])
    dmap = torch.zeros(B, H, W, Ct, device="cuda")
    ops.call("mtl_psroi_bwd", dout.cuda(), 16, 3, B, H, W, Ct, c0, D, 3, 3, 18, 18, boxes.cuda(), bi.cuda(), R, dmap)
    torch.testing.assert_close(dmap.cpu(), fr.grad, rtol=1e-4, atol=1e-5)



def test_gradient_multipliers_and_frozen_variables():
    """trainer.py:387-410 of the reference: scalar / bias gradient multipliers act on the gradient of the
    total loss (task + L2 term) BEFORE the per-tensor clip, frozen variables (regular expressions) take no update at all;
    here through ParamStore.set_gradient_policy and Trainer._apply_gradient_knobs."""
    import types
    from mtl_ssl_b200.runtime import ParamStore
    from mtl_ssl_b200.trainer import Trainer
    store = ParamStore()
    a = store.add("Net/conv/weights", (8, 3, 3, 16), l2=1e-2, init=("normal", 1.0))
    b = store.add("Net/conv/biases", (8,), init=("normal", 1.0))
    c = store.add("Net/frozen_me/weights", (4, 40), l2=1e-3, init=("normal", 1.0))
    store.finalize("cuda", seed=3)
    tc = types.SimpleNamespace(grad_multiplier=3.0, divide_grad_by_batch=True, batch_size=2, bias_grad_multiplier=2.0,
                               freeze_variables=["", "Net/frozen.*"])
    holder = types.SimpleNamespace(model=types.SimpleNamespace(param_store=store))
    Trainer._apply_gradient_knobs(holder, tc)
    assert a.grad_mult == 1.5 and b.grad_mult == 3.0 and not c.trainable and a.trainable
    ps = [a, b, c]
    w0 = [p.w.cpu().clone() for p in ps]
    gen = torch.Generator().manual_seed(4)
    grads = [torch.randn(p.shape, generator=gen) * s for p, s in zip(ps, [20.0, 0.3, 1.0])]
    lr, mu, clip = 0.1, 0.9, 10.0
    store.set_hyper(lr, mu, clip)
    for p, gr in zip(ps, grads):
        p.g.copy_(gr.cuda())
    reg = store.stats_and_reg_loss(1.0)
    store.apply(1.0)
    np.testing.assert_allclose(reg.item(), sum(p.l2 * 0.5 * (w ** 2).sum() for p, w in zip(ps, w0)).item(), rtol=1e-5)
    for p, w, gr in zip(ps, w0, grads):
        if not p.trainable:
            torch.testing.assert_close(p.w.cpu(), w, rtol=0, atol=0)            # frozen: untouched
            continue
        g = (gr + p.l2 * w) * p.grad_mult
        g = g * clip / max(g.norm().item(), clip)
        torch.testing.assert_close(p.w.cpu(), w - lr * g, rtol=1e-5, atol=1e-6)
    assert float((grads[0] + a.l2 * w0[0]).norm()) * 1.5 > clip          # the clip acted on the multiplied gradient


@pytest.mark.parametrize("R,HW,C,n,mask_hi", [(37, 49, 2048, 104, 0.0), (64, 16, 1024, 24, 6.0), (5, 9, 1536, 456, 0.0)])
def test_fused_head_forward_and_backward(R, HW, C, n, mask_hi):
    """csrc/head.cu: spatial average + FC layers (bp:470-500, :568-602) in one kernel, and their backward (logit gradient
    -> bf16 GEMM operand, bias gradient, pooled-feature gradient broadcast over the ROI grid under the ReLU / ReLU6 mask)
    against the same computation in torch fp32 with the device's rounding points (pooled features and their gradient are
    bf16 tensors).  R odd: the last block holds one ROI; n = 456: the COCO box + class head."""
    from mtl_ssl_b200 import ops
    g = torch.Generator().manual_seed(R + n)
    x = bf(torch.relu(torch.randn(R, HW, C, generator=g)) * (7.0 if mask_hi else 1.0)).cuda()
    w = bf(torch.randn(n, C, generator=g) * 0.02).cuda()
    bias = torch.randn(n, generator=g).cuda()
    pooled = torch.empty(R, C, dtype=torch.bfloat16, device="cuda")
    ld = n + 8
    out = torch.zeros(R, ld, device="cuda")
    ops.call("mtl_head_fwd", x, R, HW, C, w, bias, n, pooled, out, ld)
    want_p = (x.float().sum(1) * (1.0 / HW)).to(torch.bfloat16)
    # (sum order: positions ascending in both; the final scaling may round differently by one bf16 ulp)
    torch.testing.assert_close(pooled.float(), want_p.float(), rtol=1e-2, atol=1e-6)
    want_o = pooled.float() @ w.float().t() + bias
    torch.testing.assert_close(out[:, :n], want_o, rtol=1e-4, atol=1e-4)
    assert not out[:, n:].any()
    d_out = torch.randn(R, ld, generator=g).cuda() * 0.1
    dyb = torch.empty(R, n, dtype=torch.bfloat16, device="cuda")
    db = torch.full((n,), 0.5, device="cuda")
    dx = torch.empty(R, HW, C, dtype=torch.bfloat16, device="cuda")
    ops.call("mtl_head_bwd", d_out, ld, n, w, x, mask_hi, R, HW, C, dyb, db, dx)
    want_dy = d_out[:, :n].to(torch.bfloat16)
    assert torch.equal(dyb, want_dy)
    torch.testing.assert_close(db, 0.5 + want_dy.float().sum(0), rtol=1e-4, atol=1e-4)
    dp = (want_dy.float() @ w.float()).to(torch.bfloat16).float() * (1.0 / HW)
    alive = x.float() > 0
    if mask_hi:
        alive &= x.float() < mask_hi
    want_dx = torch.where(alive, dp[:, None, :].expand(R, HW, C), torch.zeros(())).to(torch.bfloat16)
    torch.testing.assert_close(dx.float(), want_dx.float().cuda() if not want_dx.is_cuda else want_dx.float(),
                               rtol=2e-2, atol=1e-6)
    assert torch.equal(dx == 0, want_dx.cuda() == 0) or float(((dx == 0) != (want_dx.cuda() == 0)).float().mean()) < 1e-4
    # without a gradient for the features / the biases
    ops.call("mtl_head_bwd", d_out, ld, n, w, None, 0.0, R, HW, C, dyb, None, None)
    assert torch.equal(dyb, want_dy)
