"""GPU parity: box-geometry kernels (csrc/boxes.cu) against the NumPy oracle, through the C ABI.
Integer / index outputs must be bit-exact; float outputs that use only + - * / sqrt are
bit-exact too (boxes.cu is built with -fmad=false); exp/log outputs carry a stated tolerance."""
import numpy as np
import pytest
import torch

from oracle import assign as OA
from oracle import boxes as OB
from oracle import postprocess as OP

pytestmark = pytest.mark.gpu
F = np.float32


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def rand_boxes(rng, n, h, w, min_size=4.0):
    cy = rng.uniform(0, h, n); cx = rng.uniform(0, w, n)
    bh = np.exp(rng.uniform(np.log(min_size), np.log(h * 0.8), n))
    bw = np.exp(rng.uniform(np.log(min_size), np.log(w * 0.8), n))
    b = np.stack([cy - bh / 2, cx - bw / 2, cy + bh / 2, cx + bw / 2], 1)
    b[:, [0, 2]] = np.clip(b[:, [0, 2]], 0, h)
    b[:, [1, 3]] = np.clip(b[:, [1, 3]], 0, w)
    return b.astype(F)


@pytest.mark.parametrize("hw", [(2, 2), (38, 63), (50, 84)])
def test_grid_anchors_and_prune_bit_exact(hw):
    from mtl_ssl_b200 import ops
    Hf, Wf = hw
    scales, ars = [0.25, 0.5, 1.0, 2.0], [0.5, 1.0, 2.0]
    A = len(scales) * len(ars)
    out = torch.empty(Hf * Wf * A, 4, device="cuda")
    ops.call("mtl_grid_anchors", Hf, Wf, scales, len(scales), ars, len(ars), 256.0, 256.0, 16.0, 16.0, 0.0, 0.0, out)
    want = OB.grid_anchors(Hf, Wf, scales, ars, (256, 256), (16, 16), (0, 0))
    assert np.array_equal(out.cpu().numpy(), want)
    H, W = Hf * 16.0 - 8, Wf * 16.0 - 8
    keep = torch.empty(out.shape[0], dtype=torch.int32, device="cuda")
    kept = torch.empty_like(out)
    num = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.call("mtl_prune_outside_window", out, out.shape[0], 0.0, 0.0, H, W, keep, kept, num)
    wb, widx = OB.prune_outside_window(want, (0, 0, H, W))
    n = int(num.item())
    assert n == len(widx)
    assert np.array_equal(keep[:n].cpu().numpy(), widx.astype(np.int32))
    assert np.array_equal(kept[:n].cpu().numpy(), wb)


def _decode_inputs(rng, B, Hf, Wf, A=12):
    anchors = OB.grid_anchors(Hf, Wf, [0.25, 0.5, 1.0, 2.0], [0.5, 1.0, 2.0])
    H, W = Hf * 16.0, Wf * 16.0
    kept, kidx = OB.prune_outside_window(anchors, (0, 0, H, W))
    ld = A * 6
    rpn_out = (rng.standard_normal((B, Hf * Wf, ld)) * 0.5).astype(F)
    return anchors, kept, kidx.astype(np.int32), rpn_out, H, W, ld


def test_rpn_decode_matches_oracle():
    from mtl_ssl_b200 import ops
    rng = np.random.default_rng(0)
    B, Hf, Wf, A = 2, 20, 30, 12
    anchors, kept, kidx, rpn_out, H, W, ld = _decode_inputs(rng, B, Hf, Wf)
    Nk = len(kidx)
    boxes = torch.empty(B, Nk, 4, device="cuda"); scores = torch.empty(B, Nk, device="cuda")
    keys = torch.empty(B, Nk, dtype=torch.int64, device="cuda")
    ops.call("mtl_rpn_decode", dev(rpn_out), ld, 0, A * 4, A, Hf * Wf, dev(kidx), dev(kept), Nk, B, H, W, 0.0,
             boxes, scores, keys)
    for b in range(B):
        enc = rpn_out[b].reshape(Hf * Wf, ld)[:, :A * 4].reshape(-1, 4)[kidx]
        logit = rpn_out[b].reshape(Hf * Wf, ld)[:, A * 4:].reshape(-1, 2)[kidx]
        dec = OB.box_decode(enc, kept)
        clipped, _ = OB.clip_to_window(dec, (0, 0, H, W), filter_nonoverlapping=False)
        sc = OP.softmax(logit)[:, 1]
        # expf differs from numpy's exp by a few ulp -> relative 1e-6 on sizes up to ~1e3 px
        np.testing.assert_allclose(boxes[b].cpu().numpy(), clipped, rtol=2e-6, atol=2e-4)
        np.testing.assert_allclose(scores[b].cpu().numpy(), sc, rtol=2e-6, atol=1e-7)
        # key validity flag = (score > 0) & (area > 0) on the GPU's own values
        gb = boxes[b].cpu().numpy(); gs = scores[b].cpu().numpy()
        valid = (gs > 0) & (OB.area(gb) > 0)
        assert np.array_equal(keys[b].cpu().numpy() != 0, valid)


@pytest.mark.parametrize("n,max_out", [(1, 5), (300, 50), (5000, 300), (14000, 300)])
def test_sort_and_nms_bit_exact(n, max_out):
    from mtl_ssl_b200 import ops
    rng = np.random.default_rng(n)
    B = 2
    H, W = 600.0, 1000.0
    boxes = np.stack([rand_boxes(rng, n, H, W) for _ in range(B)])
    # clustered boxes so that suppression really happens
    boxes[:, n // 2:] = boxes[:, : n - n // 2] + rng.uniform(-6, 6, (B, n - n // 2, 4)).astype(F)
    scores = rng.uniform(0.0, 1.0, (B, n)).astype(F)
    scores[:, ::7] = scores[:, 1::7][:, : scores[:, ::7].shape[1]] if n > 7 else scores[:, ::7]   # ties
    scores[:, ::11] = 0.0                                                                        # filtered
    db, ds = dev(boxes), dev(scores)
    keys = torch.empty(B, n, dtype=torch.int64, device="cuda")
    ops.call("mtl_nms_make_keys", db, ds, B, n, 0.0, 1, keys)
    order = torch.full((B, n), -1, dtype=torch.int32, device="cuda")
    nvalid = torch.zeros(B, dtype=torch.int32, device="cuda")
    ops.call("mtl_rank_sort_desc", keys, B, n, order, nvalid, torch.empty(B, n, dtype=torch.int32, device="cuda"))
    ob = torch.empty(B, max_out, 4, device="cuda"); osc = torch.empty(B, max_out, device="cuda")
    oi = torch.empty(B, max_out, dtype=torch.int32, device="cuda"); no = torch.zeros(B, dtype=torch.int32, device="cuda")
    ops.call("mtl_nms", db, ds, order, nvalid, B, n, 0.7, max_out, ob, osc, oi, no)
    for b in range(B):
        valid = np.nonzero((scores[b] > 0) & (OB.area(boxes[b]) > 0))[0]
        want_order = valid[np.argsort(-scores[b][valid], kind="stable")]
        nv = int(nvalid[b].item())
        assert nv == len(valid)
        assert np.array_equal(order[b, :nv].cpu().numpy(), want_order.astype(np.int32))
        sel = OP.nms_vectorized(boxes[b][valid], scores[b][valid], max_out, 0.7)
        k = int(no[b].item())
        assert k == len(sel)
        assert np.array_equal(oi[b, :k].cpu().numpy(), valid[sel].astype(np.int32))
        assert np.array_equal(ob[b, :k].cpu().numpy(), boxes[b][valid[sel]])
        assert np.array_equal(osc[b, :k].cpu().numpy(), scores[b][valid[sel]])
        assert not ob[b, k:].any() and not osc[b, k:].any()


def test_nms_small_matches_scalar_oracle():
    """The vectorised oracle itself is pinned to the literal TF restatement on a small case."""
    rng = np.random.default_rng(5)
    boxes = rand_boxes(rng, 200, 100, 100)
    boxes[100:] = boxes[:100] + rng.uniform(-3, 3, (100, 4)).astype(F)
    scores = rng.uniform(0.01, 1, 200).astype(F)
    a = OP.tf_non_max_suppression(boxes, scores, 50, 0.5)
    b = OP.nms_vectorized(boxes, scores, 50, 0.5)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("force,thr", [(True, (0.7, 0.3)), (False, (0.5, 0.5))])
@pytest.mark.parametrize("G", [0, 1, 7])
def test_iou_match_bit_exact(force, thr, G):
    from mtl_ssl_b200 import ops
    rng = np.random.default_rng(G + 10 * force)
    B, Gmax = 3, 8
    anchors = OB.prune_outside_window(OB.grid_anchors(20, 30, [0.25, 0.5, 1.0, 2.0], [0.5, 1.0, 2.0]),
                                      (0, 0, 320, 480))[0]
    N = len(anchors)
    gt = np.zeros((B, Gmax, 4), F); ng = np.zeros(B, np.int32)
    for b in range(B):
        g = min(G + b, Gmax) if G else 0
        ng[b] = g
        gt[b, :g] = rand_boxes(rng, g, 320, 480, 16)
        if g > 1:
            gt[b, 1] = gt[b, 0]          # duplicate GT row: tie handling of force-match
    match = torch.empty(B, N, dtype=torch.int32, device="cuda")
    miou = torch.empty(B, N, device="cuda")
    rb = torch.zeros(B, Gmax, dtype=torch.int64, device="cuda")
    ops.call("mtl_iou_match", dev(gt), dev(ng), Gmax, dev(anchors), 0, None, B, N, thr[0], thr[1], 1, int(force),
             match, miou, rb)
    for b in range(B):
        sim = OB.iou(gt[b, :ng[b]], anchors)
        want = OA.argmax_match(sim, thr[0], thr[1], True, force)
        assert np.array_equal(match[b].cpu().numpy(), want), "image %d" % b
        if ng[b]:
            assert np.array_equal(miou[b].cpu().numpy(), sim.max(0))


def test_balanced_sampler_and_gather_bit_exact():
    from mtl_ssl_b200 import ops
    rng = np.random.default_rng(3)
    B, N, P = 3, 300, 64
    match = rng.integers(-2, 4, (B, N)).astype(np.int32)
    match[2, :] = np.where(match[2] >= 0, -1, match[2])       # image without positives
    match[1, 200:] = -3                                        # absent (beyond num_proposals)
    keys = rng.random((B, N)).astype(F)
    keys[:, 5] = keys[:, 4]                                    # tie
    sampled = torch.empty(B, N, dtype=torch.uint8, device="cuda")
    counts = torch.zeros(B, 4, dtype=torch.int32, device="cuda")
    ops.call("mtl_balanced_sample", dev(match), dev(keys), B, N, P, 0.25, sampled, counts)
    boxes = np.stack([rand_boxes(rng, N, 600, 1000) for _ in range(B)])
    scores = rng.random((B, N)).astype(F)
    oa = torch.empty(B, P, 4, device="cuda"); on = torch.empty(B, P, 4, device="cuda")
    osc = torch.empty(B, P, device="cuda"); no = torch.zeros(B, dtype=torch.int32, device="cuda")
    ops.call("mtl_gather_sampled", dev(boxes), dev(scores), sampled, B, N, P, 600.0, 1000.0, oa, on, osc, no)
    for b in range(B):
        ind = (match[b] >= -1)
        lab = match[b] >= 0
        want = OA.balanced_subsample(ind, P, lab, 0.25, keys[b])
        assert np.array_equal(sampled[b].cpu().numpy().astype(bool), want)
        assert int(counts[b, 3].item()) == int(want.sum())
        idx = np.nonzero(want)[0][:P]
        nb = OB.to_normalized_coordinates(boxes[b][idx], 600, 1000)
        ab = OB.to_absolute_coordinates(nb, 600, 1000)
        k = int(no[b].item())
        assert k == len(idx)
        assert np.array_equal(on[b, :k].cpu().numpy(), nb)
        assert np.array_equal(oa[b, :k].cpu().numpy(), ab)
        assert np.array_equal(osc[b, :k].cpu().numpy(), scores[b][idx])
        assert not on[b, k:].any()


def test_detection_and_rpn_targets():
    from mtl_ssl_b200 import ops
    rng = np.random.default_rng(4)
    B, Gmax, P, K1 = 2, 6, 64, 21
    gt = np.stack([rand_boxes(rng, Gmax, 600, 1000, 32) for _ in range(B)])
    ng = np.array([4, 6], np.int32)
    props = np.stack([rand_boxes(rng, P, 600, 1000, 16) for _ in range(B)])
    props[:, :6] = gt + rng.uniform(-4, 4, gt.shape).astype(F)
    props = np.clip(props, 0, [600, 1000, 600, 1000]).astype(F)
    gcls = rng.integers(1, K1, (B, Gmax)).astype(np.int32)
    gclose = rng.random((B, Gmax, K1)).astype(F)
    match = torch.empty(B, P, dtype=torch.int32, device="cuda")
    ops.call("mtl_iou_match", dev(gt), dev(ng), Gmax, dev(props), P, None, B, P, 0.5, 0.5, 1, 0, match, None, None)
    ct = torch.empty(B, P, dtype=torch.int32, device="cuda"); rt = torch.empty(B, P, 4, device="cuda")
    rw = torch.empty(B, P, device="cuda"); cw = torch.empty(B, P, device="cuda")
    clt = torch.empty(B, P, K1, device="cuda"); clw = torch.empty(B, P, device="cuda")
    ops.call("mtl_detection_targets", match, dev(props), dev(gt), dev(gcls), dev(gclose), B, Gmax, P, K1, ct, rt, rw,
             cw, clt, clw)
    for b in range(B):
        onehot = np.zeros((ng[b], K1), F); onehot[np.arange(ng[b]), gcls[b, :ng[b]]] = 1
        want = OA.assign_detection(props[b], gt[b, :ng[b]], onehot, gclose[b, :ng[b]])
        assert np.array_equal(match[b].cpu().numpy(), want["match"])
        assert np.array_equal(ct[b].cpu().numpy(), want["cls_targets"].argmax(1))
        assert np.array_equal(rw[b].cpu().numpy(), want["reg_weights"])
        assert np.array_equal(cw[b].cpu().numpy(), want["cls_weights"])
        # logf vs numpy log: 1e-6 relative on O(1) values
        np.testing.assert_allclose(rt[b].cpu().numpy(), want["reg_targets"], rtol=2e-6, atol=2e-6)
        assert np.array_equal(clt[b].cpu().numpy(), want["closeness_targets"])
        cwant = want["reg_weights"] / max(1.0, want["reg_weights"].sum()) * want["closeness_targets"][:, 1:].sum(1)
        np.testing.assert_allclose(clw[b].cpu().numpy(), cwant, rtol=1e-6, atol=1e-8)


def test_expand_windows_bit_exact():
    from mtl_ssl_b200 import ops
    rng = np.random.default_rng(8)
    B, P = 2, 16
    y = np.sort(rng.random((B, P, 2)).astype(F), -1)
    x = np.sort(rng.random((B, P, 2)).astype(F), -1)
    pr = np.ascontiguousarray(np.stack([y[..., 0], x[..., 0], y[..., 1], x[..., 1]], -1))
    out = torch.empty(5, B, P, 4, device="cuda"); bi = torch.empty(5, B, P, dtype=torch.int32, device="cuda")
    ops.call("mtl_expand_windows", dev(pr), B, P, 4, out, bi)
    ymin, xmin, ymax, xmax = [pr[..., i] for i in range(4)]
    for e in range(5):
        want = np.stack([ymin - (ymin / F(4)) * F(e), xmin - (xmin / F(4)) * F(e),
                         ymax + ((F(1) - ymax) / F(4)) * F(e), xmax + ((F(1) - xmax) / F(4)) * F(e)], -1).astype(F)
        assert np.array_equal(out[e].cpu().numpy(), want)
    assert np.array_equal(bi.cpu().numpy(), np.broadcast_to(np.arange(B)[None, :, None], (5, B, P)))
