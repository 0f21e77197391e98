"""GPU parity of the second-stage detection path (first slice: `postprocess` of
meta_architectures/faster_rcnn_meta_arch.py:996-1053 / :1387-1469 + core/post_processing.py:25-312) through the C ABI:
decode / score conversion against the oracle to fp32 rounding, then -- on the device's own decoded boxes and
scores -- per-class NMS, merge and top-k bit-exact against oracle/postprocess.py."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(seed, B, P, K, H, W):
    rng = np.random.default_rng(seed)
    props = np.zeros((B, P, 4), np.float32)
    y0, x0 = rng.uniform(0, H * 0.8, (B, P)), rng.uniform(0, W * 0.8, (B, P))
    props[..., 0], props[..., 1] = y0, x0
    props[..., 2], props[..., 3] = y0 + rng.uniform(4, H * 0.5, (B, P)), x0 + rng.uniform(4, W * 0.5, (B, P))
    enc = rng.normal(0, 1.5, (B * P, K, 4)).astype(np.float32)
    logits = rng.normal(0, 2.0, (B * P, K + 1)).astype(np.float32)
    return props, enc, logits


@pytest.mark.parametrize("B,P,K,nprop,mode", [(2, 64, 5, (64, 40), 1), (1, 256, 20, (199,), 1), (3, 33, 3, (33, 1, 0), 2)])
def test_detection_chain_matches_oracle(B, P, K, nprop, mode):
    from mtl_ssl_b200 import ops
    from oracle import postprocess as PP
    from oracle import boxes as OB
    H, W, thr, iou, M, T = 200.0, 300.0, 0.05, 0.4, 10, 30
    props, enc, logits = _inputs(B * 100 + P, B, P, K, H, W)
    dev = "cuda"
    t = lambda a, dt=None: torch.from_numpy(np.ascontiguousarray(a)).to(dev) if dt is None else torch.from_numpy(a).to(dev, dt)
    d_props, d_enc, d_logits = t(props), t(enc), t(logits)
    d_np = torch.tensor(nprop, dtype=torch.int32, device=dev)
    boxes_n = torch.empty(B, K, P, 4, device=dev)
    scores = torch.empty(B, K, P, device=dev)
    keys = torch.empty(B, K, P, dtype=torch.int64, device=dev)
    dec = torch.empty(B, K, P, 4, device=dev)
    ops.call("mtl_detection_decode", d_enc, d_logits, d_props, d_np, B, P, K, H, W, thr, mode, boxes_n, scores, keys, dec)
    # --- stage 1: decode + score conversion vs the oracle (exp / softmax to fp32 rounding)
    dec_h = dec.cpu().numpy().transpose(0, 2, 1, 3)            # [B,P,K,4]
    sc_h = scores.cpu().numpy().transpose(0, 2, 1)             # [B,P,K]
    for b in range(B):
        for k in range(K):
            want = OB.box_decode(enc.reshape(B, P, K, 4)[b, :, k], props[b])
            np.testing.assert_allclose(dec_h[b, :, k], want, rtol=2e-6, atol=1e-3)
    lg = logits.reshape(B, P, K + 1)
    want_s = PP.softmax(lg)[..., 1:] if mode == 1 else 1.0 / (1.0 + np.exp(-lg[..., 1:]))
    np.testing.assert_allclose(sc_h, want_s, rtol=2e-6, atol=1e-7)
    # --- stage 2: NMS / merge / top-k, index exact on the device's own boxes and scores
    order = torch.empty(B * K, P, dtype=torch.int32, device=dev)
    nvalid = torch.empty(B * K, dtype=torch.int32, device=dev)
    ops.call("mtl_rank_sort_desc", keys, B * K, P, order, nvalid, torch.empty(B * K, P, dtype=torch.int32, device=dev))
    cls_b = torch.empty(B, K, M, 4, device=dev)
    cls_s = torch.empty(B, K, M, device=dev)
    cls_n = torch.empty(B * K, dtype=torch.int32, device=dev)
    ops.call("mtl_nms", boxes_n, scores, order, nvalid, B * K, P, iou, M, cls_b, cls_s, None, cls_n)
    keys2 = torch.empty(B, K * M, dtype=torch.int64, device=dev)
    ops.call("mtl_detection_merge_keys", cls_s, cls_n, B, K, M, keys2)
    order2 = torch.empty(B, K * M, dtype=torch.int32, device=dev)
    nvalid2 = torch.empty(B, dtype=torch.int32, device=dev)
    ops.call("mtl_rank_sort_desc", keys2, B, K * M, order2, nvalid2, torch.empty(B, K * M, dtype=torch.int32, device=dev))
    det_b = torch.empty(B, T, 4, device=dev)
    det_s = torch.empty(B, T, device=dev)
    det_c = torch.empty(B, T, device=dev)
    det_n = torch.empty(B, device=dev)
    ops.call("mtl_detection_gather", cls_b, cls_s, order2, nvalid2, B, K, M, T, det_b, det_s, det_c, det_n)
    wb, ws_, wc, wn = PP.second_stage_postprocess(enc, logits, props, list(nprop), (H, W), thr, iou, M, T,
                                                  decoded=dec_h, scores=sc_h)
    assert np.array_equal(det_n.cpu().numpy(), wn), (det_n.cpu().numpy(), wn)
    assert np.array_equal(det_c.cpu().numpy(), wc)
    assert np.array_equal(det_s.cpu().numpy(), ws_)
    assert np.array_equal(det_b.cpu().numpy(), wb)


def test_meta_arch_postprocess_after_forward():
    """model.postprocess(prediction_dict) on a real forward pass equals the oracle bit for bit when the oracle is fed the
    device's decoded boxes / converted scores (a random-init head gives near-equal scores, so recomputing exp() on the
    host could legitimately reorder neighbours)."""
    import test_gpu_train_step as T
    from mtl_ssl_b200 import ops
    from oracle import postprocess as PP
    H, W = 224, 320
    cfg, model, sd, examples, keys, tr = T._setup("model12.config", T.SMALL, H, W, 1)
    arrays = tr.host_arrays(examples, keys)
    image = tr._bind(arrays)
    pd = model.predict(model.preprocess(image))
    det = model.postprocess(pd)
    torch.cuda.synchronize()
    nms = cfg.model.faster_rcnn.second_stage_post_processing.batch_non_max_suppression
    enc = pd["refined_box_encodings"].contiguous().float()
    logits = pd["class_predictions_with_background"].contiguous().float()
    props, nprop = pd["proposal_boxes"], pd["num_proposals"]
    B, P, K = props.shape[0], props.shape[1], model.num_classes
    M, Tt = min(nms.max_detections_per_class, P), nms.max_total_detections
    dev = enc.device
    boxes_n = torch.empty(B, K, P, 4, device=dev)
    scores = torch.empty(B, K, P, device=dev)
    k64 = torch.empty(B, K, P, dtype=torch.int64, device=dev)
    dec = torch.empty(B, K, P, 4, device=dev)
    ops.call("mtl_detection_decode", enc, logits, props, nprop, B, P, K, float(H), float(W), float(nms.score_threshold),
             1, boxes_n, scores, k64, dec)
    wb, ws_, wc, wn = PP.second_stage_postprocess(
        enc.cpu().numpy(), logits.cpu().numpy(), props.cpu().numpy(), nprop.cpu().numpy(), (H, W), nms.score_threshold,
        nms.iou_threshold, M, Tt, decoded=dec.cpu().numpy().transpose(0, 2, 1, 3),
        scores=scores.cpu().numpy().transpose(0, 2, 1))
    n = int(det["num_detections"][0].item())
    assert det["detection_boxes"].shape == (1, Tt, 4)
    assert 0 < n <= Tt
    assert np.array_equal(det["num_detections"].cpu().numpy(), wn)
    assert np.array_equal(det["detection_scores"].cpu().numpy(), ws_)
    assert np.array_equal(det["detection_classes"].cpu().numpy(), wc)
    assert np.array_equal(det["detection_boxes"].cpu().numpy(), wb)
    s = det["detection_scores"][0].cpu().numpy()
    b = det["detection_boxes"][0].cpu().numpy()
    assert np.all(np.diff(s[:n]) <= 0) and not s[n:].any()
    assert b[:n].min() >= 0 and b[:n].max() <= 1 and not b[n:].any()


def test_inference_mode_predict_and_postprocess():
    """is_training=False graph (fmA:586-590 clipped anchors, fmA:1111-1131 unsampled proposals, max_num_proposals =
    first_stage_max_proposals) against oracle/model.py `forward(inference=True)`: proposal selection index-exact on the
    device's own RPN outputs, head outputs to bf16 rounding, then `postprocess` bit-exact on the device's decode."""
    import test_gpu_train_step as T
    from helpers import load_config, oracle_config, randomize_bn
    from mtl_ssl_b200 import ops
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    from oracle.model import Oracle
    from oracle import postprocess as PP
    H, W, B = 224, 320, 1
    cfg = load_config("model12.config", T.SMALL)
    model = model_builder.build(cfg.model, False, device="cuda", seed=0)
    sd = randomize_bn(model.param_store.state_dict(), 0)
    model.param_store.load_state_dict(sd)
    K = cfg.model.faster_rcnn.num_classes
    ex = synthetic.make_batch(7, B, H, W, K, max_boxes=4, num_windows=16)
    images = np.stack([e["image"] for e in ex]).astype(np.float32)
    image = torch.from_numpy(images).cuda()
    assert model.max_num_proposals == cfg.model.faster_rcnn.first_stage_max_proposals
    pd = model.predict(model.preprocess(image))
    det = model.postprocess(pd)
    torch.cuda.synchronize()
    P = model.max_num_proposals
    # every anchor is kept (clipped), none pruned
    Hf, Wf = pd["_feat_hw"]
    assert pd["anchors"].shape[0] == Hf * Wf * 12
    assert float(pd["anchors"].min()) >= 0 and float(pd["anchors"][:, 2].max()) <= H and float(pd["anchors"][:, 3].max()) <= W
    orc = Oracle({k: v for k, v in sd.items() if "/_pad/" not in k}, oracle_config(cfg), bf16=True)
    prop_in = (pd["rpn_box_encodings"].cpu().numpy(), pd["rpn_objectness_predictions_with_background"].cpu().numpy())
    with torch.no_grad():
        out = orc.forward(torch.from_numpy(images), None, None, H, W, proposal_inputs=prop_in, inference=True)
    assert np.array_equal(pd["num_proposals"].cpu().numpy(), out["nprop"])
    n = int(out["nprop"][0])
    assert 0 < n <= P
    np.testing.assert_allclose(pd["proposal_boxes"].cpu().numpy(), out["prop_abs"], rtol=1e-5, atol=1e-3)

    def cos(a, b):
        a, b = a.reshape(-1).double(), b.reshape(-1).double()
        return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))

    enc = pd["refined_box_encodings"].float().cpu()[:n]
    cls = pd["class_predictions_with_background"].float().cpu()[:n]
    assert cos(enc, out["refined_box_encodings"].float()[:n]) > 0.999
    assert cos(cls, out["class_predictions_with_background"].float()[:n]) > 0.999
    # detections: exact against the oracle on the device's decoded boxes / scores
    nms = cfg.model.faster_rcnn.second_stage_post_processing.batch_non_max_suppression
    e32 = pd["refined_box_encodings"].contiguous().float()
    l32 = pd["class_predictions_with_background"].contiguous().float()
    M, Tt = min(nms.max_detections_per_class, P), nms.max_total_detections
    boxes_n = torch.empty(B, K, P, 4, device="cuda")
    scores = torch.empty(B, K, P, device="cuda")
    k64 = torch.empty(B, K, P, dtype=torch.int64, device="cuda")
    dec = torch.empty(B, K, P, 4, device="cuda")
    ops.call("mtl_detection_decode", e32, l32, pd["proposal_boxes"], pd["num_proposals"], B, P, K, float(H), float(W),
             float(nms.score_threshold), 1, boxes_n, scores, k64, dec)
    wb, ws_, wc, wn = PP.second_stage_postprocess(
        e32.cpu().numpy(), l32.cpu().numpy(), pd["proposal_boxes"].cpu().numpy(), pd["num_proposals"].cpu().numpy(),
        (H, W), nms.score_threshold, nms.iou_threshold, M, Tt, decoded=dec.cpu().numpy().transpose(0, 2, 1, 3),
        scores=scores.cpu().numpy().transpose(0, 2, 1))
    assert np.array_equal(det["num_detections"].cpu().numpy(), wn) and wn[0] > 0
    assert np.array_equal(det["detection_scores"].cpu().numpy(), ws_)
    assert np.array_equal(det["detection_classes"].cpu().numpy(), wc)
    assert np.array_equal(det["detection_boxes"].cpu().numpy(), wb)


def test_evaluation_loop_metrics():
    """evaluator.evaluate: inference-mode model -> predict -> postprocess -> PASCAL-VOC + MTL metrics on synthetic
    examples (random weights: the values are small, the point is that every stage connects and the bookkeeping is right)."""
    import test_gpu_train_step as T
    from helpers import load_config, randomize_bn
    from mtl_ssl_b200 import evaluator
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    from mtl_ssl_b200.utils.detection_evaluation import ObjectDetectionEvaluation
    H, W = 224, 320
    cfg = load_config("model12.config", T.SMALL)
    model = model_builder.build(cfg.model, False, device="cuda", seed=0)
    model.param_store.load_state_dict(randomize_bn(model.param_store.state_dict(), 0))
    K = cfg.model.faster_rcnn.num_classes
    examples = synthetic.make_batch(31, 3, H, W, K, max_boxes=4, num_windows=16)
    cats = [{"id": i + 1, "name": "class%d" % (i + 1)} for i in range(K)]
    m = evaluator.evaluate(model, examples, cats)
    key = "Subset {:10} mAP@{}IOU".format("default", 0.5)
    assert key in m and (np.isnan(m[key]) or 0.0 <= m[key] <= 1.0)
    assert "CorLoc/CorLoc@0.5IOU" in m
    for k in ("mtl/window_map", "mtl/closeness_diff", "mtl/edgemask_ap"):
        assert k in m and 0.0 <= m[k] <= 1.0, (k, m.get(k))
    # the per-image results are in absolute pixels and feed the evaluator directly
    r = evaluator.run_inference(model, examples[0])
    assert len(r["detection_boxes"]) == len(r["detection_scores"]) == len(r["detection_classes"]) > 0
    assert r["detection_boxes"][:, [0, 2]].max() <= H + 1e-3 and r["detection_boxes"][:, [1, 3]].max() <= W + 1e-3
    assert r["detection_classes"].min() >= 1 and r["detection_classes"].max() <= K
    assert r["window_classes_dt"].shape == (16, K + 1) and r["closeness_dt"].shape[1] == K + 1
    assert r["edgemask_dt"].shape[-1] == 2 and np.abs(r["edgemask_dt"]).max() <= 1.0          # tanh head (mp:111)
    ev = ObjectDetectionEvaluation(K)
    ev.add_single_ground_truth_image_info(0, r["groundtruth_boxes"], r["groundtruth_classes"] - 1)
    ev.add_single_detected_image_info(0, r["detection_boxes"], r["detection_scores"], r["detection_classes"] - 1)
    ap = ev.evaluate()[0]["default"]
    present = np.unique(r["groundtruth_classes"] - 1)
    assert np.all(np.isfinite(ap[present])) and np.all(np.isnan(np.delete(ap, present)))
    # a perfect detector scores mAP 1: detections = ground truth
    lists = {"detection_boxes": [r["groundtruth_boxes"]], "detection_scores": [np.ones(len(r["groundtruth_boxes"]))],
             "detection_classes": [r["groundtruth_classes"]], "image_id": ["0"],
             "groundtruth_boxes": [r["groundtruth_boxes"]], "groundtruth_classes": [r["groundtruth_classes"]]}
    from mtl_ssl_b200 import eval_util
    assert eval_util.evaluate_detection_results_pascal_voc(lists, cats)[key] == 1.0


def test_tf_checkpoint_round_trip_through_a_model(tmp_path):
    """save_tf_checkpoint -> load_tf_checkpoint restores every variable (detection naming, TF layouts), and an
    ImageNet-style checkpoint (stage scopes stripped) initialises all block4 copies from the same keys (T5, T14)."""
    import test_gpu_train_step as T
    from helpers import load_config, randomize_bn
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.utils import checkpoint_io, tf_checkpoint
    cfg = load_config("model12.config", T.SMALL)
    a = model_builder.build(cfg.model, True, device="cuda", seed=0)
    a.param_store.load_state_dict(randomize_bn(a.param_store.state_dict(), 0))
    prefix = str(tmp_path / "model.ckpt-7")
    names = checkpoint_io.save_tf_checkpoint(a, prefix, extra={"global_step": np.asarray(7, np.int64)})
    assert "global_step" in names and not any("/_dead/" in n or "/_pad/" in n for n in names)
    r = tf_checkpoint.CheckpointReader(prefix)
    w = "FirstStageFeatureExtractor/resnet_v1_50/block2/unit_1/bottleneck_v1/conv2/weights"
    assert r.get_variable_to_shape_map()[w] == [3, 3, 128, 128]                      # HWIO on disk
    b = model_builder.build(cfg.model, True, device="cuda", seed=1)
    n, missing = checkpoint_io.load_tf_checkpoint(b, prefix)
    assert not missing and n > 300
    sa, sb = a.param_store.state_dict(), b.param_store.state_dict()
    for k in sa:
        if "/_pad/" in k:
            continue
        assert torch.equal(sa[k].cpu(), sb[k].cpu()), k
    assert torch.equal(a.param_store.by_name[w].wb, b.param_store.by_name[w].wb)     # folded bf16 copy rebuilt
    # classification-style checkpoint: first-stage trunk + ONE block4, stage scopes stripped
    cls = {}
    for k, v in sa.items():
        for sc in ("FirstStageFeatureExtractor/", "SecondStageFeatureExtractor/"):
            if k.startswith(sc) and "/_dead/" not in k and "/_pad/" not in k and "/resnet_v1_50/" in "/" + k:
                cls[k[len(sc):]] = tf_checkpoint.native_to_tf(k, v.cpu().numpy())
    prefix2 = str(tmp_path / "resnet_v1_50.ckpt")
    tf_checkpoint.write_checkpoint(prefix2, cls)
    c = model_builder.build(cfg.model, True, device="cuda", seed=2)
    n2, missing2 = checkpoint_io.load_tf_checkpoint(c, prefix2, from_detection_checkpoint=False)
    assert not missing2
    sc_ = c.param_store.state_dict()
    k2 = "resnet_v1_50/block4/unit_2/bottleneck_v1/conv2/weights"
    for scope in ("SecondStageFeatureExtractor/", "ClosenessBoxPredictor/", "WindowBoxPredictor/",
                  "FirstStageFeatureExtractor/_dead/"):
        assert torch.equal(sc_[scope + k2].cpu(), sa["SecondStageFeatureExtractor/" + k2].cpu()), scope
    k1 = "FirstStageFeatureExtractor/resnet_v1_50/block3/unit_3/bottleneck_v1/conv1/BatchNorm/moving_variance"
    assert torch.equal(sc_[k1].cpu(), sa[k1].cpu())
    # heads are not in a classification checkpoint: untouched (still the seed-2 initialisation)
    kh = "SecondStageBoxPredictor/ClassPredictor/weights"
    assert not torch.equal(sc_[kh].cpu(), sa[kh].cpu())
