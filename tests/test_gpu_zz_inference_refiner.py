"""Class refiner at inference time (the reference's evaluator, evaluator.py:145-152: `predict_with_mtl_results` on the
inference proposals, `postprocess` on the refined logits, fmA:1040-1043) against `Oracle.forward(inference=True,
inference_mtl=True)`.  Written at the end of round 1 after the GPU budget was spent: this file sorts last so that its
first run on a device cannot mask the established parity tests; `evaluator.run_inference(use_refiner=True)` is opt-in
until it has passed there."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]      # first run on a device: never hang the suite


def _cos(a, b):
    a, b = a.reshape(-1).double(), b.reshape(-1).double()
    return float(a @ b / (a.norm() * b.norm()).clamp_min(1e-30))


def test_inference_refiner_matches_oracle_and_feeds_postprocess():
    import test_gpu_train_step as T
    from helpers import load_config, oracle_config, randomize_bn
    from mtl_ssl_b200 import evaluator
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    from oracle.model import Oracle
    H, W, B = 224, 320, 1
    cfg = load_config("model12.config", T.SMALL)
    model = model_builder.build(cfg.model, False, device="cuda", seed=0)
    sd = randomize_bn(model.param_store.state_dict(), 0)
    # the refiner's FC layer starts at zero-ish weights: make it matter
    g = torch.Generator().manual_seed(5)
    for k in sd:
        if k.startswith("MTLClassRefiner/") and k.endswith("weights"):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.05
    model.param_store.load_state_dict(sd)
    K = cfg.model.faster_rcnn.num_classes
    ex = synthetic.make_batch(7, B, H, W, K, max_boxes=4, num_windows=16)
    images = np.stack([e["image"] for e in ex]).astype(np.float32)
    image = torch.from_numpy(images).cuda()
    pd = model.predict(model.preprocess(image))
    plain = model.postprocess(pd)
    plain = {k: v.clone() for k, v in plain.items()}
    pd = model.predict_with_mtl_results(pd)
    det = model.postprocess(pd)
    torch.cuda.synchronize()
    P = model.max_num_proposals
    orc = Oracle({k: v for k, v in sd.items() if "/_pad/" not in k}, oracle_config(cfg), bf16=True)
    prop_in = (pd["rpn_box_encodings"].cpu().numpy(), pd["rpn_objectness_predictions_with_background"].cpu().numpy())
    with torch.no_grad():
        out = orc.forward(torch.from_numpy(images), None, None, H, W, proposal_inputs=prop_in, inference=True,
                          inference_mtl=True)
    assert np.array_equal(pd["num_proposals"].cpu().numpy(), out["nprop"])
    n = int(out["nprop"][0])
    assert 0 < n <= P
    ref = pd["mtl_refined_class_predictions_with_background"].float().cpu()
    assert ref.shape == (B * P, K + 1)
    cls = pd["class_predictions_with_background"].float().cpu()
    # bf16 activations on the device, mirrored rounding points in the oracle: same bar as the training-step heads
    assert _cos(pd["closeness_predictions"].float().cpu()[:n], out["closeness_predictions"].float()[:n]) > 0.999
    assert _cos(ref[:n], out["mtl_refined_class_predictions_with_background"].float()[:n]) > 0.999
    assert _cos((ref - cls)[:n], (out["mtl_refined_class_predictions_with_background"]
                                  - out["class_predictions_with_background"]).float()[:n]) > 0.98
    assert float((ref - cls)[:n].abs().max()) > 1e-3               # the refiner changed the logits ...
    assert not torch.equal(det["detection_scores"], plain["detection_scores"])      # ... and the detections follow
    nd = int(det["num_detections"][0].item())
    s = det["detection_scores"][0].cpu().numpy()
    assert nd > 0 and np.all(np.diff(s[:nd]) <= 0) and not s[nd:].any()
    # the evaluator's opt-in switch takes the same path
    r0 = evaluator.run_inference(model, ex[0])
    r1 = evaluator.run_inference(model, ex[0], use_refiner=True)
    assert abs(len(r1["detection_scores"]) - nd) <= 1
    assert not np.array_equal(r0["detection_scores"][:5], r1["detection_scores"][:5])
    np.testing.assert_allclose(r1["detection_scores"][:5], s[:5], rtol=0, atol=1e-4)
