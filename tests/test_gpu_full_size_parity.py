"""Loss parity at the BASELINE.json configurations' REAL sizes: the shipped config files unchanged
(ResNet-101 with all 23 block3 units, 13 965 kept anchors through the sort + NMS, 256 + 256 + 64 trained ROIs and the
1 280 forward-only refine windows, K = 20 / 90 heads), one image per replica as in the reference's clones
(trainer.py:157-214), device (bf16 tensor-core convolutions, fp32 accumulation) against the CPU oracle:

  * `Oracle(bf16=True)` -- the oracle with the device's bf16 rounding points mirrored, post-processing the DEVICE's RPN
    outputs so the proposal path is compared index for index: every loss of fmA:1514-1589 within 1e-3 absolute
    (+ 1e-3 relative for the few-unit losses), their sum within 1e-3 relative -- the bar BASELINE.json states;
  * `Oracle(bf16=False)` -- plain fp32, the reference's arithmetic, on the same proposals: the distance is REPORTED
    (`gpurun_out/full_size_parity.json`, copied to profiles/) and bounded at 2e-2 relative on the total: that is the
    bf16 rounding of weights and activations, not an implementation difference.

COCO-shape inputs (800 x 1333) pass through the config's `keep_aspect_ratio_resizer` (600 / 1024) on the device and
through `oracle/nn.resize_bilinear` on the CPU, as in tests/test_gpu_zz_raw_size_inputs.py."""
import json
import os
import time

import numpy as np
import pytest
import torch

from helpers import load_config, oracle_config, randomize_bn

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(1200)]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (config, input H, W, example seed): seeds chosen (CPU oracle search) so that the sampled minibatch has positives,
# i.e. the second-stage localisation and closeness losses are exercised, not identically zero
CASES = [("model12.config", 600, 1000, 2), ("model22.config", 800, 1333, 15), ("model42.config", 800, 1333, 15),
         ("model62.config", 800, 1333, 15)]


def run_case(name, H, W, seed, with_fp32=True):
    """-> dict(device losses, oracle(bf16 mirror) losses, oracle(fp32) losses, counts, seconds)."""
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    from mtl_ssl_b200.meta_architectures.faster_rcnn_meta_arch import LOSS_KEYS
    from mtl_ssl_b200.trainer import Trainer
    from oracle import nn as ON
    from oracle.model import Oracle
    B = 1
    cfg = load_config(name)
    model = model_builder.build(cfg.model, True, device="cuda", seed=0)
    sd = randomize_bn(model.param_store.state_dict(), 0)
    model.param_store.load_state_dict(sd)
    K, M = cfg.model.faster_rcnn.num_classes, cfg.model.faster_rcnn.first_stage_max_proposals
    tr = Trainer(model, None, H, W, B, gmax=8, use_cuda_graph=False)
    tr.overlap_optimizer = False
    Hr, Wr = tr.Hr, tr.Wr
    examples = synthetic.make_batch(seed, B, H, W, K, max_boxes=8, num_windows=64)
    keys = synthetic.make_sampler_keys(seed + 1000, B, model.num_kept_anchors((B, H, W, 3)), M)
    arrays = tr.host_arrays(examples, keys)
    image = tr._bind(arrays)
    pd = tr._forward_backward(image)
    torch.cuda.synchronize()
    got = dict(zip(LOSS_KEYS, model.workspace.bufs["loss/values"].cpu().tolist()))
    images = torch.from_numpy(arrays["image"])
    if (Hr, Wr) != (H, W):
        images = ON.resize_bilinear(images, (Hr, Wr))
    wsb = model.workspace.bufs
    # the oracle's sort / NMS / sampler start from the device's decoded boxes + scores (index-exact claim); the decode
    # itself (exp, softmax: an ulp apart between libm and the device) is compared on its own below
    prop_in = (pd["rpn_box_encodings"].cpu().numpy(), pd["rpn_objectness_predictions_with_background"].cpu().numpy(),
               wsb["rpn/dec_boxes"].cpu().numpy(), wsb["rpn/dec_scores"].cpu().numpy())
    params = {k: v for k, v in sd.items() if "/_pad/" not in k and "/_dead/" not in k}
    res = dict(config=name, input=[H, W], resized=[Hr, Wr], device=got, kept_anchors=int(pd["_Nk"]),
               num_proposals=pd["num_proposals"].cpu().tolist())
    res["_dec_boxes_device"], res["_dec_scores_device"] = prop_in[2][0], prop_in[3][0]
    for tag, bf in (("oracle_bf16_mirror", True),) + ((("oracle_fp32", False),) if with_fp32 else ()):
        t0 = time.perf_counter()
        orc = Oracle(params, oracle_config(cfg), bf16=bf)
        with torch.no_grad():
            out = orc.forward(images, examples, keys, Hr, Wr, proposal_inputs=prop_in)
            want = {k: float(v) for k, v in orc.loss(out, examples, keys, Hr, Wr).items()}
        res[tag] = want
        res[tag + "_seconds"] = time.perf_counter() - t0
        if bf:
            res["_nprop_oracle"] = out["nprop"].tolist()
            res["_prop_abs_oracle"] = out["prop_abs"]
            res["_prop_abs_device"] = pd["proposal_boxes"].cpu().numpy()
            from oracle import boxes as OB
            from oracle import postprocess as OP
            dec = OB.clip_to_window(OB.box_decode(prop_in[0][0], out["anchors"]), (0.0, 0.0, float(Hr), float(Wr)),
                                    filter_nonoverlapping=False)[0]
            res["_decode"] = (dec, OP.softmax(prop_in[1][0])[:, 1])
            # the same chain with the oracle's OWN decode / softmax: how many of the proposals an ulp of exp() moves
            with torch.no_grad():
                own = orc.forward(images, examples, keys, Hr, Wr, proposal_inputs=prop_in[:2])
            res["proposals_moved_by_libm_vs_device_exp"] = int(
                (np.abs(own["prop_abs"] - res["_prop_abs_device"]).max(-1) > 1e-3).sum())
            if not np.allclose(res["_prop_abs_device"], res["_prop_abs_oracle"], rtol=1e-5, atol=1e-3):
                # keep what an offline look at the disagreement needs (NMS survivors + sampler state of both sides)
                os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                np.savez_compressed(
                    os.path.join(ROOT, "gpurun_out", "full_size_mismatch_%s.npz" % name.split(".")[0]),
                    rpn_box=prop_in[0], rpn_cls=prop_in[1], anchors=out["anchors"], keys2=np.asarray(keys[1]),
                    dev_nms_boxes=wsb["rpn/nms_boxes"].cpu().numpy(), dev_nms_scores=wsb["rpn/nms_scores"].cpu().numpy(),
                    dev_nms_num=wsb["rpn/nms_num"].cpu().numpy(), dev_sampled=wsb["det/sampled"].cpu().numpy(),
                    dev_match=wsb["det/sample_match"].cpu().numpy(), dev_prop=res["_prop_abs_device"],
                    orc_prop=res["_prop_abs_oracle"], orc_nms_boxes=out["nms"][0][0], orc_nms_scores=out["nms"][0][1],
                    orc_nms_num=np.asarray(out["nms"][0][2]), dev_dec_boxes=wsb["rpn/dec_boxes"].cpu().numpy(),
                    dev_dec_scores=wsb["rpn/dec_scores"].cpu().numpy(), dev_order=wsb["rpn/order"].cpu().numpy(),
                    dev_nvalid=wsb["rpn/nvalid"].cpu().numpy(),
                    gt=np.asarray(examples[0]["groundtruth_boxes"]), hw=np.asarray([Hr, Wr]))
            pos = out["targets"]["reg_w"].sum().item()
            res["positives"] = int(pos)
    keys_ = list(res["oracle_bf16_mirror"])
    res["total_device"] = sum(got[k] for k in keys_)
    res["total_oracle_bf16_mirror"] = sum(res["oracle_bf16_mirror"].values())
    if with_fp32:
        res["total_oracle_fp32"] = sum(res["oracle_fp32"].values())
    del tr, model
    torch.cuda.empty_cache()
    return res


def _record(res):
    if not torch.cuda.is_available():       # (dry-run lint of this file on a CPU box: nothing measured, nothing to record)
        return
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    path = os.path.join(out, "full_size_parity.json")
    try:
        allr = json.load(open(path))
    except Exception:
        allr = {}
    allr[res["config"]] = {k: v for k, v in res.items() if not k.startswith("_")}
    json.dump(allr, open(path, "w"), indent=1, sort_keys=True)


@pytest.mark.parametrize("name,H,W,seed", CASES)
def test_full_size_losses_match_oracle(name, H, W, seed, request):
    if request.config.getoption("--dry-run-gpu") and name != "model12.config":
        pytest.skip("dry-run lint of this body: one case is enough (the CPU oracle at full size takes minutes)")
    res = run_case(name, H, W, seed)
    _record(res)
    got, want, w32 = res["device"], res["oracle_bf16_mirror"], res["oracle_fp32"]
    # the proposal path at its real size (13 965 candidates -> sort -> NMS -> sampler), index for index
    assert res["num_proposals"] == res["_nprop_oracle"]
    np.testing.assert_allclose(res["_prop_abs_device"], res["_prop_abs_oracle"], rtol=1e-5, atol=1e-3)
    assert res["positives"] > 0, "the case must exercise the localisation / closeness losses"
    # decode + objectness softmax of all 13 965 kept anchors: exp() to an ulp (2e-6 relative, 1e-3 px on clipped boxes)
    dec, sc = res["_decode"]
    np.testing.assert_allclose(res["_dec_scores_device"], sc, rtol=2e-6, atol=1e-7)
    np.testing.assert_allclose(res["_dec_boxes_device"], dec, rtol=2e-6, atol=2e-3)
    for k, v in want.items():
        assert abs(got[k] - v) <= 1e-3 + 1e-3 * abs(v), (k, got[k], v)
    assert abs(res["total_device"] - res["total_oracle_bf16_mirror"]) <= 1e-3 * max(1.0, abs(res["total_oracle_bf16_mirror"]))
    # distance to the reference's fp32 arithmetic: reported, and bounded (bf16 weights + activations)
    assert abs(res["total_device"] - res["total_oracle_fp32"]) <= 2e-2 * max(1.0, abs(res["total_oracle_fp32"])), \
        (res["total_device"], res["total_oracle_fp32"])
    for k, v in w32.items():
        assert abs(got[k] - v) <= 2e-2 + 3e-2 * abs(v), (k, got[k], v)
