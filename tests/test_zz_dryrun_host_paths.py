"""Host orchestration of the hot path executed WITHOUT a GPU (tests/dryrun.py): every C-ABI call is marshalled against
the signatures of include/mtlssl.h, the library entry point is a stub, tensors live on the CPU.  Checks the Python side
-- buffer shapes, dictionary keys, call sequences, stream bookkeeping -- of paths that the -m gpu tests cover numerically,
and of the two paths written after this round's GPU budget was spent (refiner at inference time, shape buckets)."""
import numpy as np
import pytest
import torch

import dryrun
from helpers import load_config

SMALL = (("type: 'faster_rcnn_resnet101'", "type: 'faster_rcnn_resnet50'"),
         ("min_dimension: 600", "min_dimension: 224"), ("max_dimension: 1024", "max_dimension: 320"),
         ("first_stage_max_proposals: 300", "first_stage_max_proposals: 100"),
         ("second_stage_batch_size: 256", "second_stage_batch_size: 32"))


def _batches(model, trainer, shapes, K=20, M=100):
    from mtl_ssl_b200.data import synthetic
    for i, hw in enumerate(shapes):
        ex = synthetic.make_batch(60 + i, 1, hw[0], hw[1], K, max_boxes=4, num_windows=16)
        ky = synthetic.make_sampler_keys(70 + i, 1, model.num_kept_anchors((1, hw[0], hw[1], 3)), M)
        yield trainer.host_arrays(ex, ky)


@pytest.mark.parametrize("name", ["model12.config", "model42.config", "model52.config", "model62.config"])
def test_training_step_call_sequence(monkeypatch, name):
    """One eager training step per architecture family: Faster R-CNN ResNet, R-FCN, MobileNet, Inception-ResNet-v2."""
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.trainer import Trainer
    log = dryrun.install(monkeypatch)
    cfg = load_config(name, SMALL[1:] if name[5] in "56" else SMALL)
    model = model_builder.build(cfg.model, True, device="cpu", seed=0)
    tr = Trainer(model, cfg.train_config, 224, 320, 1, gmax=8, use_cuda_graph=False)
    (arrays,) = list(_batches(model, tr, [(224, 320)], cfg.model.faster_rcnn.num_classes))
    del log[:]
    losses = tr.step(arrays)
    assert set(losses) >= {"first_stage_localization_loss", "refined_classification_loss", "total_loss"}
    launched = set(log)
    assert {"mtl_conv_tc", "mtl_rpn_decode", "mtl_nms", "mtl_iou_match", "mtl_balanced_sample", "mtl_rpn_loss",
            "mtl_box_classifier_loss", "mtl_opt_stats", "mtl_opt_apply"} <= launched or \
        {"mtl_opt_stats_range", "mtl_opt_apply"} <= launched
    assert ("mtl_psroi_fwd" in launched) == (name == "model42.config")
    assert ("mtl_dwconv3x3_fwd" in launched) == (name == "model52.config")
    assert tr.global_step == 1 and log.count("mtl_conv_tc") > 50


def test_inference_with_refiner_call_sequence(monkeypatch):
    """evaluator.run_inference(use_refiner=True): closeness head, 5 x P expanded windows through the window tail, refiner
    concat + FC, then `postprocess` (which reads the refined logits, fmA:1040-1043)."""
    from mtl_ssl_b200 import evaluator
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    log = dryrun.install(monkeypatch)
    cfg = load_config("model12.config", SMALL)
    model = model_builder.build(cfg.model, False, device="cpu", seed=0)
    ex = synthetic.make_batch(7, 1, 224, 320, 20, max_boxes=4, num_windows=16)
    del log[:]
    r0 = evaluator.run_inference(model, ex[0])
    plain = list(log)
    del log[:]
    r1 = evaluator.run_inference(model, ex[0], use_refiner=True)
    assert "mtl_expand_windows" not in plain and "mtl_refine_concat" not in plain
    for k in ("mtl_expand_windows", "mtl_refine_concat", "mtl_fc_fwd", "mtl_detection_decode", "mtl_detection_gather"):
        assert k in log, k
    assert log.index("mtl_refine_concat") < log.index("mtl_detection_decode")        # refined logits feed postprocess
    assert log.count("mtl_crop_and_resize_fwd") == plain.count("mtl_crop_and_resize_fwd") + 1     # the 5 x P windows
    assert r1["closeness_dt"].shape == (100, 21) and r1["window_classes_dt"].shape == (16, 21)
    assert set(r0) == set(r1)


@pytest.mark.parametrize("graph", [True])       # (the eager variant runs inside the GPU-suite dry run below)
def test_mixed_shape_training_loop_with_checkpoints(monkeypatch, tmp_path, graph):
    """shape_buckets.ShapeBucketTrainer under data.loader.train_loop: two image shapes alternate over one model, the
    training state is saved and restored through the TF checkpoint format (real ParamStore, CPU tensors)."""
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import loader
    from mtl_ssl_b200.shape_buckets import ShapeBucketTrainer
    from mtl_ssl_b200.trainer import Trainer
    from mtl_ssl_b200.utils import checkpoint_io
    dryrun.install(monkeypatch)
    cfg = load_config("model12.config", SMALL)
    model = model_builder.build(cfg.model, True, device="cpu", seed=0)
    bt = ShapeBucketTrainer(model, cfg.train_config, batch_size=1, max_buckets=2, use_cuda_graph=graph, gmax=8)
    shapes = [(224, 320), (224, 288), (224, 320), (320, 224)]
    prefix = str(tmp_path / "model.ckpt")
    model.param_store.m.fill_(0.25)                # (the stubbed optimizer never touches the momenta)
    out = loader.train_loop(bt, _batches(model, bt, shapes), checkpoint_prefix=prefix)       # one save, at the end
    assert len(out) == 4 and bt.global_step == 4 and bt.evictions == 1 and list(bt.buckets) == [(224, 320), (320, 224)]
    assert checkpoint_io.latest_checkpoint(str(tmp_path)) == prefix + "-4"
    if graph:                  # every bucket captured its own graphs; the revisited shape replayed them
        assert all(tr.graph_fb is not None and tr.graph_opt is not None for tr, _ in bt.buckets.values())
        assert bt.buckets[(224, 320)][0].graph_fb.replays >= 1
    model2 = model_builder.build(cfg.model, True, device="cpu", seed=1)
    tr2 = Trainer(model2, cfg.train_config, 224, 320, 1, gmax=8, use_cuda_graph=False)
    n, slots, step = checkpoint_io.restore_training_checkpoint(tr2, prefix + "-4")
    assert step == 4 and slots == sum(1 for p in model.param_store.params if p.trainable and "/_pad/" not in p.name)
    assert torch.equal(model2.param_store.w, model.param_store.w)
    trainable = [p for p in model2.param_store.params if p.trainable and "/_pad/" not in p.name]
    assert all(bool((p.m == 0.25).all()) for p in trainable)


def test_raw_size_images_are_resized_on_the_device(monkeypatch):
    """BASELINE.json configs[0]: model51.config UNCHANGED (MobileNet, min_dimension 600) on two 300x300 images.  The
    trainer takes the images as they arrive; `preprocess` resizes them on the device (mtl_resize_bilinear_f32) and the
    anchors, the absolute ground-truth boxes and the sampler keys all belong to the resized 600x600 image."""
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    from mtl_ssl_b200.trainer import Trainer
    log = dryrun.install(monkeypatch)
    cfg = load_config("model51.config")
    model = model_builder.build(cfg.model, True, device="cpu", seed=0)
    assert model._image_resizer_fn.static_size(300, 300) == (600, 600)
    tr = Trainer(model, cfg.train_config, 300, 300, 2, gmax=8, use_cuda_graph=False)
    assert (tr.H, tr.W, tr.Hr, tr.Wr) == (300, 300, 600, 600)
    K, M = cfg.model.faster_rcnn.num_classes, cfg.model.faster_rcnn.first_stage_max_proposals
    nk = model.num_kept_anchors((2, 300, 300, 3))
    assert 0 < nk < 38 * 38 * 12                 # anchors of the 600x600 -> 38x38 map that lie inside the image
    ex = synthetic.make_batch(3, 2, 300, 300, K, max_boxes=4, num_windows=16)
    arrays = tr.host_arrays(ex, synthetic.make_sampler_keys(4, 2, nk, M))
    assert arrays["image"].shape == (2, 300, 300, 3)
    np.testing.assert_allclose(arrays["gt"][0, 0], np.asarray(ex[0]["groundtruth_boxes"][0]) * 600.0, rtol=1e-6)
    del log[:]
    tr.step(arrays)
    assert log.count("mtl_resize_bilinear_f32") >= 1 and model._gt_shape == (2, 600, 600, 3)
    assert model._last_pd["image_shape"] == (2, 600, 600, 3) and model._last_pd["_feat_hw"] == (38, 38)


def test_evaluation_loop_call_sequence(monkeypatch):
    """evaluator.evaluate on the inference-mode graph: metrics assemble from (empty) detections without a device."""
    from mtl_ssl_b200 import evaluator
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    dryrun.install(monkeypatch)
    cfg = load_config("model12.config", SMALL)
    model = model_builder.build(cfg.model, False, device="cpu", seed=0)
    examples = synthetic.make_batch(31, 2, 224, 320, 20, max_boxes=4, num_windows=16)
    cats = [{"id": i + 1, "name": "class%d" % (i + 1)} for i in range(20)]
    m = evaluator.evaluate(model, examples, cats, use_refiner=True)
    assert "Subset default    mAP@0.5IOU" in m and "mtl/window_map" in m and "mtl/edgemask_ap" in m


def test_bench_json_line_contract(monkeypatch, capsys):
    """bench.py `run_ours` end to end in dry-run mode (eager, one replica): the ONE JSON line carries every key of the
    driver's contract plus `roofline` (with the dominant launch group) -- values are meaningless here, the point is that
    the script reaches its print statement whatever was edited last."""
    import json
    import types
    import bench
    import mtl_ssl_b200.builders.model_builder as mb
    from mtl_ssl_b200 import ops
    log = dryrun.install(monkeypatch)
    monkeypatch.setattr(torch.cuda, "set_device", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "_sleep", lambda *a, **k: None, raising=False)
    monkeypatch.setattr(dryrun.FakeEvent, "elapsed_time", lambda self, other: 1.0)       # 1 ms per event pair
    real_build = mb.build
    monkeypatch.setattr(mb, "build", lambda cfg, tr, device=None, seed=0: real_build(cfg, tr, device="cpu", seed=seed))
    monkeypatch.setattr(ops, "launch_count", lambda: len(log))
    args = types.SimpleNamespace(steps=2, warmup=1, batch_per_gpu=0, no_graph=True, skip_cpu=True, gpus=1, impl="ours")
    bench.run_ours(args, 0, 1, 0)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in line, k
    assert line["metric"] == "images/sec" and line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] >= 3
    assert line["config"]["workload"].startswith("Faster R-CNN ResNet-101") and line["vs_baseline"] is None
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] > 7e6
    r = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "dominant"} <= set(r) and r["bound"] == "tensor"
    assert r["dominant"]["kernel"].startswith("tc_gemm_kernel ") and r["dominant"]["launches"] >= 1
    assert 250 <= line["gpu_launches_per_step"] <= 480            # r1: 431 on the device; r2 groups the weight-gradient GEMMs


_DP_WORKER = r"""
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import pytest, torch, torch.distributed as dist
import dryrun
from helpers import load_config
mp = pytest.MonkeyPatch()
dryrun.install(mp)
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
from mtl_ssl_b200.builders import model_builder
from mtl_ssl_b200.data import synthetic
from mtl_ssl_b200.trainer import Trainer
SMALL = %(small)r
cfg = load_config("model12.config", SMALL)
model = model_builder.build(cfg.model, True, device="cpu", seed=0)
st = model.param_store
for overlap, late in ((True, False), (True, True), (False, False)):
    tr = Trainer(model, cfg.train_config, 224, 320, 1, gmax=8, use_cuda_graph=False, world_size=2)
    tr.overlap_optimizer = overlap
    tr._hw_late = late          # late: head weight gradients after the trunk update, head bucket exchanged in 3 pieces
    if late:
        pieces = model.head_exchange_pieces(3)
        assert len(pieces) == 3 and sum(g.numel() for g, _ in pieces) == model.gradient_buckets()[0].numel()
        assert pieces[0][0].data_ptr() == model.gradient_buckets()[0].data_ptr()
        assert [a for _, (a, b) in pieces][1:] == [b for _, (a, b) in pieces][:-1]
    # the stubbed kernels leave the gradient arena untouched: plant rank-specific values where the two halves of the
    # backward pass would have written them, and look at what the exchange made of them
    fb, bt = tr._forward_backward, tr._backward_trunk
    heads, trunk = model.gradient_buckets()
    def fb2(image, prefix=None):
        r = fb(image, prefix)
        st.g.fill_(float(rank + 1))
        trunk.fill_(-1.0)                      # not final yet: must not be exchanged with the head bucket
        return r
    def bt2():
        bt()
        trunk.fill_(float(10 * (rank + 1)))
    tr._forward_backward, tr._backward_trunk = fb2, bt2
    # deferred-head schedule (overlap on): one pass fills both buckets; the trunk bucket is exchanged right after it,
    # the head bucket in finish() -- after the deferred weight-gradient launches, before the head update
    sb, sb_lo = tr._stage_t, tr._stage_t2
    _, t_hi, t_lo = model.gradient_buckets3()
    assert t_hi.numel() + t_lo.numel() == trunk.numel() and t_hi.numel() > 0 and t_lo.numel() > 0
    def sb2():
        r = sb()
        st.g.fill_(float(rank + 1))
        t_hi.fill_(float(10 * (rank + 1)))
        t_lo.fill_(-1.0)                       # not final yet: must not travel with the first trunk bucket
        return r
    def sb3():
        sb_lo()
        t_lo.fill_(float(10 * (rank + 1)))
    tr._stage_t, tr._stage_t2 = sb2, sb3
    assert tr._deferred() == overlap
    ex = synthetic.make_batch(60 + rank, 1, 224, 320, 20, max_boxes=4, num_windows=16)
    ky = synthetic.make_sampler_keys(70 + rank, 1, model.num_kept_anchors((1, 224, 320, 3)), 100)
    tr.step(tr.host_arrays(ex, ky))
    assert heads.numel() + trunk.numel() <= st.total and heads.numel() > 0 and trunk.numel() > 0
    assert bool((heads == 3.0).all()), heads.unique()            # 1 + 2: summed over the two replicas, once
    assert bool((trunk == 30.0).all()), trunk.unique()           # 10 + 20
    dead = st.g[heads.numel() + trunk.numel():]
    assert dead.numel() > 0 and bool((dead == float(rank + 1)).all())      # the dead block4 copy is not exchanged (T4)
    st.g.zero_()
# the captured variant (stand-in graphs): capture of the two backward halves, the head optimizer and the trunk optimizer,
# then pipelined steps with the all-reduces issued between the replays
for overlap, late in ((True, False), (True, True), (False, False)):
    tr = Trainer(model, cfg.train_config, 224, 320, 1, gmax=8, use_cuda_graph=True, world_size=2)
    tr.overlap_optimizer = overlap
    tr._hw_late = late
    ex = synthetic.make_batch(80 + rank, 1, 224, 320, 20, max_boxes=4, num_windows=16)
    ky = synthetic.make_sampler_keys(90 + rank, 1, model.num_kept_anchors((1, 224, 320, 3)), 100)
    arrays = tr.host_arrays(ex, ky)
    outs = [tr.step_pipelined(arrays) for _ in range(3)] + [tr.flush()]
    assert outs[0] is None and all(o is not None and "total_loss" in o for o in outs[1:])
    assert tr.graph_fb.replays >= 2 and tr.graph_opt.replays >= 2
    if overlap:         # deferred heads: first-stage graph, second-stage + backward graphs, deferred wgrads, head update
        assert tr.graph_fa.replays >= 2 and tr.graph_opt_heads.replays >= 2
        if late:
            assert len(tr.graph_hw_pieces) == 3 and all(g.replays >= 2 for g in tr.graph_hw_pieces)
        else:
            assert tr.graph_hw.replays >= 2
        assert tr.graph_ft.replays >= 2 and tr.graph_ft2.replays >= 2
        assert not tr._heads_pending                                  # flush() applied the last head update
    else:
        assert tr.graph_fb2.replays >= 2 and tr.graph_opt_heads is None
    assert tr.global_step == 3
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_replica_step_exchanges_each_gradient_bucket_once_gloo(tmp_path):
    """The N > 1 step body of Trainer (head bucket all-reduced under the trunk backward, trunk bucket after it, the
    dead stage-1 block4 copy skipped) on two gloo processes in dry-run mode: every exchanged element is summed exactly
    once, with and without the overlapped head optimizer."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "dp_worker.py"
    script.write_text(_DP_WORKER % dict(root=root, small=SMALL))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), "\n".join(outs)


@pytest.mark.parametrize("name", ["model42.config", "model52.config", "model22.config", "model11.config"])
def test_inference_with_refiner_other_architectures(monkeypatch, name):
    """R-FCN (PS-ROI windows), MobileNet, the COCO config (K = 90, crop 14 + max pool) and a baseline config without
    auxiliary heads (nothing to refine) through evaluator.run_inference(use_refiner=True)."""
    from mtl_ssl_b200 import evaluator
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    log = dryrun.install(monkeypatch)
    cfg = load_config(name, SMALL[1:] if name[5] in "56" else SMALL)
    K = cfg.model.faster_rcnn.num_classes
    model = model_builder.build(cfg.model, False, device="cpu", seed=0)
    ex = synthetic.make_batch(7, 1, 224, 320, K, max_boxes=4, num_windows=16)
    del log[:]
    r = evaluator.run_inference(model, ex[0], use_refiner=True)
    refines = bool(cfg.model.mtl.refine)
    assert ("mtl_refine_concat" in log) == refines and ("mtl_expand_windows" in log) == (refines and bool(cfg.model.mtl.window))
    assert ("mtl_psroi_fwd" in log) == (name == "model42.config")
    assert "mtl_detection_gather" in log and r["detection_boxes"].shape[1:] == (4,)


def test_full_size_step_launches_what_the_committed_profile_shows(monkeypatch):
    """BASELINE configs[1] (model12.config unchanged, 600x1000, batch 1): the step issues as many tcgen05 conv launches
    as the committed ncu launch list holds `tc_gemm_kernel` rows (profiles/r1_launches_step_final.csv: 374), so the
    profile the roofline numbers come from still describes the code."""
    import collections
    import csv
    import os
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    from mtl_ssl_b200.trainer import Trainer
    log = dryrun.install(monkeypatch)
    cfg = load_config("model12.config")
    model = model_builder.build(cfg.model, True, device="cpu", seed=0)
    tr = Trainer(model, cfg.train_config, 600, 1000, 1, gmax=16, use_cuda_graph=False)
    ex = synthetic.make_batch(1, 1, 600, 1000, 20, max_boxes=8, num_windows=64)
    ky = synthetic.make_sampler_keys(2, 1, model.num_kept_anchors((1, 600, 1000, 3)), 300)
    arrays = tr.host_arrays(ex, ky)
    # bench.py's e2e.h2d_bytes_per_step on the device; 13 965 of the 28 728 anchors lie inside the image
    assert model.num_kept_anchors((1, 600, 1000, 3)) == 13965
    assert sum(v.nbytes for v in arrays.values()) == 7297896
    tr.step(arrays)
    del log[:]
    tr.step(arrays)
    c = collections.Counter(log)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # round 2: the 118 weight-gradient GEMMs no longer launch one by one (ops_conv.WgradCollector groups them) and the
    # box heads' FC layers are fused with their pooling (csrc/head.cu), so the step issues 249 individual tcgen05
    # launches plus a handful of grouped ones
    rows = [r for r in csv.reader(open(os.path.join(root, "profiles", "r2_launches_step.csv"), errors="replace"))]
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    ki = rows[hdr].index("Kernel Name")
    profiled = collections.Counter("tc_gemm" if "tc_gemm_kernel" in r[ki] else "other" for r in rows[hdr + 1:]
                                   if len(r) > ki)
    assert c["mtl_conv_tc"] == 249 and 1 <= c["mtl_conv_tc_group_launch"] <= 12
    assert 249 < profiled["tc_gemm"] <= 249 + 12          # 258 on the device: 249 + 9 grouped launches (the stubbed
    #                                                        group key of the dry run merges some of them)
    # (the refiner's forward-only window pass hands its head the partial row sums its last conv wrote: no feature maps)
    assert c["mtl_head_fwd"] == 3 and c["mtl_head_fwd_pooled"] == 1 and c["mtl_head_bwd"] == 3
    assert c["mtl_nms"] == 1 and c["mtl_crop_and_resize_fwd"] == 3 and c["mtl_expand_windows"] == 1
    other = sum(v for k, v in c.items() if not k.startswith("mtl_conv_tc"))
    assert abs(other - profiled["other"]) <= 16          # a few library copies / casts differ


def test_algorithmic_flops_of_the_step_match_the_roofline_basis(monkeypatch):
    """2.90 TFLOP forward + 2.02 TFLOP backward per 600x1000 image.  The per-launch algorithmic FLOPs that
    bench.py divides by the measured kernel time (ops_conv.PROFILE) sum to that figure over the launches of a step (the grouped weight-gradient launches counted with their sums),
    and to bench.algorithmic_flops_per_image() which `roofline.step_tflops` uses."""
    import bench
    from mtl_ssl_b200 import ops_conv
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    from mtl_ssl_b200.trainer import Trainer
    dryrun.install(monkeypatch)
    cfg = load_config("model12.config")
    model = model_builder.build(cfg.model, True, device="cpu", seed=0)
    tr = Trainer(model, cfg.train_config, 600, 1000, 1, gmax=16, use_cuda_graph=False)
    ex = synthetic.make_batch(1, 1, 600, 1000, 20, max_boxes=8, num_windows=64)
    ky = synthetic.make_sampler_keys(2, 1, model.num_kept_anchors((1, 600, 1000, 3)), 300)
    arrays = tr.host_arrays(ex, ky)
    tr.step(arrays)
    monkeypatch.setattr(ops_conv, "PROFILE", [])
    tr.step(arrays)
    prof = ops_conv.PROFILE
    by_mode = [sum(p[1] for p in prof if p[0] == m) / 1e12 for m in range(3)]            # fprop, dgrad, wgrad
    # 249 individual launches + the grouped weight-gradient launches (each recorded once with its summed FLOPs); the
    # box heads' FC layers (0.002 TFLOP) run in the fused head kernels, outside this list
    assert 249 < len(prof) <= 249 + 12 and sum(1 for p in prof if p[0] == 2) <= 12
    assert abs(by_mode[0] - 2.90) < 0.005 and abs(by_mode[1] + by_mode[2] - 2.02) < 0.005
    assert abs(sum(by_mode) * 1e12 - bench.algorithmic_flops_per_image()) < 1e-3 * bench.algorithmic_flops_per_image()


@pytest.mark.parametrize("name,B", [("model22.config", 2), ("model42.config", 1), ("model62.config", 1)])
def test_coco_shape_inputs_of_the_other_baseline_configs(monkeypatch, name, B):
    """BASELINE.json configs[2..4] (Faster R-CNN ResNet-101 + aux heads at 2 images per GPU, R-FCN, Inception-ResNet-v2)
    on COCO-shape 800x1333 inputs with the configs UNCHANGED: their `keep_aspect_ratio_resizer` (600 / 1024) brings the
    images to 600x1000 on the device, the stride-16 map is 38x63, one step runs host-side end to end."""
    from mtl_ssl_b200 import ops_conv
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.data import synthetic
    from mtl_ssl_b200.trainer import Trainer
    log = dryrun.install(monkeypatch)
    cfg = load_config(name)
    K = cfg.model.faster_rcnn.num_classes
    model = model_builder.build(cfg.model, True, device="cpu", seed=0)
    H, W = 800, 1333
    tr = Trainer(model, cfg.train_config, H, W, B, gmax=32, use_cuda_graph=False)
    assert (tr.Hr, tr.Wr) == (600, 1000)
    ex = synthetic.make_batch(1, B, H, W, K, max_boxes=8, num_windows=64)
    ky = synthetic.make_sampler_keys(2, B, model.num_kept_anchors((B, H, W, 3)), 300)
    monkeypatch.setattr(ops_conv, "PROFILE", [])
    losses = tr.step(tr.host_arrays(ex, ky))
    assert "total_loss" in losses and model._last_pd["_feat_hw"] == (38, 63)
    assert model._last_pd["image_shape"] == (B, 600, 1000, 3) and "mtl_resize_bilinear_f32" in log
    flops = sum(p[1] for p in ops_conv.PROFILE) / 1e12
    if name == "model22.config":          # 2 images of the VOC-shape workload, K = 90, 14x14 crops + 2x2 max pool
        assert 9.5 < flops < 10.0 and model.workspace.nbytes() < 9e9


@pytest.mark.parametrize("name", ["model12.config", "model52.config", "model62.config"])
def test_classification_checkpoint_name_map_for_every_backbone(monkeypatch, tmp_path, name):
    """trainer.py:311-356 + `restore_map` (fmA:1947-2013; incres fe:173-248): an ImageNet-classification checkpoint
    (stage scopes stripped) initialises the first-stage trunk and EVERY copy of the second-stage tail (main, closeness,
    window, and the dead stage-1 copy: T5, T14) -- for ResNet, MobileNet and Inception-ResNet-v2.
    Real ParamStore on CPU tensors (dry run); the TF layouts go through the tensor-bundle writer / reader."""
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.utils import checkpoint_io, tf_checkpoint
    dryrun.install(monkeypatch)
    cfg = load_config(name, SMALL[1:] if name[5] in "56" else SMALL)
    model = model_builder.build(cfg.model, True, device="cpu", seed=0)
    name_map = checkpoint_io.variable_name_map(model, from_detection_checkpoint=False)
    st = model.param_store
    shapes = {p.name: tuple(p.shape) for p in st.params}
    kinds = {p.name: p.tf_kind for p in st.params}
    for b in st.bns:
        if b.scope is not None:
            for k in ("gamma", "beta", "moving_mean", "moving_variance"):
                shapes[b.scope + "/" + k] = (b.channels,)
    assert len(name_map) > 50 and not any(k.split("/")[0].endswith(("FeatureExtractor", "Predictor")) for k in name_map)
    rng = np.random.default_rng(0)
    ckpt = {}
    for ck, targets in name_map.items():
        shp = shapes[targets[0]]
        assert all(shapes[t] == shp for t in targets), ck              # every copy has the checkpoint tensor's shape
        v = rng.standard_normal(shp).astype(np.float32)
        if ck.endswith("moving_variance"):
            v = np.abs(v) + 0.5
        ckpt[ck] = tf_checkpoint.native_to_tf(ck, v, kinds.get(targets[0]))
        if isinstance(kinds.get(targets[0]), tuple):                   # the RGB stem conv is a real [3,3,3,K] variable
            assert ckpt[ck].shape == (3, 3, 3, shp[0]), (ck, ckpt[ck].shape)
    shared = [ck for ck, t in name_map.items() if len(t) > 1]
    multi = max(len(name_map[ck]) for ck in shared) if shared else 1
    mtl = cfg.model.mtl
    # main tail + the aux scopes' own tails + (ResNet only: slim builds block4 inside the first-stage scope too, where it
    # is dead; the MobileNet / Inception-ResNet bases stop at the stride-16 endpoint)
    has_dead = any("/_dead/" in p.name for p in st.params)
    assert has_dead == ("resnet" in cfg.model.faster_rcnn.feature_extractor.type)
    copies = 1 + int(has_dead) + int(bool(mtl.closeness)) + int(bool(mtl.window))
    assert multi == copies, (multi, copies)
    prefix = str(tmp_path / "imagenet.ckpt")
    tf_checkpoint.write_checkpoint(prefix, ckpt)
    n, missing = checkpoint_io.load_tf_checkpoint(model, prefix, from_detection_checkpoint=False)
    assert not missing and n == sum(len(t) for t in name_map.values())
    sd = st.state_dict()
    for ck, targets in name_map.items():
        for t in targets:
            want = tf_checkpoint.tf_to_native(ck, ckpt[ck], kinds.get(t), shapes[t])
            np.testing.assert_array_equal(sd[t].numpy().reshape(want.shape), want, err_msg=t)
    heads = [p.name for p in st.params if "BoxPredictor/" in p.name and p.name not in
             {t for ts in name_map.values() for t in ts} and "/_pad/" not in p.name]
    assert heads                                                         # predictor layers are not in such a checkpoint


def test_detection_checkpoint_uses_the_reference_graph_names_for_inception_resnet(monkeypatch, tmp_path):
    """Trap T18: in the reference's detection graph the second-stage block8 stack of Inception-ResNet-v2 is named
    `<scope>/InceptionResnetV2/Repeat/...` (a fresh variable scope restarts slim.repeat's counter; incres fe:133-170,
    :173-248), in a classification checkpoint `InceptionResnetV2/Repeat_2/...`.  Detection checkpoints written / read here
    use the former, the ImageNet map the latter."""
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.utils import checkpoint_io, tf_checkpoint
    dryrun.install(monkeypatch)
    cfg = load_config("model62.config", SMALL[1:])
    a = model_builder.build(cfg.model, True, device="cpu", seed=0)
    prefix = str(tmp_path / "model.ckpt-1")
    names = checkpoint_io.save_tf_checkpoint(a, prefix)
    for scope in ("SecondStageFeatureExtractor", "ClosenessBoxPredictor", "WindowBoxPredictor"):
        assert scope + "/InceptionResnetV2/Repeat/block8_1/Branch_0/Conv2d_1x1/weights" in names
        assert not any(n.startswith(scope + "/InceptionResnetV2/Repeat_2/") for n in names)
    assert "FirstStageFeatureExtractor/InceptionResnetV2/Repeat_1/block17_1/Branch_0/Conv2d_1x1/weights" in names
    b = model_builder.build(cfg.model, True, device="cpu", seed=1)
    n, missing = checkpoint_io.load_tf_checkpoint(b, prefix)
    assert not missing
    sa, sb = a.param_store.state_dict(), b.param_store.state_dict()
    assert all(torch.equal(sa[k], sb[k]) for k in sa if "/_pad/" not in k)
    cls_map = checkpoint_io.variable_name_map(a, from_detection_checkpoint=False)
    assert len(cls_map["InceptionResnetV2/Repeat_2/block8_1/Branch_0/Conv2d_1x1/weights"]) == 3


def test_every_gpu_test_body_executes_in_dry_run_mode():
    """`python -O -m pytest tests -m gpu --assert=plain --dry-run-gpu`: the bodies of ALL -m gpu tests run on the CPU with
    stubbed kernels and stripped numerical asserts (tests/conftest.py).  What this catches before a GPU is spent on it:
    C-ABI calls with the wrong number / kind of arguments, missing prediction-dict keys, shape mismatches between host
    buffers, API drift between the tests and the product -- in particular for the device tests written after this
    round's GPU budget was gone (tests/test_gpu_zz_*.py)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-O", "-m", "pytest", os.path.join(root, "tests"), "-m", "gpu", "--assert=plain",
                        "--dry-run-gpu", "-q", "-x", "-p", "no:cacheprovider"], cwd=root, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=1500)
    tail = "\n".join(r.stdout.strip().splitlines()[-15:])
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail and "error" not in tail.lower(), tail
