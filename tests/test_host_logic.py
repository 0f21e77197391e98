"""CPU tests of the host-side mirror: config drop-in, variable table, C-ABI exports, schedules,
aux-label synthesis, and the data-parallel gradient exchange (gloo, world size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import CONFIG_DIR, load_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_fixture_configs_parse_unchanged():
    for name in sorted(os.listdir(CONFIG_DIR)):
        if not name.endswith(".config"):
            continue
        cfg = load_config(name)
        fr = cfg.model.faster_rcnn
        assert cfg.model.WhichOneof("model") == "faster_rcnn"
        assert fr.num_classes in (20, 90)
        assert fr.first_stage_max_proposals == 300 and fr.first_stage_minibatch_size == 256   # default
        assert abs(fr.first_stage_nms_iou_threshold - 0.7) < 1e-6
        assert cfg.train_config.gradient_clipping_by_norm == 10.0
        assert cfg.train_config.optimizer.WhichOneof("optimizer") == "momentum_optimizer"


def test_config_proto2_semantics():
    from mtl_ssl_b200.protos import text_format
    cfg = load_config("model12.config")
    m = cfg.model.mtl
    assert m.window and m.closeness and m.edgemask and m.refine and m.refine_residue
    assert m.shared_feature == "proposal_feature_maps" and m.global_closeness is True      # defaults
    assert abs(m.closeness_loss_weight - 0.3) < 1e-7
    assert cfg.model.faster_rcnn.feature_extractor.freeze_layer == "block1"                # default
    assert cfg.model.faster_rcnn.HasField("initial_crop_size")
    assert not cfg.model.faster_rcnn.HasField("hard_example_miner")
    g = cfg.model.faster_rcnn.first_stage_anchor_generator.grid_anchor_generator
    assert list(g.scales) == [0.25, 0.5, 1.0, 2.0] and g.height == 256 and g.height_stride == 16
    with pytest.raises(text_format.ParseError):
        text_format.Merge("model { no_such_field: 1 }", text_format.Message("TrainEvalPipelineConfig"))
    c22 = load_config("model22.config")
    assert c22.model.mtl.stop_gradient_for_aux_tasks and c22.train_config.batch_size == 3


def test_learning_rate_schedule_from_config():
    from mtl_ssl_b200.utils import learning_schedules as ls
    cfg = load_config("model12.config")
    fn, mom = ls.from_optimizer_config(cfg.train_config.optimizer)
    assert abs(mom - 0.9) < 1e-7
    assert [fn(s) for s in (0, 89999, 90000, 119999, 120000, 10 ** 6)] == \
        pytest.approx([1e-3, 1e-3, 1e-4, 1e-4, 1e-5, 1e-5])
    # utils/learning_schedules_test.py: boundaries [2,3,7], rates [1,2,3,4]
    got = [ls.manual_stepping(s, [2, 3, 7], [1.0, 2.0, 3.0, 4.0]) for s in range(10)]
    assert got == [1.0, 1.0, 2.0, 3.0, 3.0, 3.0, 3.0, 4.0, 4.0, 4.0]
    with pytest.raises(ValueError):
        ls.manual_stepping(0, [3, 2], [1.0, 2.0, 3.0])


def test_variable_table_matches_reference_scopes():
    """Variable names / shapes of the model built from model12.config (no device needed)."""
    from mtl_ssl_b200.builders import model_builder
    cfg = load_config("model12.config")
    model = model_builder.build(cfg.model, True, device=None)
    st = model.param_store
    names = {p.name: p for p in st.params}
    arch = "resnet_v1_101"
    assert names["FirstStageFeatureExtractor/%s/conv1/weights" % arch].shape == (64, 7, 7, 3)
    assert not names["FirstStageFeatureExtractor/%s/conv1/weights" % arch].trainable            # rv1:216-221
    assert not names["FirstStageFeatureExtractor/%s/block1/unit_1/bottleneck_v1/conv1/weights" % arch].trainable
    assert names["FirstStageFeatureExtractor/%s/block2/unit_1/bottleneck_v1/conv1/weights" % arch].trainable
    assert names["FirstStageFeatureExtractor/%s/block3/unit_23/bottleneck_v1/conv3/weights" % arch].shape == (1024, 1, 1, 256)
    for sc in ("SecondStageFeatureExtractor", "ClosenessBoxPredictor", "WindowBoxPredictor"):
        assert names["%s/%s/block4/unit_1/bottleneck_v1/shortcut/weights" % (sc, arch)].shape == (2048, 1, 1, 1024)
    assert names["FirstStageBoxPredictor/Conv/weights"].shape == (512, 3, 3, 1024)
    assert names["FirstStageBoxPredictor/BoxEncodingPredictor/weights"].shape == (48, 1, 1, 512)
    assert names["FirstStageBoxPredictor/ClassPredictor/weights"].shape == (24, 1, 1, 512)
    assert names["SecondStageBoxPredictor/BoxEncodingPredictor/weights"].shape == (80, 1, 1, 2048)
    assert names["SecondStageBoxPredictor/ClassPredictor/weights"].shape == (21, 1, 1, 2048)
    assert names["WindowBoxPredictor/ClassPredictor/weights"].shape == (21, 1, 1, 2048)          # K+1 (T11)
    assert names["ClosenessBoxPredictor/ClassPredictor/biases"].shape == (21,)
    assert names["EdgeMaskPredictor/BoxEncodingPredictor/weights"].shape == (2, 1, 1, 1024)
    assert names["MTLClassRefiner/fc1/weights"].shape == (21, 147)
    assert abs(names["WindowBoxPredictor/ClassPredictor/weights"].l2 - 1e-3) < 1e-9
    assert abs(names["SecondStageBoxPredictor/ClassPredictor/weights"].l2 - 1e-4) < 1e-9
    n_train = sum(p.numel for p in st.params if p.trainable and "/_dead/" not in p.name)
    assert 76e6 < n_train < 78.5e6, n_train                                                     # SURVEY: ~77.1 M
    rm = model.restore_map(from_detection_checkpoint=False)
    key = "%s/block4/unit_1/bottleneck_v1/conv1/weights" % arch
    assert len(rm[key]) == 4          # dead stage-1 copy + second stage + closeness + window (T5, T14)
    base = model_builder.build(load_config("model11.config").model, True, device=None)
    assert not any(n.startswith(("WindowBoxPredictor", "ClosenessBoxPredictor", "EdgeMaskPredictor", "MTLClassRefiner"))
                   for n in base.param_store.by_name)


def test_product_path_fails_loudly_without_cuda():
    from mtl_ssl_b200.builders import model_builder
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(Exception):
        model_builder.build(load_config("model12.config").model, True, device="cuda")


def test_cabi_library_exports_every_declared_symbol():
    from mtl_ssl_b200 import _lib, ops
    hdr = open(os.path.join(ROOT, "include", "mtlssl.h")).read()
    declared = set(re.findall(r"\b(mtl_[a-z0-9_]+)\s*\(", hdr)) - {"mtl_conv_args"}
    assert len(declared) >= 40
    assert os.path.exists(_lib.LIB_PATH), "run `python -m mtl_ssl_b200.build` first"
    h = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(h, name), name
    assert set(ops._SIGS) <= declared
    # ... and the binding generator parsed every declaration (a ';' inside a parameter comment would drop one)
    hand_bound = {"mtl_conv_tc", "mtl_conv_tc_ws_bytes", "mtl_last_error_string", "mtl_abi_version",
                  "mtl_device_sm_count", "mtl_opt_chunk_size", "mtl_launch_count", "mtl_set_error", "mtl_count_launch",
                  "mtl_conv_tc_group_entry_bytes", "mtl_conv_tc_group_key", "mtl_conv_tc_group_build",
                  "mtl_conv_tc_group_launch"}
    assert declared - set(ops._SIGS) <= hand_bound, sorted(declared - set(ops._SIGS) - hand_bound)
    assert h.mtl_abi_version() == 1
    # no torch / C++ types cross the boundary: undefined symbols must not reference at:: / c10::
    out = subprocess.run(["nm", "-D", "--undefined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "c10" not in out and "at6" not in out and "torch" not in out.lower()


def test_aux_label_synthesis():
    from mtl_ssl_b200.data import aux_labels as AL
    # two unit-overlap boxes: union by inclusion-exclusion = 2*4 - 1
    assert AL.union_area([[0, 0, 2, 2], [1, 1, 3, 3]]) == pytest.approx(7.0)
    boxes = np.array([[100., 100, 300, 300], [200, 200, 500, 600]])
    classes = np.array([3, 7])
    lab, bg = AL.window_label(boxes, classes, [0, 0, 600, 1000], 20)
    assert lab.shape == (21,) and abs(lab.sum() - 1.0) < 5e-3 and lab[3] > 0 and lab[7] > lab[3] and lab[1] == 0
    a3 = 200 * 200 / 6e5; a7 = 300 * 400 / 6e5; abg = 1 - (a3 + a7 - 100 * 100 / 6e5)
    raw = np.array([np.sqrt(abg), np.sqrt(a3), np.sqrt(a7)])
    np.testing.assert_allclose([lab[0], lab[3], lab[7]], np.round(raw / raw.sum(), 3), atol=1e-6)
    cl = AL.closeness_labels(boxes, classes, 600, 1000, 20)
    assert cl.shape == (2, 21) and cl[0, 7] == 1.0 and cl[1, 3] == 1.0
    assert AL.closeness_labels(boxes[:1], classes[:1], 600, 1000, 20)[0, 0] == 1.0
    em = AL.edgemask(boxes, 600., 1000.)
    assert em.shape == (2, 64, 64) and em[0].max() == 1 and abs(em[1].mean() - 1) < 1e-5
    rng = np.random.default_rng(0)
    wb, wl = AL.random_windows(boxes, classes, 600., 1000., 20, rng, 64)
    assert wb.shape == (64, 4) and wl.shape == (64, 21) and (wl[:, 0] < 1).all() and (wb <= 1).all()


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from mtl_ssl_b200.parallel import allreduce_gradients, data_parallel_scale
dist.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=2)
r = dist.get_rank()
g = torch.arange(10, dtype=torch.float32) * (r + 1)
allreduce_gradients(g, 2)
exp = torch.arange(10, dtype=torch.float32) * 3
assert torch.equal(g, exp), g
# reference semantics (model_deploy.py:221-225, :296): task-loss gradients averaged, L2 counted once
assert data_parallel_scale(2) == 0.5
w = torch.ones(10); l2 = 0.1
total = g * data_parallel_scale(2) + l2 * w
assert torch.allclose(total, exp * 0.5 + 0.1)
dist.destroy_process_group()
print("rank", r, "ok")
"""


def test_data_parallel_gradient_exchange_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_trainer_gradient_knobs_become_per_tensor_table_entries():
    """trainer.py:387-410 of the reference: grad_multiplier / divide_grad_by_batch / bias_grad_multiplier /
    freeze_variables.  No shipped config sets them; a config that does gets them as per-tensor entries of the optimizer
    table (multiplier on the total gradient before the clip, frozen = no update), with the reference's matching rule
    (`re.match` of each regular expression on the variable name, variables_helper.py:27-55).  The arithmetic on the
    device is checked by tests/test_gpu_roi_loss_kernels.py::test_gradient_multipliers_and_frozen_variables."""
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.trainer import Trainer

    def build(*lines):
        cfg = load_config("model12.config", (("gradient_clipping_by_norm: 10.0",
                                              "gradient_clipping_by_norm: 10.0\n  " + "\n  ".join(lines)),))
        model = model_builder.build(cfg.model, True, device=None)
        Trainer(model, cfg.train_config, 600, 1000, 1)
        return model.param_store

    st = build("bias_grad_multiplier: 2.0")
    assert all(p.grad_mult == (2.0 if p.name.endswith("/biases") else 1.0) for p in st.params)
    assert any(p.name.endswith("/biases") for p in st.params)
    st = build("grad_multiplier: 0.5", "divide_grad_by_batch: true")          # batch_size 1 in model12.config
    assert all(p.grad_mult == 0.5 for p in st.params)
    st = build("freeze_variables: '.*block2.*'", "freeze_variables: 'FirstStageBoxPredictor/Conv'")
    frozen = [p.name for p in st.params if not p.trainable]
    assert any("/block2/" in n for n in frozen) and "FirstStageBoxPredictor/Conv/weights" in frozen
    base = model_builder.build(load_config("model12.config").model, True, device=None).param_store
    newly = set(frozen) - {p.name for p in base.params if not p.trainable}
    assert newly and all("/block2/" in n or n.startswith("FirstStageBoxPredictor/Conv") for n in newly)
    assert "FirstStageBoxPredictor/ClassPredictor/weights" not in newly       # re.match anchors at the start


def test_resize_to_range_output_sizes_match_reference_vectors():
    """core/preprocessor_test.py:1398-1427 (`testResizeToRangePreservesStaticSpatialShape`) and :1514-1526
    (`testResizeToRangeSameMinMax`): the output size rule of `resize_to_range` (host arithmetic in front of the
    bilinear resize kernel)."""
    from mtl_ssl_b200.core.preprocessor import _compute_new_static_size
    for (h, w), want in zip(([60, 40], [15, 30], [15, 50]), ([75, 50], [50, 100], [30, 100])):
        assert list(_compute_new_static_size(h, w, 50, 100)) == want
    for (h, w) in ([312, 312], [299, 299]):
        assert list(_compute_new_static_size(h, w, 320, 320)) == [320, 320]
    # the two workloads of BASELINE.json: VOC-shape and COCO-shape inputs are already in range
    assert list(_compute_new_static_size(600, 1000, 600, 1024)) == [600, 1000]
    assert list(_compute_new_static_size(300, 300, 600, 1024)) == [600, 600]          # model51: 300x300 upsampled


def test_label_map_util_reference_vectors(tmp_path):
    """utils/label_map_util_test.py:37-167: name -> id dictionary, ids below 1 refused, first item per id kept,
    display names, default categories, category index; plus the two label maps shipped with the reference's data/."""
    from mtl_ssl_b200.utils import label_map_util as L
    path = str(tmp_path / "label_map.pbtxt")
    open(path, "w").write("item {\n  id:2\n  name:'cat'\n}\nitem {\n  id:1\n  name:'dog'\n}\n")
    d = L.get_label_map_dict(path)
    assert d == {"dog": 1, "cat": 2} and L.get_class_indices(d) == [1, 2] and L.get_index_map_dict(d) == {1: "dog", 2: "cat", 0: "bg"}
    open(path, "w").write("item {\n id:0\n name:'class that should not be indexed at zero'\n}\nitem {\n id:2\n name:'cat'\n}\n")
    with pytest.raises(ValueError):
        L.load_labelmap(path)
    lm = L.load_labelmap("item {\n id:2\n name:'cat'\n}\nitem {\n id:1\n name:'child'\n}\nitem {\n id:1\n name:'person'\n}\n"
                         "item {\n id:1\n name:'n00007846'\n}\n")
    assert L.convert_label_map_to_categories(lm, max_num_classes=3) == [{"id": 2, "name": "cat"}, {"id": 1, "name": "child"}]
    assert L.convert_label_map_to_categories(None, max_num_classes=3) == \
        [{"name": "category_1", "id": 1}, {"name": "category_2", "id": 2}, {"name": "category_3", "id": 3}]
    gen = "".join("item {\n id: %d\n name: 'label_%d'\n display_name: '%d'\n}\n" % (i, i, i) for i in range(1, 5))
    assert L.convert_label_map_to_categories(L.load_labelmap(gen), max_num_classes=3) == \
        [{"name": "1", "id": 1}, {"name": "2", "id": 2}, {"name": "3", "id": 3}]
    assert L.convert_label_map_to_categories(L.load_labelmap(gen), 2, use_display_name=False) == \
        [{"name": "label_1", "id": 1}, {"name": "label_2", "id": 2}]
    assert L.create_category_index([{"name": "1", "id": 1}, {"name": "2", "id": 2}]) == \
        {1: {"name": "1", "id": 1}, 2: {"name": "2", "id": 2}}
    # the shape of the maps under the reference's data/ directory (pascal: 20 dense ids; mscoco: 80 names on ids 1..90)
    voc = "".join("item {\n  id: %d\n  name: 'c%d'\n}\n\n" % (i, i) for i in range(1, 21))
    assert len(L.convert_label_map_to_categories(L.load_labelmap(voc), 20)) == 20
    coco = "".join('item {\n  name: "/m/%02d"\n  id: %d\n  display_name: "thing %d"\n}\n' % (i, i, i)
                   for i in list(range(1, 12)) + [13, 90])
    cats = L.convert_label_map_to_categories(L.load_labelmap(coco), 90)
    assert cats[-1] == {"id": 90, "name": "thing 90"} and len(cats) == 13


def test_bench_dominant_conv_group():
    """bench.py's breakdown of the live per-launch timings: the (mode, geometry) group with the largest summed
    duration, its algorithmic rate and its fraction of the peak."""
    import bench
    geom_a, geom_b = (1280, 7, 7, 512, 512, 3, 1, 7, 7), (1, 38, 63, 1024, 256, 1, 1, 38, 63)
    recs = [(0, 2e12, 1.0, geom_a)] * 3 + [(0, 1e11, 0.5, geom_b)] * 4 + [(2, 2e12, 0.9, geom_a)] * 3
    d = bench.dominant_conv_group(recs, 1400.0)
    assert d["launches"] == 3 and d["kernel"].startswith("tc_gemm_kernel fprop 3x3/1 C512->K512 on 1280x7x7")
    assert abs(d["achieved"] - 2000.0) < 1e-9 and abs(d["frac"] - 2000.0 / 1400.0) < 1e-12
    assert abs(d["avg_us"] - 1000.0) < 1e-9 and abs(d["share_of_conv_time"] - 3.0 / 7.7) < 1e-12
    assert bench.dominant_conv_group([], 1400.0) is None


def test_builder_refuses_anchor_options_without_a_kernel():
    """grid_anchor_generator.proto fields 9-12 (the fork's `use_hw_scales` / `align_lefttop`): no shipped config sets
    them; the builder refuses them instead of generating the default anchors."""
    from mtl_ssl_b200.builders import model_builder
    for line in ("use_hw_scales: true", "align_lefttop: true"):
        cfg = load_config("model12.config", (("grid_anchor_generator {", "grid_anchor_generator {\n        " + line),))
        with pytest.raises(ValueError, match="use_hw_scales"):
            model_builder.build(cfg.model, True, device=None)


def test_every_fixture_config_builds_in_training_and_inference_mode():
    """builders/model_builder.py:68-380 for every config under tests/golden/configs (and, in the build container, all 18
    of the reference's configs/test/*.config): the host-side model description builds without a device in both modes;
    max_num_proposals follows fmA:475-477 (second_stage_batch_size when training, first_stage_max_proposals otherwise)."""
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.protos import text_format
    paths = [os.path.join(CONFIG_DIR, n) for n in sorted(os.listdir(CONFIG_DIR)) if n.endswith(".config")]
    ref_dir = "/root/reference/object_detection/configs/test"
    if os.path.isdir(ref_dir):
        paths += [os.path.join(ref_dir, n) for n in sorted(os.listdir(ref_dir)) if n.endswith(".config")]
        assert len(paths) >= 18
    for path in paths:
        cfg = text_format.load_pipeline_config(path)
        fr = cfg.model.faster_rcnn
        for training in (True, False):
            m = model_builder.build(cfg.model, training, device=None)
            assert m.max_num_proposals == (fr.second_stage_batch_size if training else fr.first_stage_max_proposals)
            assert m.num_classes == fr.num_classes and len(m.param_store.params) > 30


def test_shape_bucket_trainer_bookkeeping():
    """shape_buckets.ShapeBucketTrainer with stand-in trainers (the device side is tests/test_gpu_zz_shape_buckets.py):
    one trainer + workspace per image shape installed in the model before every call, a shared global_step, the losses
    of the previous call handed back in call order across a shape switch, least-recently-used eviction."""
    import types
    from mtl_ssl_b200.shape_buckets import ShapeBucketTrainer
    log = []

    class FakeWs(object):
        def __init__(self, device):
            self.device = device

    class FakeTrainer(object):
        def __init__(self, model, train_config, H, W, B, gmax=64):
            self.model, self.hw, self.global_step, self._pending, self.ws_at_build = model, (H, W), 0, None, model._ws

        def host_arrays(self, examples, keys):
            return {"image": np.stack([e["image"] for e in examples]), "keys1": keys[0], "keys2": keys[1]}

        def step_pipelined(self, arrays):
            assert self.model._ws is self.ws_at_build                     # its own workspace is installed
            log.append((self.hw, self.global_step))
            prev, self._pending = self._pending, {"total_loss": float(self.global_step)}
            self.global_step += 1
            return prev

        def flush(self):
            prev, self._pending = self._pending, None
            return prev

    model = types.SimpleNamespace(device="cpu", _ws=None, num_classes=3)
    bt = ShapeBucketTrainer(model, None, batch_size=1, max_buckets=2, trainer_cls=FakeTrainer, workspace_cls=FakeWs)
    def batch(h, w):
        ex = [dict(image=np.zeros((h, w, 3), np.float32), groundtruth_boxes=np.zeros((0, 4), np.float32),
                   groundtruth_classes=np.zeros((0, 3), np.float32))]
        return bt.host_arrays(ex, (np.zeros((1, 5), np.float32), np.zeros((1, 4), np.float32)))
    a = batch(32, 48)
    assert a.hw == (32, 48) and a["image"].shape == (1, 32, 48, 3) and a["gt"].shape == (1, 64, 4)
    seq = [(32, 48), (32, 48), (40, 40), (32, 48), (24, 64), (40, 40)]
    got = [bt.step_pipelined(batch(h, w)) for h, w in seq] + [bt.flush()]
    assert got[0] is None and [r["total_loss"] for r in got[1:]] == [0.0, 1.0, 2.0, 3.0, 4.0, 5.0]      # call order kept
    assert log == [(hw, i) for i, hw in enumerate(seq)] and bt.global_step == 6
    # (40,40) was least recently used when (24,64) arrived: evicted, rebuilt on its next use -> (32,48) evicted then
    assert bt.evictions == 2 and list(bt.buckets) == [(24, 64), (40, 40)]
    assert bt.flush() is None
    with pytest.raises(ValueError, match="different sizes"):
        bt.host_arrays([dict(image=np.zeros((8, 8, 3))), dict(image=np.zeros((8, 9, 3)))], (None, None))
    # a single bucket: every shape switch rebuilds it, the losses still come back in call order
    del log[:]
    one = ShapeBucketTrainer(model, None, batch_size=1, max_buckets=1, trainer_cls=FakeTrainer, workspace_cls=FakeWs)
    bt = one
    got = [one.step_pipelined(batch(h, w)) for h, w in [(32, 48), (40, 40), (32, 48)]] + [one.flush()]
    assert got[0] is None and [r["total_loss"] for r in got[1:]] == [0.0, 1.0, 2.0] and one.evictions == 2


def test_pack_groundtruth_accepts_images_without_boxes():
    """COCO holds images without annotations: packing must give num_gt = 0, not fail."""
    from mtl_ssl_b200.trainer import pack_groundtruth
    ex = [dict(groundtruth_boxes=np.zeros((0, 4), np.float32), groundtruth_classes=np.zeros((0, 3), np.float32)),
          dict(groundtruth_boxes=np.array([[0.1, 0.2, 0.5, 0.6]], np.float32),
               groundtruth_classes=np.array([[0, 1, 0]], np.float32))]
    out = pack_groundtruth(ex, 3, 100, 200, 4)
    assert out["num_gt"].tolist() == [0, 1] and out["gt_cls"][1, 0] == 2 and not out["gt"][0].any()
    np.testing.assert_allclose(out["gt"][1, 0], [10, 40, 50, 120], rtol=1e-6)


def test_bench_dominant_group_carries_the_committed_dram_traffic():
    """roofline.dominant.traffic: dram bytes read + written per launch from the committed ncu capture of that kernel."""
    import bench
    geom = (1280, 7, 7, 512, 512, 3, 1, 7, 7)
    d = bench.dominant_conv_group([(0, 2e12, 1.0, geom)] * 3 + [(0, 1e9, 0.1, (1, 38, 63, 64, 64, 1, 1, 38, 63))], 1400.0)
    assert d["traffic"] == 69021440 + 24559872 and "f_c3x3" in d["traffic_source"]
    d = bench.dominant_conv_group([(1, 2e12, 1.0, (7, 9, 9, 64, 64, 3, 1, 9, 9))], 1400.0)
    assert d["traffic"] is None and d["kernel"].startswith("tc_gemm_kernel dgrad")


def test_initializer_parameters_reach_the_variables():
    """hyperparams_builder._build_initializer (hyperparams_builder.py:118-146) passes factor / mode /
    uniform of variance_scaling_initializer and the means of the normal initializers on, and refuses a missing
    initializer; model22.config sets factor 1.0, uniform, FAN_AVG for the second-stage FC heads."""
    import math
    import torch
    from helpers import load_config
    from mtl_ssl_b200.builders import model_builder
    from mtl_ssl_b200.core.hyperparams import Hyperparams
    from mtl_ssl_b200.protos import text_format
    from mtl_ssl_b200.runtime import Param, _init_tensor
    cfg = load_config("model22.config")
    hp = cfg.model.faster_rcnn.second_stage_box_predictor.mask_rcnn_box_predictor.fc_hyperparams
    assert Hyperparams.from_proto(hp).init == ("variance_scaling", 1.0, "FAN_AVG", True)
    model = model_builder.build(cfg.model, True, device=None)
    p = model.param_store.by_name["SecondStageBoxPredictor/ClassPredictor/weights"]
    assert p.init == ("variance_scaling", 1.0, "FAN_AVG", True) and p.shape == (91, 1, 1, 2048)
    w = _init_tensor(p, torch.Generator().manual_seed(0))
    lim = math.sqrt(3.0 * 1.0 / ((2048 + 91) / 2.0))                     # uniform: sqrt(3 * factor / n)
    assert float(w.abs().max()) <= lim and float(w.abs().max()) > 0.98 * lim
    assert abs(float(w.var()) - lim * lim / 3.0) < 0.03 * lim * lim / 3.0
    # defaults of the proto (factor 2.0, FAN_IN, truncated normal with stddev sqrt(1.3 * factor / fan_in))
    q = Param("c/weights", (64, 3, 3, 32), 0.0, True, ("variance_scaling", 2.0, "FAN_IN", False))
    v = _init_tensor(q, torch.Generator().manual_seed(1))
    std = math.sqrt(1.3 * 2.0 / (3 * 3 * 32))
    assert float(v.abs().max()) <= 2 * std + 1e-6 and abs(float(v.std()) - 0.88 * std) < 0.03 * std
    assert torch.equal(v, _init_tensor(Param("c/weights", (64, 3, 3, 32), 0.0, True, ("variance_scaling",)),
                                       torch.Generator().manual_seed(1)))
    q = Param("c/weights", (64, 3, 3, 32), 0.0, True, ("variance_scaling", 1.0, "FAN_OUT", True))
    assert float(_init_tensor(q, torch.Generator().manual_seed(1)).abs().max()) <= math.sqrt(3.0 / (3 * 3 * 64))
    t = _init_tensor(Param("b", (4000,), 0.0, True, ("truncated_normal", 0.1, 0.5)), torch.Generator().manual_seed(2))
    assert abs(float(t.mean()) - 0.5) < 0.01 and float((t - 0.5).abs().max()) <= 0.2 + 1e-6
    t = _init_tensor(Param("b", (4000,), 0.0, True, ("normal", 0.1, -1.0)), torch.Generator().manual_seed(2))
    assert abs(float(t.mean()) + 1.0) < 0.01
    msg = text_format.Merge("op: FC regularizer { l2_regularizer { weight: 0.1 } }", text_format.Message("Hyperparams"))
    with pytest.raises(ValueError):
        Hyperparams.from_proto(msg)


def test_wgrad_collector_flushes_by_output_range(monkeypatch):
    """ops_conv.WgradCollector.flush(out_range=...): only the deferred GEMMs whose output lies in the given piece of the
    gradient arena run, the others stay pending, each flush point plans and caches its own groups (the piecewise exchange
    of the second-stage bucket with several replicas, trainer._run_step_deferred)."""
    from mtl_ssl_b200 import ops_conv as oc
    launched = []

    class FakeGroup(object):
        def __init__(self, args, target):
            self.outs = [a.out for a in args]

        def launch(self, max_ctas=0):
            launched.append((tuple(self.outs), max_ctas))

    monkeypatch.setattr(oc, "ConvGroup", FakeGroup)
    monkeypatch.setattr(oc, "group_key", lambda a: a.C)

    def arg(out, C):
        a = oc.ConvArgs()
        a.mode, a.out, a.C, a.N, a.H, a.W, a.K, a.R, a.stride = oc.WGRAD, out, C, 1, 7, 7, 8, 1, 1
        return a

    col = oc.WgradCollector()
    for rep in range(2):
        for out, C in ((1000, 64), (5000, 64), (2000, 128), (9000, 64)):
            col.add(arg(out, C))
        del launched[:]
        col.flush(max_ctas=48, key=("piece", 0), out_range=(0, 4000))
        assert sorted(o for g, _ in launched for o in g) == [1000, 2000] and len(launched) == 2     # two kernel instances
        assert all(c == 48 for _, c in launched) and [a.out for a in col.pending] == [5000, 9000]
        col.flush(max_ctas=48, key=("piece", 1), out_range=(4000, 6000))
        assert [a.out for a in col.pending] == [9000]
        col.flush(max_ctas=0, key=("piece", "rest"))
        assert not col.pending and launched[-1] == ((9000,), 0)
    assert len(col.groups) == 3                  # planned once per flush point, reused on the second pass
