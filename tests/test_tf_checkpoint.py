"""CPU tests of the TensorFlow V2 checkpoint ("tensor bundle") reader.  No TensorFlow and no checkpoint
file exist here, so the format is PARITY UNPINNED: the reader is exercised against the writer of the same module
(LevelDB table layout, prefix compression across restart points, multi-block index, CRCs) and its protobuf pieces
against the protobuf runtime elsewhere (tests/test_tfrecord.py)."""
import os
import struct

import numpy as np
import pytest

from mtl_ssl_b200.utils import tf_checkpoint as C


def _tensors(rng, n):
    out = {}
    for i in range(n):
        scope = "FirstStageFeatureExtractor/resnet_v1_101/block%d/unit_%d/bottleneck_v1/conv%d" % (i % 4 + 1, i // 12 + 1,
                                                                                              i % 3 + 1)
        out[scope + "/weights"] = rng.normal(size=(1, 1, 4 + i % 5, 8)).astype(np.float32)
        out[scope + "/BatchNorm/moving_mean"] = rng.normal(size=(8,)).astype(np.float32)
    out["global_step"] = np.asarray(12345, np.int64)
    out["some/int32"] = rng.integers(-5, 5, (3, 2)).astype(np.int32)
    out["some/half"] = rng.normal(size=(4,)).astype(np.float16)
    out["empty"] = np.zeros((0, 3), np.float32)
    return out


def test_bundle_round_trip_multi_block(tmp_path):
    rng = np.random.default_rng(0)
    tensors = _tensors(rng, 150)
    prefix = str(tmp_path / "model.ckpt-12345")
    C.write_checkpoint(prefix, tensors)
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == 0xDB4775248B80FB57 and len(raw) > 3 * 4096      # several data blocks
    assert len(raw) < sum(len(k) + 40 for k in tensors)                                           # keys are prefix-compressed
    r = C.CheckpointReader(prefix)
    assert r.num_shards == 1 and set(r.entries) == set(tensors)
    shapes = r.get_variable_to_shape_map()
    for k, v in tensors.items():
        assert shapes[k] == list(v.shape)
        got = r.get_tensor(k)
        assert got.dtype == v.dtype and np.array_equal(got, v), k
    keys = [k for k, _ in C.read_table(prefix + ".index")]
    assert keys == sorted(keys) and keys[0] == b""
    assert not r.has_tensor("nope")


def test_bundle_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(1)
    prefix = str(tmp_path / "m.ckpt")
    C.write_checkpoint(prefix, _tensors(rng, 6))
    data = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    data[10] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    r = C.CheckpointReader(prefix)
    bad = [k for k in r.entries if r.entries[k]["offset"] <= 10 < r.entries[k]["offset"] + r.entries[k]["size"]]
    assert len(bad) == 1
    with pytest.raises(ValueError):
        r.get_tensor(bad[0])
    assert C.CheckpointReader(prefix, check_crc=False).get_tensor(bad[0]) is not None
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[5] ^= 1
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError):
        C.CheckpointReader(prefix)
    open(prefix + ".index", "wb").write(b"not a table")
    with pytest.raises(ValueError):
        C.CheckpointReader(prefix)
    with pytest.raises(ValueError):
        C.write_table(str(tmp_path / "t"), [(b"b", b"1"), (b"a", b"2")])


def test_layout_conversion_and_name_map(tmp_path):
    rng = np.random.default_rng(2)
    hwio = rng.normal(size=(3, 3, 16, 32)).astype(np.float32)
    native = C.tf_to_native("x/conv2/weights", hwio)
    assert native.shape == (32, 3, 3, 16) and native[5, 1, 2, 7] == hwio[1, 2, 7, 5]
    assert np.array_equal(C.native_to_tf("x/conv2/weights", native), hwio)
    dw = rng.normal(size=(3, 3, 24, 1)).astype(np.float32)
    assert C.tf_to_native("m/Conv2d_1_depthwise/depthwise_weights", dw).shape == (24, 3, 3, 1)
    assert np.array_equal(C.native_to_tf("m/Conv2d_1_depthwise/depthwise_weights",
                                         C.tf_to_native("m/Conv2d_1_depthwise/depthwise_weights", dw)), dw)
    fc = rng.normal(size=(2048, 21)).astype(np.float32)
    assert C.tf_to_native("SecondStageBoxPredictor/ClassPredictor/weights", fc).shape == (21, 2048)
    beta = rng.normal(size=(64,)).astype(np.float32)
    assert C.tf_to_native("a/BatchNorm/beta", beta) is beta or np.array_equal(C.tf_to_native("a/BatchNorm/beta", beta), beta)
    # ImageNet-style checkpoint: one resnet block4 initialises three scopes (trap T14, trainer.py:341-348)
    prefix = str(tmp_path / "resnet_v1_101.ckpt")
    C.write_checkpoint(prefix, {"resnet_v1_101/block4/unit_1/bottleneck_v1/conv2/weights": hwio,
                                "resnet_v1_101/block4/unit_1/bottleneck_v1/conv2/BatchNorm/beta": beta[:32]})
    name_map = {
        "resnet_v1_101/block4/unit_1/bottleneck_v1/conv2/weights": [
            s + "/resnet_v1_101/block4/unit_1/bottleneck_v1/conv2/weights"
            for s in ("SecondStageFeatureExtractor", "ClosenessBoxPredictor", "WindowBoxPredictor")],
        "resnet_v1_101/block4/unit_1/bottleneck_v1/conv2/BatchNorm/beta":
            "SecondStageFeatureExtractor/resnet_v1_101/block4/unit_1/bottleneck_v1/conv2/BatchNorm/beta",
        "resnet_v1_101/logits/weights": "unused/logits",
    }
    sd, missing = C.state_dict_from_checkpoint(C.CheckpointReader(prefix), name_map)
    assert missing == ["resnet_v1_101/logits/weights"] and len(sd) == 4
    for k, v in sd.items():
        if k.endswith("weights"):
            assert v.shape == (32, 3, 3, 16) and np.array_equal(v, native)
    with pytest.raises(ValueError):
        C.state_dict_from_checkpoint(C.CheckpointReader(prefix), name_map,
                                     shapes={"WindowBoxPredictor/resnet_v1_101/block4/unit_1/bottleneck_v1/conv2/weights":
                                             (32, 3, 3, 8)})


def test_training_state_round_trip_with_momentum_slots_and_global_step(tmp_path):
    """save_training_checkpoint / restore_training_checkpoint: variables + `<variable>/Momentum` slots (TF layouts)
    + global_step through the tensor-bundle format, on a stand-in parameter store (CPU tensors; the device store has the
    same attributes and is covered by the -m gpu round trip of the variables)."""
    import types
    import torch
    from mtl_ssl_b200.utils import checkpoint_io as C, tf_checkpoint as T

    def make(seed):
        g = torch.Generator().manual_seed(seed)
        specs = [("Scope/conv/weights", (8, 3, 3, 4), True), ("Scope/conv/biases", (8,), True),
                 ("Scope/fc/weights", (5, 16), True), ("Scope/frozen/weights", (4, 1, 1, 4), False),
                 ("Scope/_pad/x", (3,), True)]
        params = [types.SimpleNamespace(name=n, shape=s, trainable=t, w=torch.randn(s, generator=g),
                                        m=torch.randn(s, generator=g)) for n, s, t in specs]
        st = types.SimpleNamespace(params=params, bns=[])
        st.state_dict = lambda: {p.name: p.w.clone() for p in params}

        def load(sd, strict=True):
            for p in params:
                if p.name in sd:
                    p.w.copy_(sd[p.name].reshape(p.shape))
        st.load_state_dict = load
        model = types.SimpleNamespace(param_store=st)
        return types.SimpleNamespace(model=model, global_step=0), params

    tr_a, pa = make(1)
    tr_a.global_step = 12345
    names = C.save_training_checkpoint(tr_a, str(tmp_path / "model.ckpt-12345"))
    assert "global_step" in names and "Scope/conv/weights/Momentum" in names and "Scope/fc/weights/Momentum" in names
    assert "Scope/frozen/weights/Momentum" not in names and not any("_pad" in n for n in names)
    r = T.CheckpointReader(str(tmp_path / "model.ckpt-12345"))
    assert r.get_variable_to_shape_map()["Scope/conv/weights/Momentum"] == [3, 3, 4, 8]        # HWIO like the variable
    assert r.get_variable_to_shape_map()["Scope/fc/weights/Momentum"] == [16, 5] and int(r.get_tensor("global_step")) == 12345
    tr_b, pb = make(2)
    n, slots, step = C.restore_training_checkpoint(tr_b, str(tmp_path / "model.ckpt-12345"))
    assert (n, slots, step) == (4, 3, 12345) and tr_b.global_step == 12345
    for a, b in zip(pa, pb):
        if "_pad" in a.name:
            continue
        assert torch.equal(a.w, b.w)
        assert torch.equal(a.m, b.m) == a.trainable
    # a variables-only checkpoint leaves the momenta and the step alone
    C.save_tf_checkpoint(tr_a.model, str(tmp_path / "weights_only"))
    tr_c, pc = make(3)
    before = [p.m.clone() for p in pc]
    assert C.restore_training_checkpoint(tr_c, str(tmp_path / "weights_only"))[1:] == (0, 0)
    assert all(torch.equal(x, p.m) for x, p in zip(before, pc))


def test_train_loop_writes_and_resumes_checkpoints(tmp_path):
    """data/loader.train_loop with `checkpoint_prefix`: `<prefix>-<global_step>` every `save_every` steps and at the
    end (the Saver's role in slim.learning.train), `latest_checkpoint` picks the newest, a second trainer resumes from it."""
    import types
    import torch
    from mtl_ssl_b200.data import loader
    from mtl_ssl_b200.utils import checkpoint_io as C

    class FakeTrainer(object):
        def __init__(self):
            self.w, self.m = torch.zeros(4), torch.zeros(4)
            p = types.SimpleNamespace(name="S/fc/biases", shape=(4,), trainable=True, w=self.w, m=self.m)
            st = types.SimpleNamespace(params=[p], bns=[])
            st.state_dict = lambda: {"S/fc/biases": self.w.clone()}
            st.load_state_dict = lambda sd, strict=True: self.w.copy_(sd["S/fc/biases"])
            self.model = types.SimpleNamespace(param_store=st)
            self.global_step, self._pending = 0, None

        def step_pipelined(self, arrays):            # losses come back one call later, like Trainer.step_pipelined
            prev, self._pending = self._pending, {"total_loss": float(arrays)}
            self.global_step += 1
            self.m.mul_(0.9).add_(float(arrays))
            self.w.sub_(0.1 * self.m)
            return prev

        def flush(self):
            prev, self._pending = self._pending, None
            return prev

    tr = FakeTrainer()
    prefix = str(tmp_path / "model.ckpt")
    out = loader.train_loop(tr, iter([1.0, 2.0, 3.0, 4.0, 5.0]), checkpoint_prefix=prefix, save_every=2)
    assert [r["total_loss"] for r in out] == [1.0, 2.0, 3.0, 4.0, 5.0]
    steps = sorted(int(f.split("-")[-1][:-len(".index")]) for f in os.listdir(str(tmp_path)) if f.endswith(".index"))
    assert steps == [2, 4, 5] and C.latest_checkpoint(str(tmp_path)) == prefix + "-5"
    assert C.latest_checkpoint(str(tmp_path), "other") is None
    tr2 = FakeTrainer()
    assert C.restore_training_checkpoint(tr2, C.latest_checkpoint(str(tmp_path))) == (1, 1, 5)
    assert torch.equal(tr2.w, tr.w) and torch.equal(tr2.m, tr.m)
    # resumed and uninterrupted runs end in the same state
    loader.train_loop(tr2, iter([6.0, 7.0]))
    loader.train_loop(tr, iter([6.0, 7.0]))
    assert torch.equal(tr2.w, tr.w) and tr2.global_step == tr.global_step == 7


def test_train_loop_numerics_check_and_jsonl_log(tmp_path):
    """train_loop's failure detection (`tf.check_numerics` on the losses, trainer.py:207-209) and per-step JSONL log
    (loss keys + sec/batch + instances/sec, learning.py:509-519)."""
    import json
    from mtl_ssl_b200.data import loader

    class T(object):
        B, world_size, global_step = 2, 4, 0

        def __init__(self):
            self._p = None

        def step_pipelined(self, a):
            prev, self._p = self._p, {"first_stage_objectness_loss": a, "total_loss": a + 1.0}
            return prev

        def flush(self):
            prev, self._p = self._p, None
            return prev

    path = str(tmp_path / "train.jsonl")
    out = loader.train_loop(T(), iter([0.5, 0.25, 0.125]), jsonl_path=path)
    rows = [json.loads(l) for l in open(path)]
    assert [r["step"] for r in rows] == [1, 2, 3] and [r["total_loss"] for r in rows] == [1.5, 1.25, 1.125] and len(out) == 3
    assert all(r["sec/batch"] >= 0 and (r["instances/sec"] is None or r["instances/sec"] > 0) for r in rows)
    with pytest.raises(FloatingPointError, match="LossTensor is inf or nan"):
        loader.train_loop(T(), iter([0.5, float("nan"), 0.1]))
    assert len(loader.train_loop(T(), iter([0.5, float("inf")]), check_numerics=False)) == 2



@pytest.mark.parametrize("name", ["model12.config", "model42.config", "model52.config", "model62.config"])
def test_exported_variables_have_the_reference_graph_shapes(name):
    """the conversion must follow the variable's TF shape, not the rank of the GEMM operand kept here.
    Real variable tables of the shipped configs (host side, no device): every exported array has the shape the
    reference's graph gives that variable --
      slim.fully_connected heads of MaskRCNNBoxPredictor (bp:482-496, :568-602)   [in, out]
      the RGB stem convs (mobilenet_v1.py Conv2d_0, inception_resnet_v2.py Conv2d_1a_3x3)   [3, 3, 3, 32]
      every other conv [R, S, C, K], depthwise [3, 3, C, 1], the refiner FC [nf, K+1], vectors unchanged --
    and converting back reproduces the native tensor bit for bit (pad columns of the packed stems stay zero)."""
    from helpers import load_config
    from mtl_ssl_b200.builders import model_builder
    cfg = load_config(name)
    model = model_builder.build(cfg.model, True, device=None, seed=0)
    st = model.param_store
    sd = st.host_state_dict(0)
    K = cfg.model.faster_rcnn.num_classes
    rfcn = cfg.model.faster_rcnn.second_stage_box_predictor.WhichOneof("box_predictor_oneof") == "rfcn_box_predictor"
    seen = set()
    for p in st.params:
        if "/_pad/" in p.name:
            continue
        v = sd[p.name].numpy()
        tf = C.native_to_tf(p.name, v, p.tf_kind)
        head = p.name.split("/")[0]
        leaf = "/".join(p.name.split("/")[1:])
        if (not rfcn and head in ("SecondStageBoxPredictor", "ClosenessBoxPredictor", "WindowBoxPredictor")
                and leaf in ("BoxEncodingPredictor/weights", "ClassPredictor/weights")):
            n = {"BoxEncodingPredictor/weights": 4 * K}.get(leaf, K + 1)
            assert tf.shape == (v.shape[-1], n) and tf.ndim == 2, (p.name, tf.shape)
            assert tf[3, 1] == v[1, 0, 0, 3]
            seen.add("fc")
        elif p.name.endswith(("/Conv2d_0/weights", "/Conv2d_1a_3x3/weights")) and head == "FirstStageFeatureExtractor" \
                and p.name.count("/") == 3:
            assert tf.shape == (3, 3, 3, 32), (p.name, tf.shape)
            assert tf[1, 2, 0, 5] == v[5, 0, 0, (1 * 3 + 2) * 3 + 0] and not v[:, 0, 0, 27:].any()
            seen.add("stem")
        elif v.ndim == 4 and p.name.endswith("depthwise_weights"):
            assert tf.shape == (3, 3, v.shape[0], 1)
        elif v.ndim == 4:
            assert tf.shape == (v.shape[1], v.shape[2], v.shape[3], v.shape[0]), (p.name, tf.shape)
        elif v.ndim == 3:                                   # depthwise [C,3,3] operand -> TF [3,3,C,1]
            seen.add("dw3")
        elif v.ndim == 2:
            assert tf.shape == (v.shape[1], v.shape[0])
        else:
            assert tf.shape == v.shape
        back = C.tf_to_native(p.name, tf, p.tf_kind, p.shape)
        assert back.shape == tuple(p.shape) or back.size == v.size, (p.name, back.shape, p.shape)
        np.testing.assert_array_equal(back.reshape(v.shape), v, err_msg=p.name)
    assert ("fc" in seen) == (not rfcn)
    assert ("stem" in seen) == (name in ("model52.config", "model62.config"))
    by = st.by_name
    if not rfcn:
        assert C.native_to_tf("x", sd["SecondStageBoxPredictor/ClassPredictor/weights"].numpy(),
                              by["SecondStageBoxPredictor/ClassPredictor/weights"].tf_kind).shape[1] == K + 1
    assert C.native_to_tf("x", sd["FirstStageBoxPredictor/ClassPredictor/weights"].numpy(),
                          by["FirstStageBoxPredictor/ClassPredictor/weights"].tf_kind).shape[:2] == (1, 1)   # RPN: conv
    if cfg.model.mtl.edgemask:
        assert C.native_to_tf("x", sd["EdgeMaskPredictor/BoxEncodingPredictor/weights"].numpy(), None).shape[:2] == (1, 1)


def test_fc_and_packed_stem_variables_through_a_written_bundle(tmp_path):
    """A bundle holding TF-shaped variables ([2048, 84] FC, [3, 3, 3, 32] stem conv) loads into the GEMM-shaped
    operands; a shape the rank-only rule would have produced is refused."""
    rng = np.random.default_rng(5)
    fc = rng.normal(size=(2048, 84)).astype(np.float32)
    stem = rng.normal(size=(3, 3, 3, 32)).astype(np.float32)
    prefix = str(tmp_path / "m.ckpt")
    C.write_checkpoint(prefix, {"SecondStageBoxPredictor/BoxEncodingPredictor/weights": fc, "M/Conv2d_0/weights": stem})
    name_map = {"SecondStageBoxPredictor/BoxEncodingPredictor/weights": "SecondStageBoxPredictor/BoxEncodingPredictor/weights",
                "M/Conv2d_0/weights": "S/M/Conv2d_0/weights"}
    shapes = {"SecondStageBoxPredictor/BoxEncodingPredictor/weights": (84, 1, 1, 2048), "S/M/Conv2d_0/weights": (32, 1, 1, 64)}
    kinds = {"SecondStageBoxPredictor/BoxEncodingPredictor/weights": "fc", "S/M/Conv2d_0/weights": ("packed_conv", 3, 3, 3)}
    sd, missing = C.state_dict_from_checkpoint(C.CheckpointReader(prefix), name_map, shapes, kinds)
    assert not missing
    w = sd["SecondStageBoxPredictor/BoxEncodingPredictor/weights"]
    assert w.shape == (84, 1, 1, 2048) and w[7, 0, 0, 100] == fc[100, 7]
    s = sd["S/M/Conv2d_0/weights"]
    assert s.shape == (32, 1, 1, 64) and s[4, 0, 0, (2 * 3 + 1) * 3 + 2] == stem[2, 1, 2, 4] and not s[..., 27:].any()
    with pytest.raises(ValueError):        # without the kinds the old rank rule gives (84, 2048) / (32, 3, 3, 3): refused
        C.state_dict_from_checkpoint(C.CheckpointReader(prefix), name_map, shapes, None)


def test_reader_against_the_independently_assembled_bundle_fixture():
    """N4: tests/golden/tf_bundle_fixture.* was assembled from the published tensor-bundle / LevelDB
    table format by tests/golden/make_bundle_fixture.py with NONE of this module's code (protobuf runtime for the two
    protos, its own table builder with 16-entry restart intervals, several data blocks and shortened index separators,
    a bit-by-bit CRC-32C).  The reader must find every variable, with dtype, shape (incl. the 0-d int64 global_step)
    and values, verify all checksums, and convert the layouts a model needs."""
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    prefix = os.path.join(here, "tf_bundle_fixture")
    exp = np.load(prefix + "_expected.npz")
    r = C.CheckpointReader(prefix)                       # check_crc=True: every block trailer
    names = {k.replace("|", "/"): k for k in exp.files}
    assert set(r.get_variable_to_shape_map()) == set(names) and len(names) == 32
    for n, k in names.items():
        a = r.get_tensor(n)                              # verifies the tensor's own CRC
        assert a.dtype == exp[k].dtype and a.shape == exp[k].shape and np.array_equal(a, exp[k]), n
    assert r.get_tensor("global_step").shape == () and int(r.get_tensor("global_step")) == 123456
    # layouts: HWIO conv -> [K,R,S,C]; slim.fully_connected [in,out] -> GEMM operand [out,1,1,in]; RGB stem conv
    # [3,3,3,32] -> packed im2col rows [32,1,1,64]; depthwise [3,3,C,1] -> [C,3,3,1]; Momentum slot like its variable
    scope = "FirstStageFeatureExtractor/resnet_v1_101/block3/unit_7/bottleneck_v1/conv2"
    name_map = {scope + "/weights": scope + "/weights", scope + "/BatchNorm/moving_variance": scope + "/BatchNorm/moving_variance",
                "SecondStageBoxPredictor/ClassPredictor/weights": "SecondStageBoxPredictor/ClassPredictor/weights",
                "FirstStageFeatureExtractor/MobilenetV1/Conv2d_0/weights": "FirstStageFeatureExtractor/MobilenetV1/Conv2d_0/weights",
                "FirstStageFeatureExtractor/MobilenetV1/Conv2d_1_depthwise/depthwise_weights":
                    "FirstStageFeatureExtractor/MobilenetV1/Conv2d_1_depthwise/depthwise_weights",
                "not/in/the/file": "x"}
    shapes = {scope + "/weights": (24, 3, 3, 16), "SecondStageBoxPredictor/ClassPredictor/weights": (21, 1, 1, 40),
              "FirstStageFeatureExtractor/MobilenetV1/Conv2d_0/weights": (32, 1, 1, 64)}
    kinds = {"SecondStageBoxPredictor/ClassPredictor/weights": "fc",
             "FirstStageFeatureExtractor/MobilenetV1/Conv2d_0/weights": ("packed_conv", 3, 3, 3)}
    sd, missing = C.state_dict_from_checkpoint(r, name_map, shapes, kinds)
    assert missing == ["not/in/the/file"]
    w = exp[names[scope + "/weights"]]
    assert sd[scope + "/weights"][5, 1, 2, 7] == w[1, 2, 7, 5]
    fc = exp[names["SecondStageBoxPredictor/ClassPredictor/weights"]]
    assert sd["SecondStageBoxPredictor/ClassPredictor/weights"][3, 0, 0, 17] == fc[17, 3]
    stem = exp[names["FirstStageFeatureExtractor/MobilenetV1/Conv2d_0/weights"]]
    assert sd["FirstStageFeatureExtractor/MobilenetV1/Conv2d_0/weights"][9, 0, 0, (2 * 3 + 0) * 3 + 1] == stem[2, 0, 1, 9]
    dw = exp[names["FirstStageFeatureExtractor/MobilenetV1/Conv2d_1_depthwise/depthwise_weights"]]
    assert sd["FirstStageFeatureExtractor/MobilenetV1/Conv2d_1_depthwise/depthwise_weights"].shape == (32, 3, 3, 1)
    assert sd["FirstStageFeatureExtractor/MobilenetV1/Conv2d_1_depthwise/depthwise_weights"][4, 1, 2, 0] == dw[1, 2, 4, 0]
    mom = C.tf_to_native(scope + "/weights", r.get_tensor(scope + "/weights/Momentum"))
    assert mom.shape == (24, 3, 3, 16)
    # a flipped byte in the data file is caught by the tensor checksum
    import shutil, tempfile
    d = tempfile.mkdtemp()
    for ext in (".index", ".data-00000-of-00001"):
        shutil.copy(prefix + ext, os.path.join(d, "b" + ext))
    raw = bytearray(open(os.path.join(d, "b.data-00000-of-00001"), "rb").read())
    raw[100] ^= 0x40
    open(os.path.join(d, "b.data-00000-of-00001"), "wb").write(bytes(raw))
    bad = C.CheckpointReader(os.path.join(d, "b"))
    with pytest.raises(ValueError):
        for n in names:
            bad.get_tensor(n)
