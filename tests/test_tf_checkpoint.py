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
