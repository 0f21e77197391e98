"""CPU tests of the TFRecord / tf.Example reader: CRC-32C known answers (RFC 3720 B.4), record
framing, the hand-written protobuf wire parser cross-checked against the protobuf runtime on a dynamically built
tf.train.Example schema, and the decoder of tf_example_decoder.py:33-124 + trainer.py:100-156 round-tripped on the
synthetic examples the trainer consumes."""
import os
import struct

import numpy as np
import pytest

from mtl_ssl_b200.data import synthetic, tfrecord as T


def test_crc32c_known_answers():
    assert T.crc32c(b"") == 0
    assert T.crc32c(b"123456789") == 0xE3069283
    assert T.crc32c(bytes(32)) == 0x8A9136AA                      # RFC 3720 B.4
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    assert T.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    c = T.crc32c(b"123456789")
    assert T.masked_crc32c(b"123456789") == (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_crc32c_lane_parallel_path_matches_bytewise():
    """Buffers >= 64 KiB go through equal lanes advanced together in NumPy and chained with the GF(2) zero-advance
    operator; it must agree with the plain table loop for every length (remainders, odd sizes)."""
    rng = np.random.default_rng(4)
    for n in (65535, 65536, 65537, 70001, 262144 + 7, 1 << 20):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert T.crc32c(d) == T._crc_update_slow(0xFFFFFFFF, d) ^ 0xFFFFFFFF, n
    a = rng.integers(0, 256, 300000, dtype=np.uint8)
    assert T.crc32c(a) == T.crc32c(a.tobytes())


def test_record_framing_round_trip_and_corruption(tmp_path):
    recs = [b"", b"a", os.urandom(1000), b"x" * 70000]
    p = str(tmp_path / "r.tfrecord")
    T.write_tfrecords(p, recs)
    assert list(T.read_tfrecords(p)) == recs
    raw = bytearray(open(p, "rb").read())
    assert struct.unpack("<Q", raw[:8])[0] == 0 and len(raw) == sum(len(r) + 16 for r in recs)
    raw[16 + 12 + 1] ^= 1                                         # flip a payload bit of the second record
    open(p, "wb").write(raw)
    with pytest.raises(ValueError):
        list(T.read_tfrecords(p))
    open(p, "wb").write(raw[:-3])
    with pytest.raises(ValueError):
        list(T.read_tfrecords(p, check_crc=False))


def _example_class():
    """tf.train.Example (tensorflow/core/example/{example,feature}.proto) rebuilt with the protobuf runtime."""
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="mtl_test_example.proto", package="mtltest", syntax="proto3")
    F = descriptor_pb2.FieldDescriptorProto

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = tname
        return m
    msg("BytesList", [("value", 1, F.TYPE_BYTES, F.LABEL_REPEATED, None)])
    msg("FloatList", [("value", 1, F.TYPE_FLOAT, F.LABEL_REPEATED, None)])
    msg("Int64List", [("value", 1, F.TYPE_INT64, F.LABEL_REPEATED, None)])
    feat = msg("Feature", [("bytes_list", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".mtltest.BytesList"),
                           ("float_list", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".mtltest.FloatList"),
                           ("int64_list", 3, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".mtltest.Int64List")])
    feat.oneof_decl.add(name="kind")
    for f in feat.field:
        f.oneof_index = 0
    feats = msg("Features", [("feature", 1, F.TYPE_MESSAGE, F.LABEL_REPEATED, ".mtltest.Features.FeatureEntry")])
    entry = feats.nested_type.add(name="FeatureEntry")
    entry.options.map_entry = True
    entry.field.add(name="key", number=1, type=F.TYPE_STRING, label=F.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=F.TYPE_MESSAGE, label=F.LABEL_OPTIONAL, type_name=".mtltest.Feature")
    msg("Example", [("features", 1, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, ".mtltest.Features")])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("mtltest.Example"))


def test_example_wire_format_against_protobuf_runtime():
    Example = _example_class()
    rng = np.random.default_rng(0)
    ex = Example()
    ex.features.feature["image/encoded"].bytes_list.value.append(os.urandom(300))
    ex.features.feature["image/object/subset"].bytes_list.value.extend([b"default", b"", b"a|b"])
    fl = rng.normal(size=37).astype(np.float32)
    ex.features.feature["image/object/bbox/ymin"].float_list.value.extend(fl.tolist())
    il = np.array([0, 1, -1, 2 ** 40, -2 ** 62, 127, 128, 300], np.int64)
    ex.features.feature["image/object/class/label"].int64_list.value.extend(il.tolist())
    ex.features.feature["empty/floats"].float_list.SetInParent()
    ex.features.feature["kind/not/set"].SetInParent()
    got = T.parse_example(ex.SerializeToString())
    assert got["image/encoded"] == list(ex.features.feature["image/encoded"].bytes_list.value)
    assert got["image/object/subset"] == [b"default", b"", b"a|b"]
    assert np.array_equal(got["image/object/bbox/ymin"], fl)
    assert np.array_equal(got["image/object/class/label"], il)
    assert len(got["empty/floats"]) == 0 and len(got["kind/not/set"]) == 0
    # and the other direction: the protobuf runtime reads what serialize_example writes
    back = Example()
    back.ParseFromString(T.serialize_example({"f": fl, "i": il, "b": [b"x", b"yz"], "s": "text"}))
    assert np.array_equal(np.asarray(back.features.feature["f"].float_list.value, np.float32), fl)
    assert list(back.features.feature["i"].int64_list.value) == il.tolist()
    assert list(back.features.feature["b"].bytes_list.value) == [b"x", b"yz"]
    assert list(back.features.feature["s"].bytes_list.value) == [b"text"]


def test_unpacked_repeated_scalars_are_accepted():
    """proto2-era writers emit one tag per element instead of a packed block."""
    def ld(num, payload):
        return T._enc_varint((num << 3) | 2) + T._enc_varint(len(payload)) + payload
    floats = b"".join(T._enc_varint((1 << 3) | 5) + struct.pack("<f", v) for v in (1.5, -2.0))
    ints = b"".join(T._enc_varint((1 << 3) | 0) + T._enc_varint(v) for v in (7, 300))
    body = ld(1, ld(1, b"f") + ld(2, ld(2, floats))) + ld(1, ld(1, b"i") + ld(2, ld(3, ints)))
    got = T.parse_example(ld(1, body))
    assert got["f"].tolist() == [1.5, -2.0] and got["i"].tolist() == [7, 300]


def test_decode_round_trip_feeds_the_trainer(tmp_path):
    from mtl_ssl_b200.trainer import pack_groundtruth
    K, H, W = 20, 96, 128
    examples = synthetic.make_batch(3, 4, H, W, K, max_boxes=5, num_windows=16)
    p = str(tmp_path / "voc.tfrecord")
    T.write_tfrecords(p, [T.encode_example(e, "png", "img%d.png" % i) for i, e in enumerate(examples)])
    decoded = list(T.TfRecordDataset(p, K))
    assert len(decoded) == len(examples)
    for e, d in zip(examples, decoded):
        assert np.array_equal(d["image"], e["image"])                       # PNG is lossless
        np.testing.assert_array_equal(d["groundtruth_boxes"], e["groundtruth_boxes"])
        np.testing.assert_array_equal(d["groundtruth_classes"], e["groundtruth_classes"])
        np.testing.assert_allclose(d["groundtruth_closeness"], e["groundtruth_closeness"], atol=5e-4)
        np.testing.assert_array_equal(d["window_boxes"], e["window_boxes"])
        np.testing.assert_allclose(d["window_classes"], e["window_classes"], atol=5e-4)
        np.testing.assert_array_equal(d["groundtruth_edgemask"], e["groundtruth_edgemask"])
        assert d["image"].dtype == np.float32 and d["groundtruth_classes"].shape == (len(e["groundtruth_boxes"]), K)
    a = pack_groundtruth(examples, K, H, W, 8)
    b = pack_groundtruth(decoded, K, H, W, 8)
    for k in a:
        np.testing.assert_allclose(b[k], a[k], atol=5e-4, err_msg=k)
    # JPEG records decode too (lossy): same geometry, pixels close on a smooth image
    smooth = dict(examples[0])
    yy, xx = np.mgrid[0:H, 0:W]
    smooth["image"] = np.stack([yy * 2.0, xx * 1.5, (yy + xx) * 0.9], -1).astype(np.float32)
    dj = T.decode_example(T.encode_example(smooth, "jpeg"), K)
    assert dj["image"].shape == (H, W, 3) and np.abs(dj["image"] - smooth["image"]).mean() < 3.0
    # labels outside 1..K are an error, as the one-hot encoding of the reference would silently drop them
    bad = T.parse_example(T.encode_example(examples[0]))
    bad["image/object/class/label"] = np.asarray([K + 1] * len(bad["image/object/class/label"]), np.int64)
    with pytest.raises(ValueError):
        T.decode_example(T.serialize_example(bad), K)


def test_horizontal_flip_reference_vectors():
    """preprocessor_test.py:47-66, :165-184 (testRandomHorizontalFlip): image columns reversed, boxes mirrored; plus the
    fork's window boxes and the edge-mask axis trap (T16)."""
    from mtl_ssl_b200.data import augment as A
    r = np.array([[128, 128, 128, 128], [0, 0, 128, 128], [0, 128, 128, 128], [192, 192, 128, 128]], np.float32)
    g = np.array([[0, 0, 128, 128], [0, 0, 128, 128], [0, 128, 192, 192], [192, 192, 128, 192]], np.float32)
    b = np.array([[128, 128, 192, 0], [0, 0, 128, 192], [0, 128, 128, 0], [192, 192, 192, 128]], np.float32)
    img = np.stack([r, g, b], -1)
    em = np.arange(2 * 3 * 4, dtype=np.float32).reshape(2, 3, 4)
    ex = dict(image=img, groundtruth_boxes=np.array([[0.0, 0.25, 0.75, 1.0], [0.25, 0.5, 0.75, 1.0]], np.float32),
              window_boxes=np.array([[0.1, 0.2, 0.3, 0.6]], np.float32), groundtruth_edgemask=em)
    f = A.horizontal_flip(ex)
    want_r = (np.array([[0, 0, 0, 0], [0, 0, -1, -1], [0, 0, 0, -1], [0, 0, 0.5, 0.5]], np.float32) + 1) * 128
    assert np.array_equal(f["image"][..., 0], want_r)                       # expectedImagesAfterMirroring, red plane
    assert np.array_equal(f["image"], img[:, ::-1])
    np.testing.assert_allclose(f["groundtruth_boxes"], [[0.0, 0.0, 0.75, 0.75], [0.25, 0.0, 0.75, 0.5]])
    np.testing.assert_allclose(f["window_boxes"], [[0.1, 0.4, 0.3, 0.8]], rtol=1e-6)
    assert np.array_equal(f["groundtruth_edgemask"], em[:, ::-1, :])        # the reference reverses the mask ROWS
    assert np.array_equal(A.horizontal_flip(ex, reference_edgemask_axis=False)["groundtruth_edgemask"], em[:, :, ::-1])
    # probability 1/2, never without boxes
    rng = np.random.default_rng(0)
    flips = sum(A.random_horizontal_flip(ex, rng) is not ex for _ in range(2000))
    assert 900 < flips < 1100
    empty = dict(ex, groundtruth_boxes=np.zeros((0, 4), np.float32))
    assert all(A.random_horizontal_flip(empty, rng) is empty for _ in range(50))
    assert [len(x) for x in A.batches(range(7), 3)] == [3, 3]
    assert [len(x) for x in A.batches(range(7), 3, drop_remainder=False)] == [3, 3, 1]


def test_input_examples_from_pipeline_config(tmp_path):
    from helpers import load_config
    from mtl_ssl_b200.data import augment as A
    K, H, W = 20, 64, 80
    examples = synthetic.make_batch(5, 3, H, W, K, max_boxes=3, num_windows=8)
    p = str(tmp_path / "train.record")
    T.write_tfrecords(p, [T.encode_example(e) for e in examples])
    cfg = load_config("model12.config", (("../data/voc/voc2007_trainval.record", p),))
    reader = cfg.train_input_reader
    assert reader.tf_record_input_reader.input_path == p and reader.shuffle is True      # proto default
    got = list(A.input_examples(reader, K, options=("random_horizontal_flip",), seed=1, epochs=2))
    assert len(got) == 6
    areas = sorted(float(np.prod(e["groundtruth_boxes"][:, 2:] - e["groundtruth_boxes"][:, :2], 1).sum()) for e in got)
    want = sorted(float(np.prod(e["groundtruth_boxes"][:, 2:] - e["groundtruth_boxes"][:, :2], 1).sum())
                  for e in examples) * 2
    np.testing.assert_allclose(areas, sorted(want), rtol=1e-5)            # flips preserve box areas
    with pytest.raises(NotImplementedError):
        next(A.input_examples(reader, K, options=("random_crop_image",)))


def test_prefetch_loader_order_errors_and_shutdown(tmp_path):
    from mtl_ssl_b200.data import loader as L
    from mtl_ssl_b200.trainer import pack_groundtruth
    K, H, W = 20, 64, 80
    examples = synthetic.make_batch(8, 6, H, W, K, max_boxes=3, num_windows=8)
    p = str(tmp_path / "t.record")
    T.write_tfrecords(p, [T.encode_example(e) for e in examples])

    def pack(group, keys):
        a = pack_groundtruth(group, K, H, W, 8)
        a["image"] = np.stack([e["image"] for e in group]).astype(np.float32)
        a["keys1"], a["keys2"] = keys
        return a
    keys_fn = L.sampler_keys_fn(3, 2, 50, 10)
    got = list(L.PrefetchLoader(T.TfRecordDataset(p, K), pack, keys_fn, batch_size=2, depth=2))
    assert len(got) == 3
    for i, a in enumerate(got):
        want = pack(examples[2 * i:2 * i + 2], keys_fn(i))
        for k in want:
            np.testing.assert_allclose(a[k], want[k], atol=5e-4, err_msg=k)
    assert not np.array_equal(got[0]["keys1"], got[1]["keys1"])          # fresh sampler keys every step
    # an exception in the producer surfaces in the consumer

    def bad():
        yield examples[0]
        yield examples[1]
        raise RuntimeError("decode failed")
    it = L.PrefetchLoader(bad(), pack, keys_fn, batch_size=1, depth=1)
    next(it); next(it)
    with pytest.raises(RuntimeError):
        next(it)
    # closing a loader whose queue is full does not hang
    it = L.PrefetchLoader(iter(examples * 50), pack, keys_fn, batch_size=1, depth=1)
    next(it)
    it.close()
    assert not it._thread.is_alive()

    class FakeTrainer(object):
        def __init__(self):
            self.pending, self.n = None, 0

        def step_pipelined(self, arrays):
            prev, self.pending = self.pending, {"total_loss": float(arrays["image"].mean())}
            self.n += 1
            return prev

        def flush(self):
            prev, self.pending = self.pending, None
            return prev
    ft = FakeTrainer()
    losses = L.train_loop(ft, L.PrefetchLoader(iter(examples), pack, keys_fn, batch_size=2, depth=2))
    assert ft.n == 3 and len(losses) == 3


_VOC_XML = """<annotation><folder>VOC2007</folder><filename>%s</filename>
<size><width>256</width><height>256</height><depth>3</depth></size>
<object><name>person</name><pose>Left</pose><truncated>0</truncated><difficult>1</difficult>
<bndbox><xmin>64</xmin><ymin>64</ymin><xmax>192</xmax><ymax>192</ymax></bndbox></object>
<object><name>notperson</name><pose>Frontal</pose><truncated>1</truncated><difficult>0</difficult>
<bndbox><xmin>10</xmin><ymin>20</ymin><xmax>110</xmax><ymax>220</ymax></bndbox></object></annotation>"""


def test_pascal_voc_records_reference_expectations(tmp_path):
    """create_pascal_tf_record_test.py:39-113 (`test_dict_to_tf_example`: the standard keys of one 256x256 image with a
    'person' box 64..192) + XML parsing, the aux-label keys, the VOC directory reader and the decoder round trip."""
    from PIL import Image
    from mtl_ssl_b200.data import pascal_voc as V
    root = tmp_path / "VOC2007"
    for d in ("JPEGImages", "Annotations", "ImageSets/Main"):
        os.makedirs(str(root / d))
    name = "tmp_image.jpg"
    rng = np.random.default_rng(0)
    Image.fromarray(rng.integers(0, 256, (256, 256, 3), dtype=np.uint8), "RGB").save(str(root / "JPEGImages" / name))
    data = {"folder": "", "filename": name, "size": {"height": 256, "width": 256},
            "object": [{"difficult": 1, "bndbox": {"xmin": 64, "ymin": 64, "xmax": 192, "ymax": 192}, "name": "person",
                        "truncated": 0, "pose": ""}]}
    label_map = {"background": 0, "person": 1, "notperson": 2}
    ex = T.parse_example(V.dict_to_tf_example(data, str(root), label_map, image_subdirectory="JPEGImages"))
    assert ex["image/height"].tolist() == [256] and ex["image/width"].tolist() == [256]
    assert ex["image/filename"] == [name.encode()] and ex["image/source_id"] == [name.encode()]
    assert ex["image/format"] == [b"jpeg"]
    for k, v in (("xmin", 0.25), ("ymin", 0.25), ("xmax", 0.75), ("ymax", 0.75)):
        assert ex["image/object/bbox/" + k].tolist() == [v]
    assert ex["image/object/class/text"] == [b"person"] and ex["image/object/class/label"].tolist() == [1]
    assert ex["image/object/difficult"].tolist() == [1] and ex["image/object/truncated"].tolist() == [0]
    assert ex["image/object/view"] == [b""] and ex["image/object/subset"] == [b"all"]
    assert len(ex["image/key/sha256"][0]) == 64
    # the fork's keys: 64 windows with K+1 soft labels, one closeness row per object, a [2,64,64] edge mask
    assert len(ex["image/window/labels/text"]) == len(ex["image/window/bbox/ymin"]) == 64
    assert len(ex["image/object/closeness/text"]) == 1
    assert ex["image/edgemask/height"].tolist() == [64] and len(ex["image/edgemask/masks"]) == 2 * 64 * 64
    with pytest.raises(ValueError):                                       # PNG on disk: "Image format not JPEG"
        Image.fromarray(np.zeros((8, 8, 3), np.uint8)).save(str(root / "JPEGImages" / "p.jpg"), format="PNG")
        V.dict_to_tf_example(dict(data, filename="p.jpg"), str(root), label_map)
    # XML -> dict -> example, the directory reader, and the record decoder agree
    open(str(root / "Annotations" / "tmp_image.xml"), "w").write(_VOC_XML % name)
    open(str(root / "ImageSets" / "Main" / "trainval.txt"), "w").write("tmp_image\n")
    parsed = V.parse_annotation(str(root / "Annotations" / "tmp_image.xml"))
    assert [o["name"] for o in parsed["object"]] == ["person", "notperson"] and parsed["size"]["width"] == "256"
    assert parsed["object"][1]["bndbox"] == {"xmin": "10", "ymin": "20", "xmax": "110", "ymax": "220"}
    ds = V.VocDataset(str(root), "trainval", label_map, num_classes=2, seed=3, num_windows=16)
    (e,) = list(ds)
    assert len(ds) == 1 and e["image"].shape == (256, 256, 3) and e["groundtruth_difficult"].tolist() == [True, False]
    np.testing.assert_allclose(e["groundtruth_boxes"], [[0.25, 0.25, 0.75, 0.75], [20 / 256, 10 / 256, 220 / 256, 110 / 256]])
    assert e["groundtruth_classes"].tolist() == [[1, 0], [0, 1]]
    assert e["window_boxes"].shape == (16, 4) and e["window_classes"].shape == (16, 3)
    assert e["groundtruth_closeness"].shape == (2, 3) and e["groundtruth_edgemask"].shape == (2, 64, 64)
    # the XML's <folder> is joined under the devkit root, like the reference (`data['folder']/JPEGImages/<file>`)
    rec = V.dict_to_tf_example(parsed, str(tmp_path), label_map, num_classes=2, rng=np.random.default_rng(3),
                               num_windows=16)
    d = T.decode_example(rec, 2)
    np.testing.assert_allclose(d["groundtruth_boxes"], e["groundtruth_boxes"], rtol=1e-6)
    np.testing.assert_array_equal(d["groundtruth_classes"], e["groundtruth_classes"])
    np.testing.assert_allclose(d["window_boxes"], e["window_boxes"], rtol=1e-6)
    np.testing.assert_allclose(d["window_classes"], e["window_classes"], atol=5e-4)
    np.testing.assert_allclose(d["groundtruth_closeness"], e["groundtruth_closeness"], atol=5e-4)
    np.testing.assert_array_equal(d["groundtruth_edgemask"], e["groundtruth_edgemask"])
    assert d["groundtruth_difficult"].tolist() == [True, False] and d["groundtruth_subset"] == ["all", "all"]
    only_easy = list(V.VocDataset(str(root), "trainval", label_map, 2, ignore_difficult_instances=True))[0]
    assert only_easy["groundtruth_classes"].tolist() == [[0, 1]]


def test_mscoco_records_and_dataset_reader(tmp_path):
    """create_mscoco_tf_record.py:87-477 key set through data/mscoco.py: annotation JSON -> CocoIndex (no pycocotools) ->
    record -> decoder, and the directory reader; the label values themselves are pinned in
    test_oracle_kats.py::test_coco_examples_against_reference_mscoco_record_writer."""
    import json
    from PIL import Image
    from mtl_ssl_b200.data import mscoco as C
    img_dir = tmp_path / "images" / "val2017"
    os.makedirs(str(img_dir))
    fname = "%012d.jpg" % 139
    rng = np.random.default_rng(0)
    Image.fromarray(rng.integers(0, 256, (200, 320, 3), dtype=np.uint8), "RGB").save(str(img_dir / fname))
    ann = {"images": [{"id": 139, "file_name": fname, "height": 200, "width": 320}],
           "categories": [{"id": 1, "name": "person"}, {"id": 3, "name": "car"}, {"id": 90, "name": "toothbrush"}],
           "annotations": [{"id": 7, "image_id": 139, "category_id": 3, "iscrowd": 0, "bbox": [240.0, 20.0, 108.0, 60.0]},
                           {"id": 8, "image_id": 139, "category_id": 90, "iscrowd": 1, "bbox": [160.0, 100.0, 80.0, 50.0]},
                           {"id": 9, "image_id": 139, "category_id": 1, "iscrowd": 0, "bbox": [100.0, 50.0, 0.0, 30.0]}]}
    open(str(tmp_path / "instances_val2017.json"), "w").write(json.dumps(ann))
    coco = C.CocoIndex(str(tmp_path / "instances_val2017.json"))
    assert coco.getAnnIds(imgIds=139) == [7, 8, 9] and coco.loadCats(3)[0]["name"] == "car"
    assert C.boundary_check([-8.0, 20.0, 108.0, 60.0], 320, 200) == (0, 20.0, 108.0, 60.0)
    assert C.boundary_check([240.0, 20.0, 108.0, 60.0], 320, 200) == (240.0, 20.0, 80.0, 60.0)
    label_map = coco.label_map_dict()
    rec = C.dict_to_tf_example(label_map, str(img_dir / fname), coco, sorted(label_map.values()),
                               rng=np.random.default_rng(1), num_windows=16)
    ex = T.parse_example(rec)
    assert ex["image/source_id"] == [b"139"] and ex["image/format"] == [b"jpg"]
    assert ex["image/height"].tolist() == [200] and ex["image/width"].tolist() == [320]
    # the third box has zero width (COCO holds a few): dropped from the ground truth, still seen by the auxiliary labels
    assert ex["image/object/class/text"] == [b"car", b"toothbrush"] and ex["image/object/class/label"].tolist() == [3, 90]
    assert ex["image/object/is_crowd"].tolist() == [0, 1]
    np.testing.assert_allclose(ex["image/object/bbox/xmin"], [0.75, 0.5])
    np.testing.assert_allclose(ex["image/object/bbox/xmax"], [1.0, 0.75])
    assert len(ex["image/object/closeness/text"]) == 2 and len(ex["image/window/labels/text"]) == 16
    assert len(ex["image/window/labels/text"][0].split()) == 91 and len(ex["image/object/closeness/text"][0].split()) == 91
    d = T.decode_example(rec, 90)
    assert d["groundtruth_classes"].shape == (2, 90) and d["groundtruth_classes"][0, 2] == 1 and d["groundtruth_classes"][1, 89] == 1
    assert d["window_classes"].shape == (16, 91) and d["groundtruth_closeness"].shape == (2, 91)
    assert d["groundtruth_edgemask"].shape == (2, 64, 64)
    ds = C.CocoDataset(coco, str(img_dir), num_classes=90, seed=1, num_windows=16)
    (e,) = list(ds)
    assert len(ds) == 1 and e["image"].shape == (200, 320, 3) and e["source_id"] == "139" == d["source_id"]
    np.testing.assert_allclose(e["groundtruth_boxes"], d["groundtruth_boxes"], rtol=1e-6)
    np.testing.assert_array_equal(e["groundtruth_classes"], d["groundtruth_classes"])
    np.testing.assert_allclose(e["window_classes"], d["window_classes"], atol=5e-4)
    np.testing.assert_allclose(e["groundtruth_closeness"], d["groundtruth_closeness"], atol=5e-4)
    # an image without annotations has no windows (`create_multi_object` returns [])
    empty = C.annotations_to_example([], np.zeros((50, 60, 3), np.uint8), {}, label_map, 90, np.random.default_rng(0))
    assert empty["window_boxes"].shape == (0, 4) and empty["groundtruth_boxes"].shape == (0, 4)
