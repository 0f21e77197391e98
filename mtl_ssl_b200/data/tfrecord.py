"""TFRecord / tf.Example input side of the training path: the wire format the reference's
`train_input_reader` feeds to the model, read without TensorFlow.

  * record framing of a TFRecord file (tensorflow/core/lib/io/record_writer.cc: u64 length, masked CRC-32C of
    the length, payload, masked CRC-32C of the payload),
  * the `tf.train.Example` protobuf (Features = map<string, Feature{bytes_list | float_list | int64_list}>),
    parsed straight from the wire encoding,
  * the key set and post-processing of /root/reference/object_detection/data_decoders/tf_example_decoder.py:33-124
    and of `trainer._get_inputs` (trainer.py:100-156): boxes as [ymin, xmin, ymax, xmax] (normalised), class
    labels shifted by label_id_offset = 1 and one-hot encoded, the fork's window / closeness labels stored as
    space-separated text of num_classes + 1 floats per row, and the edge mask stored as floats reshaped to
    [-1, height, width].

`decode_example` returns the example dict of mtl_ssl_b200/data/synthetic.py, i.e. what `Trainer.host_arrays`
consumes; `encode_example` writes the same keys the reference's create_pascal_tf_record.py writes, so that records can be
produced and round-tripped here (there is no dataset and no TensorFlow in this environment)."""
import io
import struct

import numpy as np

# ----------------------------------------------------------------------------- CRC-32C (Castagnoli)
_POLY = 0x82F63B78
_TABLE = []
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ _POLY if _c & 1 else _c >> 1
    _TABLE.append(_c)


_TABLE_NP = np.array(_TABLE, dtype=np.uint32)


def _crc_update_slow(c, data):
    for b in bytes(data):
        c = _TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c


def _gf2_apply(cols, v):
    """Matrix (32 columns as integers) times bit-vector v over GF(2)."""
    out, j = 0, 0
    while v:
        if v & 1:
            out ^= cols[j]
        v >>= 1
        j += 1
    return out


def _zero_advance_matrix(nbytes):
    """Columns of the linear map "CRC register after `nbytes` zero bytes" (the register update is linear over GF(2))."""
    cols = [_TABLE[(1 << j) & 0xFF] ^ ((1 << j) >> 8) for j in range(32)]       # one zero byte
    result = [1 << j for j in range(32)]                                        # identity
    n = nbytes
    while n:
        if n & 1:
            result = [_gf2_apply(cols, c) for c in result]
        cols = [_gf2_apply(cols, c) for c in cols]
        n >>= 1
    return result


def crc32c(data):
    """CRC-32C of a bytes-like object (reflected, init / final xor 0xFFFFFFFF).  Large buffers (checkpoint tensors)
    are cut into equal lanes whose registers advance together in NumPy; the lanes are then chained with the
    zero-advance operator: R(s, A||B) = R(0, B) xor Z(R(s, A), len B)."""
    buf = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8).reshape(-1)
    n = buf.size
    if n < (1 << 16):
        return _crc_update_slow(0xFFFFFFFF, buf.tobytes()) ^ 0xFFFFFFFF
    lanes = int(min(8192, n >> 10))
    chunk = n // lanes
    body = buf[:lanes * chunk].reshape(lanes, chunk)
    regs = np.zeros(lanes, np.uint32)
    for j in range(chunk):
        regs = _TABLE_NP[(regs ^ body[:, j]) & 0xFF] ^ (regs >> 8)
    adv = _zero_advance_matrix(chunk)
    state = 0xFFFFFFFF
    for r in regs.tolist():
        state = _gf2_apply(adv, state) ^ r
    return _crc_update_slow(state, buf[lanes * chunk:].tobytes()) ^ 0xFFFFFFFF


def masked_crc32c(data):
    """TFRecord's masked CRC: rotate right by 15 and add a constant (record_writer.h `crc32c::Mask`)."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ----------------------------------------------------------------------------- record framing
def write_tfrecords(path, records):
    with open(path, "wb") as f:
        for r in records:
            head = struct.pack("<Q", len(r))
            f.write(head)
            f.write(struct.pack("<I", masked_crc32c(head)))
            f.write(r)
            f.write(struct.pack("<I", masked_crc32c(r)))


def read_tfrecords(path, check_crc=True):
    """Yields the payload of every record; raises ValueError on a truncated file or a CRC mismatch."""
    with open(path, "rb") as f:
        while True:
            head = f.read(8)
            if not head:
                return
            if len(head) != 8:
                raise ValueError("%s: truncated record header" % path)
            (n,) = struct.unpack("<Q", head)
            crc = f.read(4)
            if len(crc) != 4 or (check_crc and struct.unpack("<I", crc)[0] != masked_crc32c(head)):
                raise ValueError("%s: corrupted record length" % path)
            data = f.read(n)
            tail = f.read(4)
            if len(data) != n or len(tail) != 4:
                raise ValueError("%s: truncated record" % path)
            if check_crc and struct.unpack("<I", tail)[0] != masked_crc32c(data):
                raise ValueError("%s: corrupted record payload" % path)
            yield data


# ----------------------------------------------------------------------------- protobuf wire format
def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if not b & 0x80:
            return out, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def _fields(buf):
    """Iterate (field number, wire type, value) over one message; length-delimited values are memoryviews."""
    buf = memoryview(buf)
    pos, n = 0, len(buf)
    while pos < n:
        tag, pos = _varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = buf[pos:pos + 8], pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v, pos = buf[pos:pos + ln], pos + ln
            if len(v) != ln:
                raise ValueError("truncated length-delimited field")
        elif wt == 5:
            v, pos = buf[pos:pos + 4], pos + 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        yield num, wt, v


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_feature(buf):
    """Feature { BytesList bytes_list = 1; FloatList float_list = 2; Int64List int64_list = 3; } (oneof)."""
    for num, wt, v in _fields(buf):
        if wt != 2:
            continue
        if num == 1:
            return [bytes(x) for n2, w2, x in _fields(v) if n2 == 1 and w2 == 2]
        if num == 2:
            vals = []
            for n2, w2, x in _fields(v):
                if n2 != 1:
                    continue
                if w2 == 2:                       # packed (the default for proto3 repeated scalars)
                    vals.append(np.frombuffer(x, dtype="<f4"))
                elif w2 == 5:
                    vals.append(np.frombuffer(x, dtype="<f4", count=1))
            return np.concatenate(vals).astype(np.float32) if vals else np.zeros((0,), np.float32)
        if num == 3:
            vals = []
            for n2, w2, x in _fields(v):
                if n2 != 1:
                    continue
                if w2 == 2:
                    p, m = 0, len(x)
                    while p < m:
                        t, p = _varint(x, p)
                        vals.append(_signed64(t))
                elif w2 == 0:
                    vals.append(_signed64(x))
            return np.asarray(vals, np.int64)
    return []                                      # kind not set: an empty feature


def parse_example(serialized):
    """tf.train.Example bytes -> {key: list of bytes | float32 array | int64 array}."""
    out = {}
    for num, wt, features in _fields(serialized):
        if num != 1 or wt != 2:
            continue
        for n2, w2, entry in _fields(features):     # map<string, Feature> feature = 1
            if n2 != 1 or w2 != 2:
                continue
            key, value = None, []
            for n3, w3, x in _fields(entry):
                if n3 == 1 and w3 == 2:
                    key = bytes(x).decode("utf-8")
                elif n3 == 2 and w3 == 2:
                    value = _parse_feature(x)
            if key is not None:
                out[key] = value
    return out


def _enc_varint(v):
    v &= (1 << 64) - 1
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(num, payload):
    return _enc_varint((num << 3) | 2) + _enc_varint(len(payload)) + payload


def serialize_example(features):
    """{key: list of bytes/str | float sequence | int sequence} -> tf.train.Example bytes (packed scalars)."""
    body = b""
    for key in sorted(features):
        v = features[key]
        if isinstance(v, (bytes, str)):
            v = [v]
        if isinstance(v, np.ndarray) and v.dtype.kind == "f":
            feat = _ld(2, _ld(1, np.asarray(v, "<f4").tobytes()) if len(v) else b"")
        elif isinstance(v, np.ndarray) and v.dtype.kind in "iub":
            feat = _ld(3, _ld(1, b"".join(_enc_varint(int(x)) for x in v)) if len(v) else b"")
        elif len(v) and isinstance(v[0], (bytes, str)):
            feat = _ld(1, b"".join(_ld(1, x.encode("utf-8") if isinstance(x, str) else x) for x in v))
        elif len(v) and isinstance(v[0], (float, np.floating)):
            feat = _ld(2, _ld(1, np.asarray(v, "<f4").tobytes()))
        elif len(v):
            feat = _ld(3, _ld(1, b"".join(_enc_varint(int(x)) for x in v)))
        else:
            feat = _ld(1, b"")
        body += _ld(1, _ld(1, key.encode("utf-8")) + _ld(2, feat))
    return _ld(1, body)


# ----------------------------------------------------------------------------- the decoder of the reference
def _text_rows(rows, width):
    """`tf.string_split` + `string_to_number` + reshape [-1, width] (trainer.py:139-150)."""
    vals = [float(t) for r in rows for t in (r.decode("utf-8") if isinstance(r, bytes) else r).split()]
    if len(vals) % width:
        raise ValueError("text labels hold %d numbers, not a multiple of %d" % (len(vals), width))
    return np.asarray(vals, np.float32).reshape(-1, width)


def _boxes(ex, prefix):
    cols = [np.asarray(ex.get(prefix + k, []), np.float32) for k in ("ymin", "xmin", "ymax", "xmax")]
    if len({len(c) for c in cols}) != 1:
        raise ValueError("%s*: coordinate lists of different lengths" % prefix)
    return np.stack(cols, 1).reshape(-1, 4)


def decode_image(encoded, fmt=b"jpeg"):
    """slim_example_decoder.Image(channels=3): uint8 [H, W, 3]."""
    from PIL import Image
    img = Image.open(io.BytesIO(encoded))
    return np.asarray(img.convert("RGB"), np.uint8)


def decode_example(serialized, num_classes, label_id_offset=1):
    """One serialized tf.Example -> the example dict of data/synthetic.py (tf_example_decoder.py:33-124 +
    trainer.py:100-156): image float32 [H,W,3] (tf.to_float of the decoded uint8), groundtruth_boxes [G,4]
    normalised, groundtruth_classes one-hot [G,K] of label - 1, groundtruth_closeness [G,K+1], window_boxes [Nw,4],
    window_classes [Nw,K+1], groundtruth_edgemask [2,h,w]; plus filename / source_id / difficult / is_crowd / subset."""
    ex = parse_example(serialized)
    K = num_classes
    enc = ex.get("image/encoded", [b""])
    fmt = ex.get("image/format", [b"jpeg"])
    image = decode_image(enc[0] if enc else b"", fmt[0] if fmt else b"jpeg").astype(np.float32)
    boxes = _boxes(ex, "image/object/bbox/")
    labels = np.asarray(ex.get("image/object/class/label", []), np.int64) - label_id_offset
    if len(labels) != len(boxes):
        raise ValueError("%d labels for %d boxes" % (len(labels), len(boxes)))
    if len(labels) and (labels.min() < 0 or labels.max() >= K):
        raise ValueError("class label outside 1..%d" % K)
    onehot = np.zeros((len(labels), K), np.float32)        # util_ops.padded_one_hot_encoding(depth=K, left_pad=0)
    onehot[np.arange(len(labels)), labels] = 1.0
    out = dict(image=image, groundtruth_boxes=boxes, groundtruth_classes=onehot)
    if ex.get("image/object/closeness/text"):
        out["groundtruth_closeness"] = _text_rows(ex["image/object/closeness/text"], K + 1)
    if "image/window/bbox/ymin" in ex:
        out["window_boxes"] = _boxes(ex, "image/window/bbox/")
        out["window_classes"] = _text_rows(ex.get("image/window/labels/text", []), K + 1)
        if len(out["window_classes"]) != len(out["window_boxes"]):
            raise ValueError("window labels and window boxes differ in length")
    if len(ex.get("image/edgemask/masks", [])):
        h, w = int(ex["image/edgemask/height"][0]), int(ex["image/edgemask/width"][0])
        out["groundtruth_edgemask"] = np.asarray(ex["image/edgemask/masks"], np.float32).reshape(-1, h, w)
    for key, name in (("image/filename", "filename"), ("image/source_id", "source_id"), ("image/key/sha256", "key")):
        if ex.get(key):
            out[name] = ex[key][0].decode("utf-8")
    for key, name, dt in (("image/object/difficult", "groundtruth_difficult", bool),
                          ("image/object/is_crowd", "groundtruth_is_crowd", bool),
                          ("image/object/area", "groundtruth_area", np.float32)):
        if key in ex and len(ex[key]):
            out[name] = np.asarray(ex[key]).astype(dt)
    if ex.get("image/object/subset"):
        out["groundtruth_subset"] = [s.decode("utf-8") for s in ex["image/object/subset"]]
    return out


def _fmt_row(row, n_round=3):
    """create_pascal_tf_record.py:121-124 `get_string_label`: values rounded to 3 decimals, space separated,
    trailing space."""
    return " ".join(str(round(float(v), n_round)) for v in row) + " "


def encode_example(example, image_format="png", filename="", quality=95):
    """Inverse of decode_example for the keys the fork's record writers emit
    (create_records/create_pascal_tf_record.py:325-421): returns serialized tf.Example bytes."""
    from PIL import Image
    img = np.clip(np.asarray(example["image"]), 0, 255).astype(np.uint8)
    buf = io.BytesIO()
    if image_format == "png":
        Image.fromarray(img).save(buf, format="PNG")
    else:
        Image.fromarray(img).save(buf, format="JPEG", quality=quality)
    boxes = np.asarray(example["groundtruth_boxes"], np.float32).reshape(-1, 4)
    labels = np.asarray(example["groundtruth_classes"]).argmax(1) + 1 if len(boxes) else np.zeros((0,), np.int64)
    f = {
        "image/encoded": [buf.getvalue()], "image/format": [image_format.encode()],
        "image/filename": [filename.encode()], "image/source_id": [filename.encode()],
        "image/height": np.asarray([img.shape[0]], np.int64), "image/width": np.asarray([img.shape[1]], np.int64),
        "image/object/class/label": labels.astype(np.int64),
    }
    for i, k in enumerate(("ymin", "xmin", "ymax", "xmax")):
        f["image/object/bbox/" + k] = boxes[:, i].astype(np.float32)
    if "groundtruth_closeness" in example:
        f["image/object/closeness/text"] = [_fmt_row(r).encode() for r in np.asarray(example["groundtruth_closeness"])]
    if "window_boxes" in example:
        wb = np.asarray(example["window_boxes"], np.float32).reshape(-1, 4)
        for i, k in enumerate(("ymin", "xmin", "ymax", "xmax")):
            f["image/window/bbox/" + k] = wb[:, i].astype(np.float32)
        f["image/window/labels/text"] = [_fmt_row(r).encode() for r in np.asarray(example["window_classes"])]
    if "groundtruth_edgemask" in example:
        em = np.asarray(example["groundtruth_edgemask"], np.float32)
        f["image/edgemask/height"] = np.asarray([em.shape[1]], np.int64)
        f["image/edgemask/width"] = np.asarray([em.shape[2]], np.int64)
        f["image/edgemask/masks"] = em.reshape(-1)
    return serialize_example(f)


class TfRecordDataset(object):
    """Iterates decoded examples of one or more TFRecord files (the `tf_record_input_reader` of the pipeline
    config); `shuffle` permutes records inside each file with a seeded generator."""

    def __init__(self, paths, num_classes, shuffle=False, seed=0):
        self.paths = [paths] if isinstance(paths, str) else list(paths)
        self.num_classes = num_classes
        self.shuffle, self.seed = shuffle, seed

    def __iter__(self):
        rng = np.random.default_rng(self.seed)
        for p in self.paths:
            recs = list(read_tfrecords(p))
            order = rng.permutation(len(recs)) if self.shuffle else range(len(recs))
            for i in order:
                yield decode_example(recs[i], self.num_classes)
