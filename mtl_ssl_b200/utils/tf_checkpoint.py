"""TensorFlow checkpoint (V2 "tensor bundle") reader without TensorFlow (host side): what
`tf.train.Saver.restore` / `slim.assign_from_checkpoint` read when the reference starts from an ImageNet or detection
checkpoint (/root/reference/object_detection/trainer.py:311-356, meta_architectures/faster_rcnn_meta_arch.py:1947-2013).

Format (tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc}, tensorflow/core/lib/io/table*.cc -- the LevelDB table):
  <prefix>.index                  an SSTable: key "" -> BundleHeaderProto {num_shards=1, endianness=2, version=3},
                                  key <variable name> -> BundleEntryProto {dtype=1, shape=2, shard_id=3, offset=4,
                                  size=5, crc32c=6 (masked CRC-32C of the bytes), slices=7}
  <prefix>.data-0000N-of-0000M    raw little-endian tensor bytes
  SSTable: data blocks | metaindex block | index block | footer (two BlockHandles, padded to 40 bytes, magic
  0xdb4775248b80fb57); a block is prefix-compressed entries (shared, non_shared, value_len varint32s + key suffix +
  value), a restart array and its length (uint32s), followed on disk by a 1-byte compression tag and a masked CRC-32C.

**Parity unpinned**: no TensorFlow and no checkpoint file exist in this environment; the reader is pinned only against
the writer below (same specification) and the CRC / protobuf pieces that are cross-checked elsewhere.  Snappy-compressed
blocks (tag 1) raise NotImplementedError (TensorFlow writes bundle indices uncompressed).

Layout conversion: TF conv kernels are HWIO and FC kernels [in, out]; this framework stores [K, R, S, C] / [out, in]
(ops_conv.py), depthwise kernels [R, S, C, 1] -> [C, R, S, 1]."""
import os
import struct

import numpy as np

from ..data.tfrecord import _enc_varint, _fields, _ld, _varint, masked_crc32c

_MAGIC = 0xDB4775248B80FB57
# tensorflow/core/framework/types.proto
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 4: np.uint8, 5: np.int16, 6: np.int8, 9: np.int64, 10: np.bool_,
           17: np.uint16, 19: np.float16, 22: np.uint32, 23: np.uint64}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


# ----------------------------------------------------------------------------- SSTable
def _read_block(buf, offset, size, check_crc=True):
    contents = buf[offset:offset + size]
    trailer = buf[offset + size:offset + size + 5]
    if len(contents) != size or len(trailer) != 5:
        raise ValueError("truncated table block")
    if check_crc and struct.unpack("<I", trailer[1:])[0] != masked_crc32c(bytes(contents) + bytes(trailer[:1])):
        raise ValueError("table block checksum mismatch")
    if trailer[0] == 1:
        raise NotImplementedError("snappy-compressed table block")
    if trailer[0] != 0:
        raise ValueError("unknown block compression tag %d" % trailer[0])
    return contents


def _block_entries(block):
    """(key, value) pairs of one block, undoing the prefix compression."""
    block = memoryview(block)
    (num_restarts,) = struct.unpack("<I", block[-4:])
    end = len(block) - 4 - 4 * num_restarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        yield key, bytes(block[pos:pos + vlen])
        pos += vlen


def _handle(buf, pos=0):
    off, pos = _varint(buf, pos)
    size, pos = _varint(buf, pos)
    return off, size, pos


def read_table(path, check_crc=True):
    """All (key, value) pairs of a LevelDB-format table file, in key order."""
    buf = memoryview(open(path, "rb").read())
    if len(buf) < 48 or struct.unpack("<Q", buf[-8:])[0] != _MAGIC:
        raise ValueError("%s: not an SSTable (bad magic)" % path)
    footer = buf[-48:]
    _, _, p = _handle(footer)                       # metaindex handle (unused)
    ioff, isize, _ = _handle(footer, p)
    out = []
    for _, hv in _block_entries(_read_block(buf, ioff, isize, check_crc)):
        off, size, _ = _handle(memoryview(hv))
        out.extend(_block_entries(_read_block(buf, off, size, check_crc)))
    return out


def _build_block(entries, restart_interval=16):
    out, restarts, prev = bytearray(), [], b""
    for i, (k, v) in enumerate(entries):
        shared = 0
        if i % restart_interval == 0:
            restarts.append(len(out))
        else:
            while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                shared += 1
        out += _enc_varint(shared) + _enc_varint(len(k) - shared) + _enc_varint(len(v)) + k[shared:] + v
        prev = k
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def write_table(path, items, block_size=4096):
    """items: iterable of (key bytes, value bytes) in strictly increasing key order."""
    items = list(items)
    if any(items[i][0] >= items[i + 1][0] for i in range(len(items) - 1)):
        raise ValueError("table keys must be strictly increasing")
    f = bytearray()

    def emit(block):
        off = len(f)
        f.extend(block)
        f.extend(b"\x00" + struct.pack("<I", masked_crc32c(block + b"\x00")))
        return _enc_varint(off) + _enc_varint(len(block))
    index, cur, cur_size = [], [], 0
    for k, v in items:
        cur.append((k, v))
        cur_size += len(k) + len(v) + 3
        if cur_size >= block_size:
            index.append((cur[-1][0], emit(_build_block(cur))))
            cur, cur_size = [], 0
    if cur:
        index.append((cur[-1][0], emit(_build_block(cur))))
    meta = emit(_build_block([]))
    idx = emit(_build_block(index, restart_interval=1))
    footer = meta + idx
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", _MAGIC)
    f.extend(footer)
    with open(path, "wb") as fh:
        fh.write(bytes(f))


# ----------------------------------------------------------------------------- tensor bundle
def _parse_entry(value):
    e = dict(dtype=0, shape=[], shard_id=0, offset=0, size=0, crc32c=None, sliced=False)
    for num, wt, v in _fields(value):
        if num == 1 and wt == 0:
            e["dtype"] = v
        elif num == 2 and wt == 2:
            for n2, w2, dim in _fields(v):                      # TensorShapeProto.dim
                if n2 == 2 and w2 == 2:
                    size = 0
                    for n3, w3, x in _fields(dim):
                        if n3 == 1 and w3 == 0:
                            size = x
                    e["shape"].append(size)
        elif num == 3 and wt == 0:
            e["shard_id"] = v
        elif num == 4 and wt == 0:
            e["offset"] = v
        elif num == 5 and wt == 0:
            e["size"] = v
        elif num == 6 and wt == 5:
            e["crc32c"] = struct.unpack("<I", v)[0]
        elif num == 7:
            e["sliced"] = True
    return e


class CheckpointReader(object):
    """`tf.train.NewCheckpointReader`-like access: get_variable_to_shape_map(), has_tensor(), get_tensor()."""

    def __init__(self, prefix, check_crc=True):
        self.prefix = prefix
        self.check_crc = check_crc
        self.entries = {}
        self.num_shards = 1
        for k, v in read_table(prefix + ".index", check_crc):
            if k == b"":
                for num, wt, x in _fields(v):
                    if num == 1 and wt == 0:
                        self.num_shards = x
                    elif num == 2 and wt == 0 and x != 0:
                        raise NotImplementedError("big-endian tensor bundle")
                continue
            self.entries[k.decode("utf-8")] = _parse_entry(v)
        self._shards = {}

    def get_variable_to_shape_map(self):
        return {k: list(e["shape"]) for k, e in self.entries.items()}

    def has_tensor(self, name):
        return name in self.entries

    def get_tensor(self, name):
        e = self.entries[name]
        if e["sliced"]:
            raise NotImplementedError("partitioned variable %s" % name)
        if e["dtype"] not in _DTYPES:
            raise NotImplementedError("dtype %d of %s" % (e["dtype"], name))
        sid = e["shard_id"]
        if sid not in self._shards:
            path = "%s.data-%05d-of-%05d" % (self.prefix, sid, self.num_shards)
            self._shards[sid] = np.memmap(path, dtype=np.uint8, mode="r") if os.path.getsize(path) else np.zeros(0, np.uint8)
        raw = np.asarray(self._shards[sid][e["offset"]:e["offset"] + e["size"]])
        dt = np.dtype(_DTYPES[e["dtype"]])
        n = int(np.prod(e["shape"])) if e["shape"] else 1
        if raw.size != e["size"] or n * dt.itemsize != e["size"]:
            raise ValueError("%s: data file does not hold %d bytes at offset %d" % (name, e["size"], e["offset"]))
        if self.check_crc and e["crc32c"] is not None and masked_crc32c(raw.tobytes()) != e["crc32c"]:
            raise ValueError("%s: tensor checksum mismatch" % name)
        return raw.view(dt.newbyteorder("<")).astype(dt).reshape(tuple(e["shape"]))


def write_checkpoint(prefix, tensors):
    """{name: ndarray} -> <prefix>.index + <prefix>.data-00000-of-00001 (one shard, for tests and export)."""
    items = [(b"", _enc_varint((1 << 3) | 0) + _enc_varint(1) + _ld(3, _enc_varint((1 << 3) | 0) + _enc_varint(1)))]
    data = bytearray()
    for name in sorted(tensors, key=lambda s: s.encode("utf-8")):
        a = np.array(tensors[name], order="C")          # (ascontiguousarray would turn a scalar into shape [1])
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
        shape = b"".join(_ld(2, _enc_varint((1 << 3) | 0) + _enc_varint(int(d))) for d in a.shape)
        entry = (_enc_varint((1 << 3) | 0) + _enc_varint(_DTYPE_IDS[a.dtype]) + _ld(2, shape) +
                 _enc_varint((4 << 3) | 0) + _enc_varint(len(data)) + _enc_varint((5 << 3) | 0) + _enc_varint(len(raw)) +
                 _enc_varint((6 << 3) | 5) + struct.pack("<I", masked_crc32c(raw)))
        items.append((name.encode("utf-8"), entry))
        data += raw
    write_table(prefix + ".index", items)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))


# ----------------------------------------------------------------------------- layout conversion
def tf_to_native(name, value, kind=None, native_shape=None):
    """A TF variable in this framework's layout.  By default the rank decides: conv HWIO -> [K,R,S,C], depthwise
    [R,S,C,1] -> [C,R,S,1], FC [in,out] -> [out,in]; vectors (biases, batch-norm statistics) unchanged.  `kind`
    (runtime.Param.tf_kind) covers the variables this framework keeps in a GEMM-specific shape:
      "fc"                      slim.fully_connected [in, out]  -> [out, 1, 1, in]  (second-stage / aux FC heads)
      ("packed_conv", R, S, C)  conv [R, S, C, K]               -> [K, 1, 1, ld] im2col rows, R*S*C real columns in
                                (r, s, c) order followed by zero padding up to ld = native_shape[-1]."""
    v = np.asarray(value)
    if kind == "fc":
        if v.ndim != 2:
            raise ValueError("%s: a fully connected variable must be [in, out], got shape %s" % (name, v.shape))
        return np.ascontiguousarray(v.T).reshape(v.shape[1], 1, 1, v.shape[0])
    if isinstance(kind, (tuple, list)) and kind[0] == "packed_conv":
        R, S, C = (int(x) for x in kind[1:4])
        if v.ndim != 4 or tuple(v.shape[:3]) != (R, S, C):
            raise ValueError("%s: expected a [%d,%d,%d,K] conv variable, got shape %s" % (name, R, S, C, v.shape))
        K = v.shape[3]
        ld = int(native_shape[-1]) if native_shape is not None else (R * S * C + 63) // 64 * 64
        out = np.zeros((K, 1, 1, ld), v.dtype)
        out[:, 0, 0, :R * S * C] = v.transpose(3, 0, 1, 2).reshape(K, R * S * C)
        return out
    if v.ndim == 4:
        if name.endswith("depthwise_weights"):
            return np.ascontiguousarray(v.transpose(2, 0, 1, 3))
        return np.ascontiguousarray(v.transpose(3, 0, 1, 2))
    if v.ndim == 2:
        return np.ascontiguousarray(v.T)
    return v


def native_to_tf(name, value, kind=None):
    """Inverse of tf_to_native: the array a `tf.train.Saver` expects for this variable."""
    v = np.asarray(value)
    if kind == "fc":
        if v.ndim == 4 and v.shape[1] == v.shape[2] == 1:
            v = v.reshape(v.shape[0], v.shape[3])
        if v.ndim != 2:
            raise ValueError("%s: a fully connected operand must be [out,1,1,in] or [out,in], got %s" % (name, v.shape))
        return np.ascontiguousarray(v.T)
    if isinstance(kind, (tuple, list)) and kind[0] == "packed_conv":
        R, S, C = (int(x) for x in kind[1:4])
        K = v.shape[0]
        rows = v.reshape(K, -1)[:, :R * S * C]
        return np.ascontiguousarray(rows.reshape(K, R, S, C).transpose(1, 2, 3, 0))
    if v.ndim == 4:
        if name.endswith("depthwise_weights"):
            return np.ascontiguousarray(v.transpose(1, 2, 0, 3))
        return np.ascontiguousarray(v.transpose(1, 2, 3, 0))
    if v.ndim == 2:
        return np.ascontiguousarray(v.T)
    return v


def state_dict_from_checkpoint(reader, name_map, shapes=None, kinds=None):
    """name_map: {checkpoint variable name: model variable name or list of names} (FasterRCNNMetaArch.restore_map gives
    the variables; batch-norm statistics map by their scope).  Returns ({model name: float32 array in native layout},
    [checkpoint names that were missing]).  `shapes`: optional {model name: expected shape} check; `kinds`: optional
    {model name: runtime.Param.tf_kind} for the variables whose layout the rank does not determine."""
    out, missing = {}, []
    kinds = kinds or {}
    for ck, targets in name_map.items():
        if not reader.has_tensor(ck):
            missing.append(ck)
            continue
        raw = reader.get_tensor(ck)
        for t in ([targets] if isinstance(targets, str) else list(targets)):
            v = tf_to_native(ck, raw, kinds.get(t), shapes.get(t) if shapes else None).astype(np.float32)
            if shapes is not None and t in shapes and tuple(shapes[t]) != v.shape:
                raise ValueError("%s: checkpoint shape %s, model shape %s" % (t, v.shape, tuple(shapes[t])))
            out[t] = v
    return out, missing
