"""Generator of tests/golden/tf_bundle_fixture.{index,data-00000-of-00001}: a TensorFlow V2 checkpoint ("tensor
bundle") assembled from the published format WITHOUT any code of mtl_ssl_b200/utils/tf_checkpoint.py:

  * BundleHeaderProto / BundleEntryProto / TensorShapeProto / VersionDef (tensorflow/core/protobuf/tensor_bundle.proto,
    framework/tensor_shape.proto, framework/versions.proto) are declared to the google.protobuf RUNTIME and serialised
    by it (the reader under test parses the wire format with its own varint code);
  * the .index file is a LevelDB table (tensorflow/core/lib/io/table_builder.cc, format.cc, block_builder.cc) written
    by the small builder below: prefix-compressed keys with a restart point every 16 entries (the builder's default
    `block_restart_interval`), a new data block whenever one exceeds 256 bytes (so the index block holds several
    handles and shortened separator keys, as table_builder.cc's FindShortestSeparator produces), an empty metaindex
    block, the 48-byte footer with the magic number;
  * every block trailer and every tensor carries a masked CRC-32C computed bit by bit here (crc32c.h: rotate right 15,
    add 0xa282ead8), not with the table-driven / lane-parallel routine the reader uses.

TensorFlow itself is not available in this environment, so this is still a restatement of the format -- but an
independent one: a disagreement between reader and fixture means one of the two misreads the specification.
Variables: names, dtypes and shapes a slim ResNet detection checkpoint holds (HWIO conv kernel, FC [in, out], batch-norm
vectors, int64 global_step scalar, a Momentum slot).  Run: python tests/golden/make_bundle_fixture.py"""
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def crc32c_bitwise(data):
    crc = 0xFFFFFFFF
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 & -(crc & 1))
    return crc ^ 0xFFFFFFFF


def masked(data):
    c = crc32c_bitwise(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def messages():
    from google.protobuf import descriptor_pb2, descriptor_pool, message_factory
    fd = descriptor_pb2.FileDescriptorProto(name="mtl_bundle_fixture.proto", package="bundlefx", syntax="proto3")
    F = descriptor_pb2.FieldDescriptorProto

    def msg(name, fields):
        m = fd.message_type.add(name=name)
        for fname, num, typ, label, tname in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=label)
            if tname:
                f.type_name = ".bundlefx." + tname
    msg("Dim", [("size", 1, F.TYPE_INT64, F.LABEL_OPTIONAL, None), ("name", 2, F.TYPE_STRING, F.LABEL_OPTIONAL, None)])
    msg("TensorShapeProto", [("dim", 2, F.TYPE_MESSAGE, F.LABEL_REPEATED, "Dim"),
                             ("unknown_rank", 3, F.TYPE_BOOL, F.LABEL_OPTIONAL, None)])
    msg("VersionDef", [("producer", 1, F.TYPE_INT32, F.LABEL_OPTIONAL, None),
                       ("min_consumer", 2, F.TYPE_INT32, F.LABEL_OPTIONAL, None)])
    msg("BundleHeaderProto", [("num_shards", 1, F.TYPE_INT32, F.LABEL_OPTIONAL, None),
                              ("endianness", 2, F.TYPE_INT32, F.LABEL_OPTIONAL, None),
                              ("version", 3, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "VersionDef")])
    msg("BundleEntryProto", [("dtype", 1, F.TYPE_INT32, F.LABEL_OPTIONAL, None),
                             ("shape", 2, F.TYPE_MESSAGE, F.LABEL_OPTIONAL, "TensorShapeProto"),
                             ("shard_id", 3, F.TYPE_INT32, F.LABEL_OPTIONAL, None),
                             ("offset", 4, F.TYPE_INT64, F.LABEL_OPTIONAL, None),
                             ("size", 5, F.TYPE_INT64, F.LABEL_OPTIONAL, None),
                             ("crc32c", 6, F.TYPE_FIXED32, F.LABEL_OPTIONAL, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("bundlefx." + n))
    return get("BundleHeaderProto"), get("BundleEntryProto")


class BlockBuilder(object):
    def __init__(self, restart_interval=16):
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""
        self.interval = restart_interval

    def add(self, key, value):
        shared = 0
        if self.count < self.interval:
            while shared < min(len(key), len(self.last)) and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += varint(shared) + varint(len(key) - shared) + varint(len(value)) + key[shared:] + value
        self.last = key
        self.count += 1

    def finish(self):
        out = bytes(self.buf)
        for r in self.restarts:
            out += struct.pack("<I", r)
        return out + struct.pack("<I", len(self.restarts))

    def empty(self):
        return not self.buf


def shortest_separator(a, b):
    """table_builder.cc -> BytewiseComparator::FindShortestSeparator: a key k with a <= k < b, as short as possible."""
    n = 0
    while n < min(len(a), len(b)) and a[n] == b[n]:
        n += 1
    if n < min(len(a), len(b)) and a[n] < 0xFF and a[n] + 1 < b[n]:
        return a[:n] + bytes([a[n] + 1])
    return a


def write_table(path, items, block_size=256):
    out = bytearray()
    index = BlockBuilder(restart_interval=1)
    pending = None                   # (last key of the finished block, its handle)

    def emit(block_bytes):
        off = len(out)
        out.extend(block_bytes)
        out.extend(b"\x00" + struct.pack("<I", masked(block_bytes + b"\x00")))
        return varint(off) + varint(len(block_bytes))

    cur = BlockBuilder()
    for key, value in items:
        if pending is not None:
            index.add(shortest_separator(pending[0], key), pending[1])
            pending = None
        cur.add(key, value)
        if len(cur.buf) >= block_size:
            pending = (key, emit(cur.finish()))
            cur = BlockBuilder()
    if not cur.empty():
        pending = (cur.last, emit(cur.finish()))
    if pending is not None:
        index.add(pending[0] + b"\x00" if False else pending[0], pending[1])       # FindShortSuccessor is optional
    meta_handle = emit(BlockBuilder().finish())
    index_handle = emit(index.finish())
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", 0xDB4775248B80FB57)
    out.extend(footer)
    open(path, "wb").write(bytes(out))


DT = {np.dtype(np.float32): 1, np.dtype(np.int32): 3, np.dtype(np.int64): 9}


def tensors():
    rng = np.random.default_rng(20261017)
    f = lambda *s: rng.standard_normal(s).astype(np.float32)
    scope = "FirstStageFeatureExtractor/resnet_v1_101/block3/unit_7/bottleneck_v1/conv2"
    t = {
        scope + "/weights": f(3, 3, 16, 24),                                   # HWIO
        scope + "/weights/Momentum": f(3, 3, 16, 24),
        scope + "/BatchNorm/gamma": f(24), scope + "/BatchNorm/beta": f(24),
        scope + "/BatchNorm/moving_mean": f(24), scope + "/BatchNorm/moving_variance": np.abs(f(24)) + 0.5,
        "SecondStageBoxPredictor/ClassPredictor/weights": f(40, 21),           # slim.fully_connected [in, out]
        "SecondStageBoxPredictor/ClassPredictor/biases": f(21),
        "FirstStageFeatureExtractor/MobilenetV1/Conv2d_0/weights": f(3, 3, 3, 32),
        "FirstStageFeatureExtractor/MobilenetV1/Conv2d_1_depthwise/depthwise_weights": f(3, 3, 32, 1),
        "global_step": np.asarray(123456, np.int64),
        "a/small_int_table": np.arange(7, dtype=np.int32),
    }
    for i in range(20):                                                        # enough keys for restarts + several blocks
        t["SecondStageFeatureExtractor/resnet_v1_101/block4/unit_%d/bottleneck_v1/conv1/BatchNorm/beta" % (i + 1)] = f(5)
    return t


def main():
    Header, Entry = messages()
    t = tensors()
    data = bytearray()
    items = []
    h = Header(num_shards=1, endianness=0)
    h.version.producer = 1
    items.append((b"", h.SerializeToString()))
    for name in sorted(t, key=lambda s: s.encode("utf-8")):
        a = np.array(t[name], order="C")          # (ascontiguousarray would turn the scalar into shape [1])
        raw = a.astype(a.dtype.newbyteorder("<")).tobytes()
        e = Entry(dtype=DT[a.dtype], shard_id=0, offset=len(data), size=len(raw), crc32c=masked(raw))
        for d in a.shape:
            e.shape.dim.add(size=int(d))
        items.append((name.encode("utf-8"), e.SerializeToString()))
        data += raw
    prefix = os.path.join(HERE, "tf_bundle_fixture")
    write_table(prefix + ".index", items)
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(data))
    np.savez_compressed(prefix + "_expected.npz", **{k.replace("/", "|"): v for k, v in t.items()})
    print("wrote", prefix + ".index", os.path.getsize(prefix + ".index"), "bytes;", len(items) - 1, "tensors,",
          len(data), "data bytes")


if __name__ == "__main__":
    main()
