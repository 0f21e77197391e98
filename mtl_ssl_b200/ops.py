"""ctypes bindings of every non-GEMM entry point of include/mtlssl.h.

torch tensors are used only as device-memory containers: each wrapper passes raw pointers,
sizes and the current CUDA stream to libmtlssl.so.  Nothing here computes on the host and
there is no fallback path: a missing library or a failing call raises MtlError.

Signature mini-language (one char per C parameter, `stream` appended automatically):
  p pointer (torch tensor / None / int address)   i int   l long long   f float
  h host float array (sequence of python floats)
"""
import ctypes

import torch

from ._lib import lib, check, MtlError

def _parse_header():
    """Derive every signature from include/mtlssl.h so that header and bindings cannot drift."""
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "include", "mtlssl.h")
    text = open(path).read()
    sigs = {}
    for m in re.finditer(r"\bint\s+(mtl_\w+)\s*\(([^;]*?)\)\s*;", text, re.S):
        name, params = m.group(1), m.group(2)
        if name == "mtl_conv_tc" or params.strip() == "void":
            continue
        sig = ""
        # split on commas that are outside comments
        parts, depth, cur, i = [], 0, "", 0
        while i < len(params):
            if params.startswith("/*", i):
                j = params.index("*/", i) + 2
                cur += params[i:j]
                i = j
                continue
            if params[i] == ",":
                parts.append(cur)
                cur = ""
            else:
                cur += params[i]
            i += 1
        parts.append(cur)
        for part in parts:
            host = "host" in part
            decl = re.sub(r"/\*.*?\*/", "", part, flags=re.S).strip()
            if "mtl_stream_t" in decl:
                continue
            if "*" in decl:
                sig += "h" if (host and "float" in decl) else "p"
            elif "long long" in decl:
                sig += "l"
            elif decl.startswith("float"):
                sig += "f"
            elif decl.startswith("int"):
                sig += "i"
            else:
                raise ValueError("mtlssl.h: cannot parse parameter %r of %s" % (part, name))
        sigs[name] = sig
    return sigs


_SIGS = _parse_header()
# 'I' is an alias of 'i' kept for readability of long signatures
_CT = {"p": ctypes.c_void_p, "i": ctypes.c_int, "I": ctypes.c_int, "l": ctypes.c_longlong,
       "f": ctypes.c_float, "h": ctypes.c_void_p}
_bound = {}


def _fn(name):
    f = _bound.get(name)
    if f is None:
        f = getattr(lib(), name)
        f.argtypes = [_CT[c] for c in _SIGS[name]] + [ctypes.c_void_p]
        f.restype = ctypes.c_int
        _bound[name] = f
    return f


def _conv(c, v):
    if c == "p":
        if v is None:
            return None
        if isinstance(v, torch.Tensor):
            if not v.is_cuda and isinstance(lib(), ctypes.CDLL):
                # a host pointer handed to a kernel is an illegal address on the device: refuse it here, loudly
                raise MtlError("host tensor (shape %s) passed to a device kernel" % (tuple(v.shape),))
            return v.data_ptr()
        return int(v)
    if c in "iIl":
        return int(v)
    if c == "f":
        return float(v)
    if c == "h":
        if v is None:
            return None
        arr = (ctypes.c_float * len(v))(*[float(x) for x in v])
        return arr
    raise ValueError(c)


# tools/timeline.py sets this to a list: (kernel name, stream handle, start event, end event, 0.0) per call
TIMELINE = None


def call(name, *args, stream=None):
    if TIMELINE is not None and stream is None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s = torch.cuda.current_stream()
        e0.record(s)
        _call(name, args, s.cuda_stream)
        e1.record(s)
        TIMELINE.append((name, s.cuda_stream, e0, e1, 0.0))
        return
    _call(name, args, stream)


def _call(name, args, stream):
    sig = _SIGS[name]
    if len(args) != len(sig):
        raise TypeError("%s expects %d arguments, got %d" % (name, len(sig), len(args)))
    keep = []
    cargs = []
    for c, v in zip(sig, args):
        cv = _conv(c, v)
        if c == "h" and cv is not None:
            keep.append(cv)
            cv = ctypes.cast(cv, ctypes.c_void_p)
        cargs.append(cv)
    if stream is None:
        stream = torch.cuda.current_stream().cuda_stream
    rc = _fn(name)(*cargs, ctypes.c_void_p(stream))
    check(rc, name)


class TensorDesc(ctypes.Structure):
    _fields_ = [("offset", ctypes.c_longlong), ("numel", ctypes.c_longlong), ("row_len", ctypes.c_longlong),
                ("scale_off", ctypes.c_longlong), ("l2_weight", ctypes.c_float), ("grad_mult", ctypes.c_float),
                ("trainable", ctypes.c_int), ("pad_", ctypes.c_int)]


class ChunkDesc(ctypes.Structure):
    _fields_ = [("tensor", ctypes.c_int), ("len", ctypes.c_int), ("start", ctypes.c_longlong)]


def opt_chunk_size():
    return int(lib().mtl_opt_chunk_size())


def launch_count():
    f = lib().mtl_launch_count
    f.restype = ctypes.c_longlong
    return int(f())


def exported_symbols():
    """Names this module binds (used by the CPU test that checks the .so exports them all)."""
    return sorted(_SIGS) + ["mtl_conv_tc", "mtl_conv_tc_ws_bytes", "mtl_last_error_string", "mtl_abi_version",
                            "mtl_device_sm_count", "mtl_opt_chunk_size", "mtl_launch_count", "mtl_conv_tc_group_entry_bytes",
                            "mtl_conv_tc_group_key"]
