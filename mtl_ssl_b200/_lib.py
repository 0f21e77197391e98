"""ctypes loader for the C-ABI CUDA library (include/mtlssl.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception
is raised (MtlError).  The oracle under oracle/ is test infrastructure and is never
imported from here.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmtlssl.so")


class MtlError(RuntimeError):
    pass


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MtlError(
                "CUDA extension %s not built; run `python -m mtl_ssl_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mtl_last_error_string.restype = ctypes.c_char_p
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().mtl_last_error_string().decode("utf-8", "replace")
        raise MtlError("%s failed (rc=%d): %s" % (what, rc, msg))


def ptr(t):
    """Raw device pointer of a torch tensor (or None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
