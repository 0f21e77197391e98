"""Builds the in-tree CUDA shared library (sm_100a only) with nvcc.

The library is a plain C-ABI .so (see include/mtlssl.h); it is loaded with ctypes by
mtl_ssl_b200._lib.  Objects are rebuilt only when their source (or a header) is newer.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmtlssl.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


# boxes.cu must round every float op like the NumPy float32 oracle: no FMA contraction.
PER_FILE_FLAGS = {"boxes.cu": ["-fmad=false"]}


def _newer(src, dst, extra):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(s) > t for s in [src] + extra)


def build(verbose=False, force=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs += [os.path.join(HERE, "..", "include", f)
             for f in os.listdir(os.path.join(HERE, "..", "include")) if f.endswith(".h")]
    jobs = []
    objs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or _newer(src, obj, hdrs):
            jobs.append(["nvcc"] + NVCC_FLAGS + PER_FILE_FLAGS.get(s, []) +
                        ["-I", os.path.join(HERE, "..", "include"), "-c", src, "-o", obj])

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if jobs or not os.path.exists(LIB):
        run(["nvcc", "-shared", "-o", LIB] + objs)
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv))
