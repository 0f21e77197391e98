"""Host-orchestration dry run WITHOUT a GPU: every C-ABI call keeps its argument-count / type marshalling
(`ops.call` against the signatures parsed from include/mtlssl.h, `mtl_conv_args` filling) but the library entry point
is replaced by a stub that returns 0, CUDA streams / events by inert stand-ins, and tensors live on the CPU.  Kernel
outputs are therefore meaningless (zeros); what a dry run checks is the Python side of the hot path -- buffer shapes,
dictionary keys, call sequences, lane / stream bookkeeping -- for code paths that could not be exercised on a device yet.
Test infrastructure only (the product path still fails loudly without the CUDA library, see
test_product_path_fails_loudly_without_cuda)."""
import contextlib

import torch


class FakeStream(object):
    cuda_stream = 0

    def wait_event(self, e):
        pass

    def wait_stream(self, s):
        pass

    def synchronize(self):
        pass

    def record_event(self, e=None):
        return e or FakeEvent()


class FakeEvent(object):
    def __init__(self, *a, **k):
        pass

    def record(self, stream=None):
        pass

    def wait(self, stream=None):
        pass

    def synchronize(self):
        pass

    def query(self):
        return True

    def elapsed_time(self, other):
        return 0.0


class _FakeLib(object):
    """Any symbol -> a function returning 0; the conv workspace query returns 0 bytes."""

    def __init__(self, log):
        self._log = log

    _CONST = {"mtl_opt_chunk_size": 8192, "mtl_device_sm_count": 148, "mtl_abi_version": 1, "mtl_launch_count": 0}

    def __getattr__(self, name):
        def fn(*a):
            if name in self._CONST:
                return self._CONST[name]
            self._log.append(name)
            return 0
        return fn


@contextlib.contextmanager
def _no_stream(s):
    yield


class FakeGraph(object):
    """torch.cuda.CUDAGraph stand-in: the body runs once inside `torch.cuda.graph(g)` (as the Python side of a real
    capture does), `replay()` only counts."""

    def __init__(self):
        self.replays = 0

    def replay(self):
        self.replays += 1


def install(monkeypatch):
    """Returns the list that records the name of every C-ABI entry point 'launched'."""
    from mtl_ssl_b200 import _lib, ops, ops_conv
    log = []
    fake = _FakeLib(log)
    cur = FakeStream()
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: FakeStream())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: cur)
    monkeypatch.setattr(torch.cuda, "stream", _no_stream)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "graph", _no_stream)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(_lib, "lib", lambda: fake)
    for mod in (ops, ops_conv):
        if hasattr(mod, "lib"):
            monkeypatch.setattr(mod, "lib", lambda: fake)
    monkeypatch.setattr(ops, "_bound", {})

    def _fn(name):
        def f(*cargs):
            log.append(name)
            return 0
        return f
    monkeypatch.setattr(ops, "_fn", _fn)
    # the two kernels whose OUTPUT the host reads back to size later buffers (the number of anchors inside the image)
    # are filled in from the oracle, so that buffer sizes in a dry run are the device's
    real_call = ops.call

    def call(name, *args, **kw):
        real_call(name, *args, **kw)
        if name == "mtl_grid_anchors":
            import numpy as np
            from oracle import boxes as OB
            hf, wf, scales, ns, ars, na, bh, bw, sh, sw, oh, ow, out = args
            out.copy_(torch.from_numpy(OB.grid_anchors(hf, wf, list(scales)[:ns], list(ars)[:na], (bh, bw), (sh, sw),
                                                       (oh, ow))).reshape(out.shape))
        elif name == "mtl_prune_outside_window":
            from oracle import boxes as OB
            allb, n, wy0, wx0, wy1, wx1, keep, kept, num = args
            boxes, idx = OB.prune_outside_window(allb.numpy(), (wy0, wx0, wy1, wx1))
            keep[:len(idx)] = torch.from_numpy(idx.astype("int32"))
            kept[:len(idx)] = torch.from_numpy(boxes)
            num.fill_(len(idx))
    monkeypatch.setattr(ops, "call", call)
    return log
