import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")


def pytest_addoption(parser):
    parser.addoption("--dry-run-gpu", action="store_true", default=False,
                     help="execute the bodies of the -m gpu tests on the CPU with stubbed kernels (tests/dryrun.py). "
                          "Use with `python -O -m pytest tests -m gpu --assert=plain --dry-run-gpu`: -O strips the "
                          "numerical asserts, what remains are Python-level errors (wrong arguments, shapes, keys).")


@pytest.fixture(autouse=True)
def _dry_run_gpu(request, monkeypatch):
    if not request.config.getoption("--dry-run-gpu") or "gpu" not in request.keywords:
        yield
        return
    import numpy as np
    import torch
    import dryrun
    import mtl_ssl_b200.builders.model_builder as mb
    dryrun.install(monkeypatch)
    real_build = mb.build
    monkeypatch.setattr(mb, "build", lambda cfg, tr, device=None, seed=0: real_build(cfg, tr, device="cpu", seed=seed))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    real_to = torch.Tensor.to
    on_cpu = lambda v: "cpu" if (isinstance(v, (str, torch.device)) and str(v).startswith("cuda")) else v
    monkeypatch.setattr(torch.Tensor, "to", lambda self, *a, **k: real_to(
        self, *[on_cpu(v) for v in a], **{kk: on_cpu(v) for kk, v in k.items()}))
    for mod, names in ((np.testing, ("assert_allclose", "assert_array_equal")), (torch.testing, ("assert_close",))):
        for n in names:
            monkeypatch.setattr(mod, n, lambda *a, **k: None)
    for n in ("empty", "zeros", "ones", "full", "tensor", "arange", "randn", "rand", "empty_like", "zeros_like",
              "randint", "linspace"):
        real = getattr(torch, n)
        monkeypatch.setattr(torch, n, (lambda real: lambda *a, **k: real(
            *a, **{kk: ("cpu" if kk == "device" and str(v).startswith("cuda") else v) for kk, v in k.items()}))(real))
    yield


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has = torch.cuda.is_available()
    except Exception:
        has = False
    if has or config.getoption("--dry-run-gpu"):
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
