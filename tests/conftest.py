import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "dnn-based-speech-enhancement-in-the-frequency-domain_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "dccrn_golden.npz"), allow_pickle=False)


@pytest.fixture(params=[0, 1], ids=["fp32", "tf32"])
def engine(request):
    """GEMM engine under test: 0 = fp32 CUDA cores (exact), 1 = tcgen05 TF32 tensor cores (the product default)."""
    from sefd import _lib
    lib = _lib.load()
    lib.sefd_set_engine(request.param)
    yield request.param
    lib.sefd_set_engine(1)


def pytest_terminal_summary(terminalreporter):
    """Report CUDA errors that were pending from outside the library when one of its launches began (include/sefd.h:
    sefd_stale_cuda_errors): absorbed instead of failing an unrelated test, but never silently."""
    mod = sys.modules.get("sefd._lib")
    lib = getattr(mod, "_lib", None) if mod else None
    try:
        n = lib.sefd_stale_cuda_errors() if lib is not None else 0
    except Exception:
        n = 0
    if n:
        terminalreporter.write_line(f"sefd: {n} stale CUDA error(s) absorbed; last: {lib.sefd_last_stale_cuda_error().decode()}")
