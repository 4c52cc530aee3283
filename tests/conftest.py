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
