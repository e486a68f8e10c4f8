import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def hk():
    """The CUDA product library through its C-ABI. Fails loudly if it is not built / no device."""
    from hierarchicalkarting_b200 import abi
    lib = abi.load_library()
    abi.check(lib.hk_init(int(os.environ.get("LOCAL_RANK", "0"))))
    return lib


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


def rel_err(got, ref):
    """SURVEY.md A.7 parity metric: |got-ref| <= tol * max(|ref|, ||ref||_inf of the same output tensor)."""
    got, ref = np.asarray(got, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    scale = np.maximum(np.abs(ref), np.max(np.abs(ref)) if ref.size else 1.0)
    scale = np.where(scale == 0, 1.0, scale)
    return float(np.max(np.abs(got - ref) / scale)) if ref.size else 0.0
