import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def golden(name: str):
    return dict(np.load(GOLDEN / name, allow_pickle=False))


@pytest.fixture(scope="session")
def native_lib():
    """The C-ABI library; builds it if this checkout has not been built yet."""
    from two_tower_model_v2_b200 import _native
    if not _native.LIB_PATH.exists():
        import __graft_entry__
        __graft_entry__.build()
    return _native.load()


def rel_err(a, b):
    """Norm-wise: max|a-b| / max|b| (the whole tensor's largest element is the scale)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def elem_rel_err(a, b, floor_frac=1.0):
    """Element-wise: max over elements of |a-b| / max(|b|, floor), floor = floor_frac * rms of that ROW of b: every
    element is held to the relative tolerance, except that elements below the row's typical magnitude (rms; 0.051 for
    a unit vector in D = 384) are held to tolerance * rms absolute - the reference's own fp32-vs-fp64 noise (2e-7 of
    the largest element) is already 6e-6 relative on an element at 0.1 rms, so a smaller floor would test fp32 itself."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    rms = np.sqrt((b * b).mean(axis=-1, keepdims=True))
    floor = np.maximum(floor_frac * rms, 1e-30)
    return float((np.abs(a - b) / np.maximum(np.abs(b), floor)).max())
