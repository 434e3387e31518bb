import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def fpa():
    """The product package, with libfpa_b200.so built in-tree."""
    entry.build()
    return entry.load_package()


@pytest.fixture(scope="session")
def gpu(fpa):
    """The package on a box with a CUDA device; GPU tests never fall back to anything."""
    if fpa._lib.device_count() < 1:
        pytest.fail("a test marked gpu ran without a CUDA device (no CPU fallback exists)")
    return fpa


@pytest.fixture(scope="session")
def golden():
    return dict(np.load(ROOT / "tests" / "golden" / "reference_golden.npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import fwm_oracle
    return fwm_oracle


@pytest.fixture(scope="session")
def nw_oracle():
    from oracle import nwave_oracle
    return nwave_oracle


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), 1e-300)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0
