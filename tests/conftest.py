import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def stages():
    return np.load(os.path.join(ROOT, "tests", "golden", "stages.npz"))


@pytest.fixture(scope="session")
def e2e_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "e2e.npz"))


@pytest.fixture(scope="session")
def prosac_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "prosac.npz"))


@pytest.fixture(scope="session")
def extra_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "extra.npz"))


@pytest.fixture(scope="session")
def port():
    from oracle import port as p
    p.build()
    return p


@pytest.fixture(scope="session")
def ref():
    """The reference binary (only where oracle/_ref exists)."""
    from oracle import build_ref, ref_wheel
    if not build_ref.build():
        pytest.skip("reference wheel not available on this box")
    return ref_wheel


@pytest.fixture(scope="session")
def ctx():
    """GPU context through the C ABI; fails (not skips) when the CUDA library cannot run."""
    from mdrp_b200 import _native as nv
    return nv.Context(0)
