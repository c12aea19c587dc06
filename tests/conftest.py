import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def shipped():
    """The reference's own fixture inputs + shipped golden outputs (tests/golden/make_golden.py)."""
    return dict(np.load(os.path.join(GOLDEN, "fixture_shipped.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def live():
    """Outputs of the live reference frozen by tests/golden/make_golden.py."""
    return dict(np.load(os.path.join(GOLDEN, "live_solver.npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def live_drivers():
    return dict(np.load(os.path.join(GOLDEN, "live_drivers.npz"), allow_pickle=False))
