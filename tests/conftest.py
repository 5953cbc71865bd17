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
def stock():
    return np.load(os.path.join(GOLDEN, "stock.npz"))


@pytest.fixture(scope="session")
def literals():
    return np.load(os.path.join(GOLDEN, "reference_literals.npz"))


@pytest.fixture(scope="session")
def runs():
    return np.load(os.path.join(GOLDEN, "reference_runs.npz"))


def run_cases(runs, prefix):
    """Seeds available for a case name in reference_runs.npz, e.g. 'block100_T200' -> ['block100_T200/s0', ...]."""
    keys = sorted({k.rsplit("/", 1)[0] for k in runs.files if k.startswith(prefix + "/s")})
    return keys
