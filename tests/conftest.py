import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def relmax(a, b):
    """max|a-b| / max|b| -- the metric BASELINE.json's tolerance is stated in."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = np.abs(b).max() if b.size else 0.0
    if den == 0.0:
        return float(np.abs(a - b).max()) if a.size else 0.0
    return float(np.abs(a - b).max() / den)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o
    o.build()
    return o
