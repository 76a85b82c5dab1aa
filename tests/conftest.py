import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def tiny_csr(rows, m=None):
    """rows: list of lists of (col, val). -> ptr, idx, val numpy arrays."""
    ptr = [0]
    idx, val = [], []
    for r in rows:
        for c, v in r:
            idx.append(c)
            val.append(v)
        ptr.append(len(idx))
    return (np.asarray(ptr, np.int32), np.asarray(idx, np.int32), np.asarray(val, np.float32))


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
