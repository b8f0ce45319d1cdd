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
def golden():
    def _load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return _load


def csr_from(d, prefix, shape=None):
    import scipy.sparse as sp
    n = len(d[f"{prefix}_indptr"]) - 1
    data = d[f"{prefix}_data"] if f"{prefix}_data" in d.files else np.ones(len(d[f"{prefix}_indices"]))
    return sp.csr_matrix((data, d[f"{prefix}_indices"], d[f"{prefix}_indptr"]), shape=shape or (n, n))
