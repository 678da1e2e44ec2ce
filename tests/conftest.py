import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def unhex_f64(s):
    return np.frombuffer(bytes.fromhex(s), dtype=np.float64).copy()


def unhex_c128(s):
    return unhex_f64(s).view(np.complex128)


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def values_equal(a, b):
    """Value identity of two complex arrays: every component compares equal as an
    IEEE double (so -0.0 == +0.0, the one representational freedom)."""
    a = np.asarray(a).view(np.float64)
    b = np.asarray(b).view(np.float64)
    return a.shape == b.shape and bool(np.all(a == b))


def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.fixture(scope="session")
def oracle_built():
    import oracle
    if not oracle.have_restatement() or (os.path.exists("/root/reference/qc_shor.c") and not oracle.have_reference()):
        oracle.build()
    return oracle


@pytest.fixture(scope="session")
def qcs():
    import quantumcomputer_b200 as q
    q.lib()
    return q
