"""pytest configuration: `-m gpu` tests need a CUDA device; everything else
runs on the CPU-only build container."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["test_mtx", "rand_wide", "rand_tall", "rand_square", "rand_sparse_rows", "long_rows", "one_row", "empty", "grid5"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def unhex(v):
    return np.array([float.fromhex(t) for t in v], dtype=np.float64)


def load_golden(name):
    with open(os.path.join(GOLDEN_DIR, name + ".json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library; building is __graft_entry__.build()'s job,
    but a missing .so is built here so a fresh checkout can run the tests."""
    import ellspmv_b200
    if not os.path.exists(ellspmv_b200.LIB_PATH):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "ellspmv_b200", "csrc"), "-j8"], check=True,
                       stdout=subprocess.DEVNULL)
    return ellspmv_b200.load_library()


def bits_equal(a, b):
    a = np.ascontiguousarray(a, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and np.array_equal(a.view(np.uint64), b.view(np.uint64))
