import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when collected on a box without a device."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def golden_prim():
    return load_golden("primitives")


@pytest.fixture(scope="session")
def golden_fs():
    return load_golden("fields_solvers")


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    n = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (n if n > 0 else 1.0)
