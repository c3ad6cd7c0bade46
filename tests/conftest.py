"""pytest configuration: registers the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` runs on the CPU-only build container (oracle vs golden vectors, host logic,
C-ABI symbol checks, gloo world_size-2 sharding).  ``-m gpu`` runs the parity tests proper on a
B200 through the C-ABI; they never read ``/root/reference`` (it does not exist on the GPU box).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden_features():
    return load_golden("features.npz")


@pytest.fixture(scope="session")
def golden_blocks():
    return load_golden("blocks.npz")


@pytest.fixture(scope="session")
def golden_e2e():
    return load_golden("e2e.npz")


@pytest.fixture(scope="session")
def golden_decode():
    return load_golden("decode.npz")


@pytest.fixture(scope="session")
def golden_helpers():
    return load_golden("helpers.npz")


def rel_err(a, b):
    """(max|a-b| / max|b|, ||a-b||_2 / ||b||_2) -- the parity metrics of SURVEY.md 8(d)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = a - b
    den_max = max(float(np.abs(b).max()), 1e-30)
    den_l2 = max(float(np.sqrt((b * b).sum())), 1e-30)
    return float(np.abs(d).max()) / den_max, float(np.sqrt((d * d).sum())) / den_l2
