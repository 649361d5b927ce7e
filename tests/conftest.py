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


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


@pytest.fixture(scope="session")
def g_kernels():
    return load_golden("kernels_k4.npz")


@pytest.fixture(scope="session")
def g_cavi():
    return load_golden("cavi_cfg1.npz")


@pytest.fixture(scope="session")
def g_k20():
    return load_golden("cavi_k20.npz")


@pytest.fixture(scope="session")
def g_project():
    return load_golden("project_cfg1.npz")


@pytest.fixture(scope="session")
def g_reinit():
    return load_golden("reinit_small.npz")


@pytest.fixture(scope="session")
def g_simul():
    return load_golden("simul_small.npz")


@pytest.fixture(scope="session")
def g_minibatch():
    return load_golden("minibatch_small.npz")


@pytest.fixture(scope="session")
def g_trials():
    return load_golden("trials_small.npz")


@pytest.fixture(scope="session")
def g_fp32():
    return load_golden("fp32_small.npz")


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def max_rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
