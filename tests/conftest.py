import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running CPU test")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def g_ssm():
    return load_golden("ssm_loglik.npz")


@pytest.fixture(scope="session")
def g_em():
    return load_golden("em_traces.npz")


@pytest.fixture(scope="session")
def g_restarts():
    return load_golden("restarts.npz")


@pytest.fixture(scope="session")
def g_pred():
    return load_golden("predictors.npz")


@pytest.fixture(scope="session")
def g_ext():
    return load_golden("extended.npz")


@pytest.fixture(scope="session")
def g_adni():
    return load_golden("adni_fold0.npz")


@pytest.fixture(scope="session")
def g_cond():
    return load_golden("conditioning.npz")
