import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def make_data(N, D, K_true, seed=1, mean_scale=4.0, noise=0.7):
    """The demos' generator (examples/crpmm_2d_demo.py:41-55), seeded like the reference tests (test_igmm.py:21-37).
    Consumes the global `random` / `np.random` streams exactly like those scripts do."""
    import random

    import numpy as np
    random.seed(seed)
    np.random.seed(seed)
    z_true = np.random.randint(0, K_true, N)
    mu = np.random.randn(D, K_true) * mean_scale
    X = mu[:, z_true] + np.random.randn(D, N) * noise
    return np.ascontiguousarray(X.T), z_true


def make_prior(D, cov="full", v_0=None, noise=0.7, mean_scale=4.0):
    import numpy as np
    v_0 = D + 3 if v_0 is None else v_0
    m_0 = np.zeros(D)
    k_0 = noise ** 2 / mean_scale ** 2
    S_0 = noise ** 2 * v_0 * (np.eye(D) if cov == "full" else np.ones(D))
    return m_0, k_0, v_0, S_0


@pytest.fixture(scope="session")
def gpu_lib():
    from pybgmm_b200 import _lib
    if _lib.device_count() < 1:
        pytest.fail("a -m gpu test ran without a CUDA device: the product path has no CPU fallback")
    return _lib
