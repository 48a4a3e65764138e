"""Host utilities with the reference's names (pybgmm/utils/utils.py)."""
import random

import numpy as np


def draw(p_k):
    """Inverse-CDF draw with one random.random() and sequential subtraction (pybgmm/utils/utils.py:7-20).
    Kept for API parity; inside a sweep the engine performs the same draw on the device from uniforms that the
    host takes from this same `random` stream (see _lib.mt19937_random)."""
    k_uni = random.random()
    for i in range(len(p_k)):
        k_uni = k_uni - p_k[i]
        if k_uni < 0:
            return i
    return len(p_k) - 1


def cluster_loss_inertia(x, assignments):
    """Sum over clusters of sqrt(sum of squared distances to the cluster mean), each term truncated to an
    integer exactly as the reference does by storing it in an int array (pybgmm/utils/utils.py:31-49)."""
    x = np.asarray(x, dtype=np.float64)
    z = np.asarray(assignments).ravel()
    uniq, inv = np.unique(z, return_inverse=True)
    cnt = np.bincount(inv).astype(np.float64)
    ssq = np.zeros(len(uniq))
    for d in range(x.shape[1]):
        mean = np.bincount(inv, weights=x[:, d]) / cnt
        ssq += np.bincount(inv, weights=np.square(x[:, d] - mean[inv]))
    unique_dist = np.zeros_like(uniq)
    unique_dist[:] = np.sqrt(ssq).astype(unique_dist.dtype)
    return np.sum(unique_dist)


def cluster_loss_from_ssq(ssq):
    """The same loss from the per-cluster sums of squared distances to the cluster mean (`bgmm_cluster_ssq`):
    the square root of each, truncated to an integer like the reference's int array does, summed."""
    return np.sum(np.sqrt(np.asarray(ssq, dtype=np.float64)).astype(np.int64))
