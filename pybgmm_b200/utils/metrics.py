"""Clustering metrics recorded once per sweep by GMM.update_record_dict (pybgmm/gmm/gmm.py:65-118).

Same definitions as pybgmm/infopy/infopy.py:19-119 (entropy, mutual information, normalised mutual information,
variation of information) but computed from one K_true x K contingency table instead of the reference's
O(K_true * K * N) Python loops, so they stay usable at N = 1e6.  Every function takes either the two label vectors
(the table is then counted here, on the host) or `table=` -- the table counted on the device by
`bgmm_contingency`, which is what the samplers' record keeping uses (the labels never leave the GPU).
"""
import math

import numpy as np


def _contingency(labels_true, labels_pred):
    lt = np.asarray(labels_true).ravel()
    lp = np.asarray(labels_pred).ravel()
    if lt.shape != lp.shape:
        raise ValueError("labels_true and labels_pred must have same size, got %d and %d" % (lt.size, lp.size))
    _, ti = np.unique(lt, return_inverse=True)
    _, pi = np.unique(lp, return_inverse=True)
    nt, npred = int(ti.max()) + 1, int(pi.max()) + 1
    table = np.bincount(ti * npred + pi, minlength=nt * npred).reshape(nt, npred)
    return table


def _entropy_counts(counts, base):
    p = counts[counts > 0] / float(counts.sum())
    return float(-(p * (np.log(p) / math.log(base))).sum())


def entropy(x, base=math.e):
    """infopy.py:19-29."""
    x = np.asarray(x).ravel()
    if len(x) == 0:
        return 1.0
    return _entropy_counts(np.unique(x, return_counts=True)[1], base)


def mutual_information(labels_true=None, labels_pred=None, normalized=False, base=math.e, table=None):
    """infopy.py:62-96 (the normaliser uses natural-log entropies whatever `base` is, as the reference does)."""
    table = _contingency(labels_true, labels_pred) if table is None else np.asarray(table)
    n = float(table.sum())
    px = table.sum(axis=1) / n
    py = table.sum(axis=0) / n
    nz = table > 0
    pxy = table[nz] / n
    outer = (px[:, None] * py[None, :])[nz]
    mi = float((pxy * (np.log(pxy / outer) / math.log(base))).sum())
    if normalized:
        h_true = _entropy_counts(table.sum(axis=1), math.e)
        h_pred = _entropy_counts(table.sum(axis=0), math.e)
        mi = mi / max(np.sqrt(h_true * h_pred), 1e-10)
    return mi


def normalized_mutual_information(labels_true=None, labels_pred=None, base=math.e, table=None):
    """infopy.py:31-60."""
    return mutual_information(labels_true, labels_pred, normalized=True, base=base, table=table)


def information_variation(labels_true=None, labels_pred=None, base=math.e, table=None):
    """infopy.py:99-119."""
    table = _contingency(labels_true, labels_pred) if table is None else np.asarray(table)
    return (_entropy_counts(table.sum(axis=1), base) + _entropy_counts(table.sum(axis=0), base)
            - 2 * mutual_information(base=base, table=table))
