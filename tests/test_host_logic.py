"""CPU: host-side pieces that need no device -- the MT19937 stream, metrics, Philox reference, fan-out helpers."""
import random

import numpy as np
import pytest

import cases


def test_mt19937_matches_cpython_random():
    from pybgmm_b200 import _lib
    for seed, n in ((1, 1), (1, 2000), (12345, 624), (7, 313), (99, 5000)):
        random.seed(seed)
        want = [random.random() for _ in range(n)]
        nxt = random.random()
        random.seed(seed)
        got = _lib.mt19937_random(n)
        assert got.tolist() == want
        assert random.random() == nxt           # the interpreter's state advanced exactly as n calls would have
    random.seed(3)
    random.random()
    a = _lib.mt19937_random(10)
    b = [random.random() for _ in range(5)]
    random.seed(3)
    ref = [random.random() for _ in range(16)]
    assert a.tolist() == ref[1:11] and b == ref[11:16]


def test_tables_match_reference_formula():
    from scipy.special import gammaln
    from pybgmm_b200 import _lib
    lg, lv = _lib.make_tables(5, 100)
    n = np.concatenate([[1], np.arange(1, 5 + 100 + 2)])   # gaussian_components.py:120
    assert len(lg) == 5 + 100 + 2
    assert (lg == gammaln(n / 2.)).all() and (lv == np.log(n)).all()


def test_metrics_against_golden():
    """nmi / mi / vi / loss recorded by the reference (infopy.py:31-119, utils.py:31-49) on its own runs."""
    from pybgmm_b200.utils import utils
    from pybgmm_b200.utils.metrics import (information_variation, mutual_information,
                                           normalized_mutual_information)
    for case in cases.golden()["samplers"]:
        X, z_true = cases.gen(case["N"], case["D"], case["K_true"], case["seed"])
        z = np.array(case["z"])
        np.testing.assert_allclose(normalized_mutual_information(z_true, z), case["nmi"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(mutual_information(z_true, z), case["mi"], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(information_variation(z_true, z, base=2), case["vi"], rtol=1e-10, atol=1e-12)
        assert float(utils.cluster_loss_inertia(X, z)) == case["loss"]
        # the same metrics from a contingency table (what bgmm_contingency returns, without its unassigned column)
        _, ti = np.unique(z_true, return_inverse=True)
        K = int(z.max()) + 1
        table = np.bincount(ti * K + z, minlength=(int(ti.max()) + 1) * K).reshape(-1, K)
        assert normalized_mutual_information(table=table) == normalized_mutual_information(z_true, z)
        assert mutual_information(table=table) == mutual_information(z_true, z)
        assert information_variation(base=2, table=table) == information_variation(z_true, z, base=2)
        # and the loss from the sufficient statistics (the arithmetic of k_cluster_ssq)
        m_0, k_0, v_0, S_0 = cases.prior_for(case["D"], case["cov"], case["v_0"])
        ssq = np.zeros(K)
        for k in range(K):
            xk = X[z == k]
            num = k_0 * m_0 + xk.sum(axis=0)
            s0d = np.diag(S_0) if np.ndim(S_0) == 2 else S_0
            sdiag = (s0d + k_0 * m_0 * m_0) + np.square(xk).sum(axis=0)
            sx = num - k_0 * m_0
            ssq[k] = max(np.sum((sdiag - (s0d + k_0 * m_0 * m_0)) - sx * sx / len(xk)), 0.0)
        assert float(utils.cluster_loss_from_ssq(ssq)) == case["loss"]


def test_draw_is_the_reference_draw():
    from pybgmm_b200.utils import utils
    random.seed(4)
    p = np.array([0.1, 0.2, 0.3, 0.4])
    got = [utils.draw(p) for _ in range(200)]
    random.seed(4)
    want = []
    for _ in range(200):
        u = random.random()
        k = 3
        for i in range(4):
            u = u - p[i]
            if u < 0:
                k = i
                break
        want.append(k)
    assert got == want


def test_shard_bounds_and_label_offsets():
    from pybgmm_b200 import fanout
    for N, W in ((10, 3), (8_000_000, 8), (7, 8), (100, 1)):
        b = [fanout.shard_bounds(N, W, r) for r in range(W)]
        assert b[0][0] == 0 and b[-1][1] == N
        assert all(b[r][1] == b[r + 1][0] for r in range(W - 1))
        assert max(hi - lo for lo, hi in b) - min(hi - lo for lo, hi in b) <= 1
    z = [np.array([0, 1, 1, -1]), np.array([2, 0, 1]), np.array([0])]
    out = fanout.offset_labels(z, [2, 3, 1])
    assert out.tolist() == [0, 1, 1, -1, 4, 2, 3, 5]


def test_betabern_struct():
    """pybgmm/prior/betabern.py:8-18: fields a, b, the tag, the refusal of a negative a; SubCRPMM's starting inclusion
    probability is the prior mean (subcrpmm.py:46-48)."""
    from pybgmm_b200.prior import BetaBern
    b = BetaBern(2, 6)
    assert (b.name, b.a, b.b) == ("Beta", 2, 6)
    assert b.mean() == 0.25
    assert b.posterior(3, 1) == (5, 7)
    import pytest
    with pytest.raises(AssertionError):
        BetaBern(-1, 1)
