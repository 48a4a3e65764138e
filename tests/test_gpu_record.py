"""GPU: the per-sweep record of GMM.update_record_dict (gmm/gmm.py:65-118) computed on the device -- contingency table
(bgmm_contingency) and per-cluster squared distances (bgmm_cluster_ssq) -- against NumPy on the same labels, and the
record dict against the host recount and the reference's recorded values."""
import numpy as np
import pytest

import cases
from conftest import make_data, make_prior

pytestmark = pytest.mark.gpu


def _chain(gpu_lib, N, D, K_true, cov="full", K_max=None, seed=3):
    X, z_true = make_data(N, D, K_true, seed=seed)
    m_0, k_0, v_0, S_0 = make_prior(D, cov)
    ch = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, K_max or (K_true + 8), covariance_type=cov)
    return ch, X, z_true


def _table_numpy(t, z, T, K):
    col = np.where(z < 0, K, z)
    return np.bincount(t * (K + 1) + col, minlength=T * (K + 1)).reshape(T, K + 1)


@pytest.mark.parametrize("N,T,K,unassigned", [
    (10007, 7, 5, True),        # ragged tail (N % 4 = 3), unassigned column in use
    (4, 1, 1, False),           # one vector, no tail
    (3, 2, 2, False),           # tail only
    (200001, 300, 400, True),   # 300 x 401 cells: too large for shared memory -> global atomics
    (1000000, 100, 100, False),  # BASELINE.json C3 label volume; shared-memory tables, 40 KB each
])
def test_contingency_is_exact(gpu_lib, N, T, K, unassigned):
    rng = np.random.RandomState(N % 1000 + T)
    z = rng.randint(0, K, N)
    if unassigned:
        z[rng.rand(N) < 0.01] = -1
    live = z >= 0
    z[live] = np.unique(z[live], return_inverse=True)[1]   # labels must be consecutive from 0
    K = int(z.max()) + 1
    ch, X, _ = _chain(gpu_lib, N, 2, 3, K_max=K + 4)
    t = rng.randint(0, T, N) * 5 - 7                      # arbitrary integers: ranked like np.unique
    ch.set_assignments(z)
    ch.set_true_labels(t)
    got = ch.contingency()
    uniq, ti = np.unique(t, return_inverse=True)
    want = _table_numpy(ti, z, len(uniq), K)
    assert got.shape == want.shape
    np.testing.assert_array_equal(got, want)              # integer work: bit-exact
    assert got.sum() == N                                 # checksum of the table
    np.testing.assert_array_equal(got.sum(axis=0)[:K], ch.get_state(z=False)["counts"][:K])


def test_contingency_needs_true_labels(gpu_lib):
    ch, X, _ = _chain(gpu_lib, 100, 2, 3)
    ch.set_assignments(np.zeros(100, np.int64))
    with pytest.raises(gpu_lib.BgmmError) as ei:
        ch.T_true = 1
        ch.contingency()
    assert ei.value.code == gpu_lib.BGMM_EINVAL
    with pytest.raises(gpu_lib.BgmmError):
        gpu_lib._check(gpu_lib.lib().bgmm_set_true_labels(ch._h, gpu_lib._ip(np.full(100, 5, np.int64)), 3))
    with pytest.raises(ValueError):
        ch.set_true_labels(np.zeros(99, np.int64))


@pytest.mark.parametrize("cov,D", [("full", 16), ("full", 3), ("diag", 8), ("full", 64), ("diag", 1)])
def test_cluster_ssq_matches_two_pass(gpu_lib, cov, D):
    from pybgmm_b200.utils import utils
    N, K = 30000, 12
    ch, X, z_true = _chain(gpu_lib, N, D, K, cov=cov)
    z = z_true.copy()
    z[:K] = np.arange(K)
    z[-1] = K                    # a cluster of one datum: squared distance exactly 0 in the reference
    ch.set_assignments(z)
    got = ch.cluster_ssq()
    want = np.array([np.sum(np.square(X[z == k] - X[z == k].mean(axis=0))) for k in range(K + 1)])
    np.testing.assert_allclose(got[:K], want[:K], rtol=1e-9)         # fp64 tolerance of north_star
    assert 0.0 <= got[K] < 1e-9
    assert utils.cluster_loss_from_ssq(got) == utils.cluster_loss_inertia(X, z)


@pytest.mark.parametrize("cls,cov,D", [("CRPMM", "full", 2), ("PCRPMM", "full", 16), ("CRPMM", "diag", 4)])
def test_record_dict_device_equals_host_recount(gpu_lib, cls, cov, D):
    """Every sweep's nmi / mi / vi / loss: device table vs recount from the labels on the host, same chain."""
    import random

    import pybgmm_b200 as P
    recs = {}
    for backend in ("device", "host"):
        X, z_true = make_data(3000, D, 6, seed=5)
        m_0, k_0, v_0, S_0 = make_prior(D, cov)
        model = getattr(P, cls)(X, P.NIW(m_0, k_0, v_0, S_0), 1.0, None, assignments="rand", K=8, K_max=64,
                                covariance_type=cov)
        model.metrics_backend = backend
        recs[backend], _ = model.collapsed_gibbs_sampler(6, z_true, num_saved=0)
    for key in ("nmi", "mi", "vi"):
        np.testing.assert_allclose(recs["device"][key], recs["host"][key], rtol=1e-12, atol=1e-14)
    assert [float(v) for v in recs["device"]["loss"]] == [float(v) for v in recs["host"]["loss"]]
    assert recs["device"]["components"] == recs["host"]["components"]
    assert recs["device"]["nk"] == recs["host"]["nk"]


def test_record_with_unassigned_data_falls_back_to_the_label_count(gpu_lib):
    """one-by-one start, metrics asked before any sweep: the -1 labels are a cluster of their own in the reference's
    metrics (np.unique / set() over the labels), which the device table reports in its last column."""
    import pybgmm_b200 as P
    from pybgmm_b200.utils.metrics import normalized_mutual_information
    X, z_true = make_data(500, 2, 4, seed=2)
    m_0, k_0, v_0, S_0 = make_prior(2)
    model = P.CRPMM(X, P.NIW(m_0, k_0, v_0, S_0), 1.0, None, assignments="one-by-one", K_max=32)
    table = model.components.contingency(z_true)
    assert table[:, -1].sum() == 499 and table[:, :-1].sum() == 1
    got = model._clustering_metrics(z_true)
    assert got["nmi"] == normalized_mutual_information(z_true, model.components.assignments)
