"""GPU parity proper: the CUDA engine (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): assignments bit-exact; log-likelihoods within 1e-9 relative; sufficient
statistics (sums) bit-identical because they are accumulated in the reference's operation order.
"""
import random

import numpy as np
import pytest

from conftest import make_data, make_prior
from oracle import oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-9  # the tolerance north_star states for predictive log-likelihoods


def _pair(gpu_lib, N, D, K_true, cov, K_init, K_max=None, seed=1, init="rand"):
    X, z_true = make_data(N, D, K_true, seed)
    m_0, k_0, v_0, S_0 = make_prior(D, cov)
    K_max = K_max or min(N, 4 * K_true + 64)
    z0 = O.init_assignments(N, init, K_init)
    orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=K_max, covariance_type=cov)
    orc.set_assignments(z0)
    ch = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, K_max, covariance_type=cov)
    ch.set_assignments(z0)
    return X, orc, ch


def _assert_state_equal(orc, ch, check_inv=True):
    st = ch.get_state()
    K = orc.K
    assert st["K"] == K
    np.testing.assert_array_equal(st["z"], orc.assignments)
    np.testing.assert_array_equal(st["counts"], orc.counts)
    # sums: same operations in the same order -> identical bits
    np.testing.assert_array_equal(st["m_num"], orc.m_N_numerators)
    np.testing.assert_array_equal(st["S_part"], orc.S_N_partials)
    np.testing.assert_allclose(st["logdet"][:K], orc.logdet_covars[:K], rtol=RTOL, atol=1e-11)
    if check_inv:
        np.testing.assert_allclose(st["inv_covar"][:K], orc.inv_covars[:K], rtol=1e-8, atol=1e-10)


@pytest.mark.parametrize("cov", ["full", "diag"])
@pytest.mark.parametrize("D", [1, 2, 3, 8, 16])
def test_build_and_log_post_pred(gpu_lib, cov, D):
    N = 400
    X, orc, ch = _pair(gpu_lib, N, D, 5, cov, K_init=6)
    _assert_state_equal(orc, ch)
    np.testing.assert_allclose(ch.log_prior(), orc.cached_log_prior, rtol=RTOL)
    idx = np.arange(0, N, 7)
    got = ch.log_post_pred(idx)
    want = np.stack([orc.log_post_pred(i) for i in idx])
    np.testing.assert_allclose(got, want, rtol=RTOL)
    np.testing.assert_allclose(ch.log_marg_k(), [orc.log_marg_k(k) for k in range(orc.K)], rtol=RTOL)
    np.testing.assert_allclose(ch.log_marg(1.3), orc.log_marg(1.3), rtol=RTOL)


@pytest.mark.parametrize("engine", ["sequential", "windows", "adaptive"])
@pytest.mark.parametrize("cov,D", [("full", 2), ("full", 16), ("diag", 1), ("diag", 8), ("full", 5)])
def test_crp_sweeps_match_oracle(gpu_lib, engine, cov, D):
    N, K_true, sweeps = 1500, 6, 6
    X, orc, ch = _pair(gpu_lib, N, D, K_true, cov, K_init=K_true)
    ch.set_engine(engine)
    random.seed(11)
    for s in range(sweeps):
        u = np.array([random.random() for _ in range(N)])
        so = orc.sweep(u, 1.0)
        sg = ch.sweep(1.0, 1.0, None, u)
        assert (sg.K, sg.moves, sg.births, sg.deaths, sg.evals) == (so.K_end, so.moves, so.births, so.deaths, so.evals), s
        np.testing.assert_array_equal(ch.assignments(), orc.assignments)
    _assert_state_equal(orc, ch)
    np.testing.assert_allclose(ch.log_marg(1.0), orc.log_marg(1.0), rtol=RTOL)


@pytest.mark.parametrize("engine", ["sequential", "adaptive"])
def test_pcrp_sweeps_match_oracle(gpu_lib, engine):
    N, D, K_true, sweeps, r = 2000, 4, 8, 6, 1.5
    X, orc, ch = _pair(gpu_lib, N, D, K_true, "full", K_init=K_true)
    ch.set_engine(engine)
    tab = O.logcount_table(N, r)
    rng = np.random.RandomState(5)
    for s in range(sweeps):
        order = rng.permutation(N)
        u = rng.random_sample(N)
        use_power = s > 0  # pcrpmm.py:105 `i_iter > power_burnin`
        so = orc.sweep(u, 1.0, order=order, logcount_tab=tab if use_power else None)
        sg = ch.sweep(1.0, r if use_power else 1.0, order, u)
        assert (sg.K, sg.moves, sg.births, sg.deaths) == (so.K_end, so.moves, so.births, so.deaths), s
        np.testing.assert_array_equal(ch.assignments(), orc.assignments)
    _assert_state_equal(orc, ch)


@pytest.mark.parametrize("init", ["each-in-own", "one-by-one"])
def test_births_deaths_and_unassigned(gpu_lib, init):
    """each-in-own: every first visit deletes a component (swap-with-last relabel, gaussian_components.py:188-205);
    one-by-one: data start unassigned (-1) and are added as they are visited (igmm.py:95-97)."""
    N, D = 120, 2
    X, orc, ch = _pair(gpu_lib, N, D, 4, "full", K_init=1, K_max=N, init=init)
    random.seed(3)
    for s in range(3):
        u = np.array([random.random() for _ in range(N)])
        so = orc.sweep(u, 1.0)
        sg = ch.sweep(1.0, 1.0, None, u)
        assert (sg.K, sg.moves, sg.births, sg.deaths) == (so.K_end, so.moves, so.births, so.deaths)
        np.testing.assert_array_equal(ch.assignments(), orc.assignments)
    _assert_state_equal(orc, ch)


def test_kmax_overflow_is_an_error(gpu_lib):
    N, D = 200, 2
    X, z_true = make_data(N, D, 8, 1, mean_scale=30.0)
    m_0, k_0, v_0, S_0 = make_prior(D)
    ch = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, K_max=2)
    ch.set_assignments(np.zeros(N, np.int64))
    with pytest.raises(gpu_lib.BgmmError) as ei:
        for _ in range(5):
            ch.sweep(50.0)
    assert ei.value.code == gpu_lib.BGMM_EKMAX


def test_add_del_item_protocol(gpu_lib):
    N, D = 60, 3
    X, orc, ch = _pair(gpu_lib, N, D, 3, "full", K_init=3, K_max=16)
    rng = np.random.RandomState(0)
    for _ in range(40):
        i = int(rng.randint(N))
        orc.del_item(i)
        ch.del_item(i)
        k = int(rng.randint(orc.K + 1))
        orc.add_item(i, k)
        ch.add_item(i, k)
    _assert_state_equal(orc, ch)


def test_philox_stream_replay(gpu_lib):
    """uniforms == NULL: device Philox stream; the oracle replays it through bgmm_get_uniforms."""
    N, D = 3000, 2
    X, orc, ch = _pair(gpu_lib, N, D, 5, "full", K_init=5)
    ch.seed(1234)
    for s in range(4):
        u = ch.get_uniforms(s)
        assert u.min() >= 0.0 and u.max() < 1.0
        orc.sweep(u, 1.0)
        ch.sweep(1.0)
        np.testing.assert_array_equal(ch.assignments(), orc.assignments)


@pytest.mark.parametrize("engine", ["generic", "generic-sequential", "generic-windows"])
def test_generic_engine_matches_oracle(gpu_lib, engine):
    """The generic (any D / any K_max) engine stays covered now that D <= 16 full covariance defaults to the
    shared-memory-resident engine."""
    N, D, K_true, sweeps = 1200, 4, 6, 4
    X, orc, ch = _pair(gpu_lib, N, D, K_true, "full", K_init=K_true)
    ch.set_engine(engine)
    random.seed(7)
    for s in range(sweeps):
        u = np.array([random.random() for _ in range(N)])
        so = orc.sweep(u, 1.0)
        sg = ch.sweep(1.0, 1.0, None, u)
        assert (sg.K, sg.moves, sg.births, sg.deaths, sg.evals) == (so.K_end, so.moves, so.births, so.deaths, so.evals), s
        np.testing.assert_array_equal(ch.assignments(), orc.assignments)
    _assert_state_equal(orc, ch)


def test_resident_capacity_overflow_hands_over_to_generic(gpu_lib):
    """More live components than the resident engine holds in shared memory: the sweep is continued by the generic
    engine from the scan position of the offending birth (bgmm_sweep_stats.generic_from), same chain as the oracle."""
    N, D = 700, 2
    X, z_true = make_data(N, D, 40, 1, mean_scale=60.0)
    m_0, k_0, v_0, S_0 = make_prior(D)
    K_max = 600
    z0 = O.init_assignments(N, "rand", 150)
    orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=K_max)
    orc.set_assignments(z0)
    ch = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, K_max)
    ch.set_assignments(z0)
    rng = np.random.RandomState(3)
    handed_over = []
    for s in range(4):
        u = rng.random_sample(N)
        so = orc.sweep(u, 1e4)
        sg = ch.sweep(1e4, 1.0, None, u)
        handed_over.append(sg.generic_from)
        assert (sg.K, sg.moves, sg.births, sg.deaths, sg.evals) == (so.K_end, so.moves, so.births, so.deaths, so.evals), s
        np.testing.assert_array_equal(ch.assignments(), orc.assignments)
    assert handed_over[0] > 0, "sweep 0 should cross the resident capacity mid-sweep"
    _assert_state_equal(orc, ch)


@pytest.mark.parametrize("engine", ["sequential", "cluster"])
def test_long_chain_drift_control(gpu_lib, engine):
    """Many rank-one record updates per component: the resident engine rebuilds a record from the bit-exact statistics
    every REFRESH_EVERY updates, the cluster step engine ends its launch when a component has taken REFRESH_CAP of them
    (records are rebuilt at every launch); the chain still follows the oracle and the statistics stay bit-identical."""
    N, D, K_true = 6000, 8, 3
    X, orc, ch = _pair(gpu_lib, N, D, K_true, "full", K_init=3, K_max=32)
    ch.set_engine(engine)
    rng = np.random.RandomState(11)
    refreshes = 0
    for s in range(3):
        u = rng.random_sample(N)
        so = orc.sweep(u, 1.0)
        sg = ch.sweep(1.0, 1.0, None, u)
        refreshes += sg.refreshes
        assert (sg.K, sg.moves, sg.births, sg.deaths, sg.evals) == (so.K_end, so.moves, so.births, so.deaths, so.evals), s
        np.testing.assert_array_equal(ch.assignments(), orc.assignments)
    assert refreshes > 0
    _assert_state_equal(orc, ch)


@pytest.mark.parametrize("D,cov", [(32, "full"), (64, "full"), (64, "diag"), (24, "full")])
def test_high_dimensional_sweeps_match_oracle(gpu_lib, D, cov):
    """D > 16 (BASELINE.json configs[3] is D = 64) runs on the generic engine: same exact parity bar."""
    N, K_true, sweeps = 500, 4, 3
    X, orc, ch = _pair(gpu_lib, N, D, K_true, cov, K_init=K_true)
    rng = np.random.RandomState(21)
    for s in range(sweeps):
        u = rng.random_sample(N)
        so = orc.sweep(u, 1.0)
        sg = ch.sweep(1.0, 1.0, None, u)
        assert (sg.K, sg.moves, sg.births, sg.deaths, sg.evals) == (so.K_end, so.moves, so.births, so.deaths, so.evals), s
        np.testing.assert_array_equal(ch.assignments(), orc.assignments)
    _assert_state_equal(orc, ch, check_inv=False)
    idx = np.arange(0, N, 11)
    np.testing.assert_allclose(ch.log_post_pred(idx), np.stack([orc.log_post_pred(i) for i in idx]), rtol=RTOL)
    np.testing.assert_allclose(ch.log_marg(1.0), orc.log_marg(1.0), rtol=RTOL)
