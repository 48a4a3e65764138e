"""GPU: the cluster step engine for dense movers at padded D <= 16 (csrc/bgmm_clu.cuh: one warp per component, the
component in registers, weights and draws exchanged through DSMEM st.async + mbarriers, no barrier in the loop; the
statistics from the move log).  Same bar as everywhere: labels, counters and sufficient statistics identical to the CPU
oracle's, log-likelihoods to 1e-9."""
import numpy as np
import pytest

from conftest import make_data, make_prior
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,D,K_true,K_init,r", [(6000, 16, 8, 20, 1.5), (5000, 2, 6, 30, 1.0), (4000, 8, 5, 12, 1.5),
                                                 (3000, 3, 4, 100, 1.0), (3000, 1, 3, 9, 1.5), (3000, 12, 6, 127, 1.0)])
def test_cluster_step_engine_matches_oracle(gpu_lib, N, D, K_true, K_init, r):
    X, _ = make_data(N, D, K_true, 5)
    m_0, k_0, v_0, S_0 = make_prior(D)
    K_max = 160
    rng = np.random.RandomState(17)
    z0 = np.unique(rng.randint(0, K_init, N), return_inverse=True)[1].astype(np.int64)
    orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=K_max)
    orc.set_assignments(z0)
    ch = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, K_max)
    ch.set_engine("cluster")
    ch.set_assignments(z0)
    tab = O.logcount_table(N, r) if r > 1 else None
    fast = 0
    for s in range(4):
        order = rng.permutation(N) if r > 1 else None
        u = rng.random_sample(N)
        use_power = r > 1 and s > 0
        so = orc.sweep(u, 1.0, order=order, logcount_tab=tab if use_power else None)
        sg = ch.sweep(1.0, r if use_power else 1.0, order, u)
        fast += sg.fast_steps
        assert (sg.K, sg.moves, sg.births, sg.deaths, sg.evals) == (so.K_end, so.moves, so.births, so.deaths, so.evals), s
        np.testing.assert_array_equal(ch.assignments(), orc.assignments)
    assert fast > 0.5 * 4 * N, "the cluster step engine should have resolved most of the data (%d of %d)" % (fast, 4 * N)
    st = ch.get_state(inv_covar=False)
    K = orc.K
    np.testing.assert_array_equal(st["counts"], orc.counts)
    np.testing.assert_array_equal(st["m_num"], orc.m_N_numerators)        # same operations in the same order: same bits
    np.testing.assert_array_equal(st["S_part"], orc.S_N_partials)
    np.testing.assert_allclose(st["logdet"][:K], orc.logdet_covars[:K], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(ch.log_marg(1.0), orc.log_marg(1.0), rtol=1e-9)
    idx = np.arange(0, N, 97)
    np.testing.assert_allclose(ch.log_post_pred(idx), np.stack([orc.log_post_pred(i) for i in idx]), rtol=1e-9)


def test_adaptive_policy_uses_the_step_engine_then_the_windows(gpu_lib):
    """A chain from the rand initial state: the first sweeps (most data move) run on the cluster step engine, the later
    ones (few movers) on the resident engine's windows; the chain is the one the sequential engine walks."""
    N, D, K_true = 60000, 16, 10
    X, _ = make_data(N, D, K_true, 3)
    prior = make_prior(D)
    rng = np.random.RandomState(4)
    z0 = rng.randint(0, 10, N).astype(np.int64)
    ins = [(rng.permutation(N), rng.random_sample(N)) for _ in range(5)]
    out = []
    for engine in ("adaptive", "sequential"):
        ch = gpu_lib.Chain(X, *prior, 96)
        ch.set_engine(engine)
        ch.set_assignments(z0)
        tr = []
        for s, (o, u) in enumerate(ins):
            sg = ch.sweep(1.0, 1.5 if s else 1.0, o, u)
            tr.append((sg.K, sg.moves, sg.births, sg.deaths, sg.evals, sg.windows, sg.launches))
        out.append((tr, ch.assignments(), ch.get_state()["S_part"], ch.get_state()["m_num"]))
        ch.close()
    assert [t[:5] for t in out[0][0]] == [t[:5] for t in out[1][0]]
    np.testing.assert_array_equal(out[0][1], out[1][1])
    np.testing.assert_array_equal(out[0][2], out[1][2])
    np.testing.assert_array_equal(out[0][3], out[1][3])
    assert out[0][0][0][5] == 0 and out[0][0][-1][5] > 0, out[0][0]   # first sweep: no windows; last sweep: windows


@pytest.mark.parametrize("N", [1, 2, 31, 32, 33, 65, 257])
def test_cluster_step_engine_ragged_sizes(gpu_lib, N):
    """Chains shorter than / not a multiple of the producer's ring (32 data per half): every datum is still visited once,
    in order, and the chain is the oracle's."""
    D = 3
    X, _ = make_data(max(N, 4), D, 2, 9)
    X = np.ascontiguousarray(X[:N])
    m_0, k_0, v_0, S_0 = make_prior(D)
    rng = np.random.RandomState(N)
    z0 = np.unique(rng.randint(0, 3, N), return_inverse=True)[1].astype(np.int64)
    orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=max(8, N))
    orc.set_assignments(z0)
    ch = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, max(8, N))
    ch.set_engine("cluster")
    ch.set_assignments(z0)
    for s in range(3):
        u = rng.random_sample(N)
        so = orc.sweep(u, 1.0)
        sg = ch.sweep(1.0, 1.0, None, u)
        assert (sg.K, sg.moves, sg.births, sg.deaths, sg.evals) == (so.K_end, so.moves, so.births, so.deaths, so.evals), s
        np.testing.assert_array_equal(ch.assignments(), orc.assignments)
    st = ch.get_state(inv_covar=False)
    np.testing.assert_array_equal(st["m_num"], orc.m_N_numerators)
    np.testing.assert_array_equal(st["S_part"], orc.S_N_partials)


def test_cluster_step_engine_with_many_births_and_deaths(gpu_lib):
    """A large alpha opens components all the time and most die again: every one of them is a hand-back to the general
    step (and past a rate of one per few hundred data the adaptive policy leaves the cluster engine for the resident one,
    which resolves them in-kernel); both policies walk the oracle's chain."""
    N, D = 4000, 2
    X, _ = make_data(N, D, 5, 6)
    m_0, k_0, v_0, S_0 = make_prior(D)
    for engine in ("cluster", "adaptive"):
        rng = np.random.RandomState(31)
        z0 = np.unique(rng.randint(0, 6, N), return_inverse=True)[1].astype(np.int64)
        orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=400)
        orc.set_assignments(z0)
        ch = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, 400)
        ch.set_engine(engine)
        ch.set_assignments(z0)
        births = 0
        for s in range(3):
            u = rng.random_sample(N)
            so = orc.sweep(u, 60.0)
            sg = ch.sweep(60.0, 1.0, None, u)
            births += so.births
            assert (sg.K, sg.moves, sg.births, sg.deaths, sg.evals) == (so.K_end, so.moves, so.births, so.deaths, so.evals), (engine, s)
            np.testing.assert_array_equal(ch.assignments(), orc.assignments)
        assert births > 50
        st = ch.get_state(inv_covar=False)
        np.testing.assert_array_equal(st["S_part"], orc.S_N_partials)
        ch.close()
