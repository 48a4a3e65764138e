"""GPU parity at larger sizes: the CUDA engine against the C oracle on long seeded chains (every window /
sequential / hand-back path of the resident engine is exercised many thousand times), and size-independent
properties at the full size BASELINE.json names (N = 1e6, D = 16, K = 100)."""
import numpy as np
import pytest

from conftest import make_data, make_prior
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _chain_pair(gpu_lib, N, D, K_true, K_init, cov="full", seed=1):
    X, _ = make_data(N, D, K_true, seed)
    m_0, k_0, v_0, S_0 = make_prior(D, cov)
    K_max = 4 * K_true + 64
    z0 = O.init_assignments(N, "rand", K_init)
    orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=K_max, covariance_type=cov)
    orc.set_assignments(z0)
    ch = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, K_max, covariance_type=cov)
    ch.set_assignments(z0)
    return X, orc, ch


@pytest.mark.parametrize("N,D,K_true,r,both", [(60000, 8, 20, 1.5, True), (40000, 16, 40, 1.0, True),
                                               (50000, 2, 30, 1.5, False)])
def test_long_chain_matches_oracle(gpu_lib, N, D, K_true, r, both):
    """Cold chain from a random initial state down to the converged regime: assignments and counters identical to the
    oracle after every sweep, sufficient statistics bit-identical at the end."""
    X, orc, ch = _chain_pair(gpu_lib, N, D, K_true, K_true)
    tab = O.logcount_table(N, r) if r > 1 else None
    rng = np.random.RandomState(9)
    modes = set()
    for s in range(6):
        order = rng.permutation(N) if r > 1 else None
        u = rng.random_sample(N)
        use_power = r > 1 and s > 0
        so = orc.sweep(u, 1.0, order=order, logcount_tab=tab if use_power else None)
        sg = ch.sweep(1.0, r if use_power else 1.0, order, u)
        assert (sg.K, sg.moves, sg.births, sg.deaths, sg.evals) == (so.K_end, so.moves, so.births, so.deaths, so.evals), s
        np.testing.assert_array_equal(ch.assignments(), orc.assignments)
        modes.add("seq" if sg.seq_data else None)
        modes.add("win" if sg.windows else None)
    if both:
        assert {"seq", "win"} <= modes, "the chain should pass through both engine modes"
    st = ch.get_state(inv_covar=False)
    K = orc.K
    np.testing.assert_array_equal(st["counts"], orc.counts)
    np.testing.assert_array_equal(st["m_num"], orc.m_N_numerators)
    np.testing.assert_array_equal(st["S_part"], orc.S_N_partials)
    np.testing.assert_allclose(st["logdet"][:K], orc.logdet_covars[:K], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(ch.log_marg(1.0), orc.log_marg(1.0), rtol=1e-9)


def test_full_size_properties(gpu_lib):
    """N = 1e6, D = 16, K = 100 (BASELINE.json C3): properties that do not need a CPU run of the same size.
    (1) the window engine and the purely sequential engine walk the same chain; (2) counts are the histogram of the
    assignments; (3) the incrementally maintained statistics equal a fresh build from the final assignments to
    rounding; (4) log_post_pred of the final state agrees with the oracle on a sample of data."""
    N, D, K_true = 1000000, 16, 100
    X, _ = make_data(N, D, K_true, 1)
    m_0, k_0, v_0, S_0 = make_prior(D)
    K_max = 4 * K_true + 64
    rng = np.random.RandomState(3)
    z0 = rng.randint(0, K_true, N).astype(np.int64)
    orders = [rng.permutation(N) for _ in range(3)]
    unis = [rng.random_sample(N) for _ in range(3)]
    chains = []
    for engine in ("adaptive", "sequential"):
        ch = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, K_max)
        ch.set_assignments(z0)
        ch.set_engine(engine)
        for s in range(3):
            ch.sweep(1.0, 1.5 if s > 0 else 1.0, orders[s], unis[s])
        chains.append(ch)
    a, b = chains
    za, zb = a.assignments(), b.assignments()
    np.testing.assert_array_equal(za, zb)                                    # (1)
    st = a.get_state(inv_covar=False)
    K = st["K"]
    assert za.min() >= 0 and za.max() == K - 1
    np.testing.assert_array_equal(np.bincount(za, minlength=K_max), st["counts"])  # (2)
    assert st["counts"].sum() == N
    fresh = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, K_max)
    fresh.set_assignments(za)
    sf = fresh.get_state(inv_covar=False)
    np.testing.assert_allclose(st["m_num"][:K], sf["m_num"][:K], rtol=1e-10, atol=1e-7)      # (3)
    np.testing.assert_allclose(st["S_part"][:K], sf["S_part"][:K], rtol=1e-10, atol=1e-6)
    np.testing.assert_allclose(st["logdet"][:K], sf["logdet"][:K], rtol=1e-9, atol=1e-9)
    idx = np.arange(0, N, 50021)
    orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=K_max)                        # (4) oracle on the final state only
    orc.set_assignments(za)
    want = np.stack([orc.log_post_pred(i) for i in idx])
    np.testing.assert_allclose(a.log_post_pred(idx), want, rtol=1e-9)


def test_replicas_are_deterministic(gpu_lib):
    """The resident engine is a replicated state machine across 148 CTAs with lock-free exchanges: run the same chain
    several times (cold sweeps, window rounds, long windows) and require bit-identical labels, counters and
    statistics every time -- a timing-dependent replica disagreement shows up here as a difference or a watchdog."""
    N, D, K_true = 100000, 16, 100
    X, _ = make_data(N, D, K_true, 1)
    m_0, k_0, v_0, S_0 = make_prior(D)
    rng = np.random.RandomState(7)
    z0 = rng.randint(0, K_true, N).astype(np.int64)
    orders = [rng.permutation(N) for _ in range(5)]
    unis = [rng.random_sample(N) for _ in range(5)]
    ref = None
    for rep in range(6):
        ch = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, 4 * K_true + 64)
        ch.set_assignments(z0)
        trace = []
        for s in range(5):
            st = ch.sweep(1.0, 1.5 if s > 0 else 1.0, orders[s], unis[s])
            trace.append((st.K, st.moves, st.births, st.deaths, st.evals))
        state = ch.get_state(inv_covar=False, logdet=False)
        got = (trace, state["z"].tobytes(), state["m_num"].tobytes(), state["S_part"].tobytes())
        if ref is None:
            ref = got
        else:
            assert got[0] == ref[0], rep
            assert got[1:] == ref[1:], rep
        ch.close()
