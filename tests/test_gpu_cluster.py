"""GPU: the cluster-resident engine for full covariance at padded D = 32 / 64 (csrc/bgmm_big.cuh: records distributed over
the shared memories of a thread-block cluster, weights exchanged through DSMEM, two hardware cluster barriers per datum;
everything unusual handed to the generic engine's step).  Same bar as everywhere: labels, counters and sufficient
statistics identical to the CPU oracle's, log-likelihoods to 1e-9."""
import numpy as np
import pytest

from conftest import make_data, make_prior
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,D,K_true,K_init,r", [(3000, 64, 6, 9, 1.5), (4000, 32, 8, 12, 1.5), (2500, 48, 5, 7, 1.0),
                                                 (1500, 20, 4, 30, 1.0)])
def test_cluster_engine_matches_oracle(gpu_lib, N, D, K_true, K_init, r):
    X, _ = make_data(N, D, K_true, 5)
    m_0, k_0, v_0, S_0 = make_prior(D)
    K_max = 96
    rng = np.random.RandomState(13)
    z0 = np.unique(rng.randint(0, K_init, N), return_inverse=True)[1].astype(np.int64)
    orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=K_max)
    orc.set_assignments(z0)
    ch = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, K_max)
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
    assert fast > 0.5 * 4 * N, "the cluster engine should have resolved most of the data (%d of %d)" % (fast, 4 * N)
    st = ch.get_state(inv_covar=False)
    K = orc.K
    np.testing.assert_array_equal(st["counts"], orc.counts)
    np.testing.assert_array_equal(st["m_num"], orc.m_N_numerators)        # same operations in the same order: same bits
    np.testing.assert_array_equal(st["S_part"], orc.S_N_partials)
    np.testing.assert_allclose(st["logdet"][:K], orc.logdet_covars[:K], rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(ch.log_marg(1.0), orc.log_marg(1.0), rtol=1e-9)
    idx = np.arange(0, N, 97)
    np.testing.assert_allclose(ch.log_post_pred(idx), np.stack([orc.log_post_pred(i) for i in idx]), rtol=1e-9)


def test_cluster_engine_and_generic_engine_walk_the_same_chain(gpu_lib):
    N, D, K_true = 2000, 64, 5
    X, _ = make_data(N, D, K_true, 8)
    prior = make_prior(D)
    rng = np.random.RandomState(2)
    z0 = np.unique(rng.randint(0, 8, N), return_inverse=True)[1].astype(np.int64)
    ins = [(rng.permutation(N), rng.random_sample(N)) for _ in range(3)]
    out = []
    for engine in ("adaptive", "generic"):
        ch = gpu_lib.Chain(X, *prior, 64)
        ch.set_engine(engine)
        ch.set_assignments(z0)
        tr = []
        for s, (o, u) in enumerate(ins):
            sg = ch.sweep(1.0, 1.5 if s else 1.0, o, u)
            tr.append((sg.K, sg.moves, sg.births, sg.deaths, sg.evals, sg.fast_steps > 0))
        out.append((tr, ch.assignments(), ch.get_state()["S_part"]))
        ch.close()
    assert [t[:5] for t in out[0][0]] == [t[:5] for t in out[1][0]]
    assert all(t[5] for t in out[0][0]) and not any(t[5] for t in out[1][0])
    np.testing.assert_array_equal(out[0][1], out[1][1])
    np.testing.assert_array_equal(out[0][2], out[1][2])
