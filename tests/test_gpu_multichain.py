"""GPU: many independent chains on one GPU (bgmm_fork + bgmm_sweep_many: one kernel launch per sweep, one thread block
per chain) and the register-resident sequential step they run on (csrc/bgmm_seq.cuh).  Every chain must be exactly the
chain the CPU oracle walks for its own seeded inputs -- labels, counters, sufficient statistics bit-identical."""
import os

import numpy as np
import pytest

from conftest import make_data, make_prior
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _oracle_chain(X, prior, K_max, z0, ins, r):
    m_0, k_0, v_0, S_0 = prior
    orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=K_max)
    orc.set_assignments(z0)
    tab = O.logcount_table(X.shape[0], r) if r > 1 else None
    trace = []
    for s, (o, u) in enumerate(ins):
        st = orc.sweep(u, 1.0, order=o, logcount_tab=tab if (r > 1 and s > 0) else None)
        trace.append(((st.K_end, st.moves, st.births, st.deaths, st.evals), orc.assignments))
    return orc, trace


@pytest.mark.parametrize("N,D,K_true,M,r", [(3000, 16, 10, 5, 1.5), (2000, 2, 6, 7, 1.0), (1500, 8, 8, 3, 1.5),
                                            (400, 4, 4, 170, 1.0), (900, 1, 3, 4, 1.0)])
def test_sweep_many_matches_per_chain_oracle(gpu_lib, N, D, K_true, M, r):
    import torch
    X, _ = make_data(N, D, K_true, 2)
    prior = make_prior(D)
    K_max = 64
    dev = torch.device("cuda", 0)
    first = gpu_lib.Chain(X, *prior, K_max)
    chains = [first] + [first.fork() for _ in range(M - 1)]
    rng = np.random.RandomState(17)
    z0s = [np.unique(rng.randint(0, K_true + (m % 3), N), return_inverse=True)[1].astype(np.int64) for m in range(M)]
    sweeps = 4
    ins = [[(rng.permutation(N) if r > 1 else None, rng.random_sample(N)) for _ in range(sweeps)] for _ in range(M)]
    for c, z0 in zip(chains, z0s):
        c.set_assignments(z0)
    group = gpu_lib.ChainGroup(chains)
    check = range(M) if M <= 8 else list(range(0, M, 23)) + [M - 1]     # oracle runs for a sample of a large group
    want = {m: _oracle_chain(X, prior, K_max, z0s[m], ins[m], r) for m in check}
    fast = 0
    for s in range(sweeps):
        d_u = torch.from_numpy(np.stack([ins[m][s][1] for m in range(M)])).to(dev)
        d_o = torch.from_numpy(np.stack([ins[m][s][0] for m in range(M)])).to(dev) if r > 1 else None
        sts = group.sweep_dev(1.0, r if s > 0 else 1.0, d_o, d_u)
        torch.cuda.synchronize()
        for m in check:
            sg = sts[m]
            assert (sg.K, sg.moves, sg.births, sg.deaths, sg.evals) == want[m][1][s][0], (m, s)
            np.testing.assert_array_equal(chains[m].assignments(), want[m][1][s][1])
            fast += sg.fast_steps
    assert fast > 0, "the register-resident step never ran"
    for m in check:
        st = chains[m].get_state(inv_covar=False)
        orc = want[m][0]
        np.testing.assert_array_equal(st["counts"], orc.counts)
        np.testing.assert_array_equal(st["m_num"], orc.m_N_numerators)
        np.testing.assert_array_equal(st["S_part"], orc.S_N_partials)
        np.testing.assert_allclose(st["logdet"][:orc.K], orc.logdet_covars[:orc.K], rtol=1e-9, atol=1e-11)
    for c in chains:
        c.close()


def test_forked_chain_equals_a_chain_of_its_own(gpu_lib):
    """A fork shares X / tables with its parent; swept on its own (bgmm_sweep) it is the chain a separately created
    handle walks."""
    N, D, K_true = 2500, 16, 8
    X, _ = make_data(N, D, K_true, 4)
    prior = make_prior(D)
    rng = np.random.RandomState(3)
    z0 = np.unique(rng.randint(0, K_true, N), return_inverse=True)[1].astype(np.int64)
    ins = [(rng.permutation(N), rng.random_sample(N)) for _ in range(3)]
    parent = gpu_lib.Chain(X, *prior, 48)
    fork = parent.fork()
    own = gpu_lib.Chain(X, *prior, 48)
    for c in (fork, own):
        c.set_assignments(z0)
    np.testing.assert_array_equal(fork.log_prior(), own.log_prior())
    for s, (o, u) in enumerate(ins):
        a = fork.sweep(1.0, 1.5 if s else 1.0, o, u)
        b = own.sweep(1.0, 1.5 if s else 1.0, o, u)
        assert (a.K, a.moves, a.evals) == (b.K, b.moves, b.evals)
    np.testing.assert_array_equal(fork.assignments(), own.assignments())
    np.testing.assert_array_equal(fork.get_state()["S_part"], own.get_state()["S_part"])
    parent.close()          # the shared buffers outlive the parent while a fork uses them
    st = fork.sweep(1.0, 1.5, *ins[0])
    assert st.K > 0
    fork.close()
    own.close()


@pytest.mark.parametrize("D,r", [(16, 1.5), (8, 1.0), (2, 1.0)])
def test_register_step_and_general_step_walk_the_same_chain(gpu_lib, D, r):
    """The sequential engine with the register-resident step (default) and with the general step only (BGMM_TUNE=8):
    both must be the oracle's chain, cold sweeps included; the first must actually use the register step."""
    N, K_true = 8000, 12
    X, _ = make_data(N, D, K_true, 6)
    prior = make_prior(D)
    rng = np.random.RandomState(11)
    z0 = np.unique(rng.randint(0, K_true, N), return_inverse=True)[1].astype(np.int64)
    ins = [(rng.permutation(N) if r > 1 else None, rng.random_sample(N)) for _ in range(3)]
    orc, trace = _oracle_chain(X, prior, 64, z0, ins, r)
    saved = os.environ.get("BGMM_TUNE")
    try:
        for tune, engine in (("0", "sequential"), ("8", "sequential"), ("0", "adaptive")):
            os.environ["BGMM_TUNE"] = tune
            ch = gpu_lib.Chain(X, *prior, 64)
            ch.set_engine(engine)
            ch.set_assignments(z0)
            fast = 0
            for s, (o, u) in enumerate(ins):
                sg = ch.sweep(1.0, r if s else 1.0, o, u)
                fast += sg.fast_steps
                assert (sg.K, sg.moves, sg.births, sg.deaths, sg.evals) == trace[s][0], (tune, engine, s)
                np.testing.assert_array_equal(ch.assignments(), trace[s][1])
            assert (fast > 0) == (tune == "0"), (tune, engine, fast)
            np.testing.assert_array_equal(ch.get_state()["S_part"], orc.S_N_partials)
            ch.close()
    finally:
        if saved is None:
            os.environ.pop("BGMM_TUNE", None)
        else:
            os.environ["BGMM_TUNE"] = saved
