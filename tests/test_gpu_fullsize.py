"""GPU parity at the sizes BASELINE.json names, against the CPU oracle run live on the same seeded inputs and against
the fixture generated from the reference itself (tests/golden/golden_large.json, make_golden_large.py):

  * C3 -- the chain bench.py times (PCRPMM r = 1.5, N = 1e6, D = 16, rand init K = 100): sweeps 0 and 1, both cold;
  * C5's shard (PCRPMM r = 1.5, N = 1e6, D = 8);
  * C4's component size (CRPMM, D = 64) on N = 5e4;
  * C2 (CRPMM, N = 1e5, D = 2): two sweeps against the reference's own output.

Labels, counters and the sufficient statistics must be identical (bit-exact); log-determinants and log_marg within the
1e-9 relative that north_star states.  The oracle runs (minutes of CPU at these sizes) are started in background threads
when the module is first used, so they overlap each other and the GPU work.
"""
import hashlib
import json
import os
import sys
import threading

import numpy as np
import pytest

from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import bench  # noqa: E402  (the generator and per-sweep inputs of the benchmarked chain)

pytestmark = pytest.mark.gpu
RTOL = 1e-9

#          name: (workload key of bench.py, N, sweeps)
LARGE = {"c3": ("c3", 1000000, 2), "c5": ("c5", 1000000, 2), "c4": ("c4", 50000, 1)}


def _inputs(name):
    wl, N, sweeps = LARGE[name]
    sampler, _, D, K_true, power, cov = bench.WORKLOADS[wl]
    X, _, z0 = bench.gen_data(N, D, K_true, 1)
    return X, z0, bench.prior_for(D, cov), 4 * K_true + 64, power, sweeps, N


class _OracleRun(threading.Thread):
    """Sweeps 0..S-1 of the benchmarked chain through the C oracle; keeps the per-sweep snapshots."""

    def __init__(self, name):
        super(_OracleRun, self).__init__(daemon=True)
        self.name_, self.snaps, self.err = name, [], None

    def run(self):
        try:
            X, z0, (m_0, k_0, v_0, S_0), K_max, power, sweeps, N = _inputs(self.name_)
            orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=K_max)
            orc.set_assignments(z0)
            tab = O.logcount_table(N, power) if power > 1 else None
            for s in range(sweeps):
                order, u = bench.step_input(N, s, power, 1)
                st = orc.sweep(u, 1.0, order=order, logcount_tab=tab if s > 0 else None)
                self.snaps.append(dict(stats=(st.K_end, st.moves, st.births, st.deaths, st.evals), z=orc.assignments))
            self.final = dict(counts=orc.counts, m_num=orc.m_N_numerators, S_part=orc.S_N_partials,
                              logdet=orc.logdet_covars, log_marg=orc.log_marg(1.0), K=orc.K)
        except Exception as e:  # pragma: no cover
            self.err = e


_RUNS = {}


def _oracle(name):
    if not _RUNS:
        for n in LARGE:
            _RUNS[n] = _OracleRun(n)
            _RUNS[n].start()
    return _RUNS[name]


@pytest.mark.parametrize("name", ["c5", "c4", "c3"])
def test_benchmarked_chain_matches_oracle(gpu_lib, name):
    run = _oracle(name)
    X, z0, (m_0, k_0, v_0, S_0), K_max, power, sweeps, N = _inputs(name)
    ch = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, K_max)
    ch.set_assignments(z0)
    got, guard = [], 0
    for s in range(sweeps):
        order, u = bench.step_input(N, s, power, 1)
        sg = ch.sweep(1.0, power if s > 0 else 1.0, order, u)
        guard += sg.guard_hits
        got.append(dict(stats=(sg.K, sg.moves, sg.births, sg.deaths, sg.evals), z=ch.assignments(),
                        min_margin=sg.min_margin))
    st = ch.get_state(z=False, inv_covar=False)
    lm = ch.log_marg(1.0)
    ch.close()
    run.join()
    assert run.err is None, run.err
    for s in range(sweeps):
        assert got[s]["stats"] == run.snaps[s]["stats"], (name, s, got[s]["stats"], run.snaps[s]["stats"])
        diff = np.nonzero(got[s]["z"] != run.snaps[s]["z"])[0]
        assert diff.size == 0, "%s sweep %d: %d labels differ, first at datum %d (min margin %.3g, guard hits %d)" % (
            name, s, diff.size, diff[0], got[s]["min_margin"], guard)
    K = run.final["K"]
    np.testing.assert_array_equal(st["counts"], run.final["counts"])
    np.testing.assert_array_equal(st["m_num"], run.final["m_num"])       # same operations in the same order: same bits
    np.testing.assert_array_equal(st["S_part"], run.final["S_part"])
    np.testing.assert_allclose(st["logdet"][:K], run.final["logdet"][:K], rtol=RTOL, atol=1e-11)
    np.testing.assert_allclose(lm, run.final["log_marg"], rtol=RTOL)
    # every committed draw kept its distance from the boundaries (below the guard the datum was redone exactly)
    assert min(g["min_margin"] for g in got) >= 0.0


def _digest(z):
    return hashlib.sha256(np.ascontiguousarray(z, dtype="<i8").tobytes()).hexdigest()


def _large_golden():
    with open(os.path.join(HERE, "golden", "golden_large.json")) as fh:
        return json.load(fh)["cases"]


@pytest.mark.parametrize("case", _large_golden(), ids=lambda c: c["name"])
def test_reference_golden_at_size(gpu_lib, case):
    """The drop-in classes against the outputs of the reference itself at N = 1e5 (C2) / 3e4: same global RNG
    consumption (random / np.random), so the SHA-256 of the labels after every sweep must be the reference's."""
    import pybgmm_b200 as P
    import cases as C
    X, z_true = C.gen(case["N"], case["D"], case["K_true"], case["seed"])
    m_0, k_0, v_0, S_0 = C.prior_for(case["D"], "full")
    cls = getattr(P, case["cls"])
    model = cls(X, P.NIW(m_0, k_0, v_0, S_0), 1.0, None, assignments="rand", K=case["K_true"], K_max=case["K_max"])
    model.metrics_every = 0
    assert _digest(model.components.assignments) == case["z0_sha256"]
    for s, want in enumerate(case["sweeps"]):
        if case["cls"] == "CRPMM":
            model.collapsed_gibbs_sampler(1, z_true, num_saved=0)
        else:
            model.collapsed_gibbs_sampler(1, z_true, num_saved=0, power_burnin=(0 if s == 0 else -1), **case["kwargs"])
        c = model.components
        assert c.K == want["K"], (s, c.K, want["K"])
        assert c.counts[:c.K].tolist() == want["counts"], s
        assert _digest(c.assignments) == want["z_sha256"], s
        np.testing.assert_allclose(model.log_marg(), want["log_marg"], rtol=RTOL)


@pytest.mark.parametrize("guard", [1e-3, 0.05])
def test_margin_guard_path_is_exact(gpu_lib, guard):
    """With the guard raised to 1e-3 / 5e-2 a large share of the draws go through the exact path (records rebuilt from
    the statistics, libm log / exp, sequential-subtract draw): the chain must still be the oracle's, in every engine
    mode, and the hits are counted."""
    from conftest import make_data, make_prior
    N, D, K_true = 6000, 16, 12
    X, _ = make_data(N, D, K_true, 3)
    m_0, k_0, v_0, S_0 = make_prior(D)
    z0 = O.init_assignments(N, "rand", K_true)
    rng = np.random.RandomState(5)
    ins = [(rng.permutation(N), rng.random_sample(N)) for _ in range(4)]
    tab = O.logcount_table(N, 1.5)
    orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=64)
    orc.set_assignments(z0)
    want = []
    for s, (o, u) in enumerate(ins):
        so = orc.sweep(u, 1.0, order=o, logcount_tab=tab if s else None)
        want.append(((so.K_end, so.moves, so.births, so.deaths, so.evals), orc.assignments))
    for engine in ("adaptive", "sequential", "windows", "generic"):
        ch = gpu_lib.Chain(X, m_0, k_0, v_0, S_0, 64)
        ch.set_guard(guard)
        ch.set_engine(engine)
        ch.set_assignments(z0)
        hits = 0
        for s, (o, u) in enumerate(ins):
            sg = ch.sweep(1.0, 1.5 if s else 1.0, o, u)
            hits += sg.guard_hits
            assert (sg.K, sg.moves, sg.births, sg.deaths, sg.evals) == want[s][0], (engine, s)
            np.testing.assert_array_equal(ch.assignments(), want[s][1])
        assert hits > 0, engine
        np.testing.assert_array_equal(ch.get_state()["S_part"], orc.S_N_partials)
        ch.close()
