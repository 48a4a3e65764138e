"""GPU: the drop-in classes (NIW / GaussianComponents{,Diag} / CRPMM / PCRPMM -> ctypes -> CUDA) replay the
fixtures generated from the reference itself and the constants written in the reference's tests."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu
GOLD = cases.golden()
RTOL = 1e-9


@pytest.mark.parametrize("engine", [None, "sequential", "windows"])
@pytest.mark.parametrize("case", GOLD["samplers"], ids=[c["name"] for c in GOLD["samplers"]])
def test_sampler_cases(gpu_lib, case, engine):
    model, rec, z0 = cases.run_sampler_case_gpu(case, engine)
    c = model.components
    np.testing.assert_array_equal(z0, case["z0"])
    np.testing.assert_array_equal(c.assignments, case["z"])                 # bit-exact integer bookkeeping
    assert c.K == case["K"]
    np.testing.assert_array_equal(c.counts[:c.K], case["counts"])
    np.testing.assert_array_equal(rec["components"], case["K_trace"])
    np.testing.assert_allclose(rec["log_marg"], case["log_marg_trace"], rtol=RTOL)
    np.testing.assert_allclose(model.log_marg(), case["log_marg"], rtol=RTOL)
    np.testing.assert_allclose([rec["nmi"][-1], rec["mi"][-1], rec["vi"][-1]], [case["nmi"], case["mi"], case["vi"]],
                               rtol=1e-9, atol=1e-12)
    assert float(rec["loss"][-1]) == case["loss"]
    assert sorted(rec.keys()) == sorted(["sample_time", "log_marg", "components", "nmi", "mi", "nk", "loss", "bic",
                                         "vi", "alpha"])   # gmm.py:45-63


def test_reference_test_constants(gpu_lib):
    by = {c["name"]: c for c in GOLD["samplers"]}
    model, _, _ = cases.run_sampler_case_gpu(by["G1_G2_test_igmm_sampling_2d"])
    np.testing.assert_array_equal(model.components.assignments, cases.G1_ASSIGNMENTS)
    np.testing.assert_almost_equal(model.log_marg(), cases.G2_LOG_MARG)
    model, _, _ = cases.run_sampler_case_gpu(by["G3_test_igmm_each_in_own"])
    np.testing.assert_array_equal(model.components.assignments, cases.G3_ASSIGNMENTS)
    model, _, _ = cases.run_sampler_case_gpu(by["G4_test_igmm_log_marg_each_in_own"])
    np.testing.assert_almost_equal(model.log_marg(), cases.G4_LOG_MARG)


def test_component_kats(gpu_lib):
    from pybgmm_b200 import NIW, GaussianComponents
    X = np.array([[-0.3406, -0.0593, -0.0686]])
    gmm = GaussianComponents(X, NIW(np.zeros(3), 0.05, 4, 0.001 * np.eye(3)))
    np.testing.assert_almost_equal(gmm.log_prior(0), cases.K1_LOG_PRIOR)
    X = np.array([[-0.3406, -0.3593, -0.0686], [-0.3381, 0.2993, 0.925], [-0.5, -0.101, 0.75]])
    gmm = GaussianComponents(X, NIW(np.zeros(3), 0.05, 6, 0.5 * np.eye(3)), [0, 0, 0])
    np.testing.assert_almost_equal(gmm.log_marg_k(0), cases.K3_LOG_MARG_K)
    gmm = GaussianComponents(np.array([[1.2, 0.9], [-0.1, 0.8], [0.5, 0.4]]),
                             NIW(m_0=np.array([0.0, 0.0]), k_0=2., v_0=5, S_0=5. * np.eye(2)))
    gmm.add_item(0, 0)
    gmm.add_item(1, 0)
    np.testing.assert_almost_equal(gmm.log_post_pred_k(2, 0), cases.K4_LOG_POST_PRED_K)
    mu, sigma = gmm.map(0)
    np.testing.assert_almost_equal(mu, cases.K2_MAP_MU)
    np.testing.assert_almost_equal(sigma, cases.K2_MAP_SIGMA)


@pytest.mark.parametrize("case", GOLD["components"], ids=[c["name"] for c in GOLD["components"]])
def test_component_cases(gpu_lib, case):
    from pybgmm_b200 import NIW, GaussianComponents, GaussianComponentsDiag
    X, _ = cases.gen(case["N"], case["D"], case["K_true"], case["seed"])
    m_0, k_0, v_0, S_0 = cases.prior_for(case["D"], case["cov"])
    cls = GaussianComponents if case["cov"] == "full" else GaussianComponentsDiag
    c = cls(X, NIW(m_0, k_0, v_0, S_0), np.array(case["z"]), K_max=16)
    K = c.K
    np.testing.assert_allclose(c.cached_log_prior, case["log_prior"], rtol=RTOL)
    np.testing.assert_allclose(c.log_post_pred_many(np.arange(case["N"])), case["log_post_pred"], rtol=RTOL)
    np.testing.assert_allclose(c.log_post_pred(3), case["log_post_pred"][3], rtol=RTOL)
    np.testing.assert_allclose([c.log_marg_k(k) for k in range(K)], case["log_marg_k"], rtol=RTOL)
    np.testing.assert_allclose(c.log_marg(), case["log_marg"], rtol=RTOL)
    np.testing.assert_array_equal(c.m_N_numerators[:K], case["m_N_numerators"])     # same op order -> same bits
    np.testing.assert_array_equal(c.S_N_partials[:K], case["S_N_partials"])
    ld = c.logdet_covars if case["cov"] == "full" else c.log_prod_vars
    iv = c.inv_covars if case["cov"] == "full" else c.inv_vars
    np.testing.assert_allclose(ld[:K], case["logdet"], rtol=RTOL, atol=1e-11)
    np.testing.assert_allclose(iv[:K], case["inv"], rtol=1e-8, atol=1e-10)
    assert not ld[K:].any() and not c.counts[K:].any()                               # gaussian_components.py:200-204
    for i, k in case["ops"]:
        c.del_item(i)
        c.add_item(i, k)
    Ka = c.K
    assert Ka == case["K_after"]
    np.testing.assert_array_equal(c.assignments, case["z_after"])
    np.testing.assert_array_equal(c.counts[:Ka], case["counts_after"])
    np.testing.assert_array_equal(c.m_N_numerators[:Ka], case["m_after"])
    np.testing.assert_array_equal(c.S_N_partials[:Ka], case["S_after"])
    ld = c.logdet_covars if case["cov"] == "full" else c.log_prod_vars
    np.testing.assert_allclose(ld[:Ka], case["logdet_after"], rtol=RTOL, atol=1e-11)


def test_cache_restore_protocol(gpu_lib):
    """cache_component_stats / del_item / restore_component_from_stats round trip (crpmm.py:62-85)."""
    from pybgmm_b200 import NIW, GaussianComponents
    X, _ = cases.gen(40, 3, 3, 5)
    m_0, k_0, v_0, S_0 = cases.prior_for(3, "full")
    c = GaussianComponents(X, NIW(m_0, k_0, v_0, S_0), np.arange(40) % 3, K_max=8)
    before = c.chain.get_state()
    k_old = int(c.assignments[7])
    stats_old = c.cache_component_stats(k_old)
    c.del_item(7)
    assert c.assignments[7] == -1 and c.counts[k_old] == stats_old[4] - 1
    c.restore_component_from_stats(k_old, *stats_old)
    after = c.chain.get_state()
    np.testing.assert_array_equal(after["m_num"], before["m_num"])
    np.testing.assert_array_equal(after["S_part"], before["S_part"])
    np.testing.assert_array_equal(after["counts"], before["counts"])
    np.testing.assert_allclose(after["logdet"], before["logdet"], rtol=1e-14)
    # the reference completes the restore on the host: components.assignments[i] = k_old (crpmm.py:84-85); here that
    # assignment is written through to the device label
    assert c.chain.assignments()[7] == -1
    c.assignments[7] = k_old
    np.testing.assert_array_equal(c.chain.assignments(), before["z"])
    np.testing.assert_array_equal(c.assignments, before["z"])
    # ... so a following sweep starts from a consistent state: same chain as an oracle that never took the detour
    from oracle import oracle as O
    orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=8)
    orc.set_assignments(np.arange(40) % 3)
    u = np.random.RandomState(5).random_sample(40)
    so = orc.sweep(u, 1.0)
    sg = c.chain.sweep(1.0, 1.0, None, u)
    assert (sg.K, sg.moves, sg.births, sg.deaths) == (so.K_end, so.moves, so.births, so.deaths)
    np.testing.assert_array_equal(c.chain.assignments(), orc.assignments)
    np.testing.assert_array_equal(c.chain.get_state()["counts"], orc.counts)
    np.testing.assert_array_equal(c.chain.get_state()["S_part"], orc.S_N_partials)


def test_errors_match_reference_conventions(gpu_lib):
    from pybgmm_b200 import NIW, CRPMM
    with pytest.raises(ValueError):                     # igmm.py:75-76
        CRPMM(np.zeros(5), NIW(np.zeros(1), 1., 2, np.eye(1)), 1., None)
    with pytest.raises(AssertionError):                 # niw.py:21
        NIW(np.zeros(3), 1., 2, np.eye(3))
    with pytest.raises(AssertionError):                 # igmm.py:111
        CRPMM(np.zeros((5, 2)), NIW(np.zeros(2), 1., 4, np.eye(2)), 1., None, covariance_type="bogus")
    with pytest.raises(AssertionError):                 # gaussian_components.py:103-105
        CRPMM(np.random.randn(5, 2), NIW(np.zeros(2), 1., 4, np.eye(2)), 1., None, assignments=np.array([0, 2, 2, 0, 0]))


@pytest.mark.parametrize("burnin,flag", [(2, True), (0, True), (-1, True), (0, False)])
def test_adapcrpmm_burnin_matches_oracle(gpu_lib, burnin, flag):
    """ADAPCRPMM (adapcrpmm.py:83-157) beyond what the reference itself can run: burn-in sweeps (its own loop stops
    with UnboundLocalError for adapcrp_burnin >= 0) are CRP sweeps, then the adaptive power; CUDA path vs oracle."""
    case = dict(cls="ADAPCRPMM", N=150, D=2, K_true=4, seed=21, cov="full", assignments="rand", K_init=6, n_iter=6,
                v_0=None, K_max=64, kwargs=dict(adapcrp_burnin=burnin, r_up=1.6, adapcrp_perct=0.1, flag_adapcrp=flag))
    orc, z0, _ = cases.run_sampler_case_oracle(case)
    model, rec, z0g = cases.run_sampler_case_gpu(case)
    np.testing.assert_array_equal(z0g, z0)
    np.testing.assert_array_equal(model.components.assignments, orc.assignments)
    np.testing.assert_array_equal(model.components.counts[:model.components.K], orc.counts[:orc.K])
    np.testing.assert_allclose(model.log_marg(), orc.log_marg(1.0), rtol=RTOL)
    powers = model.adapcrp_powers
    assert len(powers) == 6 and all(p == 1.0 for p in powers[:max(burnin + 1, 0)])
    if not flag:
        assert all(p == 1.0 for p in powers)
