"""Shared fixtures: the reference's own known-answer constants (cited) and the golden.json replay helpers."""
import json
import os
import random

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

# ---- constants copied from the reference's tests (the expected VALUES only; cited) -------------------------
# pybgmm/tests/test_igmm.py:53-59  (seed 1, N=100, D=2, K_true=4, K=3 "rand", 10 sweeps)
G1_ASSIGNMENTS = [
    1, 2, 0, 0, 2, 1, 2, 1, 2, 0, 0, 1, 0, 2, 1, 0, 1, 1, 1, 0, 1, 1, 1, 0,
    2, 0, 1, 0, 1, 1, 1, 0, 2, 2, 1, 1, 2, 1, 0, 1, 1, 1, 1, 2, 2, 1, 1, 1,
    1, 0, 0, 1, 0, 0, 1, 2, 2, 1, 1, 0, 1, 2, 2, 1, 1, 1, 1, 2, 0, 0, 1, 2,
    0, 1, 0, 0, 1, 2, 1, 1, 2, 0, 0, 1, 2, 1, 2, 2, 1, 1, 0, 1, 1, 2, 2, 1,
    2, 1, 0, 2]
G2_LOG_MARG = -411.811711231                                              # test_igmm.py:101
G3_ASSIGNMENTS = [5, 2, 4, 3, 2, 7, 2, 7, 1, 0, 4, 6, 4, 1, 6, 4, 1, 7, 1, 0]  # test_igmm.py:143
G4_LOG_MARG = -30.771535771                                               # test_igmm.py:187
K1_LOG_PRIOR = -0.472067277015                                            # test_gaussian_components.py:33
K2_MAP_MU = [0.275, 0.425]                                                # test_gaussian_components.py:48
K2_MAP_SIGMA = [[0.55886364, 0.04840909], [0.04840909, 0.52068182]]       # :49-52
K3_LOG_MARG_K = -8.42365141729                                            # :79
K4_LOG_POST_PRED_K = -2.07325364088                                       # :104


def golden():
    """golden.json (CRPMM / PCRPMM / components) with the ADAPCRPMM cases of golden_adap.json appended."""
    with open(os.path.join(HERE, "golden", "golden.json")) as fh:
        gold = json.load(fh)
    with open(os.path.join(HERE, "golden", "golden_adap.json")) as fh:
        gold["samplers"] = gold["samplers"] + json.load(fh)["samplers"]
    return gold


def gen(N, D, K_true, seed):
    """Same generator (and same consumption of the global RNG streams) as tests/golden/make_golden.py."""
    random.seed(seed)
    np.random.seed(seed)
    z_true = np.random.randint(0, K_true, N)
    mu = np.random.randn(D, K_true) * 4.0
    X = (mu[:, z_true] + np.random.randn(D, N) * 0.7).T
    return np.ascontiguousarray(X), z_true


def prior_for(D, cov, v_0=None):
    v_0 = D + 3 if v_0 is None else v_0
    S_0 = 0.7 ** 2 * v_0 * (np.eye(D) if cov == "full" else np.ones(D))
    return np.zeros(D), 0.7 ** 2 / 4.0 ** 2, v_0, S_0


def run_sampler_case_oracle(case):
    """Replay a golden sampler case through the CPU oracle; returns the Oracle."""
    from oracle import oracle as O
    X, z_true = gen(case["N"], case["D"], case["K_true"], case["seed"])
    m_0, k_0, v_0, S_0 = prior_for(case["D"], case["cov"], case["v_0"])
    z0 = O.init_assignments(case["N"], case["assignments"], case["K_init"])
    orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=case["K_max"] or case["N"], covariance_type=case["cov"])
    orc.set_assignments(z0)
    kw = case["kwargs"]
    if case["cls"] == "CRPMM":
        O.run_crpmm(orc, case["n_iter"], 1.0)
    elif case["cls"] == "ADAPCRPMM":
        O.run_adapcrpmm(orc, case["n_iter"], 1.0, r_up=kw.get("r_up", 1.3), adapcrp_perct=kw.get("adapcrp_perct", 0.04),
                        adapcrp_burnin=kw.get("adapcrp_burnin", 0), flag_adapcrp=kw.get("flag_adapcrp", True))
    else:
        O.run_pcrpmm(orc, case["n_iter"], 1.0, n_power=kw.get("n_power", 1.01),
                     power_burnin=kw.get("power_burnin", 0), flag_power=kw.get("flag_power", True))
    return orc, z0, z_true


def run_sampler_case_gpu(case, engine=None):
    """Replay a golden sampler case through the drop-in classes (CUDA engine); returns (model, record_dict)."""
    import pybgmm_b200 as P
    X, z_true = gen(case["N"], case["D"], case["K_true"], case["seed"])
    m_0, k_0, v_0, S_0 = prior_for(case["D"], case["cov"], case["v_0"])
    cls = getattr(P, case["cls"])
    model = cls(X, P.NIW(m_0, k_0, v_0, S_0), 1.0, None, assignments=case["assignments"], K=case["K_init"],
                K_max=case["K_max"] or case["N"], covariance_type=case["cov"])
    z0 = model.components.assignments.copy()
    if engine:
        model.components.chain.set_engine(engine)
    rec, _ = model.collapsed_gibbs_sampler(case["n_iter"], z_true, num_saved=0, **case["kwargs"])
    return model, rec, z0
