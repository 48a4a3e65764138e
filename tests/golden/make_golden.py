#!/usr/bin/env python
"""Generate tests/golden/golden.json by running the REFERENCE itself (mechanically shimmed to Python 3 by
oracle/make_ref.py into oracle/_ref/) on small seeded inputs.  Run in the build container (where /root/reference
exists):  python tests/golden/make_golden.py

Every case records the exact inputs (or the seeds that generate them) and the reference's outputs, so the tests can
replay them against the CPU oracle (-m "not gpu") and against the CUDA engine (-m gpu) on a box that has neither
/root/reference nor oracle/_ref.
"""
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.make_ref import build as build_ref, import_ref, import_ref_adapcrpmm  # noqa: E402

if os.path.isdir("/root/reference/pybgmm"):
    build_ref("/root/reference", os.path.join(ROOT, "oracle", "_ref"))
NIW, CRPMM, PCRPMM, GaussianComponents, GaussianComponentsDiag = import_ref()
ADAPCRPMM = import_ref_adapcrpmm()


def gen(N, D, K_true, seed):
    """examples/crpmm_2d_demo.py:41-55 == pybgmm/tests/test_igmm.py:21-37."""
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


def sampler_case(name, cls, N, D, K_true, seed, cov, assignments, K, n_iter, v_0=None, K_max=None, **kw):
    X, z_true = gen(N, D, K_true, seed)
    m_0, k_0, v_0, S_0 = prior_for(D, cov, v_0)
    model = cls(X, NIW(m_0, k_0, v_0, S_0), 1.0, None, assignments=assignments, K=K, K_max=K_max,
                covariance_type=cov)
    z0 = model.components.assignments.copy()
    rec, _ = model.collapsed_gibbs_sampler(n_iter, z_true, num_saved=0, **kw)
    c = model.components
    return {
        "name": name, "cls": cls.__name__, "N": N, "D": D, "K_true": K_true, "seed": seed, "cov": cov,
        "assignments": assignments, "K_init": K, "n_iter": n_iter, "v_0": v_0, "K_max": K_max, "kwargs": kw,
        "z0": z0.tolist(), "z": c.assignments.tolist(), "K": int(c.K), "counts": c.counts[:c.K].tolist(),
        "log_marg": float(model.log_marg()), "log_marg_trace": [float(v) for v in rec["log_marg"]],
        "K_trace": [int(v) for v in rec["components"]],
        "nmi": float(rec["nmi"][-1]), "mi": float(rec["mi"][-1]), "vi": float(rec["vi"][-1]),
        "loss": float(rec["loss"][-1]),
    }


def components_case(name, cov, N, D, K_true, seed, z):
    """Leaf functions of the components object on a fixed assignment."""
    X, _ = gen(N, D, K_true, seed)
    m_0, k_0, v_0, S_0 = prior_for(D, cov)
    cls = GaussianComponents if cov == "full" else GaussianComponentsDiag
    c = cls(X, NIW(m_0, k_0, v_0, S_0), np.array(z), K_max=16)
    out = {
        "name": name, "cov": cov, "N": N, "D": D, "K_true": K_true, "seed": seed, "z": list(map(int, z)),
        "log_prior": [float(c.log_prior(i)) for i in range(N)],
        "log_post_pred": [c.log_post_pred(i).tolist() for i in range(N)],
        "log_marg_k": [float(c.log_marg_k(k)) for k in range(c.K)],
        "log_marg": float(c.log_marg()),
        "m_N_numerators": c.m_N_numerators[:c.K].tolist(),
        "S_N_partials": c.S_N_partials[:c.K].tolist(),
        "logdet": (c.logdet_covars if cov == "full" else c.log_prod_vars)[:c.K].tolist(),
        "inv": (c.inv_covars if cov == "full" else c.inv_vars)[:c.K].tolist(),
    }
    # then a scripted add/del sequence (exercises del_component's swap-with-last, gaussian_components.py:188-205)
    rng = np.random.RandomState(7)
    ops = []
    for _ in range(30):
        i = int(rng.randint(N))
        c.del_item(i)
        k = int(rng.randint(c.K + 1))
        c.add_item(i, k)
        ops.append([i, k])
    out.update(ops=ops, z_after=c.assignments.tolist(), K_after=int(c.K), counts_after=c.counts[:c.K].tolist(),
               m_after=c.m_N_numerators[:c.K].tolist(), S_after=c.S_N_partials[:c.K].tolist(),
               logdet_after=(c.logdet_covars if cov == "full" else c.log_prod_vars)[:c.K].tolist())
    return out


def main():
    cases = {"samplers": [], "components": []}
    S = cases["samplers"]
    # the reference's own end-to-end goldens (pybgmm/tests/test_igmm.py:17-190), regenerated through the reference
    S.append(sampler_case("G1_G2_test_igmm_sampling_2d", CRPMM, 100, 2, 4, 1, "full", "rand", 3, 10, v_0=5))
    S.append(sampler_case("G3_test_igmm_each_in_own", CRPMM, 20, 2, 4, 1, "full", "each-in-own", 3, 1, v_0=5))
    S.append(sampler_case("G4_test_igmm_log_marg_each_in_own", CRPMM, 5, 2, 4, 2, "full", "each-in-own", 3, 1, v_0=5))
    # the demos (BASELINE.json configs[0]): examples/crpmm_1d_demo.py as written, and its N=300 NIX variant
    S.append(sampler_case("C1_crpmm_1d_demo", CRPMM, 100, 1, 4, 1, "full", "rand", 3, 40))
    S.append(sampler_case("C1_crpmm_1d_nix_300", CRPMM, 300, 1, 4, 1, "diag", "rand", 3, 100))
    S.append(sampler_case("crpmm_one_by_one_2d", CRPMM, 60, 2, 3, 3, "full", "one-by-one", 1, 4))
    S.append(sampler_case("crpmm_diag_3d", CRPMM, 150, 3, 4, 4, "diag", "rand", 5, 8))
    S.append(sampler_case("crpmm_full_5d", CRPMM, 200, 5, 4, 5, "full", "rand", 6, 6, K_max=64))
    S.append(sampler_case("pcrpmm_2d_r1.5", PCRPMM, 200, 2, 4, 6, "full", "rand", 5, 8, n_power=1.5, power_burnin=0))
    S.append(sampler_case("pcrpmm_2d_burnin2", PCRPMM, 150, 2, 4, 7, "full", "rand", 4, 6, n_power=1.2, power_burnin=2))
    S.append(sampler_case("pcrpmm_flag_off", PCRPMM, 120, 2, 3, 8, "full", "rand", 4, 4, flag_power=False))
    S.append(sampler_case("pcrpmm_diag_4d", PCRPMM, 160, 4, 4, 9, "diag", "rand", 5, 6, n_power=1.5, power_burnin=0))
    C = cases["components"]
    z11 = [0, 0, 0, 1, 0, 1, 3, 4, 3, 2, -1]  # pybgmm/tests/test_gaussian_components.py:120
    C.append(components_case("full_4d_test_vector", "full", 11, 4, 3, 1, z11))
    C.append(components_case("diag_4d_test_vector", "diag", 11, 4, 3, 1, z11))
    C.append(components_case("full_1d", "full", 24, 1, 3, 2, [i % 4 for i in range(24)]))
    C.append(components_case("diag_2d", "diag", 24, 2, 3, 3, [i % 3 for i in range(24)]))
    C.append(components_case("full_16d", "full", 60, 16, 3, 4, [i % 2 for i in range(60)]))
    with open(os.path.join(HERE, "golden.json"), "w") as fh:
        json.dump(cases, fh)
    print("wrote", os.path.join(HERE, "golden.json"), os.path.getsize(os.path.join(HERE, "golden.json")), "bytes")
    adaptive()
    fixedvar()


def adaptive():
    """ADAPCRPMM (pybgmm/igmm/adapcrpmm.py) cases, in their own file.  adapcrp_burnin=-1 throughout: the reference
    stops with UnboundLocalError in its first sweep for any burn-in >= 0 (adapcrpmm.py:110)."""
    A = []
    A.append(sampler_case("adapcrpmm_2d", ADAPCRPMM, 200, 2, 4, 3, "full", "rand", 6, 8, adapcrp_burnin=-1))
    A.append(sampler_case("adapcrpmm_2d_r2_perct10", ADAPCRPMM, 240, 2, 5, 11, "full", "rand", 12, 8, r_up=2.0,
                          adapcrp_perct=0.1, adapcrp_burnin=-1))
    A.append(sampler_case("adapcrpmm_diag_3d", ADAPCRPMM, 180, 3, 4, 12, "diag", "rand", 8, 6, r_up=1.5,
                          adapcrp_perct=0.08, adapcrp_burnin=-1))
    A.append(sampler_case("adapcrpmm_each_in_own", ADAPCRPMM, 60, 2, 3, 13, "full", "each-in-own", 1, 4,
                          adapcrp_burnin=-1))
    A.append(sampler_case("adapcrpmm_flag_off", ADAPCRPMM, 120, 2, 3, 14, "full", "rand", 4, 4, flag_adapcrp=False))
    with open(os.path.join(HERE, "golden_adap.json"), "w") as fh:
        json.dump({"samplers": A}, fh)
    print("wrote golden_adap.json", len(A), "cases")


def fixedvar():
    """Fixed-variance components (pybgmm/gaussian/gaussian_components_fixedvar.py) and CRPMM / PCRPMM over them
    (covariance_type="fixed", igmm.py:108-109): leaf values and sampler runs of the reference, for oracle/fixedvar.py."""
    from pybgmm.gaussian.gaussian_components_fixedvar import GaussianComponentsFixedVar, FixedVarPrior
    out = {"components": [], "samplers": []}
    for name, N, D, K_true, seed in (("fixed_3d", 30, 3, 3, 5), ("fixed_1d", 24, 1, 3, 6)):
        X, _ = gen(N, D, K_true, seed)
        rng = np.random.RandomState(seed)
        var, mu_0, var_0 = 0.3 + rng.rand(D), rng.randn(D), 4.0 + 8.0 * rng.rand(D)
        z = [i % 4 for i in range(N)]
        z[-1] = -1
        c = GaussianComponentsFixedVar(X, FixedVarPrior(var, mu_0, var_0), np.array(z), K_max=16)
        case = {"name": name, "N": N, "D": D, "K_true": K_true, "seed": seed, "var": var.tolist(), "mu_0": mu_0.tolist(),
                "var_0": var_0.tolist(), "z": z,
                "log_prior": [float(c.log_prior(i)) for i in range(N)],
                "log_post_pred": [c.log_post_pred(i).tolist() for i in range(N)],
                "log_marg_k": [float(c.log_marg_k(k)) for k in range(c.K)]}
        ops = []
        for _ in range(40):
            i = int(rng.randint(N))
            c.del_item(i)
            k = int(rng.randint(c.K + 1))
            c.add_item(i, k)
            ops.append([i, k])
        case.update(ops=ops, z_after=c.assignments.tolist(), K_after=int(c.K), counts_after=c.counts[:c.K].tolist(),
                    num_after=c.mu_N_numerators[:c.K].tolist(), prec_after=c.precision_Ns[:c.K].tolist(),
                    lpp_after=c.log_prod_precision_preds[:c.K].tolist())
        out["components"].append(case)
    for name, cls, N, D, K_true, seed, init, K, n_iter, kw in (
            ("crpmm_fixed_2d", CRPMM, 120, 2, 3, 3, "rand", 5, 6, {}),
            ("crpmm_fixed_each_in_own", CRPMM, 40, 2, 3, 4, "each-in-own", 1, 3, {}),
            ("pcrpmm_fixed_3d_r1.5", PCRPMM, 150, 3, 4, 5, "rand", 6, 6, {"n_power": 1.5, "power_burnin": 0})):
        X, z_true = gen(N, D, K_true, seed)
        var, mu_0, var_0 = 0.49 * np.ones(D), np.zeros(D), 16.0 * np.ones(D)
        model = cls(X, FixedVarPrior(var, mu_0, var_0), 1.0, None, assignments=init, K=K, covariance_type="fixed")
        z0 = model.components.assignments.copy()
        rec, _ = model.collapsed_gibbs_sampler(n_iter, z_true, num_saved=0, **kw)
        c = model.components
        out["samplers"].append({"name": name, "cls": cls.__name__, "N": N, "D": D, "K_true": K_true, "seed": seed,
                                "assignments": init, "K_init": K, "n_iter": n_iter, "kwargs": kw,
                                "var": var.tolist(), "mu_0": mu_0.tolist(), "var_0": var_0.tolist(),
                                "z0": z0.tolist(), "z": c.assignments.tolist(), "K": int(c.K),
                                "counts": c.counts[:c.K].tolist(), "log_marg": float(model.log_marg()),
                                "K_trace": [int(v) for v in rec["components"]],
                                "log_marg_trace": [float(v) for v in rec["log_marg"]]})
    with open(os.path.join(HERE, "golden_fixedvar.json"), "w") as fh:
        json.dump(out, fh)
    print("wrote golden_fixedvar.json", len(out["components"]), "+", len(out["samplers"]), "cases")


if __name__ == "__main__":
    main()
