#!/usr/bin/env python
"""Generate tests/golden/golden_cs.json by running the REFERENCE's CSCRPMM.constrained_gibbs_sample
(pybgmm/igmm/cscrpmm.py:96-485; shimmed to Python 3 into oracle/_ref by oracle/make_ref.py) on small seeded inputs:
constrained sweeps (flag_constrain: the re-draw of cscrpmm.py:342-350), the approximate step (flag_approx, :418-461),
the per-sweep adaptive power (flag_adapcrp_form2, :263-269) and the plain powered sweep (flag_power).
    python tests/golden/make_golden_cs.py
"""
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.make_ref import build as build_ref, import_ref  # noqa: E402

if os.path.isdir("/root/reference/pybgmm"):
    build_ref("/root/reference", os.path.join(ROOT, "oracle", "_ref"))
NIW = import_ref()[0]
from pybgmm.igmm import CSCRPMM  # noqa: E402


def gen(N, D, K_true, seed):
    random.seed(seed)
    np.random.seed(seed)
    z_true = np.random.randint(0, K_true, N)
    mu = np.random.randn(D, K_true) * 4.0
    X = (mu[:, z_true] + np.random.randn(D, N) * 0.7).T
    return np.ascontiguousarray(X), z_true


def case(name, N, D, K_true, seed, K, n_iter, **kw):
    X, z_true = gen(N, D, K_true, seed)
    v_0 = D + 3
    prior = NIW(np.zeros(D), 0.7 ** 2 / 4.0 ** 2, v_0, 0.7 ** 2 * v_0 * np.eye(D))
    model = CSCRPMM(X, prior, 1.0, None, assignments="rand", K=K, K_max=None, covariance_type="full")
    z0 = model.components.assignments.copy()
    rec, _ = model.constrained_gibbs_sample(n_iter, z_true, num_saved=0, **kw)
    c = model.components
    return {"name": name, "N": N, "D": D, "K_true": K_true, "seed": seed, "K_init": K, "n_iter": n_iter, "kwargs": kw,
            "z0": z0.tolist(), "z": c.assignments.tolist(), "K": int(c.K), "counts": c.counts[:c.K].tolist(),
            "log_marg": float(model.log_marg()), "K_trace": [int(v) for v in rec["components"]],
            "log_marg_trace": [float(v) for v in rec["log_marg"]],
            "tail": [random.random(), float(np.random.rand())]}   # where the global RNG streams were left


if __name__ == "__main__":
    cases = [
        case("constrain_every_2", 200, 2, 4, 11, 12, 6, flag_constrain=True, n_constrain=2, thres=0.04),
        case("constrain_each_sweep_3d", 150, 3, 3, 12, 10, 4, flag_constrain=True, n_constrain=1, thres=0.05),
        case("constrain_with_power", 200, 2, 4, 13, 12, 6, flag_constrain=True, n_constrain=3, thres=0.03,
             flag_power=True, n_power=1.4, power_burnin=0),
        case("approx_step", 180, 2, 4, 14, 12, 5, flag_approx=True, approx_thres_perct=0.04, approx_burnin=1),
        case("adapcrp_form2", 200, 2, 4, 15, 12, 6, flag_adapcrp_form2=True, r_up=1.5, adapcrp_perct=0.05, adapcrp_burnin=1),
        case("plain_power", 160, 2, 3, 16, 8, 4, flag_power=True, n_power=1.3, power_burnin=1),
    ]
    with open(os.path.join(HERE, "golden_cs.json"), "w") as fh:
        json.dump({"cases": cases}, fh)
    print("wrote golden_cs.json:", [(c["name"], c["K"]) for c in cases])
