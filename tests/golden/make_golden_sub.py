#!/usr/bin/env python
"""Generate tests/golden/golden_sub.json by running the REFERENCE's SubCRPMM.collapsed_gibbs_sampler
(pybgmm/igmm/subcrpmm.py:352-457; shimmed to Python 3 into oracle/_ref by oracle/make_ref.py) on small seeded inputs
with both mask moves (Gibbs over the dimensions :306-337, one Metropolis flip :187-291) and with the mask frozen.
    python tests/golden/make_golden_sub.py
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
from pybgmm.igmm.subcrpmm import SubCRPMM  # noqa: E402
from pybgmm.prior.betabern import BetaBern  # noqa: E402


def gen(N, D_inf, D_noise, K_true, seed):
    """K_true clusters in the first D_inf dimensions, D_noise dimensions of pure noise."""
    random.seed(seed)
    np.random.seed(seed)
    z_true = np.random.randint(0, K_true, N)
    mu = np.random.randn(D_inf, K_true) * 4.0
    X = np.hstack([(mu[:, z_true] + np.random.randn(D_inf, N) * 0.7).T, np.random.randn(N, D_noise)])
    return np.ascontiguousarray(X), z_true


def case(name, N, D_inf, D_noise, K_true, seed, K, n_iter, bern, **kw):
    X, z_true = gen(N, D_inf, D_noise, K_true, seed)
    D = X.shape[1]
    prior = NIW(np.zeros(D), 0.05, D + 3, 0.5 * np.eye(D))
    model = SubCRPMM(X, prior, 1.0, None, assignments="rand", K=K, K_max=None, covariance_type="full",
                     bern_prior=BetaBern(*bern) if bern else None, p_bern=0.3)
    z0 = model.components.assignments.copy()
    rec, _, sub = model.collapsed_gibbs_sampler(n_iter, z_true, num_saved=0, **kw)
    c = model.components
    return {"name": name, "N": N, "D_inf": D_inf, "D_noise": D_noise, "K_true": K_true, "seed": seed, "K_init": K,
            "n_iter": n_iter, "bern": bern, "kwargs": kw, "z0": z0.tolist(), "z": c.assignments.tolist(), "K": int(c.K),
            "counts": c.counts[:c.K].tolist(), "mask": [int(v) for v in model.mask],
            "components_D": int(c.X.shape[1]), "p_bern": float(model.p_bern),
            "included": [int(v) for v in sub["included_variable"]], "K_trace": [int(v) for v in rec["components"]],
            "log_marg_trace": [float(v) for v in rec["log_marg"]], "log_marg": float(model.log_marg()),
            "common_log_marg": float(model.common_component.log_marg()),
            "tail": [random.random(), float(np.random.rand())]}   # where the global RNG streams were left


if __name__ == "__main__":
    cases = [
        case("gibbs_mask", 80, 2, 3, 3, 21, 5, 6, (1, 1), burnin_mask=1, mask_update="gibbs"),
        case("metropolis_mask", 90, 2, 2, 3, 22, 6, 10, (1, 1), burnin_mask=0, mask_update="metropolis"),
        case("metropolis_fixed_p", 70, 3, 2, 2, 23, 4, 8, None, burnin_mask=1, mask_update="metropolis"),
        case("mask_frozen", 100, 2, 2, 4, 24, 8, 4, (2, 1), burnin_mask=500, mask_update="gibbs"),
    ]
    with open(os.path.join(HERE, "golden_sub.json"), "w") as fh:
        json.dump({"cases": cases}, fh)
    for c in cases:
        print(c["name"], "K", c["K_trace"], "included", c["included"], "mask", c["mask"], "D'", c["components_D"])
