#!/usr/bin/env python
"""Generate tests/golden/golden_large.json: the REFERENCE itself (shimmed to Python 3 into oracle/_ref by
oracle/make_ref.py) run at BASELINE.json's configs[1] size -- CRPMM, NIW full covariance, N = 1e5, D = 2,
K_true = 30, `rand` initial assignments -- for two sweeps (pybgmm/igmm/crpmm.py:47-88).  A run takes a few
minutes of CPU (the reference is ~3e3 data/s) and needs /root/reference, so the outputs are committed as a small
fixture: the SHA-256 of the int64 assignment vector after every sweep, K, the counts and log_marg.  The tests replay
the same seeded inputs through the C oracle (-m "not gpu") and the CUDA engine (-m gpu) and compare.

    python tests/golden/make_golden_large.py
"""
import hashlib
import json
import os
import random
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.make_ref import build as build_ref, import_ref  # noqa: E402

if os.path.isdir("/root/reference/pybgmm"):
    build_ref("/root/reference", os.path.join(ROOT, "oracle", "_ref"))
NIW, CRPMM, PCRPMM, _, _ = import_ref()


def gen(N, D, K_true, seed):
    """examples/crpmm_2d_demo.py:41-55 (the same generator as make_golden.py / tests/cases.py)."""
    random.seed(seed)
    np.random.seed(seed)
    z_true = np.random.randint(0, K_true, N)
    mu = np.random.randn(D, K_true) * 4.0
    X = (mu[:, z_true] + np.random.randn(D, N) * 0.7).T
    return np.ascontiguousarray(X), z_true


def digest(z):
    return hashlib.sha256(np.ascontiguousarray(z, dtype="<i8").tobytes()).hexdigest()


def large_case(name, cls, N, D, K_true, seed, n_iter, K_max, **kw):
    X, z_true = gen(N, D, K_true, seed)
    v_0 = D + 3
    prior = NIW(np.zeros(D), 0.7 ** 2 / 4.0 ** 2, v_0, 0.7 ** 2 * v_0 * np.eye(D))
    t0 = time.time()
    model = cls(X, prior, 1.0, None, assignments="rand", K=K_true, K_max=K_max, covariance_type="full")
    model.update_record_dict = lambda rec, i, z, t: rec      # the per-sweep metrics are O(K_true K N) Python
    out = {"name": name, "cls": cls.__name__, "N": N, "D": D, "K_true": K_true, "seed": seed, "K_max": K_max,
           "n_iter": n_iter, "kwargs": kw, "z0_sha256": digest(model.components.assignments), "sweeps": []}
    # one sweep per call continues the chain exactly (the RNG streams are global); PCRPMM's power applies from the
    # second sweep (i_iter > power_burnin, pcrpmm.py:105): power_burnin=-1 keeps it on in a one-sweep call
    for s in range(n_iter):
        if cls is CRPMM:
            model.collapsed_gibbs_sampler(1, z_true, num_saved=0)
        else:
            model.collapsed_gibbs_sampler(1, z_true, num_saved=0, power_burnin=(0 if s == 0 else -1), **kw)
        c = model.components
        out["sweeps"].append({"z_sha256": digest(c.assignments), "K": int(c.K), "counts": c.counts[:c.K].tolist(),
                              "log_marg": float(model.log_marg())})
        print(name, "sweep", s, "K", c.K, "%.0f s" % (time.time() - t0), flush=True)
    return out


if __name__ == "__main__":
    cases = [
        large_case("C2_crpmm_N1e5_D2", CRPMM, 100000, 2, 30, 1, 2, 184),
        large_case("pcrpmm_N3e4_D8_r1.5", PCRPMM, 30000, 8, 40, 3, 2, 224, n_power=1.5),
    ]
    with open(os.path.join(HERE, "golden_large.json"), "w") as fh:
        json.dump({"cases": cases}, fh, indent=1)
    print("wrote golden_large.json")
