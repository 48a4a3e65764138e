"""CPU: the C oracle against the fixture generated from the reference itself at size (tests/golden/golden_large.json:
CRPMM N = 1e5 D = 2 -- BASELINE.json configs[1] -- and PCRPMM N = 3e4 D = 8 r = 1.5, two sweeps each)."""
import hashlib
import json
import os

import numpy as np
import pytest

import cases as C
from oracle import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))


def _cases():
    with open(os.path.join(HERE, "golden", "golden_large.json")) as fh:
        return json.load(fh)["cases"]


def _digest(z):
    return hashlib.sha256(np.ascontiguousarray(z, dtype="<i8").tobytes()).hexdigest()


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_oracle_reproduces_reference_at_size(case):
    X, _ = C.gen(case["N"], case["D"], case["K_true"], case["seed"])
    m_0, k_0, v_0, S_0 = C.prior_for(case["D"], "full")
    z0 = O.init_assignments(case["N"], "rand", case["K_true"])
    assert _digest(z0) == case["z0_sha256"]
    orc = O.Oracle(X, m_0, k_0, v_0, S_0, K_max=case["K_max"])
    orc.set_assignments(z0)
    for s, want in enumerate(case["sweeps"]):
        if case["cls"] == "CRPMM":
            O.run_crpmm(orc, 1, 1.0)
        else:
            O.run_pcrpmm(orc, 1, 1.0, n_power=case["kwargs"]["n_power"], power_burnin=(0 if s == 0 else -1))
        assert orc.K == want["K"]
        assert orc.counts[:orc.K].tolist() == want["counts"]
        assert _digest(orc.assignments) == want["z_sha256"]
        np.testing.assert_allclose(orc.log_marg(1.0), want["log_marg"], rtol=1e-9)
