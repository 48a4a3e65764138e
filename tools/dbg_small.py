"""Developer probe: one small chain against the oracle, prints the first disagreement."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from pybgmm_b200 import _lib
from oracle import oracle as O
from conftest import make_data, make_prior
N, D, K, engine = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
sweeps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
X, zt = make_data(N, D, K, 1)
m0, k0, v0, S0 = make_prior(D)
z0 = O.init_assignments(N, "rand", K)
Kmax = min(N, 4 * K + 64)
orc = O.Oracle(X, m0, k0, v0, S0, K_max=Kmax); orc.set_assignments(z0)
ch = _lib.Chain(X, m0, k0, v0, S0, Kmax); ch.set_assignments(z0); ch.set_engine(engine)
rng = np.random.RandomState(5)
for s in range(sweeps):
    u = rng.random_sample(N)
    so = orc.sweep(u, 1.0)
    sg = ch.sweep(1.0, 1.0, None, u)
    same = (ch.assignments() == orc.assignments).all()
    print(s, "oracle", (so.K_end, so.moves, so.births, so.deaths, so.evals), "gpu", sg.as_dict(), "same", same, flush=True)
