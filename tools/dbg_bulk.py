import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import make_data, make_prior
from oracle import oracle as O
from pybgmm_b200 import _lib
N, D, K_true = 1500, 16, 6
X, _ = make_data(N, D, K_true, 1)
prior = make_prior(D)
z0 = O.init_assignments(N, "rand", K_true)
u = np.random.RandomState(3).random_sample(N)
orc = O.Oracle(X, *prior, K_max=88); orc.set_assignments(z0)
so = orc.sweep(u, 1.0)
for mode in ("replicated", "solo"):
    ch = _lib.Chain(X, *prior, 88); ch.set_engine("sequential"); ch.set_assignments(z0)
    try:
        if mode == "replicated":
            sg = ch.sweep(1.0, 1.0, None, u)
        else:
            import torch
            g = _lib.ChainGroup([ch])
            du = torch.from_numpy(u[None, :].copy()).cuda()
            sg = g.sweep_dev(1.0, 1.0, None, du)[0]
        print(mode, "ok K", sg.K, "moves", sg.moves, so.moves, "fast", sg.fast_steps)
    except Exception as e:
        print(mode, "EXC", e)
    st = ch.get_state(inv_covar=False, logdet=False)
    zdiff = (st["z"] != orc.assignments).sum()
    print("  labels differing:", zdiff, "counts equal:", (st["counts"] == orc.counts).all())
    K = orc.K
    dS = st["S_part"][:K] - orc.S_N_partials[:K]
    dm = st["m_num"][:K] - orc.m_N_numerators[:K]
    for k in range(K):
        print("  comp %d n=%d  max|dS|=%.3e  max|dnum|=%.3e   |S|max=%.3e" % (k, orc.counts[k], np.abs(dS[k]).max(), np.abs(dm[k]).max(), np.abs(orc.S_N_partials[k]).max()))
