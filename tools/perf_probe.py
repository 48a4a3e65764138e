"""Developer probe: per-sweep engine counters and device time for a synthetic workload (not the bench)."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pybgmm_b200 import _lib  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import make_data, make_prior  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=1000000)
ap.add_argument("--D", type=int, default=16)
ap.add_argument("--K", type=int, default=100)
ap.add_argument("--power", type=float, default=1.5)
ap.add_argument("--sweeps", type=int, default=8)
ap.add_argument("--cov", default="full")
ap.add_argument("--engine", default="adaptive")
ap.add_argument("--natural", action="store_true")
a = ap.parse_args()

t = time.time()
X, z_true = make_data(a.N, a.D, a.K, 1)
m_0, k_0, v_0, S_0 = make_prior(a.D, a.cov)
z0 = np.random.randint(0, a.K, a.N).astype(np.int64)
print("data %.1fs" % (time.time() - t), flush=True)
t = time.time()
ch = _lib.Chain(X, m_0, k_0, v_0, S_0, 4 * a.K + 64, covariance_type=a.cov)
ch.set_assignments(z0)
ch.set_engine(a.engine)
ch.seed(1)
print("create+build %.2fs K=%d" % (time.time() - t, ch.K), flush=True)
for s in range(a.sweeps):
    order = None if (a.natural or a.power <= 1) else np.random.permutation(a.N)
    t = time.time()
    st = ch.sweep(1.0, a.power if s > 0 else 1.0, order, None)
    dt = time.time() - t
    print("sweep %d: %s wall=%.3fs evals/s=%.3e" % (s, st.as_dict(), dt, st.evals / (st.device_ms * 1e-3)), flush=True)
    print("   phases(cycles):", st.phases(), flush=True)
