"""Developer probe: wall time of the first sweeps of a chain (errors of the trailing record rebuild are ignored: used with
BGMM_TUNE switches that make the statistics invalid on purpose)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import make_data, make_prior
from pybgmm_b200 import _lib
N, D, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
X, _ = make_data(N, D, K, 1)
prior = make_prior(D)
z0 = np.random.randint(0, K, N).astype(np.int64)
ch = _lib.Chain(X, *prior, 4 * K + 64)
ch.set_assignments(z0)
ch.set_engine("sequential")
for s in range(2):
    order = np.random.permutation(N)
    u = np.random.random_sample(N)
    t = time.time()
    try:
        st = ch.sweep(1.0, 1.0, order, u)
        msg = "moves %d fast %d kernel_ms %.1f" % (st.moves, st.fast_steps, st.sweep_kernel_ms)
    except Exception as e:
        msg = "EXC " + str(e)[:60]
    print("TUNE=%s sweep %d wall %.3f s  %s" % (os.environ.get("BGMM_TUNE", "0"), s, time.time() - t, msg), flush=True)
    if "EXC" in msg:
        break
