"""Developer probe: cost of the per-sweep record (GMM.update_record_dict) at C3 through the class surface --
PCRPMM(...).collapsed_gibbs_sampler(n_iter, z_true) -- with the clustering metrics counted on the device
(bgmm_contingency / bgmm_cluster_ssq) and recounted from the labels on the host.  Not the bench."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import pybgmm_b200 as P  # noqa: E402
from conftest import make_data, make_prior  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=1000000)
ap.add_argument("--D", type=int, default=16)
ap.add_argument("--K", type=int, default=100)
ap.add_argument("--power", type=float, default=1.5)
ap.add_argument("--sweeps", type=int, default=8)
a = ap.parse_args()

out = {}
for backend in ("device", "host"):
    X, z_true = make_data(a.N, a.D, a.K, 1)
    m_0, k_0, v_0, S_0 = make_prior(a.D)
    model = P.PCRPMM(X, P.NIW(m_0, k_0, v_0, S_0), 1.0, None, assignments="rand", K=a.K, K_max=4 * a.K + 64)
    model.metrics_backend = backend
    marks = []
    orig = model.update_record_dict

    def timed(rec, i_iter, true_assignments, start_time, _orig=orig, _marks=marks):
        t0 = time.time()
        r = _orig(rec, i_iter, true_assignments, start_time)
        _marks.append(time.time() - t0)
        return r
    model.update_record_dict = timed
    t0 = time.time()
    rec, _ = model.collapsed_gibbs_sampler(a.sweeps, z_true, n_power=a.power, num_saved=0)
    wall = time.time() - t0
    out[backend] = {"wall_s": wall, "sweep_s": rec["sample_time"], "record_s": marks, "nmi": rec["nmi"],
                    "vi": rec["vi"], "loss": [float(v) for v in rec["loss"]], "K": rec["components"],
                    "log_marg": rec["log_marg"]}
    print(backend, "wall %.2fs; per sweep: sweep %s | record %s" % (
        wall, ["%.3f" % v for v in rec["sample_time"]], ["%.4f" % v for v in marks]), flush=True)
same = all(np.allclose(out["device"][k], out["host"][k], rtol=1e-12, atol=1e-14) for k in ("nmi", "vi", "log_marg"))
same = same and out["device"]["loss"] == out["host"]["loss"] and out["device"]["K"] == out["host"]["K"]
print(json.dumps({"record_probe": {"N": a.N, "D": a.D, "K_true": a.K, "sweeps": a.sweeps, "records_agree": bool(same),
                                   "record_ms_device_median": 1e3 * float(np.median(out["device"]["record_s"])),
                                   "record_ms_host_median": 1e3 * float(np.median(out["host"]["record_s"])),
                                   "sweep_ms_last": 1e3 * out["device"]["sweep_s"][-1]}}))
