"""Powered-CRP mixture model on the GPU engine.

API of `pybgmm/igmm/pcrpmm.py:20-192`.  What differs from CRPMM is decided on the host, per sweep, exactly as the
reference decides it:
  * the scan order is a fresh `np.random.permutation(N)` at the top of every sweep while the power is enabled
    (`flag_power and n_power > 1`, pcrpmm.py:86-91) -- drawn from the global NumPy stream, like the reference;
  * the count prior is log(n_k ** n_power) once `i_iter > power_burnin` (strictly; pcrpmm.py:105-112), log(n_k) before;
  * the new-table weight stays log(alpha) (pcrpmm.py:116).
"""
import logging

import numpy as np

from .igmm import IGMM

logger = logging.getLogger(__name__)


class PCRPMM(IGMM):

    def collapsed_gibbs_sampler(self, n_iter, true_assignments, n_power=1.01, power_burnin=0, num_saved=3,
                                weight_first=True, flag_power=True, rng="reference"):
        powered_scan = bool(flag_power) and n_power > 1

        def schedule(i_iter, n_points):
            order = None
            if powered_scan:
                if i_iter % 20 == 0:
                    logger.info("sweep %d: random scan order, power %s", i_iter, n_power)
                order = np.random.permutation(n_points)
            use_power = bool(flag_power) and i_iter > power_burnin
            return order, (n_power if use_power else 1.0)

        return self._run_sweeps(n_iter, true_assignments, num_saved, weight_first, schedule, rng)

    gibbs_sample = collapsed_gibbs_sampler
