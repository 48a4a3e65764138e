"""Adaptive powered-CRP mixture model on the GPU engine.

API of `pybgmm/igmm/adapcrpmm.py:20-219`.  A host policy over the same sweep kernel as PCRPMM: the power of the count
prior is recomputed before every sweep from the current cluster sizes,

    power = 1 + (r_up - 1) * (fraction of clusters with at most N * adapcrp_perct members)     (adapcrpmm.py:100-104)

the scan order is a fresh `np.random.permutation(N)` whenever that power exceeds 1 (`:110-115`), and the count prior
of the sweep is log(n_k ** power) (`:132-134`).  The cluster sizes are read from the device after the previous sweep
(K_max int64 -- the labels stay on the GPU).

One deliberate difference: with `flag_adapcrp=True` the reference only defines the power once `i_iter >
adapcrp_burnin` and reads it unconditionally (`:110`), so any `adapcrp_burnin >= 0` -- including the default 0 --
dies with UnboundLocalError in the very first sweep.  Here the burn-in sweeps run as its docstring describes them
(`:48-49`: "iteration<adapcrp_burnin will set power value to 1 (i.e. CRPMM)"): power 1, data order.  For
`adapcrp_burnin < 0` (the only setting the reference runs with) the chains are identical (tests/golden).
"""
import logging

import numpy as np

from .igmm import IGMM

logger = logging.getLogger(__name__)


class ADAPCRPMM(IGMM):

    def collapsed_gibbs_sampler(self, n_iter, true_assignments, r_up=1.3, adapcrp_perct=0.04, adapcrp_burnin=0,
                                num_saved=3, weight_first=True, flag_adapcrp=True, rng="reference"):
        self.adapcrp_powers = []   # the power used by each sweep (extra; for inspection)

        def schedule(i_iter, n_points):
            power = 1.0
            if flag_adapcrp and i_iter > adapcrp_burnin:
                n_k = self.components.counts[:self.components.K]
                small = np.count_nonzero(n_k <= n_points * adapcrp_perct) * 1.0 / len(n_k)
                power = 1.0 + (r_up - 1.0) * small
                if i_iter % 20 == 0:
                    logger.info("Ada-pCRP power: %s", power)
            order = np.random.permutation(n_points) if (flag_adapcrp and power > 1) else None
            self.adapcrp_powers.append(power)
            return order, power

        return self._run_sweeps(n_iter, true_assignments, num_saved, weight_first, schedule, rng)

    gibbs_sample = collapsed_gibbs_sampler
