"""Powered-CRP mixture model (mirror of pybgmm/igmm/pcrpmm.py:20-192)."""
import logging
import time

import numpy as np

from .igmm import IGMM

logger = logging.getLogger(__name__)


class PCRPMM(IGMM):

    def __init__(self, X, kernel_prior, alpha, save_path, assignments="rand", K=1, K_max=None,
                 covariance_type="full", device=0):
        super(PCRPMM, self).__init__(X, kernel_prior, alpha, save_path, assignments=assignments, K=K, K_max=K_max,
                                     covariance_type=covariance_type, device=device)

    def collapsed_gibbs_sampler(self, n_iter, true_assignments, n_power=1.01, power_burnin=0, num_saved=3,
                                weight_first=True, flag_power=True, rng="reference"):
        """`n_iter` sweeps of the pCRP sampler (pcrpmm.py:29-192): random scan order drawn from np.random at the top
        of each sweep when the power is on (:86-91), count prior log(n_k ** n_power) once i_iter > power_burnin
        (:105-112, strict), the new-table weight stays log(alpha) (:116)."""
        record_dict = self.setup_record_dict()
        start_time = time.time()
        distribution_dict = self.setup_distribution_dict(num_saved)
        for i_iter in range(n_iter):
            if num_saved == self.components.K and i_iter > 1:
                distribution_dict = self.update_distribution_dict(distribution_dict, weight_first)
            if flag_power and n_power > 1:
                if i_iter % 20 == 0:
                    logger.info(" Permutate data; " + "Power value: {}".format(n_power))
                order = np.random.permutation(self.components.N)
            else:
                order = None
            power = n_power if (flag_power and i_iter > power_burnin) else 1.0
            self._device_sweep(power=power, order=order, rng=rng)
            record_dict = self.update_record_dict(record_dict, i_iter, true_assignments, start_time)
            start_time = time.time()
        return record_dict, distribution_dict

    gibbs_sample = collapsed_gibbs_sampler
