"""CRP mixture model: collapsed Gibbs sampling on the GPU (mirror of pybgmm/igmm/crpmm.py:15-94)."""
import time

from .igmm import IGMM


class CRPMM(IGMM):

    def __init__(self, X, kernel_prior, alpha, save_path, assignments="rand", K=1, K_max=None,
                 covariance_type="full", device=0):
        super(CRPMM, self).__init__(X, kernel_prior, alpha, save_path, assignments=assignments, K=K, K_max=K_max,
                                    covariance_type=covariance_type, device=device)

    def collapsed_gibbs_sampler(self, n_iter, true_assignments, num_saved=3, weight_first=True, rng="reference"):
        """`n_iter` sweeps (crpmm.py:23-94).  Each sweep's per-datum loop (crpmm.py:57-88) is one C-ABI call.
        Returns (record_dict, distribution_dict)."""
        record_dict = self.setup_record_dict()
        start_time = time.time()
        distribution_dict = self.setup_distribution_dict(num_saved)
        for i_iter in range(n_iter):
            if num_saved == self.components.K and i_iter > 1:
                distribution_dict = self.update_distribution_dict(distribution_dict, weight_first)
            self._device_sweep(power=1.0, order=None, rng=rng)
            record_dict = self.update_record_dict(record_dict, i_iter, true_assignments, start_time)
            start_time = time.time()
        return record_dict, distribution_dict

    gibbs_sample = collapsed_gibbs_sampler  # the name north_star and the reference's older tests use
