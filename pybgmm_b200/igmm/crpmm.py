"""CRP mixture model on the GPU engine.

API of `pybgmm/igmm/crpmm.py:15-94` (`CRPMM(X, kernel_prior, alpha, save_path, assignments, K, K_max,
covariance_type)`, `collapsed_gibbs_sampler(n_iter, true_assignments, num_saved, weight_first)` returning
`(record_dict, distribution_dict)`); the sweep loop itself is `IGMM._run_sweeps`, shared with PCRPMM.
"""
from .igmm import IGMM


def _plain_crp(i_iter, n_points):
    """Scan order and count-prior power of sweep `i_iter`: data order 0..N-1 (crpmm.py:57), log(n_k) (crpmm.py:70)."""
    return None, 1.0


class CRPMM(IGMM):

    def collapsed_gibbs_sampler(self, n_iter, true_assignments, num_saved=3, weight_first=True, rng="reference"):
        """Run `n_iter` Gibbs sweeps; every sweep is one call into the C-ABI (`bgmm_sweep`), which performs the
        reference's per-datum loop (crpmm.py:57-88) on the device."""
        return self._run_sweeps(n_iter, true_assignments, num_saved, weight_first, _plain_crp, rng)

    # the reference's older tests and BASELINE.json's north_star call it gibbs_sample
    gibbs_sample = collapsed_gibbs_sampler
