"""Diagonal-covariance (product of normal-inverse-chi^2) components, device backed.

Mirror of pybgmm/gaussian/gaussian_components_diag.py (class surface :84-353); `S_0` is a D-vector (:92).
"""
import numpy as np

from .gaussian_components import _DeviceComponents


class GaussianComponentsDiag(_DeviceComponents):
    _COV = "diag"

    def _check_prior(self):
        assert np.asarray(self.prior.S_0).shape == (self.D,)  # gaussian_components_diag.py:92

    # attribute names of the diagonal variant (gaussian_components_diag.py:96-97)
    @property
    def log_prod_vars(self):
        return self._state()["logdet"]

    @property
    def inv_vars(self):
        return self._state()["inv_covar"]

    def map(self, k):
        """MAP mean and variances of component k (gaussian_components_diag.py:310-322)."""
        k_N = self.prior.k_0 + self.counts[k]
        v_N = self.prior.v_0 + self.counts[k]
        m_N = self.m_N_numerators[k] / k_N
        var = (self.S_N_partials[k] - k_N * np.square(m_N)) / (v_N + 2)
        return m_N, var
