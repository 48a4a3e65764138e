"""Fixed-variance components whose state lives on the GPU.

Mirror of the class surface of pybgmm/gaussian/gaussian_components_fixedvar.py:16-296 (`GaussianComponentsFixedVar`) and
of `FixedVarPrior` (:304-311): components share a known diagonal variance `prior.var`; their means have independent
normal priors N(prior.mu_0, prior.var_0).  Same attribute protocol as the reference (`mu_N_numerators, precision_Ns,
log_prod_precision_preds, precision_preds, counts, assignments, K, cached_log_prior`); the computation is done by
libbgmm_b200.so (bgmm_create_fixedvar; generic sweep engine, COV_FIXED).  Selected by `covariance_type="fixed"`
(pybgmm/igmm/igmm.py:108-109).
"""
import numpy as np

from .. import _lib
from .gaussian_components import _DeviceComponents


class FixedVarPrior(object):
    """The prior parameters for a fixed diagonal covariance multivariate Gaussian (gaussian_components_fixedvar.py:304)."""

    def __init__(self, var, mu_0, var_0):
        self.var = var
        self.mu_0 = mu_0
        self.var_0 = var_0


class GaussianComponentsFixedVar(_DeviceComponents):
    _COV = "fixed"

    def _check_prior(self):
        for v in (self.prior.var, self.prior.mu_0, self.prior.var_0):
            assert np.broadcast_to(np.asarray(v, dtype=float), (self.D,)).shape == (self.D,)

    def _make_chain(self, X, device):
        pr = self.prior
        return _lib.Chain.fixed_variance(X, pr.var, pr.mu_0, pr.var_0, self.K_max, device=device)

    # the reference's attribute names for this variant (gaussian_components_fixedvar.py:84-92)
    @property
    def precision(self):
        return 1. / np.broadcast_to(np.asarray(self.prior.var, dtype=float), (self.D,))

    @property
    def mu_0(self):
        return np.broadcast_to(np.asarray(self.prior.mu_0, dtype=float), (self.D,))

    @property
    def precision_0(self):
        return 1. / np.broadcast_to(np.asarray(self.prior.var_0, dtype=float), (self.D,))

    @property
    def mu_N_numerators(self):
        return self._state()["m_num"]

    @property
    def precision_Ns(self):
        return self._state()["S_part"]

    @property
    def log_prod_precision_preds(self):
        return self._state()["logdet"]

    @property
    def precision_preds(self):
        return self._state()["inv_covar"]

    def cache_component_stats(self, k):
        s = self._state()
        return (s["m_num"][k].copy(), s["S_part"][k].copy(), s["logdet"][k], s["inv_covar"][k].copy(),
                int(s["counts"][k]))

    def cluster_ssq(self):
        """Not a function of this variant's statistics (they hold no sum of squares): the caller counts from labels."""
        raise NotImplementedError

    def rand_k(self, k):
        """A mean vector from the posterior product of normals of component k (gaussian_components_fixedvar.py:268-276)."""
        mu_N = self.mu_N_numerators[k] / self.precision_Ns[k]
        var_N = 1. / self.precision_Ns[k]
        return np.array([np.random.normal(mu_N[i], np.sqrt(var_N[i])) for i in range(self.D)])

    def map(self, k):
        """Posterior mean of component k and the (known) data variance."""
        return self.mu_N_numerators[k] / self.precision_Ns[k], 1. / self.precision
