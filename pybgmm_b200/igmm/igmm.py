"""IGMM base: initial assignments, component-class dispatch, log_marg (mirror of pybgmm/igmm/igmm.py:38-227)."""
import logging
import math

import numpy as np
from scipy import stats
from scipy.special import gammaln

from .. import _lib
from ..gaussian import GaussianComponents, GaussianComponentsDiag
from ..gmm import GMM

logger = logging.getLogger(__name__)


class IGMM(GMM):
    """Infinite Gaussian mixture model on the GPU engine.  Constructor arguments as pybgmm/igmm/igmm.py:68-71;
    `device` (extra, keyword only) selects the CUDA device."""

    def __init__(self, X, kernel_prior, alpha, save_path, assignments="rand", K=1, K_max=None,
                 covariance_type="full", device=0):
        super(IGMM, self).__init__()
        data_shape = X.shape
        if len(data_shape) < 2:
            raise ValueError('X must be at least a 2-dimensional array.')
        self.save_path = save_path
        self.alpha = alpha
        self.N, self.D = X.shape

        # Initial component assignments (igmm.py:86-102); "rand" consumes np.random exactly like the reference
        if isinstance(assignments, str) and assignments == "rand":
            assignments = np.random.randint(0, K, self.N)
            for k in range(assignments.max()):  # make the labels consecutive
                while len(np.nonzero(assignments == k)[0]) == 0:
                    assignments[np.where(assignments > k)] -= 1
                if assignments.max() == k:
                    break
        elif isinstance(assignments, str) and assignments == "one-by-one":
            assignments = -1 * np.ones(self.N, dtype="int")
            assignments[0] = 0
        elif isinstance(assignments, str) and assignments == "each-in-own":
            assignments = np.arange(self.N)
        else:
            pass  # a vector

        if covariance_type == "full":
            self.components = GaussianComponents(X, kernel_prior, assignments, K_max, device=device)
        elif covariance_type == "diag":
            self.components = GaussianComponentsDiag(X, kernel_prior, assignments, K_max, device=device)
        elif covariance_type == "fixed":
            raise NotImplementedError("fixed-variance components are not on the accelerated path (SURVEY.md 8f)")
        else:
            assert False, "Invalid covariance type."
        self.last_sweep_stats = None

    # ---- distribution dict (igmm.py:115-197); plotting is not reproduced --------------------------------
    def setup_distribution_dict(self, num_saved):
        return {"mean": np.zeros(shape=(num_saved, 0)), "variance": np.zeros(shape=(num_saved, 0)),
                "weights": np.zeros(shape=(num_saved, 0))}

    def update_distribution_dict(self, distribution_dict, weight_first):
        means, sds = [], []
        for k in range(self.components.K):
            mu, sigma = self.components.map(k)
            means.append(mu)
            sds.append(sigma)
        if weight_first:
            weights = self.gibbs_weight()
            idx = np.argsort(weights)
            sds = np.array(sds).flatten()
            means = np.array(means).flatten()
        else:
            means = np.array(means).flatten()
            idx = np.argsort(means)
            sds = np.array(sds).flatten()
            weights = self.gibbs_weight()
        means = self.label_switch(idx, means)
        sds = self.label_switch(idx, sds)
        weights = self.label_switch(idx, weights)
        self.old_mean, self.old_sigma = means, sds
        distribution_dict["mean"] = np.hstack((distribution_dict["mean"], means.reshape((-1, 1))))
        distribution_dict["variance"] = np.hstack((distribution_dict["variance"], sds.reshape((-1, 1))))
        distribution_dict["weights"] = np.hstack((distribution_dict["weights"], weights.reshape((-1, 1))))
        return distribution_dict

    def log_marg(self):
        """log p(X, z) (igmm.py:199-215): CRP term on the host, sum of log_marg_k from the device."""
        counts = self.components.counts[:self.components.K]
        facts_ = gammaln(counts)
        facts_[counts == 0] = 0
        log_prob_z = ((self.components.K - 1) * math.log(self.alpha) + gammaln(self.alpha)
                      - gammaln(np.sum(counts) + self.alpha) + np.sum(facts_))
        return log_prob_z + self.components.log_marg()

    def gibbs_weight(self):
        """igmm.py:219-227 (consumes np.random through scipy.stats, like the reference)."""
        Nk = self.components.counts[:self.components.K].tolist()
        alpha = [Nk[cid] + self.alpha / self.components.K for cid in range(self.components.K)]
        return stats.dirichlet(alpha).rvs(size=1).flatten()

    # ---- one sweep on the device --------------------------------------------------------------------------
    def _device_sweep(self, power=1.0, order=None, rng="reference"):
        """One pass over the data.  rng="reference": the uniforms are the next N values of the interpreter's
        `random` stream (what utils.draw would have consumed, utils.py:15); rng="philox": device counter-based."""
        comps = self.components
        u = _lib.mt19937_random(comps.N) if rng == "reference" else None
        try:
            st = comps.chain.sweep(self.alpha, power, order, u)
        except _lib.BgmmError as e:
            comps._dirty()
            if e.code == _lib.BGMM_EKMAX:
                raise IndexError(str(e))
            raise
        comps._dirty()
        self.last_sweep_stats = st
        return st
