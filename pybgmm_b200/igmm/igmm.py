"""IGMM base: initial assignments, component-class dispatch, log_marg (mirror of pybgmm/igmm/igmm.py:38-227)."""
import logging
import math
import time

import numpy as np
from scipy import stats
from scipy.special import gammaln

from .. import _lib
from ..gaussian import GaussianComponents, GaussianComponentsDiag, GaussianComponentsFixedVar
from ..gmm import GMM

logger = logging.getLogger(__name__)


class IGMM(GMM):
    """Infinite Gaussian mixture model whose components live on the GPU.  Constructor arguments as
    `pybgmm/igmm/igmm.py:68-71`; `device` (extra, keyword) selects the CUDA device."""

    def __init__(self, X, kernel_prior, alpha, save_path, assignments="rand", K=1, K_max=None,
                 covariance_type="full", device=0):
        super(IGMM, self).__init__()
        if np.ndim(X) < 2:
            raise ValueError('X must be at least a 2-dimensional array.')
        self.N, self.D = X.shape
        self.alpha = alpha
        self.save_path = save_path
        z0 = self._initial_assignments(assignments, K)
        # igmm.py:104-111: "full" -> NIW components, "diag" -> product of NIX, "fixed" -> known variance
        component_classes = {"full": GaussianComponents, "diag": GaussianComponentsDiag,
                             "fixed": GaussianComponentsFixedVar}
        assert covariance_type in component_classes, "Invalid covariance type."
        self.components = component_classes[covariance_type](X, kernel_prior, z0, K_max, device=device)
        self.last_sweep_stats = None

    def _initial_assignments(self, mode, K):
        """The four initialisations of igmm.py:86-102.  "rand" takes N draws from the global NumPy stream, like the
        reference, and then closes gaps in the labels; closing gaps one missing label at a time (what the reference's
        loop does) maps every label to its rank among the labels present, which is what np.unique returns."""
        if isinstance(mode, str):
            if mode == "rand":
                drawn = np.random.randint(0, K, self.N)
                return np.unique(drawn, return_inverse=True)[1].astype(drawn.dtype)
            if mode == "one-by-one":
                z = np.full(self.N, -1, dtype=int)
                z[0] = 0
                return z
            if mode == "each-in-own":
                return np.arange(self.N)
        return mode  # a caller-supplied label vector

    # ---- distribution dict (the role of igmm.py:115-197; the reference's plotting is not reproduced) ------------
    def setup_distribution_dict(self, num_saved):
        return dict((name, np.zeros(shape=(num_saved, 0))) for name in ("mean", "variance", "weights"))

    def update_distribution_dict(self, distribution_dict, weight_first):
        """Append one column (MAP mean, MAP scale, a posterior draw of the weights) ordered by weight or by mean.
        The draw consumes np.random (scipy's Dirichlet sampler), like the reference's gibbs_weight call."""
        maps = [self.components.map(k) for k in range(self.components.K)]
        means = np.array([m for m, _ in maps]).flatten()
        scales = np.array([c for _, c in maps]).flatten()
        weights = self.gibbs_weight()
        order = np.argsort(weights if weight_first else means)
        columns = {"mean": self.label_switch(order, means), "variance": self.label_switch(order, scales),
                   "weights": self.label_switch(order, weights)}
        self.old_mean, self.old_sigma = columns["mean"], columns["variance"]
        for name, col in columns.items():
            distribution_dict[name] = np.hstack((distribution_dict[name], col.reshape((-1, 1))))
        return distribution_dict

    def log_marg(self):
        """log p(X, z) (igmm.py:199-215): CRP partition term here, the sum over components of log_marg_k on the
        device.  Counts of zero contribute nothing to the partition term."""
        K = self.components.K
        n_k = self.components.counts[:K]
        occupied = gammaln(n_k)
        occupied[n_k == 0] = 0
        partition = ((K - 1) * math.log(self.alpha) + gammaln(self.alpha) - gammaln(np.sum(n_k) + self.alpha)
                     + np.sum(occupied))
        return partition + self.components.log_marg()

    def gibbs_weight(self):
        """One posterior draw of the mixture weights given the counts (igmm.py:219-227): Dirichlet(n_k + alpha / K)."""
        K = self.components.K
        concentration = [n + self.alpha / K for n in self.components.counts[:K].tolist()]
        return stats.dirichlet(concentration).rvs(size=1).flatten()

    # ---- the sweep loop shared by CRPMM / PCRPMM ---------------------------------------------------------------
    def _run_sweeps(self, n_iter, true_assignments, num_saved, weight_first, schedule, rng):
        """`schedule(i_iter, N) -> (order or None, power)` is the sampler-specific part.  Record keeping follows the
        reference: the clock of `sample_time` covers the sweep only (restarted after the bookkeeping, crpmm.py:43,92),
        the distribution dict is refreshed before a sweep when K equals `num_saved` from the third sweep on
        (crpmm.py:49)."""
        records = self.setup_record_dict()
        saved = self.setup_distribution_dict(num_saved)
        for i_iter in range(n_iter):
            if i_iter > 1 and self.components.K == num_saved:
                saved = self.update_distribution_dict(saved, weight_first)
            order, power = schedule(i_iter, self.components.N)
            tic = time.time()
            self._device_sweep(power=power, order=order, rng=rng)
            records = self.update_record_dict(records, i_iter, true_assignments, tic)
        return records, saved

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
