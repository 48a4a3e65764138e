"""Full-covariance (normal-inverse-Wishart) components whose state lives on the GPU.

Mirror of the class surface of pybgmm/gaussian/gaussian_components.py:75-344 (`GaussianComponents`): same
constructor, attribute protocol (`X, prior, N, D, K, K_max, assignments, counts, m_N_numerators, S_N_partials,
logdet_covars, inv_covars, cached_log_prior`) and methods.  The arrays are views that are synchronised lazily
from device memory (one bgmm_get_state call after each mutation); every computation is done by libbgmm_b200.so.

Differences from the reference, all deliberate:
  * `K_max=None` means min(N, 1024) rather than N (the reference allocates three N x D x D arrays,
    gaussian_components.py:81-90); overflowing it raises IndexError (the reference's add_item has no check).
  * `v_0` must be integer valued (it indexes the lgamma table, :238; float worked only on old NumPy).
  * the outer-product cache (:116-118, O(N D^2) memory) is never materialised.
"""
import numpy as np

from .. import _lib

DEFAULT_K_MAX = 1024


class _LabelView(np.ndarray):
    """`components.assignments` as the reference exposes it: an int array the caller may write a label into.  The
    reference's restore path ends with `components.assignments[i] = k_old` on the host (crpmm.py:84-85); here that
    assignment is written through to the label on the device (bgmm_set_label), so the protocol
    cache_component_stats / del_item / restore_component_from_stats / assignments[i] = k leaves the device state
    consistent.  Only the object handed out by the property writes through; copies and slices are plain arrays."""

    def __new__(cls, labels, owner):
        obj = np.asarray(labels).view(cls)
        obj._owner = owner
        return obj

    def __array_finalize__(self, obj):
        self._owner = None

    def __setitem__(self, index, value):
        np.ndarray.__setitem__(self, index, value)
        owner = self._owner
        if owner is not None:
            touched = np.arange(self.shape[0])[index]
            for i in np.atleast_1d(touched):
                owner._chain.set_label(int(i), int(np.ndarray.__getitem__(self, int(i))))
            owner._cache = None
            owner._counts = None


class _DeviceComponents(object):
    _COV = "full"

    def __init__(self, X, prior, assignments=None, K_max=None, device=0):
        X = np.ascontiguousarray(X, dtype=np.float64)
        self.X = X
        self.prior = prior
        self.N, self.D = X.shape
        if K_max is None:
            K_max = min(self.N, DEFAULT_K_MAX)
            if self.N > DEFAULT_K_MAX:
                import warnings
                warnings.warn("K_max defaults to %d here (the reference's default is N = %d, which allocates three "
                              "N x D x D arrays); a sweep that needs more components raises IndexError -- pass K_max "
                              "explicitly" % (DEFAULT_K_MAX, self.N), stacklevel=3)
        self.K_max = int(K_max)
        self._check_prior()
        self._chain = self._make_chain(X, device)
        self._cache = None
        self._labels = None
        self._counts = None
        self._true_ref = None
        self._log_prior = None
        if assignments is None:
            z = -1 * np.ones(self.N, np.int64)
        else:
            z = np.asarray(assignments, np.int64)
            assert (self.N,) == z.shape
            # Apart from unassigned (-1), components should be labelled from 0 (gaussian_components.py:103-105)
            assert set(z.tolist()).difference([-1]) == set(range(int(z.max()) + 1))
        try:
            self._chain.set_assignments(z)
        except _lib.BgmmError as e:
            if e.code == _lib.BGMM_EKMAX:
                raise IndexError(str(e))
            raise

    def _check_prior(self):
        assert np.asarray(self.prior.S_0).shape == (self.D, self.D)

    def _make_chain(self, X, device):
        prior = self.prior
        return _lib.Chain(X, prior.m_0, prior.k_0, prior.v_0, prior.S_0, self.K_max, covariance_type=self._COV,
                          device=device)

    # ---- lazily synchronised views ------------------------------------------------------------
    def _state(self):
        """Counts and per-component statistics (K_max-sized arrays; the N labels are fetched separately)."""
        if self._cache is None:
            self._cache = self._chain.get_state(z=False)
        return self._cache

    def _dirty(self):
        self._cache = None
        self._labels = None
        self._counts = None

    @property
    def chain(self):
        """The underlying `_lib.Chain` (C-ABI handle)."""
        return self._chain

    @property
    def K(self):
        return self._chain.K

    @property
    def assignments(self):
        if self._labels is None:
            self._labels = _LabelView(self._chain.assignments(), self)
        return self._labels

    # ---- per-sweep record on the device (no label traffic) ---------------------------------------
    def contingency(self, true_assignments):
        """(T, K + 1) table of (rank of the true label, component); last column = unassigned.  The true labels are
        uploaded the first time a given vector is seen."""
        if self._true_ref is not true_assignments:
            self._chain.set_true_labels(true_assignments)
            self._true_ref = true_assignments
        return self._chain.contingency()

    def cluster_ssq(self):
        return self._chain.cluster_ssq()

    @property
    def counts(self):
        """Cluster sizes (K_max int64, zero beyond K) -- fetched on their own: the per-sweep record and the adaptive
        power schedule read them after every sweep and need nothing else."""
        if self._cache is not None:
            return self._cache["counts"]
        if self._counts is None:
            self._counts = self._chain.get_state(z=False, m_num=False, S_part=False, logdet=False,
                                                 inv_covar=False)["counts"]
        return self._counts

    @property
    def m_N_numerators(self):
        return self._state()["m_num"]

    @property
    def S_N_partials(self):
        return self._state()["S_part"]

    @property
    def logdet_covars(self):
        return self._state()["logdet"]

    @property
    def inv_covars(self):
        return self._state()["inv_covar"]

    @property
    def cached_log_prior(self):
        if self._log_prior is None:
            self._log_prior = self._chain.log_prior()
        return self._log_prior

    # ---- protocol ------------------------------------------------------------------------------
    def cache_component_stats(self, k):
        """Copies of the statistics of component k (gaussian_components.py:129-142)."""
        s = self._state()
        return (s["m_num"][k].copy(), s["S_part"][k].copy(), s["logdet"][k], s["inv_covar"][k].copy(),
                int(s["counts"][k]))

    def restore_component_from_stats(self, k, m_N_numerator, S_N_partial, logdet_covar, inv_covar, count):
        """gaussian_components.py:144-152; logdet / inverse are re-derived from the restored statistics."""
        _lib._check(_lib.lib().bgmm_set_component_stats(
            self._chain._h, int(k), _lib._dp(np.ascontiguousarray(m_N_numerator, dtype=np.float64)),
            _lib._dp(np.ascontiguousarray(S_N_partial, dtype=np.float64)), int(count)))
        self._dirty()

    def add_item(self, i, k):
        """Add X[i] to component k; k == K opens a new component (gaussian_components.py:154-169)."""
        try:
            self._chain.add_item(i, k)
        except _lib.BgmmError as e:
            if e.code in (_lib.BGMM_EKMAX, _lib.BGMM_EINVAL):
                raise IndexError(str(e))
            raise
        self._dirty()

    def del_item(self, i):
        """Remove X[i] from its component; an emptied component is deleted (gaussian_components.py:171-205)."""
        self._chain.del_item(i)
        self._dirty()

    def log_prior(self, i):
        return float(self.cached_log_prior[i])

    def log_post_pred_k(self, i, k):
        return float(self._chain.log_post_pred([i])[0, k])

    def log_post_pred(self, i):
        return self._chain.log_post_pred([i])[0]

    def log_post_pred_many(self, idx):
        """(len(idx), K) matrix of log_post_pred -- the batched form the engine is built around."""
        return self._chain.log_post_pred(idx)

    def log_marg_k(self, k):
        return float(self._chain.log_marg_k()[k])

    def log_marg(self):
        """Sum over components in slot order (gaussian_components.py:278-289)."""
        log_prob_X_given_z = 0.
        for v in self._chain.log_marg_k():
            log_prob_X_given_z += v
        return log_prob_X_given_z


class GaussianComponents(_DeviceComponents):
    _COV = "full"

    def map(self, k):
        """MAP estimate of mean and covariance of component k (gaussian_components.py:305-316)."""
        k_N = self.prior.k_0 + self.counts[k]
        v_N = self.prior.v_0 + self.counts[k]
        m_N = self.m_N_numerators[k] / k_N
        sigma = (self.S_N_partials[k] - k_N * np.outer(m_N, m_N)) / (v_N + self.D + 2)
        return (m_N, sigma)

    def rand_k(self, k):
        """A draw (mu, Sigma) from the NIW posterior of component k (gaussian_components.py:291-303).
        Uses scipy.stats.invwishart, so it is distributionally -- not stream -- equivalent to the reference."""
        from scipy.stats import invwishart
        k_N = self.prior.k_0 + self.counts[k]
        v_N = self.prior.v_0 + self.counts[k]
        m_N = self.m_N_numerators[k] / k_N
        S_N = self.S_N_partials[k] - k_N * np.outer(m_N, m_N)
        sigma = np.atleast_2d(invwishart.rvs(df=v_N, scale=S_N))
        mu = np.random.multivariate_normal(m_N, sigma / k_N)
        return mu, sigma
