"""TEST INFRASTRUCTURE ONLY -- CPU restatement (NumPy) of the reference's fixed-variance components and of the CRP /
powered-CRP sweep over them, for SURVEY.md 8(f2).  Nothing in the product may import this file; only tests/ may.

Follows pybgmm/gaussian/gaussian_components_fixedvar.py: state :84-92, add_item :146-162, del_item :164-180,
del_component :182-202, log_prior :204-210, log_post_pred :221-232, log_marg_k :234-256,
_update_log_prod_precision_pred_and_precision_pred :278-286; the sweep is pybgmm/igmm/crpmm.py:57-88 (pcrpmm.py:93-131
with a scan order and log(n ** r)) and the partition term of pybgmm/igmm/igmm.py:199-215.
The uniforms and the scan order are explicit inputs (one uniform per datum, utils.py:15).

Pinned: tests/golden/golden_fixedvar.json holds outputs of the reference itself (tests/golden/make_golden.py) and
tests/test_oracle_fixedvar.py replays them.  The CUDA path for this variant is not built yet.
"""
import math

import numpy as np
from scipy.special import gammaln, logsumexp


class FixedVarState(object):
    """Per-component statistics of a mixture whose components share a known diagonal variance `var`; the component
    means have independent normal priors N(mu_0, var_0)."""

    def __init__(self, X, var, mu_0, var_0, K_max=None):
        self.X = np.asarray(X, dtype=np.float64)
        self.N, self.D = self.X.shape
        self.tau = 1.0 / np.asarray(var, dtype=np.float64)       # data precision (:80)
        self.mu_0 = np.asarray(mu_0, dtype=np.float64)
        self.tau_0 = 1.0 / np.asarray(var_0, dtype=np.float64)   # prior precision of the mean (:82)
        self.K_max = self.N if K_max is None else int(K_max)
        self.z = -np.ones(self.N, dtype=np.int64)
        self.n = np.zeros(self.K_max, dtype=np.int64)
        self.num = np.zeros((self.K_max, self.D))       # precision-weighted sums (numerator of the posterior mean)
        self.tau_N = np.zeros((self.K_max, self.D))     # posterior precision of the mean
        self.lpp = np.zeros(self.K_max)                 # sum of log predictive precisions
        self.tau_pred = np.zeros((self.K_max, self.D))  # predictive precisions
        self.K = 0
        self.c_norm = -0.5 * self.D * math.log(2. * np.pi)
        self.log_prior_all = np.array([self._log_normal(i, self.mu_0, np.log(self.tau_0).sum(), self.tau_0)
                                       for i in range(self.N)])

    # ---- building blocks -------------------------------------------------------------------------------------
    def _log_normal(self, i, mean, sum_log_prec, prec):
        d = self.X[i, :] - mean
        return self.c_norm + 0.5 * sum_log_prec - 0.5 * (np.square(d) * prec).sum()

    def _refresh_predictive(self, k):
        pred = self.tau_N[k] * self.tau / (self.tau_N[k] + self.tau)
        self.lpp[k] = np.log(pred).sum()
        self.tau_pred[k, :] = pred

    def set_assignments(self, z):
        z = np.asarray(z, dtype=np.int64)
        for k in range(int(z.max()) + 1):
            for i in np.where(z == k)[0]:
                self.add(i, k)

    def add(self, i, k):
        if k == self.K:
            if k >= self.K_max:
                raise IndexError("K_max exceeded")
            self.K += 1
            self.num[k, :] = self.tau_0 * self.mu_0
            self.tau_N[k, :] = self.tau_0
        self.num[k, :] += self.tau * self.X[i]
        self.tau_N[k, :] += self.tau
        self.n[k] += 1
        self._refresh_predictive(k)
        self.z[i] = k

    def remove(self, i):
        k = self.z[i]
        if k == -1:
            return
        self.n[k] -= 1
        self.z[i] = -1
        if self.n[k] == 0:
            last = self.K - 1
            self.K = last
            if k != last:
                for arr in (self.num, self.tau_N, self.tau_pred):
                    arr[k] = arr[last]
                self.lpp[k] = self.lpp[last]
                self.n[k] = self.n[last]
                self.z[self.z == last] = k
            for arr in (self.num, self.tau_N, self.tau_pred):
                arr[last].fill(0.)
            self.lpp[last] = 0.
            self.n[last] = 0
        else:
            self.num[k, :] -= self.tau * self.X[i]
            self.tau_N[k, :] -= self.tau
            self._refresh_predictive(k)

    def snapshot(self, k):
        return (self.num[k].copy(), self.tau_N[k].copy(), self.lpp[k], self.tau_pred[k].copy(), self.n[k])

    def restore(self, k, snap):
        self.num[k, :], self.tau_N[k, :], self.lpp[k], self.tau_pred[k, :], self.n[k] = snap

    # ---- likelihood terms ------------------------------------------------------------------------------------
    def log_post_pred(self, i):
        K = self.K
        means = self.num[:K] / self.tau_N[:K]
        d = means - self.X[i]
        return self.c_norm + 0.5 * self.lpp[:K] - 0.5 * (np.square(d) * self.tau_pred[:K]).sum(axis=1)

    def log_post_pred_k(self, i, k):
        return self._log_normal(i, self.num[k] / self.tau_N[k], self.lpp[k], self.tau_pred[k])

    def log_marg_k(self, k):
        Xk = self.X[np.where(self.z == k)]
        n = self.n[k]
        s = n / self.tau_0 + 1. / self.tau
        return np.sum(
            (n - 1) / 2. * np.log(self.tau) - 0.5 * n * math.log(2 * np.pi) - 0.5 * np.log(s)
            - 0.5 * self.tau * np.square(Xk).sum(axis=0) - 0.5 * self.tau_0 * np.square(self.mu_0)
            + 0.5 * (np.square(Xk.sum(axis=0)) * self.tau / self.tau_0 + np.square(self.mu_0) * self.tau_0 / self.tau
                     + 2 * Xk.sum(axis=0) * self.mu_0) / s)

    def log_marg(self, alpha):
        """igmm.py:199-215: CRP partition term + sum over components."""
        n_k = self.n[:self.K]
        facts = gammaln(n_k)
        facts[n_k == 0] = 0
        lpz = (self.K - 1) * math.log(alpha) + gammaln(alpha) - gammaln(np.sum(n_k) + alpha) + np.sum(facts)
        lpx = 0.
        for k in range(self.K):
            lpx += self.log_marg_k(k)
        return lpz + lpx

    # ---- one Gibbs sweep (crpmm.py:57-88 / pcrpmm.py:93-131) ---------------------------------------------------
    def sweep(self, uniforms, alpha, order=None, power=1.0):
        scan = range(self.N) if order is None else order
        for j, i in enumerate(scan):
            k_old = self.z[i]
            K_before = self.K
            snap = self.snapshot(k_old)
            self.remove(i)
            w = np.zeros(self.K + 1)
            if power != 1.0:
                w[:self.K] = np.log(np.power(self.n[:self.K], power))
            else:
                w[:self.K] = np.log(self.n[:self.K])
            w[:self.K] += self.log_post_pred(i)
            w[-1] = math.log(alpha) + self.log_prior_all[i]
            p = np.exp(w - logsumexp(w))
            u = uniforms[j]
            k = len(p) - 1
            for t in range(len(p)):          # utils.draw (utils.py:15-20)
                u = u - p[t]
                if u < 0:
                    k = t
                    break
            if k == k_old and self.K == K_before:
                self.restore(k_old, snap)
                self.z[i] = k_old
            else:
                self.add(i, k)
