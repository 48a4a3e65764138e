"""ctypes front-end of the CPU oracle (oracle/bgmm_oracle.c).

TEST INFRASTRUCTURE ONLY -- the checker, never the product.  Only tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.

`Oracle` mirrors the slice of the reference's `GaussianComponents{,Diag}` +
`CRPMM/PCRPMM.collapsed_gibbs_sampler` protocol that the hot path uses, with the
uniform stream and the scan order as *explicit inputs* (the reference takes them
from the global `random` / `np.random` state: pybgmm/utils/utils.py:15,
pybgmm/igmm/pcrpmm.py:89).
"""
import ctypes as C
import math
import os
import random
import subprocess

import numpy as np
from scipy.special import gammaln

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class SweepStats(C.Structure):
    _fields_ = [("moves", C.c_int64), ("births", C.c_int64), ("deaths", C.c_int64),
                ("evals", C.c_int64), ("K_end", C.c_int64), ("min_margin", C.c_double)]


def build(force=False):
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "bgmm_oracle.c")
    if force or not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(so) < os.path.getmtime(src)):
        subprocess.check_call(["make", "-C", HERE, "-s", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int64)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [dp, C.c_int64, C.c_int, C.c_int, dp, C.c_double, C.c_int64, dp, C.c_int,
                                 dp, dp, C.c_int64]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_assignments.argtypes = [C.c_void_p, ip]
        L.orc_add_item.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        L.orc_del_item.argtypes = [C.c_void_p, C.c_int64]
        L.orc_log_prior.restype = C.c_double
        L.orc_log_prior.argtypes = [C.c_void_p, C.c_int64]
        L.orc_log_post_pred_k.restype = C.c_double
        L.orc_log_post_pred_k.argtypes = [C.c_void_p, C.c_int64, C.c_int]
        L.orc_log_post_pred.argtypes = [C.c_void_p, C.c_int64, dp]
        L.orc_log_marg_k.restype = C.c_double
        L.orc_log_marg_k.argtypes = [C.c_void_p, C.c_int]
        L.orc_log_marg.restype = C.c_double
        L.orc_log_marg.argtypes = [C.c_void_p, C.c_double]
        L.orc_sweep.argtypes = [C.c_void_p, ip, dp, C.c_double, dp, C.POINTER(SweepStats)]
        L.orc_sweep_constrained.argtypes = [C.c_void_p, ip, dp, C.c_int64, C.c_double, dp, C.POINTER(C.c_int), C.c_int,
                                            C.POINTER(C.c_int64), C.POINTER(SweepStats)]
        L.orc_peek_probs.argtypes = [C.c_void_p, C.c_int64, C.c_double, dp, dp, dp]
        L.orc_K.argtypes = [C.c_void_p]
        for name, rt in (("orc_z", ip), ("orc_counts", ip), ("orc_num", dp), ("orc_Sp", dp),
                         ("orc_logdet", dp), ("orc_inv", dp), ("orc_cached_log_prior", dp)):
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def make_tables(v_0, N):
    """lgamma(n/2) and log(n) tables exactly as gaussian_components.py:120-122."""
    n = np.concatenate([[1], np.arange(1, int(v_0) + N + 2)])
    return gammaln(n / 2.), np.log(n)


def logcount_table(N, power):
    """log(count**r) table as pcrpmm.py:107-108 evaluates it (np.log(np.power(counts, n_power)))."""
    n = np.arange(0, N + 1)
    with np.errstate(divide="ignore"):
        return np.log(np.power(n, power))


class Oracle(object):
    """CPU oracle for one chain.  cov_type: "full" (NIW) or "diag" (NIX product)."""

    def __init__(self, X, m_0, k_0, v_0, S_0, K_max=None, covariance_type="full"):
        X = np.ascontiguousarray(X, dtype=np.float64)
        assert X.ndim == 2
        self.X = X
        self.N, self.D = X.shape
        assert float(v_0) == int(v_0), "v_0 must be integer valued (table index, gaussian_components.py:238)"
        self.v_0 = int(v_0)
        self.k_0 = float(k_0)
        self.m_0 = np.ascontiguousarray(m_0, dtype=np.float64)
        self.cov = 0 if covariance_type == "full" else 1
        self.S_0 = np.ascontiguousarray(S_0, dtype=np.float64)
        assert self.S_0.shape == ((self.D, self.D) if self.cov == 0 else (self.D,))
        assert self.v_0 >= self.D  # niw.py:21
        self.K_max = int(K_max) if K_max is not None else self.N
        lg, lv = make_tables(self.v_0, self.N)
        self._lg, self._lv = np.ascontiguousarray(lg), np.ascontiguousarray(lv)
        self._h = lib().orc_create(_dp(self.X), self.N, self.D, self.cov, _dp(self.m_0), self.k_0, self.v_0,
                                   _dp(self.S_0), self.K_max, _dp(self._lg), _dp(self._lv), len(self._lg))
        self._ss = self.D * self.D if self.cov == 0 else self.D

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    # -- state views (copies) ------------------------------------------------
    @property
    def K(self):
        return lib().orc_K(self._h)

    def _view(self, fn, n, dtype):
        ptr = getattr(lib(), fn)(self._h)
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)

    @property
    def assignments(self):
        return self._view("orc_z", self.N, np.int64)

    @property
    def counts(self):
        return self._view("orc_counts", self.K_max, np.int64)

    @property
    def m_N_numerators(self):
        return self._view("orc_num", self.K_max * self.D, np.float64).reshape(self.K_max, self.D)

    @property
    def S_N_partials(self):
        a = self._view("orc_Sp", self.K_max * self._ss, np.float64)
        return a.reshape((self.K_max, self.D, self.D) if self.cov == 0 else (self.K_max, self.D))

    @property
    def logdet_covars(self):
        return self._view("orc_logdet", self.K_max, np.float64)

    @property
    def inv_covars(self):
        a = self._view("orc_inv", self.K_max * self._ss, np.float64)
        return a.reshape((self.K_max, self.D, self.D) if self.cov == 0 else (self.K_max, self.D))

    @property
    def cached_log_prior(self):
        return self._view("orc_cached_log_prior", self.N, np.float64)

    # -- protocol --------------------------------------------------------------
    def set_assignments(self, z):
        z = np.ascontiguousarray(z, dtype=np.int64)
        rc = lib().orc_set_assignments(self._h, _ip(z))
        if rc == -1:
            raise AssertionError("assignments must be labelled 0..max (gaussian_components.py:103-105)")
        if rc == -2:
            raise IndexError("K_max overflow")

    def add_item(self, i, k):
        lib().orc_add_item(self._h, int(i), int(k))

    def del_item(self, i):
        lib().orc_del_item(self._h, int(i))

    def log_prior(self, i):
        return lib().orc_log_prior(self._h, int(i))

    def log_post_pred_k(self, i, k):
        return lib().orc_log_post_pred_k(self._h, int(i), int(k))

    def log_post_pred(self, i):
        out = np.zeros(self.K, np.float64)
        lib().orc_log_post_pred(self._h, int(i), _dp(out))
        return out

    def log_marg_k(self, k):
        return lib().orc_log_marg_k(self._h, int(k))

    def log_marg(self, alpha):
        return lib().orc_log_marg(self._h, float(alpha))

    def peek_probs(self, i, alpha, logcount_tab=None):
        """(prob_z[0..K'], lpp[0..K'-1]) for datum i under the current state, state unchanged."""
        p = np.zeros(self.K_max + 1, np.float64)
        l = np.zeros(self.K_max + 1, np.float64)
        t = None if logcount_tab is None else np.ascontiguousarray(logcount_tab, dtype=np.float64)
        K = lib().orc_peek_probs(self._h, int(i), float(alpha), None if t is None else _dp(t), _dp(p), _dp(l))
        return p[:K + 1].copy(), l[:K].copy()

    def sweep(self, uniforms, alpha, order=None, logcount_tab=None):
        """One Gibbs sweep (crpmm.py:57-88 / pcrpmm.py:93-131).  Returns a SweepStats."""
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        assert u.shape == (self.N,)
        o = None if order is None else np.ascontiguousarray(order, dtype=np.int64)
        t = None if logcount_tab is None else np.ascontiguousarray(logcount_tab, dtype=np.float64)
        st = SweepStats()
        rc = lib().orc_sweep(self._h, None if o is None else _ip(o), _dp(u), float(alpha),
                             None if t is None else _dp(t), C.byref(st))
        if rc != 0:
            raise IndexError("K_max overflow (the reference raises IndexError in add_item)")
        return st


    def sweep_constrained(self, uniforms, alpha, status, order=None, logcount_tab=None):
        """One sweep with CSCRPMM's constrained re-draw (cscrpmm.py:342-350, :455-461).  status[k]: 1 useful, 2 non-useful
        slot at the start of the sweep.  `uniforms` is the random.random() stream from the sweep's first draw on; returns
        (SweepStats, number of uniforms consumed)."""
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        o = None if order is None else np.ascontiguousarray(order, dtype=np.int64)
        t = None if logcount_tab is None else np.ascontiguousarray(logcount_tab, dtype=np.float64)
        sv = np.ascontiguousarray(status, dtype=np.int32)
        st = SweepStats()
        used = C.c_int64()
        rc = lib().orc_sweep_constrained(self._h, None if o is None else _ip(o), _dp(u), len(u), float(alpha),
                                         None if t is None else _dp(t), sv.ctypes.data_as(C.POINTER(C.c_int)), len(sv),
                                         C.byref(used), C.byref(st))
        if rc == -2:
            raise RuntimeError("the uniform stream ran out")
        if rc != 0:
            raise IndexError("K_max overflow (the reference raises IndexError in add_item)")
        return st, used.value


# ---------------------------------------------------------------------------
# Reference-shaped drivers: consume the *global* random / np.random state the
# way the reference does, so a seeded run is comparable with the shimmed
# reference end to end.
# ---------------------------------------------------------------------------
def init_assignments(N, assignments="rand", K=1):
    """igmm.py:86-99."""
    if isinstance(assignments, str) and assignments == "rand":
        z = np.random.randint(0, K, N)
        for k in range(z.max()):
            while len(np.nonzero(z == k)[0]) == 0:
                z[np.where(z > k)] -= 1
            if z.max() == k:
                break
        return z.astype(np.int64)
    if isinstance(assignments, str) and assignments == "one-by-one":
        z = -1 * np.ones(N, dtype=np.int64)
        z[0] = 0
        return z
    if isinstance(assignments, str) and assignments == "each-in-own":
        return np.arange(N, dtype=np.int64)
    return np.asarray(assignments, dtype=np.int64)


def run_crpmm(orc, n_iter, alpha):
    """crpmm.py:47-88 with num_saved=0: one random.random() per datum, natural order."""
    stats = []
    for _ in range(n_iter):
        u = np.array([random.random() for _ in range(orc.N)])
        stats.append(orc.sweep(u, alpha))
    return stats


def run_pcrpmm(orc, n_iter, alpha, n_power=1.01, power_burnin=0, flag_power=True):
    """pcrpmm.py:62-131 with num_saved=0."""
    stats = []
    tab = logcount_table(orc.N, n_power)
    for i_iter in range(n_iter):
        if flag_power and n_power > 1:
            order = np.random.permutation(range(orc.N))
        else:
            order = None
        u = np.array([random.random() for _ in range(orc.N)])
        use_power = flag_power and i_iter > power_burnin
        stats.append(orc.sweep(u, alpha, order=order, logcount_tab=tab if use_power else None))
    return stats


def run_adapcrpmm(orc, n_iter, alpha, r_up=1.3, adapcrp_perct=0.04, adapcrp_burnin=0, flag_adapcrp=True):
    """adapcrpmm.py:83-157 with num_saved=0: the power is recomputed before every sweep from the cluster sizes
    (:100-104), random scan whenever it exceeds 1 (:110-115).  Burn-in sweeps (where the reference itself stops with
    UnboundLocalError, :110) run as plain CRP sweeps, as its docstring describes."""
    stats = []
    for i_iter in range(n_iter):
        power = 1.0
        if flag_adapcrp and i_iter > adapcrp_burnin:
            n_k = np.array(orc.counts[:orc.K])
            small = len(n_k[np.where(n_k <= orc.N * adapcrp_perct)[0]]) * 1.0 / len(n_k)
            power = 1.0 + (r_up - 1.0) * small
        order = np.random.permutation(range(orc.N)) if (flag_adapcrp and power > 1) else None
        u = np.array([random.random() for _ in range(orc.N)])
        use_power = flag_adapcrp and i_iter > adapcrp_burnin
        stats.append(orc.sweep(u, alpha, order=order, logcount_tab=logcount_table(orc.N, power) if use_power else None))
    return stats


def _status_from_counts(counts, K, threshold, K_max):
    """cscrpmm.py:159-167 / :425-432: slots with more than `threshold` members are useful (1), the others non-useful
    (2); slots that do not exist yet are neither (0)."""
    st = np.zeros(K_max + 1, dtype=np.int32)
    st[:K] = np.where(np.asarray(counts[:K]) > threshold, 1, 2)
    return st


def _constrained_sweep(orc, alpha, status, order, tab):
    """One constrained sweep consuming the global `random` stream exactly as the reference would: one random.random()
    per datum plus one per re-draw (utils.py:15)."""
    state = random.getstate()
    L = 400 * orc.N + 4096      # a datum far from every useful cluster re-draws many times (1 / P(useful) on average)
    u = np.array([random.random() for _ in range(L)])
    st, used = orc.sweep_constrained(u, alpha, status, order=order, logcount_tab=tab)
    random.setstate(state)
    for _ in range(used):
        random.random()
    return st


def run_cscrpmm(orc, n_iter, alpha, flag_constrain=False, n_constrain=1000000, thres=0., flag_power=False, n_power=1,
                power_burnin=100000, flag_approx=False, approx_thres_perct=0., approx_burnin=1000000,
                flag_adapcrp_form2=False, r_up=1., adapcrp_perct=0., adapcrp_burnin=1000000):
    """cscrpmm.py:96-485 with num_saved=0, for the flags the CUDA path supports (constrained re-draw, powered scan,
    per-sweep adaptive power, approximate step)."""
    stats = []
    N = orc.N
    for i_iter in range(n_iter):
        constrained = False
        status = None
        if flag_constrain and i_iter % n_constrain == 0:
            constrained = True
            status = _status_from_counts(orc.counts, orc.K, N * thres, orc.K_max)
        power = 1.0
        if flag_adapcrp_form2 and i_iter > adapcrp_burnin:
            n_k = np.array(orc.counts[:orc.K])
            small = len(n_k[np.where(n_k <= N * adapcrp_perct)[0]]) * 1.0 / len(n_k)
            power_form2 = 1.0 + (r_up - 1.0) * small
        order = np.random.permutation(range(N)) if (flag_power and n_power > 1) else None
        if flag_power and i_iter > power_burnin:
            power = n_power
        elif flag_adapcrp_form2 and i_iter > adapcrp_burnin:
            power = power_form2
        tab = logcount_table(N, power) if power != 1.0 else None
        if constrained:
            stats.append(_constrained_sweep(orc, alpha, status, order, tab))
        else:
            u = np.array([random.random() for _ in range(N)])
            stats.append(orc.sweep(u, alpha, order=order, logcount_tab=tab))
        if flag_approx and i_iter > approx_burnin:
            status = _status_from_counts(orc.counts, orc.K, N * approx_thres_perct, orc.K_max)
            stats.append(_constrained_sweep(orc, alpha, status, None, None))
    return stats
