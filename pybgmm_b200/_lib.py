"""ctypes binding of libbgmm_b200.so (include/bgmm_b200.h) and `Chain`, a thin object over one handle.

There is no CPU fallback: if the shared library is missing, or there is no CUDA device, every entry point
raises.  The product never imports anything from oracle/.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BGMM_B200_LIB", os.path.join(HERE, "lib", "libbgmm_b200.so"))

BGMM_OK, BGMM_EINVAL, BGMM_ENODEV, BGMM_EKMAX, BGMM_ENUMERIC, BGMM_ENOMEM, BGMM_EWATCHDOG = 0, -1, -2, -3, -4, -5, -6
COV_FULL, COV_DIAG, COV_FIXED = 0, 1, 2

EXPORTS = (
    "bgmm_set_component_stats",
    "bgmm_version", "bgmm_last_error", "bgmm_device_count", "bgmm_create", "bgmm_destroy", "bgmm_set_stream",
    "bgmm_set_assignments", "bgmm_sweep", "bgmm_sweep_dev", "bgmm_set_engine", "bgmm_seed", "bgmm_get_uniforms",
    "bgmm_sweep_index", "bgmm_get_state", "bgmm_get_assignments_dev", "bgmm_K", "bgmm_log_prior",
    "bgmm_log_post_pred", "bgmm_log_marg_k", "bgmm_log_marg", "bgmm_add_item", "bgmm_del_item", "bgmm_mt19937_fill",
    "bgmm_set_true_labels", "bgmm_contingency", "bgmm_cluster_ssq", "bgmm_set_label", "bgmm_set_state", "bgmm_set_guard",
    "bgmm_fork", "bgmm_sweep_many", "bgmm_create_fixedvar", "bgmm_sweep_constrained",
)


class SweepStats(C.Structure):
    _fields_ = [("K", C.c_int64), ("moves", C.c_int64), ("births", C.c_int64), ("deaths", C.c_int64),
                ("evals", C.c_int64), ("windows", C.c_int64), ("seq_data", C.c_int64), ("wasted", C.c_int64),
                ("min_margin", C.c_double), ("device_ms", C.c_double), ("explicit_evals", C.c_int64),
                ("refreshes", C.c_int64), ("generic_from", C.c_int64), ("phase_cycles", C.c_int64 * 16),
                ("launches", C.c_int64), ("sweep_kernel_ms", C.c_double), ("guard_hits", C.c_int64),
                ("fast_steps", C.c_int64)]

    PHASES = ("stage", "head", "eval", "draw", "update", "scalars", "wineval", "barrier", "rare", "steps", "moves",
              "rounds")

    def as_dict(self):
        d = {n: getattr(self, n) for n, _ in self._fields_ if n != "phase_cycles"}
        return d

    def phases(self):
        return dict(zip(self.PHASES, list(self.phase_cycles)))


class BgmmError(RuntimeError):
    def __init__(self, code, msg):
        super(BgmmError, self).__init__("libbgmm_b200 error %d: %s" % (code, msg))
        self.code = code


_LIB = None


def lib():
    """Load the shared library (raises if it has not been built: run `python -m pybgmm_b200.build`)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError("libbgmm_b200.so not found at %s -- build it with `python -m pybgmm_b200.build` "
                          "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_void_p
    L.bgmm_version.restype = C.c_char_p
    L.bgmm_last_error.restype = C.c_char_p
    L.bgmm_device_count.restype = C.c_int
    L.bgmm_create.argtypes = [dp, C.c_int64, C.c_int32, C.c_int32, dp, C.c_double, C.c_int64, dp, C.c_int32, dp, dp,
                              C.c_int64, C.c_int32, C.POINTER(vp)]
    L.bgmm_create_fixedvar.argtypes = [dp, C.c_int64, C.c_int32, dp, dp, dp, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.bgmm_destroy.argtypes = [vp]
    L.bgmm_set_stream.argtypes = [vp, vp]
    L.bgmm_set_assignments.argtypes = [vp, ip]
    L.bgmm_sweep.argtypes = [vp, ip, dp, C.c_double, C.c_double, C.POINTER(SweepStats)]
    L.bgmm_sweep_dev.argtypes = [vp, vp, vp, C.c_double, C.c_double, C.POINTER(SweepStats)]
    L.bgmm_sweep_constrained.argtypes = [vp, ip, dp, C.c_int64, C.c_double, C.c_double, C.POINTER(C.c_int32), C.c_int32,
                                         C.POINTER(C.c_int64), C.POINTER(SweepStats)]
    L.bgmm_set_engine.argtypes = [vp, C.c_int32]
    L.bgmm_seed.argtypes = [vp, C.c_uint64]
    L.bgmm_get_uniforms.argtypes = [vp, C.c_int64, dp]
    L.bgmm_sweep_index.argtypes = [vp]
    L.bgmm_sweep_index.restype = C.c_int64
    L.bgmm_get_state.argtypes = [vp, ip, ip, C.POINTER(C.c_int32), dp, dp, dp, dp]
    L.bgmm_get_assignments_dev.argtypes = [vp, vp]
    L.bgmm_K.argtypes = [vp]
    L.bgmm_log_prior.argtypes = [vp, dp]
    L.bgmm_log_post_pred.argtypes = [vp, ip, C.c_int64, dp]
    L.bgmm_log_marg_k.argtypes = [vp, dp]
    L.bgmm_log_marg.argtypes = [vp, C.c_double, dp]
    L.bgmm_add_item.argtypes = [vp, C.c_int64, C.c_int32]
    L.bgmm_del_item.argtypes = [vp, C.c_int64]
    L.bgmm_set_component_stats.argtypes = [vp, C.c_int32, dp, dp, C.c_int64]
    L.bgmm_set_label.argtypes = [vp, C.c_int64, C.c_int32]
    L.bgmm_set_state.argtypes = [vp, ip, C.c_int32, dp, dp]
    L.bgmm_set_guard.argtypes = [vp, C.c_double]
    L.bgmm_fork.argtypes = [vp, C.POINTER(vp)]
    L.bgmm_sweep_many.argtypes = [C.POINTER(vp), C.c_int32, C.POINTER(vp), C.POINTER(vp), C.c_double, C.c_double,
                                  C.POINTER(SweepStats)]
    L.bgmm_set_true_labels.argtypes = [vp, ip, C.c_int32]
    L.bgmm_contingency.argtypes = [vp, ip]
    L.bgmm_cluster_ssq.argtypes = [vp, dp]
    L.bgmm_mt19937_fill.argtypes = [C.POINTER(C.c_uint32), dp, C.c_int64]
    _LIB = L
    return L


def _check(rc):
    if rc != 0:
        raise BgmmError(rc, lib().bgmm_last_error().decode("utf-8", "replace"))


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _ip(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_int64))


def device_count():
    return lib().bgmm_device_count()


def mt19937_random(n):
    """`n` values of random.random() taken from (and advancing) the interpreter's global `random` state
    (pybgmm/utils/utils.py:15 draws one per datum)."""
    import random
    st = random.getstate()
    words = np.array(st[1], dtype=np.uint32)
    out = np.empty(int(n), dtype=np.float64)
    _check(lib().bgmm_mt19937_fill(words.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(out), int(n)))
    random.setstate((st[0], tuple(int(w) for w in words), st[2]))
    return out


def make_tables(v_0, N):
    """lgamma(n/2) and log(n) tables exactly as gaussian_components.py:120-122 builds them (SciPy / NumPy)."""
    from scipy.special import gammaln
    n = np.concatenate([[1], np.arange(1, int(v_0) + int(N) + 2)])
    return np.ascontiguousarray(gammaln(n / 2.)), np.ascontiguousarray(np.log(n))


class Chain(object):
    """One Gibbs chain resident on one GPU (a handle of the C-ABI)."""

    def __init__(self, X, m_0, k_0, v_0, S_0, K_max, covariance_type="full", device=0, tables=True):
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError("X must be a 2-dimensional array.")
        self.N, self.D = X.shape
        if float(v_0) != int(v_0):
            raise ValueError("v_0 must be integer valued (it indexes the lgamma table, gaussian_components.py:238)")
        self.cov = {"full": COV_FULL, "diag": COV_DIAG}[covariance_type]
        self.K_max = int(K_max)
        m_0 = np.ascontiguousarray(m_0, dtype=np.float64).reshape(-1)
        S_0 = np.ascontiguousarray(S_0, dtype=np.float64)
        assert m_0.shape == (self.D,)
        assert S_0.shape == ((self.D, self.D) if self.cov == COV_FULL else (self.D,))
        lg = lv = None
        if tables:
            lg, lv = make_tables(v_0, self.N)
        h = C.c_void_p()
        _check(lib().bgmm_create(_dp(X), self.N, self.D, self.cov, _dp(m_0), float(k_0), int(v_0), _dp(S_0),
                                 self.K_max, _dp(lg), _dp(lv), 0 if lg is None else len(lg), int(device), C.byref(h)))
        self._h = h
        self._ss = self.D * self.D if self.cov == COV_FULL else self.D

    @classmethod
    def fixed_variance(cls, X, var, mu_0, var_0, K_max, device=0):
        """A chain over fixed-variance components (bgmm_create_fixedvar): known diagonal data variance `var`,
        independent normal priors N(mu_0, var_0) on the means; var, mu_0, var_0 are D-vectors (or scalars)."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim != 2:
            raise ValueError("X must be a 2-dimensional array.")
        self = object.__new__(cls)
        self.N, self.D = X.shape
        self.cov, self.K_max, self._ss = COV_FIXED, int(K_max), self.D
        vec = [np.ascontiguousarray(np.broadcast_to(np.asarray(v, dtype=np.float64), (self.D,))) for v in (var, mu_0, var_0)]
        h = C.c_void_p()
        _check(lib().bgmm_create_fixedvar(_dp(X), self.N, self.D, _dp(vec[0]), _dp(vec[1]), _dp(vec[2]), self.K_max,
                                          int(device), C.byref(h)))
        self._h = h
        return self

    def fork(self):
        """A new chain on the same data and prior: shares the device copy of X, the cached log prior and the tables
        (bgmm_fork); own labels, statistics and RNG state.  All data start unassigned."""
        other = object.__new__(Chain)
        other.N, other.D, other.cov, other.K_max, other._ss = self.N, self.D, self.cov, self.K_max, self._ss
        h = C.c_void_p()
        _check(lib().bgmm_fork(self._h, C.byref(h)))
        other._h = h
        return other

    def close(self):
        if getattr(self, "_h", None):
            lib().bgmm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def K(self):
        return lib().bgmm_K(self._h)

    def set_stream(self, cuda_stream):
        _check(lib().bgmm_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def set_engine(self, mode):
        _check(lib().bgmm_set_engine(self._h, {"adaptive": 0, "sequential": 1, "windows": 2, "generic": 3, "generic-sequential": 4,
                                                 "generic-windows": 5, "cluster": 6}.get(mode, mode)))

    def seed(self, seed):
        _check(lib().bgmm_seed(self._h, int(seed)))

    def sweep_index(self):
        return lib().bgmm_sweep_index(self._h)

    def get_uniforms(self, sweep_index):
        out = np.empty(self.N, np.float64)
        _check(lib().bgmm_get_uniforms(self._h, int(sweep_index), _dp(out)))
        return out

    def set_assignments(self, z):
        z = np.ascontiguousarray(z, dtype=np.int64)
        assert z.shape == (self.N,)
        _check(lib().bgmm_set_assignments(self._h, _ip(z)))

    def sweep(self, alpha, power=1.0, order=None, uniforms=None):
        o = None if order is None else np.ascontiguousarray(order, dtype=np.int64)
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        assert o is None or o.shape == (self.N,)
        assert u is None or u.shape == (self.N,)
        st = SweepStats()
        _check(lib().bgmm_sweep(self._h, _ip(o), _dp(u), float(alpha), float(power), C.byref(st)))
        return st

    def sweep_constrained(self, alpha, power, order, uniforms, status):
        """One sweep with CSCRPMM's constrained re-draw (bgmm_sweep_constrained).  `uniforms`: the random.random() stream
        from the sweep's first draw on (>= N values); `status[k]`: 1 useful / 2 non-useful slot at the start of the sweep.
        Returns (SweepStats, number of uniforms consumed)."""
        o = None if order is None else np.ascontiguousarray(order, dtype=np.int64)
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        sv = np.ascontiguousarray(status, dtype=np.int32)
        st = SweepStats()
        used = C.c_int64()
        _check(lib().bgmm_sweep_constrained(self._h, _ip(o), _dp(u), len(u), float(alpha), float(power),
                                            sv.ctypes.data_as(C.POINTER(C.c_int32)), len(sv), C.byref(used), C.byref(st)))
        return st, used.value

    def sweep_dev(self, alpha, power=1.0, d_order=0, d_uniforms=0):
        """`d_order` / `d_uniforms`: raw device addresses (e.g. torch.Tensor.data_ptr()) or 0."""
        st = SweepStats()
        _check(lib().bgmm_sweep_dev(self._h, C.c_void_p(int(d_order) or None), C.c_void_p(int(d_uniforms) or None),
                                    float(alpha), float(power), C.byref(st)))
        return st

    def assignments_to_device(self, d_out):
        _check(lib().bgmm_get_assignments_dev(self._h, C.c_void_p(int(d_out))))

    def get_state(self, z=True, counts=True, m_num=True, S_part=True, logdet=True, inv_covar=True):
        out = {}
        za = np.empty(self.N, np.int64) if z else None
        ca = np.empty(self.K_max, np.int64) if counts else None
        ma = np.empty((self.K_max, self.D), np.float64) if m_num else None
        shp = (self.K_max, self.D, self.D) if self.cov == COV_FULL else (self.K_max, self.D)
        sa = np.empty(shp, np.float64) if S_part else None
        la = np.empty(self.K_max, np.float64) if logdet else None
        ia = np.empty(shp, np.float64) if inv_covar else None
        K = C.c_int32()
        _check(lib().bgmm_get_state(self._h, _ip(za), _ip(ca), C.byref(K), _dp(ma), _dp(sa), _dp(la), _dp(ia)))
        out.update(K=K.value, z=za, counts=ca, m_num=ma, S_part=sa, logdet=la, inv_covar=ia)
        return out

    def assignments(self):
        return self.get_state(counts=False, m_num=False, S_part=False, logdet=False, inv_covar=False)["z"]

    def log_prior(self):
        out = np.empty(self.N, np.float64)
        _check(lib().bgmm_log_prior(self._h, _dp(out)))
        return out

    def log_post_pred(self, idx):
        idx = np.ascontiguousarray(np.atleast_1d(idx), dtype=np.int64)
        K = self.K
        out = np.empty((len(idx), K), np.float64)
        _check(lib().bgmm_log_post_pred(self._h, _ip(idx), len(idx), _dp(out)))
        return out

    def log_marg_k(self):
        out = np.empty(max(self.K, 1), np.float64)
        _check(lib().bgmm_log_marg_k(self._h, _dp(out)))
        return out[:self.K]

    def log_marg(self, alpha):
        out = C.c_double()
        _check(lib().bgmm_log_marg(self._h, float(alpha), C.cast(C.byref(out), C.POINTER(C.c_double))))
        return out.value

    def set_true_labels(self, labels):
        """Upload ground-truth labels (any integers; they are ranked like np.unique does) for `contingency`."""
        uniq, inv = np.unique(np.asarray(labels).ravel(), return_inverse=True)
        if inv.shape != (self.N,):
            raise ValueError("labels_true and labels_pred must have same size, got %d and %d" % (inv.size, self.N))
        inv = np.ascontiguousarray(inv, dtype=np.int64)
        _check(lib().bgmm_set_true_labels(self._h, _ip(inv), len(uniq)))
        self.T_true = len(uniq)

    def contingency(self):
        """(T, K + 1) int64 table of (true label, component); the last column counts unassigned data."""
        out = np.empty((self.T_true, self.K + 1), np.int64)
        _check(lib().bgmm_contingency(self._h, _ip(out)))
        return out

    def cluster_ssq(self):
        """Per component: sum of squared distances of its members to their mean (from the statistics)."""
        out = np.empty(max(self.K, 1), np.float64)
        _check(lib().bgmm_cluster_ssq(self._h, _dp(out)))
        return out[:self.K]

    def set_label(self, i, k):
        """Rewrite the label of datum i (-1: unassigned); no statistic changes (crpmm.py:84-85 does this on the host)."""
        _check(lib().bgmm_set_label(self._h, int(i), int(k)))

    def set_state(self, z, m_num=None, S_part=None):
        """Labels, and optionally the exact bits of saved statistics (K x D, K x D x D | K x D)."""
        z = np.ascontiguousarray(z, dtype=np.int64)
        K = int(z.max()) + 1
        m = None if m_num is None else np.ascontiguousarray(m_num[:K], dtype=np.float64)
        S = None if S_part is None else np.ascontiguousarray(S_part[:K], dtype=np.float64)
        _check(lib().bgmm_set_state(self._h, _ip(z), K, _dp(m), _dp(S)))

    def set_guard(self, guard):
        _check(lib().bgmm_set_guard(self._h, float(guard)))

    def add_item(self, i, k):
        _check(lib().bgmm_add_item(self._h, int(i), int(k)))

    def del_item(self, i):
        _check(lib().bgmm_del_item(self._h, int(i)))


class ChainGroup(object):
    """Independent chains resident on one GPU that are advanced together: one kernel launch per sweep, one thread block
    per chain (bgmm_sweep_many).  The chains must share device, stream, D, covariance type and K_max -- e.g. a chain and
    its forks.  Every chain walks exactly the chain `Chain.sweep_dev` would walk for the same inputs."""

    def __init__(self, chains):
        self.chains = list(chains)
        self._hs = (C.c_void_p * len(self.chains))(*[c._h for c in self.chains])

    def __len__(self):
        return len(self.chains)

    def _ptrs(self, d):
        """None, a list of device addresses (0 / None allowed), or a 2-D torch tensor (one row per chain)."""
        n = len(self.chains)
        if d is None:
            return None
        if hasattr(d, "data_ptr"):
            assert d.dim() == 2 and d.shape[0] == n and d.is_contiguous()
            step = d.shape[1] * d.element_size()
            addrs = [d.data_ptr() + c * step for c in range(n)]
        else:
            addrs = [int(a) if a else None for a in d]
            assert len(addrs) == n
        return (C.c_void_p * n)(*addrs)

    def sweep_dev(self, alpha, power=1.0, d_orders=None, d_uniforms=None):
        """One sweep of every chain.  d_orders / d_uniforms: per-chain DEVICE buffers (see _ptrs), or None (data order /
        each chain's own Philox stream).  Returns the per-chain SweepStats."""
        n = len(self.chains)
        out = (SweepStats * n)()
        _check(lib().bgmm_sweep_many(self._hs, n, self._ptrs(d_orders), self._ptrs(d_uniforms), float(alpha),
                                     float(power), out))
        return list(out)
