// bgmm_engine.cu -- auxiliary kernels + the C-ABI (include/bgmm_b200.h) of libbgmm_b200.so.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a (see __graft_entry__.build()).
#include "bgmm_ops.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>


// =============================================================================================
// auxiliary kernels
// =============================================================================================

// Initial build of the sufficient statistics in the reference's order (gaussian_components.py:109-111:
// for k ascending, for i ascending with z[i]==k: add_item).  One CTA per component, one thread per
// statistic entry, serial over the component's members so every sum is accumulated in the same order
// as the reference (bit-identical).  idx: data indices bucketed by component, ascending inside a bucket.
template <int COV>
__global__ void k_build_stats(const Params p, const long long *__restrict__ idx, const long long *__restrict__ start,
                              int DP) {
    const int k = blockIdx.x;
    const int D = p.D;
    const int SS = stat_len(DP, COV);
    const long long lo = start[k], hi = start[k + 1];
    const int nel = (COV == COV_FULL ? packed_len(D) : D) + D;
    for (int e = threadIdx.x; e < nel; e += blockDim.x) {
        int a, b;
        bool is_num = false;
        if (COV == COV_FULL) {
            if (e < packed_len(D)) decode_row_idx(e, a, b);
            else { is_num = true; a = b = e - packed_len(D); }
        } else {
            if (e < D) { a = b = e; }
            else { is_num = true; a = b = e - D; }
        }
        double acc;
        if (COV == COV_FIXED) {
            // gaussian_components_fixedvar.py:155-158: precision_0 * mu_0 (+= precision * x), precision_0 (+= precision)
            acc = is_num ? __dmul_rn(p.S0[a], p.m0[a]) : p.S0[a];
            for (long long t = lo; t < hi; ++t) {
                if (is_num) acc = __dadd_rn(acc, __dmul_rn(p.tau[a], p.X[(size_t)idx[t] * DP + a]));
                else acc = __dadd_rn(acc, p.tau[a]);
            }
            if (is_num) p.num[(size_t)k * DP + a] = acc;
            else p.S[(size_t)k * SS + a] = acc;
            continue;
        }
        if (is_num) acc = __dmul_rn(p.k0, p.m0[a]);
        else acc = __dadd_rn(p.S0[COV == COV_FULL ? row_idx(a, b) : a], __dmul_rn(p.k0, __dmul_rn(p.m0[a], p.m0[b])));
        for (long long t = lo; t < hi; ++t) {
            const long long i = idx[t];
            const double xa = p.X[(size_t)i * DP + a];
            if (is_num) acc = __dadd_rn(acc, xa);
            else acc = __dadd_rn(acc, __dmul_rn(xa, p.X[(size_t)i * DP + b]));
        }
        if (is_num) p.num[(size_t)k * DP + a] = acc;
        else p.S[(size_t)k * SS + (COV == COV_FULL ? row_idx(a, b) : a)] = acc;
    }
}

__global__ void k_relabel(const int *__restrict__ z_uid, const int *__restrict__ slot_of_uid, long long N,
                          long long *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int uid = z_uid[i];
    out[i] = uid < 0 ? -1LL : (long long)slot_of_uid[uid];
}

__global__ void k_philox(double *__restrict__ u, long long N, unsigned long long seed, unsigned long long sweep) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < N) u[j] = philox_uniform(seed, sweep, (unsigned long long)j);
}

// The scan order must be a permutation of 0..N-1 (pcrpmm.py:89 draws np.random.permutation): an index out of range
// would be an out-of-bounds read of X / z, and a repeated index would be visited twice while the replicas of the
// resident engine still read the sweep-start label copy.  One pass: every index sets its bit; a bit already set or an
// index out of range writes BGMM_EINVAL to the handle's error word (checked before the sweep kernel is launched).
__global__ void k_check_perm(const long long *__restrict__ order, long long N, unsigned int *__restrict__ bits, int *err) {
    const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N) return;
    const long long i = order[j];
    if (i < 0 || i >= N) { *err = BGMM_EINVAL; return; }
    const unsigned int m = 1u << (i & 31);
    if (atomicOr(bits + (i >> 5), m) & m) *err = BGMM_EINVAL;
}

// ---------------------------------------------------------------------------------------------
// Per-sweep record (GMM.update_record_dict, gmm/gmm.py:65-118) without moving the labels to the host.
//
// k_contingency: table[t][k] = #{i : true_label[i] == t and datum i sits in component slot k}, column K = unassigned.
// This is the one O(N) pass behind nmi / mi / vi (infopy/infopy.py:62-119 makes a pass over the data per CELL).
// HBM-bound integer work: 8 B per datum (two int32 labels, 16-B vector loads, grid-stride, a multiple of the SM
// count of CTAs); counts go to a CTA-private shared-memory table (32-bit, a CTA sees < 2^31 data) that is folded
// into the global 64-bit table with one atomic per non-zero cell; tables too large for shared memory take global
// atomics directly (they live in L2).
// ---------------------------------------------------------------------------------------------
__global__ void k_contingency(const int *__restrict__ t_true, const int *__restrict__ z_uid,
                              const int *__restrict__ slot_of_uid, long long N, int T, int K, int in_smem,
                              unsigned long long *__restrict__ table) {
    extern __shared__ __align__(16) unsigned int cell[];
    const int C = K + 1, cells = T * C;
    if (in_smem) {
        for (int e = threadIdx.x; e < cells; e += blockDim.x) cell[e] = 0u;
        __syncthreads();
    }
    const long long nvec = N >> 2;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec + 1; v += stride) {
        int tt[4], uu[4], n = 4;
        if (v < nvec) {
            const int4 a = __ldg(reinterpret_cast<const int4 *>(t_true) + v);
            const int4 b = __ldcs(reinterpret_cast<const int4 *>(z_uid) + v);
            tt[0] = a.x; tt[1] = a.y; tt[2] = a.z; tt[3] = a.w;
            uu[0] = b.x; uu[1] = b.y; uu[2] = b.z; uu[3] = b.w;
        } else {  // the ragged tail (N % 4 data), taken by exactly one thread
            n = (int)(N & 3);
            for (int q = 0; q < n; ++q) { tt[q] = t_true[(nvec << 2) + q]; uu[q] = z_uid[(nvec << 2) + q]; }
        }
        for (int q = 0; q < n; ++q) {
            const int k = uu[q] < 0 ? K : __ldg(slot_of_uid + uu[q]);
            const int e = tt[q] * C + k;
            if (in_smem) atomicAdd(cell + e, 1u);
            else atomicAdd(table + e, 1ULL);
        }
    }
    if (in_smem) {
        __syncthreads();
        for (int e = threadIdx.x; e < cells; e += blockDim.x) {
            const unsigned int c = cell[e];
            if (c) atomicAdd(table + e, (unsigned long long)c);
        }
    }
}

// k_cluster_ssq: out[k] = sum_i |x_i - mean_k|^2 over the members of component k -- the quantity under the square
// root of utils.compute_dist (utils/utils.py:52-88) -- from the sufficient statistics alone:
//   sum_a [ sum x_a^2 - (sum x_a)^2 / n ],  sum x_a = num_a - k0 m0_a,  sum x_a^2 = S_aa - (S0_aa + k0 m0_a^2).
// One warp per component; no pass over the data.
template <int COV>
__global__ void k_cluster_ssq(const Params p, int DP, double *__restrict__ out) {
    const int k = blockIdx.x, lane = threadIdx.x & 31;
    const int SS = stat_len(DP, COV);
    const double n = (double)p.counts[k];
    double acc = 0.0;
    for (int a = lane; a < p.D; a += 32) {
        const int e = (COV == COV_FULL) ? row_idx(a, a) : a;
        const double base = __dadd_rn(p.S0[e], __dmul_rn(p.k0, __dmul_rn(p.m0[a], p.m0[a])));
        const double sxx = p.S[(size_t)k * SS + e] - base;
        const double sx = p.num[(size_t)k * DP + a] - __dmul_rn(p.k0, p.m0[a]);
        acc += sxx - sx * sx / n;
    }
    for (int o = 16; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[k] = acc > 0.0 ? acc : 0.0;
}

// log_marg_k for every live component (gaussian_components.py:253-276 / _diag.py:271-289); one warp each
template <int COV>
__global__ void k_log_marg_k(const Params p, int DP, double logdet_S0, double *__restrict__ out) {
    extern __shared__ __align__(16) double A[];
    const int k = blockIdx.x, lane = threadIdx.x & 31;
    const int D = p.D;
    const long long n = p.counts[k];
    const double kap = p.k0 + (double)n;
    const long long vN = p.v0 + n;
    const double *num = p.num + (size_t)k * DP;
    const double *S = p.S + (size_t)k * stat_len(DP, COV);
    double ldN = 0.0;
    bool bad = false;
    if (COV == COV_FULL) {
        for (int a = 0; a < D; ++a) {
            const double ma = num[a] / kap;
            for (int b = lane; b <= a; b += 32) A[row_idx(a, b)] = S[row_idx(a, b)] - kap * (ma * (num[b] / kap));
        }
        __syncwarp();
        for (int j = 0; j < D; ++j) {
            const double ajj = A[row_idx(j, j)];
            if (!(ajj > 0.0)) { bad = true; break; }
            const double inv = 1.0 / sqrt(ajj);
            ldN += log(ajj);
            __syncwarp();
            for (int a = j + 1 + lane; a < D; a += 32) A[row_idx(a, j)] *= inv;
            __syncwarp();
            for (int a = j + 1 + lane; a < D; a += 32) {
                const double laj = A[row_idx(a, j)];
                for (int b = j + 1; b <= a; ++b) A[row_idx(a, b)] = fma(-laj, A[row_idx(b, j)], A[row_idx(a, b)]);
            }
            __syncwarp();
        }
    } else {
        for (int a = 0; a < D; ++a) {
            const double m = num[a] / kap;
            const double sn = S[a] - kap * (m * m);
            if (!(sn > 0.0)) bad = true;
            ldN += log(sn);
        }
    }
    if (lane == 0) {
        double r;
        if (bad) {
            r = NAN;
        } else if (COV == COV_FULL) {
            double g = 0.0;
            for (int j = 1; j <= D; ++j) g += p.lgam[vN + 1 - j] - p.lgam[p.v0 + 1 - j];
            r = -(double)n * D / 2. * p.log_pi + D / 2. * log(p.k0) - D / 2. * log(kap) + p.v0 / 2. * logdet_S0 -
                vN / 2. * ldN + g;
        } else {
            r = -(double)n * D / 2. * p.log_pi + D / 2. * log(p.k0) - D / 2. * log(kap) + p.v0 / 2. * logdet_S0 -
                vN / 2. * ldN + D * (p.lgam[vN] - p.lgam[p.v0]);
        }
        out[k] = r;
    }
}

// log_marg_k of the fixed-variance components (gaussian_components_fixedvar.py:234-256, Murphy's bayesGauss (55)): the
// formula needs sum x and sum x^2 over the component's members, which are not part of its statistics, so block k makes
// one pass over the labels (8 dimensions at a time, register accumulators, a deterministic tree over the block).
__global__ void k_fixed_log_marg_k(const Params p, int DP, double *__restrict__ out) {
    __shared__ double red[256];
    const int k = blockIdx.x, tid = threadIdx.x;
    const int D = p.D;
    const double n = (double)p.counts[k];
    double total = 0.0;
    for (int a0 = 0; a0 < D; a0 += 8) {
        double sx[8], sxx[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) sx[t] = sxx[t] = 0.0;
        for (long long i = tid; i < p.N; i += blockDim.x) {
            const int uid = p.z_uid[i];
            if (uid < 0 || p.slot_of_uid[uid] != k) continue;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                if (a0 + t < D) {
                    const double x = p.X[(size_t)i * DP + a0 + t];
                    sx[t] += x;
                    sxx[t] += x * x;
                }
            }
        }
        for (int t = 0; t < 8 && a0 + t < D; ++t) {
            double v[2] = {sx[t], sxx[t]};
            for (int w = 0; w < 2; ++w) {
                red[tid] = v[w];
                __syncthreads();
                for (int o = 128; o >= 1; o >>= 1) {
                    if (tid < o) red[tid] += red[tid + o];
                    __syncthreads();
                }
                v[w] = red[0];
                __syncthreads();
            }
            if (tid == 0) {
                const int a = a0 + t;
                const double tau = p.tau[a], tau0 = p.S0[a], mu0 = p.m0[a];
                const double s = n / tau0 + 1. / tau;
                total += (n - 1) / 2. * log(tau) - 0.5 * n * log(2 * M_PI) - 0.5 * log(s) - 0.5 * tau * v[1] -
                         0.5 * tau0 * (mu0 * mu0) +
                         0.5 * ((v[0] * v[0]) * tau / tau0 + (mu0 * mu0) * tau0 / tau + 2 * v[0] * mu0) / s;
            }
        }
    }
    if (tid == 0) out[k] = total;
}

// inv_covars / logdet_covars views (gaussian_components.py:88-89) reconstructed from the Cholesky record:
// inv = L^-T L^-1.  One warp per component; W (D x D) in shared memory.
template <int COV>
__global__ void k_inv_covar(const Params p, int DP, double *__restrict__ logdet_out, double *__restrict__ inv_out) {
    extern __shared__ __align__(16) double W[];
    const int k = blockIdx.x, lane = threadIdx.x & 31;
    const int D = p.D;
    const int R = rec_len(DP, COV);
    const double *rec = p.rec + (size_t)k * R;
    if (lane == 0) logdet_out[k] = rec[rec_sc_off(DP, COV) + SC_LOGDET];
    if (COV != COV_FULL) {   // diag: inv_vars; fixed variance: precision_preds
        for (int a = lane; a < D; a += 32) inv_out[(size_t)k * D + a] = rec[DP + a];
        return;
    }
    // column e of W = L^-1: forward substitution on the unit vector e_e
    for (int e = lane; e < D; e += 32) {
        for (int a = 0; a < D; ++a) {
            double s = (a == e) ? 1.0 : 0.0;
            for (int b = e; b < a; ++b) s -= rec[col_off(DP, b) + (a - b)] * W[b * D + e];
            W[a * D + e] = (a < e) ? 0.0 : s * rec[col_off(DP, a)];
        }
    }
    __syncwarp();
    for (int t = lane; t < D * D; t += 32) {
        const int a = t / D, b = t % D;
        double s = 0.0;
        for (int c = max(a, b); c < D; ++c) s += W[c * D + a] * W[c * D + b];
        inv_out[(size_t)k * D * D + t] = s;
    }
}

// =============================================================================================
// host side
// =============================================================================================
static thread_local std::string g_err;
int bgmm_fail(int code, const std::string &msg) {
    g_err = msg;
    return code;
}
static int fail(int code, const std::string &msg) { return bgmm_fail(code, msg); }
static int pad_dim(int D) {
    int dp = 1;
    while (dp < D) dp <<= 1;
    return dp;
}

// one getter per instantiation unit (csrc/inst/*.cu, generated); BGMM_HAVE_D<dp> says which were built
#define DECL(dp) const Ops *bgmm_ops_full_##dp(int (**prep)(bgmm_handle *)); const Ops *bgmm_ops_diag_##dp(int (**prep)(bgmm_handle *)); const Ops *bgmm_ops_fixed_##dp(int (**prep)(bgmm_handle *));
DECL(1) DECL(2) DECL(4) DECL(8) DECL(16) DECL(32) DECL(64)
#undef DECL
static const Ops *pick_ops(int cov, int DP, int (**prep)(bgmm_handle *)) {
    switch (DP) {
#define CASE(dp) case dp: return cov == BGMM_COV_FULL ? bgmm_ops_full_##dp(prep) : (cov == BGMM_COV_DIAG ? bgmm_ops_diag_##dp(prep) : bgmm_ops_fixed_##dp(prep));
#ifdef BGMM_HAVE_D1
        CASE(1)
#endif
#ifdef BGMM_HAVE_D2
        CASE(2)
#endif
#ifdef BGMM_HAVE_D4
        CASE(4)
#endif
#ifdef BGMM_HAVE_D8
        CASE(8)
#endif
#ifdef BGMM_HAVE_D16
        CASE(16)
#endif
#ifdef BGMM_HAVE_D32
        CASE(32)
#endif
#ifdef BGMM_HAVE_D64
        CASE(64)
#endif
#undef CASE
    }
    return nullptr;
}

static void free_all(bgmm_handle *h) {
    // per-chain state
    cudaFree(h->d_z); cudaFree(h->d_slot_of_uid); cudaFree(h->d_uid_of_slot); cudaFree(h->d_uid_free);
    cudaFree(h->d_counts); cudaFree(h->d_num); cudaFree(h->d_S); cudaFree(h->d_rec); cudaFree(h->d_wbuf);
    cudaFree(h->d_rec_prior); cudaFree(h->d_ctl); cudaFree(h->d_err); cudaFree(h->d_u); cudaFree(h->d_order);
    cudaFree(h->d_tmp_ll); cudaFree(h->d_recB); cudaFree(h->d_recB_prior); cudaFree(h->d_z2); cudaFree(h->d_mvbuf);
    cudaFree(h->d_true); cudaFree(h->d_table); cudaFree(h->d_pv); cudaFree(h->d_recBig); cudaFree(h->d_recClu); cudaFree(h->d_mlog); cudaFree(h->d_status); cudaFree(h->d_ubig);
    // buffers shared with the chains forked from / with this one: freed with the last of them
    if (h->sb && --h->sb->refs == 0) {
        SharedBufs *b = h->sb;
        cudaFree(b->dX); cudaFree(b->d_log_prior); cudaFree(b->d_lgam); cudaFree(b->d_logv); cudaFree(b->d_m0);
        cudaFree(b->d_S0); cudaFree(b->d_ntab); cudaFree(b->d_fmtab); cudaFree(b->d_tau);
        delete b;
    }
    h->sb = nullptr;
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev2) cudaEventDestroy(h->ev2);
    if (h->ev3) cudaEventDestroy(h->ev3);
}

// per-chain device state (labels, uid tables, statistics, records, control block, scratch): bgmm_create and bgmm_fork
static int alloc_chain_state(bgmm_handle *h) {
    const long long N = h->N;
    const int DP = h->DP, K_max = h->K_max;
    const int SS = stat_len(DP, h->cov), R = h->ops->rec_len;
#define ALLOC(ptr, bytes)                                                                             \
    do {                                                                                              \
        cudaError_t e_ = cudaMalloc((void **)&(ptr), (bytes));                                        \
        if (e_ != cudaSuccess)                                                                        \
            return fail(BGMM_ENOMEM, std::string("cudaMalloc ") + #ptr + ": " + cudaGetErrorString(e_)); \
    } while (0)
    ALLOC(h->d_z, sizeof(int) * (size_t)N);
    ALLOC(h->d_slot_of_uid, sizeof(int) * K_max);
    ALLOC(h->d_uid_of_slot, sizeof(int) * K_max);
    ALLOC(h->d_uid_free, sizeof(int) * K_max);
    ALLOC(h->d_counts, sizeof(long long) * (K_max + 1));
    ALLOC(h->d_num, sizeof(double) * (size_t)(K_max + 1) * DP);
    ALLOC(h->d_S, sizeof(double) * (size_t)(K_max + 1) * SS);
    ALLOC(h->d_rec, sizeof(double) * (size_t)(K_max + 1) * R);
    ALLOC(h->d_rec_prior, sizeof(double) * R);
    ALLOC(h->d_ctl, sizeof(Ctl));
    ALLOC(h->d_err, sizeof(int));
    ALLOC(h->d_u, sizeof(double) * (size_t)N);
    ALLOC(h->d_order, sizeof(long long) * (size_t)N);
    ALLOC(h->d_tmp_ll, sizeof(long long) * (size_t)(N + K_max + 2));
#undef ALLOC
    CU(cudaEventCreate(&h->ev0));
    CU(cudaEventCreate(&h->ev1));
    CU(cudaEventCreate(&h->ev2));
    CU(cudaEventCreate(&h->ev3));
    CU(cudaMemset(h->d_err, 0, sizeof(int)));
    CU(cudaMemset(h->d_z, 0xff, sizeof(int) * (size_t)N));
    CU(cudaMemset(h->d_counts, 0, sizeof(long long) * (K_max + 1)));
    CU(cudaMemset(h->d_num, 0, sizeof(double) * (size_t)(K_max + 1) * DP));
    CU(cudaMemset(h->d_S, 0, sizeof(double) * (size_t)(K_max + 1) * SS));
    CU(cudaMemset(h->d_rec, 0, sizeof(double) * (size_t)(K_max + 1) * R));
    std::vector<int> neg(K_max, -1), fr(K_max);
    for (int t = 0; t < K_max; ++t) fr[t] = K_max - 1 - t;
    CU(cudaMemcpy(h->d_slot_of_uid, neg.data(), sizeof(int) * K_max, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_uid_of_slot, neg.data(), sizeof(int) * K_max, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_uid_free, fr.data(), sizeof(int) * K_max, cudaMemcpyHostToDevice));
    Ctl c;
    memset(&c, 0, sizeof(c));
    c.n_free = K_max;
    c.first = POS_INF;
    c.first3[0][0] = c.first3[1][0] = c.first3[2][0] = (unsigned long long)POS_INF;
    const double one = 1.0;
    memcpy(&c.margin_bits, &one, 8);
    CU(cudaMemcpy(h->d_ctl, &c, sizeof(c), cudaMemcpyHostToDevice));
    h->K = 0;
    return 0;
}

static int check_dev_err(bgmm_handle *h, const char *what) {
    int e = 0;
    CU(cudaMemcpyAsync(&e, h->d_err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (e != 0) {
        int zero = 0;
        cudaMemcpyAsync(h->d_err, &zero, sizeof(int), cudaMemcpyHostToDevice, h->stream);
        if (e == BGMM_EINVAL) return fail(e, std::string(what) + ": the scan order is not a permutation of 0..N-1");
        return fail(e, std::string(what) + ": covariance not positive definite");
    }
    return 0;
}

extern "C" {

const char *bgmm_version(void) { return "pybgmm-b200 0.1.0 (sm_100a)"; }
const char *bgmm_last_error(void) { return g_err.c_str(); }
int bgmm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

}  // extern "C"

// bgmm_create and bgmm_create_fixedvar: var_fixed is the known data variance of the fixed-variance components (then m0 =
// mu_0, S0 = var_0, and k0 / v0 are not used) and NULL otherwise
static int create_impl(const double *X, int64_t N, int32_t D, int32_t cov_type, const double *m0, double k0, int64_t v0,
                       const double *S0, const double *var_fixed, int32_t K_max, const double *lgamma_half_tab,
                       const double *log_tab, int64_t tab_len, int32_t device, bgmm_t **out) {
    if (!out) return fail(BGMM_EINVAL, "out is NULL");
    *out = nullptr;
    if (!X || !m0 || !S0) return fail(BGMM_EINVAL, "X, m0, S0 must not be NULL");
    if (N < 1 || D < 1) return fail(BGMM_EINVAL, "N and D must be >= 1");
    if (D > 64) return fail(BGMM_EINVAL, "D > 64 is not supported by this build");
    if (cov_type != BGMM_COV_FULL && cov_type != BGMM_COV_DIAG && cov_type != BGMM_COV_FIXED)
        return fail(BGMM_EINVAL, "invalid covariance type");
    if ((cov_type == BGMM_COV_FIXED) != (var_fixed != nullptr)) return fail(BGMM_EINVAL, "invalid covariance type");
    if (v0 < D) return fail(BGMM_EINVAL, "v_0 must be >= D (prior/niw.py:21)");
    if (!(k0 > 0.0)) return fail(BGMM_EINVAL, "k_0 must be > 0");
    if (K_max < 1 || K_max > 4096) return fail(BGMM_EINVAL, "K_max must be in [1, 4096]");
    if (N >= (1LL << 31)) return fail(BGMM_EINVAL, "N must be < 2^31");
    const int64_t need = v0 + N + 2;
    if ((lgamma_half_tab || log_tab) && (!lgamma_half_tab || !log_tab || tab_len < need))
        return fail(BGMM_EINVAL, "lgamma/log tables must both be given with tab_len >= v0 + N + 2");
    if (bgmm_device_count() <= device || device < 0)
        return fail(BGMM_ENODEV, "no CUDA device " + std::to_string(device) + " (this library has no CPU fallback)");
    CU(cudaSetDevice(device));

    bgmm_handle *h = new bgmm_handle();
    h->device = device; h->N = N; h->D = D; h->DP = pad_dim(D); h->cov = cov_type; h->K_max = K_max;
    h->k0 = k0; h->v0 = v0;
    if (const char *g = getenv("BGMM_GUARD")) h->guard = atof(g);
    if (const char *w = getenv("BGMM_WATCHDOG_S")) h->watchdog_ns = (long long)(atof(w) * 1e9);
    const int DP = h->DP;
    int (*prep)(bgmm_handle *) = nullptr;
    h->ops = pick_ops(cov_type, DP, &prep);
    if (!h->ops) { delete h; return fail(BGMM_EINVAL, "this build has no kernels for D=" + std::to_string(D)); }
    const int SS = stat_len(DP, cov_type);
    const int R = h->ops->rec_len;

    // prior, padded / packed
    h->m0.assign(DP, 0.0);
    for (int a = 0; a < D; ++a) h->m0[a] = m0[a];
    h->S0p.assign(SS, 0.0);
    if (cov_type == BGMM_COV_FULL) {
        for (int a = 0; a < D; ++a)
            for (int b = 0; b <= a; ++b) h->S0p[row_idx(a, b)] = S0[a * D + b];
        // log|S_0| for log_marg_k (gaussian_components.py:271): Cholesky on the host (prior constant)
        std::vector<double> A(D * D);
        for (int a = 0; a < D; ++a) for (int b = 0; b < D; ++b) A[a * D + b] = S0[a * D + b];
        double ld = 0.0;
        for (int j = 0; j < D; ++j) {
            double d = A[j * D + j];
            for (int c = 0; c < j; ++c) d -= A[j * D + c] * A[j * D + c];
            if (!(d > 0.0)) { delete h; return fail(BGMM_ENUMERIC, "S_0 is not positive definite"); }
            const double l = sqrt(d);
            A[j * D + j] = l;
            ld += 2.0 * log(l);
            for (int a = j + 1; a < D; ++a) {
                double s = A[a * D + j];
                for (int c = 0; c < j; ++c) s -= A[a * D + c] * A[j * D + c];
                A[a * D + j] = s / l;
            }
        }
        h->logdet_S0 = ld;
    } else if (cov_type == BGMM_COV_FIXED) {
        // precision = 1 / var, precision_0 = 1 / var_0   (gaussian_components_fixedvar.py:80-82)
        h->tauv.assign(DP, 0.0);
        for (int a = 0; a < D; ++a) {
            if (!(S0[a] > 0.0) || !(var_fixed[a] > 0.0)) { delete h; return fail(BGMM_ENUMERIC, "var and var_0 must be positive"); }
            h->S0p[a] = 1. / S0[a];
            h->tauv[a] = 1. / var_fixed[a];
        }
    } else {
        double ld = 0.0;
        for (int a = 0; a < D; ++a) {
            if (!(S0[a] > 0.0)) { delete h; return fail(BGMM_ENUMERIC, "S_0 must be positive"); }
            h->S0p[a] = S0[a];
            ld += log(S0[a]);
        }
        h->logdet_S0 = ld;
    }

    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    h->num_sms = prop.multiProcessorCount;
    h->grid = h->num_sms;
    const size_t smem_max = prop.sharedMemPerBlockOptin;
    const size_t fixed = h->ops->smem_fixed(D, K_max);
    if (fixed + 1024 > smem_max) { delete h; return fail(BGMM_EINVAL, "K_max too large for shared memory"); }
    size_t kc = (smem_max - fixed - 64) / ((size_t)R * sizeof(double));
    if (kc > (size_t)K_max) kc = K_max;
    h->Kc = (int)kc;
    h->smem_bytes = fixed + (size_t)h->Kc * R * sizeof(double) + 32;
    h->smem_item = fixed + 32;
    h->smem_optin = smem_max;
    if (int rc = prep(h)) { delete h; return rc; }
    h->sb = new SharedBufs();
    SharedBufs *b = h->sb;
#define ALLOC(ptr, bytes)                                                                             \
    do {                                                                                              \
        cudaError_t e_ = cudaMalloc((void **)&(ptr), (bytes));                                        \
        if (e_ != cudaSuccess) {                                                                      \
            free_all(h); delete h;                                                                    \
            return fail(BGMM_ENOMEM, std::string("cudaMalloc ") + #ptr + ": " + cudaGetErrorString(e_)); \
        }                                                                                             \
    } while (0)
    ALLOC(b->dX, sizeof(double) * (size_t)N * DP);
    ALLOC(b->d_log_prior, sizeof(double) * (size_t)N);
    ALLOC(b->d_lgam, sizeof(double) * (size_t)need);
    ALLOC(b->d_logv, sizeof(double) * (size_t)need);
    ALLOC(b->d_m0, sizeof(double) * DP);
    ALLOC(b->d_S0, sizeof(double) * SS);
    if (cov_type == BGMM_COV_FIXED) ALLOC(b->d_tau, sizeof(double) * DP);
#undef ALLOC
    h->dX = b->dX; h->d_log_prior = b->d_log_prior; h->d_lgam = b->d_lgam; h->d_logv = b->d_logv;
    h->d_m0 = b->d_m0; h->d_S0 = b->d_S0; h->d_tau = b->d_tau;
    if (b->d_tau) CU(cudaMemcpy(b->d_tau, h->tauv.data(), sizeof(double) * DP, cudaMemcpyHostToDevice));
    if (int rc = h->ops->fast_setup(h)) { free_all(h); delete h; return rc; }
    if (int rc = h->ops->big_setup(h)) { free_all(h); delete h; return rc; }
    if (int rc = h->ops->clu_setup(h)) { free_all(h); delete h; return rc; }
    if (int rc = alloc_chain_state(h)) { free_all(h); delete h; return rc; }

    // uploads
    if (DP == D) {
        CU(cudaMemcpy(h->dX, X, sizeof(double) * (size_t)N * D, cudaMemcpyHostToDevice));
    } else {
        CU(cudaMemset(h->dX, 0, sizeof(double) * (size_t)N * DP));
        CU(cudaMemcpy2D(h->dX, sizeof(double) * DP, X, sizeof(double) * D, sizeof(double) * D, (size_t)N,
                        cudaMemcpyHostToDevice));
    }
    if (lgamma_half_tab) {
        CU(cudaMemcpy(h->d_lgam, lgamma_half_tab, sizeof(double) * (size_t)need, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->d_logv, log_tab, sizeof(double) * (size_t)need, cudaMemcpyHostToDevice));
    } else {  // n = [1, 1, 2, ..., v0+N+1]  (gaussian_components.py:120-122), libm instead of SciPy
        std::vector<double> lg((size_t)need), lv((size_t)need);
        for (int64_t t = 0; t < need; ++t) {
            const double n = (t == 0) ? 1.0 : (double)t;
            lg[t] = lgamma(n / 2.);
            lv[t] = log(n);
        }
        CU(cudaMemcpy(h->d_lgam, lg.data(), sizeof(double) * (size_t)need, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(h->d_logv, lv.data(), sizeof(double) * (size_t)need, cudaMemcpyHostToDevice));
    }
    CU(cudaMemcpy(h->d_m0, h->m0.data(), sizeof(double) * DP, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_S0, h->S0p.data(), sizeof(double) * SS, cudaMemcpyHostToDevice));
    // cached_log_prior
    Params p = make_params(h);
    if (int rc = h->ops->log_prior(h, p)) { free_all(h); delete h; return rc; }
    if (int rc = check_dev_err(h, "log_prior")) { free_all(h); delete h; return rc; }
    *out = h;
    return 0;
}

extern "C" {

int bgmm_create(const double *X, int64_t N, int32_t D, int32_t cov_type, const double *m0, double k0, int64_t v0,
                const double *S0, int32_t K_max, const double *lgamma_half_tab, const double *log_tab, int64_t tab_len,
                int32_t device, bgmm_t **out) {
    if (cov_type != BGMM_COV_FULL && cov_type != BGMM_COV_DIAG) return fail(BGMM_EINVAL, "invalid covariance type");
    return create_impl(X, N, D, cov_type, m0, k0, v0, S0, nullptr, K_max, lgamma_half_tab, log_tab, tab_len, device, out);
}

int bgmm_create_fixedvar(const double *X, int64_t N, int32_t D, const double *var, const double *mu_0, const double *var_0,
                         int32_t K_max, int32_t device, bgmm_t **out) {
    if (!var) return fail(BGMM_EINVAL, "var must not be NULL");
    return create_impl(X, N, D, BGMM_COV_FIXED, mu_0, 1.0, D, var_0, var, K_max, nullptr, nullptr, 0, device, out);
}

int bgmm_fork(bgmm_t *parent, bgmm_t **out) {
    if (!parent || !out) return fail(BGMM_EINVAL, "NULL argument");
    *out = nullptr;
    CU(cudaSetDevice(parent->device));
    bgmm_handle *h = new bgmm_handle();
    h->device = parent->device; h->stream = parent->stream;
    h->N = parent->N; h->D = parent->D; h->DP = parent->DP; h->cov = parent->cov; h->K_max = parent->K_max; h->Kc = parent->Kc;
    h->k0 = parent->k0; h->v0 = parent->v0; h->m0 = parent->m0; h->S0p = parent->S0p; h->logdet_S0 = parent->logdet_S0;
    h->num_sms = parent->num_sms; h->grid = parent->grid; h->smem_bytes = parent->smem_bytes;
    h->smem_item = parent->smem_item; h->smem_optin = parent->smem_optin; h->ops = parent->ops;
    h->guard = parent->guard; h->watchdog_ns = parent->watchdog_ns; h->engine = parent->engine;
    h->sb = parent->sb;
    h->sb->refs += 1;
    SharedBufs *b = h->sb;
    h->dX = b->dX; h->d_log_prior = b->d_log_prior; h->d_lgam = b->d_lgam; h->d_logv = b->d_logv;
    h->d_m0 = b->d_m0; h->d_S0 = b->d_S0; h->d_tau = b->d_tau;
    if (int rc = h->ops->fast_setup(h)) { free_all(h); delete h; return rc; }
    if (int rc = h->ops->big_setup(h)) { free_all(h); delete h; return rc; }
    if (int rc = h->ops->clu_setup(h)) { free_all(h); delete h; return rc; }
    if (int rc = alloc_chain_state(h)) { free_all(h); delete h; return rc; }
    // the prior's generic record (bgmm_log_prior's source) is per handle: copy the parent's
    CU(cudaMemcpy(h->d_rec_prior, parent->d_rec_prior, sizeof(double) * h->ops->rec_len, cudaMemcpyDeviceToDevice));
    *out = h;
    return 0;
}

int bgmm_destroy(bgmm_t *h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    free_all(h);
    delete h;
    return 0;
}

int bgmm_set_stream(bgmm_t *h, void *cuda_stream) {
    if (!h) return fail(BGMM_EINVAL, "handle is NULL");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    h->stream = (cudaStream_t)cuda_stream;
    return 0;
}

int bgmm_set_engine(bgmm_t *h, int32_t mode) {
    if (!h || mode < 0 || mode > 6) return fail(BGMM_EINVAL, "engine mode must be in 0..6");
    h->engine = mode;
    return 0;
}

int bgmm_seed(bgmm_t *h, uint64_t seed) {
    if (!h) return fail(BGMM_EINVAL, "handle is NULL");
    h->seed = seed;
    h->sweep_index = 0;
    return 0;
}
int64_t bgmm_sweep_index(bgmm_t *h) { return h ? h->sweep_index : -1; }

int bgmm_get_uniforms(bgmm_t *h, int64_t sweep_index, double *out) {
    if (!h || !out) return fail(BGMM_EINVAL, "NULL argument");
    CU(cudaSetDevice(h->device));
    double *tmp = (double *)h->d_tmp_ll;  // N doubles fit in the N long long scratch
    const int T = 256;
    k_philox<<<(unsigned)((h->N + T - 1) / T), T, 0, h->stream>>>(tmp, h->N, h->seed, (unsigned long long)sweep_index);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, tmp, sizeof(double) * (size_t)h->N, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int bgmm_set_assignments(bgmm_t *h, const int64_t *z) {
    if (!h || !z) return fail(BGMM_EINVAL, "NULL argument");
    CU(cudaSetDevice(h->device));
    const long long N = h->N;
    const int K_max = h->K_max, DP = h->DP;
    const int SS = stat_len(DP, h->cov), R = h->ops->rec_len;
    long long zmax = -1;
    for (long long i = 0; i < N; ++i) {
        if (z[i] < -1) return fail(BGMM_EINVAL, "assignments must be -1 or >= 0");
        zmax = std::max<long long>(zmax, z[i]);
    }
    const int K0 = (int)(zmax + 1);
    if (K0 > K_max) return fail(BGMM_EKMAX, "assignments use more than K_max components");
    std::vector<long long> start(K0 + 1, 0);
    for (long long i = 0; i < N; ++i) if (z[i] >= 0) start[z[i] + 1] += 1;
    for (int k = 0; k < K0; ++k) {
        // labels must be consecutive (gaussian_components.py:103-105 asserts)
        if (start[k + 1] == 0) return fail(BGMM_EINVAL, "assignments must be labelled 0..max without gaps");
    }
    std::vector<long long> counts(K_max + 1, 0);
    for (int k = 0; k < K0; ++k) counts[k] = start[k + 1];
    for (int k = 0; k < K0; ++k) start[k + 1] += start[k];
    std::vector<long long> idx((size_t)std::max<long long>(N, 1));
    {
        std::vector<long long> fill(start.begin(), start.end());
        for (long long i = 0; i < N; ++i) if (z[i] >= 0) idx[fill[z[i]]++] = i;
    }
    std::vector<int> zi((size_t)N);
    for (long long i = 0; i < N; ++i) zi[i] = (int)z[i];
    std::vector<int> sl(K_max, -1), fr(K_max, 0);
    for (int k = 0; k < K0; ++k) sl[k] = k;
    const int n_free = K_max - K0;
    for (int t = 0; t < n_free; ++t) fr[t] = K_max - 1 - t;

    cudaStream_t st = h->stream;
    long long *d_idx = h->d_tmp_ll, *d_start = h->d_tmp_ll + N;
    CU(cudaMemcpyAsync(d_idx, idx.data(), sizeof(long long) * (size_t)N, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(d_start, start.data(), sizeof(long long) * (K0 + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->d_z, zi.data(), sizeof(int) * (size_t)N, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->d_counts, counts.data(), sizeof(long long) * (K_max + 1), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->d_slot_of_uid, sl.data(), sizeof(int) * K_max, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->d_uid_of_slot, sl.data(), sizeof(int) * K_max, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->d_uid_free, fr.data(), sizeof(int) * K_max, cudaMemcpyHostToDevice, st));
    CU(cudaMemsetAsync(h->d_num, 0, sizeof(double) * (size_t)(K_max + 1) * DP, st));
    CU(cudaMemsetAsync(h->d_S, 0, sizeof(double) * (size_t)(K_max + 1) * SS, st));
    CU(cudaMemsetAsync(h->d_rec, 0, sizeof(double) * (size_t)(K_max + 1) * R, st));
    Ctl c;
    memset(&c, 0, sizeof(c));
    c.K = K0; c.n_free = n_free; c.first = POS_INF;
    c.first3[0][0] = c.first3[1][0] = c.first3[2][0] = (unsigned long long)POS_INF;
    const double one = 1.0;
    memcpy(&c.margin_bits, &one, 8);
    CU(cudaMemcpyAsync(h->d_ctl, &c, sizeof(c), cudaMemcpyHostToDevice, st));
    Params p = make_params(h);
    if (K0 > 0) {
        if (h->cov == BGMM_COV_FULL) k_build_stats<COV_FULL><<<K0, 256, 0, st>>>(p, d_idx, d_start, DP);
        else if (h->cov == BGMM_COV_DIAG) k_build_stats<COV_DIAG><<<K0, 256, 0, st>>>(p, d_idx, d_start, DP);
        else k_build_stats<COV_FIXED><<<K0, 256, 0, st>>>(p, d_idx, d_start, DP);
        CU(cudaGetLastError());
        if (int rc = h->ops->refactor_all(h, p, 0, K0)) return rc;
    }
    CU(cudaStreamSynchronize(st));  // host vectors go out of scope
    h->K = K0;
    h->last_gap = 0.0;
    return check_dev_err(h, "set_assignments");
}

static int run_sweep(bgmm_handle *h, const long long *d_order, const double *d_u, double alpha, double power,
                     bgmm_sweep_stats *out) {
    if (!(alpha > 0.0)) return fail(BGMM_EINVAL, "alpha must be > 0");
    if (!(power > 0.0)) return fail(BGMM_EINVAL, "power must be > 0");
    cudaStream_t st = h->stream;
    h->launches = 0;
    if (!d_u) {
        const int T = 256;
        h->launches += 1;
        k_philox<<<(unsigned)((h->N + T - 1) / T), T, 0, st>>>(h->d_u, h->N, h->seed, (unsigned long long)h->sweep_index);
        CU(cudaGetLastError());
        d_u = h->d_u;
    }
    // reset the per-sweep counters (keep K, n_free, barrier words)
    Ctl c;
    CU(cudaMemcpyAsync(&c, h->d_ctl, sizeof(c), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    c.moves = c.births = c.deaths = c.evals = c.windows = c.seq_data = c.wasted = 0;
    c.explicit_evals = c.refreshes = 0;
    c.guard_hits = c.fast_steps = 0;
    c.uextra = 0;
    c.watchdog_ns = h->watchdog_ns;
    memset(c.prof, 0, sizeof(c.prof));
    memset(c.wsum, 0, sizeof(c.wsum)); memset(c.wcnt, 0, sizeof(c.wcnt)); memset(c.wmax, 0, sizeof(c.wmax));
    memset(c.tprof, 0, sizeof(c.tprof));
    const double one = 1.0;
    memcpy(&c.margin_bits, &one, 8);
    c.error = 0; c.bar_count = 0; c.rb_count = 0; c.rb_word = 0; c.pos = 0; c.win = 0; c.first = POS_INF; c.n_dirty = 0;
    c.first3[0][0] = c.first3[1][0] = c.first3[2][0] = (unsigned long long)POS_INF;
    CU(cudaMemcpyAsync(h->d_ctl, &c, sizeof(c), cudaMemcpyHostToDevice, st));
    Params p = make_params(h);
    p.order = d_order; p.u = d_u;
    p.log_alpha = log(alpha); p.power = power;
    p.init_gap = h->last_gap;
    long long generic_from = -1;
    if (d_order) {
        // N bits of scratch: the relabel buffer is idle during a sweep
        unsigned int *bits = (unsigned int *)h->d_tmp_ll;
        CU(cudaMemsetAsync(bits, 0, sizeof(unsigned int) * (size_t)((h->N + 31) / 32), st));
        const int T = 256;
        k_check_perm<<<(unsigned)((h->N + T - 1) / T), T, 0, st>>>(d_order, h->N, bits, h->d_err);
        CU(cudaGetLastError());
        h->launches += 1;
        if (int rc = check_dev_err(h, "sweep")) return rc;
    }
    CU(cudaEventRecord(h->ev0, st));
    // engine 0..2: the replicated-state-machine engine (bgmm_fast.cuh) where it applies; 3..5: the generic engine
    const bool constrained = (h->cs_status != nullptr);   // re-draws consume a data-dependent number of uniforms: one CTA,
    if (constrained) p.engine = 1;                        // datum by datum, on the generic engine
    const bool resident = h->engine < 3 || h->engine == 6;   // 6: adaptive, the cluster step engine wherever it applies
    bool fast = h->fast_ok && resident && c.K <= h->Kcap && !constrained;
    // Cluster engines: full covariance on padded D = 32 / 64 (bgmm_big.cuh, the whole sweep) and the cluster step engine
    // for D <= 16 (bgmm_clu.cuh) while the movers are dense.  A launch walks the chain until the sweep ends, BIG_SPAN data
    // are done, a component has taken its share of rank-one changes, or a datum needs the general step (a birth, a death,
    // an explicit removal, a draw inside the margin guard); that datum is resolved by the generic engine's step, the
    // records are rebuilt from the bit-exact statistics, and the cluster continues behind it.  The step engine hands the
    // rest of the sweep to the resident engine's windows as soon as a span shows sparse movers.
    const bool use_big = !fast && h->big_ok && resident && c.K <= h->Kcap && c.K >= 1 && !constrained;
    const bool use_clu = fast && h->clu_ok && c.K >= 1 && c.K <= clu::KCH - 1 &&
                         (h->engine == 6 || (h->engine == 0 && h->last_gap < h->clu_gap));
    bool big_done = false;
    bool events_on = false;
    long long x_evals = 0, x_steps = 0, x_windows = 0, x_wasted = 0;   // data certified by k_big_window (host-side counts)
    long long resume_at = 0;   // the resident engine starts here when the step engine handed over
    if (use_big || use_clu) {
        CU(cudaEventRecord(h->ev2, st));
        events_on = true;
        long long pos = 0;
        bool to_generic = false, to_fast = false;
        long long handbacks = 0;
        long long span = (use_clu || h->engine == 0) ? 8192 : big::BIG_SPAN;   // a short first span: the regime shows early
        // D = 32 / 64, sparse movers: a data-parallel window kernel certifies the data that stay (k_big_window) and the
        // cluster steps only through the neighbourhood of the first datum that does not
        const double BIG_WIN_GAP = 16.0;
        const long long BIG_STEP_SPAN = 2;
        double gap_est = h->last_gap;
        bool win_mode = use_big && h->engine == 0 && gap_est >= BIG_WIN_GAP;
        long long wlen = 4096, since_prep = 0;
        bool need_prep = true;
        while (pos < h->N) {
            p.start_pos = pos;
            if (need_prep || use_clu || since_prep >= big::BIG_SPAN) {
                std::chrono::steady_clock::time_point tp0;
                const bool tprof = getenv("BGMM_WPROF") != nullptr && !use_clu;
                if (tprof) { CU(cudaStreamSynchronize(st)); tp0 = std::chrono::steady_clock::now(); }
                if (int rc = use_clu ? h->ops->clu_prep(h, p, c.K) : h->ops->big_prep(h, p, c.K)) return rc;
                if (int rc = check_dev_err(h, "sweep (record set-up)")) return rc;
                if (tprof) fprintf(stderr, "  prep: %.0f us\n", 1e6 * std::chrono::duration<double>(std::chrono::steady_clock::now() - tp0).count());
                need_prep = false;
                since_prep = 0;
            }
            long long stays = 0;
            if (win_mode) {
                const long long wlim = std::min<long long>(h->N, pos + wlen);
                const long long inf = POS_INF;
                CU(cudaMemcpyAsync((char *)h->d_ctl + offsetof(Ctl, first), &inf, sizeof(inf), cudaMemcpyHostToDevice, st));
                if (int rc = h->ops->big_window(h, p, c.K, pos, wlim)) return rc;
                CU(cudaMemcpyAsync(&c, h->d_ctl, sizeof(c), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                const long long f = std::min<long long>(c.first, wlim);
                stays = f - pos;
                x_evals += stays * c.K; x_steps += stays; x_windows += 1; x_wasted += wlim - f;
                since_prep += stays;
                if (getenv("BGMM_WPROF"))
                    fprintf(stderr, "  window from %lld to %lld: first non-stay %lld\n", pos, wlim, f);
                pos = f;
                if (f == wlim) { wlen = std::min<long long>(wlen * 2, 1LL << 20); continue; }
                wlen = std::max<long long>(4096, std::min<long long>(1LL << 20, (long long)(8.0 * gap_est)));   // one wave of tiles costs the same up to ~9 k positions
                p.start_pos = pos;
                span = BIG_STEP_SPAN;
            }
            const long long lim = std::min<long long>(h->N, pos + span);
            const long long moves_before = c.moves;
            if (int rc = use_clu ? h->ops->clu_sweep(h, p, lim) : h->ops->big_sweep(h, p, lim)) return rc;
            CU(cudaMemcpyAsync(&c, h->d_ctl, sizeof(c), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            if (getenv("BGMM_WPROF"))
                fprintf(stderr, "  cluster launch from %lld to %lld: moves logged %lld, %s, K=%d\n", pos, (long long)c.pos,
                        (long long)c.win, c.error == big::E_RARE ? "handed back" : "span done", c.K);
            const long long walked = c.pos - pos, moved = c.moves - moves_before;
            pos = c.pos;
            span = big::BIG_SPAN;
            since_prep += walked;
            if (use_big && h->engine == 0) {
                const double g_now = (double)(stays + walked) / (double)(moved + 1);
                gap_est = win_mode ? 0.75 * gap_est + 0.25 * g_now : g_now;
                win_mode = win_mode ? gap_est >= 0.5 * BIG_WIN_GAP : (walked >= 1024 && gap_est >= BIG_WIN_GAP);
                // between the regimes the cluster walks short spans, so that a dense patch of movers inside a sparse sweep
                // does not cost a whole BIG_SPAN of 4 us steps before the windows are tried again
                if (!win_mode && gap_est >= 4.0) span = 2048;
            }
            // the bit-exact statistics follow from the launch's move log (one CTA per component, chain order)
            if (int rc = h->ops->big_replay(h, p, c.K, c.win)) return rc;
            c.win = 0;
            if (c.error == big::E_RARE) {
                // the generic engine's (Cholesky) records follow from the statistics; its step resolves the datum
                handbacks += 1;
                c.error = 0;
                const bool tprof = getenv("BGMM_WPROF") != nullptr;
                std::chrono::steady_clock::time_point t0, t1, t2;
                if (tprof) { CU(cudaStreamSynchronize(st)); t0 = std::chrono::steady_clock::now(); }
                CU(cudaMemcpyAsync(h->d_ctl, &c, sizeof(c), cudaMemcpyHostToDevice, st));
                if (int rc = h->ops->refactor_all(h, p, 0, c.K)) return rc;
                h->launches += 1;
                need_prep = true;
                if (tprof) { CU(cudaStreamSynchronize(st)); t1 = std::chrono::steady_clock::now(); }
                if (int rc = h->ops->resolve_one(h, p, pos)) return rc;
                CU(cudaMemcpyAsync(&c, h->d_ctl, sizeof(c), cudaMemcpyDeviceToHost, st));
                CU(cudaStreamSynchronize(st));
                if (tprof) {
                    t2 = std::chrono::steady_clock::now();
                    fprintf(stderr, "  hand-back at %lld: refactor_all %.0f us, resolve_one %.0f us\n", pos,
                            1e6 * std::chrono::duration<double>(t1 - t0).count(), 1e6 * std::chrono::duration<double>(t2 - t1).count());
                }
                pos = c.pos;
                if (c.error != 0) break;
                if (c.K < 1 || c.K > (use_clu ? clu::KCH - 1 : h->Kcap)) { to_generic = true; break; }
            } else if (c.error != 0) {
                break;
            }
            if (use_clu && h->engine == 0 && walked >= 4096 && pos < h->N &&
                (double)walked / (double)(moved + 1) >= 1.25 * h->clu_gap) {
                to_fast = true;   // sparse movers: speculative windows are the better engine for the rest of the sweep
                break;
            }
            if (use_clu && h->engine == 0 && handbacks >= 16 && handbacks * 1500 > pos && pos < h->N) {
                to_fast = true;   // births / deaths every few hundred data: the resident engine resolves them in-kernel
                break;
            }
        }
        if (use_clu && (to_generic || to_fast) && pos < h->N && c.error == 0 && c.K <= h->Kcap) {
            resume_at = pos;      // the resident engine finishes the sweep (it also holds more components)
            p.init_gap = (double)h->N;   // ... starting in window mode
        } else if (to_generic && pos < h->N) {
            CU(cudaEventRecord(h->ev3, st));
            generic_from = pos;   // more live components than the cluster holds: the generic engine finishes the sweep
            fast = false;
        } else {
            CU(cudaEventRecord(h->ev3, st));
            big_done = true;
            fast = false;
            if (c.error == 0 && c.K > 0) {
                if (int rc = h->ops->refactor_all(h, p, 0, c.K)) return rc;
                h->launches += 1;
            }
        }
    }
    if (fast && !big_done) {
        // the replicas read the labels as they were at the start of the sweep; CTA 0 writes the other copy
        CU(cudaMemcpyAsync(h->d_z2, h->d_z, sizeof(int) * (size_t)h->N, cudaMemcpyDeviceToDevice, st));
        p.start_pos = resume_at;
        if (int rc = h->ops->fast_prep(h, p, c.K)) return rc;
        // a component whose S_N is not positive definite stops the sweep here, before the kernel would run on a
        // partially written record
        if (int rc = check_dev_err(h, "sweep (record set-up)")) return rc;
        if (!events_on) CU(cudaEventRecord(h->ev2, st));
        if (int rc = h->ops->fast_sweep(h, p)) return rc;
        CU(cudaEventRecord(h->ev3, st));
        h->launches += 1;
        std::swap(h->d_z, h->d_z2);
        p.z_uid = h->d_z; p.z_out = h->d_z2;
        CU(cudaMemcpyAsync(&c, h->d_ctl, sizeof(c), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        // the generic engine's (Cholesky) records, used by the auxiliary entry points, follow from the statistics
        if (c.K > 0) {
            if (int rc = h->ops->refactor_all(h, p, 0, c.K)) return rc;
            h->launches += 1;
        }
        if (c.error == fast::E_NEED_GENERIC) {  // more live components than fit in shared memory: continue generically
            generic_from = c.pos;
            c.error = 0; c.bar_count = 0;
            CU(cudaMemcpyAsync(h->d_ctl, &c, sizeof(c), cudaMemcpyHostToDevice, st));
            fast = false;
        }
    }
    const bool ran_fast = fast || generic_from >= 0 || big_done;
    if (!fast && !big_done) {
        p.start_pos = generic_from < 0 ? 0 : generic_from;
        if (!h->d_wbuf) {   // the generic engine's window scratch, on first use
            cudaError_t e_ = cudaMalloc((void **)&h->d_wbuf, sizeof(double) * (size_t)h->grid * (h->K_max + 1) * T_SWEEP);
            if (e_ != cudaSuccess) return fail(BGMM_ENOMEM, std::string("cudaMalloc d_wbuf: ") + cudaGetErrorString(e_));
        }
        p.wbuf = h->d_wbuf;
        if (!ran_fast) CU(cudaEventRecord(h->ev2, st));
        if (int rc = h->ops->sweep(h, p)) return rc;
        if (!ran_fast) CU(cudaEventRecord(h->ev3, st));
        h->launches += 1;
    }
    CU(cudaEventRecord(h->ev1, st));
    CU(cudaMemcpyAsync(&c, h->d_ctl, sizeof(c), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    c.evals += x_evals; c.seq_data += x_steps; c.fast_steps += x_steps; c.windows += x_windows; c.wasted += x_wasted;
    float ms = 0.f, ms_k = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    CU(cudaEventElapsedTime(&ms_k, h->ev2, h->ev3));
    h->K = c.K;
    h->sweep_index += 1;
    h->last_gap = (double)h->N / (double)(c.moves + 1);
    if (out) {
        out->K = c.K; out->moves = c.moves; out->births = c.births; out->deaths = c.deaths; out->evals = c.evals;
        out->windows = c.windows; out->seq_data = c.seq_data; out->wasted = c.wasted;
        memcpy(&out->min_margin, &c.margin_bits, 8);
        out->device_ms = ms;
        out->explicit_evals = c.explicit_evals; out->refreshes = c.refreshes; out->generic_from = generic_from;
        for (int t = 0; t < 16; ++t) out->phase_cycles[t] = c.prof[t];
        if (getenv("BGMM_WPROF")) {
            fprintf(stderr, "  cluster step timeline: draw warp wait-weights=%lld scan+draw=%lld post=%lld | comp(cta0,w0) ring=%lld eval=%lld wait=%lld upd=%lld | comp(cta1,w0) ring=%lld eval=%lld wait=%lld upd=%lld | comp(cta0,last) ring=%lld eval=%lld wait=%lld upd=%lld\n",
                    c.tprof[0][0], c.tprof[0][1], c.tprof[0][2], c.tprof[1][0], c.tprof[1][1], c.tprof[1][2], c.tprof[1][3],
                    c.tprof[2][0], c.tprof[2][1], c.tprof[2][2], c.tprof[2][3], c.tprof[3][0], c.tprof[3][1], c.tprof[3][2], c.tprof[3][3]);
            fprintf(stderr, "  cluster step, preparation split (cta1,w0): scalars=%lld quad=%lld log=%lld exp=%lld\n", c.tprof[1][4], c.tprof[1][5],
                    c.tprof[1][6], c.tprof[1][1]);
            static const char *tn[11] = {"stage", "A", "wait1", "finish+scan", "wait2", "draw", "wait3", "move-a", "wait4",
                                         "move-b", "wait5"};
            for (int pt = 0; pt < 4; ++pt) {
                fprintf(stderr, "  seq timeline part %d:", pt);
                for (int t = 0; t < 11; ++t) fprintf(stderr, " %s=%lld", tn[t], c.tprof[pt][t]);
                fprintf(stderr, "\n");
            }
            static const char *nm[16] = {"idle", "load", "chunk", "rowup", "stay", "cand", "full", "pick",
                                         "w-chunk", "w-patchN", "w-patch2", "w-none", "", "", "", ""};
            for (int t = 0; t < 16; ++t)
                if (c.wcnt[t])
                    fprintf(stderr, "  unit %-5s n=%10llu mean=%8.0f max=%8llu cycles\n", nm[t], c.wcnt[t],
                            (double)c.wsum[t] / (double)c.wcnt[t], c.wmax[t]);
        }
        out->launches = h->launches; out->sweep_kernel_ms = ms_k;
        out->guard_hits = c.guard_hits; out->fast_steps = c.fast_steps;
    }
    h->cs_consumed = h->N + c.uextra;
    if (c.error == BGMM_EKMAX) return fail(BGMM_EKMAX, "a new component would exceed K_max (the reference raises IndexError)");
    if (c.error == -8) return fail(BGMM_EINVAL, "constrained sweep: the uniform stream ran out (re-draws need more uniforms)");
    if (c.error != 0) return fail(c.error, "sweep: non-finite weights or covariance not positive definite");
    // the records the auxiliary entry points use were rebuilt after the sweep: report a failure of that rebuild now,
    // not from whichever call happens to look next
    return check_dev_err(h, "sweep (records after the sweep)");
}

// ---------------------------------------------------------------------------------------------
// bgmm_sweep_many: one sweep of n independent chains with ONE kernel launch, one CTA per chain (the register-resident
// sequential engine with the CTA as the chain's only replica).  Burn-in is a serial chain per datum: a single chain
// keeps one SM busy, so the GPU's other SMs run other chains (BASELINE.json configs[3] is 8 chains).
// ---------------------------------------------------------------------------------------------
int bgmm_sweep_many(bgmm_t *const *hs, int32_t n, const int64_t *const *d_orders, const double *const *d_uniforms,
                    double alpha, double power, bgmm_sweep_stats *out) {
    if (!hs || n < 1) return fail(BGMM_EINVAL, "no chains");
    if (!(alpha > 0.0) || !(power > 0.0)) return fail(BGMM_EINVAL, "alpha and power must be > 0");
    bgmm_handle *h0 = hs[0];
    if (!h0) return fail(BGMM_EINVAL, "handle is NULL");
    for (int c = 0; c < n; ++c) {
        bgmm_handle *h = hs[c];
        if (!h) return fail(BGMM_EINVAL, "handle is NULL");
        if (h->device != h0->device || h->stream != h0->stream || h->DP != h0->DP || h->cov != h0->cov ||
            h->K_max != h0->K_max || h->ops != h0->ops)
            return fail(BGMM_EINVAL, "the chains of one bgmm_sweep_many call must share device, stream, D, covariance type and K_max");
        if (!h->fast_ok) return fail(BGMM_EINVAL, "bgmm_sweep_many needs the resident engine (full covariance, D <= 16)");
        for (int d = 0; d < c; ++d) if (hs[d] == h) return fail(BGMM_EINVAL, "a chain appears twice");
    }
    CU(cudaSetDevice(h0->device));
    cudaStream_t st = h0->stream;
    if (h0->pv_cap < n) {
        cudaFree(h0->d_pv);
        h0->d_pv = nullptr;
        h0->pv_cap = 0;
        CU(cudaMalloc((void **)&h0->d_pv, sizeof(Params) * (size_t)n));
        h0->pv_cap = n;
    }
    std::vector<Params> pv(n);
    std::vector<Ctl> ctl(n);
    for (int c = 0; c < n; ++c) CU(cudaMemcpyAsync(&ctl[c], hs[c]->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    const double one = 1.0;
    for (int c = 0; c < n; ++c) {
        bgmm_handle *h = hs[c];
        h->launches = 0;
        const long long *d_order = d_orders ? (const long long *)d_orders[c] : nullptr;
        const double *d_u = d_uniforms ? d_uniforms[c] : nullptr;
        if (!d_u) {
            const int T = 256;
            k_philox<<<(unsigned)((h->N + T - 1) / T), T, 0, st>>>(h->d_u, h->N, h->seed, (unsigned long long)h->sweep_index);
            CU(cudaGetLastError());
            h->launches += 1;
            d_u = h->d_u;
        }
        Ctl &k = ctl[c];
        if (k.K > h->Kcap) return fail(BGMM_EINVAL, "a chain has more live components than the resident engine holds; sweep it with bgmm_sweep");
        k.moves = k.births = k.deaths = k.evals = k.windows = k.seq_data = k.wasted = 0;
        k.explicit_evals = k.refreshes = k.guard_hits = k.fast_steps = 0;
        k.watchdog_ns = h->watchdog_ns;
        memset(k.prof, 0, sizeof(k.prof));
        memcpy(&k.margin_bits, &one, 8);
        k.error = 0; k.bar_count = 0; k.rb_count = 0; k.rb_word = 0; k.pos = 0; k.win = 0; k.first = POS_INF; k.n_dirty = 0;
        CU(cudaMemcpyAsync(h->d_ctl, &k, sizeof(Ctl), cudaMemcpyHostToDevice, st));
        if (d_order) {
            unsigned int *bits = (unsigned int *)h->d_tmp_ll;
            CU(cudaMemsetAsync(bits, 0, sizeof(unsigned int) * (size_t)((h->N + 31) / 32), st));
            const int T = 256;
            k_check_perm<<<(unsigned)((h->N + T - 1) / T), T, 0, st>>>(d_order, h->N, bits, h->d_err);
            CU(cudaGetLastError());
            h->launches += 1;
        }
        Params p = make_params(h);
        p.order = d_order; p.u = d_u;
        p.log_alpha = log(alpha); p.power = power;
        p.init_gap = 0.0; p.engine = 1; p.solo = 1;
        CU(cudaMemcpyAsync(h->d_z2, h->d_z, sizeof(int) * (size_t)h->N, cudaMemcpyDeviceToDevice, st));
        if (int rc = h->ops->fast_prep(h, p, k.K)) return rc;
        pv[c] = p;
    }
    for (int c = 0; c < n; ++c)
        if (int rc = check_dev_err(hs[c], "sweep_many (scan order / record set-up)")) return rc;
    CU(cudaMemcpyAsync(h0->d_pv, pv.data(), sizeof(Params) * (size_t)n, cudaMemcpyHostToDevice, st));
    CU(cudaEventRecord(h0->ev2, st));
    if (int rc = h0->ops->fast_sweep_many(h0, h0->d_pv, n)) return rc;
    CU(cudaEventRecord(h0->ev3, st));
    for (int c = 0; c < n; ++c) CU(cudaMemcpyAsync(&ctl[c], hs[c]->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    float ms_k = 0.f;
    CU(cudaEventElapsedTime(&ms_k, h0->ev2, h0->ev3));
    int first_rc = 0;
    for (int c = 0; c < n; ++c) {
        bgmm_handle *h = hs[c];
        Ctl &k = ctl[c];
        std::swap(h->d_z, h->d_z2);
        Params p = pv[c];
        p.z_uid = h->d_z; p.z_out = h->d_z2;
        long long generic_from = -1;
        if (k.K > 0) {
            if (int rc = h->ops->refactor_all(h, p, 0, k.K)) return rc;
            h->launches += 1;
        }
        if (k.error == fast::E_NEED_GENERIC) {
            // more live components than the resident engine holds: this chain finishes its sweep on the generic engine
            generic_from = k.pos;
            k.error = 0; k.bar_count = 0;
            CU(cudaMemcpyAsync(h->d_ctl, &k, sizeof(Ctl), cudaMemcpyHostToDevice, st));
            if (!h->d_wbuf) {
                cudaError_t e_ = cudaMalloc((void **)&h->d_wbuf, sizeof(double) * (size_t)h->grid * (h->K_max + 1) * T_SWEEP);
                if (e_ != cudaSuccess) return fail(BGMM_ENOMEM, std::string("cudaMalloc d_wbuf: ") + cudaGetErrorString(e_));
            }
            p.wbuf = h->d_wbuf; p.solo = 0; p.engine = 0; p.start_pos = generic_from;
            if (int rc = h->ops->sweep(h, p)) return rc;
            h->launches += 1;
            CU(cudaMemcpyAsync(&k, h->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
        }
        h->K = k.K;
        h->sweep_index += 1;
        h->last_gap = (double)h->N / (double)(k.moves + 1);
        if (out) {
            bgmm_sweep_stats &o = out[c];
            memset(&o, 0, sizeof(o));
            o.K = k.K; o.moves = k.moves; o.births = k.births; o.deaths = k.deaths; o.evals = k.evals;
            o.windows = k.windows; o.seq_data = k.seq_data; o.wasted = k.wasted;
            memcpy(&o.min_margin, &k.margin_bits, 8);
            o.device_ms = ms_k; o.sweep_kernel_ms = ms_k;
            o.explicit_evals = k.explicit_evals; o.refreshes = k.refreshes; o.generic_from = generic_from;
            for (int t = 0; t < 16; ++t) o.phase_cycles[t] = k.prof[t];
            o.launches = h->launches + (c == 0 ? 1 : 0);
            o.guard_hits = k.guard_hits; o.fast_steps = k.fast_steps;
        }
        if (k.error != 0 && first_rc == 0)
            first_rc = fail(k.error == BGMM_EKMAX ? BGMM_EKMAX : k.error,
                            "sweep_many: chain " + std::to_string(c) + (k.error == BGMM_EKMAX
                                ? ": a new component would exceed K_max" : ": non-finite weights or covariance not positive definite"));
    }
    if (first_rc) return first_rc;
    for (int c = 0; c < n; ++c)
        if (int rc = check_dev_err(hs[c], "sweep_many (records after the sweep)")) return rc;
    return 0;
}

int bgmm_sweep_dev(bgmm_t *h, const int64_t *d_order, const double *d_uniforms, double alpha, double power,
                   bgmm_sweep_stats *out) {
    if (!h) return fail(BGMM_EINVAL, "handle is NULL");
    CU(cudaSetDevice(h->device));
    return run_sweep(h, (const long long *)d_order, d_uniforms, alpha, power, out);
}

int bgmm_sweep(bgmm_t *h, const int64_t *order, const double *uniforms, double alpha, double power,
               bgmm_sweep_stats *out) {
    if (!h) return fail(BGMM_EINVAL, "handle is NULL");
    CU(cudaSetDevice(h->device));
    const long long *d_order = nullptr;
    const double *d_u = nullptr;
    if (order) {
        CU(cudaMemcpyAsync(h->d_order, order, sizeof(long long) * (size_t)h->N, cudaMemcpyHostToDevice, h->stream));
        d_order = h->d_order;
    }
    if (uniforms) {
        CU(cudaMemcpyAsync(h->d_u, uniforms, sizeof(double) * (size_t)h->N, cudaMemcpyHostToDevice, h->stream));
        d_u = h->d_u;
    }
    return run_sweep(h, d_order, d_u, alpha, power, out);
}

int bgmm_sweep_constrained(bgmm_t *h, const int64_t *order, const double *uniforms, int64_t n_uniforms, double alpha,
                           double power, const int32_t *status, int32_t n_status, int64_t *consumed, bgmm_sweep_stats *out) {
    if (!h || !uniforms || !status) return fail(BGMM_EINVAL, "NULL argument");
    if (n_uniforms < h->N || n_status < 1) return fail(BGMM_EINVAL, "at least N uniforms and one status entry are needed");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    if (h->ubig_cap < n_uniforms) {
        cudaFree(h->d_ubig);
        h->d_ubig = nullptr;
        h->ubig_cap = 0;
        CU(cudaMalloc((void **)&h->d_ubig, sizeof(double) * (size_t)n_uniforms));
        h->ubig_cap = n_uniforms;
    }
    if (!h->d_status) CU(cudaMalloc((void **)&h->d_status, sizeof(int) * (size_t)(h->K_max + 1)));
    const int ns = std::min<int>(n_status, h->K_max + 1);
    CU(cudaMemcpyAsync(h->d_ubig, uniforms, sizeof(double) * (size_t)n_uniforms, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->d_status, status, sizeof(int) * (size_t)ns, cudaMemcpyHostToDevice, st));
    const long long *d_order = nullptr;
    if (order) {
        CU(cudaMemcpyAsync(h->d_order, order, sizeof(long long) * (size_t)h->N, cudaMemcpyHostToDevice, st));
        d_order = h->d_order;
    }
    h->cs_status = h->d_status; h->cs_n = ns; h->cs_ulen = n_uniforms;
    const int rc = run_sweep(h, d_order, h->d_ubig, alpha, power, out);
    h->cs_status = nullptr; h->cs_n = 0; h->cs_ulen = 0;
    if (consumed) *consumed = h->cs_consumed;
    return rc;
}

int bgmm_K(bgmm_t *h) { return h ? h->K : -1; }

int bgmm_get_assignments_dev(bgmm_t *h, int64_t *d_out) {
    if (!h || !d_out) return fail(BGMM_EINVAL, "NULL argument");
    CU(cudaSetDevice(h->device));
    const int T = 256;
    k_relabel<<<(unsigned)((h->N + T - 1) / T), T, 0, h->stream>>>(h->d_z, h->d_slot_of_uid, h->N, (long long *)d_out);
    CU(cudaGetLastError());
    return 0;
}

int bgmm_get_state(bgmm_t *h, int64_t *z, int64_t *counts, int32_t *K, double *m_num, double *S_part, double *logdet,
                   double *inv_covar) {
    if (!h) return fail(BGMM_EINVAL, "handle is NULL");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int D = h->D, DP = h->DP, K_max = h->K_max, Kl = h->K;
    const int SS = stat_len(DP, h->cov);
    const size_t ssr = (h->cov == BGMM_COV_FULL) ? (size_t)D * D : (size_t)D;
    if (K) *K = Kl;
    if (z) {
        if (int rc = bgmm_get_assignments_dev(h, (int64_t *)h->d_tmp_ll)) return rc;
        CU(cudaMemcpyAsync(z, h->d_tmp_ll, sizeof(long long) * (size_t)h->N, cudaMemcpyDeviceToHost, st));
    }
    if (counts) {
        CU(cudaMemcpyAsync(counts, h->d_counts, sizeof(long long) * K_max, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        for (int k = Kl; k < K_max; ++k) counts[k] = 0;
    }
    if (m_num) {
        std::vector<double> t((size_t)std::max(Kl, 1) * DP);
        CU(cudaMemcpyAsync(t.data(), h->d_num, sizeof(double) * (size_t)Kl * DP, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        memset(m_num, 0, sizeof(double) * (size_t)K_max * D);
        for (int k = 0; k < Kl; ++k) for (int a = 0; a < D; ++a) m_num[(size_t)k * D + a] = t[(size_t)k * DP + a];
    }
    if (S_part) {
        std::vector<double> t((size_t)std::max(Kl, 1) * SS);
        CU(cudaMemcpyAsync(t.data(), h->d_S, sizeof(double) * (size_t)Kl * SS, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        memset(S_part, 0, sizeof(double) * (size_t)K_max * ssr);
        for (int k = 0; k < Kl; ++k) {
            if (h->cov == BGMM_COV_FULL) {
                for (int a = 0; a < D; ++a)
                    for (int b = 0; b <= a; ++b) {
                        const double v = t[(size_t)k * SS + row_idx(a, b)];
                        S_part[(size_t)k * ssr + a * D + b] = v;
                        S_part[(size_t)k * ssr + b * D + a] = v;
                    }
            } else {
                for (int a = 0; a < D; ++a) S_part[(size_t)k * ssr + a] = t[(size_t)k * SS + a];
            }
        }
    }
    if (logdet || inv_covar) {
        double *d_ld = nullptr, *d_inv = nullptr;
        std::vector<double> ld(std::max(Kl, 1)), iv((size_t)std::max(Kl, 1) * ssr);
        if (Kl > 0) {
            CU(cudaMalloc((void **)&d_ld, sizeof(double) * Kl));
            CU(cudaMalloc((void **)&d_inv, sizeof(double) * (size_t)Kl * ssr));
            Params p = make_params(h);
            if (h->cov == BGMM_COV_FULL) k_inv_covar<COV_FULL><<<Kl, 32, sizeof(double) * D * D, st>>>(p, DP, d_ld, d_inv);
            else if (h->cov == BGMM_COV_DIAG) k_inv_covar<COV_DIAG><<<Kl, 32, 0, st>>>(p, DP, d_ld, d_inv);
            else k_inv_covar<COV_FIXED><<<Kl, 32, 0, st>>>(p, DP, d_ld, d_inv);
            CU(cudaGetLastError());
            CU(cudaMemcpyAsync(ld.data(), d_ld, sizeof(double) * Kl, cudaMemcpyDeviceToHost, st));
            CU(cudaMemcpyAsync(iv.data(), d_inv, sizeof(double) * (size_t)Kl * ssr, cudaMemcpyDeviceToHost, st));
            CU(cudaStreamSynchronize(st));
            cudaFree(d_ld);
            cudaFree(d_inv);
        }
        if (logdet) {
            memset(logdet, 0, sizeof(double) * K_max);
            for (int k = 0; k < Kl; ++k) logdet[k] = ld[k];
        }
        if (inv_covar) {
            memset(inv_covar, 0, sizeof(double) * (size_t)K_max * ssr);
            memcpy(inv_covar, iv.data(), sizeof(double) * (size_t)Kl * ssr);
        }
    }
    CU(cudaStreamSynchronize(st));
    return 0;
}

int bgmm_log_prior(bgmm_t *h, double *out) {
    if (!h || !out) return fail(BGMM_EINVAL, "NULL argument");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(out, h->d_log_prior, sizeof(double) * (size_t)h->N, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}

int bgmm_log_post_pred(bgmm_t *h, const int64_t *idx, int64_t n, double *out) {
    if (!h || (n > 0 && (!idx || !out))) return fail(BGMM_EINVAL, "NULL argument");
    if (n <= 0 || h->K == 0) return 0;
    for (int64_t t = 0; t < n; ++t)
        if (idx[t] < 0 || idx[t] >= h->N) return fail(BGMM_EINVAL, "datum index out of range");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    long long *d_idx = nullptr;
    double *d_out = nullptr;
    CU(cudaMalloc((void **)&d_idx, sizeof(long long) * (size_t)n));
    CU(cudaMalloc((void **)&d_out, sizeof(double) * (size_t)n * h->K));
    CU(cudaMemcpyAsync(d_idx, idx, sizeof(long long) * (size_t)n, cudaMemcpyHostToDevice, st));
    Params p = make_params(h);
    int rc = h->ops->lpp(h, p, d_idx, n, h->K, d_out);
    if (rc == 0) {
        cudaMemcpyAsync(out, d_out, sizeof(double) * (size_t)n * h->K, cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
    }
    cudaFree(d_idx);
    cudaFree(d_out);
    return rc;
}

int bgmm_log_marg_k(bgmm_t *h, double *out) {
    if (!h || !out) return fail(BGMM_EINVAL, "NULL argument");
    if (h->K == 0) return 0;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    double *d_out = nullptr;
    CU(cudaMalloc((void **)&d_out, sizeof(double) * h->K));
    Params p = make_params(h);
    if (h->cov == BGMM_COV_FULL)
        k_log_marg_k<COV_FULL><<<h->K, 32, sizeof(double) * (packed_len(h->D) + 2), st>>>(p, h->DP, h->logdet_S0, d_out);
    else if (h->cov == BGMM_COV_DIAG)
        k_log_marg_k<COV_DIAG><<<h->K, 32, 16, st>>>(p, h->DP, h->logdet_S0, d_out);
    else
        k_fixed_log_marg_k<<<h->K, 256, 0, st>>>(p, h->DP, d_out);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_out, sizeof(double) * h->K, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(d_out);
    return 0;
}

int bgmm_log_marg(bgmm_t *h, double alpha, double *out) {
    if (!h || !out) return fail(BGMM_EINVAL, "NULL argument");
    const int K = h->K;
    std::vector<double> lk(std::max(K, 1));
    std::vector<long long> cnt(h->K_max);
    if (int rc = bgmm_log_marg_k(h, lk.data())) return rc;
    CU(cudaMemcpy(cnt.data(), h->d_counts, sizeof(long long) * h->K_max, cudaMemcpyDeviceToHost));
    // igmm.py:199-215
    double facts = 0.0;
    long long tot = 0;
    for (int k = 0; k < K; ++k) {
        if (cnt[k] != 0) facts += lgamma((double)cnt[k]);
        tot += cnt[k];
    }
    const double lpz = (K - 1) * log(alpha) + lgamma(alpha) - lgamma((double)tot + alpha) + facts;
    double lpx = 0.0;
    for (int k = 0; k < K; ++k) lpx += lk[k];
    *out = lpz + lpx;
    return 0;
}

int bgmm_set_true_labels(bgmm_t *h, const int64_t *t, int32_t T) {
    if (!h || !t) return fail(BGMM_EINVAL, "NULL argument");
    if (T < 1) return fail(BGMM_EINVAL, "T must be >= 1");
    std::vector<int> t32((size_t)h->N);
    for (long long i = 0; i < h->N; ++i) {
        if (t[i] < 0 || t[i] >= T) return fail(BGMM_EINVAL, "true label out of range [0, T)");
        t32[(size_t)i] = (int)t[i];
    }
    CU(cudaSetDevice(h->device));
    if (!h->d_true) CU(cudaMalloc((void **)&h->d_true, sizeof(int) * (size_t)(h->N + 4)));
    CU(cudaMemcpyAsync(h->d_true, t32.data(), sizeof(int) * (size_t)h->N, cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->T_true = T;
    return 0;
}

int bgmm_contingency(bgmm_t *h, int64_t *table) {
    if (!h || !table) return fail(BGMM_EINVAL, "NULL argument");
    if (!h->d_true) return fail(BGMM_EINVAL, "bgmm_set_true_labels has not been called");
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int T = h->T_true, K = h->K;
    const size_t cells = (size_t)T * (K + 1);
    if (cells > h->table_cells) {
        cudaFree(h->d_table);
        h->d_table = nullptr;
        h->table_cells = 0;
        CU(cudaMalloc((void **)&h->d_table, sizeof(unsigned long long) * cells));
        h->table_cells = cells;
    }
    CU(cudaMemsetAsync(h->d_table, 0, sizeof(unsigned long long) * cells, st));
    const int TB = 256;
    const size_t smem = sizeof(unsigned int) * cells;
    const int in_smem = smem <= (size_t)200 * 1024;
    if (in_smem && smem > 48 * 1024)
        CU(cudaFuncSetAttribute(k_contingency, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // a multiple of the SM count; fewer CTAs when the table is large (every CTA clears and folds its own copy)
    const long long want = ((h->N >> 2) + TB) / TB;
    const int per_sm = smem <= 32 * 1024 ? 4 : 1;
    const int grid = (int)std::max(1LL, std::min(want, (long long)h->num_sms * per_sm));
    k_contingency<<<grid, TB, in_smem ? smem : 0, st>>>(h->d_true, h->d_z, h->d_slot_of_uid, h->N, T, K, in_smem,
                                                          h->d_table);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(table, h->d_table, sizeof(unsigned long long) * cells, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return 0;
}

int bgmm_cluster_ssq(bgmm_t *h, double *out) {
    if (!h || !out) return fail(BGMM_EINVAL, "NULL argument");
    if (h->cov == BGMM_COV_FIXED)
        return fail(BGMM_EINVAL, "the fixed-variance statistics do not hold sum x^2: count the loss from the labels");
    if (h->K == 0) return 0;
    CU(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    double *d_out = nullptr;
    CU(cudaMalloc((void **)&d_out, sizeof(double) * h->K));
    Params p = make_params(h);
    if (h->cov == BGMM_COV_FULL) k_cluster_ssq<COV_FULL><<<h->K, 32, 0, st>>>(p, h->DP, d_out);
    else k_cluster_ssq<COV_DIAG><<<h->K, 32, 0, st>>>(p, h->DP, d_out);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(out, d_out, sizeof(double) * h->K, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(d_out);
    return 0;
}

static int item_op(bgmm_handle *h, int op, long long i, int k) {
    if (!h) return fail(BGMM_EINVAL, "handle is NULL");
    if (i < 0 || i >= h->N) return fail(BGMM_EINVAL, "datum index out of range");
    CU(cudaSetDevice(h->device));
    Params p = make_params(h);
    if (int rc = h->ops->item_op(h, p, op, i, k)) return rc;
    Ctl c;
    CU(cudaMemcpyAsync(&c, h->d_ctl, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->K = c.K;
    if (c.error) {
        const int e = c.error;
        c.error = 0;
        cudaMemcpy(h->d_ctl, &c, sizeof(c), cudaMemcpyHostToDevice);
        return fail(e, e == BGMM_EKMAX ? "K_max overflow" : (e == BGMM_EINVAL ? "invalid component index" : "numeric failure"));
    }
    return 0;
}
int bgmm_set_component_stats(bgmm_t *h, int32_t k, const double *m_num, const double *S_part, int64_t count) {
    if (!h || !m_num || !S_part) return fail(BGMM_EINVAL, "NULL argument");
    if (k < 0 || k >= h->K) return fail(BGMM_EINVAL, "component index out of range");
    if (count < 1) return fail(BGMM_EINVAL, "count must be >= 1");
    CU(cudaSetDevice(h->device));
    const int D = h->D, DP = h->DP, SS = stat_len(DP, h->cov);
    std::vector<double> num(DP, 0.0), S(SS, 0.0);
    for (int a = 0; a < D; ++a) num[a] = m_num[a];
    if (h->cov == BGMM_COV_FULL) {
        for (int a = 0; a < D; ++a) for (int b = 0; b <= a; ++b) S[row_idx(a, b)] = S_part[a * D + b];
    } else {
        for (int a = 0; a < D; ++a) S[a] = S_part[a];
    }
    const long long c = count;
    cudaStream_t st = h->stream;
    CU(cudaMemcpyAsync(h->d_num + (size_t)k * DP, num.data(), sizeof(double) * DP, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->d_S + (size_t)k * SS, S.data(), sizeof(double) * SS, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(h->d_counts + k, &c, sizeof(long long), cudaMemcpyHostToDevice, st));
    Params p = make_params(h);
    if (int rc = h->ops->refactor_all(h, p, k, 1)) return rc;
    CU(cudaStreamSynchronize(st));
    return check_dev_err(h, "set_component_stats");
}
int bgmm_set_label(bgmm_t *h, int64_t i, int32_t k) {
    if (!h) return fail(BGMM_EINVAL, "handle is NULL");
    if (i < 0 || i >= h->N) return fail(BGMM_EINVAL, "datum index out of range");
    if (k < -1 || k >= h->K) return fail(BGMM_EINVAL, "component index out of range");
    CU(cudaSetDevice(h->device));
    int uid = -1;
    if (k >= 0) CU(cudaMemcpyAsync(&uid, h->d_uid_of_slot + k, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpyAsync(h->d_z + i, &uid, sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return 0;
}
int bgmm_set_state(bgmm_t *h, const int64_t *z, int32_t K, const double *m_num, const double *S_part) {
    if (!h || !z) return fail(BGMM_EINVAL, "NULL argument");
    if (int rc = bgmm_set_assignments(h, z)) return rc;
    if (K != h->K) return fail(BGMM_EINVAL, "K does not match the labels");
    if (!m_num && !S_part) return 0;
    if (!m_num || !S_part) return fail(BGMM_EINVAL, "m_num and S_part must be given together");
    std::vector<long long> cnt(h->K_max);
    CU(cudaMemcpy(cnt.data(), h->d_counts, sizeof(long long) * h->K_max, cudaMemcpyDeviceToHost));
    const size_t ssr = (h->cov == BGMM_COV_FULL) ? (size_t)h->D * h->D : (size_t)h->D;
    for (int k = 0; k < K; ++k)
        if (int rc = bgmm_set_component_stats(h, k, m_num + (size_t)k * h->D, S_part + (size_t)k * ssr, cnt[k])) return rc;
    return 0;
}
int bgmm_set_guard(bgmm_t *h, double guard) {
    if (!h || !(guard >= 0.0) || !(guard < 1.0)) return fail(BGMM_EINVAL, "guard must be in [0, 1)");
    h->guard = guard;
    return 0;
}
int bgmm_add_item(bgmm_t *h, int64_t i, int32_t k) { return item_op(h, 1, i, k); }
int bgmm_del_item(bgmm_t *h, int64_t i) { return item_op(h, 0, i, 0); }

}  // extern "C"
