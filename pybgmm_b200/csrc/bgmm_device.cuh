// bgmm_device.cuh -- device-side building blocks of the B200 collapsed-Gibbs engine.
//
// Nothing here is a translation of the reference (which is pure Python/NumPy); the comments cite the
// reference file:line whose *semantics* each block reproduces (paths relative to the reference root).
//
// Representation (see DESIGN.md "Data layout in HBM"):
//   sufficient statistics  (bit-identical to the reference, same operation order, no FMA contraction)
//       num[k][DP]                 m_N_numerators            gaussian_components.py:86
//       S[k][DP(DP+1)/2]           lower triangle of S_N_partials (symmetric by construction)  :87
//       counts[k]                                                                              :90
//   evaluation record rec[k][R]    (derived, recomputed from the statistics after every change)
//       full: Cholesky factor L of the predictive covariance (column-major packed, reciprocal diagonal),
//             the mean m_N, and the per-component constants of the Student-t log pdf  (:228-251)
//       diag: m_N, inv_vars and the constants                                (_diag.py:237-259)
//   labels are stored as component uids; slot_of_uid[] resolves them to the reference's positional
//   labels, which makes the O(N) relabel of del_component (:199) an O(1) table update.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include <stdio.h>

namespace bgmm {

constexpr int COV_FULL = 0;
constexpr int COV_DIAG = 1;
constexpr int COV_FIXED = 2;        // known diagonal variance, normal prior on the means (gaussian_components_fixedvar.py)
constexpr int T_SWEEP = 256;        // threads per CTA of the sweep kernel
constexpr int MAX_DIRTY = 8;
constexpr long long POS_INF = 0x7fffffffffffffffLL;
constexpr double OM_MIN = 1.0 / 64.0;   // below this the closed-form removal is not trusted -> explicit path
constexpr double EXP_CUTOFF = -745.2;   // exp(x) == 0.0 in fp64 for x below this

// scalar slots at the tail of a record
enum { SC_C = 0, SC_H = 1, SC_INVNU = 2, SC_LC = 3, SC_N = 4, SC_LOGDET = 5, SC_F = 6, SC_SPARE = 7, SC_COUNT = 8 };

__host__ __device__ constexpr int packed_len(int dp) { return dp * (dp + 1) / 2; }
__host__ __device__ constexpr int rec_mu_off(int dp, int cov) { return cov == COV_FULL ? packed_len(dp) : 0; }
__host__ __device__ constexpr int rec_iv_off(int dp) { return dp; }  // diag only
// fixed variance: [mean DP][predictive precision DP][precision-weighted sum DP][posterior precision DP][scalars]
__host__ __device__ constexpr int rec_sc_off(int dp, int cov) {
    return cov == COV_FULL ? packed_len(dp) + dp : (cov == COV_FIXED ? 4 * dp : 2 * dp);
}
__host__ __device__ constexpr int rec_len(int dp, int cov) { return (rec_sc_off(dp, cov) + SC_COUNT + 1) & ~1; }
__host__ __device__ constexpr int stat_len(int dp, int cov) { return cov == COV_FULL ? packed_len(dp) : dp; }
// column-major packed lower triangle (record): element (a,b), a >= b
__host__ __device__ constexpr int col_off(int dp, int b) { return b * dp - b * (b - 1) / 2; }
// row-major packed lower triangle (statistics / scratch): element (a,b), a >= b
__host__ __device__ __forceinline__ int row_idx(int a, int b) { return a * (a + 1) / 2 + b; }

// ---------------------------------------------------------------------------------------------
// control block (device global memory, one per handle)
// ---------------------------------------------------------------------------------------------
struct Ctl {
    // the barrier words and each round slot sit on their own 128-byte lines: they are hammered by every CTA
    unsigned int bar_count;
    unsigned int pad_a[31];
    unsigned int bar_gen;
    unsigned int pad_b[31];
    unsigned int rb_count;             // round barrier of the fast engine: arrivals ...
    unsigned int pad_c[31];
    unsigned long long rb_word;        // ... and (round tag << 44 | first candidate), released by the last arriver
    unsigned long long pad_d[15];
    unsigned long long first3[3][16];  // per-round atomicMin targets of the fast engine (round r uses [r % 3][0])
    int K;
    int error;          // 0, or a BGMM_E* code
    long long pos;      // next scan position
    long long win;      // length of the speculative window of this iteration; 0 = sequential run
    long long first;    // atomicMin target: first scan position in the window that does not provably stay
    int n_dirty;
    int dirty[MAX_DIRTY];
    unsigned int full_gen;  // bumped when every record must be re-staged
    int n_free;             // free uid stack size
    int pad0;
    // counters
    long long moves, births, deaths, evals, windows, seq_data, wasted;
    unsigned long long margin_bits;
    double gap;         // running estimate of the number of data between two movers
    long long explicit_evals, refreshes;
    long long guard_hits;          // draws whose margin was below Params::guard and were redone on the exact path
    long long fast_steps;          // data resolved by the register-resident sequential step (bgmm_seq.cuh)
    long long uextra;              // constrained sweep: uniforms consumed by re-draws so far (bgmm_sweep_constrained)
    long long watchdog_ns;         // spin loops give up (error word + trap) after this many ns of the global timer
    long long prof[16];            // phase clocks of CTA 0 (cycles), see bgmm_fast.cuh
    unsigned long long wsum[16], wcnt[16], wmax[16];  // evaluator unit clocks by category (profile builds)
    long long tprof[4][16];        // register step: cycles per (part, phase) of warps 0 / 4 / 8 / 12 (profile builds)
};

struct Params {
    // immutable during a sweep
    const double *X;          // N x DP (rows zero padded to DP)
    const double *log_prior;  // N
    const double *lgam;       // table gammaln(n/2)      gaussian_components.py:122
    const double *logv;       // table log(n)            :121
    const long long *order;   // N or null
    const double *u;          // N uniforms in scan order
    // mutable state
    int *z_uid;               // N
    int *z_out;               // N: labels written by the fast engine (z_uid stays the sweep's read-only input)
    int *slot_of_uid;         // K_max
    int *uid_of_slot;         // K_max
    int *uid_free;            // K_max (stack)
    long long *counts;        // K_max
    double *num;              // K_max x DP
    double *S;                // K_max x SS
    double *rec;              // K_max x R
    double *wbuf;             // grid x (K_max+1) x T scratch
    Ctl *ctl;
    // prior (pybgmm/prior/niw.py:10-23)
    const double *m0;         // DP   (fixed variance: mu_0)
    const double *S0;         // SS   (fixed variance: the prior precision of the means, 1 / var_0)
    const double *tau;        // fixed variance only: the data precision 1 / var (DP)
    double k0;
    long long v0;
    // sizes
    long long N;
    int D, K_max, Kc;         // Kc: records resident in shared memory
    double log_alpha, power;  // power == 1 -> log(n) (crpmm.py:70) else log(pow(n, power)) (pcrpmm.py:107)
    double log_pi;
    int engine;               // 0 adaptive, 1 sequential, 2 windows
    double init_gap;
    long long start_pos;      // scan position the sweep (re)starts at
    // fast engine (bgmm_fast.cuh): element-major records rec[e * KS + k] and the prior's record
    double *recB;
    double *recB_prior;
    double *mvbuf;            // per evaluator warp: the inputs of the candidate it published
    const double *fmtab;      // log / exp tables of bgmm_fastmath.cuh
    double *ntab;             // count table: 8 doubles per count n = 0..N (bgmm_fast.cuh NT_*)
    int KS, Kcap;
    float win_factor;         // window length = win_factor x running gap between movers
    int near_zone;            // waiting rows within this many windows beyond the current one are kept current (env BGMM_NEAR)
    double guard;             // a draw whose margin (probability units) is below this is redone exactly (0: never)
    int *err;                 // device error word of the handle (set by set-up kernels: BGMM_E* code)
    float gap_to_win, gap_to_seq;  // mode switches of the resident engine: windows from this gap between movers up,
                              // sequential steps from this gap down (env BGMM_GAP_WIN / BGMM_GAP_SEQ)
    // constrained re-draw (CSCRPMM, cscrpmm.py:342-350): status[k] = 1 useful / 2 non-useful slot at the start of the sweep
    // (0: neither); uniforms are then consumed in order, u_len of them are available.  NULL: ordinary sweep.
    const int *status;
    int n_status;
    long long u_len;
    int solo;                 // 1: this CTA is the chain's only replica (bgmm_sweep_many: one chain per CTA) -- grid
                              // barriers become CTA barriers, the sequential engine only
    int writer;               // set per CTA by the kernel: this CTA writes the chain's global state (CTA 0, or solo)
    int tune;                 // developer switches (env BGMM_TUNE): bit 0 = f_step evaluates with one thread per component,
                              // bit 1 = statistics by load / add / store instead of L2 reductions,
                              // bit 2 = rows about to enter the window are not kept current
};

// ---------------------------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------------------------
template <bool CG> __device__ __forceinline__ double ldr(const double *p) {
    if (CG) return __ldcg(p);
    return *p;
}
__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Watchdog of the spin loops: wall-clock based (%globaltimer, ns), configurable (Ctl::watchdog_ns, env
// BGMM_WATCHDOG_S on the host side), and independent of NDEBUG: a replica disagreement becomes an error word plus a
// trap, never a hang.  The timer is polled once per 1024 spins.
constexpr int E_WATCHDOG = -6;
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
static __device__ __noinline__ void watchdog_fail(Ctl *c, int which) {
    c->error = E_WATCHDOG;
    __threadfence_system();
    printf("bgmm watchdog %d: CTA %d waited longer than %lld ns (replicas stopped agreeing?)\n", which, (int)blockIdx.x,
           c->watchdog_ns);
    __trap();
}
struct SpinWatch {
    unsigned int spins = 0;
    unsigned long long t0 = 0;
    __device__ __forceinline__ void poll(Ctl *c, int which) {
        if ((++spins & 1023u) == 0u) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            else if ((long long)(now - t0) > c->watchdog_ns) watchdog_fail(c, which);
        }
    }
};

// Grid-wide barrier for a cooperative launch (all CTAs co-resident).  Sense-reversing on a generation word.
__device__ __forceinline__ void grid_barrier(Ctl *c) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int gen = ld_acquire_u32(&c->bar_gen);
        __threadfence();
        unsigned int prev = atomicAdd(&c->bar_count, 1u);
        if (prev == gridDim.x - 1) {
            c->bar_count = 0;
            __threadfence();
            st_release_u32(&c->bar_gen, gen + 1);
        } else {
            SpinWatch wd;
            while (ld_acquire_u32(&c->bar_gen) == gen) { __nanosleep(32); wd.poll(c, 0); }
        }
        __threadfence();
    }
    __syncthreads();
}

// m / s for 0 <= m <= s, s > 0 finite, to single-precision accuracy and without a double division (whose slow path
// is a call): both are scaled by the power of two that brings s into [1, 2) first, so sums beyond the float range
// (unnormalised probabilities exp(weight - reference) reach e^200 at D = 64) neither overflow nor flush the quotient.
// This is the margin of a draw in probability units -- a diagnostic and the margin guard's input.
__device__ __forceinline__ double margin_ratio(double m, double s) {
    const long long eb = (__double_as_longlong(s) >> 52) & 0x7ff;               // biased exponent of s
    const double sc = __longlong_as_double((2046LL - eb) << 52);                // 2^(1023 - eb): s * sc in [1, 2)
    return (double)__fdividef((float)(m * sc), (float)(s * sc));
}

// log count prior: crpmm.py:70 np.log(counts) / pcrpmm.py:107-108 np.log(np.power(counts, n_power))
__device__ __forceinline__ double log_count(double n, double power) {
    if (power == 1.0) return log(n);
    return log(pow(n, power));
}

// ---------------------------------------------------------------------------------------------
// Mahalanobis term through the Cholesky factor: q = |L^-1 (m - x)|^2      (gaussian_components.py:241-246)
// The reference contracts with the explicit inverse; forward substitution with L is the same quantity,
// backward stable, and needs no cancellation-prone expansion.  rec may live in shared or global memory.
// ---------------------------------------------------------------------------------------------
template <int DP, bool CG>
__device__ __forceinline__ double quad_full(const double *__restrict__ rec, const double (&x)[DP]) {
    constexpr int PP = packed_len(DP);
    double d[DP];
#pragma unroll
    for (int a = 0; a < DP; ++a) d[a] = ldr<CG>(rec + PP + a) - x[a];
    double q = 0.0;
#pragma unroll
    for (int b = 0; b < DP; ++b) {
        const int off = col_off(DP, b);
        const double y = d[b] * ldr<CG>(rec + off);  // reciprocal diagonal
        q = fma(y, y, q);
#pragma unroll
        for (int a = b + 1; a < DP; ++a) d[a] = fma(-ldr<CG>(rec + off + (a - b)), y, d[a]);
    }
    return q;
}

// diag: sum_d log(1 + delta_d^2 * inv_var_d * (1/v))                      (_diag.py:255-257)
template <int DP, bool CG>
__device__ __forceinline__ double logsum_diag(const double *__restrict__ rec, const double (&x)[DP], int D,
                                              double inv_nu) {
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < DP; ++a) {
        if (a < D) {
            const double dl = ldr<CG>(rec + a) - x[a];
            s += log(1.0 + (dl * dl) * ldr<CG>(rec + DP + a) * inv_nu);
        }
    }
    return s;
}

// Student-t log pdf constants for integer degrees of freedom v      (gaussian_components.py:237-249, :340-343)
__device__ __forceinline__ double t_const_full(const double *lgam, const double *logv, long long v, int D,
                                               double log_pi, double logdet) {
    return __ldg(lgam + v + D) - __ldg(lgam + v) - D / 2. * __ldg(logv + v) - D / 2. * log_pi - 0.5 * logdet;
}
// (_diag.py:247-254, :347-351)
__device__ __forceinline__ double t_const_diag(const double *lgam, const double *logv, long long v, int D,
                                               double log_pi, double log_prod_var) {
    return D * (__ldg(lgam + v + 1) - __ldg(lgam + v) - 0.5 * __ldg(logv + v) - 0.5 * log_pi) - 0.5 * log_prod_var;
}

// ---------------------------------------------------------------------------------------------
// Weight of component k for a datum NOT currently assigned to it:  log prior count + log_post_pred
// (crpmm.py:70-72).  Returns w; q/logsum evaluation included.
// ---------------------------------------------------------------------------------------------
template <int DP, int COV, bool CG>
__device__ __forceinline__ double lpp_other(const double *__restrict__ rec, const double (&x)[DP], int D) {
    constexpr int SO = rec_sc_off(DP, COV);
    const double c = ldr<CG>(rec + SO + SC_C), h = ldr<CG>(rec + SO + SC_H);
    const double inv_nu = ldr<CG>(rec + SO + SC_INVNU);
    if (COV == COV_FULL) {
        const double q = quad_full<DP, CG>(rec, x);
        return c - h * log(1.0 + inv_nu * q);
    } else if (COV == COV_FIXED) {
        // product of normals (gaussian_components_fixedvar.py:221-232): c = -D/2 log 2 pi + sum log(pred) / 2
        double s = 0.0;
#pragma unroll
        for (int a = 0; a < DP; ++a) {
            if (a < D) {
                const double dl = ldr<CG>(rec + a) - x[a];
                s += (dl * dl) * ldr<CG>(rec + DP + a);
            }
        }
        return c - 0.5 * s;
    } else {
        const double s = logsum_diag<DP, CG>(rec, x, D, inv_nu);
        return c - h * s;
    }
}
template <int DP, int COV, bool CG>
__device__ __forceinline__ double weight_other(const double *__restrict__ rec, const double (&x)[DP], int D) {
    constexpr int SO = rec_sc_off(DP, COV);
    return ldr<CG>(rec + SO + SC_LC) + lpp_other<DP, COV, CG>(rec, x, D);
}

// ---------------------------------------------------------------------------------------------
// Weight of the datum's own component with the datum removed (del_item, gaussian_components.py:171-186,
// followed by log_post_pred on the reduced component) evaluated in closed form from the CURRENT record:
//   S_N' = S_N - beta d d^T,  beta = kappa/(kappa-1), d = x - m_N            (rank-one downdate)
//   |S_N'| = |S_N| (1 - beta s),  d'^T S_N'^-1 d' = (kappa/kappa')^2 s / (1 - beta s),  s = d^T S_N^-1 d
// so a datum that stays costs no refactorisation and leaves the statistics bit-for-bit untouched
// (the reference's restore path, crpmm.py:82-85).  *ok is cleared when 1 - beta s is too small to trust;
// the caller then takes the explicit path.
// ---------------------------------------------------------------------------------------------
template <int DP, int COV, bool CG>
__device__ __forceinline__ double weight_own_removed(const double *__restrict__ rec, const double (&x)[DP],
                                                     const Params &p, bool *ok) {
    constexpr int SO = rec_sc_off(DP, COV);
    const int D = p.D;
    if (COV == COV_FIXED) {
        // the reference's own del_item arithmetic on the component's statistics (gaussian_components_fixedvar.py:
        // 164-180, :278-286), evaluated on the fly: nothing to factor, so no closed form is needed
        const double n1 = ldr<CG>(rec + SO + SC_N) - 1.0;
        double lp = 0.0, s = 0.0;
#pragma unroll
        for (int a = 0; a < DP; ++a) {
            if (a < D) {
                const double ta = p.tau[a];
                const double num1 = __dsub_rn(ldr<CG>(rec + 2 * DP + a), __dmul_rn(ta, x[a]));
                const double t1 = __dsub_rn(ldr<CG>(rec + 3 * DP + a), ta);
                const double pred = t1 * ta / (t1 + ta);
                const double dl = num1 / t1 - x[a];
                lp += log(pred);
                s += (dl * dl) * pred;
            }
        }
        return log_count(n1, p.power) + ((-0.5 * D * log(2.0 * M_PI) + 0.5 * lp) - 0.5 * s);
    }
    const double n = ldr<CG>(rec + SO + SC_N);
    const double f = ldr<CG>(rec + SO + SC_F);
    const double ld = ldr<CG>(rec + SO + SC_LOGDET);
    const double n1 = n - 1.0;
    const double kap = p.k0 + n, kap1 = p.k0 + n1;
    const double beta = kap / kap1, r = kap / kap1;
    const double lc1 = log_count(n1, p.power);
    if (COV == COV_FULL) {
        const long long nu1 = p.v0 + (long long)n1 - D + 1;
        const double f1 = (kap1 + 1.) / (kap1 * (double)nu1);
        const double q = quad_full<DP, CG>(rec, x);
        const double s = f * q;
        const double om = 1.0 - beta * s;
        if (!(om > OM_MIN)) { *ok = false; return 0.0; }
        const double ld1 = D * log(f1) + (ld - D * log(f)) + log(om);
        const double q1 = r * r * s / (f1 * om);
        const double c1 = t_const_full(p.lgam, p.logv, nu1, D, p.log_pi, ld1);
        return lc1 + (c1 - (nu1 + D) / 2. * log(1.0 + 1. / nu1 * q1));
    } else {
        const long long v1 = p.v0 + (long long)n1;
        const double f1 = (kap1 + 1.) / (kap1 * (double)v1);
        double lo = 0.0, term = 0.0;
        bool good = true;
#pragma unroll
        for (int a = 0; a < DP; ++a) {
            if (a < D) {
                const double dl = ldr<CG>(rec + a) - x[a];
                const double s = f * ((dl * dl) * ldr<CG>(rec + DP + a));
                const double om = 1.0 - beta * s;
                if (!(om > OM_MIN)) good = false;
                lo += log(om);
                term += log(1.0 + r * r * s / (f1 * (double)v1 * om));
            }
        }
        if (!good) { *ok = false; return 0.0; }
        const double lpv1 = D * log(f1) + (ld - D * log(f)) + lo;
        const double c1 = t_const_diag(p.lgam, p.logv, v1, D, p.log_pi, lpv1);
        return lc1 + (c1 - (v1 + 1) / 2. * term);
    }
}

// ---------------------------------------------------------------------------------------------
// Refactorisation of one component by ONE WARP                    (gaussian_components.py:319-331 /
// _diag.py:325-338).  Builds the predictive covariance from the statistics in the reference's operation
// order, factors it (Cholesky instead of the reference's LU inv+slogdet: same quantities to ~1e-14 for a
// positive definite matrix) and writes the evaluation record to global memory and, if rec_s != null, to
// the CTA's shared-memory copy.  A: shared scratch of packed_len(D) doubles (full only).
// mode 0: component `slot` from (num,S,n);  mode 1: the prior alone (log_prior, :207-214 / _diag :215-222).
// Returns false (all lanes) when the matrix is not positive definite.
// ---------------------------------------------------------------------------------------------
template <int DP, int COV>
__device__ bool refactor_warp(const Params &p, const double *num, const double *S, long long n_cnt, int mode,
                              double *rec_g, double *rec_s, double *A) {
    const int lane = threadIdx.x & 31;
    const int D = p.D;
    constexpr int SO = rec_sc_off(DP, COV);
    constexpr int MU = rec_mu_off(DP, COV);
    const double nn = (double)n_cnt;
    const double kap = p.k0 + nn;
    const double vN = (double)(p.v0 + n_cnt);
    bool bad = false;
    double logdet = 0.0, f;
    long long nu;
    auto put = [&](int idx, double v) {
        __stcg(rec_g + idx, v);
        if (rec_s) rec_s[idx] = v;
    };
    if (COV == COV_FULL) {
        nu = p.v0 + n_cnt - D + 1;
        f = (kap + 1.) / (kap * (vN - D + 1.));
        // covar = f * (S - kap * outer(m, m)), m = num/kap            (:326-329); prior: f * S_0 (:210)
        for (int a = 0; a < D; ++a) {
            const double ma = (mode == 0) ? __ldcg(num + a) / kap : 0.0;
            for (int b = lane; b <= a; b += 32) {
                double v;
                if (mode == 0) {
                    const double mb = __ldcg(num + b) / kap;
                    v = f * (__ldcg(S + row_idx(a, b)) - kap * (ma * mb));
                } else {
                    v = f * p.S0[row_idx(a, b)];
                }
                A[row_idx(a, b)] = v;
            }
        }
        __syncwarp();
        for (int j = 0; j < D; ++j) {
            const double ajj = A[row_idx(j, j)];
            if (!(ajj > 0.0) || !(ajj < 1e300)) { bad = true; break; }
            const double inv = 1.0 / sqrt(ajj);
            logdet += log(ajj);
            __syncwarp();
            for (int a = j + 1 + lane; a < D; a += 32) A[row_idx(a, j)] *= inv;
            if (lane == 0) A[row_idx(j, j)] = inv;
            __syncwarp();
            for (int a = j + 1 + lane; a < D; a += 32) {
                const double laj = A[row_idx(a, j)];
                for (int b = j + 1; b <= a; ++b) A[row_idx(a, b)] = fma(-laj, A[row_idx(b, j)], A[row_idx(a, b)]);
            }
            __syncwarp();
        }
        if (bad) return false;
        // record: column-major packed over DP, identity on the padded dimensions
        for (int b = 0; b < DP; ++b) {
            const int off = col_off(DP, b);
            for (int a = b + lane; a < DP; a += 32) {
                double v;
                if (a < D && b < D) v = A[row_idx(a, b)];
                else v = (a == b) ? 1.0 : 0.0;
                put(off + (a - b), v);
            }
        }
    } else if (COV == COV_FIXED) {
        // predictive precision tau_N tau / (tau_N + tau), its log product (gaussian_components_fixedvar.py:278-286);
        // the prior alone: precision 1 / var_0 (:204-210)
        nu = 0; f = 0.0;
        double lp = 0.0;
        for (int a = 0; a < D; ++a) {
            double pred, tn = 0.0, nm = 0.0;
            if (mode == 0) {
                tn = __ldcg(S + a);
                nm = __ldcg(num + a);
                pred = tn * p.tau[a] / (tn + p.tau[a]);
            } else {
                pred = p.S0[a];
            }
            if (!(pred > 0.0) || !(pred < 1e300)) bad = true;
            lp += log(pred);
            if (lane == 0) { put(DP + a, pred); put(2 * DP + a, nm); put(3 * DP + a, tn); }
        }
        if (bad) return false;
        logdet = lp;
        for (int a = D + lane; a < DP; a += 32) { put(DP + a, 0.0); put(2 * DP + a, 0.0); put(3 * DP + a, 0.0); }
    } else {
        nu = p.v0 + n_cnt;
        f = (kap + 1.) / (kap * vN);
        // var = f * (S - kap * m^2); log_prod_var = sum log var; inv_var = 1/var    (_diag.py:334-338; prior :218-220)
        double lp = 0.0;
        for (int a = 0; a < D; ++a) {  // every lane computes all dims (D is small); lane 0's values are stored
            double var;
            if (mode == 0) {
                const double m = __ldcg(num + a) / kap;
                var = f * (__ldcg(S + a) - kap * (m * m));
            } else {
                var = f * p.S0[a];
            }
            if (!(var > 0.0) || !(var < 1e300)) bad = true;
            lp += log(var);
            if (lane == 0) put(DP + a, 1. / var);
        }
        if (bad) return false;
        logdet = lp;
        for (int a = D + lane; a < DP; a += 32) put(DP + a, 0.0);
    }
    for (int a = lane; a < DP; a += 32) {
        double m = 0.0;
        if (a < D) {
            if (COV == COV_FIXED) m = (mode == 0) ? __ldcg(num + a) / __ldcg(S + a) : p.m0[a];
            else m = (mode == 0) ? __ldcg(num + a) / kap : p.m0[a];
        }
        put(MU + a, m);
    }
    if (lane == 0) {
        double c, h;
        if (COV == COV_FULL) {
            c = t_const_full(p.lgam, p.logv, nu, D, p.log_pi, logdet);
            h = (nu + D) / 2.;
        } else if (COV == COV_FIXED) {
            c = -0.5 * D * log(2.0 * M_PI) + 0.5 * logdet;
            h = 0.0;
            nu = 1;
        } else {
            c = t_const_diag(p.lgam, p.logv, nu, D, p.log_pi, logdet);
            h = (nu + 1) / 2.;
        }
        put(SO + SC_C, c);
        put(SO + SC_H, h);
        put(SO + SC_INVNU, 1. / (double)nu);
        put(SO + SC_LC, n_cnt > 0 ? log_count(nn, p.power) : 0.0);
        put(SO + SC_N, nn);
        put(SO + SC_LOGDET, logdet);
        put(SO + SC_F, f);
        put(SO + SC_SPARE, 0.0);
    }
    __syncwarp();
    return true;
}

// Philox4x32-10 (Salmon et al. 2011), one 53-bit uniform per counter, assembled like CPython's
// random.random(): (a>>5, b>>6) -> (a*2^26 + b) / 2^53.
__host__ __device__ inline void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    const uint32_t n1 = (uint32_t)p1;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    const uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}
__host__ __device__ inline double philox_uniform(uint64_t seed, uint64_t sweep, uint64_t j) {
    uint32_t c[4] = {(uint32_t)j, (uint32_t)(j >> 32), (uint32_t)sweep, (uint32_t)(sweep >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
    const uint32_t a = c[0] >> 5, b = c[1] >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}

}  // namespace bgmm
