// bgmm_fast.cuh -- the B200 sweep engine for full covariance (NIW) components with padded D <= 16.
//
// Semantics reproduced: the per-datum loop of CRPMM.collapsed_gibbs_sampler (igmm/crpmm.py:57-88) and
// PCRPMM.collapsed_gibbs_sampler (igmm/pcrpmm.py:93-131), strictly sequential over the scan order.
//
// Execution model (DESIGN.md "Engine"): a REPLICATED STATE MACHINE.  Every CTA of a cooperative grid
// (one per SM) keeps the evaluation records of all live components in its own shared memory and applies
// every state change itself, with identical arithmetic, so the replicas never need to exchange records.
//   * evaluation record of component k (element-major in shared memory, rec[e * KS + k]):
//       B = S_N^-1 (packed lower triangle), the mean m_N, and the scalars of the Student-t log pdf.
//     A datum joining / leaving a component is a rank-one change of S_N (gaussian_components.py:161-166,
//     :184-185), so B follows by Sherman-Morrison in O(D^2) and log|S_N| by the matrix determinant lemma;
//     records are rebuilt from the bit-exact sufficient statistics at the start of every sweep and after
//     REFRESH_EVERY rank-one updates of a component (drift control).
//   * window round: the CTAs split a window of upcoming scan positions (a warp per datum, lanes over
//     components), each datum is evaluated against the current records; a datum whose draw keeps it where
//     it is leaves the state untouched (the reference's restore path, crpmm.py:82-85), so every "stay"
//     in front of the first datum that does anything else is exactly the sequential chain's decision.
//     One atomicMin + ONE grid barrier publish that first position; every CTA then resolves it itself.
//   * sequential batch: when movers are dense, every CTA walks the scan datum by datum (no barriers).
//   * CTA 0 is the only writer of global state: labels, the bit-exact statistics (same operation order
//     as the reference: one rounded multiply and one rounded add per element), counters.
#pragma once
#include "bgmm_sweep.cuh"

namespace bgmm {
namespace fast {

constexpr int TF = 512;               // threads per CTA
constexpr int NWARP = TF / 32;
constexpr int NSC = 16;               // scalars per record
constexpr int SEQ_BATCH = 16;         // data staged per sequential batch
constexpr int REFRESH_EVERY = 1024;   // rank-one updates of a record before it is rebuilt from the statistics
constexpr int NB_MAX = 6;             // weights per lane in the draw: supports K + 1 <= 192
constexpr int E_NEED_GENERIC = 1;     // internal: a birth would exceed the resident capacity -> generic engine
constexpr double GAP_TO_WIN = 6.0, GAP_TO_SEQ = 3.0;
constexpr int WIN_PASSES_MAX = 8;

// record scalars
enum {
    F_N = 0,   // count n (as double)
    F_LDS,     // log|S_N|
    F_TC,      // Student-t constant of nu = v0 + n - D + 1 (gaussian_components.py:237-249 without logdet)
    F_LF,      // log f(n), f = (kappa + 1) / (kappa nu)   (:324-329)
    F_LC,      // log count prior of n (crpmm.py:70 / pcrpmm.py:107)
    F_TCM, F_LFM, F_LCM,  // the same three for n - 1 (the datum's own component with the datum removed)
    F_CW,      // LC + TC - (D LF + LDS) / 2
    F_G,       // kappa / (kappa + 1) = 1 / (f nu)
    F_H,       // (nu + D) / 2
    F_BETA,    // kappa / (kappa - 1)
    F_CWO,     // LCM + TCM - (D LFM + LDS) / 2
    F_HO,      // (nu - 1 + D) / 2 - 1 / 2
    F_CNT,     // rank-one updates since the record was rebuilt
    F_SPARE
};

// grid barrier with a watchdog: replicas that stopped agreeing would otherwise spin forever
__device__ __forceinline__ void f_grid_barrier(Ctl *c) {
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
            const long long t0 = clock64();
            while (ld_acquire_u32(&c->bar_gen) == gen) {
                __nanosleep(20);
                if (clock64() - t0 > 8000000000LL) __trap();
            }
        }
        __threadfence();
    }
    __syncthreads();
}

template <int DP> struct Lay {
    static constexpr int PP = DP * (DP + 1) / 2;
    static constexpr int MU = PP;
    static constexpr int SC = PP + DP;
    static constexpr int R = PP + DP + NSC;
    static constexpr int NS = PP + DP;  // statistics per component: S (packed) then num
};

struct FSh {
    // replicated chain state
    int K, n_free, error, mode;
    long long pos;
    double gap;
    long long moves, births, deaths, evals, windows, seq_data, wasted, explicit_evals, refreshes;
    unsigned long long margin_bits;
    // datum being resolved
    long long i;
    int uid, k_old, k_new, died, need_explicit, explicit_done, refresh_a, refresh_b;
    double n_old, u, lp, margin;
    // update scratch
    double fresh[2][3];   // [a|b][TC, LF, LC]
    double lds_new[2], n_new[2];
    int upd_a, upd_b, birth;
    unsigned int round;
    unsigned long long mbar;
};

template <int DP> struct FSmem {
    double *rec;      // R * KS
    double *prior;    // R
    double *tmprec;   // R  (exact record of the datum's own component with the datum removed)
    double *wrow;     // NWARP * WS
    double *xb;       // SEQ_BATCH * DP
    double *ub, *lpb; // SEQ_BATCH
    long long *ib;    // SEQ_BATCH
    int *uidb;        // SEQ_BATCH
    double *dv;       // 2 * DP   (d of the two updated components)
    double *vv;       // 2 * DP   (v = B d)
    double *A;        // PP       (exact refactor scratch)
    double *W;        // DP * DP
    double *mm;       // DP
    int *slot_of_uid, *uid_of_slot, *uid_free;  // K_max each
    unsigned short *rc;  // PP: (a << 8) | b of packed element e
    FSh *sh;
    int WS;
};

__host__ __device__ inline int fast_ws(int KS) { return (KS + 1 + 3) & ~3; }

template <int DP> __host__ __device__ inline size_t fast_smem_bytes(int KS, int K_max) {
    using Ly = Lay<DP>;
    size_t d = (size_t)Ly::R * KS + 2 * (size_t)Ly::R + (size_t)NWARP * fast_ws(KS) + (size_t)SEQ_BATCH * DP +
               2 * SEQ_BATCH + SEQ_BATCH /*ib*/ + SEQ_BATCH / 2 /*uidb*/ + 4 * DP + Ly::PP + DP * DP + DP;
    size_t b = d * sizeof(double);
    b += 3 * (size_t)K_max * sizeof(int);
    b += ((size_t)Ly::PP * sizeof(unsigned short) + 15) & ~(size_t)15;
    b += sizeof(FSh) + 64;
    return (b + 15) & ~(size_t)15;
}

template <int DP> __device__ inline FSmem<DP> fast_carve(double *base, const Params &p) {
    using Ly = Lay<DP>;
    FSmem<DP> s;
    s.WS = fast_ws(p.KS);
    double *q = base;
    s.rec = q; q += (size_t)Ly::R * p.KS;
    q = (double *)(((uintptr_t)q + 15) & ~(uintptr_t)15);
    s.prior = q; q += Ly::R;
    s.tmprec = q; q += Ly::R;
    s.wrow = q; q += (size_t)NWARP * s.WS;
    s.xb = q; q += SEQ_BATCH * DP;
    s.ub = q; q += SEQ_BATCH;
    s.lpb = q; q += SEQ_BATCH;
    s.ib = (long long *)q; q += SEQ_BATCH;
    s.uidb = (int *)q; q += SEQ_BATCH / 2;
    s.dv = q; q += 2 * DP;
    s.vv = q; q += 2 * DP;
    s.A = q; q += Ly::PP;
    s.W = q; q += DP * DP;
    s.mm = q; q += DP;
    int *t = (int *)q;
    s.slot_of_uid = t; t += p.K_max;
    s.uid_of_slot = t; t += p.K_max;
    s.uid_free = t; t += p.K_max;
    unsigned short *r = (unsigned short *)t;
    s.rc = r; r += Ly::PP;
    s.sh = (FSh *)(((uintptr_t)r + 15) & ~(uintptr_t)15);
    return s;
}

// ---------------------------------------------------------------------------------------------
// scalar pieces of a record
// ---------------------------------------------------------------------------------------------
// Student-t constant for integer nu (gaussian_components.py:237-249): lgamma((nu+D)/2) - lgamma(nu/2)
// - D/2 log(nu) - D/2 log(pi), tables indexed like the reference's (index = the integer itself).
__device__ __forceinline__ double f_tc(const Params &p, long long nu) {
    const int D = p.D;
    return __ldg(p.lgam + nu + D) - __ldg(p.lgam + nu) - D / 2. * __ldg(p.logv + nu) - D / 2. * p.log_pi;
}
__device__ __forceinline__ double f_lf(const Params &p, double n) {
    const double kap = p.k0 + n;
    const double nu = (double)(p.v0 - p.D + 1) + n;
    return log((kap + 1.) / (kap * nu));
}
// piece `which` (0 TC, 1 LF, 2 LC) of count n; n < 1 gives 0 (never used)
__device__ __forceinline__ double f_piece(const Params &p, int which, double n) {
    if (which == 0) return f_tc(p, p.v0 - p.D + 1 + (long long)n);
    if (which == 1) return f_lf(p, n);
    return n >= 1.0 ? log_count(n, p.power) : 0.0;
}

// derived scalars from (n, lds, current pieces, minus-one pieces); writes the NSC scalars with stride `st`
__device__ __forceinline__ void f_write_scalars(const Params &p, double *sc, int st, double n, double lds, double tc,
                                                double lf, double lc, double tcm, double lfm, double lcm, double cnt) {
    const int D = p.D;
    const double kap = p.k0 + n;
    const double nu = (double)(p.v0 - D + 1) + n;
    sc[F_N * st] = n;
    sc[F_LDS * st] = lds;
    sc[F_TC * st] = tc; sc[F_LF * st] = lf; sc[F_LC * st] = lc;
    sc[F_TCM * st] = tcm; sc[F_LFM * st] = lfm; sc[F_LCM * st] = lcm;
    sc[F_CW * st] = lc + tc - 0.5 * (D * lf + lds);
    sc[F_G * st] = kap / (kap + 1.);
    sc[F_H * st] = (nu + D) / 2.;
    sc[F_BETA * st] = kap / (kap - 1.);
    sc[F_CWO * st] = lcm + tcm - 0.5 * (D * lfm + lds);
    sc[F_HO * st] = (nu - 1. + D) / 2. - 0.5;
    sc[F_CNT * st] = cnt;
    sc[F_SPARE * st] = 0.0;
}

// ---------------------------------------------------------------------------------------------
// evaluation: delta^T B delta with delta = m - x; col points at element 0 of the component, stride st
// ---------------------------------------------------------------------------------------------
template <int DP>
__device__ __forceinline__ double f_quad(const double *__restrict__ col, int st, const double (&x)[DP]) {
    using Ly = Lay<DP>;
    double d[DP];
    const double *pm = col + (size_t)Ly::MU * st;
#pragma unroll
    for (int a = 0; a < DP; ++a) d[a] = pm[(size_t)a * st] - x[a];
    const double *pb = col;
    double q = 0.0;
#pragma unroll
    for (int a = 0; a < DP; ++a) {
        double r = 0.0;
#pragma unroll
        for (int b = 0; b < a; ++b) { r = fma(*pb, d[b], r); pb += st; }
        r = fma(0.5 * (*pb), d[a], r);
        pb += st;
        q = fma(d[a], r, q);
    }
    return 2.0 * q;
}

// weight of a component the datum is not in: log count prior + log_post_pred (crpmm.py:70-72)
template <int DP>
__device__ __forceinline__ double f_weight_other(const double *__restrict__ col, int st, const double (&x)[DP]) {
    using Ly = Lay<DP>;
    const double *sc = col + (size_t)Ly::SC * st;
    const double q = f_quad<DP>(col, st, x);
    return sc[F_CW * st] - sc[F_H * st] * log(1.0 + sc[F_G * st] * q);
}
// weight of the datum's own component with the datum removed (del_item then log_post_pred,
// gaussian_components.py:171-186, :228-251) in closed form: S_N' = S_N - beta d d^T, so
// |S_N'| = |S_N| om, om = 1 - beta d^T B d, and 1 + q'/nu' = 1 / om.
template <int DP>
__device__ __forceinline__ double f_weight_own(const double *__restrict__ col, int st, const double (&x)[DP], bool *ok) {
    using Ly = Lay<DP>;
    const double *sc = col + (size_t)Ly::SC * st;
    const double q = f_quad<DP>(col, st, x);
    const double om = 1.0 - sc[F_BETA * st] * q;
    if (!(om > OM_MIN)) { *ok = false; return 0.0; }
    return sc[F_CWO * st] + sc[F_HO * st] * log(om);
}

// ---------------------------------------------------------------------------------------------
// logsumexp + draw (crpmm.py:75-78, utils.py:7-20) by one warp over n = K + 1 weights in shared memory.
// Lane l owns the nb consecutive weights starting at l * nb.  Returns (all lanes) the drawn index; *margin is the
// distance of u to the nearest boundary of the drawn interval (probability units); *bad is set when the
// normaliser is not finite / positive.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int f_warp_draw(const double *__restrict__ w, int n, double u, double *margin, bool *bad) {
    const int lane = threadIdx.x & 31;
    const int nb = (n + 31) >> 5;
    const int lo = lane * nb;
    double v[NB_MAX];
    float fm = -INFINITY;
#pragma unroll
    for (int t = 0; t < NB_MAX; ++t) {
        v[t] = (t < nb && lo + t < n) ? w[lo + t] : -INFINITY;
        fm = fmaxf(fm, __double2float_rn(v[t]));
    }
    int key = __float_as_int(fm);
    key ^= (key >> 31) & 0x7fffffff;
    key = __reduce_max_sync(0xffffffffu, key);
    key ^= (key >> 31) & 0x7fffffff;
    const double M = (double)__int_as_float(key);  // within float rounding of the true maximum: a safe scale
    double c[NB_MAX];
    double run = 0.0;
#pragma unroll
    for (int t = 0; t < NB_MAX; ++t) {
        const double dlt = v[t] - M;
        const double e = (dlt < EXP_CUTOFF) ? 0.0 : exp(dlt);
        run += (t < nb) ? e : 0.0;
        c[t] = run;
    }
    double incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    double excl = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) excl = 0.0;
    const double s = __shfl_sync(0xffffffffu, incl, 31);
    const double t0 = u * s;
    int cand = -1;
    double lower = excl, mg = 0.0;
#pragma unroll
    for (int t = 0; t < NB_MAX; ++t) {
        if (cand < 0 && t < nb && lo + t < n) {
            const double upper = excl + c[t];
            if (upper > t0) { cand = lo + t; mg = fmin(t0 - lower, upper - t0); }
            lower = upper;
        }
    }
    const unsigned who = __ballot_sync(0xffffffffu, cand >= 0);
    int k;
    if (who == 0u) {
        k = n - 1;  // utils.py:20 fallback: the last index
        mg = 0.0;
    } else {
        const int src = __ffs(who) - 1;
        k = __shfl_sync(0xffffffffu, cand, src);
        mg = __shfl_sync(0xffffffffu, mg, src) / s;
    }
    *margin = mg;
    *bad = !(s > 0.0) || !(s < INFINITY);
    return k;
}

// ---------------------------------------------------------------------------------------------
// exact record of one component by ONE WARP from the bit-exact statistics (gaussian_components.py:319-331):
// S_N = S - kappa m m^T, Cholesky, log|S_N|, B = S_N^-1.  mode 0: component statistics in global memory,
// optionally with datum x removed first (the reference's del_item arithmetic, :184-185); mode 1: the prior
// alone (n = 0, :161-164).  out has stride ost.  Returns false (all lanes) if S_N is not positive definite.
// ---------------------------------------------------------------------------------------------
template <int DP>
__device__ bool f_exact_record_warp(const Params &p, int mode, const double *num_g, const double *S_g, double n_after,
                                    const double *xrm, const unsigned short *rc, double *A, double *W, double *mm,
                                    double *out, int ost) {
    using Ly = Lay<DP>;
    const int lane = threadIdx.x & 31;
    const int D = p.D;
    const double kap = p.k0 + n_after;
    for (int r = lane; r < DP; r += 32) {
        double v = 0.0;
        if (r < D) {
            if (mode == 1) v = __dmul_rn(p.k0, p.m0[r]);
            else { v = __ldcg(num_g + r); if (xrm) v = __dsub_rn(v, xrm[r]); }
        }
        mm[r] = v / kap;
    }
    __syncwarp();
    for (int e = lane; e < Ly::PP; e += 32) {
        const int a = rc[e] >> 8, b = rc[e] & 0xff;
        double v;
        if (a < D) {  // b <= a
            double sv;
            if (mode == 1) sv = __dadd_rn(p.S0[e], __dmul_rn(p.k0, __dmul_rn(p.m0[a], p.m0[b])));
            else { sv = __ldcg(S_g + e); if (xrm) sv = __dsub_rn(sv, __dmul_rn(xrm[a], xrm[b])); }
            v = sv - kap * (mm[a] * mm[b]);
        } else {
            v = (a == b) ? 1.0 : 0.0;
        }
        A[e] = v;
    }
    __syncwarp();
    double lds = 0.0;
    bool bad = false;
    for (int j = 0; j < DP; ++j) {
        const double ajj = A[row_idx(j, j)];
        if (!(ajj > 0.0) || !(ajj < 1e300)) { bad = true; break; }
        const double inv = 1.0 / sqrt(ajj);
        lds += log(ajj);
        __syncwarp();
        for (int a = j + 1 + lane; a < DP; a += 32) A[row_idx(a, j)] *= inv;
        if (lane == 0) A[row_idx(j, j)] = inv;  // reciprocal of L_jj
        __syncwarp();
        for (int a = j + 1 + lane; a < DP; a += 32) {
            const double laj = A[row_idx(a, j)];
            for (int b = j + 1; b <= a; ++b) A[row_idx(a, b)] = fma(-laj, A[row_idx(b, j)], A[row_idx(a, b)]);
        }
        __syncwarp();
    }
    if (bad) return false;
    // W = L^-1, one column per lane
    for (int e = lane; e < DP; e += 32) {
        for (int a = 0; a < DP; ++a) {
            double sacc = (a == e) ? 1.0 : 0.0;
            for (int b = e; b < a; ++b) sacc = fma(-A[row_idx(a, b)], W[b * DP + e], sacc);
            W[a * DP + e] = (a < e) ? 0.0 : sacc * A[row_idx(a, a)];
        }
    }
    __syncwarp();
    for (int e = lane; e < Ly::PP; e += 32) {
        const int a = rc[e] >> 8, b = rc[e] & 0xff;
        double sacc = 0.0;
        for (int c2 = a; c2 < DP; ++c2) sacc = fma(W[c2 * DP + a], W[c2 * DP + b], sacc);
        out[(size_t)e * ost] = sacc;
    }
    for (int r = lane; r < DP; r += 32) out[(size_t)(Ly::MU + r) * ost] = (r < D) ? mm[r] : 0.0;
    // scalar pieces: lanes 0..5
    double piece = 0.0;
    if (lane < 3) piece = f_piece(p, lane, n_after);
    else if (lane < 6) piece = (n_after >= 2.0) ? f_piece(p, lane - 3, n_after - 1.0) : 0.0;
    const double tc = __shfl_sync(0xffffffffu, piece, 0), lf = __shfl_sync(0xffffffffu, piece, 1);
    const double lc = __shfl_sync(0xffffffffu, piece, 2), tcm = __shfl_sync(0xffffffffu, piece, 3);
    const double lfm = __shfl_sync(0xffffffffu, piece, 4), lcm = __shfl_sync(0xffffffffu, piece, 5);
    if (lane == 0) f_write_scalars(p, out + (size_t)Ly::SC * ost, ost, n_after, lds, tc, lf, lc, tcm, lfm, lcm, 0.0);
    __syncwarp();
    return true;
}

// ---------------------------------------------------------------------------------------------
// rank-one update of one record by ONE WARP.  sign = -1: the datum leaves (del_item), sign = +1: it joins
// (add_item).  x in shared memory.  Writes B, m; leaves log|S_N|' and n' in sh.lds_new / sh.n_new [which].
// ---------------------------------------------------------------------------------------------
template <int DP>
__device__ void f_rank_one_warp(const Params &p, const FSmem<DP> &s, int k, int sign, const double *x, int which) {
    using Ly = Lay<DP>;
    const int lane = threadIdx.x & 31;
    const int st = p.KS;
    double *col = s.rec + k;
    double *dv = s.dv + which * DP, *vv = s.vv + which * DP;
    const double n = col[(size_t)(Ly::SC + F_N) * st];
    const double kap = p.k0 + n;
    if (lane < DP) dv[lane] = x[lane] - col[(size_t)(Ly::MU + lane) * st];
    __syncwarp();
    if (lane < DP) {
        double acc = 0.0;
#pragma unroll
        for (int b = 0; b < DP; ++b) {
            const int e = (lane >= b) ? row_idx(lane, b) : row_idx(b, lane);
            acc = fma(col[(size_t)e * st], dv[b], acc);
        }
        vv[lane] = acc;
    }
    __syncwarp();
    double sq = 0.0;
#pragma unroll
    for (int b = 0; b < DP; ++b) sq = fma(dv[b], vv[b], sq);
    double gam, kap2, lds_add;
    if (sign < 0) {
        const double beta = kap / (kap - 1.);
        const double om = 1.0 - beta * sq;
        gam = beta / om;
        kap2 = kap - 1.;
        lds_add = log(om);
    } else {
        const double beta = kap / (kap + 1.);
        const double den = 1.0 + beta * sq;
        gam = -beta / den;
        kap2 = kap + 1.;
        lds_add = log(den);
    }
    for (int e = lane; e < Ly::PP; e += 32) {
        const int a = s.rc[e] >> 8, b = s.rc[e] & 0xff;
        double *pe = col + (size_t)e * st;
        *pe = fma(gam * vv[a], vv[b], *pe);
    }
    if (lane < DP) {
        double *pm = col + (size_t)(Ly::MU + lane) * st;
        *pm = (sign < 0) ? (*pm - dv[lane] / kap2) : (*pm + dv[lane] / kap2);
    }
    if (lane == 0) {
        s.sh->lds_new[which] = col[(size_t)(Ly::SC + F_LDS) * st] + lds_add;
        s.sh->n_new[which] = n + (double)sign;
    }
    __syncwarp();
}

// bit-exact statistics update in global memory by `nthr` threads of CTA 0 (thread rank t):
// S -= / += fl(x_a x_b), num -= / += x_a   (gaussian_components.py:165-166, :184-185)
template <int DP>
__device__ __forceinline__ void f_stats_axpy(const Params &p, const unsigned short *rc, int slot, const double *x,
                                             int sign, bool init_prior, int t, int nthr) {
    using Ly = Lay<DP>;
    const int D = p.D;
    double *S = p.S + (size_t)slot * Ly::PP;
    double *num = p.num + (size_t)slot * DP;
    for (int e = t; e < Ly::NS; e += nthr) {
        if (e < Ly::PP) {
            const int a = rc[e] >> 8, b = rc[e] & 0xff;
            if (a >= D) { if (init_prior) __stcg(S + e, 0.0); continue; }
            const double o = __dmul_rn(x[a], x[b]);
            const double v = init_prior ? __dadd_rn(p.S0[e], __dmul_rn(p.k0, __dmul_rn(p.m0[a], p.m0[b]))) : __ldcg(S + e);
            __stcg(S + e, sign > 0 ? __dadd_rn(v, o) : __dsub_rn(v, o));
        } else {
            const int a = e - Ly::PP;
            if (a >= D) { if (init_prior) __stcg(num + a, 0.0); continue; }
            const double v = init_prior ? __dmul_rn(p.k0, p.m0[a]) : __ldcg(num + a);
            __stcg(num + a, sign > 0 ? __dadd_rn(v, x[a]) : __dsub_rn(v, x[a]));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Resolve ONE datum (whole CTA, every CTA identically).  Inputs in the staging buffers at index jj.
// ---------------------------------------------------------------------------------------------
template <int DP>
__device__ void f_step(const Params &p, const FSmem<DP> &s, int jj) {
    using Ly = Lay<DP>;
    FSh &sh = *s.sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int st = p.KS;
    const bool cta0 = (blockIdx.x == 0);
    const double *xs = s.xb + jj * DP;

    if (tid == 0) {
        const int uid = s.uidb[jj];
        int k_old = -1;
        double n_old = 0.0;
        if (uid >= 0) { k_old = s.slot_of_uid[uid]; n_old = s.rec[(size_t)(Ly::SC + F_N) * st + k_old]; }
        sh.i = s.ib[jj]; sh.uid = uid; sh.k_old = k_old; sh.n_old = n_old;
        sh.u = s.ub[jj]; sh.lp = s.lpb[jj];
        sh.died = 0; sh.need_explicit = 0; sh.explicit_done = 0; sh.refresh_a = 0; sh.refresh_b = 0;
    }
    __syncthreads();
    const int k_old = sh.k_old;
    const long long i = sh.i;
    bool died = false;
    if (k_old >= 0 && sh.n_old == 1.0) {
        // del_item empties the component: del_component (gaussian_components.py:188-205), swap with the last slot
        const int L = sh.K - 1;
        if (k_old != L) {
            for (int e = tid; e < Ly::R; e += TF) s.rec[(size_t)e * st + k_old] = s.rec[(size_t)e * st + L];
            if (cta0) {
                for (int e = tid; e < Ly::PP; e += TF)
                    __stcg(p.S + (size_t)k_old * Ly::PP + e, __ldcg(p.S + (size_t)L * Ly::PP + e));
                for (int e = tid; e < DP; e += TF)
                    __stcg(p.num + (size_t)k_old * DP + e, __ldcg(p.num + (size_t)L * DP + e));
            }
        }
        __syncthreads();
        if (tid == 0) {
            const int uid_dead = s.uid_of_slot[k_old];
            if (k_old != L) {
                const int uid_l = s.uid_of_slot[L];
                s.uid_of_slot[k_old] = uid_l;
                s.slot_of_uid[uid_l] = k_old;
            }
            s.uid_of_slot[L] = -1;
            s.slot_of_uid[uid_dead] = -1;
            s.uid_free[sh.n_free] = uid_dead;
            sh.n_free += 1;
            sh.K = L;
            sh.deaths += 1;
            sh.died = 1;
        }
        died = true;
        __syncthreads();
    }
    const int K = sh.K;

    // weights of the live components (crpmm.py:68-74): thread k evaluates component k
    if (tid < K) {
        double x[DP];
#pragma unroll
        for (int a = 0; a < DP; ++a) x[a] = xs[a];
        double w;
        if (tid == k_old && !died) {
            bool ok = true;
            w = f_weight_own<DP>(s.rec + tid, st, x, &ok);
            if (!ok) sh.need_explicit = 1;
        } else {
            w = f_weight_other<DP>(s.rec + tid, st, x);
        }
        s.wrow[tid] = w;
    }
    if (tid == K) s.wrow[K] = p.log_alpha + sh.lp;
    __syncthreads();

    if (sh.need_explicit) {
        // the closed form is not trusted: build the reduced component's record exactly from the statistics
        // (every CTA; CTA 0's statistics writes are made visible by the barrier)
        f_grid_barrier(p.ctl);
        if (warp == 0) {
            const bool okf = f_exact_record_warp<DP>(p, 0, p.num + (size_t)k_old * DP, p.S + (size_t)k_old * Ly::PP,
                                                     sh.n_old - 1.0, xs, s.rc, s.A, s.W, s.mm, s.tmprec, 1);
            if (!okf) { if (lane == 0) sh.error = -4; }
            else if (lane == 0) {
                double x[DP];
#pragma unroll
                for (int a = 0; a < DP; ++a) x[a] = xs[a];
                s.wrow[k_old] = f_weight_other<DP>(s.tmprec, 1, x);
                sh.explicit_done = 1;
                sh.explicit_evals += 1;
            }
        }
        // CTA 0 must not touch the statistics again before every replica has read them
        f_grid_barrier(p.ctl);
    }

    if (warp == 0) {
        double mg;
        bool bad;
        const int k = f_warp_draw(s.wrow, K + 1, sh.u, &mg, &bad);
        if (lane == 0) {
            sh.k_new = k;
            sh.margin = mg;
            if (bad) sh.error = -4;
            sh.evals += K;
            const unsigned long long mb = (unsigned long long)__double_as_longlong(mg);
            if (mb < sh.margin_bits) sh.margin_bits = mb;
        }
    }
    __syncthreads();
    if (sh.error) return;
    const int k_new = sh.k_new;
    if (k_new == k_old && !died) return;  // stay: nothing was touched (crpmm.py:82-85)

    // ---- the datum moves: add_item (gaussian_components.py:154-169) ----
    const bool birth = (k_new == K);
    if (birth) {
        if (K >= p.K_max) { if (tid == 0) sh.error = -3; __syncthreads(); return; }
        if (K >= p.Kcap) {  // nothing has been changed for this datum yet: the generic engine redoes it
            if (tid == 0) { sh.error = E_NEED_GENERIC; sh.evals -= K; }
            __syncthreads();
            return;
        }
        for (int e = tid; e < Ly::R; e += TF) s.rec[(size_t)e * st + K] = s.prior[e];
        if (tid == 0) {
            const int uid = s.uid_free[sh.n_free - 1];
            sh.n_free -= 1;
            s.uid_of_slot[K] = uid;
            s.slot_of_uid[uid] = K;
            sh.K = K + 1;
            sh.births += 1;
        }
    }
    const bool remove_now = (k_old >= 0) && !died;
    const bool expl = sh.explicit_done != 0;
    __syncthreads();
    // update phase: warps 0/1 the two records, warps 2..7 the fresh scalar pieces, CTA 0 warps 8.. the statistics
    const double n_a = remove_now ? sh.n_old - 1.0 : 0.0;
    const double n_b = s.rec[(size_t)(Ly::SC + F_N) * st + k_new] + 1.0;
    if (warp == 0) {
        if (remove_now) {
            if (expl) {
                for (int e = lane; e < Ly::R; e += 32) s.rec[(size_t)e * st + k_old] = s.tmprec[e];
            } else {
                f_rank_one_warp<DP>(p, s, k_old, -1, xs, 0);
            }
        }
    } else if (warp == 1) {
        f_rank_one_warp<DP>(p, s, k_new, +1, xs, 1);
    } else if (warp < 5) {
        // component a after the removal: its minus-one pieces are new (count n_a - 1)
        if (lane == 0 && remove_now && !expl) sh.fresh[0][warp - 2] = (n_a >= 2.0) ? f_piece(p, warp - 2, n_a - 1.0) : 0.0;
    } else if (warp < 8) {
        // component b after the addition: its current pieces are new (count n_b)
        if (lane == 0) sh.fresh[1][warp - 5] = f_piece(p, warp - 5, n_b);
    } else if (cta0) {
        const int t = tid - 8 * 32, nthr = TF - 8 * 32;
        if (remove_now) f_stats_axpy<DP>(p, s.rc, k_old, xs, -1, false, t, nthr);
        f_stats_axpy<DP>(p, s.rc, k_new, xs, +1, birth, t, nthr);
        if (t == 0) __stcg(p.z_out + i, s.uid_of_slot[k_new]);  // replicas keep reading the sweep's input labels
    }
    __syncthreads();
    if (tid == 0 && remove_now && !expl) {
        double *sc = s.rec + (size_t)Ly::SC * st + k_old;
        const double cnt = sc[F_CNT * st] + 1.0;
        // the old minus-one pieces become the current ones
        f_write_scalars(p, sc, st, sh.n_new[0], sh.lds_new[0], sc[F_TCM * st], sc[F_LFM * st], sc[F_LCM * st],
                        sh.fresh[0][0], sh.fresh[0][1], sh.fresh[0][2], cnt);
        if (cnt >= (double)REFRESH_EVERY) sh.refresh_a = 1;
    }
    if (tid == 32) {
        double *sc = s.rec + (size_t)Ly::SC * st + k_new;
        const double cnt = sc[F_CNT * st] + 1.0;
        // the old current pieces become the minus-one ones
        const double tcm = birth ? 0.0 : sc[F_TC * st], lfm = birth ? 0.0 : sc[F_LF * st], lcm = birth ? 0.0 : sc[F_LC * st];
        f_write_scalars(p, sc, st, sh.n_new[1], sh.lds_new[1], sh.fresh[1][0], sh.fresh[1][1], sh.fresh[1][2], tcm, lfm,
                        lcm, cnt);
        if (cnt >= (double)REFRESH_EVERY) sh.refresh_b = 1;
        sh.moves += 1;
    }
    __syncthreads();
    if (sh.refresh_a || sh.refresh_b) {
        // drift control: rebuild the record(s) from the bit-exact statistics (every CTA; one barrier)
        f_grid_barrier(p.ctl);
        if (warp == 0) {
            bool okf = true;
            if (sh.refresh_a)
                okf = f_exact_record_warp<DP>(p, 0, p.num + (size_t)k_old * DP, p.S + (size_t)k_old * Ly::PP, n_a, nullptr,
                                              s.rc, s.A, s.W, s.mm, s.rec + k_old, st);
            if (okf && sh.refresh_b)
                okf = f_exact_record_warp<DP>(p, 0, p.num + (size_t)k_new * DP, p.S + (size_t)k_new * Ly::PP, n_b, nullptr,
                                              s.rc, s.A, s.W, s.mm, s.rec + k_new, st);
            if (lane == 0) {
                if (!okf) sh.error = -4;
                sh.refreshes += (sh.refresh_a ? 1 : 0) + (sh.refresh_b ? 1 : 0);
            }
        }
        f_grid_barrier(p.ctl);  // as above: statistics stay frozen until every replica has read them
    }
}

// stage `nb` data starting at scan position pos into the buffers, then resolve them in order
template <int DP>
__device__ int f_run(const Params &p, const FSmem<DP> &s, long long pos, int nb) {
    const int tid = threadIdx.x;
    for (int t = tid; t < nb * DP; t += TF) {
        const int jj = t / DP, a = t % DP;
        const long long j = pos + jj;
        const long long i = p.order ? p.order[j] : j;
        s.xb[jj * DP + a] = p.X[(size_t)i * DP + a];
        if (a == 0) {
            s.ib[jj] = i;
            s.uidb[jj] = __ldcg(p.z_uid + i);
            s.ub[jj] = p.u[j];
            s.lpb[jj] = p.log_prior[i];
        }
    }
    __syncthreads();
    int done = 0;
    for (int jj = 0; jj < nb; ++jj) {
        f_step<DP>(p, s, jj);
        if (s.sh->error) break;
        done = jj + 1;
    }
    return done;  // data fully resolved (uniform over the CTA)
}

// ---------------------------------------------------------------------------------------------
// speculative evaluation of the window [pos, pos + win): a warp per datum, lanes over components.
// Scan position j is owned by CTA (j % grid), warp ((j / grid) % NWARP) -- a fixed owner, so the rows of X a
// warp re-evaluates after a mover are already in its SM's L1.
// ---------------------------------------------------------------------------------------------
template <int DP>
__device__ void f_window_eval(const Params &p, const FSmem<DP> &s, long long pos, long long win, int K,
                              unsigned long long *first_slot) {
    using Ly = Lay<DP>;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int st = p.KS;
    const long long G = gridDim.x;
    const long long end = pos + win;
    double *wrow = s.wrow + (size_t)warp * s.WS;
    // first scan position >= pos owned by this warp
    const long long stride = G * NWARP;
    const long long own = (long long)blockIdx.x + G * warp;  // j % stride == own
    long long j = pos - (pos % stride) + own;
    if (j < pos) j += stride;
    double my_margin = 1.0;
    for (; j < end; j += stride) {
        long long known = 0;
        if (lane == 0) known = (long long)__ldcg(first_slot);
        known = __shfl_sync(0xffffffffu, known, 0);
        if (known < j) break;  // an earlier candidate is already known: this datum would be redone
        const long long i = p.order ? p.order[j] : j;
        const int uid = __ldcg(p.z_uid + i);
        bool cand = (uid < 0);
        int k_old = -1;
        if (!cand) {
            k_old = s.slot_of_uid[uid];
            if (s.rec[(size_t)(Ly::SC + F_N) * st + k_old] == 1.0) cand = true;  // the component would die
        }
        if (!cand) {
            double x[DP];
            const double *xr = p.X + (size_t)i * DP;
            if (DP >= 2) {
#pragma unroll
                for (int a = 0; a < DP; a += 2) {
                    const double2 v = *reinterpret_cast<const double2 *>(xr + a);
                    x[a] = v.x; x[a + 1] = v.y;
                }
            } else {
                x[0] = xr[0];
            }
            bool ok = true;
            for (int k = lane; k < K; k += 32) {
                double w;
                if (k == k_old) w = f_weight_own<DP>(s.rec + k, st, x, &ok);
                else w = f_weight_other<DP>(s.rec + k, st, x);
                wrow[k] = w;
            }
            if (lane == 0) wrow[K] = p.log_alpha + p.log_prior[i];
            ok = __all_sync(0xffffffffu, ok);
            __syncwarp();
            if (!ok) {
                cand = true;
            } else {
                double mg;
                bool bad;
                const int k_new = f_warp_draw(wrow, K + 1, p.u[j], &mg, &bad);
                if (k_new != k_old || bad) cand = true;
                else my_margin = fmin(my_margin, mg);
            }
            __syncwarp();
        }
        if (cand) {
            if (lane == 0) atomicMin(first_slot, (unsigned long long)j);
            break;
        }
    }
    if (lane == 0 && my_margin < 1.0) {
        // committed and discarded evaluations alike: a lower bound of the chain's true minimum margin
        atomicMin(&p.ctl->margin_bits, (unsigned long long)__double_as_longlong(my_margin));
    }
}

// ---------------------------------------------------------------------------------------------
// the sweep kernel: cooperative grid, one CTA per SM
// ---------------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(TF, 1) k_fast_sweep(const Params p) {
    extern __shared__ __align__(16) double smem_raw[];
    using Ly = Lay<DP>;
    const FSmem<DP> s = fast_carve<DP>(smem_raw, p);
    FSh &sh = *s.sh;
    Ctl *ctl = p.ctl;
    const int tid = threadIdx.x;
    const bool cta0 = (blockIdx.x == 0);
    const int st = p.KS;

    // ---- prologue: replicate the chain state ----
    if (tid == 0) {
        sh.K = __ldcg(&ctl->K);
        sh.n_free = __ldcg(&ctl->n_free);
        sh.error = 0;
        sh.pos = p.start_pos;
        sh.gap = p.init_gap;
        sh.moves = sh.births = sh.deaths = sh.evals = sh.windows = sh.seq_data = sh.wasted = 0;
        sh.explicit_evals = sh.refreshes = 0;
        const double one = 1.0;
        sh.margin_bits = (unsigned long long)__double_as_longlong(one);
        sh.round = 0;
        sh.mode = (p.engine == 2) ? 1 : ((p.engine == 1) ? 0 : (p.init_gap >= GAP_TO_WIN ? 1 : 0));
        const uint32_t mb = smem_u32(&sh.mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < Ly::PP; e += TF) {
        int a, b;
        decode_row_idx(e, a, b);
        s.rc[e] = (unsigned short)((a << 8) | b);
    }
    for (int t = tid; t < p.K_max; t += TF) {
        s.slot_of_uid[t] = __ldcg(p.slot_of_uid + t);
        s.uid_of_slot[t] = __ldcg(p.uid_of_slot + t);
        s.uid_free[t] = __ldcg(p.uid_free + t);
    }
    for (int e = tid; e < Ly::R; e += TF) s.prior[e] = __ldcg(p.recB_prior + e);
    __syncthreads();
    {
        // records: one TMA bulk copy of the whole element-major table (cp.async.bulk + mbarrier complete_tx)
        // whole table, rounded up to the 16-byte granule (the global buffer and the shared region are padded)
        const uint32_t bytes = (uint32_t)(((size_t)Ly::R * st * sizeof(double) + 15) & ~(size_t)15);
        const uint32_t mb = smem_u32(&sh.mbar);
        if (tid == 0) {
            asm volatile("fence.proxy.async;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
            uint32_t done = 0;
            while (done < bytes) {
                const uint32_t chunk = min(bytes - done, 32768u);
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                        smem_u32(s.rec) + done),
                    "l"((const char *)p.recB + done), "r"(chunk), "r"(mb)
                    : "memory");
                done += chunk;
            }
        }
        uint32_t okw = 0;
        const long long t0 = clock64();
        while (!okw) {
            if (clock64() - t0 > 8000000000LL) __trap();
            asm volatile(
                "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(okw)
                : "r"(mb), "r"(0u)
                : "memory");
        }
    }
    // no replica may still be reading the initial state when CTA 0 starts changing it
    f_grid_barrier(ctl);

    // ---- main loop ----
    while (true) {
        const long long pos = sh.pos;
        if (pos >= p.N || sh.error != 0) break;
        const int mode = sh.mode;
        __syncthreads();
        if (mode == 0) {
            const int nb = (int)min((long long)SEQ_BATCH, p.N - pos);
            const long long moves0 = sh.moves;
            const int done = f_run<DP>(p, s, pos, nb);
            __syncthreads();
            if (tid == 0) {
                const long long mv = sh.moves - moves0;
                sh.gap = 0.5 * sh.gap + 0.5 * (double)nb / ((double)mv + 0.5);
                sh.seq_data += done;
                sh.pos = pos + done;
                if (p.engine == 0 && sh.gap >= GAP_TO_WIN) sh.mode = 1;
            }
        } else {
            const unsigned int r = sh.round;
            unsigned long long *slot = &ctl->first3[r % 3u];
            if (cta0 && tid == 0) __stcg(&ctl->first3[(r + 1u) % 3u], (unsigned long long)POS_INF);
            const int K = sh.K;
            const long long wcap = (long long)gridDim.x * NWARP * WIN_PASSES_MAX;
            long long win = (long long)fmin(fmax(3.0 * sh.gap, (double)gridDim.x), (double)wcap);
            if (win > p.N - pos) win = p.N - pos;
            f_window_eval<DP>(p, s, pos, win, K, slot);
            f_grid_barrier(ctl);
            const long long f = (long long)__ldcg(slot);
            const long long end = pos + win;
            if (f < end) {
                if (tid == 0) { sh.evals += (f - pos) * (long long)K; sh.wasted += end - (f + 1); }
                const int done = f_run<DP>(p, s, f, 1);
                __syncthreads();
                if (tid == 0) {
                    sh.pos = f + done;
                    sh.gap = 0.7 * sh.gap + 0.3 * (double)(f - pos + 1);
                }
            } else if (tid == 0) {
                sh.evals += win * (long long)K;
                sh.pos = end;
                sh.gap = fmax(sh.gap, 0.7 * sh.gap + 0.3 * 2.0 * (double)win);
            }
            if (tid == 0) {
                sh.windows += 1;
                sh.round = r + 1u;
                if (p.engine == 0 && sh.gap < GAP_TO_SEQ) sh.mode = 0;
            }
        }
        __syncthreads();
    }

    // ---- epilogue: CTA 0 publishes the chain state ----
    __syncthreads();
    if (cta0) {
        const int K = sh.K;
        for (int t = tid; t < p.K_max; t += TF) {
            __stcg(p.slot_of_uid + t, s.slot_of_uid[t]);
            __stcg(p.uid_of_slot + t, s.uid_of_slot[t]);
            __stcg(p.uid_free + t, s.uid_free[t]);
            __stcg(p.counts + t, t < K ? (long long)s.rec[(size_t)(Ly::SC + F_N) * st + t] : 0LL);
        }
        if (tid == 0) {
            __stcg(&ctl->K, K);
            __stcg(&ctl->n_free, sh.n_free);
            __stcg(&ctl->pos, sh.pos);
            __stcg(&ctl->error, sh.error);
            __stcg(&ctl->moves, __ldcg(&ctl->moves) + sh.moves);
            __stcg(&ctl->births, __ldcg(&ctl->births) + sh.births);
            __stcg(&ctl->deaths, __ldcg(&ctl->deaths) + sh.deaths);
            __stcg(&ctl->evals, __ldcg(&ctl->evals) + sh.evals);
            __stcg(&ctl->windows, __ldcg(&ctl->windows) + sh.windows);
            __stcg(&ctl->seq_data, __ldcg(&ctl->seq_data) + sh.seq_data);
            __stcg(&ctl->wasted, __ldcg(&ctl->wasted) + sh.wasted);
            __stcg(&ctl->explicit_evals, __ldcg(&ctl->explicit_evals) + sh.explicit_evals);
            __stcg(&ctl->refreshes, __ldcg(&ctl->refreshes) + sh.refreshes);
            atomicMin(&ctl->margin_bits, sh.margin_bits);
            __stcg(&ctl->gap, sh.gap);
        }
    }
}

// records of all live components (and the prior) in the engine's format, from the bit-exact statistics.
// grid = K + 1 blocks of one warp; block K builds the prior's record.
template <int DP> __global__ void k_fast_prep(const Params p, int K, int *err) {
    using Ly = Lay<DP>;
    __shared__ double A[Ly::PP], W[DP * DP], mm[DP];
    __shared__ unsigned short rc[Ly::PP];
    const int k = blockIdx.x;
    for (int e = threadIdx.x; e < Ly::PP; e += 32) {
        int a, b;
        decode_row_idx(e, a, b);
        rc[e] = (unsigned short)((a << 8) | b);
    }
    __syncwarp();
    bool ok;
    if (k < K) {
        ok = f_exact_record_warp<DP>(p, 0, p.num + (size_t)k * DP, p.S + (size_t)k * Ly::PP, (double)p.counts[k], nullptr, rc,
                                     A, W, mm, p.recB + k, p.KS);
    } else {
        ok = f_exact_record_warp<DP>(p, 1, nullptr, nullptr, 0.0, nullptr, rc, A, W, mm, p.recB_prior, 1);
    }
    if (!ok && threadIdx.x == 0) *err = -4;
}

}  // namespace fast
}  // namespace bgmm
