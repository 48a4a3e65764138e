// bgmm_fast.cuh -- the B200 sweep engine for full covariance (NIW) components with padded D <= 16.
//
// Semantics reproduced: the per-datum loop of CRPMM.collapsed_gibbs_sampler (igmm/crpmm.py:57-88) and
// PCRPMM.collapsed_gibbs_sampler (igmm/pcrpmm.py:93-131), strictly sequential over the scan order.
//
// Execution model (DESIGN.md "Engine"): a REPLICATED STATE MACHINE.  Every CTA of a cooperative grid
// (one per SM) keeps the evaluation records of all live components in its own shared memory and applies
// every state change itself, with identical arithmetic, so the replicas never exchange records.
//   * evaluation record of component k (element-major in shared memory, rec[e * KS + k]):
//       B = S_N^-1 (packed lower triangle), the mean m_N, and the scalars of the Student-t log pdf.
//     A datum joining / leaving a component is a rank-one change of S_N (gaussian_components.py:161-166,
//     :184-185), so B follows by Sherman-Morrison in O(D^2) and log|S_N| by the matrix determinant lemma;
//     records are rebuilt from the bit-exact sufficient statistics at the start of every sweep and after
//     REFRESH_EVERY rank-one updates of a component (drift control).
//   * window round: the CTAs split a window of upcoming scan positions (a warp per datum, lanes over
//     components); each datum is evaluated against the current records.  A datum whose draw keeps it where
//     it is leaves the state untouched (the reference's restore path, crpmm.py:82-85), so every "stay"
//     in front of the first datum that does anything else is exactly the sequential chain's decision.
//     One atomicMin + ONE grid barrier publish that first position (and the warp that holds its inputs);
//     every CTA then resolves it itself.  Warps keep the inputs of the datum they own across rounds.
//   * sequential batch: when movers are dense, every CTA walks the scan datum by datum (no barriers); one step
//     puts the whole CTA on the datum (four threads per component, CTA-wide draw: f_step).
//   * CTA 0 is the only writer of global state: labels (into the sweep's output copy), the bit-exact
//     statistics (same operation order as the reference: one rounded multiply and one rounded add per
//     element), counters.  Replicas read mutable global state only between two grid barriers.
#pragma once
#include "bgmm_sweep.cuh"
#include "bgmm_fastmath.cuh"

namespace bgmm {
namespace fast {

constexpr int TF = 512;               // threads per CTA
constexpr int NWARP = TF / 32;
constexpr int SEQ_BATCH = 16;         // data staged per sequential batch
constexpr int REFRESH_EVERY = 1024;   // rank-one updates of a record before it is rebuilt from the statistics
constexpr int NB_MAX = 6;             // weights per lane in the draw: supports K + 1 <= 192
constexpr int E_NEED_GENERIC = 1;     // internal: a birth would exceed the resident capacity -> generic engine
// a window round costs about as much as two sequential steps, so windows pay from a gap of ~2 data between movers
constexpr double GAP_TO_WIN = 2.5, GAP_TO_SEQ = 1.7;   // (general step; with the register step of bgmm_seq.cuh: Params::gap_to_win / gap_to_seq)
constexpr int WIN_PASSES_MAX = 8;
constexpr int BULK_PASSES_MAX = 4;     // passes of the thread-per-datum evaluator in one window
constexpr int BULK_MIN_ROWS = TF / 2;   // windows of at least this many data per SM use it: a pass of the
                                        // thread-per-datum evaluator has a long fixed latency (K evaluations in series)
constexpr int DLOG = 32;              // versions of the dirty log (power of two)
constexpr int PREP_ZONE = 6;          // rows are prepared from PREP_ZONE windows ahead of the chain (a quiet round can double the window)
constexpr int PATCH_LAG = 10;         // a complete row that waits is brought up to date once it is this many versions behind
constexpr int MV_EXTRA = 5;           // mover slot: x[DP], u, log prior, i, uid, drawn component (-1: full step)
// phase clocks (CTA 0, thread 0; cycles): reported through bgmm_sweep_stats.phase_cycles
enum { PH_STAGE = 0, PH_HEAD, PH_EVAL, PH_DRAW, PH_UPDATE, PH_SCALARS, PH_WINEVAL, PH_BARRIER, PH_RARE, PH_STEPS, PH_MOVES,
       PH_ROUNDS, PH_COUNT = 16 };
#ifdef BGMM_PROFILE
#define F_PROF(slot)                                                  \
    do {                                                              \
        if (threadIdx.x == 0 && blockIdx.x == 0) {                    \
            const long long t_ = clock64();                           \
            s.sh->prof[slot] += t_ - s.sh->prof_last;                 \
            s.sh->prof_last = t_;                                     \
        }                                                             \
    } while (0)
#define F_COUNT(slot) do { if (threadIdx.x == 0 && blockIdx.x == 0) s.sh->prof[slot] += 1; } while (0)
// evaluator unit categories: 0 idle, 1 load inputs, 2 evaluate a chunk, 3 keep a waiting row fresh, 4 decide: stay,
// 5 decide: candidate, 6 decide after a full evaluation
#define F_WCAT(cat) wcat_ = (cat)
#define F_WPROF_BEGIN() int wcat_ = 0; const long long wt0_ = clock64()
#define F_WSUB_BEGIN() const long long ws0_ = clock64()
#define F_WSUB_END(cat)                                                                       \
    do {                                                                                      \
        if ((threadIdx.x & 31) == 0) {                                                        \
            const unsigned long long dt_ = (unsigned long long)(clock64() - ws0_);            \
            atomicAdd(&p.ctl->wsum[cat], dt_);                                                \
            atomicAdd(&p.ctl->wcnt[cat], 1ULL);                                               \
            atomicMax(&p.ctl->wmax[cat], dt_);                                                \
        }                                                                                     \
    } while (0)
#define F_WPROF_END()                                                                         \
    do {                                                                                      \
        if ((threadIdx.x & 31) == 0) {                                                        \
            const unsigned long long dt_ = (unsigned long long)(clock64() - wt0_);            \
            atomicAdd(&p.ctl->wsum[wcat_], dt_);                                              \
            atomicAdd(&p.ctl->wcnt[wcat_], 1ULL);                                             \
            atomicMax(&p.ctl->wmax[wcat_], dt_);                                              \
        }                                                                                     \
    } while (0)
#else
#define F_WCAT(cat) do { } while (0)
#define F_WPROF_BEGIN() do { } while (0)
#define F_WPROF_END() do { } while (0)
#define F_WSUB_BEGIN() do { } while (0)
#define F_WSUB_END(cat) do { } while (0)
#define F_PROF(slot) do { } while (0)
#define F_COUNT(slot) do { } while (0)
#endif

// Everything in a record that depends on the component's count n alone comes from the count table
// Params::ntab (8 doubles per n, built once per chain / per power by k_fast_ntab):
//   NT_CN    LC(n) + TC(n) - D/2 LF(n), with
//              TC  Student-t constant of nu = v0 + n - D + 1 (gaussian_components.py:237-249 without logdet)
//              LF  log f(n), f = (kappa + 1) / (kappa nu)                      (:324-329)
//              LC  log count prior of n (crpmm.py:70 / pcrpmm.py:107)
//   NT_G     kappa / (kappa + 1) = 1 / (f nu)
//   NT_H     (nu + D) / 2
//   NT_BETA  kappa / (kappa - 1)
//   NT_RK    1 / kappa
enum { NT_CN = 0, NT_G, NT_H, NT_BETA, NT_RK, NT_W = 8 };
// record scalars
enum {
    F_N = 0,   // count n (as double)
    F_LDS,     // log|S_N|
    F_CNT,     // rank-one updates since the record was rebuilt
    F_CW,      // CN(n) - LDS / 2
    F_G, F_H, F_BETA,
    F_CWO,     // CN(n - 1) - LDS / 2   (the datum's own component without the datum)
    NSC = 8
};

template <int DP> struct Lay {
    static constexpr int PP = DP * (DP + 1) / 2;
    static constexpr int MU = PP;
    static constexpr int SC = PP + DP;
    static constexpr int R = PP + DP + NSC;
    static constexpr int NS = PP + DP;  // statistics per component: S (packed) then num
    // compile-time record stride (odd: conflict-free column writes) and resident capacity
    static constexpr int KS = (DP == 16) ? 145 : 191;
    static constexpr int KCAP = KS - 1;
    static constexpr int WS = (KS + 1 + 3) & ~3;
};

// grid barrier with a watchdog: replicas that stopped agreeing would otherwise spin forever.  A solo chain (one CTA is
// the chain's only replica, bgmm_sweep_many) needs only its own writes to be complete: CTA barrier + fence.
static __device__ __noinline__ void f_grid_barrier(const Params &p) {
    Ctl *c = p.ctl;
    __syncthreads();
    if (p.solo) {
        if (threadIdx.x == 0) __threadfence();
        __syncthreads();
        return;
    }
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
            while (ld_acquire_u32(&c->bar_gen) == gen) wd.poll(c, 1);
        }
        __threadfence();
    }
    __syncthreads();
}

struct FSh {
    unsigned long long fv;   // the round's first-candidate word, read once per CTA (f_round_barrier)
    long long win;           // length of the next window (set by thread 0 with the state update)
    // replicated chain state
    int K, n_free, error, mode;
    long long pos;
    double gap;
    long long moves, births, deaths, evals, windows, seq_data, wasted, explicit_evals, refreshes;
    long long guard_hits, fast_steps, seq_moves0;
    double last_mg;          // margin of the draw just made (f_step)
    unsigned long long margin_bits;
    // datum being resolved
    int k_new, need_explicit, explicit_done, refresh_a, refresh_b;
    int rare_seq;   // bgmm_seq.cuh: step sequence number of a datum that must go through the general step
    // record version: bumped by every change of the records; the change from ver - 1 to ver touched only the
    // components logged in dlog_a / dlog_b (-1: none) unless it is <= dall_ver
    int ver, dall_ver;               // dall_ver: last version whose change was not confined to two components
    int dlog_a[DLOG], dlog_b[DLOG];  // components touched by the change into version v, at index v % DLOG
    unsigned int round;
    unsigned long long mbar;
    // CTA-wide draw of f_step: inclusive totals, first hit and its margin per warp (4 warps x 32 choices)
    double wtot[4], wmg[4];
    int wcand[4];
    long long prof[PH_COUNT], prof_last;
    long long tprof[4][16];   // bgmm_seq.cuh timeline (profile builds)
};

// The per-round barrier.  Arrival is one acq_rel atomic (its release orders this CTA's candidate publication, made
// visible to thread 0 by the __syncthreads before it); the last CTA to arrive reads the round's candidate word --
// every atomicMin on it happened before some arrival it has now acquired -- and releases
//     (round tag << 44) | candidate word
// in ONE 64-bit word, so a waiting CTA learns "everyone arrived" and "who is first" from the same load.  Critical
// path: atomic (1 L2 trip) + slot read by the last arriver (1) + visibility of the word (1); no standalone fences.
// `tag` is the index of this barrier within the sweep + 1 (20 bits, consecutive tags always differ).
__device__ __forceinline__ unsigned int atom_add_acq_rel_u32(unsigned int *p, unsigned int v) {
    unsigned int r;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
    return r;
}
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
static __device__ __noinline__ void f_round_barrier(Ctl *c, const unsigned long long *slot, unsigned int tag,
                                                    unsigned long long *fv_out) {
    constexpr unsigned long long M44 = (1ULL << 44) - 1;
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned long long want = (unsigned long long)(tag & 0xfffffu);
        const unsigned int prev = atom_add_acq_rel_u32(&c->rb_count, 1u);
        unsigned long long res;
        if (prev == gridDim.x - 1) {
            res = __ldcg(slot) & M44;
            c->rb_count = 0;
            st_release_u64(&c->rb_word, (want << 44) | res);
        } else {
            SpinWatch wd;
            unsigned long long w;
            while (((w = ld_acquire_u64(&c->rb_word)) >> 44) != want) wd.poll(c, 4);
            res = w & M44;
        }
        *fv_out = res;
    }
    __syncthreads();
}

template <int DP> struct FSmem {
    double *rec;      // R * KS
    double *prior;    // R
    double *tmprec;   // R  (exact record of the datum's own component with the datum removed)
    double *ew;       // (NWARP + 1) * WS   (per warp: exp(weight - reference) of the K + 1 choices; last row: f_step)
    double *xw;       // NWARP * DP   (per warp: the datum it owns in the window)
    double *xb;       // SEQ_BATCH * DP
    double *ub, *lpb; // SEQ_BATCH
    long long *ib;    // SEQ_BATCH
    int *uidb;        // SEQ_BATCH
    double *dv;       // 2 * DP   (d of the two updated components)
    double *vv;       // 2 * DP   (v = B d)
    double *nt;       // 2 * 2 * NT_W  (count-table rows n2 - 1, n2 of the two updated components)
    double *A;        // PP       (exact refactor scratch)
    double *W;        // DP * DP
    double *mm;       // DP
    int *slot_of_uid, *uid_of_slot, *uid_free;  // K_max each
    unsigned short *rc;  // PP: (a << 8) | b of packed element e
    double *fm;          // fm::TAB_LEN: the log / exp tables (a global-memory table costs an L1/L2 trip per log and exp)
    FSh *sh;
};

template <int DP> __host__ __device__ inline size_t fast_smem_bytes(int K_max) {
    using Ly = Lay<DP>;
    size_t d = (size_t)Ly::R * Ly::KS + 2 + 2 * (size_t)Ly::R + (size_t)(NWARP + 1) * Ly::WS + (size_t)NWARP * DP +
               (size_t)SEQ_BATCH * DP + 2 * SEQ_BATCH + SEQ_BATCH /*ib*/ + SEQ_BATCH / 2 /*uidb*/ + 4 * DP + 4 * NT_W +
               Ly::PP + DP * DP + DP + fm::TAB_LEN;
    size_t b = d * sizeof(double) + 16;
    b += 3 * (size_t)K_max * sizeof(int);
    b += (((size_t)Ly::PP + 7) & ~(size_t)7) * sizeof(unsigned short);
    b += ((sizeof(FSh) + 15) & ~(size_t)15) + 64;
    return (b + 15) & ~(size_t)15;
}

template <int DP> __device__ inline FSmem<DP> fast_carve(double *base, const Params &p) {
    using Ly = Lay<DP>;
    FSmem<DP> s;
    double *q = base;
    s.rec = q; q += (size_t)Ly::R * Ly::KS;
    q = (double *)(((uintptr_t)q + 15) & ~(uintptr_t)15);
    s.prior = q; q += Ly::R;
    s.tmprec = q; q += Ly::R;
    s.ew = q; q += (size_t)(NWARP + 1) * Ly::WS;
    s.xw = q; q += (size_t)NWARP * DP;
    s.xb = q; q += SEQ_BATCH * DP;
    s.ub = q; q += SEQ_BATCH;
    s.lpb = q; q += SEQ_BATCH;
    s.ib = (long long *)q; q += SEQ_BATCH;
    s.uidb = (int *)q; q += SEQ_BATCH / 2;
    s.dv = q; q += 2 * DP;
    s.vv = q; q += 2 * DP;
    s.nt = q; q += 4 * NT_W;
    s.A = q; q += Ly::PP;
    s.W = q; q += DP * DP;
    s.mm = q; q += DP;
    s.fm = q; q += fm::TAB_LEN;
    // the chain state block sits at a compile-time offset (SOff<DP>::SH); only the three tables behind it are
    // sized at run time
    q = (double *)(((uintptr_t)q + 15) & ~(uintptr_t)15);
    s.sh = (FSh *)q;
    unsigned short *r = (unsigned short *)((char *)q + ((sizeof(FSh) + 15) & ~(size_t)15));
    s.rc = r; r += (Ly::PP + 7) & ~7;
    int *t = (int *)r;
    s.slot_of_uid = t; t += p.K_max;
    s.uid_of_slot = t; t += p.K_max;
    s.uid_free = t; t += p.K_max;
    return s;
}

// Compile-time offsets (in doubles, from the start of the dynamic shared array) of the regions fast_carve lays out in
// front of the run-time sized tables: code on a latency-critical path addresses shared memory as smem_raw[OFF + ...]
// (immediate offsets, no pointer registers).  fast_sweep_body checks them against fast_carve once per launch.
template <int DP> struct SOff {
    using Ly = Lay<DP>;
    static constexpr int REC = 0;
    static constexpr int PRIOR = ((Ly::R * Ly::KS) + 1) & ~1;
    static constexpr int TMPREC = PRIOR + Ly::R;
    static constexpr int EW = TMPREC + Ly::R;
    static constexpr int XW = EW + (NWARP + 1) * Ly::WS;
    static constexpr int XB = XW + NWARP * DP;
    static constexpr int UB = XB + SEQ_BATCH * DP;
    static constexpr int LPB = UB + SEQ_BATCH;
    static constexpr int IB = LPB + SEQ_BATCH;
    static constexpr int UIDB = IB + SEQ_BATCH;
    static constexpr int DV = UIDB + SEQ_BATCH / 2;
    static constexpr int VV = DV + 2 * DP;
    static constexpr int NT = VV + 2 * DP;
    static constexpr int A = NT + 4 * NT_W;
    static constexpr int W = A + Ly::PP;
    static constexpr int MM = W + DP * DP;
    static constexpr int FM = MM + DP;
    static constexpr int SH = (FM + fm::TAB_LEN + 1) & ~1;
};

// The FSmem pointers are generic (the struct is filled at run time), so loads through them are generic loads: they
// go through the LSU's local/global path and wait on the long scoreboard.  On latency-critical single-warp paths the
// same address is rebuilt from the kernel's dynamic shared array (the records start it), which the compiler can prove
// to be shared memory: LDS / STS.
template <int DP, class T> __device__ __forceinline__ T *f_sh(const FSmem<DP> &s, T *ptr) {
    extern __shared__ __align__(16) double smem_raw[];
    return reinterpret_cast<T *>(reinterpret_cast<char *>(smem_raw) +
                                 (reinterpret_cast<const char *>(ptr) - reinterpret_cast<const char *>(s.rec)));
}

// ---------------------------------------------------------------------------------------------
// count table: row n of Params::ntab (see the NT_* enum).  One thread per count; `with_fixed` also writes the
// entries that do not depend on the power.
// ---------------------------------------------------------------------------------------------
static __global__ void k_fast_ntab(const Params p, long long rows, int with_fixed) {
    const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= rows) return;
    const int D = p.D;
    const double nn = (double)n;
    const double kap = p.k0 + nn;
    const long long nu = p.v0 - D + 1 + n;
    double *row = p.ntab + (size_t)n * NT_W;
    // Student-t constant for integer nu (gaussian_components.py:237-249): lgamma((nu+D)/2) - lgamma(nu/2)
    // - D/2 log(nu) - D/2 log(pi), tables indexed like the reference's (index = the integer itself)
    const double tc = p.lgam[nu + D] - p.lgam[nu] - D / 2. * p.logv[nu] - D / 2. * p.log_pi;
    const double lf = log((kap + 1.) / (kap * (double)nu));
    const double lc = n >= 1 ? log_count(nn, p.power) : 0.0;
    row[NT_CN] = lc + tc - 0.5 * (D * lf);
    if (with_fixed) {
        row[NT_G] = kap / (kap + 1.);
        row[NT_H] = ((double)nu + D) / 2.;
        row[NT_BETA] = kap / (kap - 1.);
        row[NT_RK] = 1. / kap;
        row[5] = row[6] = row[7] = 0.0;
    }
}

// scalars of a record with count n and log|S_N| = lds from the count-table rows of n - 1 (r0) and n (r1)
__device__ __forceinline__ void f_write_scalars(double *sc, int st, double n, double lds, double cnt, const double *r0,
                                                const double *r1) {
    sc[(size_t)F_N * st] = n;
    sc[(size_t)F_LDS * st] = lds;
    sc[(size_t)F_CNT * st] = cnt;
    sc[(size_t)F_CW * st] = r1[NT_CN] - 0.5 * lds;
    sc[(size_t)F_G * st] = r1[NT_G];
    sc[(size_t)F_H * st] = r1[NT_H];
    sc[(size_t)F_BETA * st] = r1[NT_BETA];
    sc[(size_t)F_CWO * st] = r0[NT_CN] - 0.5 * lds;
}

// ---------------------------------------------------------------------------------------------
// evaluation of one (datum, component) pair by one thread: exp(weight - wref), or weight - wref when want_log.
//   other component: weight = log count prior + log_post_pred (crpmm.py:70-72)
//   own component  : the same with the datum removed (del_item then log_post_pred, gaussian_components.py:171-186,
//                    :228-251) in closed form: S_N' = S_N - beta d d^T, so |S_N'| = |S_N| om with
//                    om = 1 - beta d^T B d, and 1 + q'/nu' = 1 / om.  NaN when om is too small to trust.
// col points at element 0 of the component (compile-time stride ST); x in shared memory.
// ---------------------------------------------------------------------------------------------
template <int DP, int ST>
__device__ __noinline__ double f_eval_lane(const double *__restrict__ col, const double *__restrict__ x, int own,
                                           double wref, int want_log, const double *__restrict__ fmtab) {
    using Ly = Lay<DP>;
    // note: operands stay generic pointers on purpose.  Passing shared-window offsets (true LDS) was measured 15-40%
    // slower: the compiler then schedules each load right before its use instead of hoisting the batch.
    double d[DP];
#pragma unroll
    for (int a = 0; a < DP; ++a) d[a] = col[(Ly::MU + a) * ST] - x[a];
    double q = 0.0;
#pragma unroll
    for (int a = 0; a < DP; ++a) {
        double r = 0.0;
#pragma unroll
        for (int b = 0; b < a; ++b) r = fma(col[(a * (a + 1) / 2 + b) * ST], d[b], r);
        r = fma(0.5 * col[(a * (a + 1) / 2 + a) * ST], d[a], r);
        q = fma(d[a], r, q);
    }
    q *= 2.0;
    const double *sc = col + Ly::SC * ST;
    double arg, hh, cc;
    if (own) {
        arg = 1.0 - sc[F_BETA * ST] * q;
        if (!(arg > OM_MIN)) return NAN;
        hh = 1.0 - sc[F_H * ST];
        cc = sc[F_CWO * ST];
    } else {
        arg = 1.0 + sc[F_G * ST] * q;
        hh = sc[F_H * ST];
        cc = sc[F_CW * ST];
    }
    const double t = (cc - hh * fm::f_log(arg, fmtab)) - wref;
    if (want_log) return t;
    return (t < EXP_CUTOFF) ? 0.0 : fm::f_exp(t, fmtab);
}

// ---------------------------------------------------------------------------------------------
// The same quantity for TWO components at once, rows of B over the 16 lanes of each half-warp (half 0: component
// ka, half 1: kb; -1 = none): the latency-critical form used to bring a cached row up to date after ONE move --
// 16-term dot products and a 4-step shuffle reduction instead of a 136-term chain per lane.  Same formula as
// f_eval_lane, different summation order (agreement to rounding, ~1e-16).  Whole warp must call.
// ---------------------------------------------------------------------------------------------
template <int DP>
__device__ __forceinline__ void f_eval_rows2(const double *__restrict__ rec, const double *__restrict__ x, int ka, int kb,
                                          int k_old, double wref, double *__restrict__ ew,
                                          const double *__restrict__ fmtab) {
    using Ly = Lay<DP>;
    constexpr int ST = Ly::KS;
    const int lane = threadIdx.x & 31;
    const int r = lane & 15, half = lane >> 4;
    const int k = half ? kb : ka;
    const bool act = (k >= 0) && (r < DP);
    const double *col = rec + (k >= 0 ? k : 0);
    const double dr = act ? col[(Ly::MU + r) * ST] - x[r] : 0.0;
    const double *row = col + (size_t)(r * (r + 1) / 2) * ST;   // element (r, 0)
    double a0 = 0.0, a1 = 0.0;
#pragma unroll
    for (int b = 0; b < DP; b += 2) {
        const double d0 = __shfl_sync(0xffffffffu, dr, (half << 4) + b, 32);
        const double d1 = __shfl_sync(0xffffffffu, dr, (half << 4) + ((b + 1) & 15), 32);
        if (act && b < r) a0 = fma(row[(size_t)b * ST], d0, a0);
        if (act && b + 1 < r) a1 = fma(row[(size_t)(b + 1) * ST], d1, a1);
    }
    double t = act ? dr * fma(0.5 * row[(size_t)r * ST], dr, a0 + a1) : 0.0;
#pragma unroll
    for (int o = 8; o >= 1; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (r == 0 && k >= 0) {
        const double q = 2.0 * t;
        const double *sc = col + Ly::SC * ST;
        const bool own = (k == k_old);
        const double arg = own ? 1.0 - sc[F_BETA * ST] * q : 1.0 + sc[F_G * ST] * q;
        double e;
        if (own && !(arg > OM_MIN)) {
            e = NAN;
        } else {
            const double hh = own ? 1.0 - sc[F_H * ST] : sc[F_H * ST];
            const double cc = own ? sc[F_CWO * ST] : sc[F_CW * ST];
            const double tt = (cc - hh * fm::f_log(arg, fmtab)) - wref;
            e = (tt < EXP_CUTOFF) ? 0.0 : fm::f_exp(tt, fmtab);
        }
        ew[k] = e;
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// f_step's evaluation for D = 16 and K + 1 <= 128: the quadratic form of ONE component is split over four threads that
// sit in four different warps (warp-uniform PART, lanes over components: no divergence, conflict-free loads), so all
// 16 warps work on the one datum of a sequential step instead of 4.  d^T B d over the packed lower triangle, rows and
// columns split in halves L = 0..7, H = 8..15:
//   PART 0: block LL (36 terms)   PART 1: block HH (36)   PART 2: rows 8..11 x L (32)   PART 3: rows 12..15 x L (32)
// returns this part of  sum_a d_a (sum_{b<a} B_ab d_b + B_aa d_a / 2);  q = 2 * (sum of the four).  Same formula as
// f_eval_lane, different summation order (agreement to rounding).
// ---------------------------------------------------------------------------------------------
template <int PART, int ST>
__device__ __forceinline__ double f_quad_part16(const double *__restrict__ col, const double *__restrict__ x) {
    using Ly = Lay<16>;
    double q = 0.0;
    if constexpr (PART <= 1) {
        constexpr int O = PART * 8;
        double d[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) d[a] = col[(Ly::MU + O + a) * ST] - x[O + a];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            const int ra = (O + a) * (O + a + 1) / 2 + O;   // packed index of element (O + a, O)
            double r = 0.0;
#pragma unroll
            for (int b = 0; b < a; ++b) r = fma(col[(ra + b) * ST], d[b], r);
            r = fma(0.5 * col[(ra + a) * ST], d[a], r);
            q = fma(d[a], r, q);
        }
    } else {
        constexpr int O = 8 + (PART - 2) * 4;
        double dl[8], dh[4];
#pragma unroll
        for (int b = 0; b < 8; ++b) dl[b] = col[(Ly::MU + b) * ST] - x[b];
#pragma unroll
        for (int a = 0; a < 4; ++a) dh[a] = col[(Ly::MU + O + a) * ST] - x[O + a];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int ra = (O + a) * (O + a + 1) / 2;       // packed index of element (O + a, 0)
            double r = 0.0;
#pragma unroll
            for (int b = 0; b < 8; ++b) r = fma(col[(ra + b) * ST], dl[b], r);
            q = fma(dh[a], r, q);
        }
    }
    return q;
}

// exp(weight - wref) from the quadratic form q (the tail of f_eval_lane); NaN when the closed form of the own
// component is not trusted
template <int ST>
__device__ __forceinline__ double f_finish_weight(const double *__restrict__ sc, double q, int own, double wref,
                                                  const double *__restrict__ fmtab) {
    double arg, hh, cc;
    if (own) {
        arg = 1.0 - sc[F_BETA * ST] * q;
        if (!(arg > OM_MIN)) return NAN;
        hh = 1.0 - sc[F_H * ST];
        cc = sc[F_CWO * ST];
    } else {
        arg = 1.0 + sc[F_G * ST] * q;
        hh = sc[F_H * ST];
        cc = sc[F_CW * ST];
    }
    const double t = (cc - hh * fm::f_log(arg, fmtab)) - wref;
    return (t < EXP_CUTOFF) ? 0.0 : fm::f_exp(t, fmtab);
}

// ---------------------------------------------------------------------------------------------
// draw (utils.py:7-20) by one warp from the n = K + 1 unnormalised probabilities e[] in shared memory: the first
// index whose inclusive cumulative sum exceeds u * total, else the last index.  Lane l owns the nb consecutive
// entries starting at l * nb.  Returns the index, or -2 when the total is not finite / positive (the caller falls
// back to the log-domain draw).  *margin: distance of u to the nearest boundary of the drawn interval
// (probability units, float accuracy -- a diagnostic).
// ---------------------------------------------------------------------------------------------
// The lane's (at most NB_MAX) entries are read once into registers: both passes below then run on registers, in the
// same order as a loop over shared memory would (bit-identical sums; trailing zeros add nothing).
static __device__ __forceinline__ int f_warp_pick_inl(const double *__restrict__ e, int n, double u, double *margin) {
    const int lane = threadIdx.x & 31;
    const int nb = (n + 31) >> 5;
    const int lo = lane * nb;
    double v[NB_MAX];
#pragma unroll
    for (int t = 0; t < NB_MAX; ++t) v[t] = (t < nb && lo + t < n) ? e[lo + t] : 0.0;
    double run = 0.0;
#pragma unroll
    for (int t = 0; t < NB_MAX; ++t) run += v[t];
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
    double lower = excl, mg = 0.0, cum = 0.0;
#pragma unroll
    for (int t = 0; t < NB_MAX; ++t) {
        if (t < nb && lo + t < n && cand < 0) {
            cum += v[t];
            const double upper = excl + cum;
            if (upper > t0) { cand = lo + t; mg = fmin(t0 - lower, upper - t0); }
            else lower = upper;
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
        mg = __shfl_sync(0xffffffffu, mg, src);
        mg = margin_ratio(mg, s);
    }
    *margin = mg;
    if (!(s > 0.0) || !(s < INFINITY)) k = -2;
    return k;
}

static __device__ __noinline__ int f_warp_pick(const double *__restrict__ e, int n, double u, double *margin) {
    return f_warp_pick_inl(e, n, u, margin);
}

// log-domain version (crpmm.py:75: exp(w - logsumexp(w))) for weights whose spread overflows the fast scaling:
// w[] holds weight - wref; scaled by (a float rounding of) the maximum.  In place: w[] becomes the exponentials.
static __device__ __noinline__ int f_warp_draw_log(double *__restrict__ w, int n, double u, double *margin) {
    const int lane = threadIdx.x & 31;
    float fm = -INFINITY;
    for (int t = lane; t < n; t += 32) fm = fmaxf(fm, __double2float_rn(w[t]));
    int key = __float_as_int(fm);
    key ^= (key >> 31) & 0x7fffffff;
    key = __reduce_max_sync(0xffffffffu, key);
    key ^= (key >> 31) & 0x7fffffff;
    const double M = (double)__int_as_float(key);
    for (int t = lane; t < n; t += 32) {
        const double dlt = w[t] - M;
        w[t] = (dlt < EXP_CUTOFF) ? 0.0 : exp(dlt);
    }
    __syncwarp();
    return f_warp_pick(w, n, u, margin);
}

// ---------------------------------------------------------------------------------------------
// exact record of one component by ONE WARP from the bit-exact statistics (gaussian_components.py:319-331):
// S_N = S - kappa m m^T, Cholesky, log|S_N|, B = S_N^-1.  mode 0: component statistics in global memory,
// optionally with datum x removed first (the reference's del_item arithmetic, :184-185); mode 1: the prior
// alone (n = 0, :161-164).  out has stride ost.  Returns false (all lanes) if S_N is not positive definite.
// ---------------------------------------------------------------------------------------------
template <int DP>
__device__ __noinline__ bool f_exact_record_warp(const Params &p, int mode, const double *num_g, const double *S_g,
                                                 double n_after, const double *xrm, const unsigned short *rc, double *A,
                                                 double *W, double *mm, double *out, int ost) {
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
    if (lane == 0) {
        const long long n1 = (long long)n_after;
        const double *r1 = p.ntab + (size_t)n1 * NT_W;
        const double *r0 = p.ntab + (size_t)(n1 > 0 ? n1 - 1 : 0) * NT_W;
        f_write_scalars(out + (size_t)Ly::SC * ost, ost, n_after, lds, 0.0, r0, r1);
    }
    __syncwarp();
    return true;
}

// ---------------------------------------------------------------------------------------------
// rank-one update of one record by ONE WARP, including its scalars.  sign = -1: the datum leaves (del_item),
// sign = +1: it joins (add_item).  x in shared memory.  which = 0 (removal) / 1 (addition) selects the scratch.
// ---------------------------------------------------------------------------------------------
template <int DP>
__device__ __noinline__ void f_rank_one_warp(const Params &p, const FSmem<DP> &s, int k, int sign, const double *x,
                                             int which, int seq) {
    using Ly = Lay<DP>;
    constexpr int ST = Ly::KS;
    const int lane = threadIdx.x & 31;
    double *col = f_sh<DP>(s, s.rec) + k;
    double *dv = f_sh<DP>(s, s.dv) + which * DP, *vv = f_sh<DP>(s, s.vv) + which * DP;
    double *nt = f_sh<DP>(s, s.nt) + which * 2 * NT_W;
    const unsigned short *rc = f_sh<DP>(s, s.rc);
    x = f_sh<DP>(s, x);   // the datum sits in the staging buffers (shared memory)
    double *sc = col + Ly::SC * ST;
    const double n = sc[F_N * ST];
    const double n2 = n + (double)sign;
    // count-table rows n2 - 1 and n2 (adjacent: 2 * NT_W doubles), in flight while the matrix work runs
    double ntv = 0.0;
    if (lane < 2 * NT_W) ntv = __ldg(p.ntab + (size_t)((long long)n2 - 1) * NT_W + lane);
    // d = x - m, v = B d: lane (part, r) sums its share of row r, parts combined by shuffles
    const int r = lane % DP, part = lane / DP;
    constexpr int NH = 32 / DP;
    constexpr int BS = (DP + NH - 1) / NH;
    const double dr = x[r] - col[(Ly::MU + r) * ST];
    if (part == 0) dv[r] = dr;
    __syncwarp();
    double acc = 0.0;
#pragma unroll
    for (int t = 0; t < BS; ++t) {
        const int b = part * BS + t;
        if (b < DP) {
            const int e = (r >= b) ? row_idx(r, b) : row_idx(b, r);
            acc = fma(col[e * ST], dv[b], acc);
        }
    }
#pragma unroll
    for (int o = DP; o < 32; o <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (part == 0) vv[r] = acc;
    double sq = dr * acc;
#pragma unroll
    for (int o = DP / 2; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    __syncwarp();
    // beta = kappa / (kappa -+ 1): BETA(n) for a removal, G(n) for an addition
    const double beta = (sign < 0) ? sc[F_BETA * ST] : sc[F_G * ST];
    const double den = (sign < 0) ? 1.0 - beta * sq : 1.0 + beta * sq;
    const double gam = (sign < 0) ? beta / den : -beta / den;
    const double lg_den = fm::f_log(den, p.fmtab);   // den is om in (1/64, 1] or 1 + beta q >= 1
#pragma unroll
    for (int e0 = 0; e0 < Ly::PP; e0 += 32) {
        const int e = e0 + lane;
        if (e < Ly::PP) {
            const int a = rc[e] >> 8, b = rc[e] & 0xff;
            double *pe = col + e * ST;
            *pe = fma(gam * vv[a], vv[b], *pe);
        }
    }
    // the count-table rows (an L2 round trip issued at the top) are first needed here
    if (lane < 2 * NT_W) nt[lane] = ntv;
    __syncwarp();
    const double rk2 = nt[NT_W + NT_RK];   // 1 / kappa(n2)
    if (part == 0) {
        double *pm = col + (Ly::MU + r) * ST;
        *pm = fma((sign < 0) ? -dr : dr, rk2, *pm);
    }
    if (lane == 0) {
        const double cnt = sc[F_CNT * ST] + 1.0;
        f_write_scalars(sc, ST, n2, sc[F_LDS * ST] + lg_den, cnt, nt, nt + NT_W);
        FSh &sh = *f_sh<DP>(s, s.sh);
        if (cnt >= (double)REFRESH_EVERY) { if (which == 0) sh.refresh_a = seq; else sh.refresh_b = seq; }
        if (which == 1) sh.moves += 1;
    }
    __syncwarp();
}

// bit-exact statistics update in global memory by `nthr` threads of CTA 0 (thread rank t):
// S -= / += fl(x_a x_b), num -= / += x_a   (gaussian_components.py:165-166, :184-185)
// The addition is a fire-and-forget reduction performed by the L2 (RED.ADD.F64: one IEEE round-to-nearest add, the
// same bits as __dadd_rn; v - o == v + (-o)), so the move phase of CTA 0 does not wait for an L2 round trip per
// element.  Updates of one address are issued by different threads in different moves; they are ordered by the CTA
// barriers between moves (coherence order of same-address atomics follows happens-before), and every reader of the
// statistics sits behind a grid barrier or a CTA barrier.
__device__ __forceinline__ void red_add_f64(double *addr, double v) {
    asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
template <int DP>
__device__ __forceinline__ void f_stats_axpy_inl(const Params &p, const unsigned short *rc, int slot, const double *x,
                                                 int sign, int init_prior, int t, int nthr) {
    using Ly = Lay<DP>;
    const int D = p.D;
    double *S = p.S + (size_t)slot * Ly::PP;
    double *num = p.num + (size_t)slot * DP;
#pragma unroll 2
    for (int e = t; e < Ly::NS; e += nthr) {
        if (e < Ly::PP) {
            const int a = rc[e] >> 8, b = rc[e] & 0xff;
            if (a >= D) { if (init_prior) __stcg(S + e, 0.0); continue; }
            const double o = __dmul_rn(x[a], x[b]);
            if (init_prior) {
                const double v = __dadd_rn(p.S0[e], __dmul_rn(p.k0, __dmul_rn(p.m0[a], p.m0[b])));
                __stcg(S + e, sign > 0 ? __dadd_rn(v, o) : __dsub_rn(v, o));
            } else if (p.tune & 2) {
                const double v = __ldcg(S + e);
                __stcg(S + e, sign > 0 ? __dadd_rn(v, o) : __dsub_rn(v, o));
            } else {
                red_add_f64(S + e, sign > 0 ? o : -o);
            }
        } else {
            const int a = e - Ly::PP;
            if (a >= D) { if (init_prior) __stcg(num + a, 0.0); continue; }
            if (init_prior) {
                const double v = __dmul_rn(p.k0, p.m0[a]);
                __stcg(num + a, sign > 0 ? __dadd_rn(v, x[a]) : __dsub_rn(v, x[a]));
            } else if (p.tune & 2) {
                const double v = __ldcg(num + a);
                __stcg(num + a, sign > 0 ? __dadd_rn(v, x[a]) : __dsub_rn(v, x[a]));
            } else {
                red_add_f64(num + a, sign > 0 ? x[a] : -x[a]);
            }
        }
    }
}

template <int DP>
__device__ __noinline__ void f_stats_axpy(const Params &p, const unsigned short *rc, int slot, const double *x, int sign,
                                          int init_prior, int t, int nthr) {
    f_stats_axpy_inl<DP>(p, rc, slot, x, sign, init_prior, t, nthr);
}

// ---------------------------------------------------------------------------------------------
// rare paths of a step, out of line
// ---------------------------------------------------------------------------------------------
// del_item empties the component: del_component (gaussian_components.py:188-205), swap with the last slot
template <int DP> __device__ __noinline__ void f_delete_component(const Params &p, const FSmem<DP> &s, int k_old) {
    using Ly = Lay<DP>;
    constexpr int ST = Ly::KS;
    FSh &sh = *s.sh;
    const int tid = threadIdx.x;
    const int L = sh.K - 1;
    __syncthreads();
    if (k_old != L) {
        for (int e = tid; e < Ly::R; e += TF) s.rec[(size_t)e * ST + k_old] = s.rec[(size_t)e * ST + L];
        if (p.writer) {
            for (int e = tid; e < Ly::PP; e += TF)
                __stcg(p.S + (size_t)k_old * Ly::PP + e, __ldcg(p.S + (size_t)L * Ly::PP + e));
            for (int e = tid; e < DP; e += TF) __stcg(p.num + (size_t)k_old * DP + e, __ldcg(p.num + (size_t)L * DP + e));
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
    }
    __syncthreads();
}

// the closed form of the own-component weight is not trusted: build the reduced component's record exactly from
// the statistics (every CTA; CTA 0's statistics writes are made visible by the first barrier, and stay frozen until
// every replica has read them by the second).  Leaves the entry of k_old in ew[].
template <int DP>
__device__ __noinline__ void f_explicit_own(const Params &p, const FSmem<DP> &s, int k_old, double n_old,
                                            const double *xs, double wref, double *ew, int seq) {
    using Ly = Lay<DP>;
    FSh &sh = *s.sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    f_grid_barrier(p);
    if (warp == 0) {
        const bool okf = f_exact_record_warp<DP>(p, 0, p.num + (size_t)k_old * DP, p.S + (size_t)k_old * Ly::PP,
                                                 n_old - 1.0, xs, s.rc, s.A, s.W, s.mm, s.tmprec, 1);
        if (lane == 0) {
            if (!okf) {
                sh.error = -4;
            } else {
                ew[k_old] = f_eval_lane<DP, 1>(s.tmprec, xs, 0, wref, 0, p.fmtab);
                sh.explicit_done = seq;
                sh.explicit_evals += 1;
            }
        }
    }
    f_grid_barrier(p);
}

// the weights' spread overflowed the exp scale of the fast draw: redo the datum in the log domain
template <int DP>
__device__ __noinline__ void f_log_domain_draw(const Params &p, const FSmem<DP> &s, int K, int k_old, bool own_live,
                                               bool expl, const double *xs, double wref, double u, double *ew) {
    using Ly = Lay<DP>;
    FSh &sh = *s.sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < K) {
        if (expl && tid == k_old) ew[tid] = f_eval_lane<DP, 1>(s.tmprec, xs, 0, wref, 1, p.fmtab);
        else ew[tid] = f_eval_lane<DP, Ly::KS>(s.rec + tid, xs, (own_live && tid == k_old) ? 1 : 0, wref, 1, p.fmtab);
    } else if (tid == K) {
        ew[K] = 0.0;
    }
    __syncthreads();
    if (warp == 0) {
        double mg;
        const int k = f_warp_draw_log(ew, K + 1, u, &mg);
        if (lane == 0) {
            sh.k_new = k;
            sh.last_mg = mg;
            if (k < 0) sh.error = -4;
        }
    }
    __syncthreads();
}


// ---------------------------------------------------------------------------------------------
// Margin guard.  The engine's weights differ from the reference's by rounding (Sherman-Morrison records, table-driven
// log / exp, a different summation order: ~1e-12 in probability units at worst), so a draw is the reference's draw only
// while the uniform keeps a distance from the boundaries of the drawn interval.  When that margin is below
// Params::guard the datum is redone on the exact path (every CTA identically):
//   * every live record is rebuilt from the bit-exact statistics (what the start of a sweep does), the datum's own
//     component with the datum removed by the reference's del_item arithmetic (gaussian_components.py:184-185);
//   * weights in the log domain with libdevice log, normalised like crpmm.py:75 (exp(w - logsumexp(w))), and the
//     reference's sequential-subtract draw (utils.py:15-20) by one thread.
// Leaves the drawn index in sh.k_new and the reduced own component in tmprec (the caller treats it as `expl`).
// ---------------------------------------------------------------------------------------------
template <int DP, int ST>
__device__ __noinline__ double f_weight_exact(const double *__restrict__ col, const double *__restrict__ x) {
    using Ly = Lay<DP>;
    double d[DP];
#pragma unroll
    for (int a = 0; a < DP; ++a) d[a] = col[(Ly::MU + a) * ST] - x[a];
    double q = 0.0;
#pragma unroll
    for (int a = 0; a < DP; ++a) {
        double r = 0.0;
#pragma unroll
        for (int b = 0; b < a; ++b) r = fma(col[(a * (a + 1) / 2 + b) * ST], d[b], r);
        r = fma(0.5 * col[(a * (a + 1) / 2 + a) * ST], d[a], r);
        q = fma(d[a], r, q);
    }
    q *= 2.0;
    const double *sc = col + Ly::SC * ST;
    return sc[F_CW * ST] - sc[F_H * ST] * log(1.0 + sc[F_G * ST] * q);
}

template <int DP>
__device__ __noinline__ void f_exact_redo(const Params &p, const FSmem<DP> &s, int K, int k_old, double n_old,
                                          bool own_live, const double *xs, double wref, double u, double *ew) {
    using Ly = Lay<DP>;
    constexpr int ST = Ly::KS;
    FSh &sh = *s.sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    f_grid_barrier(p);   // CTA 0's statistics are complete and stay frozen until the second barrier
    if (warp == 0) {
        bool okf = true;
        for (int k = 0; k < K && okf; ++k) {
            const double n_k = s.rec[(Ly::SC + F_N) * ST + k];
            okf = f_exact_record_warp<DP>(p, 0, p.num + (size_t)k * DP, p.S + (size_t)k * Ly::PP, n_k, nullptr, s.rc, s.A,
                                          s.W, s.mm, s.rec + k, ST);
        }
        if (okf && own_live)
            okf = f_exact_record_warp<DP>(p, 0, p.num + (size_t)k_old * DP, p.S + (size_t)k_old * Ly::PP, n_old - 1.0, xs,
                                          s.rc, s.A, s.W, s.mm, s.tmprec, 1);
        if (lane == 0) {
            if (!okf) sh.error = -4;
            sh.guard_hits += 1;
            sh.refreshes += K;
            // every record changed (by rounding): cached rows of the window evaluators are void
            sh.ver += 1;
            sh.dall_ver = sh.ver;
        }
    }
    f_grid_barrier(p);
    if (sh.error) return;
    if (tid < K) {
        ew[tid] = ((own_live && tid == k_old) ? f_weight_exact<DP, 1>(s.tmprec, xs) : f_weight_exact<DP, ST>(s.rec + tid, xs)) -
                  wref;
    } else if (tid == K) {
        ew[K] = 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        double M = -INFINITY;
        for (int k = 0; k <= K; ++k) M = fmax(M, ew[k]);
        double ssum = 0.0;
        for (int k = 0; k <= K; ++k) ssum += exp(ew[k] - M);
        const double lse = M + log(ssum);
        double t = u, mg = 0.0;
        int kk = K;   // utils.py:20: the last index when the running value never goes negative
        for (int k = 0; k <= K; ++k) {
            const double tb = t;
            t -= exp(ew[k] - lse);
            if (t < 0.0) { kk = k; mg = fmin(fabs(tb), -t); break; }
        }
        if (!(ssum > 0.0) || !(ssum < INFINITY)) { kk = -2; sh.error = -4; }
        sh.k_new = kk;
        sh.last_mg = mg;
    }
    __syncthreads();
}

// drift control: rebuild record(s) from the bit-exact statistics (every CTA; two barriers as above)
template <int DP>
__device__ __noinline__ void f_refresh(const Params &p, const FSmem<DP> &s, int ka, double n_a, int kb, double n_b) {
    using Ly = Lay<DP>;
    FSh &sh = *s.sh;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    f_grid_barrier(p);
    if (warp == 0) {
        bool okf = true;
        int cnt = 0;
        if (ka >= 0) {
            okf = f_exact_record_warp<DP>(p, 0, p.num + (size_t)ka * DP, p.S + (size_t)ka * Ly::PP, n_a, nullptr, s.rc, s.A,
                                          s.W, s.mm, s.rec + ka, Ly::KS);
            ++cnt;
        }
        if (okf && kb >= 0) {
            okf = f_exact_record_warp<DP>(p, 0, p.num + (size_t)kb * DP, p.S + (size_t)kb * Ly::PP, n_b, nullptr, s.rc, s.A,
                                          s.W, s.mm, s.rec + kb, Ly::KS);
            ++cnt;
        }
        if (lane == 0) {
            if (!okf) sh.error = -4;
            sh.refreshes += cnt;
        }
    }
    f_grid_barrier(p);
}

// a new component opens in slot K, initialised with the prior (gaussian_components.py:161-164).  false: error set
template <int DP> __device__ __noinline__ bool f_open_component(const Params &p, const FSmem<DP> &s, int K) {
    using Ly = Lay<DP>;
    FSh &sh = *s.sh;
    const int tid = threadIdx.x;
    if (K >= p.K_max || K >= Ly::KCAP) {
        // K_max: the reference would raise IndexError.  Resident capacity: nothing has been changed for this datum
        // yet, the generic engine redoes it.
        __syncthreads();
        if (tid == 0) {
            if (K >= p.K_max) sh.error = -3;
            else { sh.error = E_NEED_GENERIC; sh.evals -= K; }
        }
        __syncthreads();
        return false;
    }
    for (int e = tid; e < Ly::R; e += TF) s.rec[(size_t)e * Ly::KS + K] = s.prior[e];
    if (tid == 0) {
        const int nuid = s.uid_free[sh.n_free - 1];
        sh.n_free -= 1;
        s.uid_of_slot[K] = nuid;
        s.slot_of_uid[nuid] = K;
        sh.K = K + 1;
        sh.births += 1;
    }
    __syncthreads();
    return true;
}

// the state change of a move (whole CTA): the datum leaves k_old (if remove_now) and joins k_new
template <int DP>
__device__ __noinline__ void f_move_phase(const Params &p, const FSmem<DP> &s, int jj, int k_old, int k_new,
                                          bool remove_now, bool birth, bool expl, bool died, int seq) {
    using Ly = Lay<DP>;
    constexpr int ST = Ly::KS;
    FSh &sh = *f_sh<DP>(s, s.sh);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *xs = s.xb + jj * DP;
    if (warp == 0) {
        if (remove_now) {
            if (expl) {
                for (int e = lane; e < Ly::R; e += 32) s.rec[(size_t)e * ST + k_old] = s.tmprec[e];
            } else {
                f_rank_one_warp<DP>(p, s, k_old, -1, xs, 0, seq);
            }
        }
    } else if (warp == 1) {
        f_rank_one_warp<DP>(p, s, k_new, +1, xs, 1, seq);
    } else if (warp == 2) {
        if (lane == 0) {
            // what this step changed, for the evaluators' cached rows
            const int v = sh.ver + 1;
            sh.ver = v;
            sh.dlog_a[v & (DLOG - 1)] = remove_now ? k_old : -1;
            sh.dlog_b[v & (DLOG - 1)] = k_new;
            if (died || expl) sh.dall_ver = v;
        }
    }
    __syncthreads();
    if (p.writer && warp >= 8) {
        // The writer CTA: the bit-exact statistics and the label (warps 8..11 the removal, 12..15 the addition), issued
        // behind the move's barrier: a CTA barrier waits for the warp's outstanding global reductions (an L2 round trip
        // behind ~150 of them), and from here the next barrier is an evaluation away.  Reductions of consecutive moves to
        // one address stay ordered by the barriers between them; every reader of the statistics sits behind one.
        if (warp < 12) {
            if (remove_now) f_stats_axpy<DP>(p, s.rc, k_old, xs, -1, 0, tid - 256, 128);
        } else {
            f_stats_axpy<DP>(p, s.rc, k_new, xs, +1, birth ? 1 : 0, tid - 384, 128);
            if (tid == 384) __stcg(p.z_out + s.ib[jj], s.uid_of_slot[k_new]);  // replicas keep reading the input labels
        }
    }
    F_PROF(PH_UPDATE);
    F_COUNT(PH_MOVES);
    const bool ra = (sh.refresh_a == seq), rb = (sh.refresh_b == seq);
    if (ra || rb) {
        const double n_a = ra ? s.rec[(Ly::SC + F_N) * ST + k_old] : 0.0;
        const double n_b = s.rec[(Ly::SC + F_N) * ST + k_new];
        f_refresh<DP>(p, s, ra ? k_old : -1, n_a, rb ? k_new : -1, n_b);
        if (tid == 0) sh.dall_ver = sh.ver;
        __syncthreads();
        F_PROF(PH_RARE);
    }
}

// ---------------------------------------------------------------------------------------------
// Resolve ONE datum (whole CTA, every CTA identically).  Inputs in the staging buffers at index jj; seq is a
// per-step sequence number (> 0) used to tag the rare-path flags so they never need resetting.
// ---------------------------------------------------------------------------------------------
template <int DP>
__device__ __noinline__ void f_step(const Params &p, const FSmem<DP> &s, int jj, int seq) {
    using Ly = Lay<DP>;
    constexpr int ST = Ly::KS;
    FSh &sh = *f_sh<DP>(s, s.sh);   // shared-memory addressing on the step's dependent chain (see f_sh)
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double *xs = s.xb + jj * DP;
    double *ew = f_sh<DP>(s, s.ew) + (size_t)NWARP * Ly::WS;

    // head: every thread derives the same values from the replicated state
    const int uid = f_sh<DP>(s, s.uidb)[jj];
    int k_old = -1;
    double n_old = 0.0;
    if (uid >= 0) {
        k_old = f_sh<DP>(s, s.slot_of_uid)[uid];
        n_old = f_sh<DP>(s, s.rec)[(Ly::SC + F_N) * ST + k_old];
    }
    bool died = false;
    if (k_old >= 0 && n_old == 1.0) {
        f_delete_component<DP>(p, s, k_old);
        died = true;
    }
    const int K = sh.K;
    const double wref = p.log_alpha + f_sh<DP>(s, s.lpb)[jj];   // the new-table weight (crpmm.py:74) is the exp scale
    const double u_draw = f_sh<DP>(s, s.ub)[jj];
    const bool own_live = (k_old >= 0) && !died;

    // phase A: exp(weight - wref) of every live component (crpmm.py:68-75)
    bool cta_draw = false;   // the K + 1 choices sit one per thread in warps 0..3 (uniform over the CTA)
    double e_mine = 0.0, incl = 0.0;
    if constexpr (DP == 16) {
        if (K < 128 && !(p.tune & 1)) {
            // four threads per component, one in each quarter of the CTA (f_quad_part16); partial sums of parts
            // 1..3 through the refactor scratch (A and W are contiguous: 392 doubles, idle outside the rare paths)
            cta_draw = true;
            // the records start the kernel's dynamic shared array: addressing everything through it (not through the
            // generic pointers of FSmem) makes the loads below LDS -- all 16 warps load at once here, and generic
            // loads queue in the LSU's local/global path (see f_bulk_eval)
            extern __shared__ __align__(16) double smem_raw[];
            const double *rec_sh = smem_raw;
            const double *xs_sh = smem_raw + ((s.xb - s.rec) + jj * DP);
            double *psum = smem_raw + (s.A - s.rec);
            const int k = tid & 127, part = warp >> 2;
            double pq = 0.0;
            if (k < K) {
                const double *col = rec_sh + k;
                switch (part) {
                    case 0: pq = f_quad_part16<0, ST>(col, xs_sh); break;
                    case 1: pq = f_quad_part16<1, ST>(col, xs_sh); break;
                    case 2: pq = f_quad_part16<2, ST>(col, xs_sh); break;
                    default: pq = f_quad_part16<3, ST>(col, xs_sh); break;
                }
                if (part > 0) psum[(part - 1) * 128 + k] = pq;
            }
            __syncthreads();
            if (warp < 4) {
                if (tid < K) {
                    const double q = 2.0 * ((pq + psum[tid]) + (psum[128 + tid] + psum[256 + tid]));
                    e_mine = f_finish_weight<ST>(rec_sh + tid + Ly::SC * ST, q, (own_live && tid == k_old) ? 1 : 0, wref,
                                                 p.fmtab);
                    if (e_mine != e_mine) sh.need_explicit = seq;
                } else if (tid == K) {
                    e_mine = 1.0;
                }
                ew[tid] = e_mine;   // for the rare paths
                // inclusive scan of this warp's 32 choices, straight from the registers
                incl = e_mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane == 31) sh.wtot[warp] = incl;
            }
        }
    }
    if (!cta_draw) {
        // thread k evaluates component k
        if (tid < K) {
            const double e = f_eval_lane<DP, ST>(s.rec + tid, xs, (own_live && tid == k_old) ? 1 : 0, wref, 0, p.fmtab);
            if (e != e) sh.need_explicit = seq;
            ew[tid] = e;
        } else if (tid == K) {
            ew[K] = 1.0;
        }
    }
    __syncthreads();
    F_PROF(PH_EVAL);
    bool expl = false;
    if (sh.need_explicit == seq) {
        f_explicit_own<DP>(p, s, k_old, n_old, xs, wref, ew, seq);
        expl = (sh.explicit_done == seq);
        cta_draw = false;   // ew[k_old] changed: the one-warp draw below reads ew[]
        F_PROF(PH_RARE);
    }

    // phase B: the draw (crpmm.py:75-78, utils.py:7-20): first index whose cumulative sum exceeds u * total
    int k_drawn;
    double mg_draw = 0.0;   // margin of the draw (uniform over the CTA)
    if (cta_draw) {
        const double w0 = sh.wtot[0], w1 = sh.wtot[1], w2 = sh.wtot[2], w3 = sh.wtot[3];
        const double p1 = w0, p2 = w0 + w1, p3 = p2 + w2, tot = p3 + w3;
        const double t0 = u_draw * tot;
        if (warp < 4) {
            const double pre = (warp == 0) ? 0.0 : (warp == 1) ? p1 : (warp == 2) ? p2 : p3;
            double excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 0.0;
            const double upper = pre + incl, lower = pre + excl;   // lower == the previous choice's upper, bit for bit
            const unsigned who = __ballot_sync(0xffffffffu, upper > t0);
            if (who != 0u && lane == __ffs(who) - 1) {
                sh.wcand[warp] = tid;
                sh.wmg[warp] = margin_ratio(fmin(t0 - lower, upper - t0), tot);
            } else if (who == 0u && lane == 0) {
                sh.wcand[warp] = 1 << 20;
            }
        }
        __syncthreads();
        int k = min(min(sh.wcand[0], sh.wcand[1]), min(sh.wcand[2], sh.wcand[3]));
        double mg = 0.0;
        if (k > K) k = K;                       // utils.py:20 fallback: the last index (also absorbs the zero padding)
        else mg = sh.wmg[k >> 5];
        if (!(tot > 0.0) || !(tot < INFINITY)) k = -2;
        k_drawn = k;
        mg_draw = mg;
        if (tid == 0) sh.k_new = k;
    } else {
        if (warp == 0) {
            double mg;
            const int k = f_warp_pick(ew, K + 1, s.ub[jj], &mg);
            if (lane == 0) { sh.k_new = k; sh.last_mg = mg; }
        }
        __syncthreads();
        k_drawn = sh.k_new;
        mg_draw = sh.last_mg;
    }
    if (k_drawn == -2) {
        if (cta_draw) __syncthreads();   // thread 0's sh.k_new = -2 must not land after the fallback's result
        f_log_domain_draw<DP>(p, s, K, k_old, own_live, expl, xs, wref, s.ub[jj], ew);
        k_drawn = sh.k_new;
        mg_draw = sh.last_mg;
    }
    if (k_drawn >= 0 && mg_draw < p.guard) {
        // the uniform is too close to a boundary of the drawn interval for the fast arithmetic: exact path
        __syncthreads();
        f_exact_redo<DP>(p, s, K, k_old, n_old, own_live, xs, wref, s.ub[jj], ew);
        k_drawn = sh.k_new;
        mg_draw = sh.last_mg;
        expl = own_live;
        F_PROF(PH_RARE);
    }
    if (tid == 0 && k_drawn >= 0) {
        sh.evals += K;
        const unsigned long long mb = (unsigned long long)__double_as_longlong(mg_draw);
        if (mb < sh.margin_bits) sh.margin_bits = mb;
    }
    F_PROF(PH_DRAW);
    F_COUNT(PH_STEPS);
    if (sh.error) return;
    const int k_new = k_drawn;
    if (k_new == k_old && !died) return;  // stay: nothing was touched (crpmm.py:82-85)

    // phase C: the datum moves: add_item (gaussian_components.py:154-169)
    const bool birth = (k_new == K);
    if (birth) {
        if (!f_open_component<DP>(p, s, K)) return;
    }
    f_move_phase<DP>(p, s, jj, k_old, k_new, own_live, birth, expl, died, seq);
}

// stage `nb` data starting at scan position pos into the buffers, then resolve them in order
template <int DP>
__device__ int f_run(const Params &p, const FSmem<DP> &s, long long pos, int nb, int &seq) {
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
    F_PROF(PH_STAGE);
    int done = 0;
    for (int jj = 0; jj < nb; ++jj) {
        seq += 1;
        f_step<DP>(p, s, jj, seq);
        if (s.sh->error) break;
        done = jj + 1;
    }
    return done;  // data fully resolved (uniform over the CTA)
}

// ---------------------------------------------------------------------------------------------
// speculative evaluation of the window [pos, pos + win): a warp per datum, lanes over components.
// Scan position j is owned by CTA (j % grid), warp ((j / grid) % NWARP) -- a fixed owner, so after a mover the
// warp still holds the datum it is asked to evaluate again: its inputs (WCache + xw) and its row of
// exponentials ew[], of which only the entries of the components the mover touched are recomputed.
// A candidate (anything but a provable "stay") is published with its inputs: mover slot of this warp, then
// atomicMin of (position << 12 | global warp id).
// ---------------------------------------------------------------------------------------------
struct WCache {
    long long nj;     // first scan position owned by this warp that has not been passed yet
    long long j, i;   // scan position being prepared / held (-1: none), its datum
    int uid;
    int stage;        // 0: nothing, 1: inputs loaded, 1 + c: c chunks of 32 components evaluated into the row ew[]
    int ver;          // record version the row's oldest chunk was evaluated at
    int K;            // live components when the row was started
    double u, lp;
};

// inputs of scan position c.j: datum index, label, uniform, cached log prior, and the row of X (into xw)
template <int DP>
__device__ __forceinline__ void f_load_inputs(const Params &p, WCache &c, double *xw) {
    const int lane = threadIdx.x & 31;
    const long long i = p.order ? p.order[c.j] : c.j;
    c.i = i;
    c.uid = __ldcg(p.z_uid + i);
    c.u = p.u[c.j];
    c.lp = p.log_prior[i];
    __syncwarp();
    if (lane < DP) xw[lane] = p.X[(size_t)i * DP + lane];
    __syncwarp();
    c.stage = 1;
}

// one chunk (32 components) of the row of exponentials of the held datum
template <int DP>
__device__ __forceinline__ void f_eval_chunk(const Params &p, const FSmem<DP> &s, WCache &c, int K, int ver,
                                             const double *xw, double *ew) {
    using Ly = Lay<DP>;
    const int lane = threadIdx.x & 31;
    const int chunk = c.stage - 1;
    if (chunk == 0) { c.ver = ver; c.K = K; }
    const int k = chunk * 32 + lane;
    if (k < K) {
        const int k_old = (c.uid >= 0) ? s.slot_of_uid[c.uid] : -1;
        ew[k] = f_eval_lane<DP, Ly::KS>(s.rec + k, xw, k == k_old ? 1 : 0, p.log_alpha + c.lp, 0, p.fmtab);
    }
    __syncwarp();
    c.stage += 1;
}

// is the (partial) row still usable at record version `ver` with K live components?
__device__ __forceinline__ bool f_row_usable(const FSh &sh, const WCache &c, int K, int ver) {
    return c.stage >= 2 && c.K <= K && c.ver >= sh.dall_ver && ver - c.ver <= DLOG - 1;
}

// bring a complete row evaluated at version c.ver up to the current version: re-evaluate the entries of the components
// touched since (two per version, from the dirty log)
template <int DP>
__device__ __forceinline__ void f_row_update(const Params &p, const FSmem<DP> &s, WCache &c, int K, int ver, int k_old,
                                             double wref, const double *xw, double *ew) {
    using Ly = Lay<DP>;
    const FSh &sh = *f_sh<DP>(s, s.sh);
    const int lane = threadIdx.x & 31;
    const int nd = 2 * (ver - c.ver);
    if (nd == 2) {
        // the common case on the critical path (a row re-examined right after one move): both components at once,
        // rows over half-warps
        int ka = sh.dlog_a[ver & (DLOG - 1)], kb = sh.dlog_b[ver & (DLOG - 1)];
        if (ka >= K) ka = -1;
        if (kb >= K) kb = -1;
        f_eval_rows2<DP>(f_sh<DP>(s, s.rec), xw, ka, kb, k_old, wref, ew, p.fmtab);
        c.ver = ver;
        return;
    }
    for (int base = 0; base < nd; base += 32) {   // 16 versions per pass; duplicates write the same value
        const int t = base + lane;
        if (t < nd) {
            const int v = c.ver + 1 + (t >> 1);
            const int k = (t & 1) ? sh.dlog_b[v & (DLOG - 1)] : sh.dlog_a[v & (DLOG - 1)];
            if (k >= 0 && k < K) ew[k] = f_eval_lane<DP, Ly::KS>(s.rec + k, xw, k == k_old ? 1 : 0, wref, 0, p.fmtab);
        }
    }
    __syncwarp();
    c.ver = ver;
}

// ---------------------------------------------------------------------------------------------
// One round of a warp's evaluator duty.  The warp owns the scan positions congruent to its global id; `c` holds
// the next one (c.nj) in some stage of preparation.
//  * not yet inside the window [pos, end): advance the preparation by ONE unit (load the inputs, or evaluate one
//    chunk of 32 components), so that rows are ready rounds before the chain reaches them and no full evaluation
//    sits on a round's critical path.  Rows go stale while they wait: the records of the components touched since
//    are re-evaluated from the dirty log when the row is used (exact: all other entries saw unchanged records).
//  * inside the window: finish the row, bring it up to date, draw.  A candidate (anything but a provable "stay")
//    is published with its inputs and, for a plain move, the drawn component: mover slot of this warp, then
//    atomicMin of (position << 12 | global warp id).
// ---------------------------------------------------------------------------------------------
template <int DP>
__device__ void f_window_eval(const Params &p, const FSmem<DP> &s, long long pos, long long win, int K,
                              unsigned long long *first_slot, unsigned int round, WCache &c, double &my_margin) {
    using Ly = Lay<DP>;
    constexpr int ST = Ly::KS;
    // shared-memory addressing for the evaluator's own dependent chain (see f_sh); f_eval_lane keeps generic operands
    const FSh &sh = *f_sh<DP>(s, s.sh);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long G = gridDim.x;
    const long long end = pos + win;
    double *ew = f_sh<DP>(s, s.ew) + (size_t)warp * Ly::WS;
    double *xw = f_sh<DP>(s, s.xw) + (size_t)warp * DP;
    const long long stride = G * NWARP;   // this warp owns the positions congruent to blockIdx.x + G * warp
    const unsigned long long gw = (unsigned long long)(blockIdx.x * NWARP + warp);
    while (c.nj < pos) c.nj += stride;
    const int ver = sh.ver;
    const int nch = (K + 31) >> 5;
    if (c.j != c.nj) { c.j = c.nj; c.stage = 0; }
    F_WPROF_BEGIN();
    if (c.j >= end) {
        // not needed this round: one unit of preparation
        if (c.j < p.N) {
            if (c.stage == 0) {
                F_WCAT(1);
                f_load_inputs<DP>(p, c, xw);
            } else if (c.j - pos < PREP_ZONE * win) {
                // evaluate ahead of the chain, but not as soon as possible: a waiting row ages with every move.  (Keeping
                // waiting rows current round by round was measured slower: every warp busy every round slows the
                // critical one.)  The zone is wide enough that a window which doubles after a quiet round still finds
                // its rows prepared -- an unprepared row puts nch full evaluations on that round's critical path, and
                // one in ~15 was unprepared with a zone of 3 windows.
                if (c.stage >= 2 && !f_row_usable(sh, c, K, ver)) c.stage = 1;
                if (c.stage - 1 < nch) {
                    F_WCAT(2);
                    f_eval_chunk<DP>(p, s, c, K, ver, xw, ew);
                } else if (c.uid >= 0 && c.ver != ver &&
                           (ver - c.ver >= PATCH_LAG || c.j - end < (long long)(p.tune & 4 ? 0 : p.near_zone) * win)) {
                    // a complete row that waits is brought up to date (a) before it outlives the dirty log and (b)
                    // every round once it is within one window of the chain: its first in-window examination then
                    // costs what a re-examination costs (one two-component update) instead of a pass over everything
                    // touched since it was prepared -- the slowest first examination is what a round waits for
                    F_WCAT(3);
                    f_row_update<DP>(p, s, c, K, ver, f_sh<DP>(s, s.slot_of_uid)[c.uid], p.log_alpha + c.lp, xw, ew);
                    c.K = K;
                }
            }
        }
        F_WPROF_END();
        return;
    }
    bool first_pass = true;
    for (long long j = c.j; j < end; j += stride) {
        if (!first_pass) {
            // later passes of a long window: stop once an earlier candidate is known (this datum would be redone)
            long long known = 0;
            if (lane == 0) known = (long long)(__ldcg(first_slot) >> 12);
            known = __shfl_sync(0xffffffffu, known, 0);
            if (known < j) break;
            c.j = j; c.stage = 0;
        }
        first_pass = false;
        if (c.stage == 0) f_load_inputs<DP>(p, c, xw);
        const int uid = c.uid;
        bool cand = (uid < 0);
        int k_old = -1, drawn = -1;
        if (!cand) {
            k_old = f_sh<DP>(s, s.slot_of_uid)[uid];
            if (f_sh<DP>(s, s.rec)[(Ly::SC + F_N) * ST + k_old] == 1.0) cand = true;  // the component would die
        }
        if (!cand) {
            const double wref = p.log_alpha + c.lp;
            if (c.stage >= 2 && !f_row_usable(sh, c, K, ver)) c.stage = 1;
            const bool fresh = (c.stage == 1);
            F_WCAT(fresh ? 6 : 4);
            // (profile builds: 8 = chunks evaluated inside the window, 9 / 10 = row brought up to date over several
            // versions / over one, 11 = row already current)
            while (c.stage - 1 < nch) {
                F_WSUB_BEGIN();
                f_eval_chunk<DP>(p, s, c, K, ver, xw, ew);
                F_WSUB_END(8);
            }
            if (!fresh && c.ver != ver) {
                F_WSUB_BEGIN();
                const int nver_ = ver - c.ver;
                (void)nver_;
                f_row_update<DP>(p, s, c, K, ver, k_old, wref, xw, ew);
                F_WSUB_END(nver_ > 1 ? 9 : 10);
            } else {
                F_WSUB_BEGIN();
                F_WSUB_END(11);
            }
            if (lane == 0) ew[K] = 1.0;
            __syncwarp();
            c.ver = ver;
            c.K = K;
            double mg;
            int k_new;
            {
                F_WSUB_BEGIN();
                k_new = f_warp_pick_inl(ew, K + 1, c.u, &mg);
                F_WSUB_END(7);
            }
            if (k_new >= 0 && mg < p.guard) {
                cand = true;   // margin guard: the full step redoes this datum (on the exact path if it agrees)
            } else if (k_new != k_old) {
                cand = true;   // includes -2 (an untrusted / overflowed entry): the step redoes it in full
                // a plain move between two live components needs no second evaluation if it turns out to be
                // the first candidate: every datum in front of it stayed, so the records it saw are current
                if (k_new >= 0 && k_new < K) { drawn = k_new; my_margin = fmin(my_margin, mg); }
            } else {
                my_margin = fmin(my_margin, mg);
            }
            __syncwarp();
        }
        if (cand) {
            F_WCAT(5);
            // mover slots are double-buffered by round parity: a replica that lags (nothing bounds by how much) may
            // still be fetching the previous round's winner from this warp's other slot; it cannot lag two rounds
            double *mv = p.mvbuf + ((size_t)(round & 1u) * G * NWARP + gw) * (DP + MV_EXTRA);
            if (lane < DP) __stcg(mv + lane, xw[lane]);
            if (lane == 0) {
                __stcg(mv + DP, c.u);
                __stcg(mv + DP + 1, c.lp);
                __stcg(mv + DP + 2, __longlong_as_double(c.i));
                __stcg(mv + DP + 3, __longlong_as_double((long long)uid));
                __stcg(mv + DP + 4, __longlong_as_double((long long)drawn));
            }
            // no fence here: the slot is read only after the round's grid barrier, whose release (a __threadfence by
            // this CTA's thread 0 after __syncthreads) already orders these stores before the barrier arrival
            if (lane == 0) atomicMin(first_slot, ((unsigned long long)j << 12) | gw);
            break;
        }
    }
    F_WPROF_END();
}

// ---------------------------------------------------------------------------------------------
// Long windows (movers far apart): a THREAD per datum, components in a loop, so every record element is one
// broadcast shared-memory read for 32 data and the evaluation is FP64-issue bound instead of latency bound.
// Only the stay test is needed here, and it needs no stored row: with e_k = exp(weight_k - wref) accumulated in
// component order, total s, prefix P = sum_{k < k_old} e_k and e_own decide "stay" iff P <= u s < P + e_own
// (utils.py:15-20).  Anything else is a candidate, published without inputs (low bits 4095): every CTA then stages
// the datum from global memory and resolves it in full.
// Scan position j is owned by CTA (j % grid), thread ((j / grid) % TF).
// ---------------------------------------------------------------------------------------------
constexpr unsigned long long BULK_TAG = 4095ULL;

template <int DP>
__device__ __noinline__ void f_bulk_eval(const Params &p, const FSmem<DP> &s, long long pos, long long win, int K,
                                         unsigned long long *first_slot, double &my_margin) {
    using Ly = Lay<DP>;
    constexpr int ST = Ly::KS;
    const long long G = gridDim.x;
    const long long end = pos + win;
    const long long stride = G * TF;
    long long j = pos + (long long)blockIdx.x + G * threadIdx.x;   // relative ownership: windows are long here
    // the records start the kernel's dynamic shared array: addressing them through it (not through the generic
    // pointer s.rec) makes the loads below LDS broadcasts instead of generic loads that queue in the LSU's
    // local/global path (stall_lg_throttle in the converged-sweep profile)
    extern __shared__ __align__(16) double smem_raw[];
    const double *rec = smem_raw;
    const double *fmtab = p.fmtab;
    bool first_pass = true;
    for (; j < end; j += stride) {
        if (!first_pass) {
            // later passes: stop once a candidate in front of this datum is known (it would be redone)
            if ((long long)(__ldcg(first_slot) >> 12) < j) break;
        }
        first_pass = false;
        const long long i = p.order ? p.order[j] : j;
        const int uid = __ldcg(p.z_uid + i);
        bool cand = (uid < 0);
        int k_old = -1;
        if (!cand) {
            k_old = s.slot_of_uid[uid];
            if (rec[(Ly::SC + F_N) * ST + k_old] == 1.0) cand = true;  // the component would die
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
            const double wref = p.log_alpha + p.log_prior[i];
            const double u = p.u[j];
            double ssum = 0.0, pre = 0.0, eown = 0.0;
            bool ok = true;
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                const double *col = rec + k;   // uniform over the warp: broadcast reads
                double q = 0.0;
#pragma unroll
                for (int a = 0; a < DP; ++a) {
                    double r = 0.0;
#pragma unroll
                    for (int b = 0; b < a; ++b)
                        r = fma(col[(a * (a + 1) / 2 + b) * ST], col[(Ly::MU + b) * ST] - x[b], r);
                    const double da = col[(Ly::MU + a) * ST] - x[a];
                    r = fma(0.5 * col[(a * (a + 1) / 2 + a) * ST], da, r);
                    q = fma(da, r, q);
                }
                q *= 2.0;
                const double *sc = col + Ly::SC * ST;
                const bool own = (k == k_old);
                const double arg = own ? 1.0 - sc[F_BETA * ST] * q : 1.0 + sc[F_G * ST] * q;
                if (own && !(arg > OM_MIN)) { ok = false; break; }
                const double hh = own ? 1.0 - sc[F_H * ST] : sc[F_H * ST];
                const double cc = own ? sc[F_CWO * ST] : sc[F_CW * ST];
                const double t = (cc - hh * fm::f_log(arg, fmtab)) - wref;
                const double e = (t < EXP_CUTOFF) ? 0.0 : fm::f_exp(t, fmtab);
                if (k < k_old) pre += e;
                if (own) eown = e;
                ssum += e;
            }
            ssum += 1.0;   // the new-table entry: exp(wref - wref)
            const double t0 = u * ssum;
            if (!ok || !(ssum < INFINITY) || !(t0 >= pre) || !(t0 < pre + eown)) {
                cand = true;
            } else {
                const double mg = margin_ratio(fmin(t0 - pre, pre + eown - t0), ssum);
                if (mg < p.guard) cand = true;   // margin guard: resolved in full by every CTA
                else my_margin = fmin(my_margin, mg);
            }
        }
        if (cand) {
            atomicMin(first_slot, ((unsigned long long)j << 12) | BULK_TAG);
            break;
        }
    }
}

// length of the next window: about twice the running gap between movers, whole rows of one datum per SM
__device__ __forceinline__ long long f_next_window(double gap, long long pos, long long N, float factor) {
    // 32-bit / single precision on purpose: this runs on one thread between two rounds
    const int G = (int)gridDim.x;
    const int wcap = G * TF * BULK_PASSES_MAX;
    // window length in gaps: longer windows pay once movers are sparse (measured at C3: 3 gaps beat 2 by 2-3 % from a
    // gap of ~80 data, 2 gaps beat 3 by 2 % at a gap of ~40); factor > 0 (env BGMM_WIN_FACTOR) overrides
    if (factor <= 0.0f) factor = 2.0f + fminf(fmaxf(((float)gap - 40.0f) * (1.0f / 30.0f), 0.0f), 1.0f);
    int win = (int)fminf(factor * (float)gap, (float)wcap);
    win = ((win + G - 1) / G) * G;
    if (win < G) win = G;
    const long long left = N - pos;
    return left < (long long)win ? left : (long long)win;
}

// ---------------------------------------------------------------------------------------------
// Window mode: speculative rounds (f_window_eval / f_bulk_eval -> one grid barrier -> every CTA applies the first
// candidate) until the running gap between movers says sequential steps pay, the sweep ends, or an error.  Out of line
// and self-contained (its own evaluator cache: after sequential batches the cached rows are void anyway) so that its
// working set does not share a register allocation with the call sites of the sequential engine -- what does not fit
// in registers goes to local memory, and local memory has next to no L1 under ~220 KB of shared memory.
// ---------------------------------------------------------------------------------------------
template <int DP>
__device__ __noinline__ int f_window_run(const Params &p, const FSmem<DP> &s, int seq) {
    FSh &sh = *s.sh;
    Ctl *ctl = p.ctl;
    const int tid = threadIdx.x;
    const bool cta0 = (p.writer != 0);
    WCache cache;
    cache.nj = (long long)blockIdx.x + (long long)gridDim.x * (tid >> 5);
    cache.j = -1; cache.i = 0; cache.uid = -1; cache.stage = 0; cache.ver = -1; cache.K = 0; cache.u = 0.0; cache.lp = 0.0;
    // minimum margin over this warp's window evaluations (committed and discarded alike: a lower bound of the
    // chain's true minimum margin); folded into the control block on the way out
    double win_margin = 1.0;
    while (true) {
        const long long pos = sh.pos;
        if (pos >= p.N || sh.error != 0 || sh.mode != 1) break;
        {
            const unsigned int r = sh.round;
            unsigned long long *slot = &ctl->first3[r % 3u][0];
            if (cta0 && tid == 0) __stcg(&ctl->first3[(r + 1u) % 3u][0], ~0ULL >> 1);
            const int K = sh.K;
            const long long win = sh.win;
            F_PROF(PH_HEAD);
            const bool bulk = (win >= (long long)gridDim.x * BULK_MIN_ROWS);
            if (bulk) f_bulk_eval<DP>(p, s, pos, win, K, slot, win_margin);
            else f_window_eval<DP>(p, s, pos, win, K, slot, r, cache, win_margin);
#ifdef BGMM_PROFILE
            __syncthreads();   // only to attribute the waiting to the right phase clock
#endif
            F_PROF(PH_WINEVAL);
            f_round_barrier(ctl, slot, r + 1u, &sh.fv);
            F_PROF(PH_BARRIER);
            F_COUNT(PH_ROUNDS);
            const unsigned long long fv = sh.fv;
            const long long f = (long long)(fv >> 12);
            const long long end = pos + win;
            if (f < end && (fv & 4095ULL) == BULK_TAG) {
                // candidate of the thread-per-datum evaluator: stage it from global memory and resolve it in full
                if (tid == 0) { sh.evals += (f - pos) * (long long)K; sh.wasted += end - (f + 1); }
                __syncthreads();
                f_run<DP>(p, s, f, 1, seq);
                __syncthreads();
                if (tid == 0) {
                    sh.pos = f + (sh.error ? 0 : 1);
                    sh.gap = 0.7 * sh.gap + 0.3 * (double)(f - pos + 1);
                }
            } else if (f < end) {
                // every CTA resolves the first candidate itself, from the inputs its evaluator published
                const double *mv = p.mvbuf + ((size_t)(r & 1u) * gridDim.x * NWARP + (size_t)(fv & 4095ULL)) * (DP + MV_EXTRA);
                if (tid < DP + MV_EXTRA) {
                    const double v = __ldcg(mv + tid);
                    if (tid < DP) s.xb[tid] = v;
                    else if (tid == DP) s.ub[0] = v;
                    else if (tid == DP + 1) s.lpb[0] = v;
                    else if (tid == DP + 2) s.ib[0] = __double_as_longlong(v);
                    else if (tid == DP + 3) s.uidb[0] = (int)__double_as_longlong(v);
                    else sh.k_new = (int)__double_as_longlong(v);
                }
                if (tid == 0) { sh.evals += (f - pos) * (long long)K; sh.wasted += end - (f + 1); }
                __syncthreads();
                const int drawn = sh.k_new;
                F_PROF(PH_STAGE);
                seq += 1;
                if (drawn >= 0) {
                    // a plain move, already drawn by its evaluator against the current records
                    if (tid == 0) sh.evals += K;
                    f_move_phase<DP>(p, s, 0, s.slot_of_uid[s.uidb[0]], drawn, true, false, false, false, seq);
                } else {
                    f_step<DP>(p, s, 0, seq);
                    __syncthreads();   // f_step can return without a trailing barrier (stay / error paths)
                }
                if (tid == 0) {
                    sh.pos = f + (sh.error ? 0 : 1);
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
                if (p.engine == 0 && sh.gap < (double)p.gap_to_seq) sh.mode = 0;
                sh.win = f_next_window(sh.gap, sh.pos, p.N, p.win_factor);
            }
                }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) win_margin = fmin(win_margin, __shfl_xor_sync(0xffffffffu, win_margin, o));
    if ((tid & 31) == 0 && win_margin < 1.0)
        atomicMin(&ctl->margin_bits, (unsigned long long)__double_as_longlong(win_margin));
    return seq;
}

}  // namespace fast
}  // namespace bgmm
#include "bgmm_seq.cuh"
namespace bgmm {
namespace fast {

// ---------------------------------------------------------------------------------------------
// the sweep kernel: cooperative grid, one CTA per SM -- or, for bgmm_sweep_many, one CTA per chain (solo)
// ---------------------------------------------------------------------------------------------
template <int DP>
__device__ __forceinline__ void fast_sweep_body(const Params &p_in) {
    extern __shared__ __align__(16) double smem_raw[];
    using Ly = Lay<DP>;
    constexpr int ST = Ly::KS;
    // the out-of-line device functions take these by reference: keep one copy in shared memory, not one per thread
    // in local memory
    __shared__ Params p_sh;
    __shared__ FSmem<DP> s_sh;
    if (threadIdx.x == 0) {
        p_sh = p_in;
        p_sh.writer = (p_in.solo || blockIdx.x == 0) ? 1 : 0;
        s_sh = fast_carve<DP>(smem_raw, p_in);
    }
    __syncthreads();
    const Params &p = p_sh;
    const FSmem<DP> &s = s_sh;
    FSh &sh = *s.sh;
    Ctl *ctl = p.ctl;
    const int tid = threadIdx.x;
    const bool cta0 = (p.writer != 0);
    {
        using O = SOff<DP>;
        const bool ok = s.prior == smem_raw + O::PRIOR && s.tmprec == smem_raw + O::TMPREC && s.ew == smem_raw + O::EW &&
                        s.xw == smem_raw + O::XW && s.xb == smem_raw + O::XB && s.ub == smem_raw + O::UB &&
                        s.lpb == smem_raw + O::LPB && (double *)s.ib == smem_raw + O::IB &&
                        (double *)s.uidb == smem_raw + O::UIDB && s.dv == smem_raw + O::DV && s.vv == smem_raw + O::VV &&
                        s.nt == smem_raw + O::NT && s.A == smem_raw + O::A && s.W == smem_raw + O::W &&
                        s.mm == smem_raw + O::MM && s.fm == smem_raw + O::FM && (double *)s.sh == smem_raw + O::SH;
        if (!ok) {   // fast_carve and SOff disagree: a build error, reported instead of silently corrupting memory
            if (tid == 0 && cta0) __stcg(&ctl->error, -7);
            return;
        }
    }

    // ---- prologue: replicate the chain state ----
    if (tid == 0) {
        sh.K = __ldcg(&ctl->K);
        sh.n_free = __ldcg(&ctl->n_free);
        sh.error = 0;
        sh.pos = p.start_pos;
        sh.gap = p.init_gap;
        sh.moves = sh.births = sh.deaths = sh.evals = sh.windows = sh.seq_data = sh.wasted = 0;
        sh.explicit_evals = sh.refreshes = 0;
        sh.guard_hits = sh.fast_steps = 0;
        sh.last_mg = 1.0;
        const double one = 1.0;
        sh.margin_bits = (unsigned long long)__double_as_longlong(one);
        sh.round = 0;
        sh.k_new = 0; sh.need_explicit = 0; sh.explicit_done = 0; sh.refresh_a = 0; sh.refresh_b = 0;
        sh.rare_seq = 0;
        sh.ver = 1; sh.dall_ver = 1;
        for (int t = 0; t < DLOG; ++t) sh.dlog_a[t] = sh.dlog_b[t] = -1;
        for (int t = 0; t < PH_COUNT; ++t) sh.prof[t] = 0;
        for (int t = 0; t < 64; ++t) (&sh.tprof[0][0])[t] = 0;
        sh.prof_last = clock64();
        sh.mode = (p.engine == 2) ? 1 : ((p.engine == 1) ? 0 : (p.init_gap >= (double)p.gap_to_win ? 1 : 0));
        if (p.solo) sh.mode = 0;
        sh.win = f_next_window(p.init_gap, p.start_pos, p.N, p.win_factor);
        const uint32_t mb = smem_u32(&sh.mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int e = tid; e < fm::TAB_LEN; e += TF) s.fm[e] = __ldg(p_in.fmtab + e);
    if (tid == 0) p_sh.fmtab = s.fm;   // every later use of p.fmtab reads the shared-memory copy (visible after the barrier below)
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
        // records: one TMA bulk copy of the whole element-major table (cp.async.bulk + mbarrier complete_tx),
        // rounded up to the 16-byte granule (the global buffer and the shared region are padded)
        const uint32_t bytes = (uint32_t)(((size_t)Ly::R * ST * sizeof(double) + 15) & ~(size_t)15);
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
        SpinWatch wd;
        while (!okw) {
            wd.poll(ctl, 3);
            asm volatile(
                "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(okw)
                : "r"(mb), "r"(0u)
                : "memory");
        }
    }
    // no replica may still be reading the initial state when CTA 0 starts changing it
    f_grid_barrier(p);

    int seq = 0;

    // ---- main loop ----
    while (true) {
        const long long pos = sh.pos;
        if (pos >= p.N || sh.error != 0) break;
        const int mode = sh.mode;
        if (mode == 0 && sh.K < SEQ_KMAX && !(p.tune & 8)) {
            // dense movers: the register-resident sequential step (bgmm_seq.cuh); runs batches until the engine should
            // change mode, K outgrows its layout, or the sweep ends
            __syncthreads();
            seq = f_seq_run<DP>(p, s, seq);
        } else if (mode == 0) {
            const int nb = (int)min((long long)SEQ_BATCH, p.N - pos);
            const long long moves0 = sh.moves;
            __syncthreads();
            const int done = f_run<DP>(p, s, pos, nb, seq);
            __syncthreads();
            if (tid == 0) {
                const long long mv = sh.moves - moves0;
                sh.gap = 0.5 * sh.gap + 0.5 * (double)nb / ((double)mv + 0.5);
                sh.seq_data += done;
                sh.pos = pos + done;
                sh.dall_ver = sh.ver;   // several changes since the evaluators last looked
                if (p.engine == 0 && sh.gap >= (double)p.gap_to_win) sh.mode = 1;
                sh.win = f_next_window(sh.gap, sh.pos, p.N, p.win_factor);
            }
        } else {
            // sparse movers: speculative window rounds until the engine should change mode or the sweep ends
            __syncthreads();
            seq = f_window_run<DP>(p, s, seq);
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
            __stcg(p.counts + t, t < K ? (long long)s.rec[(Ly::SC + F_N) * ST + t] : 0LL);
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
            __stcg(&ctl->guard_hits, __ldcg(&ctl->guard_hits) + sh.guard_hits);
            __stcg(&ctl->fast_steps, __ldcg(&ctl->fast_steps) + sh.fast_steps);
            atomicMin(&ctl->margin_bits, sh.margin_bits);
            __stcg(&ctl->gap, sh.gap);
            for (int t = 0; t < PH_COUNT; ++t) __stcg(&ctl->prof[t], sh.prof[t]);
            for (int t = 0; t < 64; ++t) __stcg(&ctl->tprof[0][0] + t, (&sh.tprof[0][0])[t]);
        }
    }
}

// One entry point for both launch shapes (ONE copy of the device functions: ptxas 12.9 was seen to encode the TMA
// reduction of bgmm_seq.cuh as an integer add -- UBLKRED.ADD.U64 instead of .ADD.F64.RN -- in the second kernel's clone
// of the same PTX function; tests/test_abi.py checks the shipped SASS):
//   pv == nullptr: the cooperative grid advances the one chain of p_in (replicated state machine);
//   pv != nullptr: bgmm_sweep_many -- CTA c advances chain c (its Params at pv[c]) on its own, the sequential engine with
//                  the CTA as the chain's only replica; chains beyond the number of resident CTAs simply queue.
template <int DP> __global__ void __launch_bounds__(TF, 1) k_fast_sweep(const Params p_in, const Params *__restrict__ pv) {
    __shared__ Params p_mine;
    if (threadIdx.x == 0) p_mine = pv ? pv[blockIdx.x] : p_in;
    __syncthreads();
    fast_sweep_body<DP>(p_mine);
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
                                     A, W, mm, p.recB + k, Ly::KS);
    } else {
        ok = f_exact_record_warp<DP>(p, 1, nullptr, nullptr, 0.0, nullptr, rc, A, W, mm, p.recB_prior, 1);
    }
    if (!ok && threadIdx.x == 0) *err = -4;
}

}  // namespace fast
}  // namespace bgmm
