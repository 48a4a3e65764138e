// bgmm_big.cuh -- the cluster-resident sweep engine for full covariance (NIW) components with padded D = 32 / 64.
//
// Semantics: the per-datum loop of CRPMM / PCRPMM.collapsed_gibbs_sampler (igmm/crpmm.py:57-88, igmm/pcrpmm.py:93-131),
// strictly sequential over the scan order -- the common case of a step (the datum stays, or moves between two live
// components); everything else ends the launch and is resolved by the generic engine (bgmm_sweep.cuh) on the host's
// next turn.
//
// At D = 64 one component's evaluation record (B = S_N^-1 packed, the mean, the scalars) is 17 KB and K = 100 of them are
// 1.7 MB: no SM holds them.  A thread-block CLUSTER does: 16 CTAs (4 at D = 32) on 16 SMs of one GPC, component k owned
// by CTA k mod C.  One step of the chain:
//   1. every CTA evaluates the quadratic forms and weights of the components it owns (a group of D/2 lanes per
//      component, rows r and D-1-r of the triangle per lane: D + 1 elements each, stored lane-interleaved so that the
//      loads are conflict-free), and stores exp(weight - reference) straight into CTA 0's shared memory
//      (st.shared::cluster through DSMEM);
//   2. cluster barrier (hardware: barrier.cluster arrive.release / wait.acquire);
//   3. CTA 0 scans the K + 1 choices and draws (utils.py:7-20), and writes the result into every CTA's shared memory;
//   4. cluster barrier;
//   5. the owner(s) of the two touched components apply the rank-one update of B (Sherman-Morrison), the mean and the
//      scalars; CTA 0 appends (datum, from, to) to a move log.
// Two cluster barriers per datum, no global memory on the critical path.  The bit-exact statistics (D (D + 1) / 2 + D
// doubles per component: 2144 at D = 64) are not touched per move at all: k_big_replay applies the log afterwards, one CTA
// per component with the component's statistics in registers, every element receiving the reference's -= fl(x_a x_b) /
// += fl(x_a x_b) in chain order (gaussian_components.py:165-166, :184-185) -- the same bits, a few ms per sweep instead of
// an L2 reduction of 2 x 2144 doubles inside every step (measured: 55 us per move through TMA reductions).  Records are rebuilt from the bit-exact
// statistics at every launch (k_big_prep), and a launch covers at most BIG_SPAN data, which bounds the drift of the
// incrementally updated records like REFRESH_EVERY does in bgmm_fast.cuh.
#pragma once
#include "bgmm_fast.cuh"

namespace bgmm {
namespace big {

using fast::F_N; using fast::F_LDS; using fast::F_CNT; using fast::F_CW; using fast::F_G; using fast::F_H;
using fast::F_BETA; using fast::F_CWO; using fast::NSC;
using fast::NT_CN; using fast::NT_G; using fast::NT_H; using fast::NT_BETA; using fast::NT_W;

constexpr int TB = 512;              // threads per CTA
constexpr int SB = 16;               // data staged per batch
constexpr int KCH = 128;             // choices (K + 1) held by warps 0..3 of CTA 0
constexpr int E_RARE = 2;            // internal: the datum at Ctl::pos needs the general step
constexpr long long BIG_SPAN = 1 << 16;   // data per launch (records are rebuilt from the statistics between launches)

template <int DP> struct BL {
    static constexpr int PP = DP * (DP + 1) / 2;
    static constexpr int NL = DP / 2;          // lanes per component in the evaluation
    static constexpr int EPL = DP + 1;         // matrix elements per lane: rows r and DP - 1 - r
    static constexpr int MU = PP, SC = PP + DP, R = PP + DP + NSC;
    static constexpr int C = (DP == 64) ? 16 : 4;        // CTAs per cluster
    static constexpr int LMAX = KCH / C;                 // components per CTA
    static constexpr int SLOTS = TB / NL;                // groups of NL lanes per CTA
    static constexpr int SPC = SLOTS / LMAX;             // lane groups per component (they split the elements of a lane)
    static constexpr int SEGS = TB / DP;                 // v = B d: threads per row
    static constexpr int CPS = DP / SEGS;                //          columns per thread
    static_assert(SPC >= 1 && SLOTS % LMAX == 0, "lane groups must tile the components");
    static_assert(SEGS <= 32 && DP % SEGS == 0, "a row's threads sit in one warp");
};

// storage index of element (a, b), a >= b, of the packed triangle: lane-interleaved rows r / DP - 1 - r
template <int DP> __host__ __device__ __forceinline__ int pidx(int a, int b) {
    constexpr int NL = BL<DP>::NL;
    const int lane = a < NL ? a : DP - 1 - a;
    const int j = a < NL ? b : b + lane + 1;
    return j * NL + lane;
}

__device__ __forceinline__ unsigned cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(const void *smem_ptr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(smem_ptr)), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_remote_f64(uint32_t addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void st_remote_u32(uint32_t addr, unsigned v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

struct BSh {
    // result of the draw, written by CTA 0 into every CTA (remote stores), read after the second cluster barrier
    double res_mg;
    int res_k, res_rare;
    // CTA-local
    double wtot[4];
    double gam, den, rk;
    long long moves, evals, steps, n_log;
    unsigned long long margin_bits;
    int K, stop;
};

template <int DP> struct BSmem {
    double *rec;      // LMAX * R: the records this CTA owns, local component l = k / C
    double *dsh;      // LMAX * DP: d = m - x of the datum for every owned component
    double *qpart;    // SLOTS partial quadratic forms
    double *qloc;     // LMAX
    double *ebuf;     // KCH: exp(weight - reference) of every choice (used in CTA 0)
    double *vbuf;     // DP: v = B d of the component being updated
    double *xb, *ub, *lpb;
    long long *ib;
    int *uidb, *kob;
    double *fm;
    int *slot_of_uid, *uid_of_slot;
    unsigned short *rc;   // PP: (a << 8) | b of row-major packed element e (the statistics' order)
    BSh *sh;
};

template <int DP> __host__ __device__ inline size_t big_smem_bytes(int K_max) {
    using L = BL<DP>;
    size_t d = (size_t)L::LMAX * L::R + (size_t)L::LMAX * DP + L::SLOTS + L::LMAX + KCH + DP + 2 * (size_t)(L::PP + DP) +
               (size_t)SB * DP + 2 * SB + SB + SB + fm::TAB_LEN + 8;
    size_t b = d * sizeof(double) + 2 * (size_t)K_max * sizeof(int) + (((size_t)L::PP + 7) & ~(size_t)7) * sizeof(unsigned short) +
               ((sizeof(BSh) + 15) & ~(size_t)15) + 64;
    return (b + 15) & ~(size_t)15;
}

template <int DP> __device__ inline BSmem<DP> big_carve(double *base, int K_max) {
    using L = BL<DP>;
    BSmem<DP> s;
    double *q = base;
    s.rec = q; q += (size_t)L::LMAX * L::R;
    s.dsh = q; q += (size_t)L::LMAX * DP;
    s.qpart = q; q += L::SLOTS;
    s.qloc = q; q += L::LMAX;
    s.ebuf = q; q += KCH;
    s.vbuf = q; q += DP;
    s.xb = q; q += SB * DP;
    s.ub = q; q += SB;
    s.lpb = q; q += SB;
    s.ib = (long long *)q; q += SB;
    s.uidb = (int *)q; q += SB / 2;
    s.kob = (int *)q; q += SB / 2;
    s.fm = q; q += fm::TAB_LEN;
    q = (double *)(((uintptr_t)q + 15) & ~(uintptr_t)15);
    s.sh = (BSh *)q;
    unsigned short *r = (unsigned short *)((char *)q + ((sizeof(BSh) + 15) & ~(size_t)15));
    s.rc = r; r += (L::PP + 7) & ~7;
    int *t = (int *)r;
    s.slot_of_uid = t; t += K_max;
    s.uid_of_slot = t; t += K_max;
    return s;
}

// ---------------------------------------------------------------------------------------------
// records of all live components in this engine's layout, from the bit-exact statistics: one warp per component
// (fast::f_exact_record_warp: Cholesky of S_N, inverse, scalars from the count table), then the triangle permuted into
// the lane-interleaved order.  Dynamic shared memory: PP + DP * DP + DP doubles + PP shorts.
// ---------------------------------------------------------------------------------------------
template <int DP> __global__ void k_big_prep(const Params p, int K, double *__restrict__ rec_out, int *err) {
    using L = BL<DP>;
    extern __shared__ __align__(16) double sm[];
    double *A = sm, *W = A + L::PP, *mm = W + DP * DP, *tmp = mm + DP;
    unsigned short *rc = (unsigned short *)(tmp + L::R);
    const int k = blockIdx.x, lane = threadIdx.x;
    for (int e = lane; e < L::PP; e += 32) {
        int a, b;
        decode_row_idx(e, a, b);
        rc[e] = (unsigned short)((a << 8) | b);
    }
    __syncwarp();
    if (k >= K) return;
    const bool ok = fast::f_exact_record_warp<DP>(p, 0, p.num + (size_t)k * DP, p.S + (size_t)k * L::PP, (double)p.counts[k],
                                                  nullptr, rc, A, W, mm, tmp, 1);
    if (!ok) { if (lane == 0) *err = -4; return; }
    double *out = rec_out + (size_t)k * L::R;
    for (int e = lane; e < L::PP; e += 32) out[pidx<DP>(rc[e] >> 8, rc[e] & 0xff)] = tmp[e];
    for (int e = L::PP + lane; e < L::R; e += 32) out[e] = tmp[e];
}

// element (hi, lo) of the symmetric matrix for any pair of indices
template <int DP> __device__ __forceinline__ double bsym(const double *__restrict__ B, int a, int b) {
    return a >= b ? B[pidx<DP>(a, b)] : B[pidx<DP>(b, a)];
}

// ---------------------------------------------------------------------------------------------
// the sweep kernel: one cluster per chain
// ---------------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(TB, 1) k_big_sweep(const Params p_in, double *__restrict__ rec_in, long long pos_limit,
                                                    int4 *__restrict__ mlog) {
    using L = BL<DP>;
    constexpr int C = L::C, NL = L::NL, R = L::R;
    extern __shared__ __align__(16) double smem_raw[];
    __shared__ Params p_sh;
    __shared__ BSmem<DP> s_sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { p_sh = p_in; s_sh = big_carve<DP>(smem_raw, p_in.K_max); }
    __syncthreads();
    const Params &p = p_sh;
    const BSmem<DP> &s = s_sh;
    BSh &sh = *s.sh;
    Ctl *ctl = p.ctl;
    const int c = (int)cluster_rank();

    // ---- prologue ----
    const int K = __ldcg(&ctl->K);
    const int Lc = (K > c) ? (K - c + C - 1) / C : 0;     // components this CTA owns: k = l * C + c
    if (tid == 0) {
        sh.K = K; sh.stop = 0; sh.moves = sh.evals = sh.steps = 0; sh.n_log = 0;
        const double one = 1.0;
        sh.margin_bits = (unsigned long long)__double_as_longlong(one);
        sh.res_k = 0; sh.res_rare = 0; sh.res_mg = 1.0;
    }
    for (int l = 0; l < Lc; ++l) {
        const double *src = rec_in + (size_t)(l * C + c) * R;
        for (int e = tid; e < R; e += TB) s.rec[(size_t)l * R + e] = __ldcg(src + e);
    }
    for (int e = tid; e < fm::TAB_LEN; e += TB) s.fm[e] = __ldg(p.fmtab + e);
    for (int e = tid; e < L::PP; e += TB) {
        int a, b;
        decode_row_idx(e, a, b);
        s.rc[e] = (unsigned short)((a << 8) | b);
    }
    for (int t = tid; t < p.K_max; t += TB) {
        s.slot_of_uid[t] = __ldcg(p.slot_of_uid + t);
        s.uid_of_slot[t] = __ldcg(p.uid_of_slot + t);
    }
    for (int e = tid; e < KCH; e += TB) s.ebuf[e] = 0.0;
    __syncthreads();
    cluster_sync_all();   // every CTA's shared memory is initialised before anyone stores into it remotely

    const uint32_t ebuf0 = map_to_cta(s.ebuf, 0);       // CTA 0's choice buffer
    long long pos = p.start_pos;
    bool stop = false;
    while (pos < pos_limit && !stop) {
        const int nb = (int)min((long long)SB, pos_limit - pos);
        __syncthreads();
        for (int t = tid; t < nb * DP; t += TB) {
            const int jj = t / DP, a = t % DP;
            const long long j = pos + jj;
            const long long i = p.order ? p.order[j] : j;
            s.xb[jj * DP + a] = p.X[(size_t)i * DP + a];
            if (a == 0) {
                s.ib[jj] = i;
                const int uid = __ldcg(p.z_uid + i);
                s.uidb[jj] = uid;
                s.ub[jj] = p.u[j];
                s.lpb[jj] = p.log_prior[i];
                s.kob[jj] = uid >= 0 ? s.slot_of_uid[uid] : -1;
            }
        }
        __syncthreads();
        int done = 0;
        for (int jj = 0; jj < nb; ++jj) {
            const int k_old = s.kob[jj];
            if (k_old < 0) { stop = true; if (tid == 0) sh.res_rare = 5; break; }   // an unassigned datum: the general step
            const double *x = s.xb + jj * DP;
            const double wref = p.log_alpha + s.lpb[jj];
            // ---- 1. quadratic forms and weights of the components this CTA owns ----
            for (int t = tid; t < Lc * DP; t += TB) {
                const int l = t / DP, a = t % DP;
                s.dsh[t] = s.rec[(size_t)l * R + L::MU + a] - x[a];
            }
            __syncthreads();
            {
                const int slot = tid / NL, ln = tid % NL;
                const int l = slot / L::SPC, seg = slot % L::SPC;
                double acc = 0.0;
                if (l < Lc) {
                    const double *B = s.rec + (size_t)l * R;
                    const double *d = s.dsh + l * DP;
                    // lane ln: row r = ln (columns 0..r at j = 0..r), then row r2 = DP - 1 - ln (columns 0..r2);
                    // sum_a d_a (sum_{b<a} B_ab d_b + B_aa d_a / 2), the lane's D + 1 elements split over SPC groups
                    const int r = ln, r2 = DP - 1 - ln;
                    constexpr int JS = (L::EPL + L::SPC - 1) / L::SPC;
                    const int j0 = seg * JS, j1 = min(L::EPL, j0 + JS);
                    double a1 = 0.0, a2 = 0.0;
                    for (int j = j0; j < j1; ++j) {
                        const double bv = B[j * NL + ln];
                        if (j <= r) {
                            a1 = fma(j == r ? 0.5 * bv : bv, d[j], a1);
                        } else {
                            const int b = j - r - 1;
                            a2 = fma(b == r2 ? 0.5 * bv : bv, d[b], a2);
                        }
                    }
                    acc = d[r] * a1 + d[r2] * a2;
                }
#pragma unroll
                for (int o = NL / 2; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (ln == 0) s.qpart[slot] = acc;
            }
            __syncthreads();
            if (tid < Lc) {
                const int l = tid, k = l * C + c;
                double qs = 0.0;
#pragma unroll
                for (int g = 0; g < L::SPC; ++g) qs += s.qpart[l * L::SPC + g];
                const double q = 2.0 * qs;
                const double *sc = s.rec + (size_t)l * R + L::SC;
                const int own = (k == k_old) ? 1 : 0;
                double e = fast::f_finish_weight<1>(sc, q, own, wref, s.fm);
                if (own && sc[F_N] == 1.0) e = NAN;   // the datum is its component's last member: the general step
                s.qloc[l] = q;
                st_remote_f64(ebuf0 + (uint32_t)k * 8u, e);
            }
            cluster_sync_all();                                                       // A: all weights are in CTA 0
            // ---- 3. CTA 0 draws (crpmm.py:75-78, utils.py:7-20) ----
            if (c == 0 && warp < 4) {
                const int kk = tid;
                double e = (kk < K) ? s.ebuf[kk] : (kk == K ? 1.0 : 0.0);
                double incl = e;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane == 31) sh.wtot[warp] = incl;
                asm volatile("bar.sync 1, 128;" ::: "memory");
                const double w0 = sh.wtot[0], w1 = sh.wtot[1], w2 = sh.wtot[2], w3 = sh.wtot[3];
                const double p1 = w0, p2 = w0 + w1, p3 = p2 + w2, tot = p3 + w3;
                const double t0 = s.ub[jj] * tot;
                const int hitw = (p1 > t0) ? 0 : (p2 > t0) ? 1 : (p3 > t0) ? 2 : (tot > t0) ? 3 : 4;
                int k_new = -1;
                double mg = 0.0;
                bool have = false;
                if (warp == hitw) {
                    const double pre = (warp == 0) ? 0.0 : (warp == 1) ? p1 : (warp == 2) ? p2 : p3;
                    double excl = __shfl_up_sync(0xffffffffu, incl, 1);
                    if (lane == 0) excl = 0.0;
                    const double upper = pre + incl, lower = pre + excl;
                    const unsigned who = __ballot_sync(0xffffffffu, upper > t0);
                    if (who != 0u && lane == __ffs(who) - 1) {
                        k_new = kk;
                        mg = margin_ratio(fmin(t0 - lower, upper - t0), tot);
                        have = true;
                    }
                }
                if (tid == 0 && hitw == 4) { k_new = K; mg = 0.0; have = true; }   // utils.py:20 fallback: the last index
                if (have) {
                    // anything but a stay or a plain move between two live components ends the launch: a birth, a draw
                    // inside the margin guard, an untrusted closed form / a component that would die (NaN), overflow
                    // (the code says why: 1 sum not finite / positive, 2 birth, 3 margin guard, 4 NaN weight)
                    const int rare = (tot != tot) ? 4 : (!(tot > 0.0) || !(tot < INFINITY)) ? 1 : (k_new >= K) ? 2 : (mg < p.guard) ? 3 : 0;
                    for (int t = 0; t < C; ++t) {
                        st_remote_f64(map_to_cta(&sh.res_mg, t), mg);
                        st_remote_u32(map_to_cta(&sh.res_k, t), (unsigned)k_new);
                        st_remote_u32(map_to_cta(&sh.res_rare, t), (unsigned)rare);
                    }
                }
            }
            cluster_sync_all();                                                       // B: the draw is in every CTA
            const int k_new = sh.res_k;
            if (sh.res_rare) { stop = true; break; }
            if (tid == 0 && c == 0) {
                sh.evals += K;
                sh.steps += 1;
                const unsigned long long mb = (unsigned long long)__double_as_longlong(sh.res_mg);
                if (mb < sh.margin_bits) sh.margin_bits = mb;
            }
            done = jj + 1;
            if (k_new == k_old) continue;   // stay: nothing was touched (crpmm.py:82-85)

            // ---- 5. the datum moves: the owners update their records (add_item / del_item as rank-one changes of S_N) ----
            for (int side = 0; side < 2; ++side) {
                const int k = side ? k_new : k_old;
                if (k % C != c) continue;       // uniform over the CTA
                const int l = k / C;
                double *B = s.rec + (size_t)l * R;
                double *sc = B + L::SC;
                const double *d = s.dsh + l * DP;   // m - x, still this datum's
                // v = B d: SEGS threads per row, CPS columns each
                {
                    const int a = tid / L::SEGS, sg = tid % L::SEGS;
                    double acc = 0.0;
#pragma unroll
                    for (int t = 0; t < L::CPS; ++t) {
                        const int b = sg * L::CPS + t;
                        acc = fma(bsym<DP>(B, a, b), d[b], acc);
                    }
#pragma unroll
                    for (int o = L::SEGS / 2; o >= 1; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (sg == 0) s.vbuf[a] = acc;
                }
                // scalars of the update and the count-table rows of the new count (thread 0)
                double cn0 = 0.0, cn1 = 0.0, g1 = 0.0, h1 = 0.0, b1 = 0.0, n2 = 0.0;
                if (tid == 0) {
                    const double n = sc[F_N];
                    n2 = n + (side ? 1.0 : -1.0);
                    const double *r0 = p.ntab + (size_t)((long long)n2 - 1) * NT_W;
                    cn0 = __ldg(r0 + NT_CN); cn1 = __ldg(r0 + NT_W + NT_CN); g1 = __ldg(r0 + NT_W + NT_G);
                    h1 = __ldg(r0 + NT_W + NT_H); b1 = __ldg(r0 + NT_W + NT_BETA);
                    const double beta = side ? sc[F_G] : sc[F_BETA];
                    const double den = side ? 1.0 + beta * s.qloc[l] : 1.0 - beta * s.qloc[l];
                    sh.gam = side ? -beta / den : beta / den;
                    sh.den = den;
                    sh.rk = (side ? -1.0 : 1.0) / (p.k0 + n2);   // m' = m -+ d / kappa(n2), d = m - x
                }
                __syncthreads();
                const double gam = sh.gam, rk = sh.rk;
                // rank-one update of the triangle in its storage order: index = j * NL + lane
                for (int idx = tid; idx < L::PP; idx += TB) {
                    const int ln = idx % NL, j = idx / NL;
                    int a, b;
                    if (j <= ln) { a = ln; b = j; } else { a = DP - 1 - ln; b = j - ln - 1; }
                    B[idx] = fma(gam * s.vbuf[a], s.vbuf[b], B[idx]);
                }
                __syncthreads();   // every reader of d is done before the mean changes
                if (tid < DP) B[L::MU + tid] = fma(d[tid], rk, B[L::MU + tid]);
                if (tid == 0) {
                    const double lds = sc[F_LDS] + fm::f_log(sh.den, s.fm);   // matrix determinant lemma
                    sc[F_N] = n2;
                    sc[F_LDS] = lds;
                    sc[F_CNT] = sc[F_CNT] + 1.0;
                    sc[F_CW] = cn1 - 0.5 * lds;
                    sc[F_G] = g1;
                    sc[F_H] = h1;
                    sc[F_BETA] = b1;
                    sc[F_CWO] = cn0 - 0.5 * lds;
                }
                __syncthreads();
            }
            if (c == 0 && tid == 0) {
                const long long i = s.ib[jj];
                __stcg(p.z_uid + i, s.uid_of_slot[k_new]);   // one writer, no replicas: the labels change in place
                __stcg(mlog + sh.n_log, make_int4((int)i, k_old, k_new, 0));   // the statistics follow from the log
                sh.n_log += 1;
                sh.moves += 1;
            }
        }
        pos += done;
    }

    // ---- epilogue ----
    __syncthreads();
    for (int l = tid; l < Lc; l += TB) __stcg(p.counts + (l * C + c), (long long)s.rec[(size_t)l * R + L::SC + F_N]);
    // the records as they are now, for the next launch (the host rebuilds them from the statistics every BIG_SPAN data or
    // after the general step changed anything)
    for (int l = 0; l < Lc; ++l) {
        double *dst = rec_in + (size_t)(l * C + c) * R;
        for (int e = tid; e < R; e += TB) __stcg(dst + e, s.rec[(size_t)l * R + e]);
    }
    if (c == 0 && tid == 0) {
        __stcg(&ctl->pos, pos);
        __stcg(&ctl->error, stop ? E_RARE : 0);
        __stcg(&ctl->win, sh.n_log);   // entries of the move log (Ctl::win is idle in this engine)
        if (stop) __stcg(&ctl->prof[sh.res_rare & 7], __ldcg(&ctl->prof[sh.res_rare & 7]) + 1);   // why it was handed back
        __stcg(&ctl->moves, __ldcg(&ctl->moves) + sh.moves);
        __stcg(&ctl->evals, __ldcg(&ctl->evals) + sh.evals);
        __stcg(&ctl->seq_data, __ldcg(&ctl->seq_data) + sh.steps);
        __stcg(&ctl->fast_steps, __ldcg(&ctl->fast_steps) + sh.steps);
        atomicMin(&ctl->margin_bits, sh.margin_bits);
    }
    cluster_sync_all();   // no CTA exits while another may still store into its shared memory
}

// ---------------------------------------------------------------------------------------------
// Sparse movers: which of the data at scan positions [pos0, pos1) provably STAY, given the records as they are now?
// Data-parallel over the whole GPU: a CTA takes a tile of WT consecutive positions; for every component the full
// symmetric B (DP x DP) is staged in shared memory once per tile and Y = B (m - X_tile) is a register-blocked product
// (thread = RPT rows x one datum, B broadcast across the warp: the FP64 pipe is the limit, not shared memory); then
// the weights, the cumulative sums in the reference's order and the draw (crpmm.py:70-78, utils.py:7-20).  A datum whose
// draw is its own component with a margin of at least Params::guard stays whatever arithmetic is used; the first
// position that does not is the atomicMin target Ctl::first -- the host hands the chain over to k_big_sweep there,
// and everything behind it is evaluated again later (its weights may change with that datum's move).
// Dynamic shared memory: big_window_smem<DP>().
// ---------------------------------------------------------------------------------------------
constexpr int WT = 64;   // positions per tile: every thread works on two of them (t and t + 32), so a staged element of B
                         // loaded from shared memory feeds two FMAs

template <int DP> constexpr size_t big_window_smem() {
    return sizeof(double) * ((size_t)DP * (DP + 2) + (size_t)DP * WT + (size_t)KCH * WT + 8 * WT + DP + NSC + fm::TAB_LEN);
}

template <int DP>
__global__ void __launch_bounds__(256) k_big_window(const Params p, const double *__restrict__ rec_in, int K, long long pos0,
                                                    long long pos1) {
    using L = BL<DP>;
    constexpr int NL = L::NL, R = L::R, RPT = DP / 8;
    constexpr int BS = DP + 2;              // row stride of the staged matrix: the transposed stores of the staging spread
                                            // over the banks (stride DP: 32-way conflicts), rows stay 16-byte aligned
    extern __shared__ __align__(16) double wsm[];
    double *Bt = wsm;                       // Bt[b * BS + a] = B[a][b]
    double *Xs = Bt + DP * BS;              // Xs[b * WT + t]
    double *E = Xs + DP * WT;               // E[k * WT + t]
    double *qp = E + KCH * WT;              // partial quadratic forms: qp[g * WT + t]
    double *mk = qp + 8 * WT;               // mean (DP) and scalars (NSC) of the component being evaluated
    double *fmt = mk + DP + NSC;
    __shared__ long long js[WT];
    __shared__ int kos[WT];
    __shared__ double us[WT], lps[WT];
    __shared__ int stop_s;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    Ctl *ctl = p.ctl;
    for (int e = tid; e < fm::TAB_LEN; e += 256) fmt[e] = __ldg(p.fmtab + e);
    const long long n_tiles = (pos1 - pos0 + WT - 1) / WT;
    unsigned long long margin_bits;
    { const double one = 1.0; margin_bits = (unsigned long long)__double_as_longlong(one); }
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long jb = pos0 + tile * WT;
        __syncthreads();
        if (tid == 0) stop_s = (jb > __ldcg(&ctl->first)) ? 1 : 0;   // an earlier position already ends the window
        __syncthreads();
        if (stop_s) break;
        if (tid < WT) {
            const long long j = jb + tid;
            const bool valid = j < pos1;
            long long i = 0;
            if (valid) i = p.order ? p.order[j] : j;
            js[tid] = valid ? i : -1;
            int ko = -1;
            if (valid) { const int uid = __ldcg(p.z_uid + i); ko = uid >= 0 ? __ldcg(p.slot_of_uid + uid) : -1; }
            kos[tid] = ko;
            us[tid] = valid ? p.u[j] : 0.0;
            lps[tid] = valid ? p.log_prior[i] : 0.0;
        }
        __syncthreads();
        for (int e = tid; e < DP * WT; e += 256) {
            const int t = e / DP, b = e % DP;
            const long long i = js[t];
            Xs[b * WT + t] = i >= 0 ? p.X[(size_t)i * DP + b] : 0.0;
        }
        // component k + 1's packed triangle, mean and scalars travel from L2 into registers while component k is being
        // multiplied: the staging below is shared-memory stores only
        constexpr int NPRE = (L::PP + 255) / 256;
        double pre[NPRE], premk = 0.0;
#pragma unroll
        for (int t = 0; t < NPRE; ++t) pre[t] = (K > 0 && tid + t * 256 < L::PP) ? __ldcg(rec_in + tid + t * 256) : 0.0;
        if (K > 0 && tid < DP + NSC) premk = __ldcg(rec_in + L::MU + tid);
        for (int k = 0; k < K; ++k) {
            __syncthreads();
            if ((k & 7) == 7) {   // an earlier position ended the window meanwhile: this tile's answer is not needed
                if (tid == 0) stop_s = (jb > __ldcg(&ctl->first)) ? 1 : 0;
                __syncthreads();
                if (stop_s) break;
            }
#pragma unroll
            for (int t = 0; t < NPRE; ++t) {   // the packed triangle (lane-interleaved rows, see pidx) -> full matrix
                const int e = tid + t * 256;
                if (e < L::PP) {
                    const int ln = e % NL, j = e / NL;
                    int a, b;
                    if (j <= ln) { a = ln; b = j; } else { a = DP - 1 - ln; b = j - ln - 1; }
                    Bt[b * BS + a] = pre[t];
                    Bt[a * BS + b] = pre[t];
                }
            }
            if (tid < DP + NSC) mk[tid] = premk;
            __syncthreads();
            if (k + 1 < K) {
                const double *rn = rec_in + (size_t)(k + 1) * R;
#pragma unroll
                for (int t = 0; t < NPRE; ++t) pre[t] = (tid + t * 256 < L::PP) ? __ldcg(rn + tid + t * 256) : 0.0;
                if (tid < DP + NSC) premk = __ldcg(rn + L::MU + tid);
            }
            double y0[RPT], y1[RPT];
#pragma unroll
            for (int r = 0; r < RPT; ++r) { y0[r] = 0.0; y1[r] = 0.0; }
#pragma unroll 4
            for (int b = 0; b < DP; ++b) {
                const double m = mk[b];
                const double d0 = m - Xs[b * WT + tx], d1 = m - Xs[b * WT + 32 + tx];
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    const double bv = Bt[b * BS + ty * RPT + r];
                    y0[r] = fma(bv, d0, y0[r]);
                    y1[r] = fma(bv, d1, y1[r]);
                }
            }
            double part0 = 0.0, part1 = 0.0;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const double m = mk[ty * RPT + r];
                part0 = fma(m - Xs[(ty * RPT + r) * WT + tx], y0[r], part0);
                part1 = fma(m - Xs[(ty * RPT + r) * WT + 32 + tx], y1[r], part1);
            }
            qp[ty * WT + tx] = part0;
            qp[ty * WT + 32 + tx] = part1;
            __syncthreads();
            if (ty < 2) {
                const int t = ty * 32 + tx;
                double q = 0.0;
#pragma unroll
                for (int g = 0; g < 8; ++g) q += qp[g * WT + t];
                const double *sc = mk + DP;
                const int own = (k == kos[t]) ? 1 : 0;
                double e = fast::f_finish_weight<1>(sc, q, own, p.log_alpha + lps[t], fmt);
                if (own && sc[F_N] == 1.0) e = NAN;   // the datum is its component's last member: not a plain stay
                E[k * WT + t] = e;
            }
        }
        __syncthreads();
        if (stop_s) break;
        if (tid < WT && js[tid] >= 0) {
            const int ko = kos[tid];
            double tot = 0.0;
            for (int k = 0; k < K; ++k) tot += E[k * WT + tid];
            tot += 1.0;   // the new component: exp(reference - reference)
            const double t0 = us[tid] * tot;
            double lower = 0.0, upper = 0.0;
            int kd = K;
            for (int k = 0; k < K; ++k) {
                upper = lower + E[k * WT + tid];
                if (upper > t0) { kd = k; break; }
                lower = upper;
            }
            bool stays = false;
            if (kd == ko && ko >= 0 && tot > 0.0 && tot < INFINITY) {
                const double gapw = fmin(t0 - lower, upper - t0);
                if (gapw >= p.guard * tot) {
                    stays = true;
                    const unsigned long long mb = (unsigned long long)__double_as_longlong(margin_ratio(gapw, tot));
                    if (mb < margin_bits) margin_bits = mb;
                }
            }
            if (!stays) atomicMin(&ctl->first, jb + tid);
        }
    }
    if (tid < WT) atomicMin(&ctl->margin_bits, margin_bits);
}

// ---------------------------------------------------------------------------------------------
// The bit-exact sufficient statistics from the move log: CTA k holds component k's statistics in registers (EPT elements
// per thread: S packed row-major, then num), scans the log in chain order and applies, for every move that touches k,
//     S_ab -= fl(x_a x_b), num_a -= x_a   (the datum leaves: del_item, gaussian_components.py:184-185)
//     S_ab += fl(x_a x_b), num_a += x_a   (it joins: add_item, :165-166)
// one rounded multiply and one rounded add per element, in the order the chain made the moves: the reference's bits.
// Matching entries are gathered a tile at a time and their rows of X fetched G at a time, so the row loads overlap.
// ---------------------------------------------------------------------------------------------
template <int DP> __global__ void __launch_bounds__(256) k_big_replay(const Params p, const int4 *__restrict__ mlog, long long n_log) {
    using L = fast::Lay<DP>;   // any padded D (bgmm_clu.cuh replays through this kernel too)
    constexpr int T = 256, NS = L::PP + DP, EPT = (NS + T - 1) / T, TILE = 1024, G = 8;
    __shared__ int4 ent[TILE];
    __shared__ int hit[TILE];          // matching entries of the tile: (index in tile) << 1 | joins
    __shared__ int nhit;
    __shared__ double xs[G][DP];
    __shared__ unsigned short rc[L::PP];
    const int k = blockIdx.x, tid = threadIdx.x;
    for (int e = tid; e < L::PP; e += T) {
        int a, b;
        decode_row_idx(e, a, b);
        rc[e] = (unsigned short)((a << 8) | b);
    }
    double *S = p.S + (size_t)k * L::PP, *num = p.num + (size_t)k * DP;
    double acc[EPT];
    int ea[EPT], eb[EPT];
    __syncthreads();
#pragma unroll
    for (int t = 0; t < EPT; ++t) {
        const int e = tid + t * T;
        acc[t] = 0.0; ea[t] = 0; eb[t] = -1;
        if (e < L::PP) { acc[t] = S[e]; ea[t] = rc[e] >> 8; eb[t] = rc[e] & 0xff; }
        else if (e < NS) { acc[t] = num[e - L::PP]; ea[t] = e - L::PP; eb[t] = -1; }
    }
    for (long long base = 0; base < n_log; base += TILE) {
        const int nt = (int)min((long long)TILE, n_log - base);
        __syncthreads();
        if (tid == 0) nhit = 0;
        for (int t = tid; t < nt; t += T) ent[t] = mlog[base + t];
        __syncthreads();
        if (tid == 0) {   // chain order is kept by a serial gather (a tile is 1024 entries; matches are ~2 %)
            int n = 0;
            for (int t = 0; t < nt; ++t) {
                const int4 m = ent[t];
                if (m.y == k) hit[n++] = t << 1;
                if (m.z == k) hit[n++] = (t << 1) | 1;
            }
            nhit = n;
        }
        __syncthreads();
        const int nh = nhit;
        for (int h0 = 0; h0 < nh; h0 += G) {
            const int ng = min(G, nh - h0);
            __syncthreads();
            for (int t = tid; t < ng * DP; t += T) {
                const int g = t / DP, a = t % DP;
                xs[g][a] = p.X[(size_t)ent[hit[h0 + g] >> 1].x * DP + a];
            }
            __syncthreads();
            for (int g = 0; g < ng; ++g) {
                const bool joins = hit[h0 + g] & 1;
#pragma unroll
                for (int t = 0; t < EPT; ++t) {
                    const double o = eb[t] >= 0 ? __dmul_rn(xs[g][ea[t]], xs[g][eb[t]]) : xs[g][ea[t]];
                    acc[t] = joins ? __dadd_rn(acc[t], o) : __dsub_rn(acc[t], o);
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < EPT; ++t) {
        const int e = tid + t * T;
        if (e < L::PP) S[e] = acc[t];
        else if (e < NS) num[e - L::PP] = acc[t];
    }
}

}  // namespace big
}  // namespace bgmm
