// bgmm_clu.cuh -- the cluster step engine: dense movers, full covariance (NIW) components, padded D <= 16.
//
// Semantics: the per-datum loop of CRPMM / PCRPMM.collapsed_gibbs_sampler (igmm/crpmm.py:57-88, igmm/pcrpmm.py:93-131),
// strictly sequential over the scan order, for the common case of a step (the datum stays, or moves between two live
// components).  Anything else (a birth, a death, a draw inside the margin guard, a weight that is not finite) ends the
// launch at that datum, which the generic engine's step resolves (bgmm_ops.cuh k_resolve_one), like bgmm_big.cuh.
//
// When most data move, the chain is one long dependency: weights of datum j -> draw -> rank-one changes of two components
// -> weights of datum j + 1.  bgmm_seq.cuh runs that chain inside ONE CTA (replicated 148 times): 16 warps share K ~ 100
// components, six or seven quadratic forms deep per warp, CTA barriers between the phases -- 4.5 us per mover at D = 16.
// Here a thread-block CLUSTER of C CTAs runs it with ONE WARP PER COMPONENT and no barrier of any kind inside the loop:
//   * component k lives in the registers of warp k / C of CTA k % C: B = S_N^-1 (full matrix, row a in lanes a G .. a G + G - 1,
//     CP = D / G columns each), the mean, the scalars, and the count-table rows its next change will need (prefetched);
//   * a producer warp per CTA keeps a ring of the next 32 .. 64 data (row of X, uniform, log prior, current component) in
//     shared memory, handed over through mbarriers (full / empty per half);
//   * every component warp evaluates its quadratic form and weight and sends exp(weight - reference) straight into the
//     leader CTA's choice buffer: st.async through DSMEM, completing bytes on the leader's mbarrier;
//   * the draw warp of the leader CTA waits for the K weights, scans, draws (utils.py:7-20) and sends the result word into
//     every CTA the same way; it also owns the labels, the move log and the counters;
//   * the warps of the two touched components apply the rank-one change of B (Sherman-Morrison), the mean and the scalars
//     in registers -- v = B d is already there from the evaluation -- while everybody else is already evaluating the
//     next datum.
// The critical path of a step is: quadratic form + log / exp (one warp) -> DSMEM -> scan of <= 128 choices (one warp) ->
// DSMEM -> rank-one update (one warp).  The bit-exact statistics follow from the move log (big::k_big_replay), records are
// rebuilt from them at every launch (k_clu_prep) and a launch ends when a component has taken REFRESH_CAP rank-one changes,
// which bounds the drift of the incrementally updated records like REFRESH_EVERY does in bgmm_fast.cuh.
#pragma once
#include "bgmm_big.cuh"

namespace bgmm {
namespace clu {

using fast::F_N; using fast::F_LDS; using fast::F_CNT; using fast::F_CW; using fast::F_G; using fast::F_H;
using fast::F_BETA; using fast::F_CWO; using fast::NSC;
using fast::NT_CN; using fast::NT_G; using fast::NT_H; using fast::NT_BETA; using fast::NT_RK; using fast::NT_W;
using big::cluster_rank; using big::cluster_sync_all; using big::map_to_cta;

constexpr int KCH = 128;             // choices (K + 1) the draw warp holds: 4 per lane
constexpr int RS = 64;               // ring slots: two halves of 32 data
constexpr int E_RARE = 2;            // internal: the datum at Ctl::pos needs the general step
constexpr int REFRESH_CAP = 2048;    // rank-one changes of one component per launch
#ifndef BGMM_CLU_C
#define BGMM_CLU_C 16
#endif

template <int DP> struct CL {
    static constexpr int PP = DP * (DP + 1) / 2;
    static constexpr int MU = PP, SC = PP + DP, R = PP + DP + NSC;
    static constexpr int C = BGMM_CLU_C;                 // CTAs per cluster
    static constexpr int WPC = KCH / C;                  // component warps per CTA
    static constexpr int NW = WPC + 2;                   // + the producer warp + the draw warp (active in the leader)
    static constexpr int TB = NW * 32;
    static constexpr int G = (32 / DP < DP) ? 32 / DP : DP;   // lanes per row of B
    static constexpr int CP = DP / G;                    // columns per lane
    static constexpr int ACT = DP * G;                   // lanes that hold matrix elements
    static_assert(G * CP == DP && ACT <= 32, "rows must tile a warp");
};

// ---- mbarrier / DSMEM primitives ----
__device__ __forceinline__ void mbar_init(void *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(void *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ unsigned mbar_try_wait(void *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}
// the phase was completed by st.async writes of another CTA of the cluster: acquire at cluster scope
__device__ __forceinline__ unsigned mbar_try_wait_cluster(void *bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok;
}
// 8 bytes into the shared memory of a CTA of the cluster, completing 8 bytes on that CTA's mbarrier
__device__ __forceinline__ void st_async_u64(uint32_t remote_addr, unsigned long long v, uint32_t remote_bar) {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(remote_addr), "l"(v),
                 "r"(remote_bar) : "memory");
}

template <int DP> struct alignas(16) CSh {
    double ebuf[KCH];                 // leader: exp(weight - reference) of every choice
    double xr[RS][DP];                // ring: rows of X
    double ur[RS], lpr[RS];           //       the uniform of the step, log prior of the datum
    long long ir[RS];                 //       datum index
    int kor[RS];                      //       the component it sits in (-1: unassigned)
    int uidk[KCH];                    // leader: uid of every live slot (labels are uids)
    double fm[fm::TAB_LEN];
    unsigned long long full[2], empty[2];   // ring hand-over (producer -> consumers -> producer)
    unsigned long long ebar;          // leader: K weights have arrived
    unsigned long long rbar;          // every CTA: the result word has arrived
    unsigned long long res;           // k_new | rare << 16 | stop_after << 24
    volatile int stop;
};

// ---------------------------------------------------------------------------------------------
// records of all live components from the bit-exact statistics: one warp per component (fast::f_exact_record_warp:
// Cholesky of S_N, inverse, scalars from the count table); layout [B packed row-major | mean | scalars].
// Dynamic shared memory: PP + DP * DP + DP + R doubles + PP shorts.
// ---------------------------------------------------------------------------------------------
template <int DP> __global__ void k_clu_prep(const Params p, int K, double *__restrict__ rec_out, int *err) {
    using L = CL<DP>;
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
    __syncwarp();
    double *out = rec_out + (size_t)k * L::R;
    for (int e = lane; e < L::R; e += 32) out[e] = tmp[e];
}

template <int DP> constexpr size_t clu_prep_smem() {
    using L = CL<DP>;
    return sizeof(double) * (L::PP + DP * DP + DP + L::R) + sizeof(unsigned short) * ((L::PP + 7) & ~7);
}

// ---------------------------------------------------------------------------------------------
// the sweep kernel: one cluster per chain; data [p.start_pos, pos_limit)
// ---------------------------------------------------------------------------------------------
template <int DP>
__global__ void __launch_bounds__(CL<DP>::TB, 1) k_clu_sweep(const Params p, const double *__restrict__ rec_in, long long pos_limit,
                                                             int4 *__restrict__ mlog) {
    using L = CL<DP>;
    constexpr int C = L::C, G = L::G, CP = L::CP, R = L::R;
    __shared__ CSh<DP> S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Ctl *ctl = p.ctl;
    const int c = (int)cluster_rank();
    const int K = __ldcg(&ctl->K);
    const long long j0 = p.start_pos, j1 = pos_limit;
    const int n_comp_warps = (K > c) ? min(L::WPC, (K - c + C - 1) / C) : 0;   // live component warps of this CTA
    const int n_cons = n_comp_warps + (c == 0 ? 1 : 0);                        // warps that read the ring

    // ---- prologue ----
    for (int e = tid; e < fm::TAB_LEN; e += L::TB) S.fm[e] = __ldg(p.fmtab + e);
    for (int e = tid; e < KCH; e += L::TB) { S.ebuf[e] = 0.0; S.uidk[e] = e < K ? __ldcg(p.uid_of_slot + e) : -1; }
    if (tid == 0) {
        S.stop = 0;
        S.res = 0ull;
        mbar_init(&S.full[0], 1); mbar_init(&S.full[1], 1);
        mbar_init(&S.empty[0], (unsigned)max(n_cons, 1)); mbar_init(&S.empty[1], (unsigned)max(n_cons, 1));
        mbar_init(&S.ebar, 1);
        mbar_init(&S.rbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arrive_expect_tx(&S.rbar, 8);          // phase 0 of the result barrier
    }
    __syncthreads();
    cluster_sync_all();   // every CTA's barriers and buffers exist before anyone stores into them remotely

    const bool is_comp = warp < n_comp_warps;
    const bool is_prod = warp == L::WPC && n_cons > 0;
    const bool is_draw = warp == L::WPC + 1 && c == 0;
    const long long total = j1 - j0;

    if (is_prod) {
        // ---- producer: the ring of upcoming data ----
        SpinWatch wd;
        bool quit = false;
        for (long long ch = 0; ch * 32 < total && !quit; ++ch) {
            const int hf = (int)(ch & 1);
            const unsigned par = (unsigned)(((ch >> 1) & 1) ^ 1);
            while (!mbar_try_wait(&S.empty[hf], par)) {
                if (S.stop) { quit = true; break; }
                wd.poll(ctl, 11);
                __nanosleep(64);
            }
            if (quit) break;
            const long long j = j0 + ch * 32 + lane;
            const bool valid = j < j1;
            long long i = 0;
            if (valid) i = p.order ? __ldg(p.order + j) : j;
            const int slot = hf * 32 + lane;
            if (valid) {
                const int uid = __ldcg(p.z_uid + i);
                S.ir[slot] = i;
                S.kor[slot] = uid >= 0 ? __ldcg(p.slot_of_uid + uid) : -1;
                S.ur[slot] = __ldg(p.u + j);
                S.lpr[slot] = __ldg(p.log_prior + i);
            }
            constexpr int RPI = 32 / DP;   // rows per iteration
#pragma unroll 4
            for (int r0 = 0; r0 < 32; r0 += RPI) {
                const int r = r0 + lane / DP, a = lane % DP;
                const long long ir = __shfl_sync(0xffffffffu, i, r & 31);
                const int vr = __shfl_sync(0xffffffffu, (int)valid, r & 31);
                if (vr && lane < RPI * DP) S.xr[hf * 32 + r][a] = __ldg(p.X + (size_t)ir * DP + a);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.full[hf]);
        }
    } else if (is_comp) {
        // ---- component warp: component k in registers ----
        const int k = warp * C + c;
        const int a = (lane < L::ACT) ? lane / G : 0, h = (lane < L::ACT) ? lane % G : 0;
        const bool act = lane < L::ACT;
        const double *rk_g = rec_in + (size_t)k * R;
        double B[CP], mb[CP];
#pragma unroll
        for (int t = 0; t < CP; ++t) {
            const int b = h * CP + t;
            const int hi = a > b ? a : b, lo = a > b ? b : a;
            B[t] = act ? __ldcg(rk_g + hi * (hi + 1) / 2 + lo) : 0.0;
            mb[t] = __ldcg(rk_g + L::MU + b);
        }
        double ma = __ldcg(rk_g + L::MU + a);
        double sc[NSC];
#pragma unroll
        for (int t = 0; t < NSC; ++t) sc[t] = __ldcg(rk_g + L::SC + t);
        const long long nt_len = (p.N + 1) * NT_W;
        auto fetch_rows = [&](double n) -> double {   // lane l: word l of the count-table rows n - 2 .. n + 1
            const long long idx = ((long long)n - 2) * NT_W + lane;
            return (idx >= 0 && idx < nt_len) ? __ldg(p.ntab + idx) : 0.0;
        };
        double trow = fetch_rows(sc[F_N]);
        const uint32_t e_dst = map_to_cta(&S.ebuf[k], 0), e_bar = map_to_cta(&S.ebar, 0);
        const bool armer = (warp == 0);
        SpinWatch wd;
        for (long long j = j0; j < j1; ++j) {
            const long long s = j - j0;
            const int slot = (int)(s & (RS - 1));
            if ((s & 31) == 0) {
                const long long ch = s >> 5;
                while (!mbar_try_wait(&S.full[ch & 1], (unsigned)((ch >> 1) & 1))) wd.poll(ctl, 12);
            }
            // d = m - x, v = B d, q = d' B d
            const double xa = S.xr[slot][a];
            const double lp = S.lpr[slot];
            const int ko = S.kor[slot];
            double db[CP];
#pragma unroll
            for (int t = 0; t < CP; ++t) db[t] = mb[t] - S.xr[slot][h * CP + t];
            const double da = ma - xa;
            double v = 0.0;
#pragma unroll
            for (int t = 0; t < CP; ++t) v = fma(B[t], db[t], v);
#pragma unroll
            for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            double q = act ? da * v : 0.0;
#pragma unroll
            for (int o = G; o < 32; o <<= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
            const int own = (k == ko) ? 1 : 0;
            const double wref = p.log_alpha + lp;
            double e = fast::f_finish_weight<1>(sc, q, own, wref, S.fm);
            if (own && sc[F_N] == 1.0) e = NAN;   // the datum is its component's last member: the general step
            if (lane == 0) st_async_u64(e_dst, (unsigned long long)__double_as_longlong(e), e_bar);
            if ((s & 31) == 31 || j == j1 - 1) {   // done with this half of the ring
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.empty[(s >> 5) & 1]);
            }
            // the draw
            while (!mbar_try_wait_cluster(&S.rbar, (unsigned)(s & 1))) wd.poll(ctl, 13);
            const unsigned long long res = *(volatile unsigned long long *)&S.res;
            __syncwarp();
            if (armer && lane == 0) mbar_arrive_expect_tx(&S.rbar, 8);   // next phase
            const int k_new = (int)(res & 0xffffu), rare = (int)((res >> 16) & 0xffu), stop_after = (int)((res >> 24) & 1u);
            if (rare) break;
            if (k_new != ko && (k == ko || k == k_new)) {
                // the datum moves and this component is one of the two: del_item / add_item as a rank-one change of S_N
                // (gaussian_components.py:154-186) -- Sherman-Morrison on B, the determinant lemma on log|S_N|
                const int side = (k == k_new) ? 1 : 0;
                const double beta = side ? sc[F_G] : sc[F_BETA];
                const double den = side ? fma(beta, q, 1.0) : fma(-beta, q, 1.0);
                const double rd = fast::seq_recip(den);
                const double gam = side ? -(beta * rd) : beta * rd;
                const int base = side ? 16 : 0;   // rows (n2 - 1, n2) of the new count n2 = n -+ 1 inside trow
                const double cn0 = __shfl_sync(0xffffffffu, trow, base + NT_CN);
                const double cn1 = __shfl_sync(0xffffffffu, trow, base + 8 + NT_CN);
                const double g1 = __shfl_sync(0xffffffffu, trow, base + 8 + NT_G);
                const double h1 = __shfl_sync(0xffffffffu, trow, base + 8 + NT_H);
                const double b1 = __shfl_sync(0xffffffffu, trow, base + 8 + NT_BETA);
                const double rkk = __shfl_sync(0xffffffffu, trow, base + 8 + NT_RK);
                const double rk = side ? -rkk : rkk;   // m' = m -+ d / kappa(n2), d = m - x
                const double gv = gam * v;
#pragma unroll
                for (int t = 0; t < CP; ++t) {
                    const double vb = __shfl_sync(0xffffffffu, v, ((h * CP + t) * G) & 31);
                    B[t] = fma(gv, vb, B[t]);
                    mb[t] = fma(db[t], rk, mb[t]);
                }
                ma = fma(da, rk, ma);
                const double n2 = sc[F_N] + (side ? 1.0 : -1.0);
                const double lds = sc[F_LDS] + fm::f_log(den, S.fm);
                sc[F_N] = n2;
                sc[F_LDS] = lds;
                sc[F_CW] = cn1 - 0.5 * lds;
                sc[F_G] = g1;
                sc[F_H] = h1;
                sc[F_BETA] = b1;
                sc[F_CWO] = cn0 - 0.5 * lds;
                trow = fetch_rows(n2);
            }
            if (stop_after) break;
        }
        if (armer && lane == 0) S.stop = 1;
        if (lane == 0) __stcg(p.counts + k, (long long)sc[F_N]);
    } else if (is_draw) {
        // ---- the draw warp of the leader CTA ----
        uint32_t r_dst = 0, r_bar = 0;
        const int n_cta = min(C, K);            // CTAs that own components
        if (lane < n_cta) { r_dst = map_to_cta(&S.res, lane); r_bar = map_to_cta(&S.rbar, lane); }
        int cnt[4] = {0, 0, 0, 0};
        long long moves = 0, steps = 0, n_log = 0, pos = j1;
        unsigned long long margin_bits;
        { const double one = 1.0; margin_bits = (unsigned long long)__double_as_longlong(one); }
        int why = 0;
        bool stopped = false, pending_stop = false;
        SpinWatch wd;
        for (long long j = j0; j < j1; ++j) {
            const long long s = j - j0;
            const int slot = (int)(s & (RS - 1));
            if (lane == 0) mbar_arrive_expect_tx(&S.ebar, (unsigned)K * 8u);
            if ((s & 31) == 0) {
                const long long ch = s >> 5;
                while (!mbar_try_wait(&S.full[ch & 1], (unsigned)((ch >> 1) & 1))) wd.poll(ctl, 14);
            }
            const double u = S.ur[slot];
            const int ko = S.kor[slot];
            const long long i = S.ir[slot];
            while (!mbar_try_wait_cluster(&S.ebar, (unsigned)(s & 1))) wd.poll(ctl, 15);
            // crpmm.py:75-78, utils.py:7-20: the first choice whose cumulative weight exceeds u * total
            double e[4];
            asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(e[0]), "=d"(e[1]) : "r"(smem_u32(&S.ebuf[4 * lane])) : "memory");
            asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(e[2]), "=d"(e[3]) : "r"(smem_u32(&S.ebuf[4 * lane + 2])) : "memory");
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int idx = 4 * lane + t;
                if (idx == K) e[t] = 1.0;
                else if (idx > K) e[t] = 0.0;
            }
            const double run = (e[0] + e[1]) + (e[2] + e[3]);
            double incl = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            double excl = __shfl_up_sync(0xffffffffu, incl, 1);
            if (lane == 0) excl = 0.0;
            const double tot = __shfl_sync(0xffffffffu, incl, 31);
            const double t0 = u * tot;
            int cand = -1;
            double lower = excl, upper = excl;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                if (cand < 0 && 4 * lane + t <= K) {
                    upper = lower + e[t];
                    if (upper > t0) cand = 4 * lane + t;
                    else lower = upper;
                }
            }
            const unsigned who = __ballot_sync(0xffffffffu, cand >= 0);
            // margin of the draw: distance of the target to the nearest boundary of the drawn interval; inside the guard
            // (relative to the total) the datum goes through the exact path.  No cumulative weight exceeds the target:
            // utils.py:20 falls back to the last index -- a birth, the general step's business either way.
            const double gapw = fmin(t0 - lower, upper - t0);
            int pick = (cand >= 0) ? (cand | ((gapw >= p.guard * tot) ? 0 : 0x10000)) : 0;
            int k_new = K, in_guard = 0, src = 0;
            if (who != 0u) {
                src = __ffs(who) - 1;
                pick = __shfl_sync(0xffffffffu, pick, src);
                k_new = pick & 0xffff;
                in_guard = pick >> 16;
            }
            // anything but a stay or a plain move between two live components ends the launch (the code says why:
            // 1 sum not finite / positive, 2 birth, 3 margin guard, 4 NaN weight, 5 unassigned datum)
            const int rare = (ko < 0) ? 5 : (tot != tot) ? 4 : (!(tot > 0.0) || !(tot < INFINITY)) ? 1 : (k_new >= K) ? 2 : in_guard ? 3 : 0;
            const bool moved = !rare && k_new != ko;
            const int stop_after = (!rare && pending_stop) ? 1 : 0;
            const unsigned long long word = (unsigned long long)(unsigned)(k_new & 0xffff) | ((unsigned long long)rare << 16) |
                                            ((unsigned long long)stop_after << 24);
            if (lane < n_cta) st_async_u64(r_dst, word, r_bar);
            // ---- off the critical path from here ----
            double mg = 0.0;
            if (who != 0u) mg = __shfl_sync(0xffffffffu, margin_ratio(gapw, tot), src);
            if (moved) {   // a component that has taken its share of rank-one changes ends the launch one step later
                int over = 0;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int idx = 4 * lane + t;
                    if (idx == ko || idx == k_new) { cnt[t] += 1; if (cnt[t] >= REFRESH_CAP) over = 1; }
                }
                if (__any_sync(0xffffffffu, over)) pending_stop = true;
            }
            if ((s & 31) == 31 || j == j1 - 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&S.empty[(s >> 5) & 1]);
            }
            if (rare) { pos = j; why = rare; stopped = true; break; }
            steps += 1;
            {
                const unsigned long long mbits = (unsigned long long)__double_as_longlong(mg);
                if (mbits < margin_bits) margin_bits = mbits;
            }
            if (moved) {
                if (lane == 0) {
                    __stcg(p.z_uid + i, S.uidk[k_new]);   // one writer: the labels change in place
                    __stcg(mlog + n_log, make_int4((int)i, ko, k_new, 0)); // the statistics follow from the log
                }
                n_log += 1;
                moves += 1;
            }
            if (stop_after) { pos = j + 1; break; }
        }
        if (lane == 0) {
            __stcg(&ctl->pos, pos);
            __stcg(&ctl->error, stopped ? E_RARE : 0);
            __stcg(&ctl->win, n_log);   // entries of the move log (Ctl::win is idle in this engine)
            if (stopped) __stcg(&ctl->prof[why & 7], __ldcg(&ctl->prof[why & 7]) + 1);   // why it was handed back
            __stcg(&ctl->moves, __ldcg(&ctl->moves) + moves);
            __stcg(&ctl->evals, __ldcg(&ctl->evals) + steps * K);
            __stcg(&ctl->seq_data, __ldcg(&ctl->seq_data) + steps);
            __stcg(&ctl->fast_steps, __ldcg(&ctl->fast_steps) + steps);
            atomicMin(&ctl->margin_bits, margin_bits);
        }
    }
    __syncwarp();
    cluster_sync_all();   // no CTA exits while another may still store into its shared memory
}

}  // namespace clu
}  // namespace bgmm
