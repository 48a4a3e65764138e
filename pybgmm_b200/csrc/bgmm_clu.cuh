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
constexpr float PRE_MARGIN = 1e-4f;  // single-precision draw: accepted when the target is this far (x total) from a boundary
#ifndef BGMM_CLU_C
#define BGMM_CLU_C 16
#endif

#ifdef BGMM_PROFILE
#define CLU_TDECL long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long tlast = clock64()
#define CLU_T(slot) do { const long long now_ = clock64(); tacc[slot] += now_ - tlast; tlast = now_; } while (0)
#define CLU_TFLUSH(row) do { if (lane == 0) for (int t_ = 0; t_ < 8; ++t_) __stcg(&ctl->tprof[row][t_], __ldcg(&ctl->tprof[row][t_]) + tacc[t_]); } while (0)
#else
#define CLU_TDECL do { } while (0)
#define CLU_T(slot) do { } while (0)
#define CLU_TFLUSH(row) do { } while (0)
#endif

template <int DP> struct CL {
    static constexpr int PP = DP * (DP + 1) / 2;
    static constexpr int MU = PP, SC = PP + DP, R = PP + DP + NSC;
    static constexpr int C = BGMM_CLU_C;                 // CTAs per cluster: rank 0 draws, ranks 1 .. C - 1 hold the components
    static constexpr int CC = C - 1;                     // component CTAs
    static constexpr int WPC = (KCH + CC - 1) / CC;      // component warps per CTA
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
#ifdef BGMM_CLU_ACQ_CTA
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#else
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
#endif
    return ok;
}
// 8 bytes into the shared memory of a CTA of the cluster, completing 8 bytes on that CTA's mbarrier
// 1 / a for a normal positive a: the hardware's single-precision estimate (2^-22) and two Newton steps (2^-44, 2^-88)
__device__ __forceinline__ double recip2(double a) {
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)a));
    double r = (double)rf;
    r = fma(r, fma(-a, r, 1.0), r);
    r = fma(r, fma(-a, r, 1.0), r);
    return r;
}
// fm::f_log / fm::f_exp with the polynomials in Estrin form: the same tables, arguments and accuracy, four / two
// dependent operations fewer on the chain every step waits for
__device__ __forceinline__ double c_log(double x, const double *__restrict__ tab) {
    const long long bits = __double_as_longlong(x);
    const int e = (int)((bits >> 52) & 0x7ff) - 1023;
    const double m = __longlong_as_double((bits & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);  // [1, 2)
    const int i = (int)((bits >> (52 - 7)) & (fm::LOG_N - 1));
    const double inv = tab[2 * i], lc = tab[2 * i + 1];
    const double r = fma(m, inv, -1.0);                 // |r| <= 2^-8
    // log1p(r) = r + r^2 P(r),  P = -1/2 + r/3 - r^2/4 + r^3/5 - r^4/6 + r^5/7 - r^6/8
    const double r2 = r * r;
    const double pa = fma(r, 1.0 / 3.0, -1.0 / 2.0), pb = fma(r, 1.0 / 5.0, -1.0 / 4.0), pc = fma(r, 1.0 / 7.0, -1.0 / 6.0);
    const double r4 = r2 * r2;
    const double P = fma(r4, fma(r2, -1.0 / 8.0, pc), fma(r2, pb, pa));
    const double l1p = fma(r2, P, r);
    const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
    const double ed = (double)e;
    return fma(ed, LN2_HI, lc) + fma(ed, LN2_LO, l1p);
}
__device__ __forceinline__ double c_exp(double t, const double *__restrict__ tab) {
    const double INV = 9.23324826168936568e+01;          // 64 / ln 2
    const double C_HI = 1.08304244931787252e-02;         // ln2 / 64, high part (27 trailing zero bits)
    const double C_LO = 2.03070420217029510e-10;         //           low part
    const double kd = rint(t * INV);
    const int k = (int)kd;
    double r = fma(-kd, C_HI, t);
    r = fma(-kd, C_LO, r);                               // |r| <= ln2/128
    // exp(r) - 1 = r + r^2 Q(r),  Q = 1/2 + r/6 + r^2/24 + r^3/120 + r^4/720
    const double r2 = r * r;
    const double qa = fma(r, 1.0 / 6.0, 0.5), qb = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    const double Q = fma(r2, fma(r2, 1.0 / 720.0, qb), qa);
    const double pm1 = fma(r2, Q, r);
    const double sv = tab[2 * fm::LOG_N + (k & (fm::EXP_N - 1))];
    const double v = fma(sv, pm1, sv);                   // 2^(j/64) * exp(r)
    const int q = k >> 6;                                // floor division: k = 64 q + j
    const int q1 = q / 2, q2 = q - q1;
    const double s1 = __longlong_as_double((long long)(q1 + 1023) << 52);
    const double s2 = __longlong_as_double((long long)(q2 + 1023) << 52);
    return (v * s1) * s2;
}
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
    double vx[CL<DP>::WPC][DP];       // per component warp: v = B d of the datum in flight, for the rank-one update
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
    constexpr int C = L::C, CC = L::CC, G = L::G, CP = L::CP, R = L::R;
    __shared__ CSh<DP> S;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Ctl *ctl = p.ctl;
    const int c = (int)cluster_rank();
    const int K = __ldcg(&ctl->K);
    const long long j0 = p.start_pos, j1 = pos_limit;
    // component k: warp k / CC of CTA 1 + k % CC (the leader CTA keeps its SM for the draw warp)
    const int n_comp_warps = (c >= 1 && K > c - 1) ? min(L::WPC, (K - (c - 1) + CC - 1) / CC) : 0;   // live component warps here
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

    const long long total = j1 - j0;
    const bool is_comp = warp < n_comp_warps && total > 0;
    const bool is_prod = warp == L::WPC && n_cons > 0;
    const bool is_draw = warp == L::WPC + 1 && c == 0;

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
        // While the draw of datum s is in flight the warp prepares datum s + 1 under BOTH outcomes it can see: its
        // component unchanged, and changed by datum s (removed if s sits in it, added otherwise).  The changed
        // component is never formed: with v = B d_s, q_s = d_s' B d_s of datum s (kept from its own evaluation),
        //     B' = B + gam v v',   m' = m + rk d_s,   d' = d + rk d_s   (d = m - x_{s+1})
        //     sigma = v' d' = v' d + rk q_s
        //     d'' B' d' = d' B d + rk (2 v' d + rk q_s) + gam sigma^2,       B' d' = B d + (rk + gam sigma) v
        // so ONE matrix-vector product serves both outcomes, and the scalar tails (log, exp) run once with the two
        // outcomes in different lanes.  When the result arrives the matching weight is sent at once: the chain's
        // critical path holds no evaluation at all.
        const int k = warp * CC + (c - 1);
        const bool act = lane < L::ACT;
        const int a = act ? lane / G : 0, h = act ? lane % G : 0;
        const int sel = lane & 3;   // scalar tails: 0 the component as it is, 1 the other outcome, 2 the determinant lemma
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
        double n = __ldcg(rk_g + L::SC + F_N), lds = __ldcg(rk_g + L::SC + F_LDS);
        // window of the count table: rows wc - 8 .. wc + 7, lane l holds (l & 1 ? RK : CN) of row wc - 8 + l / 2
        const long long nt_rows = p.N + 1;
        auto fetch_win = [&](long long c0) -> double {
            const long long row = c0 - 8 + (lane >> 1);
            return (row >= 0 && row < nt_rows) ? __ldg(p.ntab + row * NT_W + ((lane & 1) ? NT_RK : NT_CN)) : 0.0;
        };
        long long wc = (long long)n, wc2 = 0;
        double tw = fetch_win(wc), tw2 = 0.0;
        int pend = 0;   // steps until the requested window tw2 replaces tw (the load has two steps to land)
        const double hh0 = 0.5 * (double)(p.v0 + 1);   // H(n) = (nu + D) / 2 = (v0 + 1 + n) / 2
        // everything in a record that depends on the count alone (bgmm_fast.cuh NT_*) for the counts n - 1, n, n + 1,
        // out of the window whenever n changes: cnw[i] = CN(n - 2 + i), rkw[i] = 1 / kappa(n - 2 + i)
        struct NS { double cn_n, cn_m, g, beta, hh; };
        double cnw[4], rkw[5];
        auto load_counts = [&]() {
            const int r = (int)((long long)n - (wc - 8));   // row of n inside the window
#pragma unroll
            for (int i = 0; i < 4; ++i) cnw[i] = __shfl_sync(0xffffffffu, tw, (2 * (r - 2 + i)) & 31);
#pragma unroll
            for (int i = 0; i < 5; ++i) rkw[i] = __shfl_sync(0xffffffffu, tw, (2 * (r - 2 + i) + 1) & 31);
        };
        // count nn = n + dn, dn in {-1, 0, 1}: G = kappa / (kappa + 1) = 1 - 1 / kappa(nn + 1), BETA = kappa / (kappa - 1)
        // = 1 + 1 / kappa(nn - 1)
        auto ns_of = [&](int dn) -> NS {
            NS o;
            o.cn_n = dn < 0 ? cnw[1] : dn == 0 ? cnw[2] : cnw[3];
            o.cn_m = dn < 0 ? cnw[0] : dn == 0 ? cnw[1] : cnw[2];
            o.g = 1.0 - (dn < 0 ? rkw[2] : dn == 0 ? rkw[3] : rkw[4]);
            o.beta = 1.0 + (dn < 0 ? rkw[0] : dn == 0 ? rkw[1] : rkw[2]);
            o.hh = hh0 + 0.5 * (n + (double)dn);
            return o;
        };
        load_counts();
        NS ns_st = ns_of(0);
        const uint32_t e_dst = map_to_cta(&S.ebuf[k], 0), e_bar = map_to_cta(&S.ebar, 0);
        const bool armer = (warp == 0);
        double *vx = S.vx[warp];
        SpinWatch wd;
        CLU_TDECL;
        // w = (B d)[row of the lane], q = d' B d, sg = v_s' d for the datum in ring slot `slot1` (d = m - x)
        auto quad = [&](double v_s, int slot1, double &q1, double &sg1, double &w) {
            double u0 = 0.0, u1 = 0.0;   // two partial chains
#pragma unroll
            for (int t = 0; t < CP; ++t) {
                const double d1 = mb[t] - S.xr[slot1][h * CP + t];
                if (t & 1) u1 = fma(B[t], d1, u1);
                else u0 = fma(B[t], d1, u0);
            }
            w = u0 + u1;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
            const double da1 = ma - S.xr[slot1][a];
            q1 = act ? da1 * w : 0.0;
            sg1 = act ? da1 * v_s : 0.0;
#pragma unroll
            for (int o = G; o < 32; o <<= 1) {
                q1 += __shfl_xor_sync(0xffffffffu, q1, o);
                sg1 += __shfl_xor_sync(0xffffffffu, sg1, o);
            }
        };

        // datum 0
        while (!mbar_try_wait(&S.full[0], 0u)) wd.poll(ctl, 12);
        double cur_q, cur_v;
        {
            double sg0;
            quad(0.0, 0, cur_q, sg0, cur_v);
            double sc[NSC];
            sc[F_N] = n; sc[F_LDS] = lds; sc[F_CNT] = 0.0;
            sc[F_CW] = ns_st.cn_n - 0.5 * lds; sc[F_CWO] = ns_st.cn_m - 0.5 * lds;
            sc[F_G] = ns_st.g; sc[F_BETA] = ns_st.beta; sc[F_H] = ns_st.hh;
            const int own0 = (k == S.kor[0]) ? 1 : 0;
            double e = fast::f_finish_weight<1>(sc, cur_q, own0, p.log_alpha + S.lpr[0], S.fm);
            if (own0 && n == 1.0) e = NAN;   // the datum is its component's last member: the general step
            if (lane == 0) st_async_u64(e_dst, (unsigned long long)__double_as_longlong(e), e_bar);
        }
        const int n_steps = (int)total;   // a launch covers at most BIG_SPAN data
        for (int s = 0; s < n_steps; ++s) {
            const int slot = s & (RS - 1), slot1 = (s + 1) & (RS - 1);
            const bool has_next = s + 1 < n_steps;
            const int ko = S.kor[slot];
            if (pend && --pend == 0) { tw = tw2; wc = wc2; }
            if ((s & 31) == 0 && s > 0 && lane == 0)   // done with the previous half of the ring (its last datum was committed)
                mbar_arrive(&S.empty[((s - 1) >> 5) & 1]);
            if (((s + 1) & 31) == 0 && has_next) {
                const int ch = (s + 1) >> 5;
                while (!mbar_try_wait(&S.full[ch & 1], (unsigned)((ch >> 1) & 1))) wd.poll(ctl, 12);
            }
            CLU_T(0);   // ring
            // The other outcome of datum s for this component: it left (side 0, s sits here) / it joined (side 1):
            // del_item / add_item as a rank-one change of S_N (gaussian_components.py:154-186), Sherman-Morrison on B,
            // the determinant lemma on log|S_N|.
            const int side = (k == ko) ? 0 : 1;
            const bool alt_ok = side || n > 1.0;
            const double beta = side ? ns_st.g : ns_st.beta;
            const double den = side ? fma(beta, cur_q, 1.0) : fma(-beta, cur_q, 1.0);
            const double rd = recip2(den);
            const double gam = side ? -(beta * rd) : beta * rd;
            const double n2 = alt_ok ? n + (side ? 1.0 : -1.0) : n;
            const double n_sel = (sel == 1) ? n2 : n;
            const int dn = alt_ok ? (side ? 1 : -1) : 0;
            const NS nsx = ns_of((sel == 1) ? dn : 0);
            const double rk = side ? -rkw[3] : rkw[1];   // m' = m -+ d_s / kappa(n2), d_s = m - x_s
            CLU_T(4);   // scalars of the other outcome
            // (past the last datum of the launch the ring slot holds stale data: evaluated like any other, never sent)
            double e_u, e_a, q_u, q_a, sg, w, al_lds;
            {
                double sg_u;
                quad(cur_v, slot1, q_u, sg_u, w);
                sg = fma(rk, cur_q, sg_u);
                q_a = fma(gam * sg, sg, fma(rk, fma(rk, cur_q, 2.0 * sg_u), q_u));
                const int own1 = (k == S.kor[slot1]) ? 1 : 0;
                const double wref1 = p.log_alpha + S.lpr[slot1];
                const double q_sel = (sel == 1) ? q_a : q_u;
                const double arg = own1 ? fma(-nsx.beta, q_sel, 1.0) : fma(nsx.g, q_sel, 1.0);
                const double x = (sel == 2) ? den : arg;
                CLU_T(5);   // quadratic forms
                // one logarithm for the three arguments (the two weights' and the determinant lemma's), one exponential
                const double Lg = c_log(x, S.fm);
                al_lds = lds + __shfl_sync(0xffffffffu, Lg, 2);
                CLU_T(6);   // logarithm
                const double lds_sel = (sel == 1) ? al_lds : lds;
                const double cc = (own1 ? nsx.cn_m : nsx.cn_n) - 0.5 * lds_sel, hh = own1 ? 1.0 - nsx.hh : nsx.hh;
                const double tt = (cc - hh * Lg) - wref1;
                double e = (tt < EXP_CUTOFF) ? 0.0 : c_exp(tt, S.fm);
                // the own component's closed form is not trusted / the datum is its last member: the general step
                if (own1 && (!(x > OM_MIN) || n_sel == 1.0)) e = NAN;
                e_u = __shfl_sync(0xffffffffu, e, 0);
                e_a = alt_ok ? __shfl_sync(0xffffffffu, e, 1) : NAN;
            }
            CLU_T(1);   // preparation of datum s + 1
            // the draw of datum s
            while (!mbar_try_wait_cluster(&S.rbar, (unsigned)(s & 1))) wd.poll(ctl, 13);
            const unsigned long long res = *(volatile unsigned long long *)&S.res;
            const int k_new = (int)(res & 0xffffu);
            const unsigned ending = (unsigned)(res >> 16) & 0x1ffu;   // rare code (the datum is not resolved here) | stop_after << 8
            const bool rare = (ending & 0xffu) != 0u;
            const bool touched = !rare && (k_new != ko) && (k == ko || k == k_new);
            if (ending == 0u && has_next && lane == 0)   // first thing after the result: the chain waits for this store
                st_async_u64(e_dst, (unsigned long long)__double_as_longlong(touched ? e_a : e_u), e_bar);
            __syncwarp();
            if (armer && lane == 0) mbar_arrive_expect_tx(&S.rbar, 8);   // next phase
            CLU_T(2);   // waiting for the draw
            if (touched) {
                // the component takes the other outcome: B += gam v v', m += rk (m - x_s)
                __syncwarp();
                if (act && h == 0) vx[a] = cur_v;
                __syncwarp();
                const double gv = gam * cur_v;
#pragma unroll
                for (int t = 0; t < CP; ++t) {
                    const int b = h * CP + t;
                    B[t] = fma(gv, vx[b], B[t]);
                    mb[t] = fma(mb[t] - S.xr[slot][b], rk, mb[t]);
                }
                ma = fma(ma - S.xr[slot][a], rk, ma);
                n = n2;
                lds = al_lds;
                load_counts();
                ns_st = ns_of(0);
                const long long n_now = (long long)n;
                if (pend == 0 && (n_now - wc >= 3 || wc - n_now >= 3)) { wc2 = n_now; tw2 = fetch_win(n_now); pend = 2; }
                CLU_T(3);   // commit
            }
            cur_v = touched ? fma(fma(gam, sg, rk), cur_v, w) : w;   // B' d' = B d + (rk + gam sigma) v
            cur_q = touched ? q_a : q_u;
            if (ending) break;
        }
        if (warp == 0 && c <= 2) CLU_TFLUSH(c);
        if (warp == n_comp_warps - 1 && c == 1) CLU_TFLUSH(3);
        if (armer && lane == 0) S.stop = 1;
        if (lane == 0) __stcg(p.counts + k, (long long)n);
    } else if (is_draw) {
        // ---- the draw warp of the leader CTA ----
        uint32_t r_dst = 0, r_bar = 0;
        const int n_cta = min(CC, K);           // CTAs that own components: ranks 1 .. n_cta
        if (lane < n_cta) { r_dst = map_to_cta(&S.res, lane + 1); r_bar = map_to_cta(&S.rbar, lane + 1); }
        int cnt[4] = {0, 0, 0, 0};
        long long moves = 0, steps = 0, n_log = 0, pos = j1, slow_draws = 0;
        unsigned long long margin_bits;
        { const double one = 1.0; margin_bits = (unsigned long long)__double_as_longlong(one); }
        int why = 0, refreshed = 0;
        bool stopped = false, pending_stop = false;
        SpinWatch wd;
        CLU_TDECL;
        for (long long j = j0; j < j1; ++j) {
            CLU_T(2);   // bookkeeping of the previous step
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
            CLU_T(0);   // waiting for the weights
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
            // First in single precision: prefix sums of <= 128 non-negative terms are good to ~1e-6 of the total, so a
            // target at least PRE_MARGIN (1e-4) of the total away from both boundaries of the interval it falls in picks
            // the same choice as any more accurate arithmetic -- the reference's included (utils.py:15-20).  One step in
            // ~5000 is closer than that (or overflows a float) and takes the double precision scan below.
            int k_new = K, rare = 0;
            double mg = 0.0;
            bool decided = false;
            if (!(p.tune & 128)) {
                const float f0 = (float)e[0], f1 = (float)e[1], f2 = (float)e[2], f3 = (float)e[3];
                float inclf = (f0 + f1) + (f2 + f3);
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const float t = __shfl_up_sync(0xffffffffu, inclf, o);
                    if (lane >= o) inclf += t;
                }
                float lowf = __shfl_up_sync(0xffffffffu, inclf, 1);
                if (lane == 0) lowf = 0.0f;
                const float totf = __shfl_sync(0xffffffffu, inclf, 31);
                const float t0f = (float)u * totf;
                int candf = -1;
                float upf = lowf;
                const float ff[4] = {f0, f1, f2, f3};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if (candf < 0 && 4 * lane + t <= K) {
                        upf = lowf + ff[t];
                        if (upf > t0f) candf = 4 * lane + t;
                        else lowf = upf;
                    }
                }
                const unsigned whof = __ballot_sync(0xffffffffu, candf >= 0);
                const float gapf = fminf(t0f - lowf, upf - t0f);
                int pickf = (candf >= 0 && gapf >= PRE_MARGIN * totf && totf < INFINITY) ? candf : -1;
                if (whof != 0u) {
                    const int srcf = __ffs(whof) - 1;
                    pickf = __shfl_sync(0xffffffffu, pickf, srcf);
                    if (pickf >= 0) {
                        decided = true;
                        k_new = pickf;
                        rare = (ko < 0) ? 5 : (k_new >= K) ? 2 : 0;
                        mg = (double)__shfl_sync(0xffffffffu, __fdividef(gapf, totf), srcf);   // a diagnostic (float accuracy)
                    }
                }
            }
            if (!decided) {
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
                // margin of the draw: distance of the target to the nearest boundary of the drawn interval; inside the
                // guard (relative to the total) the datum goes through the exact path.  No cumulative weight exceeds the
                // target: utils.py:20 falls back to the last index -- a birth, the general step's business either way.
                const double gapw = fmin(t0 - lower, upper - t0);
                int pick = (cand >= 0) ? (cand | ((gapw >= p.guard * tot) ? 0 : 0x10000)) : 0;
                int in_guard = 0;
                if (who != 0u) {
                    const int src = __ffs(who) - 1;
                    pick = __shfl_sync(0xffffffffu, pick, src);
                    k_new = pick & 0xffff;
                    in_guard = pick >> 16;
                    mg = __shfl_sync(0xffffffffu, margin_ratio(gapw, tot), src);
                }
                // anything but a stay or a plain move between two live components ends the launch (the code says why:
                // 1 sum not finite / positive, 2 birth, 3 margin guard, 4 NaN weight, 5 unassigned datum)
                rare = (ko < 0) ? 5 : (tot != tot) ? 4 : (!(tot > 0.0) || !(tot < INFINITY)) ? 1 : (k_new >= K) ? 2 : in_guard ? 3 : 0;
                slow_draws += 1;
            }
            const bool moved = !rare && k_new != ko;
            const int stop_after = (!rare && pending_stop) ? 1 : 0;
            const unsigned long long word = (unsigned long long)(unsigned)(k_new & 0xffff) | ((unsigned long long)rare << 16) |
                                            ((unsigned long long)stop_after << 24);
            if (lane < n_cta) st_async_u64(r_dst, word, r_bar);
            CLU_T(1);   // scan + draw
            // ---- off the critical path from here ----
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
            if (stop_after) { pos = j + 1; refreshed = 1; break; }
        }
        CLU_TFLUSH(0);
        if (lane == 0) {
            S.stop = 1;   // this CTA's producer
            __stcg(&ctl->pos, pos);
            __stcg(&ctl->error, stopped ? E_RARE : 0);
            __stcg(&ctl->win, n_log);   // entries of the move log (Ctl::win is idle in this engine)
            if (stopped) __stcg(&ctl->prof[why & 7], __ldcg(&ctl->prof[why & 7]) + 1);   // why it was handed back
            __stcg(&ctl->moves, __ldcg(&ctl->moves) + moves);
            __stcg(&ctl->evals, __ldcg(&ctl->evals) + steps * K);
            __stcg(&ctl->seq_data, __ldcg(&ctl->seq_data) + steps);
            __stcg(&ctl->fast_steps, __ldcg(&ctl->fast_steps) + steps);
            __stcg(&ctl->refreshes, __ldcg(&ctl->refreshes) + refreshed);   // the launch ended for a rebuild of the records
            __stcg(&ctl->prof[8], __ldcg(&ctl->prof[8]) + slow_draws);   // draws the single-precision scan did not decide
            atomicMin(&ctl->margin_bits, margin_bits);
        }
    }
    __syncwarp();
    cluster_sync_all();   // no CTA exits while another may still store into its shared memory
}

}  // namespace clu
}  // namespace bgmm
