// bgmm_seq.cuh -- the register-resident sequential step of the B200 sweep engine (dense movers: cold chains, D = 2).
//
// Semantics: the per-datum loop of CRPMM / PCRPMM.collapsed_gibbs_sampler (igmm/crpmm.py:57-88, igmm/pcrpmm.py:93-131),
// one datum after the other, exactly as bgmm_fast.cuh's f_step resolves it -- this file is the common case of that step
// (a datum that stays, or moves between two live components) made short; everything else (a component dies, a birth,
// an unassigned datum, an untrusted closed form, overflow, a draw inside the margin guard, drift refresh) is handed to
// f_step unchanged.
//
// Why it exists: when most data move, the chain is one serial step per datum and the step's latency is the whole
// story.  f_step streams every live record through shared memory for every datum (K x 1.2 KB at D = 16: the 128 B/clk
// shared-memory port alone is ~1.4 k cycles per datum) and chains five CTA barriers through out-of-line phases.  Here
//   * the matrices B_k = S_N^-1 of all live components live (mostly) in REGISTERS for the whole run: thread (k, part)
//     of the CTA owns one of four blocks of B_k -- the two diagonal triangles LL / HH and the two row-halves of the
//     off-diagonal block (D = 16: 36 / 36 / 32 / 32 doubles, of which 24 sit in registers and the rest is read
//     from the shared-memory record: more would spill, and spills go to L2 here); `part` is warp-uniform
//     (warps 4 part .. 4 part + 3), lanes run over components: no divergence, and an evaluation reads only the mean
//     and x from shared memory (8-12 doubles per thread instead of 44);
//   * a datum that stays costs three barriers (partial sums, scan of the 4 x 32 choices, draw); a move adds two:
//     the four owners of each of the two touched components rebuild v = B d from their blocks (partial products
//     meet in shared memory), then update their registers (Sherman-Morrison, gaussian_components.py:161-166 /
//     :184-185 as a rank-one change of S_N), the mean and the scalars; no thread outside those eight does anything;
//   * the bit-exact statistics are updated by the writer CTA with L2 reductions, off the critical path, as before.
// The quadratic form is summed in the same order as f_quad_part16, so at D = 16 the weights are bit-identical to
// f_step's and handing a datum over never changes its draw.  Registers are written back to the shared-memory records
// whenever anything else needs them (rare paths, window mode, the end of the sweep).
#pragma once

namespace bgmm {
namespace fast {

constexpr int SEQ_KMAX = 128;   // the K + 1 choices sit one per thread in warps 0..3

// profile builds: a timeline of the step for the first warp of each part (lane 0 of warps 0, 4, 8, 12): cycles from the
// previous mark to this one, summed per (part, phase) in FSh::tprof
#ifdef BGMM_PROFILE
#define SEQ_T(idx)                                                         \
    do {                                                                   \
        if (lane == 0 && grp == 0) {                                       \
            const long long t_ = clock64();                                \
            sh.tprof[PART][idx] += t_ - tlast_;                       \
            tlast_ = t_;                                                   \
        }                                                                  \
    } while (0)
#define SEQ_T_DECL() long long tlast_ = clock64()
#else
#define SEQ_T(idx) do { } while (0)
#define SEQ_T_DECL() do { } while (0)
#endif

// geometry of block PART of the symmetric DP x DP matrix B (packed lower triangle, row-major: e = a (a + 1) / 2 + b)
//   L = [0, H), U = [H, DP);  part 0: triangle LL, part 1: triangle UU, part 2 / 3: rows [R0, R0 + RN) of the block UL
template <int DP, int PART> struct Blk {
    static constexpr int H = (DP + 1) / 2;
    static constexpr int NU = DP - H;
    static constexpr int R2 = (NU + 1) / 2;
    static constexpr bool TRI = PART <= 1;
    static constexpr int O = PART == 0 ? 0 : H;            // TRI: first row / column
    static constexpr int TN = PART == 0 ? H : NU;          // TRI: size
    static constexpr int R0 = PART == 2 ? H : H + R2;      // RECT: first row
    static constexpr int RN = PART == 2 ? R2 : NU - R2;    // RECT: rows (columns are L)
    static constexpr int NE = TRI ? TN * (TN + 1) / 2 : RN * H;
    // of the block's NE elements the first NR (in the block's own packed order) live in registers, the rest stay in
    // the shared-memory record: 24 doubles + the step's working set fit the 128-register budget of a 512-thread CTA
    // without spilling (local memory has almost no L1 here -- shared memory takes ~220 KB of the SM's 256 KB)
    static constexpr int RCAP = PART == 0 ? 20 : 24;
    static constexpr int NR = NE > RCAP ? RCAP : NE;
    static constexpr int NEA = NR > 0 ? NR : 1;
    // local packed index / global packed index (row-major lower triangle of the DP x DP matrix) of element (a, b)
    __host__ __device__ static constexpr int le(int a, int b) { return TRI ? a * (a + 1) / 2 + b : a * H + b; }
    __host__ __device__ static constexpr int ge(int a, int b) {
        return TRI ? (O + a) * (O + a + 1) / 2 + O + b : (R0 + a) * (R0 + a + 1) / 2 + b;
    }
    static constexpr int TNA = TN > 0 ? TN : 1;
    static constexpr int RNA = RN > 0 ? RN : 1;
};

// Warp w runs on scheduler w % 4.  The four parts are four different instruction streams (the block geometry is
// compile-time), and a scheduler that hosts a warp of every part would keep four streams in its instruction cache;
// with this mapping every scheduler hosts two parts (two warps of each): schedulers 0, 1 the triangles, 2, 3 the
// rectangles.  The four warps of a part hold the component groups 0..3 (32 components each).
__device__ __forceinline__ int seq_part_of_warp(int w) { return ((w & 3) >> 1) * 2 + ((w >> 2) & 1); }
__device__ __forceinline__ int seq_group_of_warp(int w) { return (w & 1) * 2 + (w >> 3); }

__device__ __forceinline__ void bar_sync_all() { asm volatile("bar.sync 0;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_front() { asm volatile("bar.sync 1, 128;" ::: "memory"); }   // warps 0..3

// element (a, b) of this thread's block: a register, or the shared-memory record (col = element 0 of the component,
// element-major rec[e * ST + k]).  a, b are compile-time after unrolling, so the choice costs nothing.
template <int DP, int PART, int ST>
__device__ __forceinline__ double bget(const double (&B)[Blk<DP, PART>::NEA], const double *__restrict__ col, int a, int b) {
    using G = Blk<DP, PART>;
    const int l = G::le(a, b);
    return l < G::NR ? B[l < G::NR ? l : 0] : col[G::ge(a, b) * ST];
}
template <int DP, int PART, int ST>
__device__ __forceinline__ void bset(double (&B)[Blk<DP, PART>::NEA], double *__restrict__ col, int a, int b, double v) {
    using G = Blk<DP, PART>;
    const int l = G::le(a, b);
    if (l < G::NR) B[l < G::NR ? l : 0] = v;
    else col[G::ge(a, b) * ST] = v;
}

// registers <-> shared-memory records: only the register-resident elements move
template <int DP, int PART, int ST>
__device__ __forceinline__ void seq_load_block(const double *__restrict__ col, double (&B)[Blk<DP, PART>::NEA]) {
    using G = Blk<DP, PART>;
    if constexpr (G::TRI) {
#pragma unroll
        for (int a = 0; a < G::TN; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b)
                if (G::le(a, b) < G::NR) B[G::le(a, b) < G::NR ? G::le(a, b) : 0] = col[G::ge(a, b) * ST];
    } else {
#pragma unroll
        for (int a = 0; a < G::RN; ++a)
#pragma unroll
            for (int b = 0; b < G::H; ++b)
                if (G::le(a, b) < G::NR) B[G::le(a, b) < G::NR ? G::le(a, b) : 0] = col[G::ge(a, b) * ST];
    }
}
template <int DP, int PART, int ST>
__device__ __forceinline__ void seq_store_block(double *__restrict__ col, const double (&B)[Blk<DP, PART>::NEA]) {
    using G = Blk<DP, PART>;
    if constexpr (G::TRI) {
#pragma unroll
        for (int a = 0; a < G::TN; ++a)
#pragma unroll
            for (int b = 0; b <= a; ++b)
                if (G::le(a, b) < G::NR) col[G::ge(a, b) * ST] = B[G::le(a, b) < G::NR ? G::le(a, b) : 0];
    } else {
#pragma unroll
        for (int a = 0; a < G::RN; ++a)
#pragma unroll
            for (int b = 0; b < G::H; ++b)
                if (G::le(a, b) < G::NR) col[G::ge(a, b) * ST] = B[G::le(a, b) < G::NR ? G::le(a, b) : 0];
    }
}

// this block's share of  sum_a d_a (sum_{b<a} B_ab d_b + B_aa d_a / 2),  d = m - x;  q = 2 * (sum over the four parts).
// Same operation order as f_quad_part16.
template <int DP, int PART, int ST>
__device__ __forceinline__ double seq_quad_part(const double (&B)[Blk<DP, PART>::NEA], const double *__restrict__ col,
                                                const double *__restrict__ mu, const double *__restrict__ x) {
    using G = Blk<DP, PART>;
    double q = 0.0;
    if constexpr (G::TRI) {
        double d[G::TNA];
#pragma unroll
        for (int a = 0; a < G::TN; ++a) d[a] = mu[(G::O + a) * ST] - x[G::O + a];
#pragma unroll
        for (int a = 0; a < G::TN; ++a) {
            double r = 0.0;
#pragma unroll
            for (int b = 0; b < a; ++b) r = fma(bget<DP, PART, ST>(B, col, a, b), d[b], r);
            r = fma(0.5 * bget<DP, PART, ST>(B, col, a, a), d[a], r);
            q = fma(d[a], r, q);
        }
    } else {
        double dl[G::H], dh[G::RNA];
#pragma unroll
        for (int b = 0; b < G::H; ++b) dl[b] = mu[b * ST] - x[b];
#pragma unroll
        for (int a = 0; a < G::RN; ++a) dh[a] = mu[(G::R0 + a) * ST] - x[G::R0 + a];
#pragma unroll
        for (int a = 0; a < G::RN; ++a) {
            double r = 0.0;
#pragma unroll
            for (int b = 0; b < G::H; ++b) r = fma(bget<DP, PART, ST>(B, col, a, b), dl[b], r);
            q = fma(dh[a], r, q);
        }
    }
    return q;
}

// this block's share of v = B d (d = m - x), into the three partial-product slots vp[slot * DP + a]:
//   slot 0: LL -> [0, H), UU -> [H, DP);  slot 1: part 2 -> [0, H) and its rows, part 3 -> its rows;
//   slot 2: part 3 -> [0, H).  v_a = (slot0 + slot1) + slot2 (slot 2 only for a < H).  Empty blocks write zeros.
template <int DP, int PART, int ST>
__device__ __forceinline__ void seq_partial_v(const double (&B)[Blk<DP, PART>::NEA], const double *__restrict__ col,
                                              const double *__restrict__ mu, const double *__restrict__ x,
                                              double *__restrict__ vp) {
    using G = Blk<DP, PART>;
    if constexpr (G::TRI) {
        double d[G::TNA];
#pragma unroll
        for (int a = 0; a < G::TN; ++a) d[a] = mu[(G::O + a) * ST] - x[G::O + a];
#pragma unroll
        for (int a = 0; a < G::TN; ++a) {
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < G::TN; ++b)
                acc = fma(a >= b ? bget<DP, PART, ST>(B, col, a, b) : bget<DP, PART, ST>(B, col, b, a), d[b], acc);
            vp[G::O + a] = acc;
        }
    } else {
        double dl[G::H], dh[G::RNA];
#pragma unroll
        for (int b = 0; b < G::H; ++b) dl[b] = mu[b * ST] - x[b];
#pragma unroll
        for (int a = 0; a < G::RN; ++a) dh[a] = mu[(G::R0 + a) * ST] - x[G::R0 + a];
#pragma unroll
        for (int a = 0; a < G::RN; ++a) {
            double acc = 0.0;
#pragma unroll
            for (int b = 0; b < G::H; ++b) acc = fma(bget<DP, PART, ST>(B, col, a, b), dl[b], acc);
            vp[DP + G::R0 + a] = acc;
        }
#pragma unroll
        for (int b = 0; b < G::H; ++b) {
            double acc = 0.0;
#pragma unroll
            for (int a = 0; a < G::RN; ++a) acc = fma(bget<DP, PART, ST>(B, col, a, b), dh[a], acc);
            vp[(PART == 2 ? DP : 2 * DP) + b] = acc;
        }
    }
}

// B += gam v v^T on this block, v combined from the partial-product slots
template <int DP, int PART, int ST>
__device__ __forceinline__ void seq_rank_one(double (&B)[Blk<DP, PART>::NEA], double *__restrict__ col,
                                             const double *__restrict__ vp, double gam) {
    using G = Blk<DP, PART>;
    auto vat = [&](int a) -> double {
        const double t = vp[a] + vp[DP + a];
        return (a < G::H) ? t + vp[2 * DP + a] : t;
    };
    if constexpr (G::TRI) {
        double v[G::TNA];
#pragma unroll
        for (int a = 0; a < G::TN; ++a) v[a] = vat(G::O + a);
#pragma unroll
        for (int a = 0; a < G::TN; ++a) {
            const double ga = gam * v[a];
#pragma unroll
            for (int b = 0; b <= a; ++b) bset<DP, PART, ST>(B, col, a, b, fma(ga, v[b], bget<DP, PART, ST>(B, col, a, b)));
        }
    } else {
        double vl[G::H], vh[G::RNA];
#pragma unroll
        for (int b = 0; b < G::H; ++b) vl[b] = vat(b);
#pragma unroll
        for (int a = 0; a < G::RN; ++a) vh[a] = vat(G::R0 + a);
#pragma unroll
        for (int a = 0; a < G::RN; ++a) {
            const double ga = gam * vh[a];
#pragma unroll
            for (int b = 0; b < G::H; ++b) bset<DP, PART, ST>(B, col, a, b, fma(ga, vl[b], bget<DP, PART, ST>(B, col, a, b)));
        }
    }
}

// 1 / a for a > 0 (normal range) without a division: single-precision seed, three Newton steps (~1 ulp)
__device__ __forceinline__ double seq_recip(double a) {
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)a));
    double r = (double)rf;
    r = fma(r, fma(-a, r, 1.0), r);
    r = fma(r, fma(-a, r, 1.0), r);
    r = fma(r, fma(-a, r, 1.0), r);
    return r;
}
// count-table rows n2 - 1 and n2 (n2 = n + 1: the datum joins; n2 = n - 1 with `leave`) into 8 doubles of shared
// memory, asynchronously: dst = [r0.CN, r0.G, r1.CN, r1.G, r1.H, r1.BETA, r1.RK, -]
__device__ __forceinline__ void seq_fetch_rows(const double *ntab, long long n_cur, double *dst_sh, bool leave = false) {
    const double *r0 = ntab + (size_t)(n_cur + (leave ? -2 : 0)) * NT_W;
    const uint32_t dst = smem_u32(dst_sh);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(r0) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16), "l"(r0 + NT_W) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 32), "l"(r0 + NT_W + 2) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 48), "l"(r0 + NT_W + 4) : "memory");
}

// TMA bulk reduction: global[0..n) += shared[0..n) (doubles), element-wise IEEE adds at the L2, asynchronous
__device__ __forceinline__ void seq_bulk_add(double *gdst, const double *ssrc, int n) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(ssrc)), "r"(n * 8)
                 : "memory");
}
// the bulk group committed `m` groups ago (m >= 1: the latest is 1) has completed, i.e. at most m - 1 of the most
// recent groups are still pending; nothing older than 8 groups is ever left pending
__device__ __forceinline__ void seq_bulk_wait_older(int m) {
    switch (m) {
        case 1: asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); break;
        case 2: asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); break;
        case 3: asm volatile("cp.async.bulk.wait_group 2;" ::: "memory"); break;
        case 4: asm volatile("cp.async.bulk.wait_group 3;" ::: "memory"); break;
        case 5: asm volatile("cp.async.bulk.wait_group 4;" ::: "memory"); break;
        case 6: asm volatile("cp.async.bulk.wait_group 5;" ::: "memory"); break;
        case 7: asm volatile("cp.async.bulk.wait_group 6;" ::: "memory"); break;
        default: asm volatile("cp.async.bulk.wait_group 7;" ::: "memory"); break;
    }
}
// all bulk groups of this thread have completed (their global writes are performed); then make this CTA's shared
// memory writes visible to the TMA for the next group
__device__ __forceinline__ void seq_bulk_wait() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// Sequential batches with the B blocks in registers, for the warps of part PART.  Every warp of the CTA calls its own
// instantiation; all take the same (CTA-uniform) decisions and meet at the same barriers.  Returns when the sweep is
// over, on error, when the engine should go to windows, or when K outgrows the one-choice-per-thread layout.
// ---------------------------------------------------------------------------------------------
template <int DP, int PART>
__device__ __noinline__ int f_seq_part(const Params &p, const FSmem<DP> &s, int seq) {
    using Ly = Lay<DP>;
    using G = Blk<DP, PART>;
    constexpr int ST = Ly::KS;
    extern __shared__ __align__(16) double smem_raw[];
    // everything on the step's path is addressed as smem_raw[compile-time offset + ...]: immediate offsets, no pointer
    // registers next to the 72 that hold the B block
    using O = SOff<DP>;
    double *rec = smem_raw;
    FSh &sh = *reinterpret_cast<FSh *>(smem_raw + O::SH);
    const double *xb = smem_raw + O::XB;
    const double *ub = smem_raw + O::UB, *lpb = smem_raw + O::LPB;
    const long long *ib = reinterpret_cast<const long long *>(smem_raw + O::IB);
    // the window evaluators' rows (ew rows 0..NWARP-1) are idle, and void, while the chain runs sequentially:
    double *psum = smem_raw + O::EW;                         // 3 x 128 partial sums of the quadratic forms,
    int *kob = reinterpret_cast<int *>(psum + 3 * SEQ_KMAX); // slots of the staged data's own components (SEQ_BATCH ints)
    double *qv = smem_raw + O::EW + NWARP * Ly::WS;          // f_step's row: the quadratic forms of this datum
    double *vpb = smem_raw + O::DV;                          // dv, vv, nt are contiguous: 2 x 3 x DP partial products
    double *gdb = psum + 3 * SEQ_KMAX + SEQ_BATCH / 2;       // (gam, den, -+1/kappa') of the two rank-one updates (6, pad 8)
    // count-table rows (cp.async targets, 16-B aligned): [0, 8) / [8, 16) of the component the datum may leave, by the
    // parity of the datum (a request that is never used must not land on the next datum's), [16, 24) of the one it joins
    double *ntb = gdb + 8;
    // the writer CTA's statistics deltas of a move: [parity of the move][leaves / joins][S packed (PP), num (DP)]
    constexpr bool BULK = (Ly::PP % 2 == 0) && (DP % 2 == 0);   // TMA bulk operands are 16-byte granules
    constexpr int NSP = Ly::PP + DP;
    double *dbuf = ntb + 24;
    // move counter of the last bulk group that touched each component (the issuing thread's bookkeeping)
    int *lastmv = reinterpret_cast<int *>(dbuf + 4 * NSP);
    int mvcount = 0;
    int mvpar = 0;
    const bool bulk = BULK && !(p.tune & 16);   // developer switch: bit 4 = per-thread reductions instead of the TMA
    const bool stats_on = !(p.tune & 64);       // developer switch: bit 6 = no statistics at all (timing experiments only)
    const double *fmtab = smem_raw + O::FM;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // warp -> (part, group of 32 components): see seq_part_of_warp; this thread's component is k
    const int grp = seq_group_of_warp(warp);
    const int k = grp * 32 + lane;
    if (tid < SEQ_KMAX) lastmv[tid] = -1000;
    double *col = rec + k;
    const double *mu = col + Ly::MU * ST;
    double *sc = col + Ly::SC * ST;

    double B[G::NEA];
    B[0] = 0.0;
    seq_load_block<DP, PART, ST>(col, B);
    SEQ_T_DECL();

    while (true) {
        const long long pos = sh.pos;
        if (pos >= p.N || sh.error != 0 || sh.mode != 0 || sh.K >= SEQ_KMAX) break;
        const int nb = (int)min((long long)SEQ_BATCH, p.N - pos);
        if (tid == 0) sh.seq_moves0 = sh.moves;
        bar_sync_all();   // everyone has read the loop state before thread 0 changes it again
        // ---- stage the batch (as f_run) ----
        for (int t = tid; t < nb * DP; t += TF) {
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
                kob[jj] = uid >= 0 ? s.slot_of_uid[uid] : -1;
            }
        }
        bar_sync_all();
        F_PROF(PH_STAGE);
        SEQ_T(0);   // staging + loop control
        int K = sh.K;
        int done = 0;
        for (int jj = 0; jj < nb; ++jj) {
            seq += 1;
            const int k_old = kob[jj];
            bool rare = (k_old < 0) || (K >= SEQ_KMAX);
            int k_new = -1;
            if (!rare) {
                const double *x = xb + jj * DP;
                if constexpr (PART == 0) {
                    // the count-table rows the component would need if the datum leaves it: requested now, used (if at
                    // all) at the end of the move
                    if (k == k_old) {
                        const long long n_cur = (long long)sc[F_N * ST];
                        if (n_cur >= 2) seq_fetch_rows(p.ntab, n_cur, ntb + (jj & 1) * 8, true);
                    }
                }
                // ---- phase A: the quadratic forms, four threads per component ----
                double pq = 0.0;
                if (k < K) pq = seq_quad_part<DP, PART, ST>(B, col, mu, x);
                if constexpr (PART > 0) psum[(PART - 1) * SEQ_KMAX + k] = pq;
                SEQ_T(1);   // head + phase A
                bar_sync_all();                                                        // #1
                SEQ_T(2);   // wait at #1
                if constexpr (PART == 0) {
                    // ---- finish the K + 1 weights, scan them, draw (crpmm.py:68-78, utils.py:7-20) ----
                    const double wref = p.log_alpha + lpb[jj];   // the new-table weight (crpmm.py:74) is the exp scale
                    double e = 0.0;
                    if (k < K) {
                        const double q = 2.0 * ((pq + psum[k]) + (psum[SEQ_KMAX + k] + psum[2 * SEQ_KMAX + k]));
                        const int own = (k == k_old) ? 1 : 0;
                        e = f_finish_weight<ST>(sc, q, own, wref, fmtab);
                        qv[k] = q;
                        // untrusted closed form, or the datum is its component's last member: the general step
                        if (e != e || (own && sc[F_N * ST] == 1.0)) sh.rare_seq = seq;
                    } else if (k == K) {
                        e = 1.0;
                    }
                    double incl = e;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const double t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    if (lane == 31) sh.wtot[grp] = incl;
                    SEQ_T(3);   // finish + scan
                    bar_sync_front();                                                  // #2 (warps 0..3)
                    SEQ_T(4);   // wait at #2
                    const double w0 = sh.wtot[0], w1 = sh.wtot[1], w2 = sh.wtot[2], w3 = sh.wtot[3];
                    const double p1 = w0, p2 = w0 + w1, p3 = p2 + w2, tot = p3 + w3;
                    const double t0 = ub[jj] * tot;
                    const int hitw = (p1 > t0) ? 0 : (p2 > t0) ? 1 : (p3 > t0) ? 2 : (tot > t0) ? 3 : 4;
                    if (grp == hitw) {
                        const double pre = (grp == 0) ? 0.0 : (grp == 1) ? p1 : (grp == 2) ? p2 : p3;
                        double excl = __shfl_up_sync(0xffffffffu, incl, 1);
                        if (lane == 0) excl = 0.0;
                        const double upper = pre + incl, lower = pre + excl;
                        const unsigned who = __ballot_sync(0xffffffffu, upper > t0);
                        if (who != 0u && lane == __ffs(who) - 1) {
                            sh.k_new = k;
                            sh.last_mg = margin_ratio(fmin(t0 - lower, upper - t0), tot);
                        }
                    }
                    if (tid == 0) {
                        if (hitw == 4) { sh.k_new = K; sh.last_mg = 0.0; }   // utils.py:20 fallback: the last index
                        if (!(tot > 0.0) || !(tot < INFINITY)) sh.rare_seq = seq;
                    }
                }
                SEQ_T(5);   // draw
                bar_sync_all();                                                        // #3
                SEQ_T(6);   // wait at #3
                F_PROF(PH_EVAL);
                k_new = sh.k_new;
                const double mg = sh.last_mg;
                // a birth, a draw inside the margin guard, or anything flagged above: the general step redoes the datum
                rare = (sh.rare_seq == seq) || (k_new >= K) || (mg < p.guard);
                if (!rare && tid == 0) {
                    sh.evals += K;
                    sh.fast_steps += 1;
                    const unsigned long long mb = (unsigned long long)__double_as_longlong(mg);
                    if (mb < sh.margin_bits) sh.margin_bits = mb;
                }
            }
            if (rare) {
                // registers back to the shared-memory records, the general step (bgmm_fast.cuh), registers again; the
                // general step reads and writes the statistics itself: outstanding bulk reductions complete first
                seq_store_block<DP, PART, ST>(col, B);
                if (bulk && p.writer && tid == 384) seq_bulk_wait();
                bar_sync_all();
                f_step<DP>(p, s, jj, seq);
                bar_sync_all();
                F_PROF(PH_RARE);
                if (sh.error) break;
                seq_load_block<DP, PART, ST>(col, B);
                K = sh.K;
                // deaths renumber slots: the staged data's own components again
                if (tid > jj && tid < nb) { const int uid = s.uidb[tid]; kob[tid] = uid >= 0 ? s.slot_of_uid[uid] : -1; }
                bar_sync_all();
                done = jj + 1;
                continue;
            }
            F_COUNT(PH_STEPS);
            done = jj + 1;
            if (k_new == k_old) continue;   // stay: nothing was touched (crpmm.py:82-85)

            // ---- the datum moves from k_old to k_new, both live (add_item / del_item, gaussian_components.py:154-186) ----
            // Nothing but the B block lives across barrier #4: what the second half needs is re-read from shared memory.
            const bool mine = (k == k_old) || (k == k_new);
            const int which = (k == k_new) ? 1 : 0;
            double *vp = vpb + which * 3 * DP;
            if (mine) {
                const double *x = xb + jj * DP;
                if constexpr (PART == 0) {
                    // count-table rows n2 - 1 and n2 of the component the datum joins, straight into shared memory
                    // (cp.async: an L2 round trip that costs no register; first needed for the scalars at the very end of
                    // the move).  Those of the component it leaves were requested at the head of the step.
                    if (which) seq_fetch_rows(p.ntab, (long long)sc[F_N * ST], ntb + 16);
                }
                if constexpr (PART == 1) {
                    // the scalars of the two rank-one updates, by the other triangle's owner (it has no rows to fetch):
                    // beta = kappa / (kappa -+ 1): BETA(n) for the removal, G(n) for the addition
                    const double beta = which ? sc[F_G * ST] : sc[F_BETA * ST];
                    const double sq = qv[k];
                    const double den = which ? 1.0 + beta * sq : 1.0 - beta * sq;
                    const double n2 = sc[F_N * ST] + (which ? 1.0 : -1.0);
                    // reciprocals without a division (its slow path is a call, and a call with the B block live would
                    // put part of the block on the stack; nor an L2 trip for 1 / kappa): single-precision seed, three
                    // Newton steps
                    const double rd = seq_recip(den);
                    const double rk = seq_recip(p.k0 + n2);
                    gdb[which * 3] = which ? -(beta * rd) : beta * rd;
                    gdb[which * 3 + 1] = den;
                    gdb[which * 3 + 2] = which ? -rk : rk;   // m' = m -+ d / kappa(n2), d = m - x
                }
                seq_partial_v<DP, PART, ST>(B, col, mu, x, vp);
            }
            if (bulk && stats_on) {
                if (p.writer && warp >= 8) {
                    // the bit-exact statistics change as two delta vectors in shared memory (warps 8..11: the component
                    // the datum leaves, 12..15: the one it joins): -+ fl(x_a x_b) and -+ x_a, the reference's operands
                    // (gaussian_components.py:165-166, :184-185); added to global memory by the TMA below
                    const double *xg = xb + jj * DP;
                    const int side = warp >= 12 ? 1 : 0;
                    double *db = dbuf + (mvpar * 2 + side) * NSP;
                    const unsigned short *rcs = s.rc;
                    for (int e = tid - (side ? 384 : 256); e < NSP; e += 128) {
                        double v;
                        if (e < Ly::PP) { const int a = rcs[e] >> 8, b = rcs[e] & 0xff; v = __dmul_rn(xg[a], xg[b]); }
                        else v = xg[e - Ly::PP];
                        db[e] = side ? v : -v;
                    }
                    // the TMA reads shared memory through the async proxy: every writing thread orders its own generic
                    // writes in front of it (the issuing thread's fence alone does not cover other threads' writes)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                }
            }
            SEQ_T(7);   // decision + first half of the move
            bar_sync_all();                                                            // #4
            SEQ_T(8);   // wait at #4
            if (mine) {
                seq_rank_one<DP, PART, ST>(B, col, vp, gdb[which * 3]);
                if constexpr (G::TRI) {
                    // this thread is the only writer of these means
                    const double *x = xb + jj * DP;
                    const double rk2 = gdb[which * 3 + 2];
#pragma unroll
                    for (int a = 0; a < G::TN; ++a) {
                        double *pm = col + (Ly::MU + G::O + a) * ST;
                        const double m = *pm;
                        *pm = fma(m - x[G::O + a], rk2, m);
                    }
                }
                if constexpr (PART == 0) {
                    const double n2 = sc[F_N * ST] + (which ? 1.0 : -1.0);
                    const double lds = sc[F_LDS * ST] + fm::f_log(gdb[which * 3 + 1], fmtab);   // matrix determinant lemma
                    const double cnt = sc[F_CNT * ST] + 1.0;
                    sc[F_N * ST] = n2;
                    sc[F_LDS * ST] = lds;
                    sc[F_CNT * ST] = cnt;
                    asm volatile("cp.async.wait_all;" ::: "memory");   // this thread's own copies
                    const double *nt = ntb + (which ? 16 : (jj & 1) * 8);   // r0: CN, G | r1: CN, G, H, BETA, RK
                    sc[F_CW * ST] = nt[2] - 0.5 * lds;
                    sc[F_G * ST] = nt[3];
                    sc[F_H * ST] = nt[4];
                    sc[F_BETA * ST] = nt[5];
                    sc[F_CWO * ST] = nt[0] - 0.5 * lds;
                    if (cnt >= (double)REFRESH_EVERY) { if (which == 0) sh.refresh_a = seq; else sh.refresh_b = seq; }
                    if (which == 1) sh.moves += 1;
                }
            }
            SEQ_T(9);   // second half of the move
            bar_sync_all();                                                            // #5
            SEQ_T(10);  // wait at #5
            if (!stats_on) {
                if (p.writer && tid == 384) __stcg(p.z_out + ib[jj], s.uid_of_slot[k_new]);
            } else if (bulk) {
                // One thread hands the four delta vectors to the TMA: element-wise IEEE round-to-nearest adds performed at
                // the L2 (cp.reduce.async.bulk .add.f64, SASS UBLKRED.ADD.F64.RN) -- the same bits as the reference's
                // `+=` / `-=`, and no warp waits for them: per-thread global reductions (RED.ADD.F64) in front of a CTA
                // barrier cost ~2 k cycles per move here, because the barrier waits for the warp's outstanding reductions.
                // Moves that touch one component stay ordered: the last group that touched either component has completed
                // before this one is issued (wait_group on its age; it usually is many steps old), and the parity buffers
                // keep a group's source intact until it has been read.
                if (p.writer && tid == 384) {
                    // order: the last group that touched either component has completed (it usually is many groups old)
                    seq_bulk_wait_older(mvcount - max(lastmv[k_old], lastmv[k_new]));
                    lastmv[k_old] = mvcount;
                    lastmv[k_new] = mvcount;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    const double *db = dbuf + mvpar * 2 * NSP;
                    seq_bulk_add(p.S + (size_t)k_old * Ly::PP, db, Ly::PP);
                    seq_bulk_add(p.num + (size_t)k_old * DP, db + Ly::PP, DP);
                    seq_bulk_add(p.S + (size_t)k_new * Ly::PP, db + NSP, Ly::PP);
                    seq_bulk_add(p.num + (size_t)k_new * DP, db + NSP + Ly::PP, DP);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    // the other parity's buffers are refilled by the next move: their group's source reads are done
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    if (p.tune & 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // developer switch: synchronous
                    __stcg(p.z_out + ib[jj], s.uid_of_slot[k_new]);   // the label (replicas keep reading the input copy)
                }
                mvpar ^= 1;
                mvcount += 1;
            } else if (p.writer && warp >= 8) {
                // D <= 2 (the statistics of a component are not a whole number of 16-byte granules): per-thread global
                // reductions (warps 8..11 the removal, 12..15 the addition), issued behind the move's last barrier
                const double *xg = s.xb + jj * DP;
                if (warp < 12) {
                    f_stats_axpy_inl<DP>(p, s.rc, k_old, xg, -1, 0, tid - 256, 128);
                } else {
                    f_stats_axpy_inl<DP>(p, s.rc, k_new, xg, +1, 0, tid - 384, 128);
                    if (tid == 384) __stcg(p.z_out + ib[jj], s.uid_of_slot[k_new]);
                }
            }
            F_PROF(PH_UPDATE);
            F_COUNT(PH_MOVES);
            const bool ra = (sh.refresh_a == seq), rb = (sh.refresh_b == seq);
            if (ra || rb) {
                // drift control: the record(s) again from the bit-exact statistics (needs the shared-memory records)
                seq_store_block<DP, PART, ST>(col, B);
                if (bulk && p.writer && tid == 384) seq_bulk_wait();
                bar_sync_all();
                const double n_a = ra ? rec[(Ly::SC + F_N) * ST + k_old] : 0.0;
                const double n_b = rec[(Ly::SC + F_N) * ST + k_new];
                f_refresh<DP>(p, s, ra ? k_old : -1, n_a, rb ? k_new : -1, n_b);
                seq_load_block<DP, PART, ST>(col, B);
                // the rebuild used the scratch that holds the staged data's own components: look them up again
                if (tid > jj && tid < nb) { const int uid = s.uidb[tid]; kob[tid] = uid >= 0 ? s.slot_of_uid[uid] : -1; }
                bar_sync_all();
                F_PROF(PH_RARE);
            }
        }
        bar_sync_all();
        if (tid == 0) {
            const long long mv = sh.moves - sh.seq_moves0;
            sh.gap = 0.5 * sh.gap + 0.5 * (double)nb / ((double)mv + 0.5);
            sh.seq_data += done;
            sh.pos = sh.pos + done;
            if (p.engine == 0 && sh.gap >= (double)p.gap_to_win) sh.mode = 1;
            sh.win = f_next_window(sh.gap, sh.pos, p.N, p.win_factor);
        }
        bar_sync_all();
    }
    // registers back to the shared-memory records; the window evaluators' cached rows are void
    seq_store_block<DP, PART, ST>(col, B);
    if (tid == 0) { sh.ver += 1; sh.dall_ver = sh.ver; }
    if (bulk && p.writer && tid == 384) seq_bulk_wait();   // the deltas' buffers are the window evaluators' rows
    bar_sync_all();
    return seq;
}

// dispatch by warp (seq_part_of_warp)
template <int DP> __device__ __forceinline__ int f_seq_run(const Params &p, const FSmem<DP> &s, int seq) {
    switch (seq_part_of_warp(threadIdx.x >> 5)) {
        case 0: return f_seq_part<DP, 0>(p, s, seq);
        case 1: return f_seq_part<DP, 1>(p, s, seq);
        case 2: return f_seq_part<DP, 2>(p, s, seq);
        default: return f_seq_part<DP, 3>(p, s, seq);
    }
}

}  // namespace fast
}  // namespace bgmm
