// bgmm_sweep.cuh -- the persistent sweep kernel: one launch = one Gibbs sweep of one chain.
//
// Semantics reproduced: the per-datum loop of CRPMM.collapsed_gibbs_sampler (igmm/crpmm.py:57-88) and
// PCRPMM.collapsed_gibbs_sampler (igmm/pcrpmm.py:93-131) -- strictly sequential over the scan order.
//
// Execution (DESIGN.md "Engine"): every CTA of a cooperative grid speculatively evaluates a window of
// upcoming data against the *current* component records (thread per datum, records broadcast from shared
// memory).  A datum whose draw keeps it where it is leaves the state bit-for-bit unchanged (the
// reference's restore path, crpmm.py:82-85), so all "stay" decisions in front of the first datum that
// does anything else are exactly the sequential chain's decisions.  CTA 0 then resolves that first
// mover sequentially (full general path: deaths, births, explicit refactorisation), and the next window
// starts behind it.  When movers are frequent the kernel drops to a purely sequential run on CTA 0.
#pragma once
#include "bgmm_device.cuh"

namespace bgmm {

constexpr int SEQ_RUN = 1024;          // data per sequential run
constexpr double SEQ_GAP = 24.0;       // below this mean gap between movers the window path does not pay
constexpr long long WIN_MIN = 64;

struct Sh {
    long long i;
    long long n_old;
    int K;
    int k_old, k_new, uid;
    int need_explicit, died, error;
    int n_dirty, reload_all;
    int dirty[MAX_DIRTY];
    double margin;
    double u;
    // CTA 0's working copy of the control-block counters (loaded after barrier 1, stored before barrier 2)
    long long moves, births, deaths, evals, windows, seq_data, wasted, guard_hits;
    long long uextra;        // constrained sweep: uniforms consumed by re-draws so far
    unsigned long long margin_bits;
    double gap;
    int n_free;
    unsigned long long mbar;
    unsigned int mbar_phase;
};

struct Smem {
    double *rec;   // Kc * R
    double *w;     // K_max + 2
    double *x;     // DP
    double *A0;    // packed_len(D)
    double *A1;
    Sh *sh;
};

template <int DP, int COV> __host__ __device__ inline size_t smem_fixed_doubles(int D, int K_max) {
    size_t n = (size_t)(K_max + 2) + DP;
    if (COV == COV_FULL) n += 2 * (size_t)packed_len(D);
    n += (sizeof(Sh) + 7) / 8 + 2;
    return (n + 1) & ~(size_t)1;
}

template <int DP, int COV> __device__ inline Smem carve(double *base, const Params &p) {
    constexpr int R = rec_len(DP, COV);
    Smem s;
    s.rec = base;
    double *q = base + (size_t)p.Kc * R;
    s.w = q; q += p.K_max + 2;
    s.x = q; q += DP;
    if (COV == COV_FULL) { s.A0 = q; q += packed_len(p.D); s.A1 = q; q += packed_len(p.D); }
    else { s.A0 = s.A1 = nullptr; }
    q = (double *)(((uintptr_t)q + 15) & ~(uintptr_t)15);
    s.sh = (Sh *)q;
    return s;
}

__device__ __forceinline__ void decode_row_idx(int e, int &a, int &b) {
    a = (int)floorf((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
    while ((a + 1) * (a + 2) / 2 <= e) ++a;
    while (a * (a + 1) / 2 > e) --a;
    b = e - a * (a + 1) / 2;
}

__device__ __forceinline__ void add_dirty(Sh &sh, int k) {
    for (int t = 0; t < sh.n_dirty; ++t) if (sh.dirty[t] == k) return;
    if (sh.n_dirty < MAX_DIRTY) sh.dirty[sh.n_dirty++] = k;
    else sh.reload_all = 1;
}

// record pointer of slot k: the CTA's shared copy when resident, else global (read with ld.cg)
template <int DP, int COV> __device__ __forceinline__ const double *rec_ptr(const Params &p, const double *rec_s,
                                                                            int k, bool &cg) {
    constexpr int R = rec_len(DP, COV);
    if (k < p.Kc) { cg = false; return rec_s + (size_t)k * R; }
    cg = true;
    return p.rec + (size_t)k * R;
}

// ---------------------------------------------------------------------------------------------
// statistics updates (whole CTA).  sign=-1: del_item's "-= X[i], -= outer" (gaussian_components.py:184-185),
// sign=+1: add_item's "+=" (:165-166).  The operand is fl(x_a*x_b) (the reference's precomputed
// _cached_outer, :116-118), so multiply and add are rounded separately.
// ---------------------------------------------------------------------------------------------
template <int DP, int COV> __device__ void stats_axpy(const Params &p, int slot, const double *x, int sign) {
    constexpr int SS = stat_len(DP, COV);
    const int D = p.D;
    double *num = p.num + (size_t)slot * DP;
    double *S = p.S + (size_t)slot * SS;
    if (COV == COV_FIXED) {
        // mu_N_numerators +-= precision * X[i]; precision_Ns +-= precision   (gaussian_components_fixedvar.py:157-158, :177-178)
        for (int a = threadIdx.x; a < D; a += blockDim.x) {
            const double tx = __dmul_rn(p.tau[a], x[a]);
            const double v = __ldcg(num + a), t = __ldcg(S + a);
            __stcg(num + a, sign > 0 ? __dadd_rn(v, tx) : __dsub_rn(v, tx));
            __stcg(S + a, sign > 0 ? __dadd_rn(t, p.tau[a]) : __dsub_rn(t, p.tau[a]));
        }
        return;
    }
    for (int a = threadIdx.x; a < D; a += blockDim.x) {
        const double v = __ldcg(num + a);
        __stcg(num + a, sign > 0 ? __dadd_rn(v, x[a]) : __dsub_rn(v, x[a]));
    }
    if (COV == COV_FULL) {
        const int P = packed_len(D);
        for (int e = threadIdx.x; e < P; e += blockDim.x) {
            int a, b;
            decode_row_idx(e, a, b);
            const double o = __dmul_rn(x[a], x[b]);
            const double v = __ldcg(S + e);
            __stcg(S + e, sign > 0 ? __dadd_rn(v, o) : __dsub_rn(v, o));
        }
    } else {
        for (int a = threadIdx.x; a < D; a += blockDim.x) {
            const double o = __dmul_rn(x[a], x[a]);
            const double v = __ldcg(S + a);
            __stcg(S + a, sign > 0 ? __dadd_rn(v, o) : __dsub_rn(v, o));
        }
    }
}

// new component initialised with the prior: num = k_0 m_0, S = S_0 + k_0 outer(m_0)   (gaussian_components.py:161-164)
template <int DP, int COV> __device__ void stats_init_prior(const Params &p, int slot) {
    constexpr int SS = stat_len(DP, COV);
    const int D = p.D;
    double *num = p.num + (size_t)slot * DP;
    double *S = p.S + (size_t)slot * SS;
    if (COV == COV_FIXED) {   // precision_0 * mu_0, precision_0   (gaussian_components_fixedvar.py:155-156)
        for (int a = threadIdx.x; a < DP; a += blockDim.x) {
            __stcg(num + a, a < D ? __dmul_rn(p.S0[a], p.m0[a]) : 0.0);
            __stcg(S + a, a < D ? p.S0[a] : 0.0);
        }
        return;
    }
    for (int a = threadIdx.x; a < DP; a += blockDim.x) __stcg(num + a, a < D ? __dmul_rn(p.k0, p.m0[a]) : 0.0);
    if (COV == COV_FULL) {
        const int P = packed_len(D);
        for (int e = threadIdx.x; e < SS; e += blockDim.x) {
            double v = 0.0;
            if (e < P) {
                int a, b;
                decode_row_idx(e, a, b);
                v = __dadd_rn(p.S0[e], __dmul_rn(p.k0, __dmul_rn(p.m0[a], p.m0[b])));
            }
            __stcg(S + e, v);
        }
    } else {
        for (int a = threadIdx.x; a < DP; a += blockDim.x)
            __stcg(S + a, a < D ? __dadd_rn(p.S0[a], __dmul_rn(p.k0, __dmul_rn(p.m0[a], p.m0[a]))) : 0.0);
    }
}

template <int DP, int COV> __device__ void stats_copy(const Params &p, int dst, int src) {
    constexpr int SS = stat_len(DP, COV);
    for (int a = threadIdx.x; a < DP; a += blockDim.x)
        __stcg(p.num + (size_t)dst * DP + a, __ldcg(p.num + (size_t)src * DP + a));
    for (int e = threadIdx.x; e < SS; e += blockDim.x)
        __stcg(p.S + (size_t)dst * SS + e, __ldcg(p.S + (size_t)src * SS + e));
}

// del_component (gaussian_components.py:188-205): swap-with-last.  The O(N) relabel (:199) becomes a uid-table update.
template <int DP, int COV> __device__ void delete_component(const Params &p, double *rec_s, Sh &sh, int k) {
    constexpr int R = rec_len(DP, COV);
    const int L = sh.K - 1;
    __syncthreads();
    if (k != L) {
        stats_copy<DP, COV>(p, k, L);
        bool cg;
        const double *src = rec_ptr<DP, COV>(p, rec_s, L, cg);
        for (int e = threadIdx.x; e < R; e += blockDim.x) {
            const double v = cg ? __ldcg(src + e) : src[e];
            __stcg(p.rec + (size_t)k * R + e, v);
            if (k < p.Kc) rec_s[(size_t)k * R + e] = v;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int uid_dead = __ldcg(p.uid_of_slot + k);
        if (k != L) {
            const int uid_l = __ldcg(p.uid_of_slot + L);
            __stcg(p.uid_of_slot + k, uid_l);
            __stcg(p.slot_of_uid + uid_l, k);
            __stcg(p.counts + k, __ldcg(p.counts + L));
            add_dirty(sh, k);
        }
        __stcg(p.counts + L, 0LL);
        __stcg(p.uid_of_slot + L, -1);
        __stcg(p.slot_of_uid + uid_dead, -1);
        const int nf = sh.n_free;
        __stcg(p.uid_free + nf, uid_dead);
        sh.n_free = nf + 1;
        sh.K = L;
        sh.deaths += 1;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Sequential resolution of the datum at scan position j by the whole CTA (general path).
// ---------------------------------------------------------------------------------------------
template <int DP, int COV>
__device__ void resolve(const Params &p, const Smem &sm, long long j) {
    constexpr int R = rec_len(DP, COV);
    Sh &sh = *sm.sh;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const long long i = p.order ? p.order[j] : j;
        sh.i = i;
        const int uid = __ldcg(p.z_uid + i);
        sh.uid = uid;
        int k_old = -1;
        long long n_old = 0;
        if (uid >= 0) { k_old = __ldcg(p.slot_of_uid + uid); n_old = __ldcg(p.counts + k_old); }
        sh.k_old = k_old; sh.n_old = n_old;
        sh.need_explicit = 0; sh.died = 0;
        if (p.status) {
            // constrained sweep: the uniforms are one stream, consumed in order (one per datum plus one per re-draw)
            if (j + sh.uextra >= p.u_len) { sh.error = -8; sh.u = 0.5; }
            else sh.u = p.u[j + sh.uextra];
        } else {
            sh.u = p.u[j];
        }
    }
    __syncthreads();
    if (sh.error) return;
    const long long i = sh.i;
    if (tid < DP) sm.x[tid] = p.X[(size_t)i * DP + tid];
    const int k_old = sh.k_old;
    const long long n_old = sh.n_old;
    __syncthreads();
    if (k_old >= 0 && n_old == 1) {  // del_item empties the component (gaussian_components.py:179-181)
        delete_component<DP, COV>(p, sm.rec, sh, k_old);
        if (tid == 0) { sh.died = 1; __stcg(p.z_uid + i, -1); }
        __syncthreads();
    }
    const bool died = sh.died != 0;
    int K = sh.K;
    double x[DP];
#pragma unroll
    for (int a = 0; a < DP; ++a) x[a] = sm.x[a];

    // Margin guard: when the uniform lands too close to a boundary of the drawn interval for the closed-form own
    // weight (Params::guard), the datum is redone with the own component removed explicitly (the reference's own
    // del_item arithmetic, gaussian_components.py:171-186) -- pass 1.
    bool expl = false;
    for (int pass = 0; pass < 2; ++pass) {
        // weights of the live components (crpmm.py:68-74)
        for (int k = tid; k < K; k += blockDim.x) {
            bool cg;
            const double *rec = rec_ptr<DP, COV>(p, sm.rec, k, cg);
            double w;
            if (k == k_old && !died) {
                bool ok = true;
                w = cg ? weight_own_removed<DP, COV, true>(rec, x, p, &ok) : weight_own_removed<DP, COV, false>(rec, x, p, &ok);
                if (!ok) sh.need_explicit = 1;
            } else {
                w = cg ? weight_other<DP, COV, true>(rec, x, p.D) : weight_other<DP, COV, false>(rec, x, p.D);
            }
            sm.w[k] = w;
        }
        if (tid == 0) sm.w[K] = p.log_alpha + p.log_prior[i];
        __syncthreads();

        if (sh.need_explicit != 0 && !expl) {
            expl = true;
            // explicit del_item: save the statistics (scratch slot K_max), remove, refactor, re-evaluate the column
            stats_copy<DP, COV>(p, p.K_max, k_old);
            __syncthreads();
            stats_axpy<DP, COV>(p, k_old, sm.x, -1);
            if (tid == 0) __stcg(p.counts + k_old, n_old - 1);
            __syncthreads();
            if (warp == 0) {
                const bool okf = refactor_warp<DP, COV>(p, p.num + (size_t)k_old * DP, p.S + (size_t)k_old * stat_len(DP, COV),
                                                        n_old - 1, 0, p.rec + (size_t)k_old * R,
                                                        k_old < p.Kc ? sm.rec + (size_t)k_old * R : nullptr, sm.A0);
                if (!okf && lane == 0) sh.error = -4;
            }
            __syncthreads();
            if (tid == 0) {
                bool cg;
                const double *rec = rec_ptr<DP, COV>(p, sm.rec, k_old, cg);
                sm.w[k_old] = cg ? weight_other<DP, COV, true>(rec, x, p.D) : weight_other<DP, COV, false>(rec, x, p.D);
                add_dirty(sh, k_old);
            }
            __syncthreads();
        }

        // logsumexp + draw (crpmm.py:75-78, utils.py:7-20) by warp 0 over the K+1 weights
        if (warp == 0) {
            const int n = K + 1;
            const int per = (n + 31) / 32;
            const int lo = lane * per, hi = min(n, lo + per);
            double M = -INFINITY;
            for (int k = lo; k < hi; ++k) M = fmax(M, sm.w[k]);
    #pragma unroll
            for (int o = 16; o > 0; o >>= 1) M = fmax(M, __shfl_xor_sync(0xffffffffu, M, o));
            double sl = 0.0;
            for (int k = lo; k < hi; ++k) {
                const double dlt = sm.w[k] - M;
                const double e = (dlt < EXP_CUTOFF) ? 0.0 : exp(dlt);
                sm.w[k] = e;
                sl += e;
            }
            double inc = sl;
    #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const double v = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += v;
            }
            const double s = __shfl_sync(0xffffffffu, inc, 31);
            double t = sh.u * s - (inc - sl);
            int cand = 0x7fffffff;
            double marg = 1.0;
            for (int k = lo; k < hi; ++k) {
                const double tb = t;
                t -= sm.w[k];
                if (t < 0.0) { cand = k; marg = fmin(fabs(tb), -t) / s; break; }
            }
            int best = cand;
    #pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            const unsigned who = __ballot_sync(0xffffffffu, cand == best && cand != 0x7fffffff);
            if (best == 0x7fffffff) {
                if (lane == 0) { sh.k_new = K; sh.margin = 0.0; }  // utils.py:20 fallback: the last index
            } else if (lane == (int)(__ffs(who) - 1)) {
                sh.k_new = best; sh.margin = marg;
            }
            if (!(s > 0.0) || !(s < INFINITY)) { if (lane == 0) sh.error = -4; }
        }
        __syncthreads();
        if (pass == 0 && !expl && k_old >= 0 && !died && sh.margin < p.guard && sh.error == 0) {
            __syncthreads();
            if (tid == 0) { sh.need_explicit = 1; sh.guard_hits += 1; }
            __syncthreads();
            continue;
        }
        break;
    }
    if (p.status && sh.error == 0) {
        // CSCRPMM's constrained re-draw (cscrpmm.py:342-350, :455-461): a datum whose old slot was a non-useful cluster when
        // the sweep started draws again -- the same probabilities, a fresh uniform each time (utils.py:15) -- until the slot
        // drawn was a useful one; any other datum goes back to its old slot index.  sm.w holds the unnormalised
        // probabilities of the draw above.
        if (warp == 0) {
            const int st_old = (k_old >= 0 && k_old < p.n_status) ? p.status[k_old] : 0;
            if (st_old == 2) {
                const int n = K + 1;
                const int per = (n + 31) / 32;
                const int lo = lane * per, hi = min(n, lo + per);
                double sl = 0.0;
                for (int k = lo; k < hi; ++k) sl += sm.w[k];
                double inc = sl;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const double v = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += v;
                }
                const double s = __shfl_sync(0xffffffffu, inc, 31);
                long long extra = sh.uextra;
                int k_sel = -1;
                while (true) {
                    extra += 1;
                    const long long ui = j + extra;
                    if (ui >= p.u_len) { if (lane == 0) sh.error = -8; break; }
                    const double u2 = p.u[ui];
                    double t = u2 * s - (inc - sl);
                    int cand = 0x7fffffff;
                    for (int k = lo; k < hi; ++k) {
                        t -= sm.w[k];
                        if (t < 0.0) { cand = k; break; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
                    k_sel = (cand == 0x7fffffff) ? K : cand;   // utils.py:20 fallback: the last index
                    if (k_sel < p.n_status && p.status[k_sel] == 1) break;
                }
                if (lane == 0) { sh.uextra = extra; sh.k_new = k_sel < 0 ? K : k_sel; }
            } else if (lane == 0) {
                sh.k_new = k_old;
            }
        }
        __syncthreads();
        if (sh.error) return;
    }
    const int k_new = sh.k_new;
    if (tid == 0) {
        sh.evals += K;
        const unsigned long long mb = (unsigned long long)__double_as_longlong(sh.margin);
        if (mb < sh.margin_bits) sh.margin_bits = mb;
    }
    const bool stay = (k_new == k_old) && !died;
    if (stay) {
        if (expl) {  // bitwise restore (crpmm.py:82-85), then the record follows from the restored statistics
            stats_copy<DP, COV>(p, k_old, p.K_max);
            if (tid == 0) __stcg(p.counts + k_old, n_old);
            __syncthreads();
            if (warp == 0)
                refactor_warp<DP, COV>(p, p.num + (size_t)k_old * DP, p.S + (size_t)k_old * stat_len(DP, COV), n_old, 0,
                                       p.rec + (size_t)k_old * R, k_old < p.Kc ? sm.rec + (size_t)k_old * R : nullptr,
                                       sm.A0);
            __syncthreads();
        }
        return;
    }
    // ---- the datum moves: add_item (gaussian_components.py:154-169) ----
    const bool remove_now = (k_old >= 0) && !died && !expl;
    if (remove_now) {
        stats_axpy<DP, COV>(p, k_old, sm.x, -1);
        if (tid == 0) __stcg(p.counts + k_old, n_old - 1);
    }
    if (k_new == K) {  // open a new component
        if (K >= p.K_max) {
            if (tid == 0) sh.error = -3;
            __syncthreads();
            return;
        }
        __syncthreads();
        stats_init_prior<DP, COV>(p, K);
        if (tid == 0) {
            const int nf = sh.n_free - 1;
            const int uid = __ldcg(p.uid_free + nf);
            sh.n_free = nf;
            __stcg(p.uid_of_slot + K, uid);
            __stcg(p.slot_of_uid + uid, K);
            __stcg(p.counts + K, 0LL);
            sh.K = K + 1;
            sh.births += 1;
        }
    }
    __syncthreads();
    stats_axpy<DP, COV>(p, k_new, sm.x, +1);
    long long n_new = 0;
    if (tid == 0) {
        n_new = __ldcg(p.counts + k_new) + 1;
        __stcg(p.counts + k_new, n_new);
        __stcg(p.z_uid + i, __ldcg(p.uid_of_slot + k_new));
        sh.moves += 1;
        add_dirty(sh, k_new);
        if (remove_now) add_dirty(sh, k_old);
        sh.n_old = n_new;  // broadcast slot
    }
    __syncthreads();
    n_new = sh.n_old;
    if (warp == 0 && remove_now) {
        const bool okf = refactor_warp<DP, COV>(p, p.num + (size_t)k_old * DP, p.S + (size_t)k_old * stat_len(DP, COV),
                                                n_old - 1, 0, p.rec + (size_t)k_old * R,
                                                k_old < p.Kc ? sm.rec + (size_t)k_old * R : nullptr, sm.A0);
        if (!okf && lane == 0) sh.error = -4;
    }
    if (warp == 1) {
        const bool okf = refactor_warp<DP, COV>(p, p.num + (size_t)k_new * DP, p.S + (size_t)k_new * stat_len(DP, COV),
                                                n_new, 0, p.rec + (size_t)k_new * R,
                                                k_new < p.Kc ? sm.rec + (size_t)k_new * R : nullptr, sm.A1);
        if (!okf && lane == 0) sh.error = -4;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// record staging: global -> shared.  Full reload through the TMA bulk-copy engine (cp.async.bulk +
// mbarrier complete_tx), dirty records with plain ld.cg.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int DP, int COV> __device__ void stage_all_records(const Params &p, const Smem &sm, int K) {
    constexpr int R = rec_len(DP, COV);
    Sh &sh = *sm.sh;
    const int kc = min(K, p.Kc);
    const uint32_t bytes = (uint32_t)((size_t)kc * R * sizeof(double));
    __syncthreads();
    if (bytes == 0) return;
    const uint32_t mb = smem_u32(&sh.mbar);
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        uint32_t done = 0;
        while (done < bytes) {
            const uint32_t chunk = min(bytes - done, 32768u);
            asm volatile(
                "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    smem_u32(sm.rec) + done),
                "l"((const char *)p.rec + done), "r"(chunk), "r"(mb)
                : "memory");
            done += chunk;
        }
    }
    const uint32_t phase = sh.mbar_phase;
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(mb), "r"(phase)
            : "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) sh.mbar_phase = phase ^ 1u;
    __syncthreads();
}

template <int DP, int COV> __device__ void stage_dirty_records(const Params &p, const Smem &sm, int n_dirty,
                                                               const int *dirty) {
    constexpr int R = rec_len(DP, COV);
    for (int t = 0; t < n_dirty; ++t) {
        const int k = dirty[t];
        if (k < 0 || k >= p.Kc) continue;
        for (int e = threadIdx.x; e < R; e += blockDim.x) sm.rec[(size_t)k * R + e] = __ldcg(p.rec + (size_t)k * R + e);
    }
}

// ---------------------------------------------------------------------------------------------
// speculative evaluation of the window [pos, pos+win): thread per datum.
// ---------------------------------------------------------------------------------------------
template <int DP, int COV>
__device__ void window_eval(const Params &p, const Smem &sm, long long pos, long long win, int K) {
    constexpr int R = rec_len(DP, COV);
    const int T = blockDim.x, G = gridDim.x;
    double *wb = p.wbuf + (size_t)blockIdx.x * (size_t)(p.K_max + 1) * T + threadIdx.x;
    long long my_first = POS_INF;
    double my_margin = 1.0;
    for (long long off = (long long)threadIdx.x * G + blockIdx.x; off < win; off += (long long)G * T) {
        const long long j = pos + off;
        if (__ldcg(&p.ctl->first) < j) break;  // an earlier mover is already known: this datum will be redone
        const long long i = p.order ? p.order[j] : j;
        const int uid = __ldcg(p.z_uid + i);
        bool flag = (uid < 0);
        int k_old = -1;
        if (!flag) {
            k_old = __ldcg(p.slot_of_uid + uid);
            if (__ldcg(p.counts + k_old) <= 1) flag = true;  // the component would die: general path
        }
        if (!flag) {
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
            const double wK = p.log_alpha + p.log_prior[i];
            double M = wK;
            bool ok = true;
            for (int k = 0; k < K; ++k) {
                double w;
                if (k < p.Kc) {
                    const double *rec = sm.rec + (size_t)k * R;
                    w = (k == k_old) ? weight_own_removed<DP, COV, false>(rec, x, p, &ok)
                                     : weight_other<DP, COV, false>(rec, x, p.D);
                } else {
                    const double *rec = p.rec + (size_t)k * R;
                    w = (k == k_old) ? weight_own_removed<DP, COV, true>(rec, x, p, &ok)
                                     : weight_other<DP, COV, true>(rec, x, p.D);
                }
                wb[(size_t)k * T] = w;
                M = fmax(M, w);
            }
            if (!ok) {
                flag = true;
            } else {
                double s = 0.0;
                for (int k = 0; k < K; ++k) {
                    const double dlt = wb[(size_t)k * T] - M;
                    const double e = (dlt < EXP_CUTOFF) ? 0.0 : exp(dlt);
                    wb[(size_t)k * T] = e;
                    s += e;
                }
                const double eK = exp(wK - M);
                s += eK;
                double t = p.u[j] * s;
                int k_new = K;
                double marg = 0.0;
                bool hit = false;
                for (int k = 0; k < K; ++k) {
                    const double tb = t;
                    t -= wb[(size_t)k * T];
                    if (t < 0.0) { k_new = k; marg = fmin(fabs(tb), -t) / s; hit = true; break; }
                }
                if (!hit) { const double tb = t; t -= eK; marg = (t < 0.0) ? fmin(fabs(tb), -t) / s : 0.0; }
                if (k_new != k_old || !(s > 0.0) || !(s < INFINITY) || marg < p.guard) flag = true;
                else my_margin = fmin(my_margin, marg);
            }
        }
        if (flag) { my_first = j; break; }
    }
    // one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_first = min(my_first, __shfl_xor_sync(0xffffffffu, my_first, o));
        my_margin = fmin(my_margin, __shfl_xor_sync(0xffffffffu, my_margin, o));
    }
    if ((threadIdx.x & 31) == 0) {
        if (my_first != POS_INF) atomicMin(&p.ctl->first, my_first);
        if (my_margin < 1.0)
            atomicMin(&p.ctl->margin_bits, (unsigned long long)__double_as_longlong(my_margin));
    }
}

// ---------------------------------------------------------------------------------------------
// the sweep kernel
// ---------------------------------------------------------------------------------------------
template <int DP, int COV>
__global__ void __launch_bounds__(T_SWEEP, 1) k_sweep(const Params p) {
    extern __shared__ __align__(16) double smem_raw[];
    constexpr int R = rec_len(DP, COV);
    const Smem sm = carve<DP, COV>(smem_raw, p);
    Sh &sh = *sm.sh;
    Ctl *ctl = p.ctl;
    const int tid = threadIdx.x;
    const bool cta0 = (blockIdx.x == 0);

    if (tid == 0) {
        sh.mbar_phase = 0;
        sh.error = 0;
        sh.n_dirty = 0;
        sh.reload_all = 0;
        const uint32_t mb = smem_u32(&sh.mbar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // sweep prologue (CTA 0): the count-prior term of every record follows this sweep's power
    // (crpmm.py:70 / pcrpmm.py:105-112), and the scan restarts.
    if (cta0) {
        const int K0 = __ldcg(&ctl->K);
        constexpr int SO = rec_sc_off(DP, COV);
        for (int k = tid; k < K0; k += blockDim.x) {
            const double n = __ldcg(p.rec + (size_t)k * R + SO + SC_N);
            __stcg(p.rec + (size_t)k * R + SO + SC_LC, log_count(n, p.power));
        }
        if (tid == 0) {
            __stcg(&ctl->pos, p.start_pos);
            __stcg(&ctl->first, POS_INF);
            __stcg(&ctl->n_dirty, 0);
            __stcg(&ctl->gap, p.init_gap);
            long long w = 0;
            if (p.engine == 2 || (p.engine == 0 && p.init_gap >= SEQ_GAP)) {
                const long long wmax = 4LL * gridDim.x * blockDim.x;
                w = (long long)fmin(fmax(2.0 * p.init_gap, (double)WIN_MIN), (double)wmax);
                if (w > p.N - p.start_pos) w = p.N - p.start_pos;
            }
            __stcg(&ctl->win, w);
        }
    }
    grid_barrier(ctl);
    int K = __ldcg(&ctl->K);
    stage_all_records<DP, COV>(p, sm, K);
    unsigned int my_full_gen = __ldcg(&ctl->full_gen);
    if (tid == 0) sh.K = K;
    __syncthreads();

    while (true) {
        const long long pos = __ldcg(&ctl->pos);
        const long long win = __ldcg(&ctl->win);
        if (pos >= p.N || __ldcg(&ctl->error) != 0) break;
        if (win > 0) window_eval<DP, COV>(p, sm, pos, win, K);
        grid_barrier(ctl);
        if (cta0) {
            if (tid == 0) {
                sh.n_dirty = 0; sh.reload_all = 0; sh.K = K;
                sh.moves = __ldcg(&ctl->moves); sh.births = __ldcg(&ctl->births); sh.deaths = __ldcg(&ctl->deaths);
                sh.evals = __ldcg(&ctl->evals); sh.windows = __ldcg(&ctl->windows);
                sh.seq_data = __ldcg(&ctl->seq_data); sh.wasted = __ldcg(&ctl->wasted);
                sh.guard_hits = __ldcg(&ctl->guard_hits);
                sh.uextra = __ldcg(&ctl->uextra);
                sh.margin_bits = __ldcg(&ctl->margin_bits); sh.gap = __ldcg(&ctl->gap);
                sh.n_free = __ldcg(&ctl->n_free);
                sh.n_old = __ldcg(&ctl->first);  // broadcast slot
            }
            __syncthreads();
            long long new_pos;
            double gap = sh.gap;
            if (win > 0) {
                const long long f = sh.n_old;
                const long long end = pos + win;
                __syncthreads();
                if (f < end) {
                    if (tid == 0) { sh.evals += (f - pos) * (long long)K; sh.wasted += end - (f + 1); }
                    resolve<DP, COV>(p, sm, f);
                    new_pos = f + 1;
                    gap = 0.7 * gap + 0.3 * (double)(f - pos + 1);
                } else {
                    if (tid == 0) sh.evals += win * (long long)K;
                    new_pos = end;
                    gap = fmax(gap, 0.7 * gap + 0.3 * 2.0 * (double)win);
                }
                if (tid == 0) sh.windows += 1;
            } else {
                const long long run = min((long long)SEQ_RUN, p.N - pos);
                const long long moves0 = sh.moves;
                __syncthreads();
                for (long long s = 0; s < run; ++s) {
                    resolve<DP, COV>(p, sm, pos + s);
                    if (sh.error) break;
                }
                new_pos = pos + run;
                const long long mv = sh.moves - moves0;
                gap = 0.5 * gap + 0.5 * (double)run / (double)(mv + 1);
                __syncthreads();
                if (tid == 0) { sh.seq_data += run; sh.reload_all = 1; }
            }
            __syncthreads();
            if (tid == 0) {
                long long w = 0;
                if (p.engine == 2 || (p.engine == 0 && gap >= SEQ_GAP)) {
                    const long long wmax = 4LL * gridDim.x * blockDim.x;
                    w = (long long)fmin(fmax(2.0 * gap, (double)WIN_MIN), (double)wmax);
                }
                if (w > p.N - new_pos) w = p.N - new_pos;
                if (w < 0) w = 0;
                __stcg(&ctl->gap, gap);
                __stcg(&ctl->win, w);
                __stcg(&ctl->pos, new_pos);
                __stcg(&ctl->first, POS_INF);
                __stcg(&ctl->K, sh.K);
                __stcg(&ctl->moves, sh.moves); __stcg(&ctl->births, sh.births); __stcg(&ctl->deaths, sh.deaths);
                __stcg(&ctl->evals, sh.evals); __stcg(&ctl->windows, sh.windows);
                __stcg(&ctl->seq_data, sh.seq_data); __stcg(&ctl->wasted, sh.wasted);
                __stcg(&ctl->guard_hits, sh.guard_hits);
                __stcg(&ctl->uextra, sh.uextra);
                __stcg(&ctl->margin_bits, sh.margin_bits);
                __stcg(&ctl->n_free, sh.n_free);
                if (sh.error) __stcg(&ctl->error, sh.error);
                if (sh.reload_all) { __stcg(&ctl->full_gen, __ldcg(&ctl->full_gen) + 1u); __stcg(&ctl->n_dirty, 0); }
                else {
                    __stcg(&ctl->n_dirty, sh.n_dirty);
                    for (int t = 0; t < sh.n_dirty; ++t) __stcg(&ctl->dirty[t], sh.dirty[t]);
                }
            }
        }
        grid_barrier(ctl);
        K = __ldcg(&ctl->K);
        const unsigned int fg = __ldcg(&ctl->full_gen);
        if (fg != my_full_gen) {
            my_full_gen = fg;
            if (!cta0) stage_all_records<DP, COV>(p, sm, K);
        } else if (!cta0) {
            const int nd = __ldcg(&ctl->n_dirty);
            int dl[MAX_DIRTY];
            for (int t = 0; t < MAX_DIRTY; ++t) dl[t] = (t < nd) ? __ldcg(&ctl->dirty[t]) : -1;
            stage_dirty_records<DP, COV>(p, sm, nd, dl);
        }
        __syncthreads();
    }
}

}  // namespace bgmm
