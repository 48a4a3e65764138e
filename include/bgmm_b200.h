/*
 * bgmm_b200.h -- C-ABI of libbgmm_b200.so, the B200 (sm_100a) collapsed-Gibbs engine
 * for the CRP / powered-CRP Gaussian mixture model.
 *
 * The reference (junlulocky/PyBGMM) has no FFI: its seam is a duck-typed Python
 * protocol.  Every entry point below replaces one piece of that protocol; the
 * citation after "replaces:" is the reference file:line (relative to the reference
 * repository root).  INTEGRATION.md shows the ctypes stub a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; host buffers are caller owned, device memory is
 *     library owned (except the *_dev entry points, which take caller device pointers);
 *   - every function returns 0 on success or a negative BGMM_E* code; the message of the
 *     last failure on the calling thread is bgmm_last_error();
 *   - one host thread per handle (the reference is single threaded, igmm.py / crpmm.py);
 *   - all floating point is IEEE fp64, labels are int64, X is C-order (N, D).
 *   - there is NO CPU fallback: without a CUDA device every call that needs one fails
 *     with BGMM_ENODEV.
 */
#ifndef BGMM_B200_H
#define BGMM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGMM_OK 0
#define BGMM_EINVAL (-1)   /* bad argument                                   (reference: assert / ValueError) */
#define BGMM_ENODEV (-2)   /* no CUDA device / CUDA runtime failure                                            */
#define BGMM_EKMAX (-3)    /* a birth would exceed K_max                     (reference: IndexError, gaussian_components.py:161) */
#define BGMM_ENUMERIC (-4) /* covariance not positive definite / NaN         (reference: LinAlgError or NaN weights) */
#define BGMM_ENOMEM (-5)
#define BGMM_EWATCHDOG (-6) /* a device spin loop timed out (replicas of the sweep kernel stopped agreeing)          */

#define BGMM_COV_FULL 0 /* NIW, full covariance      -- pybgmm/gaussian/gaussian_components.py      */
#define BGMM_COV_DIAG 1 /* product of NIX, diagonal  -- pybgmm/gaussian/gaussian_components_diag.py */
#define BGMM_COV_FIXED 2 /* known diagonal variance  -- pybgmm/gaussian/gaussian_components_fixedvar.py (bgmm_create_fixedvar) */

typedef struct bgmm_handle bgmm_t;

/* Per-sweep counters (no counterpart in the reference; diagnostics for parity and the metric). */
typedef struct {
    int64_t K;          /* live components after the sweep(s)                                        */
    int64_t moves;      /* data that went through add_item (crpmm.py:88) rather than the restore path */
    int64_t births;     /* new components opened (gaussian_components.py:161-164)                     */
    int64_t deaths;     /* components deleted (gaussian_components.py:188-205)                        */
    int64_t evals;      /* datum-component log_post_pred evaluations = sum over data of live K       */
    int64_t windows;    /* speculative windows evaluated (engine diagnostic)                          */
    int64_t seq_data;   /* data resolved by the sequential path (engine diagnostic)                   */
    int64_t wasted;     /* speculative datum evaluations discarded and redone (engine diagnostic)     */
    double min_margin;  /* min distance of a uniform to the CDF boundary it was compared with         */
    double device_ms;   /* CUDA-event time of the sweep kernel(s), summed                             */
    int64_t explicit_evals; /* own-component weights rebuilt exactly from the statistics (engine diagnostic)  */
    int64_t refreshes;      /* records rebuilt from the statistics for drift control, incl. cluster launches ended for it  */
    int64_t generic_from;   /* scan position the generic engine took over at, or -1 (engine diagnostic)       */
    int64_t phase_cycles[16]; /* SM cycles CTA 0 spent per engine phase in the last sweep (engine diagnostic)   */
    int64_t launches;         /* kernels launched by this call                                                  */
    double sweep_kernel_ms;   /* CUDA-event time of the sweep kernel alone (device_ms also covers record set-up) */
    int64_t guard_hits;       /* draws whose margin was below the guard and were redone on the exact path (bgmm_set_guard) */
    int64_t fast_steps;       /* data resolved by a resident sequential step: the cluster step engines (and the data their
                                 window kernel certified as stays) or the in-CTA register step (engine diagnostic)      */
} bgmm_sweep_stats;

const char *bgmm_version(void);
const char *bgmm_last_error(void);
/* number of visible CUDA devices (0 when there is none; never fails) */
int bgmm_device_count(void);

/*
 * replaces: GaussianComponents.__init__ + _cache  (gaussian_components.py:75-127) and
 *           GaussianComponentsDiag.__init__ + _cache (gaussian_components_diag.py:84-135), prior = NIW (prior/niw.py:8-23).
 * S0 is D*D (full) or D (diag).  v0 must be integer valued (it indexes the lgamma table,
 * gaussian_components.py:238).  lgamma_half_tab[n] / log_tab[n] are the tables of
 * gaussian_components.py:120-122 (index 0 is the dud entry), length tab_len >= v0 + N + 2;
 * pass NULL to have the library build them with libm.  K_max is the component capacity.
 * Allocates device state, uploads X, computes cached_log_prior[i] for all i; all data start unassigned.
 */
int bgmm_create(const double *X, int64_t N, int32_t D, int32_t cov_type, const double *m0, double k0, int64_t v0,
                const double *S0, int32_t K_max, const double *lgamma_half_tab, const double *log_tab,
                int64_t tab_len, int32_t device, bgmm_t **out);
/*
 * replaces: GaussianComponentsFixedVar.__init__ + _cache (gaussian_components_fixedvar.py:75-130), prior =
 * FixedVarPrior(var, mu_0, var_0) (:304-311): components with a known diagonal variance `var` (D) and independent normal
 * priors N(mu_0, var_0) (D each) on their means; IGMM's covariance_type="fixed" (igmm/igmm.py:108-109).  The handle
 * behaves like any other: in bgmm_get_state / bgmm_set_component_stats, m_num is mu_N_numerators (K_max x D), S_part is
 * precision_Ns (K_max x D), logdet is log_prod_precision_preds, inv_covar is precision_preds (K_max x D).
 */
int bgmm_create_fixedvar(const double *X, int64_t N, int32_t D, const double *var, const double *mu_0, const double *var_0,
                         int32_t K_max, int32_t device, bgmm_t **out);
int bgmm_destroy(bgmm_t *h);

/* All kernels of this handle are enqueued on `cuda_stream` (a cudaStream_t; NULL = the legacy default stream). */
int bgmm_set_stream(bgmm_t *h, void *cuda_stream);

/*
 * replaces: the component build loop of gaussian_components.py:95-111 -- for k ascending, for i ascending
 * with z[i]==k: add_item(i,k), so the sufficient statistics are accumulated in the reference's order
 * (bit-identical sums).  z[i] in {-1, 0..max}, labels must be consecutive (the reference asserts, :103-105).
 */
int bgmm_set_assignments(bgmm_t *h, const int64_t *z);

/*
 * replaces: one pass of the per-datum loop of CRPMM.collapsed_gibbs_sampler (igmm/crpmm.py:57-88) or
 * PCRPMM.collapsed_gibbs_sampler (igmm/pcrpmm.py:93-131), including utils.draw (utils/utils.py:7-20).
 *   order    : scan order, N int64 (pcrpmm.py:86-91); NULL = 0..N-1 (crpmm.py:57)
 *   uniforms : uniforms[j] is the random.random() consumed for the j-th visited datum (utils.py:15);
 *              NULL = counter-based Philox4x32-10 stream keyed by (seed, sweep index), see bgmm_get_uniforms
 *   alpha    : CRP concentration (crpmm.py:74)
 *   power    : count prior is log(pow(n_k, power)) (pcrpmm.py:107-108); 1.0 = plain CRP log(n_k) (crpmm.py:70)
 * Host buffers; the host->device copies are part of the call.  `out` may be NULL.
 */
int bgmm_sweep(bgmm_t *h, const int64_t *order, const double *uniforms, double alpha, double power,
               bgmm_sweep_stats *out);
/* Same, with `d_order` / `d_uniforms` already resident in device memory (either may be NULL as above). */
int bgmm_sweep_dev(bgmm_t *h, const int64_t *d_order, const double *d_uniforms, double alpha, double power,
                   bgmm_sweep_stats *out);
/*
 * Many independent chains on one GPU (no counterpart in the reference, which runs one chain in one Python process;
 * BASELINE.json configs[3] "8 independent chains" is the multi-GPU form of the same fan-out).
 *   bgmm_fork        a new chain on the SAME data and prior as `parent`: shares the device copy of X, the cached log
 *                    prior and the tables; own labels, statistics and RNG state.  All data start unassigned.  The shared
 *                    buffers live until the last chain of the family is destroyed.
 *   bgmm_sweep_many  one sweep (crpmm.py:57-88 / pcrpmm.py:93-131) of each of the n chains with ONE kernel launch, one
 *                    thread block per chain; d_orders / d_uniforms are arrays of n DEVICE pointers (either array, or any
 *                    entry, may be NULL with the meaning of bgmm_sweep).  Every chain walks exactly the chain bgmm_sweep
 *                    would walk for the same inputs.  The chains must share device, stream, D, covariance type and K_max
 *                    (full covariance, D <= 16).  out: n entries, or NULL.
 */
int bgmm_fork(bgmm_t *parent, bgmm_t **out);
int bgmm_sweep_many(bgmm_t *const *handles, int32_t n, const int64_t *const *d_orders, const double *const *d_uniforms,
                    double alpha, double power, bgmm_sweep_stats *out);
/*
 * replaces: one pass of the per-datum loop of CSCRPMM.constrained_gibbs_sample with the constrained re-draw
 * (igmm/cscrpmm.py:290-358; the approximate step :418-461 is the same loop with CRP weights).  status[k] says what slot k
 * was when the sweep started: 1 = useful cluster, 2 = non-useful cluster (cscrpmm.py:159-167), 0 = neither.  After the
 * ordinary draw, a datum whose old slot is non-useful draws again -- same probabilities, the next uniform of the stream
 * (utils.draw, utils/utils.py:7-20) -- until the slot drawn is a useful one (:344-347); any other datum goes back to its
 * old slot index (:349).  `uniforms` is the random.random() stream from the sweep's first draw on, n_uniforms >= N values;
 * *consumed returns how many were used (so the caller can leave the interpreter's generator exactly where the
 * reference's loop would have).  Runs datum by datum on the generic engine (the number of uniforms a datum consumes
 * depends on the draws before it).  BGMM_EINVAL if the stream runs out.
 */
int bgmm_sweep_constrained(bgmm_t *h, const int64_t *order, const double *uniforms, int64_t n_uniforms, double alpha,
                           double power, const int32_t *status, int32_t n_status, int64_t *consumed,
                           bgmm_sweep_stats *out);
/* Engine policy: 0 = adaptive (default), 1 = always the sequential per-datum path, 2 = always speculative windows;
 * 3..5 = the same three policies on the generic engine (any D, any K_max) instead of the shared-memory-resident one;
 * 6 = adaptive, with every sweep started on the cluster step engine (full covariance, D <= 16) instead of only the
 * sweeps that follow a dense-mover sweep. */
int bgmm_set_engine(bgmm_t *h, int32_t mode);

/* Philox stream used when uniforms == NULL: seed it, and replay the stream of a given sweep index on the host
 * (so a CPU oracle can consume identical uniforms).  The sweep counter starts at 0 and increments per sweep. */
int bgmm_seed(bgmm_t *h, uint64_t seed);
int bgmm_get_uniforms(bgmm_t *h, int64_t sweep_index, double *out /* N */);
int64_t bgmm_sweep_index(bgmm_t *h);

/*
 * replaces: the attribute protocol of GaussianComponents (gaussian_components.py:78-98):
 * assignments[N], counts[K_max], K, m_N_numerators[K_max*D], S_N_partials[K_max*D*D | K_max*D],
 * logdet_covars[K_max] (diag: log_prod_vars), inv_covars[K_max*D*D] (diag: inv_vars[K_max*D]).
 * Slots >= K are zero, as in the reference (gaussian_components.py:200-204).  Any pointer may be NULL.
 */
int bgmm_get_state(bgmm_t *h, int64_t *z, int64_t *counts, int32_t *K, double *m_num, double *S_part, double *logdet,
                   double *inv_covar);
/* Relabelled assignments written to a caller DEVICE buffer of N int64 (for NCCL gathers; no host copy). */
int bgmm_get_assignments_dev(bgmm_t *h, int64_t *d_out);
int bgmm_K(bgmm_t *h);

/* replaces: cached_log_prior (gaussian_components.py:125-127) = log_prior(i) for all i (:207-214; diag :215-222) */
int bgmm_log_prior(bgmm_t *h, double *out /* N */);
/* replaces: log_post_pred(i) (gaussian_components.py:228-251; diag :237-259) for n data; out is n x K row-major */
int bgmm_log_post_pred(bgmm_t *h, const int64_t *idx, int64_t n, double *out);
/* replaces: log_marg_k(k) for all k < K (gaussian_components.py:253-276; diag :271-289); out has K entries */
int bgmm_log_marg_k(bgmm_t *h, double *out);
/* replaces: IGMM.log_marg (igmm/igmm.py:199-215) with the CRP term evaluated with libm lgamma */
int bgmm_log_marg(bgmm_t *h, double alpha, double *out);

/*
 * replaces: the clustering metrics of GMM.update_record_dict (gmm/gmm.py:81-106) -- normalized_mutual_information,
 * mutual_information, information_variation (infopy/infopy.py:31-119: one pass over the data per (true, predicted)
 * cell) and utils.cluster_loss_inertia (utils/utils.py:31-88) -- by ONE device pass over the labels:
 *   bgmm_set_true_labels  uploads the ground-truth labels once (values in [0, T));
 *   bgmm_contingency      table[t * (K + 1) + k] = #{i: true[i] == t, assignments[i] == k}; column K counts the
 *                         unassigned (-1) data.  T x (K + 1) int64, row-major.  Every metric above is a function of it.
 *   bgmm_cluster_ssq      out[k] = sum over members of component k of |x_i - mean_k|^2 (utils.py:52-88 takes the
 *                         square root and truncates to an integer), from the sufficient statistics; K entries.
 * The labels never leave the device.
 */
int bgmm_set_true_labels(bgmm_t *h, const int64_t *t /* N */, int32_t T);
int bgmm_contingency(bgmm_t *h, int64_t *table /* T * (K + 1) */);
int bgmm_cluster_ssq(bgmm_t *h, double *out /* K */);

/* replaces: add_item(i,k) (gaussian_components.py:154-169) and del_item(i) (:171-186, incl. del_component :188-205) */
int bgmm_add_item(bgmm_t *h, int64_t i, int32_t k);
int bgmm_del_item(bgmm_t *h, int64_t i);
/* replaces: restore_component_from_stats(k, ...) (gaussian_components.py:144-152): overwrite the sufficient statistics
 * of live component k (m_num: D, S_part: D*D full | D diag, count) -- logdet/inv are re-derived from them. */
int bgmm_set_component_stats(bgmm_t *h, int32_t k, const double *m_num, const double *S_part, int64_t count);

/* replaces: the caller-side `components.assignments[i] = k_old` that completes the reference's restore path
 * (crpmm.py:84-85): rewrite the label of datum i (k = -1: unassigned) without touching any statistic. */
int bgmm_set_label(bgmm_t *h, int64_t i, int32_t k);
/* Resume from a saved state (SURVEY.md 8b bgmm_set_state; the reference's in-memory snapshots are copy.deepcopy of
 * the components, cscrpmm.py:173): bgmm_set_assignments(z), then -- when m_num / S_part (K x D, K x D*D | K x D) are
 * given -- the exact bits of the saved statistics per component (bgmm_set_component_stats). */
int bgmm_set_state(bgmm_t *h, const int64_t *z, int32_t K, const double *m_num, const double *S_part);
/* Margin guard of the draw (no counterpart in the reference): a draw whose uniform is closer than `guard`
 * (probability units) to a boundary of the drawn interval is redone on the exact path (records rebuilt from the
 * statistics, libm log / exp, sequential-subtract draw of utils.py:15-20).  Default 1e-9 (env BGMM_GUARD at create);
 * 0 disables.  bgmm_sweep_stats.guard_hits counts the redone draws. */
int bgmm_set_guard(bgmm_t *h, double guard);

/*
 * replaces: the stream of random.random() calls made by utils.draw (utils/utils.py:15).  `state` is the 625-word
 * tuple of random.getstate()[1] (624 words of MT19937 state + position); it is advanced in place, so
 * random.setstate() with it leaves the interpreter exactly where the reference's n calls would have.
 * Pure host code (no device needed).
 */
int bgmm_mt19937_fill(uint32_t *state /* 625 */, double *out, int64_t n);

#ifdef __cplusplus
}
#endif
#endif /* BGMM_B200_H */
