/* pfb200.h — C ABI of libpfb200.so, the B200-native ELBO-and-resample engine.
 *
 * The reference (mlcolab/Pathfinder.jl v0.10.7) has no FFI layer: its hot path is entered
 * through ordinary Julia functions.  Each entry point below names the reference call it
 * replaces (paths relative to the reference repo).  INTEGRATION.md shows the Julia `ccall`
 * shim and the Python ctypes binding (pathfinder_b200/_lib.py) that a maintainer would add.
 *
 * Conventions
 *   - all reals are IEEE double, all matrices column-major (Julia layout), indices 1-based
 *     where the reference's are (best_iter, inds, ids);
 *   - the caller owns every host buffer; the engine never keeps a host pointer after return;
 *   - return 0 = OK, < 0 = argument/shape error (ArgumentError / DimensionMismatch),
 *     > 0 = CUDA runtime error code; pfb_last_error() gives the text;
 *   - numerical failure is data, not an error: a non-positive-definite iteration gets
 *     ELBO = NaN (the reference throws PosDefException from cholesky, src/woodbury.jl:205),
 *     non-finite log densities propagate as in src/elbo.jl:16-17, success[] mirrors
 *     src/singlepath.jl:309-314.
 */
#ifndef PFB200_H
#define PFB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PFB_OK 0
#define PFB_ERR_ARG (-1)
#define PFB_ERR_SHAPE (-2)
#define PFB_ERR_UNSUPPORTED (-3)
#define PFB_ERR_STATE (-4)
#define PFB_ERR_NUMERIC (-5) /* importance resampling with no usable weight at all */

/* registered device-side target log densities */
#define PFB_MODEL_ISONORMAL 0 /* logp(x) = -|x|^2/2          test/singlepath.jl:15          */
#define PFB_MODEL_FUNNEL 1    /* Neal's funnel               docs/src/examples/quickstart.md:229-234 */
#define PFB_MODEL_DIAGNORMAL 2 /* independent normals; blob = { mean[n], sd[n] }  test/elbo.jl:8-12 */
#define PFB_MODEL_DENSENORMAL 3 /* logp = -(x-m)'P(x-m)/2; blob = { m[n], P[n x n] column-major }
                                   docs/src/examples/quickstart.md:17-24 (BASELINE config 5)         */
#define PFB_MODEL_HLOGISTIC 4  /* hierarchical logistic regression, theta = (log tau, b0, b[n-2]):
                                   log tau ~ N(0,1), b0 ~ N(0,2.5^2), b_j ~ N(0,tau^2),
                                   y_i ~ Bernoulli(sigmoid(b0 + x_i'b)); blob = { nobs,
                                   X[nobs x (n-2)] column-major, y[nobs] }   (BASELINE config 4)     */
#define PFB_MODEL_HOSTCALLBACK 5 /* arbitrary target evaluated by a HOST callback (pfb_register_host_model) */
#define PFB_MODEL_EXTERNAL 100 /* internal: log p evaluated outside the sampling kernel             */

typedef struct pfb_engine* pfb_handle;

/* Keyword defaults of the reference (src/Pathfinder.jl:24-27, src/inverse_hessian.jl:25). */
typedef struct {
    int32_t device;          /* CUDA device ordinal                                          */
    int32_t history_length;  /* J, DEFAULT_HISTORY_LENGTH = 6                                 */
    int32_t ndraws_elbo;     /* K, DEFAULT_NDRAWS_ELBO = 5                                    */
    int32_t materialize_all; /* 1: keep the draws of every iteration on the device (mode M,
                                the ELBOEstimate.draws payload of src/elbo.jl:19); 0: lean    */
    int32_t elbo_mode;       /* 0: auto — the lean ELBO stage generates each normal once and gets
                                log p from quadratic-form statistics (all registered families are
                                diagonal-quadratic); 1: always the generic two-pass kernel     */
    int32_t reserved;        /* 0                                                              */
    double eps;              /* curvature tolerance, 1e-12                                    */
} pfb_config;

/* Outputs of pfb_elbo_batch / pfb_batch_download.  Any pointer may be NULL (skipped).
 * U = sum_p L_p units; unit order = path-major, iteration-minor (elbo_estimates[l] of path p). */
typedef struct {
    double* elbo;        /* [U]        ELBOEstimate.value                      src/elbo.jl:17 */
    double* elbo_se;     /* [U]        ELBOEstimate.std_err                    src/elbo.jl:18 */
    double* logp;        /* [K x U]    log_densities_target per iteration      src/elbo.jl:15 */
    double* logq;        /* [K x U]    log_densities_fit                        src/mvnormal.jl:36 */
    int64_t* best_iter;  /* [P]        fit_iteration, 1-based, 0 = none         src/elbo.jl:8  */
    int32_t* success;    /* [P]                                       src/singlepath.jl:309-314 */
    int64_t* n_rejected; /* [P]        num_bfgs_updates_rejected      src/inverse_hessian.jl:57 */
    double* draws;       /* [n x K x P] draws of the best iteration  src/singlepath.jl:231-232 */
    double* draws_logp;  /* [K x P]                                                            */
    double* draws_logq;  /* [K x P]                                                            */
    /* best-iteration fit_distribution in the reference's WoodburyPDMat form (f4):            */
    double* fit_mu;      /* [n x P]    MvNormal.mu                              src/mvnormal.jl:17 */
    double* fit_alpha;   /* [n x P]    diag(A)                                                 */
    double* fit_vh;      /* [n x KP x P] Householder reflectors of F.Q (unit diagonal explicit) */
    double* fit_T;       /* [KP x KP x P] compact-WY T of F.Q, row-major: the FULL k x k factor (LAPACK's
                            geqrt, which Julia's qr calls with nb = min(k, 36), stores its diagonal blocks) */
    double* fit_Vc;      /* [KP x KP x P] F.V (upper Cholesky factor), row-major               */
    double* fit_logdet;  /* [P]        logdet(Sigma)                     src/woodbury.jl:77-80 */
    int32_t* fit_jeff;   /* [P]        history_length_effective                                */
    double* all_draws;   /* [n x K x U] every iteration's draws (only if materialize_all)       */
} pfb_elbo_out;

/* Outputs of the PSIS + resample stage. Any pointer may be NULL. */
typedef struct {
    double* log_weights; /* [N] normalised smoothed log weights     PSISResult.log_weights     */
    double* weights;     /* [N] exp(log_weights)                    PSISResult.weights         */
    double* pareto_k;    /* [1] PSISResult.pareto_shape                                        */
    int64_t* tail_len;   /* [1] PSISResult.tail_length                                         */
    int64_t* inds;       /* [ndraws] sample_inds, 1-based           src/resample.jl:60-67      */
    int64_t* ids;        /* [ndraws] draw_component_ids             src/resample.jl:70         */
    double* draws;       /* [n x ndraws]                            src/resample.jl:68         */
} pfb_resample_out;

/* lifetime ---------------------------------------------------------------------------------- */
int pfb_create(pfb_handle* out, const pfb_config* cfg);
int pfb_destroy(pfb_handle h);
const char* pfb_last_error(pfb_handle h);
int pfb_kp(pfb_handle h); /* padded reflector count KP: 12, 20, 24 for history_length <= 6, 10, 12 (tensor-core
                             kernels); 2 * history_length for 13..64 (generic runtime-width kernels K2g / K3g) */

/* Target density: replaces the Julia closure logp(x) (src/singlepath.jl:186, src/multipath.jl:159)
 * by a registered device-side family + parameter blob (doubles). */
int pfb_register_model(pfb_handle h, int family, int n, const double* blob, size_t ndoubles);

/* Row f2 — arbitrary target density: the Julia closure `logp` itself (src/singlepath.jl:186,
 * src/multipath.jl:159; any LogDensityProblems / Turing model) evaluated on the HOST.  The ELBO
 * stage then materialises the draws of a chunk of (path, iteration) units on the device, streams
 * them to pinned host memory, and calls `cb(user, x, n, m, logp_out)` with x = n x m column-major
 * draws (what `logp.(eachcol(x))` consumes, src/elbo.jl:15; src/resample.jl:90-92) while the next
 * chunk is being sampled and copied; logp_out[m] goes back to the device for the ELBO reduction.
 * The callback runs on the thread that called pfb_batch_run / pfb_draw_from_fits (for the fallback draws
 * of a FAILED path: the first call that consumes the pool afterwards — pfb_batch_download,
 * pfb_psis_resample, ...), never concurrently with itself.  Non-finite values propagate as in src/elbo.jl:16-17. */
typedef void (*pfb_logp_callback)(void* user, const double* x, int64_t n, int64_t m, double* logp_out);
int pfb_register_host_model(pfb_handle h, int n, pfb_logp_callback cb, void* user);

/* One call for the whole ELBO stage of P paths.  Replaces, batched over (path x iteration):
 *   fit_mvnormals(optim_trace.points, optim_trace.gradients; history_length)   src/singlepath.jl:301-303
 *   maximize_elbo(rng, logp, fit_distributions[2:end], ndraws_elbo, ntasks)    src/singlepath.jl:306-308
 *   success / draw selection                                             src/singlepath.jl:309-314, 224-233
 * offsets[P+1]: first column of each path in positions/gradients (path p has L_p + 1 points);
 * positions, gradients: n x offsets[P] column-major (gradients of the LOG density, src/optimize.jl:96);
 * seeds[U]: the UInt64 seeds drawn as in src/elbo.jl:2, unit order;
 * normals_or_null: parity mode, u[n x K x U] supplied by the host RNG (src/mvnormal.jl:30). */
int pfb_elbo_batch(pfb_handle h, int n, int P, const int64_t* offsets, const double* positions,
                   const double* gradients, const uint64_t* seeds, const double* normals_or_null,
                   pfb_elbo_out* out);

/* Failed paths (success = 0: no iteration at all, or a NaN / -Inf best ELBO).  The reference then returns
 *   rand(rng, fit_distributions[fit_iteration + 1], ndraws)            src/singlepath.jl:224-228
 * — fresh draws from the fit of the "best" iteration (the identity fit N(theta_0 + grad_0, I) of iteration
 * 0 when there is none) — and these enter the PSIS pool with their own log densities
 * (src/multipath.jl:217, src/resample.jl:81-95).  The engine does the same with one UInt64 seed per path
 * (drawn from the path's rng by the caller; consumed by the next pfb_batch_run / pfb_elbo_batch; when not
 * set, a fixed per-path default is used): pfb_elbo_out.draws / draws_logp / draws_logq of a failed path
 * are those fresh draws.  A NaN log ratio in the pool gets zero importance weight. */
int pfb_set_fallback_seeds(pfb_handle h, int P, const uint64_t* seeds);

/* The same, split so that callers can keep inputs resident / overlap / time the stages. */
int pfb_batch_upload(pfb_handle h, int n, int P, const int64_t* offsets, const double* positions,
                     const double* gradients, const uint64_t* seeds, const double* normals_or_null);
/* K1..K5, asynchronous on the engine stream: returns without a host round trip (with a host-callback
 * target it returns once the last chunk has been evaluated).  Whether a path failed is resolved by the first
 * consumer of the pool: pfb_psis_resample and pfb_batch_download queue up behind the ELBO stage and repeat
 * their pool part only when a path did fail; pfb_batch_device_view, pfb_pool_*, pfb_draw_from_fits and the
 * pool exchange wait for the flags first. */
int pfb_batch_run(pfb_handle h);
int pfb_batch_sync(pfb_handle h);
int pfb_batch_download(pfb_handle h, pfb_elbo_out* out);

/* Row f1 — batched device L-BFGS for every registered device-side family.  Replaces the per-path trajectory producer
 *   optimize_with_trace(prob, optimizer; maxiters)                src/optimize.jl:35-59
 *   (default_optimizer = Optim.LBFGS(m = history_length, ...),    src/Pathfinder.jl:29-35)
 * mapped over the runs by _chunk_tmap (src/multipath.jl:190-208): one CTA per path runs the whole
 * optimisation (contract: pathfinder_b200/csrc/pf_lbfgs.h) and leaves the trace
 * (OptimizationTrace points / log_densities / gradients, src/optimize.jl:110-114) resident on the
 * device.  x0[n x P]: initial points.  npoints[P] = L_p + 1 trace points; status[P] = PF_LBFGS_*
 * (0 gradient tolerance, 1 objective tolerance, 2 maxiters, 3 line search failed, 4 non-finite:
 * the point is recorded and the run stops, src/optimize.jl:103-105); nevals[P] density evaluations
 * (may be NULL).  PFB_ERR_UNSUPPORTED for a host-callback target (optimise it on the host). */
typedef struct {
    int32_t maxiters;   /* 1000, src/optimize.jl:40                                       */
    int32_t max_points; /* capacity of the per-path trace on the device (<= maxiters + 1) */
    double gtol;        /* max |gradient| tolerance, 1e-8 (Optim g_abstol)                */
    double ftol;        /* relative objective decrease tolerance                          */
} pfb_lbfgs_opts;
int pfb_lbfgs_batch(pfb_handle h, int n, int P, const double* x0, const pfb_lbfgs_opts* opts,
                    int64_t* npoints, int32_t* status, int32_t* nevals);
/* Makes the traces of the last pfb_lbfgs_batch the current batch (device-to-device pack into the
 * n x T layout; what pfb_batch_upload does from host buffers).  seeds[U], U = sum(npoints) - P. */
int pfb_batch_from_lbfgs(pfb_handle h, const uint64_t* seeds);
/* The traces of the last pfb_lbfgs_batch, packed path after path: positions / gradients
 * [n x T] column-major, log_densities[T], T = sum(npoints).  Any pointer may be NULL. */
int pfb_lbfgs_download(pfb_handle h, double* positions, double* gradients, double* log_densities);
int pfb_lbfgs_ms(pfb_handle h, double* ms); /* kernel time of the last pfb_lbfgs_batch */

/* K1 + K2 only, the best iteration of every path given by the caller (1-based, 0 = none): rebuilds
 * the fitted normals of a stored result for resample() re-entry (src/resample.jl:20-46) without
 * an ELBO stage.  Needs pfb_batch_upload first. */
int pfb_batch_fit_only(pfb_handle h, const int64_t* best_iter);

/* K_new fresh draws per path from its best-iteration normal, one UInt64 seed per path:
 * rand(rng, fit_distribution, K_new) of src/singlepath.jl:228-230 (top-up draws when ndraws >
 * ndraws_elbo) and src/resample.jl:102-109 (resample with ndraws_per_run), together with logp(x)
 * and logq = logpdf(fit, x) (src/resample.jl:81-95).  Host outputs draws[n x K_new x P],
 * logp / logq[K_new x P] may be NULL.  keep_as_pool != 0 makes these draws the device pool of
 * pfb_psis_resample (N = P * K_new). */
int pfb_draw_from_fits(pfb_handle h, int K_new, const uint64_t* seeds, double* draws, double* logp,
                       double* logq, int keep_as_pool);

/* The ELBOEstimate payload of src/elbo.jl:22-29 on demand: draws[n x K x nunits], logp / logq
 * [K x nunits] of arbitrary units (0-based unit indices, path-major / iteration-minor) of the
 * current batch, regenerated from their seeds — bit-identical to what the ELBO stage evaluated.
 * Any output pointer may be NULL. */
int pfb_unit_draws(pfb_handle h, int nunits, const int32_t* units, double* draws, double* logp, double* logq);

/* fit_distributions[l + 1] (src/singlepath.jl:64: the reference keeps every iteration's) of arbitrary
 * units (0-based, path-major / iteration-minor) of the current batch, on demand, in the layout of the
 * fit_* fields of pfb_elbo_out with P replaced by nunits.  Any output pointer may be NULL. */
int pfb_unit_fits(pfb_handle h, int nunits, const int32_t* units, double* mu, double* alpha, double* vh,
                  double* T, double* Vc, double* logdet, int32_t* jeff);

/* Paths [p0, p1) of the device pool (best-iteration draws [n x K x (p1-p0)], their logp / logq
 * [K x (p1-p0)]; K = ndraws_elbo, or K_new after pfb_draw_from_fits(keep_as_pool)): the lazy form of
 * pfb_elbo_out.draws — PathfinderResult.draws (src/singlepath.jl:231-232) fetched on first use.
 * Any pointer may be NULL. */
int pfb_pool_download(pfb_handle h, int p0, int p1, double* draws, double* logp, double* logq);

/* Replaces _compute_psis_result + _resample (src/multipath.jl:220-225, src/resample.jl:58-95)
 * on the pool produced by the last batch (N = P * K draws; log ratios reuse the ELBO stage's
 * logp - logq, which src/resample.jl:81-95 recomputes).  importance = 0: uniform resampling
 * (psis_result === nothing, src/resample.jl:61).  replace = 0: sampling without replacement
 * (the `replace` keyword of resample, src/resample.jl:25 -> StatsBase.sample(...; replace),
 * src/resample.jl:61-66); ndraws > N is then an argument error like StatsBase's. */
int pfb_psis_resample(pfb_handle h, uint64_t seed, int ndraws, int importance, int replace,
                      pfb_resample_out* out);

/* PSIS + resampling on caller-supplied log ratios (host): log_ratios[N]; pool_or_null[n x N].
 * Used by resample() re-entry (src/resample.jl:20-46) and by the parity tests. */
int pfb_psis_resample_host(pfb_handle h, int n, int64_t N, int K_run, const double* log_ratios,
                           const double* pool_or_null, uint64_t seed, int ndraws, int importance,
                           int replace, pfb_resample_out* out);

/* Device views for multi-GPU plumbing (NCCL all-gather of the pool by the host layer). */
typedef struct {
    void* pool_draws; /* double [n x K x P]  */
    void* pool_logp;  /* double [K x P]      */
    void* pool_logq;  /* double [K x P]      */
    void* elbo;       /* double [U]          */
    void* stream;     /* cudaStream_t        */
    int64_t n, K, P, U;
} pfb_device_view;
int pfb_batch_device_view(pfb_handle h, pfb_device_view* view);

/* The pool's log densities are always resident after pfb_batch_run; its DRAWS are materialised on
 * demand (a draws download, pfb_pool_download, or this call) — pfb_psis_resample regenerates only the
 * columns it selects.  Call before reading pfb_device_view.pool_draws. */
int pfb_pool_materialize(pfb_handle h);
/* Multi-GPU resampling without materialising the pool: d_inds[m] = 1-based indices into the GLOBAL
 * pool (device, int64), this engine's runs cover [base, base + P K); its columns are regenerated into
 * d_out [n x m] (device), the others left untouched (zero d_out first, then sum-reduce over the ranks). */
int pfb_pool_columns_device(pfb_handle h, int m, const void* d_inds, int64_t base, void* d_out);

/* ---- multi-GPU: paths shard across GPUs, one exchange for the PSIS pool (src/multipath.jl:190-225) -------
 * One handle per GPU.  The communicator is NCCL, loaded lazily with dlopen("libnccl.so.2") (inside a
 * PyTorch process that is the NCCL torch already loaded), so a Julia / C caller gets multi-GPU without
 * MPI bindings of its own:
 *   - one process per GPU: rank 0 calls pfb_comm_unique_id, ships the 128 bytes to the other ranks by any
 *     means (MPI, torch.distributed, a file), every rank calls pfb_comm_init(h, id, rank, world);
 *   - one process, all GPUs: pfb_comm_init_all(handles, ndev) (handle i on device i is rank i), then the
 *     *_all form of the exchange from the same thread.
 * pfb_pool_exchange_resample replaces _compute_psis_result + _resample (src/multipath.jl:220-225) over
 * the runs of ALL ranks: paths_per_rank[world] runs per rank in run order (the pool keeps the
 * reference's component order, src/multipath.jl:217).  It all-gathers the per-draw log densities of
 * the pools (16 B per pool draw), runs PSIS and the index draw replicated on every rank (deterministic
 * kernels + counter RNG => identical results everywhere), lets every rank produce the selected columns
 * it owns (regenerated from their seeds; the pools' draws never move) and sum-reduces the n x ndraws
 * result.  No host synchronisation between the steps; every rank receives the same outputs. */
#define PFB_COMM_ID_BYTES 128
int pfb_comm_unique_id(void* id128);
int pfb_comm_init(pfb_handle h, const void* id128, int rank, int world);
int pfb_comm_init_all(pfb_handle* handles, int ndev);
int pfb_comm_destroy(pfb_handle h);
int pfb_pool_exchange_resample(pfb_handle h, const int32_t* paths_per_rank, uint64_t seed, int ndraws,
                               int importance, int replace, pfb_resample_out* out);
int pfb_pool_exchange_resample_all(pfb_handle* handles, int ndev, const int32_t* paths_per_rank, uint64_t seed,
                                   int ndraws, int importance, int replace, pfb_resample_out* outs /*[ndev]*/);
/* A device pool from HOST arrays (pools assembled on the host: top-up draws beyond ndraws_elbo, retried
 * paths, a rank that owns no run: P = 0): draws[n x K_run x P], logp / logq[K_run x P]; it then takes
 * part in pfb_psis_resample / pfb_pool_exchange_resample like the pool of a batch. */
int pfb_pool_set(pfb_handle h, int P, int K_run, const double* draws, const double* logp, const double* logq);

/* PSIS + resampling on DEVICE buffers (an all-gathered pool): d_logp/d_logq [N], d_pool [n x N];
 * outputs in `out` are HOST pointers. */
int pfb_psis_resample_device(pfb_handle h, int n, int64_t N, int K_run, const void* d_logp,
                             const void* d_logq, const void* d_pool, uint64_t seed, int ndraws,
                             int importance, int replace, pfb_resample_out* out);

/* Page-lock (cudaHostRegister) / unlock a caller-owned host buffer.  Optional: output buffers a
 * caller reuses across batches then receive their copies at full PCIe rate. */
int pfb_host_register(void* p, size_t bytes);
int pfb_host_unregister(void* p);

/* Timings of the last batch in milliseconds (CUDA events on the engine stream):
 * ms[0..5] = K1, K2, K3, K4, K5, total; returns the number of kernels launched. */
int pfb_get_timings(pfb_handle h, double* ms6);

/* Measurement utility: FP64 FMA peak of `device` in TFLOP/s (DFMA chains, best of reps). */
int pfb_measure_fp64_fma_tflops(int device, int reps, double* tflops);
/* The same for the FP64 tensor cores (mma.sync m8n8k4 f64, SASS DMMA): the roofline denominator
 * of K3, whose Q-apply runs on them. */
int pfb_measure_fp64_dmma_tflops(int device, int reps, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* PFB200_H */
