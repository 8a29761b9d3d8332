/*
 * gpx.h -- C ABI of the B200-native exact-GP hot path (libgpx.so).
 *
 * The reference (tmpethick/thesis-code) has no FFI of its own on this path: `GPModel` calls the Python
 * library GPy 1.9.6, which calls LAPACK/BLAS.  Each entry point below therefore cites the *reference call
 * site* (paths relative to /root/reference) whose arithmetic it replaces.  A maintainer binds these with the
 * ctypes stub shown in INTEGRATION.md (the in-repo binding is thesis_code_b200/_gpx.py).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary.
 *   - every array argument is float64, C-contiguous (row-major), exactly as GPModel receives numpy arrays
 *     (src/models/core_models.py:58-63, 185-187).
 *   - `mem` says where the caller's buffers live: GPX_MEM_HOST (numpy arrays, copied with cudaMemcpyAsync on the
 *     handle's stream; the small-batch predict path stages through a persistent pinned buffer) or GPX_MEM_DEVICE
 *     (pointers into this handle's GPU, e.g. torch tensors).
 *   - streams: every call runs on the handle's own streams and returns after they have drained, so results are
 *     visible to any stream afterwards.  Device inputs produced on another stream must be ordered BEFORE the call:
 *     either synchronise that stream or hand it to gpx_set_stream(), after which every call first waits (on the
 *     device, via an event) for the work already enqueued there.
 *   - return 0 = OK; > 0 = LAPACK-style info (1-based index of the first non-positive pivot of Ky, cf. GPy
 *     util.linalg.jitchol, which the Python layer mimics by retrying with more jitter); < 0 = bad argument /
 *     CUDA / NCCL error, text in gpx_last_error().  No exceptions cross the ABI.
 *   - one handle is used by one host thread at a time.  The handle owns the factor storage (packed lower
 *     triangular 128x128 tiles, factored in place), the inverted diagonal tiles, alpha, workspaces, streams
 *     and the NCCL communicator.
 *   - there is NO CPU fallback: every entry point that computes launches sm_100a kernels.
 *
 * Differences from the ABI sketched in SURVEY.md section 8(b) (a proposal: the reference has no FFI to mirror):
 *   - no per-call cudaStream_t argument: a handle owns three streams (update / panel factorisation / communication) whose
 *     interplay is the look-ahead, so a caller's stream is attached once with gpx_set_stream() and every call orders itself
 *     after it on the device; buffers may be host or device memory (`mem`), so numpy callers need no torch at all;
 *   - no `precision` argument: the path is fp64 only.  The optional fp32-with-jitter mode cannot meet its own 1e-4 tolerance on
 *     the BASELINE configs (measured: DESIGN.md section 1, profiles/r02_fp32_probe.txt);
 *   - gpx_export_factor is gpx_export (alpha and, for small N, the dense factor); gpx_create takes one device (one process
 *     per GPU; the ranks are joined by gpx_comm_init).
 */
#ifndef GPX_H_
#define GPX_H_

#include <stdint.h>

#if defined(__GNUC__)
#define GPX_API __attribute__((visibility("default")))
#else
#define GPX_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpx_handle gpx_t;

/* kernel ids: the names src/kernels.py:4-7 exposes to the experiment config dict */
enum { GPX_KERNEL_RBF = 0, GPX_KERNEL_MATERN52 = 1, GPX_KERNEL_MATERN32 = 2, GPX_KERNEL_EXPONENTIAL = 3 };
enum { GPX_MEM_HOST = 0, GPX_MEM_DEVICE = 1 };
/* phase ids for gpx_phase_ms() */
enum { GPX_PHASE_KBUILD = 0, GPX_PHASE_CHOLESKY = 1, GPX_PHASE_SOLVE = 2, GPX_PHASE_LML = 3,
       GPX_PHASE_PREDICT_KSTAR = 4, GPX_PHASE_PREDICT_TRSM = 5, GPX_PHASE_PREDICT_REDUCE = 6,
       GPX_PHASE_H2D = 7, GPX_PHASE_D2H = 8, GPX_NUM_PHASES = 9 };

GPX_API int gpx_version(void);

/* Create / destroy a handle bound to CUDA device `device` (one process per GPU). */
GPX_API int gpx_create(gpx_t** out, int device);
GPX_API void gpx_destroy(gpx_t* h);
GPX_API const char* gpx_last_error(const gpx_t* h);

/* Multi-GPU (block-column-cyclic factor over `world` ranks of one node, one process per GPU).
 * gpx_comm_unique_id fills a 128-byte NCCL unique id on rank 0; the caller ships it to the other ranks
 * (torch.distributed broadcast) and every rank calls gpx_comm_init.  No reference counterpart: the reference
 * is single-process CPU code. */
GPX_API int gpx_comm_unique_id(char* out128);
GPX_API int gpx_comm_init(gpx_t* h, int rank, int world, const char* uid128);

/*
 * FIT.  Replaces GPy.models.GPRegression(X, Y, kernel, mean_function) + Gaussian_noise.fix(noise) as called at
 * src/models/core_models.py:201-209, i.e. ExactGaussianInference.inference:
 *     Ky = K(X,X) + (noise + jitter) I ;  L = chol(Ky) (lower) ;  alpha = Ky^-1 (Y - mean_const) ;
 *     logdet = 2 sum log L_ii ;  LML = 0.5 (-N O log 2pi - O logdet - sum(alpha * (Y - mean_const))).
 * X is (N, d), Y is (N, O).  `ard` is the kernel kwarg `ARD` (core_models.py:167-174,192): 1 -> `lengthscale` has
 * n_ls = d entries and the inputs are divided by them BEFORE the Gram-trick distance; 0 -> n_ls = 1 and the distance of
 * the unscaled inputs is divided by it AFTERWARDS -- GPy's Stationary._scaled_dist, operation for operation (the two
 * orders differ by ~1e-9 in r^2 when |x / l| is large).
 * `noise` = Gaussian_noise.variance; `jitter` = GPy's fixed 1e-8; `mean_const` = mean(Y) when the config sets
 * mean_prior (core_models.py:194-196), else 0.
 */
GPX_API int gpx_fit(gpx_t* h, const double* X, int64_t N, int d, const double* Y, int O, int kernel_id,
            double variance, const double* lengthscale, int n_ls, int ard, double noise, double jitter,
            double mean_const, int mem);

/*
 * Fused NormalizerModel (src/models/normalizer.py:36-48,51-114; notebook_header.py:17-26 wraps every model in it).
 * After gpx_set_normalizer(h, nx, ny) every gpx_fit computes the column means / population standard deviations of the
 * raw X (nx) and Y (ny) on the device, z-scores X in the input loader of the K-build and Y in the residual kernel, and
 * every predict z-scores the test inputs in its loader and de-normalises in its epilogue (mean * std_y + mean_y,
 * var * std_y^2) -- no normalised copy of the data exists on the host.  gpx_get_normalizer returns the statistics of
 * the last fit (any pointer may be NULL; x_* have d entries, y_* have O).
 */
GPX_API int gpx_set_normalizer(gpx_t* h, int normalize_input, int normalize_output);
GPX_API int gpx_get_normalizer(gpx_t* h, double* x_mean, double* x_std, double* y_mean, double* y_std);

/* Order every later call after the work already enqueued on `stream` (a cudaStream_t; NULL = the legacy default
 * stream; call with has_stream = 0 to switch the waiting off again). */
GPX_API int gpx_set_stream(gpx_t* h, void* stream, int has_stream);

/*
 * APPEND k observations to the fitted model without refactorising (single-GPU handles, no normaliser).  O(k N^2)
 * instead of the O(N^3) refit that BaseModel.add_observations performs (src/models/core_models.py:65-77; called once per
 * Bayesian-optimisation step from src/algorithms.py:120-143).  Up to 8 points that fit into the partially filled last
 * 128-row tile: k forward substitutions in one pass over L, the k x k Schur complement and the new rows of L / of the
 * inverted diagonal tile in one finishing CTA, the backward substitution -- no tile work.  Otherwise: the tile rows that
 * receive observations are rebuilt, solved against the resident factor (the DMMA triangular solve of the predictive
 * variance, in place), the Schur complement is factored, alpha and the LML are recomputed.  Returns 0, > 0 (non-positive
 * pivot: the caller refits with jitter), -8 when the handle cannot append (sharded / normalised: the caller refits).
 */
GPX_API int gpx_append(gpx_t* h, const double* X_new, int64_t k, const double* Y_new, int mem);

/* log marginal likelihood / log-determinant / quadratic form of the last fit (host scalars).
 * Replaces gpy_model.log_likelihood() (name from MarginalLogLikelihoodMixin, core_models.py:12-14). */
GPX_API int gpx_lml(gpx_t* h, double* lml, double* logdet, double* quad);

/*
 * PREDICT.  Replaces gpy_model.predict(Xnew, full_cov=True) followed by np.diagonal, as called from
 * GPModel._predict (src/models/core_models.py:249-263, 270-271):
 *     mean[j, o] = sum_i k(x_i, x*_j) alpha[i, o] + mean_const
 *     var[j]     = variance - sum_i T_ij^2 + (include_noise ? noise : 0),   T = L^-1 K(X, X*)
 * Xs is (Ns, d); mean is (Ns, O); var is NULL (skip the variance pass), (Ns) with var_cols = 1 (GPy: one variance
 * shared by all outputs) or (Ns, O) with var_cols = O (required when the output normaliser is on and O > 1: each
 * output's variance is scaled by its own std^2).
 * Ns <= 128 on a single-GPU handle takes the latency path (src/growth_model_GPR/ipopt_wrapper.py:55 and
 * src/acquisition_functions.py:38,79 predict one or a few rows at a time, by the million): persistent workspaces and a
 * pinned staging buffer (no allocation, no cudaMemGetInfo), the variance of <= 8 rows by GEMV-style forward
 * substitution (L is read once, no tensor-core tile work) or, once the explicit inverse of the factor exists (option
 * "latency_inverse"), by one triangular GEMV over it (then up to 64 rows go that way, in groups of 8); for <= 8 rows the
 * loader, k(x*, X) and the mean are one kernel; host buffers are read / written in place through a mapped pinned buffer;
 * the whole launch chain is replayed as a CUDA graph.
 */
GPX_API int gpx_predict(gpx_t* h, const double* Xs, int64_t Ns, double* mean, double* var, int var_cols,
                int include_noise, int mem);

/* Full predictive covariance (Ns x Ns, row-major) for small Ns: core_models.py:244-259 (`full_cov=True`). */
GPX_API int gpx_predict_cov(gpx_t* h, const double* Xs, int64_t Ns, double* mean, double* cov, int include_noise,
                    int mem);

/*
 * Gradient of the log marginal likelihood of the last fit w.r.t. [variance, lengthscale (1 or d), noise, mean_const]
 * (n_ls + 3 values; returns that count, or < 0).  Replaces the gradient GPy's optimize() evaluates each L-BFGS step when
 * GPModel is built with do_optimize=True (src/models/core_models.py:211-223): dL_dK = 0.5 (alpha alpha^T - O Ky^-1),
 * Ky^-1 from the tiled factor (L^-T by the DMMA triangular solve, then Y Y^T), reduced against dK/dtheta on the fly.
 * The workspaces stay in the handle between calls (an optimiser calls this once per iteration).  On a sharded handle the
 * call is a collective: it needs the whole factor resident on every rank (the panel cache of gpx_fit; -6 otherwise), splits
 * the rows of L^-T and the row groups of Ky^-1 over the ranks and all-reduces the sums -- every rank gets identical bits.
 */
GPX_API int gpx_lml_grad(gpx_t* h, double* grad, int capacity);

/*
 * Moments of the posterior mean for its Jacobian / Hessian (GPModel.predict_jacobian / predict_hessian,
 * src/models/core_models.py:313-380, consumed by src/acquisition_functions.py:86,113), first output only like the
 * reference.  With w_ji = k(x_j, x*_i) alpha_j and dx_jik = x*_ik - x_jk in the coordinates the kernel sees
 * (z-scored / ARD-scaled if applicable):
 *     S0[i] = sum_j w_ji,   S1[i][k] = sum_j w_ji dx_jik,   S2[i][k][l] = sum_j w_ji dx_jik dx_jil  (NULL to skip).
 * The host layer turns them into the reference's expressions.  Fixed summation order (deterministic).
 */
GPX_API int gpx_predict_moments(gpx_t* h, const double* Xs, int64_t Ns, double* S0, double* S1, double* S2, int mem);

/*
 * Self-check of the fitted state on the device (bench.py prints it; SURVEY.md section 8(d) "Parity check"): with Ky
 * rebuilt tile by tile from the inputs (never stored),
 *     out[0] = || Ky alpha - (y - m) ||_2 / || y - m ||_2          (the solve)
 *     out[1] = || L (L^T v) - Ky v ||_2 / || Ky v ||_2              (the factor; v = fixed pseudo-random vector)
 *     out[2] = || Ky v ||_2,  out[3] = || y - m ||_2.
 * Collective on a sharded handle (every rank calls it; every rank gets the same numbers).
 */
GPX_API int gpx_selfcheck(gpx_t* h, double* out, int capacity);

/* Export for pickling / tests: alpha (N x O row-major) and, if L != NULL, the dense row-major lower factor
 * (N x N, upper triangle zeroed; only sensible for small N).  Host buffers. */
GPX_API int gpx_export(gpx_t* h, double* alpha, double* L);
/* Export K(X,X)+(noise+jitter)I as built by the K-build kernel (before factorisation), dense row-major, small N
 * only: builds into scratch, does not disturb the fitted state.  Host buffer.  Used by the parity tests. */
GPX_API int gpx_kernel_matrix(gpx_t* h, const double* X, int64_t N, int d, const double* X2, int64_t M, int kernel_id,
                      double variance, const double* lengthscale, int n_ls, int ard, double diag_add, double* K_out);

/* Timing of the last fit / predict, per phase, measured with CUDA events on the handle's stream (ms). */
GPX_API double gpx_phase_ms(const gpx_t* h, int phase);
/* Kernel launches issued by this handle since creation (the `gpu_launches` claim in bench.py). */
GPX_API int64_t gpx_launch_count(const gpx_t* h);
/* Tuning knobs (tests exercise several): "outer_tiles" (tile columns per outer Cholesky panel = ownership unit of the
 * block-cyclic layout and K/128 of the trailing update; 0 = auto: 4 when sharded, else min(16, max(4, nt/24))),
 * "predict_batch_tiles" (test tiles per variance batch), "sync_phases" (1: bracket phases with events),
 * "gemm_timing" (1: per-launch events on the tile-GEMM kernel, see gpx_gemm_stats), "lookahead" (0: serialise the
 * factor / communication / update streams), "gemm_max_chunk" (cap on tile slots per CTA), "group_panels" (panels per delayed far update when sharded, default 1),
 * "nccl_lo_ctas" (CTA cap of the factorisation communicator, default 4; before gpx_comm_init), "cache_mb" (sharded
 * handles keep received panels resident for later predicts: -1 = as many as fit beside a reserve, 0 = none, else a cap
 * in MB), "trace" (see gpx_trace), "reserve_rows" (row capacity laid out beyond N at fit time so that gpx_append works in
 * place; single GPU), "predict_graph" (1, default: replay the small-batch predict launch chain as a CUDA graph),
 * "panel_inverse" (in-panel variance solve through the inverted W x W-tile diagonal blocks: -1 auto = predicts with >= 64 test
 * tiles on one GPU, 0 off, 1 on), "gemm_strip" (rasterisation strip width of the tile GEMM, 0 = 16), "gemm_interleave" (1,
 * default: resident CTAs walk neighbouring tile slots), "nccl_full_from_pct" (panel broadcasts switch from the low-CTA to the
 * full-width communicator at this percentage of the panels, default 72), "latency_inverse" (explicit inverse of the factor,
 * L^-T as a packed-upper tile matrix, for the latency path: the variance of <= 8 rows and the <= 8-point gpx_append become
 * triangular GEMV sweeps at HBM rate instead of nt dependent substitution steps; costs N^3/3 flop and the factor's memory
 * once per fit; appends keep it up to date (re-layouts of the storage drop it).  -1 auto = built by the 16th small variance predict of a
 * fit if it fits in 60 % of the free memory, 0 = never (drops a built one), 1 = at the first such predict). */
GPX_API int gpx_set_option(gpx_t* h, const char* name, int64_t value);
/* The CUDA stream (cudaStream_t, returned as void*) every kernel of this handle is launched on, so that a caller can
 * bracket calls with its own CUDA events (bench.py does, via torch.cuda.ExternalStream). */
GPX_API void* gpx_stream(const gpx_t* h);
/* With option "gemm_timing"=1 every launch of the DMMA tile-GEMM kernel on the handle's main stream (the trailing /
 * far updates of the Cholesky and the whole variance solve; the small panel-factorisation GEMMs run concurrently on a
 * second stream and are left out so that no wall time is counted twice) is bracketed by CUDA events; this returns
 * (and resets) the accumulated device time in ms, the fp64 flops those launches performed (2*128^3 per output tile
 * and k-tile) and their count.  Used for the roofline of the dominant kernel. */
GPX_API int gpx_gemm_stats(gpx_t* h, double* total_ms, double* issued_flops, int64_t* launches);
/* With option "trace"=1 the Cholesky records a per-panel timeline: 6 doubles per outer panel (ms since the start of
 * the factorisation on this rank): factor start, factor end (owner only), panel ready on the update stream, next-panel
 * update done, fused forward steps done, trailing update done.  Returns the number of values (copies up to capacity). */
GPX_API int64_t gpx_trace(const gpx_t* h, double* out, int64_t capacity);
/* Device-memory footprint in bytes of the current fit (factor + workspaces). */
GPX_API int64_t gpx_bytes_allocated(const gpx_t* h);

#ifdef GPX_ENABLE_DEBUG_API
/* Development aids (tools/gemm_bench.py), only exported when the library is built with `make DEBUG_API=1`:
 * isolated timing of the DMMA tile GEMM (ms per launch) and of the diagonal-tile POTRF+inverse kernel (us). */
GPX_API double gpx_debug_gemm(gpx_t* h, int rows, int cols, int ktiles, int overwrite, int max_chunk, int reps);
GPX_API double gpx_debug_potrf(gpx_t* h, int reps);
#endif

#ifdef __cplusplus
}
#endif
#endif /* GPX_H_ */
