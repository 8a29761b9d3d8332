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
 *   - `mem` says where the caller's buffers live: GPX_MEM_HOST (numpy arrays; the library stages them through
 *     pinned memory on its stream) or GPX_MEM_DEVICE (pointers into this handle's GPU, e.g. torch tensors).
 *   - return 0 = OK; > 0 = LAPACK-style info (1-based index of the first non-positive pivot of Ky, cf. GPy
 *     util.linalg.jitchol, which the Python layer mimics by retrying with more jitter); < 0 = bad argument /
 *     CUDA / NCCL error, text in gpx_last_error().  No exceptions cross the ABI.
 *   - one handle is used by one host thread at a time.  The handle owns the factor storage (packed lower
 *     triangular 128x128 tiles, factored in place), the inverted diagonal tiles, alpha, workspaces, streams
 *     and the NCCL communicator.
 *   - there is NO CPU fallback: every entry point that computes launches sm_100a kernels.
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
 * X is (N, d), Y is (N, O).  `lengthscale` has n_ls = 1 (isotropic) or d (ARD) entries
 * (kernel kwargs `variance`, `lengthscale`, `ARD` of the config dict, core_models.py:167-174,192).
 * `noise` = Gaussian_noise.variance; `jitter` = GPy's fixed 1e-8; `mean_const` = mean(Y) when the config sets
 * mean_prior (core_models.py:194-196), else 0.
 */
GPX_API int gpx_fit(gpx_t* h, const double* X, int64_t N, int d, const double* Y, int O, int kernel_id,
            double variance, const double* lengthscale, int n_ls, double noise, double jitter,
            double mean_const, int mem);

/* log marginal likelihood / log-determinant / quadratic form of the last fit (host scalars).
 * Replaces gpy_model.log_likelihood() (name from MarginalLogLikelihoodMixin, core_models.py:12-14). */
GPX_API int gpx_lml(gpx_t* h, double* lml, double* logdet, double* quad);

/*
 * PREDICT.  Replaces gpy_model.predict(Xnew, full_cov=True) followed by np.diagonal, as called from
 * GPModel._predict (src/models/core_models.py:249-263, 270-271):
 *     mean[j, o] = sum_i k(x_i, x*_j) alpha[i, o] + mean_const
 *     var[j]     = variance - sum_i T_ij^2 + (include_noise ? noise : 0),   T = L^-1 K(X, X*)
 * Xs is (Ns, d); mean is (Ns, O); var is (Ns) or NULL to skip the variance pass.
 */
GPX_API int gpx_predict(gpx_t* h, const double* Xs, int64_t Ns, double* mean, double* var, int include_noise, int mem);

/* Full predictive covariance (Ns x Ns, row-major) for small Ns: core_models.py:244-259 (`full_cov=True`). */
GPX_API int gpx_predict_cov(gpx_t* h, const double* Xs, int64_t Ns, double* mean, double* cov, int include_noise,
                    int mem);

/*
 * Gradient of the log marginal likelihood of the last fit w.r.t. [variance, lengthscale (1 or d), noise, mean_const]
 * (n_ls + 3 values; returns that count, or < 0).  Replaces the gradient GPy's optimize() evaluates each L-BFGS step when
 * GPModel is built with do_optimize=True (src/models/core_models.py:211-223): dL_dK = 0.5 (alpha alpha^T - O Ky^-1),
 * Ky^-1 from the tiled factor (L^-T by the DMMA triangular solve, then Y Y^T), reduced against dK/dtheta on the fly.
 */
GPX_API int gpx_lml_grad(gpx_t* h, double* grad, int capacity);

/* Export for pickling / tests: alpha (N x O row-major) and, if L != NULL, the dense row-major lower factor
 * (N x N, upper triangle zeroed; only sensible for small N).  Host buffers. */
GPX_API int gpx_export(gpx_t* h, double* alpha, double* L);
/* Export K(X,X)+(noise+jitter)I as built by the K-build kernel (before factorisation), dense row-major, small N
 * only: builds into scratch, does not disturb the fitted state.  Host buffer.  Used by the parity tests. */
GPX_API int gpx_kernel_matrix(gpx_t* h, const double* X, int64_t N, int d, const double* X2, int64_t M, int kernel_id,
                      double variance, const double* lengthscale, int n_ls, double diag_add, double* K_out);

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
 * in MB), "trace" (see gpx_trace). */
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

#ifdef __cplusplus
}
#endif
#endif /* GPX_H_ */
