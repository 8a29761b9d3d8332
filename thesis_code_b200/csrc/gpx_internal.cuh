// Internal declarations shared by the translation units that implement the C ABI (gpx_api.cu: handle, fit, predict;
// gpx_update.cu: incremental append, latency path, self-check, posterior-mean derivatives).  Not installed.
#pragma once
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gpx.h"
#include "gpx_common.cuh"

// launchers from the other translation units
void gpx_kbuild_init();
int gpx_kbuild_max_d();
int gpx_grad_max_d();
void gpx_launch_kbuild_lower(TileMat L, int nt, int jfirst, int jstep, int ncols, const double* Xs, const double* xsq,
                             int64_t N, int d, int kernel_id, double variance, double rdiv, double diag_add, cudaStream_t st);
// ystats: null, or [O means | O stds] of the fused output normaliser (mean -> mean * std + mean_y, var -> var * std^2)
void gpx_launch_mean_reduce(const double* mean_partial, int nbt, int nt, int O, double mean_const, const double* ystats,
                            double* mean, int64_t row0, int64_t Ns, cudaStream_t st);
int gpx_var_reduce_splits(int nbt, int nt, int num_sms);
void gpx_launch_var_reduce(TileMat Tt, int nbt, int nt, int S, double* partial, double prior_var, const double* ystats,
                           int var_cols, double* var, int64_t row0, int64_t Ns, cudaStream_t st);
bool gpx_small_kstar_ok(int64_t ns, int d, int O);
void gpx_launch_small_kstar(const double* Xin, int ns, int d, const double* ls, int ard, const double* shift, const double* scale,
                            const double* Xs, const double* xsq, int64_t N, int nt, int kernel_id, double variance, double rdiv,
                            const double* alpha, int O, int64_t Np, double mean_const, const double* ystats, double* vec_out,
                            double* partial, unsigned int* ticket, double* mean, cudaStream_t st);
void gpx_launch_detile(TileMat M, int lower, int64_t rows, int64_t cols, double* out, cudaStream_t st);
void gpx_launch_fwd_step(TileMat L, const double* linv, int nt, int J, double* b, double* z, int64_t Np, int O,
                         int max_ctas, cudaStream_t st);
void gpx_launch_bwd_step(TileMat L, const double* linv, int I, double* z, double* a, int64_t Np, int O,
                         int compute_alpha, int ncols, int jbase, int jW, int jstride, int max_ctas, cudaStream_t st);
void gpx_launch_lml(const double* linv, const double* alpha, const double* resid, int64_t n_rows, int64_t Np, int O,
                    double* out2, cudaStream_t st);
void gpx_launch_resid(const double* Y, int64_t N, int64_t Np, int O, double mean_const, const double* ystats, double* resid,
                      cudaStream_t st);
void gpx_launch_untranspose(const double* in, int64_t N, int64_t Np, int O, double* out, cudaStream_t st);
void gpx_launch_identity_tiles(TileMat T, int nt, cudaStream_t st);
void gpx_launch_identity_block(TileMat T, int w, int j0, cudaStream_t st);            // T(i, j0 + j) = (i == j) I, i, j < w
void gpx_launch_transpose_block(TileMat Y, TileMat M, int p0, int w, cudaStream_t st);  // M(p0+j, p0+k) = Y(k, p0+j)^T, k <= j < w
void gpx_launch_lml_grad(TileMat Kinv, int nt, const double* Xs, const double* xsq, int64_t N, int d, int kernel_id,
                         double variance, double rdiv, const double* ls, int n_ls, const double* alpha, int O, int64_t Np,
                         double* partial, double* out, cudaStream_t st, int rg = 8, int world = 1, int rank = 0);


struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

// NCCL is resolved at run time (dlopen) so that the library also loads where no libnccl is installed; the
// single-GPU path never touches it.
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, ncclConfig_t*) = nullptr;  // optional (NCCL >= 2.18)
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool tried = false, ok = false;
};


NcclApi& nccl();


struct gpx_handle {
    int device = 0;
    int num_sms = 148;
    cudaStream_t st = nullptr;    // main / update stream (U)
    cudaStream_t st_f = nullptr;  // panel factorisation (F), high priority
    cudaStream_t st_c = nullptr;  // communication (C), high priority
    std::string err;
    int64_t launches = 0;
    int64_t bytes_alloc = 0;

    // distribution
    int rank = 0, world = 1;
    ncclComm_t comm = nullptr;     // default configuration: bandwidth-bound transfers (predict re-broadcast, all-reduce)
    ncclComm_t comm_lo = nullptr;  // at most 4 CTAs: the panel broadcasts that overlap the trailing updates
    int nccl_lo_ctas = 4;
    int nccl_full_from_pct = 72;   // panel broadcasts use the full-width communicator from this percentage of the panels on

    // options
    int outer_tiles = 0;          // 0 = auto: 4 when sharded, else min(16, max(4, nt / 24)) (K of the trailing update)
    int predict_batch_tiles = 0;  // 0 = auto
    int sync_phases = 1;
    int gemm_timing = 0;
    int lookahead = 1;
    int gemm_max_chunk = 0;               // 0 = auto: 2 on one GPU, 4 when sharded (shorter CTAs let the panel
                                          // factorisation and NCCL kernels of the other streams in sooner)
    int gemm_strip = 0;                   // option "gemm_strip": rasterisation strip width of the tile GEMM (0 = default)
    int gemm_interleave = 1;              // option "gemm_interleave": interleaved slot assignment of the tile GEMM (L2 locality)
    int panel_inverse = -1;               // option "panel_inverse": in-panel variance solve through inverted diagonal blocks
                                          // (-1 auto: predicts with >= 64 test tiles on one GPU; 0 off; 1 on)
    int minv_active = 0;                  // set by gpx_predict around its batch loop
    DevBuf zkeep;                         // z = L^-1 (Y - m) of the last fit / append (the backward pass consumes work2)
    bool zkeep_valid = false;
    int minv_valid = 0;                   // leading panels whose inverted diagonal blocks are up to date
    int minv_layout_cap = -1, minv_layout_w = -1;
    DevBuf minv, minv_coloff, ywork, ywork_coloff;
    int trace = 0;                        // 1: per-panel timeline events during the Cholesky (gpx_trace)
    std::vector<cudaEvent_t> tev;         // 6 per panel + 1 base, timing enabled
    std::vector<float> trace_ms;          // 6 per panel: factor start, factor end, panel ready, (a) end, fwd end, (b) end
    std::vector<cudaEvent_t> gemm_ev;  // pairs, grown on demand
    size_t gemm_ev_used = 0;
    double gemm_ms = 0, gemm_flops = 0;
    int64_t gemm_launches = 0;

    // problem
    bool fitted = false;
    int64_t N = 0, Np = 0;           // Np = cap_nt * 128: padded length and stride of the [o][Np] vectors
    int nt = 0, d = 0, O = 0, kernel_id = 0, n_ls = 0;
    int cap_nt = 0;                  // tile rows the factor / vector storage is laid out for (>= nt; gpx_append grows into it)
    int64_t reserve_rows = 0;        // option "reserve_rows": extra row capacity laid out at fit time (single GPU)
    uint64_t gen = 0;                // bumped by every fit / append: captured predict graphs of older generations are stale
    int predict_graph = 1;           // option "predict_graph": replay the small-batch predict chain as a CUDA graph
    int W = 4, npanels = 0;  // W = outer_tiles frozen at fit time
    double variance = 1, noise = 1, jitter = 1e-8, mean_const = 0;
    double lml = 0, logdet = 0, quad = 0;
    int ard = 0;                     // 1: inputs are divided by the lengthscales before the Gram product (GPy ARD=True)
    double rdiv = 1.0;               // isotropic: r = sqrt(r^2 of the unscaled inputs) / rdiv (GPy _scaled_dist); ARD: 1
    std::vector<double> h_ls;        // host copy of the lengthscales

    // fused NormalizerModel (reference src/models/normalizer.py): column z-scores of X / Y computed on the device at fit
    // time and applied in the input loader (prep_x), the residual kernel and the predict epilogues
    int norm_x = 0, norm_y = 0;      // requested by gpx_set_normalizer; frozen into fit_norm_* by gpx_fit
    int fit_norm_x = 0, fit_norm_y = 0;
    DevBuf xstats, ystats;           // [d means | d stds], [O means | O stds]
    std::vector<double> h_xstats, h_ystats;

    cudaStream_t caller_stream = nullptr;   // gpx_set_stream: work already enqueued there is waited for at call entry
    bool has_caller_stream = false;
    cudaEvent_t ev_caller = nullptr;

    DevBuf Xs, xsq, ls, L, linv, coloff, vcoloff, linv_coloff, resid, work1, work2, alpha, scal, info;
    DevBuf stage;
    // persistent scratch of the calls an optimiser / acquisition loop repeats (cudaMalloc + cudaFree of the gradient's
    // workspaces cost 3x the device time of the gradient itself at N = 20000, 10x at N = 1000): Y staging of fit / append,
    // gpx_lml_grad (Ky^-1, offsets, per-tile partial sums, result), gpx_predict_cov, gpx_predict_moments
    DevBuf stage_y, g_kinv, g_kinv_off, g_partial, g_out, c_dm, c_dc, c_covt, c_covoff, m_xt, m_xtsq, m_o0, m_o1, m_o2;
    DevBuf gbuf[2];                     // group buffers (world > 1): the M panels of a group, contiguous; double buffered
    int group_panels = 1;               // M: panels per delayed far update when sharded (measured: no gain from M = 4,
                                        // 2587 vs 2591 ms at 4 GPUs / N = 100k, so the simpler M = 1 is the default)
    DevBuf cache;                       // received panels kept resident after the fit (world > 1) so that predict does not
    DevBuf ucoloff;                     // have to re-broadcast them; ucoloff = column offsets of own + cached columns
    std::vector<int64_t> h_ucoloff;     // relative to the base of L (pointer arithmetic across the two allocations)
    int cache_panels = 0;               // panels [0, cache_panels) are resident on every rank
    int64_t cache_limit_mb = -1;        // option "cache_mb": -1 auto (free memory minus a reserve), 0 disables
    std::vector<int64_t> h_coloff;      // local offsets (-1: not owned)
    std::vector<int64_t> h_vcoloff;     // virtual full packed offsets
    std::vector<cudaEvent_t> pev;       // per-panel events (3 per panel), no timing

    // predict workspaces
    DevBuf Tt, tt_coloff, Xt, xtsq, mean_partial, dmean, dvar, var_partial;
    // latency path (gpx_predict with <= 128 rows, gpx_update.cu): one persistent workspace, a pinned staging buffer and
    // the captured launch chains
    DevBuf sp;
    int64_t sp_key[3] = {-1, -1, -1};
    void* pin = nullptr;
    size_t pin_bytes = 0;
    struct SmallGraph { uint64_t gen; int64_t Ns; int want_var, var_cols, include_noise, nodes, hostio; cudaGraphExec_t exec; };
    std::vector<SmallGraph> graphs;
    // explicit inverse of the factor, Y = L^-T as a packed-upper tile matrix (column J holds the tile rows 0 .. end of J's
    // panel): the variance of <= 8 rows becomes ONE triangular GEMV over it instead of nt dependent substitution steps
    int latency_inverse = -1;   // option "latency_inverse": -1 auto (built by the 16th small variance predict of a fit), 0 off, 1 at the first
    bool finv_valid = false, finv_failed = false;
    int finv_calls = 0, finv_C = 1, finv_S = 1;
    int finv_layout_cap = -1, finv_layout_w = -1;
    DevBuf finv, finv_coloff, finv_partial;

    cudaEvent_t ev[2 * GPX_NUM_PHASES];
    float phase_ms[GPX_NUM_PHASES];
};


inline int fail(gpx_t* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define GPX_CUDA(h, call)                                                                                \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(h, -2, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #call); \
    } while (0)

#define GPX_NCCL(h, call)                                                                                \
    do {                                                                                                 \
        ncclResult_t r_ = (call);                                                                        \
        if (r_ != ncclSuccess)                                                                           \
            return fail(h, -7, "NCCL error %s at %s:%d (%s)", nccl().GetErrorString(r_), __FILE__, __LINE__, #call); \
    } while (0)

inline int ensure(gpx_t* h, DevBuf& b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return 0;
    if (b.p) {
        cudaFree(b.p);
        h->bytes_alloc -= (int64_t)b.bytes;
        b.p = nullptr;
        b.bytes = 0;
    }
    if (bytes == 0) return 0;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(h, -3, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    b.bytes = bytes;
    h->bytes_alloc += (int64_t)bytes;
    return 0;
}

inline void release(gpx_t* h, DevBuf& b);
// every persistent scratch buffer above (released by gpx_destroy and by the out-of-memory reclaim of gpx_fit)
inline void release_scratch(gpx_t* h) {
    for (DevBuf* b : {&h->stage_y, &h->g_kinv, &h->g_kinv_off, &h->g_partial, &h->g_out, &h->c_dm, &h->c_dc, &h->c_covt, &h->c_covoff,
                      &h->m_xt, &h->m_xtsq, &h->m_o0, &h->m_o1, &h->m_o2})
        release(h, *b);
}
inline void release(gpx_t* h, DevBuf& b) {
    if (b.p) {
        cudaFree(b.p);
        h->bytes_alloc -= (int64_t)b.bytes;
    }
    b.p = nullptr;
    b.bytes = 0;
}

template <typename T>
inline T* ptr(const DevBuf& b) { return reinterpret_cast<T*>(b.p); }

inline void phase_begin(gpx_t* h, int ph) { if (h->sync_phases) cudaEventRecord(h->ev[2 * ph], h->st); }
inline void phase_end(gpx_t* h, int ph) { if (h->sync_phases) cudaEventRecord(h->ev[2 * ph + 1], h->st); }
inline void phase_collect(gpx_t* h, int ph) {
    if (!h->sync_phases) return;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[2 * ph], h->ev[2 * ph + 1]) == cudaSuccess) h->phase_ms[ph] = ms;
    else cudaGetLastError();
}

// stage a caller buffer onto the device (host pointer -> device copy on the main stream, or passthrough)
inline int to_device(gpx_t* h, const double* src, size_t count, int mem, DevBuf& staging, const double** out) {
    if (mem == GPX_MEM_DEVICE) { *out = src; return 0; }
    size_t bytes = count * sizeof(double);
    int rc = ensure(h, staging, bytes);
    if (rc) return rc;
    GPX_CUDA(h, cudaMemcpyAsync(staging.p, src, bytes, cudaMemcpyHostToDevice, h->st));
    *out = ptr<double>(staging);
    return 0;
}

// fused normaliser views (null when the corresponding z-score is off for the current fit)
inline const double* x_shift(const gpx_t* h) { return h->fit_norm_x ? reinterpret_cast<const double*>(h->xstats.p) : nullptr; }
inline const double* x_scale(const gpx_t* h) { return h->fit_norm_x ? reinterpret_cast<const double*>(h->xstats.p) + h->d : nullptr; }
inline const double* y_stats(const gpx_t* h) { return h->fit_norm_y ? reinterpret_cast<const double*>(h->ystats.p) : nullptr; }
// load + (z-score) + (ARD scaling) + |x|^2 of `rows` input rows into (Xs_out, xsq_out), padded to rows_padded
inline void prep_inputs(gpx_t* h, const double* X, int64_t rows, int64_t rows_padded, double* Xs_out, double* xsq_out,
                        cudaStream_t st) {
    gpx_launch_prep_x(X, rows, rows_padded, h->d, reinterpret_cast<const double*>(h->ls.p), h->ard, x_shift(h), x_scale(h),
                      Xs_out, xsq_out, st);
    h->launches++;
}
// make the handle's main stream wait for everything the caller has enqueued on the stream given to gpx_set_stream
inline int wait_for_caller(gpx_t* h) {
    if (!h->has_caller_stream) return 0;
    if (!h->ev_caller) GPX_CUDA(h, cudaEventCreateWithFlags(&h->ev_caller, cudaEventDisableTiming));
    GPX_CUDA(h, cudaEventRecord(h->ev_caller, h->caller_stream));
    GPX_CUDA(h, cudaStreamWaitEvent(h->st, h->ev_caller, 0));
    return 0;
}

void gemm(gpx_t* h, const GemmJob& job_in, cudaStream_t st);
void gemm_collect(gpx_t* h);

// var = prior - rowsumsq(Tt) for nb test tiles (two launches; the partial-sum buffer is kept in the handle)
inline int var_reduce(gpx_t* h, TileMat Tt, int nb, double prior_var, int var_cols, double* var, int64_t row0, int64_t Ns) {
    const int S = gpx_var_reduce_splits(nb, h->nt, h->num_sms);
    int rc = ensure(h, h->var_partial, (size_t)nb * S * GPX_TILE * 8);
    if (rc) return rc;
    gpx_launch_var_reduce(Tt, nb, h->nt, S, ptr<double>(h->var_partial), prior_var, y_stats(h), var_cols, var, row0, Ns, h->st);
    h->launches += 2;
    return 0;
}

// ---- distribution helpers ---------------------------------------------------------------------------------------
inline int panel_begin(const gpx_t* h, int p) { return p * h->W; }
inline int panel_end(const gpx_t* h, int p) { int e = (p + 1) * h->W; return e < h->nt ? e : h->nt; }
inline int panel_owner(const gpx_t* h, int p) { return p % h->world; }
inline bool owns_panel(const gpx_t* h, int p) { return panel_owner(h, p) == h->rank; }
inline int64_t panel_elems(const gpx_t* h, int p) {
    int64_t e = 0;
    for (int c = panel_begin(h, p); c < panel_end(h, p); c++) e += (int64_t)(h->nt - c) * GPX_TILE_ELEMS;
    return e;
}
// first panel >= q owned by this rank (may be >= npanels)
inline int next_owned_panel(const gpx_t* h, int q) {
    int r = q % h->world;
    int add = (h->rank - r + h->world) % h->world;
    return q + add;
}
// number of tile columns in owned panels q, q+world, q+2 world, ... (q must be owned)
inline int owned_cols_from(const gpx_t* h, int q) {
    int n = 0;
    for (int p = q; p < h->npanels; p += h->world) n += panel_end(h, p) - panel_begin(h, p);
    return n;
}

// receive buffers alternate over the panels this rank does NOT own
inline int recv_index(const gpx_t* h, int p) {
    int owned_before = (p > h->rank) ? (p - h->rank + h->world - 1) / h->world : 0;
    return p - owned_before;
}
inline double* recv_buf(gpx_t* h, int p) { return reinterpret_cast<double*>(h->gbuf[recv_index(h, p) & 1].p); }  // predict sweep
// the previous non-owned panel that used the same receive buffer as panel p (-1: none)
inline int prev_same_buffer_panel(const gpx_t* h, int p) {
    int seen = 0;
    for (int q = p - 1; q >= 0; q--) {
        if (!owns_panel(h, q) && ++seen == 2) return q;
    }
    return -1;
}

inline TileMat lmat(gpx_t* h) { return TileMat{ptr<double>(h->L), ptr<int64_t>(h->coloff), 1}; }
inline TileMat linvmat(gpx_t* h) { return TileMat{ptr<double>(h->linv), ptr<int64_t>(h->linv_coloff), 1}; }
// own + cached columns through one table (offsets relative to L's base)
inline TileMat umat(gpx_t* h) { return TileMat{ptr<double>(h->L), ptr<int64_t>(h->ucoloff), 1}; }
inline bool panel_cached(const gpx_t* h, int p) { return h->world > 1 && p < h->cache_panels && !owns_panel(h, p); }
// where a non-owned panel is received: its resident cache slot, or one of the two rotating buffers
inline double* panel_recv_ptr(gpx_t* h, int p) {
    if (panel_cached(h, p)) return ptr<double>(h->L) + h->h_ucoloff[panel_begin(h, p)];
    return recv_buf(h, p);
}
// the factored panel p as an operand: the local storage on its owner, the cache or a receive buffer elsewhere
inline TileMat panel_mat(gpx_t* h, int p) {
    if (owns_panel(h, p)) return lmat(h);
    if (panel_cached(h, p)) return umat(h);
    return TileMat{recv_buf(h, p) - h->h_vcoloff[panel_begin(h, p)], ptr<int64_t>(h->vcoloff), 1};
}

// per-panel events: 0 panel ready (factored / received), 3 factored on stream F; per group (stored at its first panel):
// 1 far update of the next group's panels done, 2 far update done (group buffer free).  The predict sweep re-uses 0 / 2.
inline cudaEvent_t pevent(gpx_t* h, int p, int which) { return h->pev[(size_t)4 * p + which]; }

inline int ensure_panel_events(gpx_t* h, int npanels) {
    size_t need = (size_t)4 * npanels;
    while (h->pev.size() < need) {
        cudaEvent_t e;
        GPX_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->pev.push_back(e);
    }
    return 0;
}


void var_trsm_panel(gpx_t* h, TileMat Tt, int nb_in, TileMat P, int p, bool upper_rows = false, int row0 = 0);
int var_trsm(gpx_t* h, TileMat Tt, int nb);
int backward_solve(gpx_t* h);
int ensure_panel_inverse(gpx_t* h, int upto_panel);
// step-wise in-panel solve of the tile rows [i0, nb) of X against the columns [p0, p1) of P
void trsm_inner_stepwise(gpx_t* h, TileMat X, int nb, TileMat P, int p0, int p1, int i0 = 0);   // inverted W x W-tile diagonal blocks of the panels [0, upto_panel)
// fits / appends invalidate the explicit inverse of the factor (gpx_update.cu)
inline void finv_invalidate(gpx_t* h) { h->finv_valid = false; h->finv_failed = false; h->finv_calls = 0; }
int gpx_predict_small(gpx_t* h, const double* Xs, int64_t Ns, double* mean, double* var, int var_cols, int include_noise,
                      int mem);
void gpx_launch_kbuild_rows(TileMat L, int nt, int i_first, const double* Xs, const double* xsq, int64_t N, int d,
                            int kernel_id, double variance, double rdiv, double diag_add, cudaStream_t st);
void factor_panel(gpx_t* h, int p);
void factor_cols(gpx_t* h, int c0, int c1, int nt, cudaStream_t st);
