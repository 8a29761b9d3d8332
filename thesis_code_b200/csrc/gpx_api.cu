// C ABI (include/gpx.h) and host-side orchestration of the exact-GP hot path.
//
// The handle owns: the scaled training inputs, the packed lower-triangular tile storage of Ky (factored in place
// into L), the inverted diagonal tiles, alpha, the predict workspaces and one CUDA stream.  The blocked
// right-looking Cholesky is driven from here: per tile column  POTRF(tile) -> TRSM (DMMA GEMM against the inverted
// tile) -> rank-128 update of the rest of the current outer panel; per outer panel one SYRK/GEMM trailing update
// with K = outer_tiles*128 (all DMMA).  See DESIGN.md.
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
void gpx_launch_kbuild_lower(TileMat L, int nt, int jfirst, int jstep, int ncols, const double* Xs, const double* xsq,
                             int64_t N, int d, int kernel_id, double variance, double diag_add, cudaStream_t st);
void gpx_launch_mean_reduce(const double* mean_partial, int nbt, int nt, int O, double mean_const, double* mean,
                            int64_t row0, int64_t Ns, cudaStream_t st);
void gpx_launch_var_reduce(TileMat Tt, int nbt, int nt, double prior_var, double* var, int64_t row0, int64_t Ns,
                           cudaStream_t st);
void gpx_launch_detile(TileMat M, int lower, int64_t rows, int64_t cols, double* out, cudaStream_t st);
void gpx_launch_fwd_step(TileMat L, const double* linv, int nt, int J, double* b, double* z, int64_t Np, int O,
                         int max_ctas, cudaStream_t st);
void gpx_launch_bwd_step(TileMat L, const double* linv, int I, double* z, double* a, int64_t Np, int O, int max_ctas,
                         cudaStream_t st);
void gpx_launch_lml(TileMat L, int nt, const int* owned, const double* alpha, const double* resid, int64_t Np, int O,
                    double* out2, cudaStream_t st);
void gpx_launch_resid(const double* Y, int64_t N, int64_t Np, int O, double mean_const, double* resid, cudaStream_t st);
void gpx_launch_untranspose(const double* in, int64_t N, int64_t Np, int O, double* out, cudaStream_t st);

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct gpx_handle {
    int device = 0;
    int num_sms = 148;
    cudaStream_t st = nullptr;
    std::string err;
    int64_t launches = 0;
    int64_t bytes_alloc = 0;

    // options
    int outer_tiles = 4;
    int predict_batch_tiles = 0;  // 0 = auto
    int sync_phases = 1;
    int gemm_timing = 0;
    std::vector<cudaEvent_t> gemm_ev;   // pairs, grown on demand
    size_t gemm_ev_used = 0;
    double gemm_ms = 0, gemm_flops = 0;
    int64_t gemm_launches = 0;

    // problem
    bool fitted = false;
    int64_t N = 0, Np = 0;
    int nt = 0, d = 0, O = 0, kernel_id = 0, n_ls = 0;
    double variance = 1, noise = 1, jitter = 1e-8, mean_const = 0;
    double lml = 0, logdet = 0, quad = 0;

    DevBuf Xs, xsq, ls, L, linv, coloff, linv_coloff, resid, work1, work2, alpha, scal, info;
    DevBuf stage;  // device staging for host inputs / outputs
    std::vector<int64_t> h_coloff;
    void* pinned = nullptr;
    size_t pinned_bytes = 0;

    // predict workspaces
    DevBuf Tt, tt_coloff, Xt, xtsq, mean_partial, dmean, dvar;

    cudaEvent_t ev[2 * GPX_NUM_PHASES];
    float phase_ms[GPX_NUM_PHASES];
};

namespace {

int fail(gpx_t* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    return code;
}

#define GPX_CUDA(h, call)                                                                              \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) return fail(h, -2, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), \
                                           __FILE__, __LINE__, #call);                                 \
    } while (0)

int ensure(gpx_t* h, DevBuf& b, size_t bytes) {
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

void release(gpx_t* h, DevBuf& b) {
    if (b.p) {
        cudaFree(b.p);
        h->bytes_alloc -= (int64_t)b.bytes;
    }
    b.p = nullptr;
    b.bytes = 0;
}

int ensure_pinned(gpx_t* h, size_t bytes) {
    if (h->pinned_bytes >= bytes) return 0;
    if (h->pinned) cudaFreeHost(h->pinned);
    h->pinned = nullptr;
    h->pinned_bytes = 0;
    GPX_CUDA(h, cudaMallocHost(&h->pinned, bytes));
    h->pinned_bytes = bytes;
    return 0;
}

template <typename T>
T* ptr(const DevBuf& b) { return reinterpret_cast<T*>(b.p); }

void phase_begin(gpx_t* h, int ph) { if (h->sync_phases) cudaEventRecord(h->ev[2 * ph], h->st); }
void phase_end(gpx_t* h, int ph) { if (h->sync_phases) cudaEventRecord(h->ev[2 * ph + 1], h->st); }
void phase_collect(gpx_t* h, int ph) {
    if (!h->sync_phases) return;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev[2 * ph], h->ev[2 * ph + 1]) == cudaSuccess) h->phase_ms[ph] = ms;
    else cudaGetLastError();
}

// stage a caller buffer onto the device (host -> pinned -> device, or device pointer passthrough)
int to_device(gpx_t* h, const double* src, size_t count, int mem, DevBuf& staging, const double** out) {
    if (mem == GPX_MEM_DEVICE) { *out = src; return 0; }
    size_t bytes = count * sizeof(double);
    int rc = ensure(h, staging, bytes);
    if (rc) return rc;
    // pageable numpy memory: cudaMemcpyAsync stages through the driver's pinned buffers; one copy per input
    GPX_CUDA(h, cudaMemcpyAsync(staging.p, src, bytes, cudaMemcpyHostToDevice, h->st));
    *out = ptr<double>(staging);
    return 0;
}

// every DMMA tile-GEMM launch goes through here (optionally bracketed by events for the roofline numbers)
void gemm(gpx_t* h, const GemmJob& job) {
    if (!h->gemm_timing) {
        long long steps = gpx_launch_gemm(job, h->num_sms, h->st, &h->launches);
        h->gemm_flops += 2.0 * 128 * 128 * 128 * (double)steps;
        if (steps) h->gemm_launches++;
        return;
    }
    if (h->gemm_ev_used + 2 > h->gemm_ev.size()) {
        size_t old = h->gemm_ev.size();
        h->gemm_ev.resize(old + 1024);
        for (size_t i = old; i < h->gemm_ev.size(); i++) cudaEventCreate(&h->gemm_ev[i]);
    }
    cudaEventRecord(h->gemm_ev[h->gemm_ev_used], h->st);
    long long steps = gpx_launch_gemm(job, h->num_sms, h->st, &h->launches);
    if (steps) {
        cudaEventRecord(h->gemm_ev[h->gemm_ev_used + 1], h->st);
        h->gemm_ev_used += 2;
        h->gemm_flops += 2.0 * 128 * 128 * 128 * (double)steps;
        h->gemm_launches++;
    }
}

// fold the recorded per-launch event pairs into gemm_ms (the stream must be idle)
void gemm_collect(gpx_t* h) {
    for (size_t i = 0; i + 1 < h->gemm_ev_used; i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->gemm_ev[i], h->gemm_ev[i + 1]) == cudaSuccess) h->gemm_ms += ms;
        else cudaGetLastError();
    }
    h->gemm_ev_used = 0;
}

TileMat lmat(gpx_t* h) { return TileMat{ptr<double>(h->L), ptr<int64_t>(h->coloff), 1}; }
TileMat linvmat(gpx_t* h) { return TileMat{ptr<double>(h->linv), ptr<int64_t>(h->linv_coloff), 1}; }

// ---- blocked right-looking Cholesky of the packed tile storage (in place) -------------------------------------
int cholesky(gpx_t* h) {
    const int nt = h->nt;
    TileMat L = lmat(h), Li = linvmat(h);
    int* info = ptr<int>(h->info);
    const int W = h->outer_tiles < 1 ? 1 : h->outer_tiles;
    for (int p0 = 0; p0 < nt; p0 += W) {
        const int p1 = (p0 + W < nt) ? p0 + W : nt;
        for (int k = p0; k < p1; k++) {
            double* diag = ptr<double>(h->L) + h->h_coloff[k];
            gpx_launch_potrf_tile(diag, ptr<double>(h->linv) + (int64_t)k * GPX_TILE_ELEMS, info, k, h->st);
            h->launches++;
            if (k + 1 < nt) {
                // TRSM: L(I,k) <- L(I,k) * Linv_k^T for I > k
                GemmJob t{};
                t.A = L; t.B = Li; t.C = L;
                t.tri = 0; t.i0 = k + 1; t.i1 = nt; t.j0 = k; t.j1 = k + 1; t.k0 = k; t.k1 = k + 1;
                t.bdiag = 1; t.overwrite = 1;
                gemm(h, t);
            }
            if (k + 1 < p1) {
                // update the remaining columns of this outer panel with column k
                GemmJob u{};
                u.A = L; u.B = L; u.C = L;
                u.tri = 1; u.i0 = k + 1; u.i1 = nt; u.j0 = k + 1; u.j1 = p1; u.k0 = k; u.k1 = k + 1;
                u.bdiag = 0; u.overwrite = 0;
                gemm(h, u);
            }
        }
        if (p1 < nt) {
            GemmJob u{};
            u.A = L; u.B = L; u.C = L;
            u.tri = 1; u.i0 = p1; u.i1 = nt; u.j0 = p1; u.j1 = nt; u.k0 = p0; u.k1 = p1;
            u.bdiag = 0; u.overwrite = 0;
            gemm(h, u);
        }
    }
    return 0;
}

}  // namespace

extern "C" {

int gpx_version(void) { return 100; }

int gpx_create(gpx_t** out, int device) {
    if (!out) return -1;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return -4; }
    if (device < 0 || device >= ndev) return -1;
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    gpx_t* h = new gpx_handle();
    h->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete h; return -2; }
    if (prop.major != 10) {
        // sm_100a-only library: refuse loudly rather than fall back
        delete h;
        return -5;
    }
    h->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking) != cudaSuccess) { delete h; return -2; }
    for (auto& e : h->ev) cudaEventCreate(&e);
    for (auto& m : h->phase_ms) m = 0.f;
    gpx_gemm_init();
    gpx_potrf_init();
    gpx_kbuild_init();
    *out = h;
    return 0;
}

void gpx_destroy(gpx_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->st);
    DevBuf* all[] = {&h->Xs, &h->xsq, &h->ls, &h->L, &h->linv, &h->coloff, &h->linv_coloff, &h->resid, &h->work1,
                     &h->work2, &h->alpha, &h->scal, &h->info, &h->stage, &h->Tt, &h->tt_coloff, &h->Xt, &h->xtsq,
                     &h->mean_partial, &h->dmean, &h->dvar};
    for (DevBuf* b : all) release(h, *b);
    if (h->pinned) cudaFreeHost(h->pinned);
    for (auto& e : h->ev) cudaEventDestroy(e);
    for (auto& e : h->gemm_ev) cudaEventDestroy(e);
    cudaStreamDestroy(h->st);
    delete h;
}

const char* gpx_last_error(const gpx_t* h) { return h ? h->err.c_str() : "null handle"; }

int gpx_comm_unique_id(char* out128) {
    (void)out128;
    return -6;  // multi-GPU path: see gpx_dist.cu (not linked in this build)
}
int gpx_comm_init(gpx_t* h, int rank, int world, const char* uid128) {
    (void)uid128;
    if (world == 1 && rank == 0) return 0;
    return fail(h, -6, "multi-GPU communicator not available in this build");
}

int gpx_set_option(gpx_t* h, const char* name, int64_t value) {
    if (!h || !name) return -1;
    if (!strcmp(name, "outer_tiles")) { if (value < 1 || value > 64) return fail(h, -1, "outer_tiles out of range"); h->outer_tiles = (int)value; return 0; }
    if (!strcmp(name, "predict_batch_tiles")) { if (value < 0) return fail(h, -1, "bad predict_batch_tiles"); h->predict_batch_tiles = (int)value; return 0; }
    if (!strcmp(name, "sync_phases")) { h->sync_phases = value ? 1 : 0; return 0; }
    if (!strcmp(name, "gemm_timing")) { h->gemm_timing = value ? 1 : 0; return 0; }
    return fail(h, -1, "unknown option %s", name);
}

double gpx_phase_ms(const gpx_t* h, int phase) {
    if (!h || phase < 0 || phase >= GPX_NUM_PHASES) return -1.0;
    return h->phase_ms[phase];
}
int64_t gpx_launch_count(const gpx_t* h) { return h ? h->launches : -1; }
void* gpx_stream(const gpx_t* h) { return h ? (void*)h->st : nullptr; }
int gpx_gemm_stats(gpx_t* h, double* total_ms, double* issued_flops, int64_t* launches) {
    if (!h) return -1;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->st);
    gemm_collect(h);
    if (total_ms) *total_ms = h->gemm_ms;
    if (issued_flops) *issued_flops = h->gemm_flops;
    if (launches) *launches = h->gemm_launches;
    h->gemm_ms = 0; h->gemm_flops = 0; h->gemm_launches = 0;
    return 0;
}
int64_t gpx_bytes_allocated(const gpx_t* h) { return h ? h->bytes_alloc : -1; }

int gpx_fit(gpx_t* h, const double* X, int64_t N, int d, const double* Y, int O, int kernel_id, double variance,
            const double* lengthscale, int n_ls, double noise, double jitter, double mean_const, int mem) {
    if (!h) return -1;
    if (!X || !Y || !lengthscale) return fail(h, -1, "null argument");
    if (N < 1 || d < 1 || O < 1) return fail(h, -1, "bad shape N=%lld d=%d O=%d", (long long)N, d, O);
    if (d > gpx_kbuild_max_d()) return fail(h, -1, "input dimension %d exceeds the supported maximum %d", d, gpx_kbuild_max_d());
    if (kernel_id < 0 || kernel_id > 3) return fail(h, -1, "unknown kernel id %d", kernel_id);
    if (n_ls != 1 && n_ls != d) return fail(h, -1, "lengthscale must have 1 or d entries");
    if (!(variance > 0) || !(noise + jitter > 0)) return fail(h, -1, "variance and noise+jitter must be positive");
    for (int i = 0; i < n_ls; i++) if (!(lengthscale[i] > 0)) return fail(h, -1, "lengthscale must be positive");
    GPX_CUDA(h, cudaSetDevice(h->device));
    h->fitted = false;
    const int nt = (int)((N + GPX_TILE - 1) / GPX_TILE);
    const int64_t Np = (int64_t)nt * GPX_TILE;
    h->N = N; h->Np = Np; h->nt = nt; h->d = d; h->O = O; h->kernel_id = kernel_id; h->n_ls = n_ls;
    h->variance = variance; h->noise = noise; h->jitter = jitter; h->mean_const = mean_const;
    for (auto& m : h->phase_ms) m = 0.f;

    // packed lower-triangular tile storage
    const int64_t ntiles = (int64_t)nt * (nt + 1) / 2;
    int rc;
    if ((rc = ensure(h, h->L, (size_t)ntiles * GPX_TILE_BYTES))) return rc;
    if ((rc = ensure(h, h->linv, (size_t)nt * GPX_TILE_BYTES))) return rc;
    if ((rc = ensure(h, h->coloff, (size_t)nt * 8))) return rc;
    if ((rc = ensure(h, h->linv_coloff, (size_t)nt * 8))) return rc;
    if ((rc = ensure(h, h->Xs, (size_t)Np * d * 8))) return rc;
    if ((rc = ensure(h, h->xsq, (size_t)Np * 8))) return rc;
    if ((rc = ensure(h, h->ls, (size_t)n_ls * 8))) return rc;
    if ((rc = ensure(h, h->resid, (size_t)Np * O * 8))) return rc;
    if ((rc = ensure(h, h->work1, (size_t)Np * O * 8))) return rc;
    if ((rc = ensure(h, h->work2, (size_t)Np * O * 8))) return rc;
    if ((rc = ensure(h, h->alpha, (size_t)Np * O * 8))) return rc;
    if ((rc = ensure(h, h->scal, 64))) return rc;
    if ((rc = ensure(h, h->info, 64))) return rc;
    h->h_coloff.resize(nt);
    std::vector<int64_t> lc(nt);
    int64_t off = 0;
    for (int c = 0; c < nt; c++) { h->h_coloff[c] = off; off += (int64_t)(nt - c) * GPX_TILE_ELEMS; lc[c] = (int64_t)c * GPX_TILE_ELEMS; }
    GPX_CUDA(h, cudaMemcpyAsync(h->coloff.p, h->h_coloff.data(), (size_t)nt * 8, cudaMemcpyHostToDevice, h->st));
    GPX_CUDA(h, cudaMemcpyAsync(h->linv_coloff.p, lc.data(), (size_t)nt * 8, cudaMemcpyHostToDevice, h->st));
    GPX_CUDA(h, cudaMemcpyAsync(h->ls.p, lengthscale, (size_t)n_ls * 8, cudaMemcpyHostToDevice, h->st));
    GPX_CUDA(h, cudaMemsetAsync(h->info.p, 0, 64, h->st));
    GPX_CUDA(h, cudaStreamSynchronize(h->st));  // lc / lengthscale are host temporaries

    // inputs
    phase_begin(h, GPX_PHASE_H2D);
    const double *dX, *dY;
    DevBuf stageY;
    if ((rc = to_device(h, X, (size_t)N * d, mem, h->stage, &dX))) return rc;
    if ((rc = to_device(h, Y, (size_t)N * O, mem, stageY, &dY))) { release(h, stageY); return rc; }
    phase_end(h, GPX_PHASE_H2D);

    // K-build
    phase_begin(h, GPX_PHASE_KBUILD);
    gpx_launch_scale_x(dX, N, Np, d, ptr<double>(h->ls), n_ls, ptr<double>(h->Xs), ptr<double>(h->xsq), h->st);
    gpx_launch_resid(dY, N, Np, O, mean_const, ptr<double>(h->resid), h->st);
    gpx_launch_kbuild_lower(lmat(h), nt, 0, 1, nt, ptr<double>(h->Xs), ptr<double>(h->xsq), N, d, kernel_id, variance,
                            noise + jitter, h->st);
    h->launches += 3;
    phase_end(h, GPX_PHASE_KBUILD);

    // Cholesky
    phase_begin(h, GPX_PHASE_CHOLESKY);
    cholesky(h);
    phase_end(h, GPX_PHASE_CHOLESKY);

    // alpha = Ky^-1 resid
    phase_begin(h, GPX_PHASE_SOLVE);
    GPX_CUDA(h, cudaMemcpyAsync(h->work1.p, h->resid.p, (size_t)Np * O * 8, cudaMemcpyDeviceToDevice, h->st));
    const int max_ctas = 2 * h->num_sms;
    for (int J = 0; J < nt; J++) {
        gpx_launch_fwd_step(lmat(h), ptr<double>(h->linv), nt, J, ptr<double>(h->work1), ptr<double>(h->work2), Np, O,
                            max_ctas, h->st);
        h->launches++;
    }
    for (int I = nt - 1; I >= 0; I--) {
        gpx_launch_bwd_step(lmat(h), ptr<double>(h->linv), I, ptr<double>(h->work2), ptr<double>(h->alpha), Np, O,
                            max_ctas, h->st);
        h->launches++;
    }
    phase_end(h, GPX_PHASE_SOLVE);

    // log-determinant and quadratic form
    phase_begin(h, GPX_PHASE_LML);
    gpx_launch_lml(lmat(h), nt, nullptr, ptr<double>(h->alpha), ptr<double>(h->resid), Np, O, ptr<double>(h->scal), h->st);
    h->launches++;
    phase_end(h, GPX_PHASE_LML);

    double hs[2];
    int hinfo = 0;
    GPX_CUDA(h, cudaMemcpyAsync(hs, h->scal.p, 16, cudaMemcpyDeviceToHost, h->st));
    GPX_CUDA(h, cudaMemcpyAsync(&hinfo, h->info.p, 4, cudaMemcpyDeviceToHost, h->st));
    cudaError_t e = cudaStreamSynchronize(h->st);
    release(h, stageY);
    if (e != cudaSuccess) return fail(h, -2, "CUDA error during fit: %s", cudaGetErrorString(e));
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, -2, "CUDA launch error during fit: %s", cudaGetErrorString(e));
    for (int ph : {GPX_PHASE_H2D, GPX_PHASE_KBUILD, GPX_PHASE_CHOLESKY, GPX_PHASE_SOLVE, GPX_PHASE_LML}) phase_collect(h, ph);
    gemm_collect(h);
    if (hinfo != 0) {
        fail(h, hinfo, "Ky is not positive definite: non-positive pivot at row %d", hinfo);
        return hinfo;
    }
    h->logdet = hs[0];
    h->quad = hs[1];
    const double log2pi = 1.8378770664093454835606594728112;
    h->lml = 0.5 * (-(double)N * O * log2pi - (double)O * h->logdet - h->quad);
    h->fitted = true;
    return 0;
}

int gpx_lml(gpx_t* h, double* lml, double* logdet, double* quad) {
    if (!h) return -1;
    if (!h->fitted) return fail(h, -1, "gpx_lml called before a successful gpx_fit");
    if (lml) *lml = h->lml;
    if (logdet) *logdet = h->logdet;
    if (quad) *quad = h->quad;
    return 0;
}

// variance solve for one batch of nbt test tiles held in Tt (rows = test points): Tt <- Tt L^-T, blocked.
static void var_trsm(gpx_t* h, TileMat Tt, int nbt) {
    const int nt = h->nt;
    TileMat L = lmat(h), Li = linvmat(h);
    const int W = h->outer_tiles < 1 ? 1 : h->outer_tiles;
    for (int p0 = 0; p0 < nt; p0 += W) {
        const int p1 = (p0 + W < nt) ? p0 + W : nt;
        for (int k = p0; k < p1; k++) {
            GemmJob t{};
            t.A = Tt; t.B = Li; t.C = Tt;
            t.tri = 0; t.i0 = 0; t.i1 = nbt; t.j0 = k; t.j1 = k + 1; t.k0 = k; t.k1 = k + 1; t.bdiag = 1; t.overwrite = 1;
            gemm(h, t);
            if (k + 1 < p1) {
                GemmJob u{};
                u.A = Tt; u.B = L; u.C = Tt;
                u.tri = 0; u.i0 = 0; u.i1 = nbt; u.j0 = k + 1; u.j1 = p1; u.k0 = k; u.k1 = k + 1; u.bdiag = 0; u.overwrite = 0;
                gemm(h, u);
            }
        }
        if (p1 < nt) {
            GemmJob u{};
            u.A = Tt; u.B = L; u.C = Tt;
            u.tri = 0; u.i0 = 0; u.i1 = nbt; u.j0 = p1; u.j1 = nt; u.k0 = p0; u.k1 = p1; u.bdiag = 0; u.overwrite = 0;
            gemm(h, u);
        }
    }
}

int gpx_predict(gpx_t* h, const double* Xs, int64_t Ns, double* mean, double* var, int include_noise, int mem) {
    if (!h) return -1;
    if (!h->fitted) return fail(h, -1, "gpx_predict called before a successful gpx_fit");
    if (!Xs || !mean) return fail(h, -1, "null argument");
    if (Ns < 1) return fail(h, -1, "bad Ns");
    GPX_CUDA(h, cudaSetDevice(h->device));
    const int nt = h->nt, d = h->d, O = h->O;
    const int64_t Np = h->Np;
    for (int ph : {GPX_PHASE_PREDICT_KSTAR, GPX_PHASE_PREDICT_TRSM, GPX_PHASE_PREDICT_REDUCE, GPX_PHASE_H2D, GPX_PHASE_D2H}) h->phase_ms[ph] = 0.f;

    const int total_tiles = (int)((Ns + GPX_TILE - 1) / GPX_TILE);
    // batch size: a multiple of the SM count when possible, bounded by free memory when the variance is wanted
    int cap = h->predict_batch_tiles > 0 ? h->predict_batch_tiles : h->num_sms;
    if (var) {
        size_t fr = 0, tot = 0;
        GPX_CUDA(h, cudaMemGetInfo(&fr, &tot));
        fr += h->Tt.bytes;  // reusable
        size_t per_tile_row = (size_t)nt * GPX_TILE_BYTES;
        size_t can = (size_t)(0.8 * (double)fr) / per_tile_row;
        if (can < 1) return fail(h, -3, "not enough device memory for the variance workspace");
        if ((size_t)cap > can) cap = (int)can;
    }
    if (cap > total_tiles) cap = total_tiles;
    const int nbatch = (total_tiles + cap - 1) / cap;
    const int per = (total_tiles + nbatch - 1) / nbatch;

    int rc;
    phase_begin(h, GPX_PHASE_H2D);
    const double* dXs;
    if ((rc = to_device(h, Xs, (size_t)Ns * d, mem, h->stage, &dXs))) return rc;
    phase_end(h, GPX_PHASE_H2D);
    double* dmean = mean;
    double* dvar = var;
    if (mem == GPX_MEM_HOST) {
        if ((rc = ensure(h, h->dmean, (size_t)Ns * O * 8))) return rc;
        dmean = ptr<double>(h->dmean);
        if (var) { if ((rc = ensure(h, h->dvar, (size_t)Ns * 8))) return rc; dvar = ptr<double>(h->dvar); }
    }
    if ((rc = ensure(h, h->Xt, (size_t)per * GPX_TILE * d * 8))) return rc;
    if ((rc = ensure(h, h->xtsq, (size_t)per * GPX_TILE * 8))) return rc;
    if ((rc = ensure(h, h->mean_partial, (size_t)nt * per * GPX_TILE * O * 8))) return rc;
    TileMat Tt{nullptr, nullptr, 0};
    if (var) {
        if ((rc = ensure(h, h->Tt, (size_t)per * nt * GPX_TILE_BYTES))) return rc;
        if ((rc = ensure(h, h->tt_coloff, (size_t)nt * 8))) return rc;
        std::vector<int64_t> tc(nt);
        for (int c = 0; c < nt; c++) tc[c] = (int64_t)c * per * GPX_TILE_ELEMS;
        GPX_CUDA(h, cudaMemcpyAsync(h->tt_coloff.p, tc.data(), (size_t)nt * 8, cudaMemcpyHostToDevice, h->st));
        GPX_CUDA(h, cudaStreamSynchronize(h->st));
        Tt = TileMat{ptr<double>(h->Tt), ptr<int64_t>(h->tt_coloff), 0};
    }
    const double prior_var = h->variance + (include_noise ? h->noise : 0.0);

    // per-phase events only bracket the first batch exactly; accumulate by timing whole loops per phase instead:
    // phases are interleaved per batch, so record begin at the first batch and end at the last for each phase is
    // wrong -> time each batch's phases and accumulate on the host after a sync per batch (cheap: few batches).
    for (int b = 0; b < nbatch; b++) {
        const int t0 = b * per;
        const int nb = (t0 + per <= total_tiles) ? per : (total_tiles - t0);
        if (nb <= 0) break;
        const int64_t row0 = (int64_t)t0 * GPX_TILE;
        const int64_t valid = (Ns - row0 < (int64_t)nb * GPX_TILE) ? (Ns - row0) : (int64_t)nb * GPX_TILE;
        phase_begin(h, GPX_PHASE_PREDICT_KSTAR);
        gpx_launch_scale_x(dXs + row0 * d, valid, (int64_t)nb * GPX_TILE, d, ptr<double>(h->ls), h->n_ls, ptr<double>(h->Xt),
                           ptr<double>(h->xtsq), h->st);
        gpx_launch_kcross(Tt, nb, per, nt, ptr<double>(h->Xt), ptr<double>(h->xtsq), valid, ptr<double>(h->Xs),
                          ptr<double>(h->xsq), h->N, d, h->kernel_id, h->variance, ptr<double>(h->alpha), O, Np,
                          ptr<double>(h->mean_partial), h->st);
        gpx_launch_mean_reduce(ptr<double>(h->mean_partial), per, nt, O, h->mean_const, dmean, row0, Ns, h->st);
        h->launches += 3;
        phase_end(h, GPX_PHASE_PREDICT_KSTAR);
        if (var) {
            phase_begin(h, GPX_PHASE_PREDICT_TRSM);
            var_trsm(h, Tt, nb);
            phase_end(h, GPX_PHASE_PREDICT_TRSM);
            phase_begin(h, GPX_PHASE_PREDICT_REDUCE);
            gpx_launch_var_reduce(Tt, nb, nt, prior_var, dvar, row0, Ns, h->st);
            h->launches++;
            phase_end(h, GPX_PHASE_PREDICT_REDUCE);
        }
        if (h->sync_phases) {
            GPX_CUDA(h, cudaStreamSynchronize(h->st));
            for (int ph : {GPX_PHASE_PREDICT_KSTAR, GPX_PHASE_PREDICT_TRSM, GPX_PHASE_PREDICT_REDUCE}) {
                if (!var && ph != GPX_PHASE_PREDICT_KSTAR) continue;
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, h->ev[2 * ph], h->ev[2 * ph + 1]) == cudaSuccess) h->phase_ms[ph] += ms;
                else cudaGetLastError();
            }
        }
    }
    if (mem == GPX_MEM_HOST) {
        phase_begin(h, GPX_PHASE_D2H);
        GPX_CUDA(h, cudaMemcpyAsync(mean, dmean, (size_t)Ns * O * 8, cudaMemcpyDeviceToHost, h->st));
        if (var) GPX_CUDA(h, cudaMemcpyAsync(var, dvar, (size_t)Ns * 8, cudaMemcpyDeviceToHost, h->st));
        phase_end(h, GPX_PHASE_D2H);
    }
    cudaError_t e = cudaStreamSynchronize(h->st);
    if (e != cudaSuccess) return fail(h, -2, "CUDA error during predict: %s", cudaGetErrorString(e));
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, -2, "CUDA launch error during predict: %s", cudaGetErrorString(e));
    phase_collect(h, GPX_PHASE_H2D);
    if (mem == GPX_MEM_HOST) phase_collect(h, GPX_PHASE_D2H);
    gemm_collect(h);
    return 0;
}

int gpx_predict_cov(gpx_t* h, const double* Xs, int64_t Ns, double* mean, double* cov, int include_noise, int mem) {
    (void)Xs; (void)Ns; (void)mean; (void)cov; (void)include_noise; (void)mem;
    return fail(h, -6, "gpx_predict_cov: not implemented yet");
}

int gpx_export(gpx_t* h, double* alpha, double* Lout) {
    if (!h) return -1;
    if (!h->fitted) return fail(h, -1, "gpx_export called before a successful gpx_fit");
    GPX_CUDA(h, cudaSetDevice(h->device));
    int rc;
    if (alpha) {
        DevBuf tmp;
        if ((rc = ensure(h, tmp, (size_t)h->N * h->O * 8))) return rc;
        gpx_launch_untranspose(ptr<double>(h->alpha), h->N, h->Np, h->O, ptr<double>(tmp), h->st);
        cudaError_t e = cudaMemcpyAsync(alpha, tmp.p, (size_t)h->N * h->O * 8, cudaMemcpyDeviceToHost, h->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
        release(h, tmp);
        if (e != cudaSuccess) return fail(h, -2, "CUDA error in gpx_export: %s", cudaGetErrorString(e));
    }
    if (Lout) {
        DevBuf tmp;
        if ((rc = ensure(h, tmp, (size_t)h->N * h->N * 8))) return rc;
        gpx_launch_detile(lmat(h), 1, h->N, h->N, ptr<double>(tmp), h->st);
        cudaError_t e = cudaMemcpyAsync(Lout, tmp.p, (size_t)h->N * h->N * 8, cudaMemcpyDeviceToHost, h->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
        release(h, tmp);
        if (e != cudaSuccess) return fail(h, -2, "CUDA error in gpx_export: %s", cudaGetErrorString(e));
    }
    return 0;
}

int gpx_kernel_matrix(gpx_t* h, const double* X, int64_t N, int d, const double* X2, int64_t M, int kernel_id,
                      double variance, const double* lengthscale, int n_ls, double diag_add, double* K_out) {
    if (!h) return -1;
    if (!X || !lengthscale || !K_out) return fail(h, -1, "null argument");
    if (N < 1 || d < 1 || d > gpx_kbuild_max_d() || (n_ls != 1 && n_ls != d)) return fail(h, -1, "bad shape");
    GPX_CUDA(h, cudaSetDevice(h->device));
    const bool sym = (X2 == nullptr);
    if (sym) M = N;
    const int ntr = (int)((N + GPX_TILE - 1) / GPX_TILE), ntc = (int)((M + GPX_TILE - 1) / GPX_TILE);
    DevBuf dX, dXs, dsq, dX2, dX2s, dsq2, dls, tiles, off, dense;
    int rc = 0;
    auto cleanup = [&]() { for (DevBuf* b : {&dX, &dXs, &dsq, &dX2, &dX2s, &dsq2, &dls, &tiles, &off, &dense}) release(h, *b); };
    std::vector<int64_t> co(sym ? ntr : ntc);
    int64_t ntiles;
    if (sym) { int64_t o = 0; for (int c = 0; c < ntr; c++) { co[c] = o; o += (int64_t)(ntr - c) * GPX_TILE_ELEMS; } ntiles = o / GPX_TILE_ELEMS; }
    else { for (int c = 0; c < ntc; c++) co[c] = (int64_t)c * ntr * GPX_TILE_ELEMS; ntiles = (int64_t)ntr * ntc; }
    if ((rc = ensure(h, dX, (size_t)N * d * 8)) || (rc = ensure(h, dXs, (size_t)ntr * GPX_TILE * d * 8)) ||
        (rc = ensure(h, dsq, (size_t)ntr * GPX_TILE * 8)) || (rc = ensure(h, dls, (size_t)n_ls * 8)) ||
        (rc = ensure(h, tiles, (size_t)ntiles * GPX_TILE_BYTES)) || (rc = ensure(h, off, co.size() * 8)) ||
        (rc = ensure(h, dense, (size_t)N * M * 8))) { cleanup(); return rc; }
    cudaMemcpyAsync(dX.p, X, (size_t)N * d * 8, cudaMemcpyHostToDevice, h->st);
    cudaMemcpyAsync(dls.p, lengthscale, (size_t)n_ls * 8, cudaMemcpyHostToDevice, h->st);
    cudaMemcpyAsync(off.p, co.data(), co.size() * 8, cudaMemcpyHostToDevice, h->st);
    gpx_launch_scale_x(ptr<double>(dX), N, (int64_t)ntr * GPX_TILE, d, ptr<double>(dls), n_ls, ptr<double>(dXs), ptr<double>(dsq), h->st);
    if (sym) {
        TileMat Lm{ptr<double>(tiles), ptr<int64_t>(off), 1};
        gpx_launch_kbuild_lower(Lm, ntr, 0, 1, ntr, ptr<double>(dXs), ptr<double>(dsq), N, d, kernel_id, variance, diag_add, h->st);
        gpx_launch_detile(Lm, 1, N, N, ptr<double>(dense), h->st);
    } else {
        if ((rc = ensure(h, dX2, (size_t)M * d * 8)) || (rc = ensure(h, dX2s, (size_t)ntc * GPX_TILE * d * 8)) ||
            (rc = ensure(h, dsq2, (size_t)ntc * GPX_TILE * 8))) { cleanup(); return rc; }
        cudaMemcpyAsync(dX2.p, X2, (size_t)M * d * 8, cudaMemcpyHostToDevice, h->st);
        gpx_launch_scale_x(ptr<double>(dX2), M, (int64_t)ntc * GPX_TILE, d, ptr<double>(dls), n_ls, ptr<double>(dX2s), ptr<double>(dsq2), h->st);
        TileMat Tm{ptr<double>(tiles), ptr<int64_t>(off), 0};
        // rows = X (as "test"), cols = X2 (as "train")
        gpx_launch_kcross(Tm, ntr, ntr, ntc, ptr<double>(dXs), ptr<double>(dsq), N, ptr<double>(dX2s), ptr<double>(dsq2), M, d,
                          kernel_id, variance, nullptr, 0, 0, nullptr, h->st);
        gpx_launch_detile(Tm, 0, N, M, ptr<double>(dense), h->st);
    }
    cudaError_t e = cudaMemcpyAsync(K_out, dense.p, (size_t)N * M * 8, cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    if (e != cudaSuccess) return fail(h, -2, "CUDA error in gpx_kernel_matrix: %s", cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
