// The parts of the C ABI around the core fit / predict (gpx_api.cu):
//   * gpx_set_normalizer / gpx_get_normalizer / gpx_set_stream        (boundary plumbing)
//   * gpx_predict_small  -- latency path of gpx_predict for <= 128 test rows
//   * gpx_append         -- rank-k append to the resident factor        
//   * gpx_selfcheck      -- residual / factor checks on the device      (SURVEY.md 8(d) "Parity check", C5 clause)
//   * gpx_predict_moments-- sums behind predict_jacobian / predict_hessian
#include "gpx_internal.cuh"

namespace {

// resid[o][row0 + i] = (i < k) ? Y[i][o] - mean_const : 0   for i in [0, rows_padded)
__global__ void resid_append_kernel(const double* __restrict__ Y, int64_t k, int O, double mean_const,
                                    double* __restrict__ resid, int64_t Np, int64_t row0, int64_t rows_padded) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= rows_padded * O) return;
    int64_t o = e / rows_padded, i = e - o * rows_padded;
    resid[o * Np + row0 + i] = (i < k) ? Y[i * O + o] - mean_const : 0.0;
}

// var[s][*] = prior - sum_i z[s][i]^2 : one CTA per test row, fixed-order tree (deterministic)
__global__ void __launch_bounds__(256) vec_var_kernel(const double* __restrict__ z, int64_t n_rows, int64_t Np, double prior,
                                                      const double* __restrict__ ystats, int var_cols,
                                                      double* __restrict__ var) {
    __shared__ double red[256];
    const int s = blockIdx.x, tid = threadIdx.x;
    const double* zs = z + (int64_t)s * Np;
    double a = 0.0;
    for (int64_t i = tid; i < n_rows; i += 256) a = fma(zs[i], zs[i], a);
    red[tid] = a;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) { if (tid < off) red[tid] += red[tid + off]; __syncthreads(); }
    if (tid < var_cols) {
        const double v = prior - red[0];
        const double sd = ystats ? ystats[var_cols + tid] : 1.0;
        var[(int64_t)s * var_cols + tid] = ystats ? v * (sd * sd) : v;
    }
}

// Incremental forward substitution of gpx_append: partial[(i * S + s)][o][r] = sum over the tile columns J of split s
// (J < jend) of (L(i0 + i, J) z_o[J])[r]  -- the new tile rows against the z the fit left behind.  grid (rows, S).
// upper = 1: the tile columns of row i0 + i start at the diagonal, J in [i0 + i, jend) -- the rows of the packed-upper
// explicit inverse Y = L^-T times a vector.
constexpr int kRowRhs = 8;
__global__ void __launch_bounds__(256) row_gemv_partial_kernel(TileMat L, int i0, int jend, int S, const double* __restrict__ z,
                                                               int64_t Np, int O, double* __restrict__ partial, int upper) {
    __shared__ double sz[kRowRhs][GPX_TILE], scratch[kRowRhs][GPX_TILE];
    const int i = blockIdx.x, sidx = blockIdx.y, tid = threadIdx.x, r = tid & 127, half = tid >> 7;
    const int jlo = upper ? i0 + i : 0, jn = jend - jlo;
    const int J0 = jlo + (int)((long long)jn * sidx / S), J1 = jlo + (int)((long long)jn * (sidx + 1) / S);
    double acc[kRowRhs];
#pragma unroll
    for (int o = 0; o < kRowRhs; o++) acc[o] = 0.0;
    for (int J = J0; J < J1; J++) {
        __syncthreads();
        for (int e = tid; e < O * GPX_TILE; e += 256) sz[e >> 7][e & 127] = z[(int64_t)(e >> 7) * Np + (int64_t)J * GPX_TILE + (e & 127)];
        __syncthreads();
        const double* tp = tile_ptr(L, i0 + i, J);
#pragma unroll
        for (int kc = 0; kc < 8; kc++) {
            const int ch = half * 8 + kc;
            const double2* p = reinterpret_cast<const double2*>(tp + ch * 1024 + r * 8);
            const double2 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
#pragma unroll
            for (int o = 0; o < kRowRhs; o++) {
                if (o < O) {
                    const double* x = &sz[o][ch * 8];
                    double a = acc[o];
                    a = fma(v0.x, x[0], a); a = fma(v0.y, x[1], a); a = fma(v1.x, x[2], a); a = fma(v1.y, x[3], a);
                    a = fma(v2.x, x[4], a); a = fma(v2.y, x[5], a); a = fma(v3.x, x[6], a); a = fma(v3.y, x[7], a);
                    acc[o] = a;
                }
            }
        }
    }
#pragma unroll
    for (int o = 0; o < kRowRhs; o++) if (o < O && half == 1) scratch[o][r] = acc[o];
    __syncthreads();
#pragma unroll
    for (int o = 0; o < kRowRhs; o++)
        if (o < O && half == 0) partial[(((int64_t)i * S + sidx) * O + o) * GPX_TILE + r] = acc[o] + scratch[o][r];
}

// b[o][(i0 + i) * 128 + r] -= sum_s partial[(i * S + s)][o][r]   (fixed order; assign = 1: b = + the sum)
__global__ void row_gemv_apply_kernel(const double* __restrict__ partial, int rows, int S, int O, int i0, int64_t Np,
                                      double* __restrict__ b, int assign) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)rows * O * GPX_TILE) return;
    const int r = (int)(e & 127), o = (int)((e >> 7) % O), i = (int)((e >> 7) / O);
    double s = 0.0;
    for (int k = 0; k < S; k++) s += partial[(((int64_t)i * S + k) * O + o) * GPX_TILE + r];
    double* dst = b + (int64_t)o * Np + (int64_t)(i0 + i) * GPX_TILE + r;
    *dst = assign ? s : *dst - s;
}

// ---- GEMV-style append of k <= 8 observations that fit into the partially filled last tile -----------------------------
// After the forward substitutions z_a = L^-1 k(X, x_new_a) (fwd_step, k right-hand sides) one CTA finishes the job:
//   S = K(new, new) + (noise + jitter) I - Z Z^T,  C = chol(S)                  (the k x k Schur complement)
//   new rows of L:      L[N0 + a][j] = z_a[j] (j < N0),  L[N0 + a][N0 + b] = C[a][b]
//   new rows of Linv_t: [-C^-1 (B A^-1), C^-1]  with B = the new rows' entries inside the last tile, A^-1 = its old inverse block
//   new entries of z_y: C^-1 (resid_new - Z z_y)                                 (z_y = L^-1 (Y - m), kept from the fit)
// The kernel entries follow the K-build's arithmetic (Gram trick, exact zero on the diagonal, sqrt, / lengthscale).
constexpr int kAppMax = 8;
struct AppendArgs {
    const double* z;        // [k][Np]
    const double* xt;       // [k][d] loaded new inputs
    const double* xtsq;     // [k]
    double* zy;             // [O][Np]  (zkeep)
    const double* resid;    // [O][Np]
    double* linv_tile;      // inverted diagonal tile of the last tile column
    double* cinv_out;       // null, or kAppMax x kAppMax: C^-1 (lower) for the update of the explicit inverse of the factor
    int* info;
    int64_t N0, Np;
    int k, d, O, t0, r0, kernel_id;
    double variance, rdiv, diag_add;
};

__device__ __forceinline__ double block_sum_1024(double v, double* red) {
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < 32; w++) s += red[w];   // every thread sums the 32 warp partials in the same order
    return s;
}

__global__ void __launch_bounds__(1024) append_finish_kernel(TileMat L, AppendArgs a) {
    __shared__ double red[32];
    __shared__ double S[kAppMax][kAppMax], C[kAppMax][kAppMax], Ci[kAppMax][kAppMax], D[kAppMax][kAppMax], T[kAppMax][GPX_TILE];
    __shared__ int bad;
    const int tid = threadIdx.x, k = a.k;
    if (tid == 0) bad = 0;
    // Schur complement and Z z_y
    for (int p = 0; p < k; p++)
        for (int q = 0; q <= p; q++) {
            double s = 0.0;
            for (int64_t j = tid; j < a.N0; j += 1024) s = fma(a.z[(int64_t)p * a.Np + j], a.z[(int64_t)q * a.Np + j], s);
            s = block_sum_1024(s, red);
            if (tid == 0) {
                double dot = 0.0;
                for (int c = 0; c < a.d; c++) dot = fma(a.xt[p * a.d + c], a.xt[q * a.d + c], dot);
                double r2 = fmax(-2.0 * dot + (a.xtsq[p] + a.xtsq[q]), 0.0);
                if (p == q) r2 = 0.0;
                double kv = gpx_k_of_r(a.kernel_id, a.variance, gpx_scaled_r(r2, a.rdiv));
                if (p == q) kv += a.diag_add;
                S[p][q] = kv - s;
            }
        }
    for (int p = 0; p < k; p++)
        for (int o = 0; o < a.O; o++) {
            double s = 0.0;
            for (int64_t j = tid; j < a.N0; j += 1024) s = fma(a.z[(int64_t)p * a.Np + j], a.zy[(int64_t)o * a.Np + j], s);
            s = block_sum_1024(s, red);
            if (tid == 0) D[p][o] = s;
        }
    __syncthreads();
    if (tid == 0) {
        for (int p = 0; p < k; p++)
            for (int q = 0; q <= p; q++) {
                double s = S[p][q];
                for (int m = 0; m < q; m++) s -= C[p][m] * C[q][m];
                if (p == q) {
                    if (!(s > 0.0)) { atomicCAS(a.info, 0, (int)(a.N0 + p + 1)); bad = 1; s = 1.0; }
                    C[p][p] = sqrt(s);
                } else {
                    C[p][q] = s / C[q][q];
                }
            }
        for (int p = 0; p < k; p++)       // inverse of the lower-triangular C
            for (int q = 0; q <= p; q++) {
                double s = (p == q) ? 1.0 : 0.0;
                for (int m = q; m < p; m++) s -= C[p][m] * Ci[m][q];
                Ci[p][q] = s / C[p][p];
            }
        if (a.cinv_out)
            for (int p = 0; p < kAppMax; p++)
                for (int q = 0; q < kAppMax; q++) a.cinv_out[p * kAppMax + q] = (p < k && q <= p) ? Ci[p][q] : 0.0;
        for (int p = 0; p < k; p++)       // z_y of the new rows
            for (int o = 0; o < a.O; o++) {
                double s = 0.0;
                for (int q = 0; q <= p; q++) s += Ci[p][q] * (a.resid[(int64_t)o * a.Np + a.N0 + q] - D[q][o]);
                a.zy[(int64_t)o * a.Np + a.N0 + p] = s;
            }
    }
    __syncthreads();
    if (bad) return;
    // new rows of L
    for (int p = 0; p < k; p++) {
        const int row = a.r0 + p;
        for (int64_t j = tid; j < a.N0; j += 1024)
            tile_ptr(L, a.t0, (int)(j >> 7))[gpx_tile_idx(row, (int)(j & 127))] = a.z[(int64_t)p * a.Np + j];
        if (tid < GPX_TILE - a.r0) {     // the tile columns of the new block and of the identity padding behind it
            const int q = tid;
            const double v = q < k ? (q <= p ? C[p][q] : 0.0) : 0.0;
            tile_ptr(L, a.t0, a.t0)[gpx_tile_idx(row, a.r0 + q)] = v;
        }
    }
    // new rows of the inverted diagonal tile: T = B A^-1 (A^-1 lower triangular), rows = [-C^-1 T, C^-1]
    if (tid < a.r0) {
        const int c = tid;
        for (int p = 0; p < k; p++) {
            double s = 0.0;
            for (int cc = c; cc < a.r0; cc++)
                s = fma(a.z[(int64_t)p * a.Np + (int64_t)a.t0 * GPX_TILE + cc], a.linv_tile[gpx_tile_idx(cc, c)], s);
            T[p][c] = s;
        }
    }
    __syncthreads();
    if (tid < GPX_TILE) {
        const int c = tid;
        for (int p = 0; p < k; p++) {
            double v = 0.0;
            if (c < a.r0) { for (int q = 0; q <= p; q++) v -= Ci[p][q] * T[q][c]; }
            else if (c - a.r0 < k) { const int q = c - a.r0; v = q <= p ? Ci[p][q] : 0.0; }
            a.linv_tile[gpx_tile_idx(a.r0 + p, c)] = v;
        }
    }
    gpx_publish_for_async_proxy();
}

// v[i] = fixed pseudo-random value in (-1, 1) for i < N, 0 for the padding (splitmix64 of the index)
__global__ void fill_random_kernel(double* __restrict__ v, int64_t N, int64_t Np) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Np) return;
    uint64_t x = (uint64_t)i + 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    x ^= x >> 31;
    v[i] = i < N ? ((double)(x >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0 : 0.0;
}

// y (+)= op(L) x over the owned tile columns of the packed lower factor (check code: atomics, not deterministic).
//   trans = 0:  y_I += L_IJ x_J ;  trans = 1:  y_J += L_IJ^T x_I.   grid (nt, ncols); column ordinal -> J as in GemmJob.
__global__ void __launch_bounds__(256) tri_matvec_kernel(TileMat L, int nt, int jbase, int jW, int jstride,
                                                         const double* __restrict__ x, double* __restrict__ y, int trans) {
    const int I = blockIdx.x, m = blockIdx.y;
    const int J = jbase + (m / jW) * jstride + (m % jW);
    if (J >= nt || I < J) return;
    __shared__ double sx[GPX_TILE];
    const int tid = threadIdx.x, q = tid & 127, half = tid >> 7;
    if (tid < GPX_TILE) sx[tid] = x[(int64_t)(trans ? I : J) * GPX_TILE + tid];
    __syncthreads();
    const double* tp = tile_ptr(L, I, J);
    double s = 0.0;
    if (!trans) {   // row q of the tile, columns of chunks half*8 .. half*8+7 (lower triangle only on the diagonal tile)
        for (int c = half * 64; c < half * 64 + 64; c++) {
            if (I == J && c > q) break;
            s = fma(tp[gpx_tile_idx(q, c)], sx[c], s);
        }
        atomicAdd(y + (int64_t)I * GPX_TILE + q, s);
    } else {        // column q of the tile, rows half*64 .. half*64+63
        for (int r = half * 64; r < half * 64 + 64; r++) {
            if (I == J && r < q) continue;
            s = fma(tp[gpx_tile_idx(r, q)], sx[r], s);
        }
        atomicAdd(y + (int64_t)J * GPX_TILE + q, s);
    }
}

// out[0] = sum (a + c * b - t)^2, out[1] = sum t^2   (a = K x, b = x, c = noise + jitter, t = target); one CTA
__global__ void __launch_bounds__(1024) resid_norm_kernel(const double* __restrict__ a, const double* __restrict__ b, double c,
                                                          const double* __restrict__ t, int64_t n, double* __restrict__ out) {
    __shared__ double red[2][32];
    double s0 = 0.0, s1 = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const double r = a[i] + c * b[i] - t[i];
        s0 = fma(r, r, s0);
        s1 = fma(t[i], t[i], s1);
    }
    for (int off = 16; off > 0; off >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, off); s1 += __shfl_xor_sync(0xffffffffu, s1, off); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double x0 = 0.0, x1 = 0.0;
        for (int w = 0; w < 32; w++) { x0 += red[0][w]; x1 += red[1][w]; }
        out[0] = x0; out[1] = x1;
    }
}

// t = a + c * b
__global__ void axpy_kernel(const double* __restrict__ a, const double* __restrict__ b, double c, double* __restrict__ t, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) t[i] = a[i] + c * b[i];
}

// Moments of the posterior mean (see gpx.h: gpx_predict_moments).  One CTA per test point, 256 training points per
// step: phase A evaluates w_j = k(x_j, x*) alpha_j (one per thread, the K-build's own Gram-trick arithmetic), phase B
// lets thread m accumulate moment m over the step's 256 points in index order (deterministic).
__global__ void __launch_bounds__(256) moments_kernel(const double* __restrict__ Xt, const double* __restrict__ xtsq,
                                                      const double* __restrict__ Xs, const double* __restrict__ xsq, int64_t N,
                                                      int d, int kernel_id, double variance, double rdiv,
                                                      const double* __restrict__ alpha, double* __restrict__ S0,
                                                      double* __restrict__ S1, double* __restrict__ S2) {
    extern __shared__ __align__(16) double sm[];
    double* w = sm;                 // [256]
    double* xi = w + 256;           // [d]
    double* xj = xi + d;            // [256][d]
    const int i = blockIdx.x, tid = threadIdx.x;
    for (int k = tid; k < d; k += 256) xi[k] = Xt[(int64_t)i * d + k];
    const double xisq = xtsq[i];
    const int M = 1 + (S1 ? d : 0) + (S2 ? d * d : 0);
    constexpr int kMaxPer = 5;      // moments per thread: 256 * 5 >= 1 + 32 + 32^2 ... capped by the host check d <= 32 -> 1057 < 1280
    double acc[kMaxPer];
#pragma unroll
    for (int u = 0; u < kMaxPer; u++) acc[u] = 0.0;
    __syncthreads();
    for (int64_t j0 = 0; j0 < N; j0 += 256) {
        const int64_t j = j0 + tid;
        double wj = 0.0;
        if (j < N) {
            double dot = 0.0;
            for (int k = 0; k < d; k++) { const double v = Xs[j * d + k]; xj[tid * d + k] = v; dot = fma(xi[k], v, dot); }
            const double r2 = fmax(-2.0 * dot + (xisq + xsq[j]), 0.0);
            wj = gpx_k_of_r(kernel_id, variance, gpx_scaled_r(r2, rdiv)) * alpha[j];
        }
        w[tid] = wj;
        __syncthreads();
        const int cnt = (N - j0) < 256 ? (int)(N - j0) : 256;
#pragma unroll
        for (int u = 0; u < kMaxPer; u++) {
            const int m = tid + u * 256;
            if (m >= M) break;
            double a = acc[u];
            if (m == 0) {
                for (int jj = 0; jj < cnt; jj++) a += w[jj];
            } else if (S1 && m <= d) {
                const int k = m - 1;
                for (int jj = 0; jj < cnt; jj++) a = fma(w[jj], xi[k] - xj[jj * d + k], a);
            } else {
                const int e = m - 1 - (S1 ? d : 0), k = e / d, l = e - k * d;
                for (int jj = 0; jj < cnt; jj++) a = fma(w[jj], (xi[k] - xj[jj * d + k]) * (xi[l] - xj[jj * d + l]), a);
            }
            acc[u] = a;
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < kMaxPer; u++) {
        const int m = tid + u * 256;
        if (m >= M) break;
        if (m == 0) S0[i] = acc[u];
        else if (S1 && m <= d) S1[(int64_t)i * d + (m - 1)] = acc[u];
        else S2[(int64_t)i * d * d + (m - 1 - (S1 ? d : 0))] = acc[u];
    }
}

constexpr int kMomentsMaxD = 32;
size_t moments_smem(int d) { return (size_t)(256 + d + 256 * d) * sizeof(double); }

// ---- host helpers --------------------------------------------------------------------------------------------------

struct SmallWs {   // views into h->sp (one allocation, re-made when cap_nt / d / O change)
    double *Xin, *Xt, *xtsq, *mean_partial, *b, *z, *mean, *var, *kpartial;
    unsigned int* ticket;
};
constexpr int kFinvAutoCalls = 16;
constexpr int kInvRows = 64;  // with the explicit inverse resident, up to 64 rows go through it in groups of 8 (128 rows: the tile route wins)
constexpr int kVecRows = 8;   // variance of <= 8 rows: GEMV-style substitution; above: one tile row through the DMMA solve

SmallWs small_views(gpx_t* h) {
    const int64_t d = h->d, O = h->O, Np = h->Np, cap = h->cap_nt;
    double* p = ptr<double>(h->sp);
    SmallWs w;
    w.Xin = p; p += GPX_TILE * d;
    w.Xt = p; p += GPX_TILE * d;
    w.xtsq = p; p += GPX_TILE;
    w.mean_partial = p; p += cap * GPX_TILE * O;
    w.b = p; p += kVecRows * Np;
    w.z = p; p += kVecRows * Np;
    w.mean = p; p += GPX_TILE * O;
    w.var = p; p += GPX_TILE * O;
    w.kpartial = p; p += cap * 64;            // per-tile partial sums of the fused small-batch K* kernel
    w.ticket = reinterpret_cast<unsigned int*>(p);
    return w;
}

int small_workspace(gpx_t* h) {
    const int64_t d = h->d, O = h->O, Np = h->Np, cap = h->cap_nt;
    const size_t doubles = (size_t)(2 * GPX_TILE * d + GPX_TILE + cap * GPX_TILE * O + 2 * kVecRows * Np + 2 * GPX_TILE * O + cap * 64 + 8);
    if (h->sp.p && h->sp_key[0] == cap && h->sp_key[1] == d && h->sp_key[2] == O) return 0;
    release(h, h->sp);
    int rc = ensure(h, h->sp, doubles * 8);
    if (rc) return rc;
    h->sp_key[0] = cap; h->sp_key[1] = d; h->sp_key[2] = O;
    GPX_CUDA(h, cudaMemsetAsync(small_views(h).ticket, 0, 64, h->st));
    h->gen++;   // captured graphs point into the old workspace
    const size_t pin_bytes = (size_t)(GPX_TILE * d + 2 * GPX_TILE * O) * 8;
    if (h->pin_bytes < pin_bytes) {
        if (h->pin) cudaFreeHost(h->pin);
        h->pin = nullptr; h->pin_bytes = 0;
        GPX_CUDA(h, cudaHostAlloc(&h->pin, pin_bytes, cudaHostAllocMapped));   // the latency path's kernels read / write it in place
        h->pin_bytes = pin_bytes;
        h->gen++;   // captured graphs point into the old buffer
    }
    return 0;
}


// ---- explicit inverse of the factor for the 1..8-row variance --------------------------------------------------------
// Y = L^-T is kept as a packed-upper tile matrix; v = L^-1 k* is then v_J = sum_{i <= J} Y(i, J)^T k*_i: no dependent chain,
// every tile is read exactly once, at HBM rate.  grid (nt, S): CTA (J, c) sums the tile rows i in [c C, min((c + 1) C, J + 1)).
// Thread map: warp w owns the column chunks w and w + 8 of the tile, lane = rsub * 4 + kp reads the double2 (r, 2 kp) of
// the rows r = 8 r8 + rsub -- a warp load is 512 contiguous bytes -- and the row sum is finished with three shuffles.
template <int RHS>
__global__ void __launch_bounds__(256) finv_gemv_kernel(TileMat Y, int C, int S, const double* __restrict__ b, int64_t Np,
                                                        double* __restrict__ partial) {
    const int J = blockIdx.x, c = blockIdx.y, i0 = c * C;
    if (i0 > J) return;
    const int i1 = (i0 + C < J + 1) ? i0 + C : J + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, rsub = lane >> 2, kp = lane & 3;
    double acc[2][RHS][2];
#pragma unroll
    for (int hh = 0; hh < 2; hh++)
#pragma unroll
        for (int q = 0; q < RHS; q++) acc[hh][q][0] = acc[hh][q][1] = 0.0;
    for (int i = i0; i < i1; i++) {
        const double* tp = tile_ptr(Y, i, J) + rsub * 8 + 2 * kp;
        const double* bi = b + (int64_t)i * GPX_TILE + rsub;
#pragma unroll 4
        for (int r8 = 0; r8 < 16; r8++) {
            const double2 t0 = __ldcs(reinterpret_cast<const double2*>(tp + warp * 1024 + r8 * 64));
            const double2 t1 = __ldcs(reinterpret_cast<const double2*>(tp + (warp + 8) * 1024 + r8 * 64));
#pragma unroll
            for (int q = 0; q < RHS; q++) {
                const double bv = __ldg(bi + (int64_t)q * Np + r8 * 8);
                acc[0][q][0] = fma(t0.x, bv, acc[0][q][0]);
                acc[0][q][1] = fma(t0.y, bv, acc[0][q][1]);
                acc[1][q][0] = fma(t1.x, bv, acc[1][q][0]);
                acc[1][q][1] = fma(t1.y, bv, acc[1][q][1]);
            }
        }
    }
    double* out = partial + ((int64_t)J * S + c) * RHS * GPX_TILE;
#pragma unroll
    for (int hh = 0; hh < 2; hh++)
#pragma unroll
        for (int q = 0; q < RHS; q++) {
            double x = acc[hh][q][0], y = acc[hh][q][1];
#pragma unroll
            for (int off = 4; off < 32; off <<= 1) {
                x += __shfl_xor_sync(0xffffffffu, x, off);
                y += __shfl_xor_sync(0xffffffffu, y, off);
            }
            if (rsub == 0) *reinterpret_cast<double2*>(out + q * GPX_TILE + (warp + 8 * hh) * 8 + 2 * kp) = make_double2(x, y);
        }
}

// var[s] = prior - sum_J | sum_c partial(J, c)[s] |^2.  grid (nt, rows): CTA (J, s) finishes the 128 entries of v_J, squares
// and sums them into colsq[s][J]; the CTA that arrives last (ticket counter, self-resetting) adds the nt column sums in a
// fixed order -- deterministic, and no single CTA walks the whole partial-sum buffer (that took 52 us of a 370 us call).
// (more than 4 right-hand sides are computed as two groups of 4: partial + group * gstride)
__global__ void __launch_bounds__(128) finv_var_kernel(const double* __restrict__ partial, int nt, int C, int S, int rhs,
                                                       int64_t gstride, double prior, const double* __restrict__ ystats,
                                                       int var_cols, double* __restrict__ colsq, int colsq_stride,
                                                       unsigned int* __restrict__ tickets, double* __restrict__ var) {
    __shared__ double red[4];
    __shared__ bool last;
    const int J = blockIdx.x, s = blockIdx.y, col = threadIdx.x, nc = J / C + 1;
    const double* p = partial + (int64_t)(s / rhs) * gstride + (((int64_t)J * S) * rhs + (s % rhs)) * GPX_TILE + col;
    double v = 0.0;
    for (int c = 0; c < nc; c++) v += p[(int64_t)c * rhs * GPX_TILE];
    double a = v * v;
    for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if ((col & 31) == 0) red[col >> 5] = a;
    __syncthreads();
    if (col == 0) {
        colsq[(int64_t)s * colsq_stride + J] = (red[0] + red[1]) + (red[2] + red[3]);
        __threadfence();
        last = atomicAdd(&tickets[s], 1u) == (unsigned)(nt - 1);
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (col < 32) {
        double t = 0.0;
        for (int j = col; j < nt; j += 32) t += __ldcg(&colsq[(int64_t)s * colsq_stride + j]);
        for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(0xffffffffu, t, off);
        if (col == 0) tickets[s] = 0u;
        if (col < var_cols) {
            const double x = prior - t;
            const double sd = ystats ? ystats[var_cols + col] : 1.0;
            var[(int64_t)s * var_cols + col] = ystats ? x * (sd * sd) : x;
        }
    }
}

// z[s][J * 128 + col] = sum_c partial(J, c)[s][col]: the vectors L^-1 b themselves (gpx_append).  grid (nt, nvec), 128 threads
__global__ void finv_collect_kernel(const double* __restrict__ partial, int C, int S, int rhs, int64_t gstride, int64_t Np,
                                    double* __restrict__ z) {
    const int J = blockIdx.x, s = blockIdx.y, col = threadIdx.x, nc = J / C + 1;
    const double* p = partial + (int64_t)(s / rhs) * gstride + (((int64_t)J * S) * rhs + (s % rhs)) * GPX_TILE + col;
    double v = 0.0;
    for (int c = 0; c < nc; c++) v += p[(int64_t)c * rhs * GPX_TILE];
    z[(int64_t)s * Np + (int64_t)J * GPX_TILE + col] = v;
}

// gpx_append of k <= 8 rows inside the last tile keeps the explicit inverse up to date: with L_new = [L 0; Z^T C],
//   Y_new = L_new^-T = [ Y  -U C^-T ; 0  C^-T ],  U = Y Z  (the rows of Y times the new rows of L, row_gemv kernels),
// i.e. k new columns inside the last tile column t0 (local columns r0 .. r0 + k - 1).  One thread per row j < N0 + k.
__global__ void finv_append_cols_kernel(TileMat Y, const double* __restrict__ U, const double* __restrict__ cinv, int64_t N0,
                                        int64_t Np, int t0, int r0, int k) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= N0 + k) return;
    double* tp = tile_ptr(Y, (int)(j >> 7), t0);
    const int r = (int)(j & 127);
    for (int p = 0; p < k; p++) {
        double v = 0.0;
        if (j < N0) { for (int q = 0; q <= p; q++) v = fma(-U[(int64_t)q * Np + j], cinv[p * kAppMax + q], v); }
        else { const int q = (int)(j - N0); v = q <= p ? cinv[p * kAppMax + q] : 0.0; }
        tp[gpx_tile_idx(r, r0 + p)] = v;
    }
}

// diagonal tiles of a zeroed tile matrix <- I
__global__ void diag_identity_kernel(TileMat T, int j0) {
    double* tp = tile_ptr(T, j0 + blockIdx.x, j0 + blockIdx.x);
    if (threadIdx.x < GPX_TILE) tp[gpx_tile_idx(threadIdx.x, threadIdx.x)] = 1.0;
    gpx_publish_for_async_proxy();
}

// behind the partial sums in finv_partial: per-row column sums of squares [kVecRows][cap_nt] and the ticket counters
inline size_t finv_partial_doubles(const gpx_t* h, int cap, int S) { return (size_t)cap * S * kVecRows * GPX_TILE; }
inline double* finv_colsq(gpx_t* h) { return ptr<double>(h->finv_partial) + finv_partial_doubles(h, h->finv_layout_cap, h->finv_S); }
inline unsigned int* finv_tickets(gpx_t* h) { return reinterpret_cast<unsigned int*>(finv_colsq(h) + (size_t)kVecRows * h->finv_layout_cap); }
inline TileMat finvmat(gpx_t* h) { return TileMat{ptr<double>(h->finv), ptr<int64_t>(h->finv_coloff), 0}; }
// right-hand sides per launch: 1, 2 or 4; 5..8 vectors run as two launches of 4 (the 8-wide instantiation needs 96 registers
// and measured 3.5x slower per tile than the 4-wide one: too few loads in flight)
inline int finv_rhs(int64_t nvec) { return nvec <= 1 ? 1 : nvec <= 2 ? 2 : 4; }
inline int64_t finv_gstride(const gpx_t* h, int rhs) { return (int64_t)h->nt * h->finv_S * rhs * GPX_TILE; }
// partial sums of Y^T b for nvec <= 8 vectors b[q][Np]; returns the number of launches
int launch_finv_gemv(gpx_t* h, int nvec, const double* b, cudaStream_t st) {
    const TileMat Y = finvmat(h);
    const dim3 grid(h->nt, h->nt / h->finv_C + 1);
    const int rhs = finv_rhs(nvec), groups = (nvec + rhs - 1) / rhs;
    for (int g = 0; g < groups; g++) {
        double* part = ptr<double>(h->finv_partial) + g * finv_gstride(h, rhs);
        const double* bg = b + (int64_t)g * rhs * h->Np;
        if (rhs == 1) finv_gemv_kernel<1><<<grid, 256, 0, st>>>(Y, h->finv_C, h->finv_S, bg, h->Np, part);
        else if (rhs == 2) finv_gemv_kernel<2><<<grid, 256, 0, st>>>(Y, h->finv_C, h->finv_S, bg, h->Np, part);
        else finv_gemv_kernel<4><<<grid, 256, 0, st>>>(Y, h->finv_C, h->finv_S, bg, h->Np, part);
    }
    return groups;
}
// out[o][*] = Y v_o for nrhs <= 8 vectors (the rows of the packed-upper inverse times a vector): two launches
int finv_row_apply(gpx_t* h, const double* v, int nrhs, double* out, cudaStream_t st) {
    const int nt = h->nt;
    int S2 = (8 * h->num_sms + nt - 1) / nt;
    if (S2 > nt) S2 = nt;
    if (S2 > h->finv_S) S2 = h->finv_S;    // the partial-sum buffer is sized for cap_nt * finv_S * 8 vectors
    row_gemv_partial_kernel<<<dim3(nt, S2), 256, 0, st>>>(finvmat(h), 0, nt, S2, v, h->Np, nrhs, ptr<double>(h->finv_partial), 1);
    row_gemv_apply_kernel<<<(int)(((int64_t)nt * nrhs * GPX_TILE + 255) / 256), 256, 0, st>>>(ptr<double>(h->finv_partial), nt, S2, nrhs, 0,
                                                                                           h->Np, out, 1);
    return 2;
}

// Build Y = L^-T with the blocked DMMA solve of the predictive variance applied to the identity (N^3/3 flop, once per fit
// generation; the same job list as gpx_lml_grad's first step).  Running out of memory is not an error: the substitution
// route stays in use for this fit.
int build_factor_inverse(gpx_t* h) {
    const int W = h->W, cap = h->cap_nt, nt = h->nt;
    std::vector<int64_t> off(cap);
    int64_t tiles = 0;
    for (int c = 0; c < cap; c++) {
        int p1 = (c / W + 1) * W;
        if (p1 > cap) p1 = cap;
        off[c] = tiles * GPX_TILE_ELEMS;
        tiles += p1;
    }
    // tile rows per CTA: about 8 CTAs per SM over the nt (nt + 1) / 2 tiles of a full sweep, at capacity
    const int64_t total = (int64_t)cap * (cap + 1) / 2;
    int C = (int)((total + 8LL * h->num_sms - 1) / (8LL * h->num_sms));
    if (C < 1) C = 1;
    const int S = (cap - 1) / C + 1;
    const size_t part_bytes = (finv_partial_doubles(h, cap, S) + (size_t)kVecRows * cap + 8) * 8;
    if (h->finv_layout_cap != cap || h->finv_layout_w != W || !h->finv.p) {
        release(h, h->finv); release(h, h->finv_partial);
        size_t fr = 0, tot = 0;
        GPX_CUDA(h, cudaMemGetInfo(&fr, &tot));
        const double need = (double)tiles * GPX_TILE_BYTES + (double)part_bytes;
        if (need > 0.6 * (double)fr || ensure(h, h->finv, (size_t)tiles * GPX_TILE_BYTES) || ensure(h, h->finv_partial, part_bytes) ||
            ensure(h, h->finv_coloff, (size_t)cap * 8)) {
            release(h, h->finv); release(h, h->finv_partial);
            h->finv_layout_cap = -1;
            h->finv_failed = true;
            h->err.clear();
            return 0;
        }
        GPX_CUDA(h, cudaMemcpyAsync(h->finv_coloff.p, off.data(), (size_t)cap * 8, cudaMemcpyHostToDevice, h->st));
        GPX_CUDA(h, cudaStreamSynchronize(h->st));   // off is a host temporary
        h->finv_layout_cap = cap; h->finv_layout_w = W;
    }
    h->finv_C = C; h->finv_S = S;
    GPX_CUDA(h, cudaMemsetAsync(finv_tickets(h), 0, 64, h->st));
    TileMat Y{ptr<double>(h->finv), ptr<int64_t>(h->finv_coloff), 0};
    GPX_CUDA(h, cudaMemsetAsync(h->finv.p, 0, (size_t)tiles * GPX_TILE_BYTES, h->st));
    diag_identity_kernel<<<nt, GPX_TILE, 0, h->st>>>(Y, 0);
    h->launches++;
    const int keep = h->minv_active;
    h->minv_active = 0;
    for (int p = 0; p < h->npanels; p++) var_trsm_panel(h, Y, nt, lmat(h), p, true);
    h->minv_active = keep;
    h->finv_valid = true;
    h->gen++;   // captured predict graphs of this fit still run the substitution chain
    return 0;
}

// gpx_append through the tile route (new tile rows from t0 on) with the inverse resident: the tile columns of Y = L^-T left
// of t0's panel are unchanged (leading blocks of a triangular inverse), so only the columns [c0, nt) are recomputed --
// zeroed, unit diagonal, the far updates of the earlier panels restricted to those columns, then the regular panel steps.
// O(N^2 * new columns) like the append itself.
int refresh_factor_inverse(gpx_t* h, int t0) {
    const int W = h->W, nt = h->nt, cap = h->cap_nt, p0 = t0 / W, c0 = p0 * W;
    TileMat Y = finvmat(h);
    int64_t tiles = 0;
    for (int c = 0; c < nt; c++) {
        int p1 = (c / W + 1) * W;
        if (p1 > cap) p1 = cap;
        if (c >= c0) GPX_CUDA(h, cudaMemsetAsync(ptr<double>(h->finv) + tiles * GPX_TILE_ELEMS, 0, (size_t)p1 * GPX_TILE_BYTES, h->st));
        tiles += p1;
    }
    diag_identity_kernel<<<nt - c0, GPX_TILE, 0, h->st>>>(Y, c0);
    h->launches++;
    const int keep = h->minv_active;
    h->minv_active = 0;
    for (int q = 0; q < p0; q++) {
        GemmJob u{};
        u.A = Y; u.B = lmat(h); u.C = Y;
        u.tri = 0; u.i0 = 0; u.i1 = (q + 1) * W; u.ncols = nt - c0; u.jbase = c0; u.jW = GPX_JW_CONTIG; u.jstride = 0;
        u.k0 = q * W; u.k1 = (q + 1) * W; u.bdiag = 0; u.overwrite = 0;
        gemm(h, u, h->st);
    }
    for (int p = p0; p < h->npanels; p++) var_trsm_panel(h, Y, nt, lmat(h), p, true);
    h->minv_active = keep;
    return 0;
}

// the graph-capturable part of the latency path: loader -> K* (+ mean partials, + right-hand-side vectors) -> mean
// -> [forward substitution -> variance].  Returns the number of kernels launched.
// xin / mean_out / var_out: the workspace copies, or (host callers) the mapped pinned buffer itself -- the loader reads the
// test rows and the reduce kernels write the results straight through it, which saves three cudaMemcpyAsync calls per predict.
int small_chain(gpx_t* h, const SmallWs& w, int64_t Ns, bool want_var, int var_cols, int include_noise, cudaStream_t st,
                const double* xin, double* mean_out, double* var_out) {
    const int nt = h->nt, d = h->d, O = h->O;
    const int64_t Np = h->Np;
    int n = 0;
    if (want_var && Ns > kVecRows) {
        // 9 .. kInvRows rows with the explicit inverse resident: groups of 8 rows, each its own fused K* + GEMV + variance chain
        // (the workspaces are re-used group after group on the stream)
        for (int64_t r0 = 0; r0 < Ns; r0 += kVecRows) {
            const int64_t ns = Ns - r0 < kVecRows ? Ns - r0 : kVecRows;
            n += small_chain(h, w, ns, true, var_cols, include_noise, st, xin + r0 * d, mean_out + r0 * O, var_out + r0 * var_cols);
        }
        return n;
    }
    if (gpx_small_kstar_ok(Ns, d, O)) {   // <= 8 rows: loader + K* + mean in one kernel
        gpx_launch_small_kstar(xin, (int)Ns, d, ptr<double>(h->ls), h->ard, x_shift(h), x_scale(h), ptr<double>(h->Xs),
                               ptr<double>(h->xsq), h->N, nt, h->kernel_id, h->variance, h->rdiv, ptr<double>(h->alpha), O, Np,
                               h->mean_const, y_stats(h), want_var ? w.b : nullptr, w.kpartial, w.ticket, mean_out, st); n++;
    } else {
        gpx_launch_prep_x(xin, Ns, GPX_TILE, d, ptr<double>(h->ls), h->ard, x_shift(h), x_scale(h), w.Xt, w.xtsq, st); n++;
        gpx_launch_kcross(TileMat{nullptr, nullptr, 0}, 1, 1, nt, w.Xt, w.xtsq, Ns, ptr<double>(h->Xs), ptr<double>(h->xsq), h->N, d,
                          h->kernel_id, h->variance, h->rdiv, ptr<double>(h->alpha), O, Np, w.mean_partial, st,
                          want_var ? w.b : nullptr, (int)Ns); n++;
        gpx_launch_mean_reduce(w.mean_partial, 1, nt, O, h->mean_const, y_stats(h), mean_out, 0, Ns, st); n++;
    }
    if (want_var) {
        const double prior = h->variance + (include_noise ? h->noise : 0.0);
        if (h->finv_valid) {   // explicit inverse: one triangular GEMV, no dependent chain
            const int rhs = finv_rhs(Ns);
            n += launch_finv_gemv(h, (int)Ns, w.b, st);
            finv_var_kernel<<<dim3(nt, (int)Ns), 128, 0, st>>>(ptr<double>(h->finv_partial), nt, h->finv_C, h->finv_S, rhs,
                                                              finv_gstride(h, rhs), prior, y_stats(h), var_cols, finv_colsq(h),
                                                              h->cap_nt, finv_tickets(h), var_out); n++;
        } else {
            for (int J = 0; J < nt; J++) {
                gpx_launch_fwd_step(lmat(h), ptr<double>(h->linv), nt, J, w.b, w.z, Np, (int)Ns, 2 * h->num_sms, st); n++;
            }
            vec_var_kernel<<<(int)Ns, 256, 0, st>>>(w.z, (int64_t)nt * GPX_TILE, Np, prior, y_stats(h), var_cols, var_out); n++;
        }
    }
    return n;
}

struct GraphKey { uint64_t gen; int64_t Ns; int want_var, var_cols, include_noise; };

// column-packed factor storage for `cap` tile rows: offsets of all columns (single GPU: every column is owned)
int64_t packed_offsets(int cap, std::vector<int64_t>& off) {
    off.resize(cap);
    int64_t o = 0;
    for (int c = 0; c < cap; c++) { off[c] = o; o += (int64_t)(cap - c) * GPX_TILE_ELEMS; }
    return o;
}

// re-lay the factor and the vectors out for new_cap tile rows (gpx_append ran out of reserve)
int relayout(gpx_t* h, int new_cap) {
    const int nt = h->nt, O = h->O, d = h->d;
    const int64_t oldNp = h->Np, newNp = (int64_t)new_cap * GPX_TILE, rows = (int64_t)nt * GPX_TILE;
    std::vector<int64_t> noff, lc(new_cap);
    const int64_t elems = packed_offsets(new_cap, noff);
    for (int c = 0; c < new_cap; c++) lc[c] = (int64_t)c * GPX_TILE_ELEMS;
    DevBuf nL, nlinv, nXs, nxsq, nres, nw1, nw2, nal, ncoloff, nlcoloff;
    h->zkeep_valid = false;   // the kept forward solution is simply recomputed by the next append (full forward pass)
    int rc;
    auto bail = [&](int code) { for (DevBuf* b : {&nL, &nlinv, &nXs, &nxsq, &nres, &nw1, &nw2, &nal, &ncoloff, &nlcoloff}) release(h, *b); return code; };
    if ((rc = ensure(h, nL, (size_t)elems * 8)) || (rc = ensure(h, nlinv, (size_t)new_cap * GPX_TILE_BYTES)) ||
        (rc = ensure(h, nXs, (size_t)newNp * d * 8)) || (rc = ensure(h, nxsq, (size_t)newNp * 8)) ||
        (rc = ensure(h, nres, (size_t)newNp * O * 8)) || (rc = ensure(h, nw1, (size_t)newNp * O * 8)) ||
        (rc = ensure(h, nw2, (size_t)newNp * O * 8)) || (rc = ensure(h, nal, (size_t)newNp * O * 8)) ||
        (rc = ensure(h, ncoloff, (size_t)new_cap * 8)) || (rc = ensure(h, nlcoloff, (size_t)new_cap * 8))) return bail(rc);
    cudaStream_t st = h->st;
    for (int c = 0; c < nt; c++)
        GPX_CUDA(h, cudaMemcpyAsync(ptr<double>(nL) + noff[c], ptr<double>(h->L) + h->h_coloff[c], (size_t)(nt - c) * GPX_TILE_BYTES,
                                    cudaMemcpyDeviceToDevice, st));
    GPX_CUDA(h, cudaMemcpyAsync(nlinv.p, h->linv.p, (size_t)nt * GPX_TILE_BYTES, cudaMemcpyDeviceToDevice, st));
    GPX_CUDA(h, cudaMemsetAsync(nXs.p, 0, nXs.bytes, st));
    GPX_CUDA(h, cudaMemsetAsync(nxsq.p, 0, nxsq.bytes, st));
    GPX_CUDA(h, cudaMemcpyAsync(nXs.p, h->Xs.p, (size_t)rows * d * 8, cudaMemcpyDeviceToDevice, st));
    GPX_CUDA(h, cudaMemcpyAsync(nxsq.p, h->xsq.p, (size_t)rows * 8, cudaMemcpyDeviceToDevice, st));
    DevBuf* olds[] = {&h->resid, &h->work1, &h->work2, &h->alpha};
    DevBuf* news[] = {&nres, &nw1, &nw2, &nal};
    for (int i = 0; i < 4; i++) {
        GPX_CUDA(h, cudaMemsetAsync(news[i]->p, 0, news[i]->bytes, st));
        GPX_CUDA(h, cudaMemcpy2DAsync(news[i]->p, (size_t)newNp * 8, olds[i]->p, (size_t)oldNp * 8, (size_t)rows * 8, (size_t)O,
                                      cudaMemcpyDeviceToDevice, st));
    }
    GPX_CUDA(h, cudaMemcpyAsync(ncoloff.p, noff.data(), (size_t)new_cap * 8, cudaMemcpyHostToDevice, st));
    GPX_CUDA(h, cudaMemcpyAsync(nlcoloff.p, lc.data(), (size_t)new_cap * 8, cudaMemcpyHostToDevice, st));
    GPX_CUDA(h, cudaStreamSynchronize(st));
    auto swap_in = [&](DevBuf& dst, DevBuf& src) { release(h, dst); dst = src; src = DevBuf(); };
    swap_in(h->L, nL); swap_in(h->linv, nlinv); swap_in(h->Xs, nXs); swap_in(h->xsq, nxsq);
    swap_in(h->resid, nres); swap_in(h->work1, nw1); swap_in(h->work2, nw2); swap_in(h->alpha, nal);
    swap_in(h->coloff, ncoloff); swap_in(h->linv_coloff, nlcoloff);
    h->h_coloff = noff;
    h->cap_nt = new_cap;
    h->Np = newNp;
    h->gen++;
    finv_invalidate(h);
    return 0;
}

// out[row tiles t0..t1) = K(X, X) vec  (K rebuilt tile by tile with the cross-kernel builder; vec, out: [Np])
int kmatvec(gpx_t* h, const double* vec, double* out, DevBuf& partial, int t0, int t1) {
    const int nt = h->nt, d = h->d;
    const int nbmax = 64;
    int rc = ensure(h, partial, (size_t)nt * nbmax * GPX_TILE * 8);
    if (rc) return rc;
    for (int t = t0; t < t1; t += nbmax) {
        const int nb = (t1 - t < nbmax) ? (t1 - t) : nbmax;
        const int64_t row0 = (int64_t)t * GPX_TILE;
        int64_t valid = h->N - row0;
        if (valid > (int64_t)nb * GPX_TILE) valid = (int64_t)nb * GPX_TILE;
        if (valid < 0) valid = 0;
        gpx_launch_kcross(TileMat{nullptr, nullptr, 0}, nb, nb, nt, ptr<double>(h->Xs) + row0 * d, ptr<double>(h->xsq) + row0, valid,
                          ptr<double>(h->Xs), ptr<double>(h->xsq), h->N, d, h->kernel_id, h->variance, h->rdiv, vec, 1, h->Np,
                          ptr<double>(partial), h->st);
        gpx_launch_mean_reduce(ptr<double>(partial), nb, nt, 1, 0.0, nullptr, out, row0, row0 + (int64_t)nb * GPX_TILE, h->st);
        h->launches += 2;
    }
    return 0;
}

}  // namespace

// gpx_append for k <= 8 observations that fit into the partially filled last tile: k forward substitutions (one pass over
// L for all of them), one finishing CTA, the backward substitution -- no 128-row tile work.  Returns 1000 if it does not apply.
static int append_vec(gpx_t* h, const double* Xn, int64_t k, const double* Yn, int mem) {
    const int d = h->d, O = h->O, nt = h->nt, W = h->W;
    const int64_t N0 = h->N, Np = h->Np;
    const int r0 = (int)(N0 % GPX_TILE), t0 = (int)(N0 / GPX_TILE);
    if (k > kAppMax || r0 == 0 || r0 + k > GPX_TILE || O > kAppMax || !h->zkeep_valid || t0 != nt - 1) return 1000;
    int rc = small_workspace(h);
    if (rc) return rc;
    const SmallWs w = small_views(h);
    cudaStream_t st = h->st;
    double* pin = reinterpret_cast<double*>(h->pin);
    double* dY = w.var;   // k * O <= 128 * O doubles of scratch
    if (mem == GPX_MEM_HOST) {
        memcpy(pin, Xn, (size_t)k * d * 8);
        memcpy(pin + GPX_TILE * d, Yn, (size_t)k * O * 8);
        GPX_CUDA(h, cudaMemcpyAsync(w.Xin, pin, (size_t)k * d * 8, cudaMemcpyHostToDevice, st));
        GPX_CUDA(h, cudaMemcpyAsync(dY, pin + GPX_TILE * d, (size_t)k * O * 8, cudaMemcpyHostToDevice, st));
    } else {
        GPX_CUDA(h, cudaMemcpyAsync(w.Xin, Xn, (size_t)k * d * 8, cudaMemcpyDeviceToDevice, st));
        GPX_CUDA(h, cudaMemcpyAsync(dY, Yn, (size_t)k * O * 8, cudaMemcpyDeviceToDevice, st));
    }
    h->fitted = false;
    GPX_CUDA(h, cudaMemsetAsync(h->info.p, 0, 64, st));
    prep_inputs(h, w.Xin, k, k, ptr<double>(h->Xs) + N0 * d, ptr<double>(h->xsq) + N0, st);          // into the training set
    prep_inputs(h, w.Xin, k, GPX_TILE, w.Xt, w.xtsq, st);                                             // aligned copy for the tile builder
    resid_append_kernel<<<(int)((k * O + 255) / 256), 256, 0, st>>>(dY, k, O, h->mean_const, ptr<double>(h->resid), Np, N0, k);
    gpx_launch_kcross(TileMat{nullptr, nullptr, 0}, 1, 1, nt, w.Xt, w.xtsq, k, ptr<double>(h->Xs), ptr<double>(h->xsq), N0, d,
                      h->kernel_id, h->variance, h->rdiv, nullptr, 0, Np, nullptr, st, w.b, (int)k);
    // forward substitutions z_a = L^-1 k(X, x_new_a): one triangular GEMV when the explicit inverse of the factor is there
    // (it is then kept up to date below, and alpha comes from it too: no dependent chain at all), else nt substitution steps
    const bool use_finv = h->finv_valid && h->finv_layout_cap == h->cap_nt;
    double* cinv = w.mean;   // kAppMax^2 <= 128 * O doubles of scratch
    if (use_finv) {
        const int rhs = finv_rhs(k);
        h->launches += launch_finv_gemv(h, (int)k, w.b, st) + 1;
        finv_collect_kernel<<<dim3(nt, (int)k), GPX_TILE, 0, st>>>(ptr<double>(h->finv_partial), h->finv_C, h->finv_S, rhs,
                                                                  finv_gstride(h, rhs), Np, w.z);
    } else {
        for (int J = 0; J < nt; J++)
            gpx_launch_fwd_step(lmat(h), ptr<double>(h->linv), nt, J, w.b, w.z, Np, (int)k, 2 * h->num_sms, st);
        h->launches += nt;
    }
    AppendArgs a{};
    a.z = w.z; a.xt = w.Xt; a.xtsq = w.xtsq; a.zy = ptr<double>(h->zkeep); a.resid = ptr<double>(h->resid);
    a.linv_tile = ptr<double>(h->linv) + (int64_t)t0 * GPX_TILE_ELEMS; a.info = ptr<int>(h->info);
    a.cinv_out = use_finv ? cinv : nullptr;
    a.N0 = N0; a.Np = Np; a.k = (int)k; a.d = d; a.O = O; a.t0 = t0; a.r0 = r0; a.kernel_id = h->kernel_id;
    a.variance = h->variance; a.rdiv = h->rdiv; a.diag_add = h->noise + h->jitter;
    append_finish_kernel<<<1, 1024, 0, st>>>(lmat(h), a);
    h->launches += 3;
    h->N = N0 + k;
    h->gen++;
    if (h->minv_valid > t0 / W) h->minv_valid = t0 / W;
    if (use_finv) {
        // the k new columns of Y = L^-T (U = Y Z lands in w.b), then alpha = L^-T z_y = Y z_y
        h->launches += finv_row_apply(h, w.z, (int)k, w.b, st) + 1;
        finv_append_cols_kernel<<<(int)((N0 + k + 255) / 256), 256, 0, st>>>(finvmat(h), w.b, cinv, N0, Np, t0, r0, (int)k);
        h->launches += finv_row_apply(h, ptr<double>(h->zkeep), O, ptr<double>(h->alpha), st);
        rc = 0;
    } else {
        finv_invalidate(h);
        // alpha: backward substitution over the whole factor (it consumes its right-hand side: a copy of the kept z), then the LML
        GPX_CUDA(h, cudaMemcpyAsync(h->work2.p, h->zkeep.p, (size_t)Np * O * 8, cudaMemcpyDeviceToDevice, st));
        rc = backward_solve(h);
    }
    gpx_launch_lml(ptr<double>(h->linv), ptr<double>(h->alpha), ptr<double>(h->resid), (int64_t)nt * GPX_TILE, Np, O, ptr<double>(h->scal), st);
    h->launches++;
    double hs[2] = {0, 0};
    int hinfo = 0;
    cudaMemcpyAsync(hs, h->scal.p, 16, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&hinfo, h->info.p, 4, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc) return rc;
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, -2, "CUDA error during append: %s", cudaGetErrorString(e));
    if (hinfo != 0) {
        h->zkeep_valid = false;
        finv_invalidate(h);
        fail(h, hinfo, "append: Ky is not positive definite: non-positive pivot at row %d", hinfo);
        return hinfo;
    }
    h->logdet = hs[0];
    h->quad = hs[1];
    const double log2pi = 1.8378770664093454835606594728112;
    h->lml = 0.5 * (-(double)h->N * O * log2pi - (double)O * h->logdet - h->quad);
    h->fitted = true;
    return 0;
}

// ====================================================================================================================
int gpx_predict_small(gpx_t* h, const double* Xs, int64_t Ns, double* mean, double* var, int var_cols, int include_noise,
                      int mem) {
    const int nt = h->nt, d = h->d, O = h->O;
    for (int ph : {GPX_PHASE_PREDICT_KSTAR, GPX_PHASE_PREDICT_TRSM, GPX_PHASE_PREDICT_REDUCE, GPX_PHASE_H2D, GPX_PHASE_D2H}) h->phase_ms[ph] = 0.f;
    int rc = small_workspace(h);
    if (rc) return rc;
    const SmallWs w = small_views(h);
    cudaStream_t st = h->st;
    double* pin = reinterpret_cast<double*>(h->pin);
    const bool want_var = var != nullptr;
    // host callers of the chain route: zero-copy through the mapped pinned buffer [X 128 d | mean 128 O | var 128 O]
    // variance routes: <= 8 rows GEMV-style (substitution chain or explicit inverse); 9 .. 64 rows through the explicit inverse
    // when it is resident; otherwise one tile row through the blocked DMMA solve
    auto chain_route = [&]() { return !want_var || Ns <= kVecRows || (h->finv_valid && Ns <= kInvRows); };
    double *pin_mean = pin + GPX_TILE * d, *pin_var = pin_mean + GPX_TILE * O;
    if (mem == GPX_MEM_HOST) {
        memcpy(pin, Xs, (size_t)Ns * d * 8);   // (copied on to the workspace below unless the kernels read it in place)
    } else {
        GPX_CUDA(h, cudaMemcpyAsync(w.Xin, Xs, (size_t)Ns * d * 8, cudaMemcpyDeviceToDevice, st));
    }
    if (want_var && Ns <= kInvRows && !h->finv_valid && !h->finv_failed && h->latency_inverse != 0) {
        // the acquisition / IPOPT loops predict one row at a time, thousands of times per fit: after a few such calls (or at
        // the first with latency_inverse = 1) pay N^3/3 flop once for the explicit inverse of the factor
        h->finv_calls++;
        if (h->latency_inverse > 0 || h->finv_calls >= kFinvAutoCalls) {
            if ((rc = build_factor_inverse(h))) return rc;
        }
    }
    const bool zc = mem == GPX_MEM_HOST && chain_route();
    if (mem == GPX_MEM_HOST && !zc) GPX_CUDA(h, cudaMemcpyAsync(w.Xin, pin, (size_t)Ns * d * 8, cudaMemcpyHostToDevice, st));
    const double* xin = zc ? pin : w.Xin;
    double *mean_out = zc ? pin_mean : w.mean, *var_out = zc ? pin_var : w.var;
    if (chain_route()) {
        if (h->predict_graph) {
            // replay (or capture once per fit generation and shape) the whole launch chain as one CUDA graph
            gpx_handle::SmallGraph* g = nullptr;
            for (auto& e : h->graphs)
                if (e.gen == h->gen && e.Ns == Ns && e.want_var == (int)want_var && e.var_cols == var_cols && e.include_noise == include_noise &&
                    e.hostio == (int)zc) g = &e;
            if (!g) {
                for (auto it = h->graphs.begin(); it != h->graphs.end();) {   // drop stale generations, bound the cache
                    if (it->gen != h->gen || h->graphs.size() >= 8) { cudaGraphExecDestroy(it->exec); it = h->graphs.erase(it); }
                    else ++it;
                }
                cudaGraph_t graph = nullptr;
                GPX_CUDA(h, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                const int nodes = small_chain(h, w, Ns, want_var, var_cols, include_noise, st, xin, mean_out, var_out);
                cudaError_t ec = cudaStreamEndCapture(st, &graph);
                if (ec != cudaSuccess || !graph) { cudaGetLastError(); return fail(h, -2, "CUDA graph capture of the predict chain failed: %s", cudaGetErrorString(ec)); }
                gpx_handle::SmallGraph e{};
                e.gen = h->gen; e.Ns = Ns; e.want_var = want_var; e.var_cols = var_cols; e.include_noise = include_noise; e.nodes = nodes; e.hostio = zc;
                ec = cudaGraphInstantiate(&e.exec, graph, 0);
                cudaGraphDestroy(graph);
                if (ec != cudaSuccess) { cudaGetLastError(); return fail(h, -2, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ec)); }
                h->graphs.push_back(e);
                g = &h->graphs.back();
            }
            GPX_CUDA(h, cudaGraphLaunch(g->exec, st));
            h->launches += g->nodes;
        } else {
            h->launches += small_chain(h, w, Ns, want_var, var_cols, include_noise, st, xin, mean_out, var_out);
        }
    } else {
        // 9..128 rows with variance: one tile row through the blocked DMMA solve (workspace kept between calls)
        if ((rc = ensure(h, h->Tt, (size_t)h->cap_nt * GPX_TILE_BYTES))) return rc;
        TileMat Tt{ptr<double>(h->Tt), ptr<int64_t>(h->linv_coloff), 0};   // one tile row: column c at c * 128 * 128
        gpx_launch_prep_x(w.Xin, Ns, GPX_TILE, d, ptr<double>(h->ls), h->ard, x_shift(h), x_scale(h), w.Xt, w.xtsq, st);
        gpx_launch_kcross(Tt, 1, 1, nt, w.Xt, w.xtsq, Ns, ptr<double>(h->Xs), ptr<double>(h->xsq), h->N, d, h->kernel_id, h->variance,
                          h->rdiv, ptr<double>(h->alpha), O, h->Np, w.mean_partial, st);
        gpx_launch_mean_reduce(w.mean_partial, 1, nt, O, h->mean_const, y_stats(h), w.mean, 0, Ns, st);
        h->launches += 3;
        if ((rc = var_trsm(h, Tt, 1))) return rc;
        if ((rc = var_reduce(h, Tt, 1, h->variance + (include_noise ? h->noise : 0.0), var_cols, w.var, 0, Ns))) return rc;
    }
    const size_t mb = (size_t)Ns * O * 8, vb = want_var ? (size_t)Ns * var_cols * 8 : 0;
    if (zc) {
        // results are already on their way into the pinned buffer
    } else if (mem == GPX_MEM_HOST) {
        GPX_CUDA(h, cudaMemcpyAsync(pin_mean, w.mean, mb, cudaMemcpyDeviceToHost, st));
        if (want_var) GPX_CUDA(h, cudaMemcpyAsync(pin_var, w.var, vb, cudaMemcpyDeviceToHost, st));
    } else {
        GPX_CUDA(h, cudaMemcpyAsync(mean, w.mean, mb, cudaMemcpyDeviceToDevice, st));
        if (want_var) GPX_CUDA(h, cudaMemcpyAsync(var, w.var, vb, cudaMemcpyDeviceToDevice, st));
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, -2, "CUDA error during predict (latency path): %s", cudaGetErrorString(e));
    if (mem == GPX_MEM_HOST) {
        memcpy(mean, pin_mean, mb);
        if (want_var) memcpy(var, pin_var, vb);
    }
    gemm_collect(h);
    return 0;
}

extern "C" {

int gpx_set_normalizer(gpx_t* h, int normalize_input, int normalize_output) {
    if (!h) return -1;
    h->norm_x = normalize_input ? 1 : 0;
    h->norm_y = normalize_output ? 1 : 0;
    return 0;
}

int gpx_get_normalizer(gpx_t* h, double* x_mean, double* x_std, double* y_mean, double* y_std) {
    if (!h) return -1;
    if (!h->fitted) return fail(h, -1, "gpx_get_normalizer called before a successful gpx_fit");
    const int d = h->d, O = h->O;
    for (int k = 0; k < d; k++) {
        if (x_mean) x_mean[k] = h->fit_norm_x ? h->h_xstats[k] : 0.0;
        if (x_std) x_std[k] = h->fit_norm_x ? h->h_xstats[d + k] : 1.0;
    }
    for (int o = 0; o < O; o++) {
        if (y_mean) y_mean[o] = h->fit_norm_y ? h->h_ystats[o] : 0.0;
        if (y_std) y_std[o] = h->fit_norm_y ? h->h_ystats[O + o] : 1.0;
    }
    return 0;
}

int gpx_set_stream(gpx_t* h, void* stream, int has_stream) {
    if (!h) return -1;
    h->caller_stream = reinterpret_cast<cudaStream_t>(stream);
    h->has_caller_stream = has_stream != 0;
    return 0;
}

int gpx_append(gpx_t* h, const double* Xn, int64_t k, const double* Yn, int mem) {
    if (!h) return -1;
    if (!h->fitted) return fail(h, -1, "gpx_append called before a successful gpx_fit");
    if (!Xn || !Yn || k < 1) return fail(h, -1, "bad argument");
    if (h->world > 1) return fail(h, -8, "gpx_append: not available on a sharded handle (refit)");
    if (h->fit_norm_x || h->fit_norm_y) return fail(h, -8, "gpx_append: the normaliser statistics change with the data (refit)");
    GPX_CUDA(h, cudaSetDevice(h->device));
    { int rcw = wait_for_caller(h); if (rcw) return rcw; }
    {   // a handful of points inside the current tile: the GEMV-style path
        const int rv = append_vec(h, Xn, k, Yn, mem);
        if (rv != 1000) return rv;
    }
    const int d = h->d, O = h->O, W = h->W;
    const int64_t N0 = h->N, N1 = N0 + k;
    const int t0 = (int)(N0 / GPX_TILE), nt1 = (int)((N1 + GPX_TILE - 1) / GPX_TILE);
    int rc;
    if (nt1 > h->cap_nt) {
        int grow = h->cap_nt / 4 > 1 ? h->cap_nt / 4 : 1;
        int new_cap = h->cap_nt + grow > nt1 ? h->cap_nt + grow : nt1;
        if ((rc = relayout(h, new_cap))) return rc;
    }
    h->fitted = false;   // until the append has succeeded
    cudaStream_t st = h->st;
    const double *dX, *dY;
    DevBuf& stageY = h->stage_y;
    if ((rc = to_device(h, Xn, (size_t)k * d, mem, h->stage, &dX))) return rc;
    if ((rc = to_device(h, Yn, (size_t)k * O, mem, stageY, &dY))) return rc;
    const int64_t Np = h->Np, rows_padded = (int64_t)nt1 * GPX_TILE - N0;
    prep_inputs(h, dX, k, rows_padded, ptr<double>(h->Xs) + N0 * d, ptr<double>(h->xsq) + N0, st);
    resid_append_kernel<<<(int)((rows_padded * O + 255) / 256), 256, 0, st>>>(dY, k, O, h->mean_const, ptr<double>(h->resid), Np, N0,
                                                                            rows_padded);
    h->launches++;
    h->N = N1; h->nt = nt1; h->npanels = (nt1 + W - 1) / W;
    h->gen++;
    const bool had_finv = h->finv_valid && h->finv_layout_cap == h->cap_nt && h->finv_layout_w == W;   // (a re-layout dropped it)
    finv_invalidate(h);
    GPX_CUDA(h, cudaMemsetAsync(h->info.p, 0, 64, st));
    // 1. the new rows of Ky (tile rows t0 .. nt1-1; a partially filled tile row t0 is rebuilt whole)
    gpx_launch_kbuild_rows(lmat(h), nt1, t0, ptr<double>(h->Xs), ptr<double>(h->xsq), N1, d, h->kernel_id, h->variance, h->rdiv,
                           h->noise + h->jitter, st);
    h->launches++;
    // 2. solve them against the resident factor: L(R, J) = (K(R, J) - sum_{k<J} L(R,k) L(J,k)^T) Linv_J^T for J < t0, and
    //    subtract their outer product from the new diagonal block -- panel by panel, exactly the predictive-variance solve
    //    Panels that lie entirely below the new rows go through their inverted diagonal block (one row-wise GEMM instead
    //    of 2 W - 1 single-tile launches: an append is latency-bound); the blocks are cached across appends, only the panels
    //    the append touches are invalidated.
    TileMat L = lmat(h);
    const int full_panels = t0 / W;
    if (h->minv_valid > full_panels) h->minv_valid = full_panels;
    if (full_panels > 0 && (rc = ensure_panel_inverse(h, full_panels))) return rc;
    TileMat Mi{ptr<double>(h->minv), ptr<int64_t>(h->minv_coloff), 1};
    for (int c0 = 0; c0 < t0; c0 += W) {
        const int c1 = c0 + W < t0 ? c0 + W : t0;
        if (c1 - c0 == W && c0 / W < full_panels) {
            GemmJob t{};
            t.A = L; t.B = Mi; t.C = L;
            t.tri = 0; t.i0 = t0; t.i1 = nt1; t.ncols = W; t.jbase = c0; t.jW = GPX_JW_CONTIG; t.jstride = 0;
            t.k0 = c0; t.k1 = c1; t.bdiag = 0; t.overwrite = 1; t.rowwise = 1;
            gemm(h, t, st);
        } else {
            trsm_inner_stepwise(h, L, nt1, L, c0, c1, t0);
        }
        if (c1 < t0) {
            GemmJob u{};
            u.A = L; u.B = L; u.C = L;
            u.tri = 0; u.i0 = t0; u.i1 = nt1; u.ncols = t0 - c1; u.jbase = c1; u.jW = GPX_JW_CONTIG; u.jstride = 0;
            u.k0 = c0; u.k1 = c1; u.bdiag = 0; u.overwrite = 0;
            gemm(h, u, st);
        }
        GemmJob s{};
        s.A = L; s.B = L; s.C = L;
        s.tri = 1; s.i0 = t0; s.i1 = nt1; s.ncols = nt1 - t0; s.jbase = t0; s.jW = GPX_JW_CONTIG; s.jstride = 0;
        s.k0 = c0; s.k1 = c1; s.bdiag = 0; s.overwrite = 0;
        gemm(h, s, st);
    }
    // 3. Cholesky of the Schur complement (the new diagonal block)
    factor_cols(h, t0, nt1, nt1, st);
    // 4. alpha and the LML.  Forward substitution: only the new tile rows -- z of the old rows is what the fit (or the
    //    previous append) left in zkeep; b_R = resid_R - sum_{J < t0} L(R, J) z_J by a split row GEMV, then the few steps inside
    //    the new block.  (More than kRowRhs outputs, or no kept z: the plain full pass.)  Backward substitution: all of L.
    GPX_CUDA(h, cudaMemcpyAsync(h->work1.p, h->resid.p, (size_t)Np * O * 8, cudaMemcpyDeviceToDevice, st));
    if ((rc = ensure(h, h->zkeep, (size_t)Np * O * 8))) return rc;
    int jfirst = 0;
    if (h->zkeep_valid && O <= kRowRhs && t0 > 0) {
        const int rows = nt1 - t0;
        int S = (2 * h->num_sms + rows - 1) / rows;
        if (S > t0) S = t0;
        if ((rc = ensure(h, h->var_partial, (size_t)rows * S * O * GPX_TILE * 8))) return rc;
        GPX_CUDA(h, cudaMemcpyAsync(h->work2.p, h->zkeep.p, (size_t)Np * O * 8, cudaMemcpyDeviceToDevice, st));
        row_gemv_partial_kernel<<<dim3(rows, S), 256, 0, st>>>(L, t0, t0, S, ptr<double>(h->work2), Np, O, ptr<double>(h->var_partial), 0);
        row_gemv_apply_kernel<<<(int)(((int64_t)rows * O * GPX_TILE + 255) / 256), 256, 0, st>>>(ptr<double>(h->var_partial), rows, S, O, t0, Np,
                                                                                              ptr<double>(h->work1), 0);
        h->launches += 2;
        jfirst = t0;
    }
    for (int J = jfirst; J < nt1; J++) {
        gpx_launch_fwd_step(L, ptr<double>(h->linv), nt1, J, ptr<double>(h->work1), ptr<double>(h->work2), Np, O, 2 * h->num_sms, st);
        h->launches++;
    }
    GPX_CUDA(h, cudaMemcpyAsync(h->zkeep.p, h->work2.p, (size_t)Np * O * 8, cudaMemcpyDeviceToDevice, st));
    h->zkeep_valid = true;
    rc = backward_solve(h);
    gpx_launch_lml(ptr<double>(h->linv), ptr<double>(h->alpha), ptr<double>(h->resid), (int64_t)nt1 * GPX_TILE, Np, O,
                   ptr<double>(h->scal), st);
    h->launches++;
    // keep the explicit inverse of the factor (latency path) up to date: only its tile columns from t0's panel on change
    if (!rc && had_finv && !(rc = refresh_factor_inverse(h, t0))) h->finv_valid = true;
    double hs[2] = {0, 0};
    int hinfo = 0;
    cudaMemcpyAsync(hs, h->scal.p, 16, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&hinfo, h->info.p, 4, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (rc) { finv_invalidate(h); return rc; }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) { finv_invalidate(h); return fail(h, -2, "CUDA error during append: %s", cudaGetErrorString(e)); }
    gemm_collect(h);
    if (hinfo != 0) {
        finv_invalidate(h);
        fail(h, hinfo, "append: Ky is not positive definite: non-positive pivot at row %d", hinfo);
        return hinfo;
    }
    h->logdet = hs[0];
    h->quad = hs[1];
    const double log2pi = 1.8378770664093454835606594728112;
    h->lml = 0.5 * (-(double)N1 * O * log2pi - (double)O * h->logdet - h->quad);
    h->fitted = true;
    return 0;
}

int gpx_selfcheck(gpx_t* h, double* out, int capacity) {
    if (!h) return -1;
    if (!h->fitted) return fail(h, -1, "gpx_selfcheck called before a successful gpx_fit");
    if (!out || capacity < 4) return fail(h, -1, "output buffer too small: need 4");
    GPX_CUDA(h, cudaSetDevice(h->device));
    const int nt = h->nt, G = h->world, W = h->W;
    const int64_t Np = h->Np, n = (int64_t)nt * GPX_TILE;
    const double c = h->noise + h->jitter;
    cudaStream_t st = h->st;
    DevBuf partial, ka, v, w, u, kv, tgt, sc;
    int rc = 0;
    auto cleanup = [&]() { for (DevBuf* b : {&partial, &ka, &v, &w, &u, &kv, &tgt, &sc}) release(h, *b); };
    if ((rc = ensure(h, ka, (size_t)Np * 8)) || (rc = ensure(h, v, (size_t)Np * 8)) || (rc = ensure(h, w, (size_t)Np * 8)) ||
        (rc = ensure(h, u, (size_t)Np * 8)) || (rc = ensure(h, kv, (size_t)Np * 8)) || (rc = ensure(h, tgt, (size_t)Np * 8)) ||
        (rc = ensure(h, sc, 64))) { cleanup(); return rc; }
    // rows of the K mat-vecs are sharded over the ranks like the test tiles of gpx_predict
    const int t0 = (int)((int64_t)nt * h->rank / G), t1 = (int)((int64_t)nt * (h->rank + 1) / G);
    auto allreduce = [&](DevBuf& b) -> int {
        if (G > 1) GPX_NCCL(h, nccl().AllReduce(b.p, b.p, (size_t)Np, ncclDouble, ncclSum, h->comm, st));
        return 0;
    };
    // (a) || Ky alpha - (y - m) || / || y - m ||   (first output)
    GPX_CUDA(h, cudaMemsetAsync(ka.p, 0, (size_t)Np * 8, st));
    if ((rc = kmatvec(h, ptr<double>(h->alpha), ptr<double>(ka), partial, t0, t1)) || (rc = allreduce(ka))) { cleanup(); return rc; }
    resid_norm_kernel<<<1, 1024, 0, st>>>(ptr<double>(ka), ptr<double>(h->alpha), c, ptr<double>(h->resid), n, ptr<double>(sc));
    // (b) || L (L^T v) - Ky v || / || Ky v ||
    fill_random_kernel<<<(int)((Np + 255) / 256), 256, 0, st>>>(ptr<double>(v), h->N, Np);
    GPX_CUDA(h, cudaMemsetAsync(w.p, 0, (size_t)Np * 8, st));
    GPX_CUDA(h, cudaMemsetAsync(u.p, 0, (size_t)Np * 8, st));
    GPX_CUDA(h, cudaMemsetAsync(kv.p, 0, (size_t)Np * 8, st));
    const int q0 = next_owned_panel(h, 0);
    const int ncols = q0 < h->npanels ? owned_cols_from(h, q0) : 0;
    const int jbase = q0 < h->npanels ? panel_begin(h, q0) : 0;
    if (ncols > 0) tri_matvec_kernel<<<dim3(nt, ncols), 256, 0, st>>>(lmat(h), nt, jbase, W, W * G, ptr<double>(v), ptr<double>(w), 1);
    if ((rc = allreduce(w))) { cleanup(); return rc; }
    if (ncols > 0) tri_matvec_kernel<<<dim3(nt, ncols), 256, 0, st>>>(lmat(h), nt, jbase, W, W * G, ptr<double>(w), ptr<double>(u), 0);
    if ((rc = allreduce(u))) { cleanup(); return rc; }
    if ((rc = kmatvec(h, ptr<double>(v), ptr<double>(kv), partial, t0, t1)) || (rc = allreduce(kv))) { cleanup(); return rc; }
    axpy_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(ptr<double>(kv), ptr<double>(v), c, ptr<double>(tgt), n);
    // || u - tgt ||: reuse resid_norm with a = u, b = anything, c = 0
    resid_norm_kernel<<<1, 1024, 0, st>>>(ptr<double>(u), ptr<double>(u), 0.0, ptr<double>(tgt), n, ptr<double>(sc) + 2);
    h->launches += 6;
    double hs[4] = {0, 0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(hs, sc.p, 32, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    if (e != cudaSuccess) return fail(h, -2, "CUDA error in gpx_selfcheck: %s", cudaGetErrorString(e));
    out[0] = hs[1] > 0 ? sqrt(hs[0] / hs[1]) : 0.0;
    out[1] = hs[3] > 0 ? sqrt(hs[2] / hs[3]) : 0.0;
    out[2] = sqrt(hs[3]);
    out[3] = sqrt(hs[1]);
    return 0;
}

int gpx_predict_moments(gpx_t* h, const double* Xs, int64_t Ns, double* S0, double* S1, double* S2, int mem) {
    if (!h) return -1;
    if (!h->fitted) return fail(h, -1, "gpx_predict_moments called before a successful gpx_fit");
    if (!Xs || !S0 || Ns < 1) return fail(h, -1, "bad argument");
    if (h->world > 1) return fail(h, -6, "gpx_predict_moments: single-GPU handles only");
    const int d = h->d;
    if (d > kMomentsMaxD) return fail(h, -1, "gpx_predict_moments supports input dimension <= %d", kMomentsMaxD);
    GPX_CUDA(h, cudaSetDevice(h->device));
    { int rcw = wait_for_caller(h); if (rcw) return rcw; }
    cudaStream_t st = h->st;
    static bool attr_done[64] = {false};
    if (h->device < 64 && !attr_done[h->device]) {
        cudaFuncSetAttribute(moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)moments_smem(kMomentsMaxD));
        attr_done[h->device] = true;
    }
    int rc;
    const double* dXs;
    if ((rc = to_device(h, Xs, (size_t)Ns * d, mem, h->stage, &dXs))) return rc;
    DevBuf &xt = h->m_xt, &xtsq = h->m_xtsq, &o0 = h->m_o0, &o1 = h->m_o1, &o2 = h->m_o2;   // kept between calls
    auto cleanup = [&]() {};
    if ((rc = ensure(h, xt, (size_t)Ns * d * 8)) || (rc = ensure(h, xtsq, (size_t)Ns * 8))) { cleanup(); return rc; }
    double *d0 = S0, *d1 = S1, *d2 = S2;
    if (mem == GPX_MEM_HOST) {
        if ((rc = ensure(h, o0, (size_t)Ns * 8)) || (S1 && (rc = ensure(h, o1, (size_t)Ns * d * 8))) ||
            (S2 && (rc = ensure(h, o2, (size_t)Ns * d * d * 8)))) { cleanup(); return rc; }
        d0 = ptr<double>(o0); d1 = S1 ? ptr<double>(o1) : nullptr; d2 = S2 ? ptr<double>(o2) : nullptr;
    }
    prep_inputs(h, dXs, Ns, Ns, ptr<double>(xt), ptr<double>(xtsq), st);
    moments_kernel<<<(int)Ns, 256, moments_smem(d), st>>>(ptr<double>(xt), ptr<double>(xtsq), ptr<double>(h->Xs), ptr<double>(h->xsq),
                                                          h->N, d, h->kernel_id, h->variance, h->rdiv, ptr<double>(h->alpha), d0, d1, d2);
    h->launches++;
    if (mem == GPX_MEM_HOST) {
        cudaMemcpyAsync(S0, d0, (size_t)Ns * 8, cudaMemcpyDeviceToHost, st);
        if (S1) cudaMemcpyAsync(S1, d1, (size_t)Ns * d * 8, cudaMemcpyDeviceToHost, st);
        if (S2) cudaMemcpyAsync(S2, d2, (size_t)Ns * d * d * 8, cudaMemcpyDeviceToHost, st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    if (e != cudaSuccess) return fail(h, -2, "CUDA error in gpx_predict_moments: %s", cudaGetErrorString(e));
    return 0;
}

}  // extern "C"
