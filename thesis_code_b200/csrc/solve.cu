// Triangular solves for alpha = Ky^-1 (Y - m) on the tiled lower factor, plus the log-determinant / quadratic-form
// reduction of the marginal likelihood.
//
// Forward  L z = b   : right-looking over tile columns J:  z_J = Linv_JJ b_J ;  b_I -= L_IJ z_J  (I > J)
// Backward L^T a = z : right-looking over tile rows I (descending):  a_I = Linv_II^T z_I ;  z_J -= L_IJ^T a_I (J < I)
// One launch per step; every CTA recomputes the 128-vector of the step from the inverted diagonal tile (L2
// resident) so no second launch / grid sync is needed, and CTA 0 publishes it.  No atomics: the result is
// deterministic.  Roofline: HBM (L is read twice: 2 * 8 * N(N+1)/2 bytes), plus ~2*nt launch latencies.
#include "gpx_common.cuh"

namespace {

// v_out[128] = Linv (or Linv^T) * v_in[128], computed by all 256 threads into shared memory `dst`.
__device__ __forceinline__ void diag_apply(const double* __restrict__ inv_tile, const double* __restrict__ vin,
                                           bool transpose, double* sv_in, double* dst, double* scratch) {
    const int tid = threadIdx.x, r = tid & 127, half = tid >> 7;
    if (tid < GPX_TILE) sv_in[tid] = vin[tid];
    __syncthreads();
    double s = 0.0;
    if (!transpose) {
        // row r of Linv: sum_c Linv[r][c] v[c], this thread covers chunks half*8 .. half*8+7
#pragma unroll
        for (int kc = 0; kc < 8; kc++) {
            const int ch = half * 8 + kc;
            const double2* p = reinterpret_cast<const double2*>(inv_tile + ch * 1024 + r * 8);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                double2 v = p[u];
                s = fma(v.x, sv_in[ch * 8 + 2 * u], s);
                s = fma(v.y, sv_in[ch * 8 + 2 * u + 1], s);
            }
        }
    } else {
        // column r of Linv: sum_i Linv[i][r] v[i], rows half*64 .. half*64+63 ; element (i, r) at (r>>3)*1024 + i*8 + (r&7)
        const double* p = inv_tile + (r >> 3) * 1024 + (r & 7);
        for (int i = half * 64; i < half * 64 + 64; i++) s = fma(p[i * 8], sv_in[i], s);
    }
    if (half == 1) scratch[r] = s;
    __syncthreads();
    if (half == 0) dst[r] = s + scratch[r];
    __syncthreads();
}

// forward step J.  b, z: [O][Np].  grid.x CTAs stride over tiles I = J+1 .. nt-1.
// Right-hand sides are handled kFwdRhs at a time: every tile of L (and the inverted diagonal tile) is read ONCE per group
// and applied to all of the group's vectors from registers -- with one vector per pass (round 1) a multi-output alpha solve
// or the latency-path variance of several test rows re-read L once per vector.
constexpr int kFwdRhs = 8;
__global__ void __launch_bounds__(256) fwd_step_kernel(TileMat L, const double* __restrict__ linv, int nt, int J,
                                                       double* __restrict__ b, double* __restrict__ z, int64_t Np, int O) {
    __shared__ double sv[kFwdRhs][GPX_TILE], sz[kFwdRhs][GPX_TILE], scratch[kFwdRhs][GPX_TILE];
    const int tid = threadIdx.x, r = tid & 127, half = tid >> 7;
    const double* inv_tile = linv + (int64_t)J * GPX_TILE_ELEMS;
    for (int o0 = 0; o0 < O; o0 += kFwdRhs) {
        const int Oc = (O - o0 < kFwdRhs) ? (O - o0) : kFwdRhs;
        // z_J = Linv_JJ b_J for the group (every CTA recomputes it: no grid-wide dependency)
        for (int e = tid; e < Oc * GPX_TILE; e += 256) sv[e >> 7][e & 127] = b[(int64_t)(o0 + (e >> 7)) * Np + (int64_t)J * GPX_TILE + (e & 127)];
        __syncthreads();
        {
            double acc[kFwdRhs];
#pragma unroll
            for (int o = 0; o < kFwdRhs; o++) acc[o] = 0.0;
#pragma unroll
            for (int kc = 0; kc < 8; kc++) {
                const int ch = half * 8 + kc;
                const double2* p = reinterpret_cast<const double2*>(inv_tile + ch * 1024 + r * 8);
                const double2 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
#pragma unroll
                for (int o = 0; o < kFwdRhs; o++) {
                    if (o < Oc) {
                        const double* x = &sv[o][ch * 8];
                        double s = acc[o];
                        s = fma(v0.x, x[0], s); s = fma(v0.y, x[1], s); s = fma(v1.x, x[2], s); s = fma(v1.y, x[3], s);
                        s = fma(v2.x, x[4], s); s = fma(v2.y, x[5], s); s = fma(v3.x, x[6], s); s = fma(v3.y, x[7], s);
                        acc[o] = s;
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < kFwdRhs; o++) if (o < Oc && half == 1) scratch[o][r] = acc[o];
            __syncthreads();
#pragma unroll
            for (int o = 0; o < kFwdRhs; o++) if (o < Oc && half == 0) sz[o][r] = acc[o] + scratch[o][r];
            __syncthreads();
        }
        if (blockIdx.x == 0)
            for (int e = tid; e < Oc * GPX_TILE; e += 256) z[(int64_t)(o0 + (e >> 7)) * Np + (int64_t)J * GPX_TILE + (e & 127)] = sz[e >> 7][e & 127];
        // b_I -= L_IJ z_J
        for (int I = J + 1 + blockIdx.x; I < nt; I += gridDim.x) {
            const double* tp = tile_ptr(L, I, J);
            double acc[kFwdRhs];
#pragma unroll
            for (int o = 0; o < kFwdRhs; o++) acc[o] = 0.0;
#pragma unroll
            for (int kc = 0; kc < 8; kc++) {
                const int ch = half * 8 + kc;
                const double2* p = reinterpret_cast<const double2*>(tp + ch * 1024 + r * 8);
                const double2 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
#pragma unroll
                for (int o = 0; o < kFwdRhs; o++) {
                    if (o < Oc) {
                        const double* x = &sz[o][ch * 8];
                        double s = acc[o];
                        s = fma(v0.x, x[0], s); s = fma(v0.y, x[1], s); s = fma(v1.x, x[2], s); s = fma(v1.y, x[3], s);
                        s = fma(v2.x, x[4], s); s = fma(v2.y, x[5], s); s = fma(v3.x, x[6], s); s = fma(v3.y, x[7], s);
                        acc[o] = s;
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < kFwdRhs; o++) if (o < Oc && half == 1) scratch[o][r] = acc[o];
            __syncthreads();
#pragma unroll
            for (int o = 0; o < kFwdRhs; o++)
                if (o < Oc && half == 0) b[(int64_t)(o0 + o) * Np + (int64_t)I * GPX_TILE + r] -= acc[o] + scratch[o][r];
            __syncthreads();
        }
    }
}

// backward step I.  z is updated in place for the columns this rank owns; a receives / provides the solution.
// Owned columns below I:  J = jbase + (m / jW) * jstride + (m % jW),  m in [0, ncols)  (see GemmJob).
// compute_alpha = 1: a_I = Linv_II^T z_I is computed here (the owner of column I; z_I is final) and published by
// CTA 0; compute_alpha = 0: a_I was received from its owner and is only applied.
__global__ void __launch_bounds__(256) bwd_step_kernel(TileMat L, const double* __restrict__ linv, int I,
                                                       double* __restrict__ z, double* __restrict__ a, int64_t Np, int O,
                                                       int compute_alpha, int ncols, int jbase, int jW, int jstride) {
    __shared__ double sv_in[GPX_TILE], sa[GPX_TILE], scratch[GPX_TILE];
    __shared__ double part[8][GPX_TILE];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int o = 0; o < O; o++) {
        double* zo = z + (int64_t)o * Np;
        if (compute_alpha) {
            diag_apply(linv + (int64_t)I * GPX_TILE_ELEMS, zo + (int64_t)I * GPX_TILE, true, sv_in, sa, scratch);
            if (blockIdx.x == 0 && tid < GPX_TILE) a[(int64_t)o * Np + (int64_t)I * GPX_TILE + tid] = sa[tid];
        } else {
            if (tid < GPX_TILE) sa[tid] = a[(int64_t)o * Np + (int64_t)I * GPX_TILE + tid];
            __syncthreads();
        }
        for (int m = blockIdx.x; m < ncols; m += gridDim.x) {
            const int J = jbase + (m / jW) * jstride + (m % jW);
            if (J >= I) continue;
            const double* tp = tile_ptr(L, I, J);
            // out[c] = sum_r L[r][c] a[r]: warp w covers rows w*16..w*16+15; lane -> column pair 2*(lane&3) of the
            // chunks {lane>>2, (lane>>2)+8}
            const int t = lane & 3, cg = lane >> 2;
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int kc = cg + 8 * h;
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int i = 0; i < 16; i++) {
                    const int row = warp * 16 + i;
                    double2 v = *reinterpret_cast<const double2*>(tp + kc * 1024 + row * 8 + 2 * t);
                    s0 = fma(v.x, sa[row], s0);
                    s1 = fma(v.y, sa[row], s1);
                }
                part[warp][kc * 8 + 2 * t] = s0;
                part[warp][kc * 8 + 2 * t + 1] = s1;
            }
            __syncthreads();
            if (tid < GPX_TILE) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < 8; w++) s += part[w][tid];
                zo[(int64_t)J * GPX_TILE + tid] -= s;
            }
            __syncthreads();
        }
    }
}

// logdet = 2 sum_i log L_ii = -2 sum_i log (Linv_ii)  (read from the inverted diagonal tiles, which every rank holds);
// quad = sum alpha * resid.  One CTA, fixed order.
// n_rows = nt * 128 rows have a diagonal tile; Np >= n_rows is the padded vector length (zeros beyond N).
__global__ void __launch_bounds__(1024) lml_kernel(const double* __restrict__ linv, const double* __restrict__ alpha,
                                                   const double* __restrict__ resid, int64_t n_rows, int64_t Np, int O,
                                                   double* __restrict__ out2) {
    __shared__ double red[2][32];
    double ld = 0.0, q = 0.0;
    for (int64_t i = threadIdx.x; i < n_rows; i += blockDim.x) {
        int J = (int)(i >> 7), r = (int)(i & 127);
        ld -= log(linv[(int64_t)J * GPX_TILE_ELEMS + gpx_tile_idx(r, r)]);
    }
    for (int64_t i = threadIdx.x; i < Np * O; i += blockDim.x) q = fma(alpha[i], resid[i], q);
    for (int off = 16; off > 0; off >>= 1) {
        ld += __shfl_xor_sync(0xffffffffu, ld, off);
        q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = ld; red[1][threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) { a += red[0][w]; b += red[1][w]; }
        out2[0] = 2.0 * a;
        out2[1] = b;
    }
}

// resid[o][i] = (i < N) ? Y'[i][o] - mean_const : 0  (transposes the caller's (N, O) row-major Y);
// Y' = (Y - mean_o) / std_o when the fused output normaliser is on (ystats = [O means | O stds]), else Y.
__global__ void resid_kernel(const double* __restrict__ Y, int64_t N, int64_t Np, int O, double mean_const,
                             const double* __restrict__ ystats, double* __restrict__ resid) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Np * O) return;
    int64_t o = e / Np, i = e - o * Np;
    double v = 0.0;
    if (i < N) {
        v = Y[i * O + o];
        if (ystats) v = (v - ystats[o]) / ystats[O + o];
        v -= mean_const;
    }
    resid[e] = v;
}

// out[i][o] = in[o][i] for i < N
__global__ void untranspose_kernel(const double* __restrict__ in, int64_t N, int64_t Np, int O, double* __restrict__ out) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= N * O) return;
    int64_t i = e / O, o = e - i * O;
    out[e] = in[o * Np + i];
}

}  // namespace

void gpx_launch_fwd_step(TileMat L, const double* linv, int nt, int J, double* b, double* z, int64_t Np, int O,
                         int max_ctas, cudaStream_t st) {
    int work = nt - J - 1;
    int grid = work < 1 ? 1 : (work < max_ctas ? work : max_ctas);
    fwd_step_kernel<<<grid, 256, 0, st>>>(L, linv, nt, J, b, z, Np, O);
}

void gpx_launch_bwd_step(TileMat L, const double* linv, int I, double* z, double* a, int64_t Np, int O,
                         int compute_alpha, int ncols, int jbase, int jW, int jstride, int max_ctas, cudaStream_t st) {
    int grid = ncols < 1 ? 1 : (ncols < max_ctas ? ncols : max_ctas);
    bwd_step_kernel<<<grid, 256, 0, st>>>(L, linv, I, z, a, Np, O, compute_alpha, ncols, jbase, jW, jstride);
}

void gpx_launch_lml(const double* linv, const double* alpha, const double* resid, int64_t n_rows, int64_t Np, int O,
                    double* out2, cudaStream_t st) {
    lml_kernel<<<1, 1024, 0, st>>>(linv, alpha, resid, n_rows, Np, O, out2);
}

void gpx_launch_resid(const double* Y, int64_t N, int64_t Np, int O, double mean_const, const double* ystats, double* resid,
                      cudaStream_t st) {
    int64_t total = Np * O;
    resid_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(Y, N, Np, O, mean_const, ystats, resid);
}

void gpx_launch_untranspose(const double* in, int64_t N, int64_t Np, int O, double* out, cudaStream_t st) {
    int64_t total = N * O;
    if (total <= 0) return;
    untranspose_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>(in, N, Np, O, out);
}
