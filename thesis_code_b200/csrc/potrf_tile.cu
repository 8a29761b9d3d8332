// Diagonal-tile kernel of the blocked Cholesky: factor one 128x128 fp64 tile in shared memory (lower Cholesky,
// LAPACK dpotrf semantics) and invert the triangular factor, so that the panel TRSM, the alpha solves and the
// predictive-variance solve all become DMMA GEMMs / GEMVs against the inverted tile.
//
// One CTA, 512 threads, the tile held in shared memory as S[128][129] (odd pitch = conflict-free column walks).
// Both sweeps are "right-looking with deferred scaling": one __syncthreads per column.
//   Cholesky : after step j the trailing block holds  S[r][c] -= S[r][j] * S[c][j] / S[j][j];  L[r][j] = S[r][j] / sqrt(S[j][j])
//   inverse  : U (stored transposed in the strict upper triangle) accumulates  U[i][c] -= L[i][k] * U[k][c] / L[k][k]
// info: first non-positive (or NaN) pivot, reported LAPACK-style as a 1-based global row index.
#include "gpx_common.cuh"

namespace {
constexpr int kPitch = 129;
constexpr int kThreadsP = 512;
constexpr int kSmemP = GPX_TILE * kPitch * 8 + 64;

__global__ void __launch_bounds__(kThreadsP, 1) potrf_tile_kernel(double* __restrict__ tile, double* __restrict__ inv_tile,
                                                                  int* __restrict__ info, int col_index) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* S = reinterpret_cast<double*>(smem_raw);
    const int tid = threadIdx.x;

    // load (chunked layout -> S[r][c]); coalesced 16 B global reads
    for (int e = tid * 2; e < GPX_TILE_ELEMS; e += kThreadsP * 2) {
        double2 v = *reinterpret_cast<const double2*>(tile + e);
        int chunk = e >> 10, r = (e >> 3) & 127, c = chunk * 8 + (e & 7);
        S[r * kPitch + c] = v.x;
        S[r * kPitch + c + 1] = v.y;
    }
    __syncthreads();

    const int r = tid & 127, q = tid >> 7;  // row owner, column phase (4 phases)
    // ---- Cholesky, deferred scaling -------------------------------------------------------------------------
    for (int j = 0; j < GPX_TILE; j++) {
        double d = S[j * kPitch + j];
        if (!(d > 0.0)) {
            if (tid == 0) atomicCAS(info, 0, col_index * GPX_TILE + j + 1);
            d = 1.0;
        }
        if (r > j) {
            double f = S[r * kPitch + j] * (1.0 / d);
            for (int c = j + 1 + q; c <= r; c += 4) S[r * kPitch + c] -= f * S[c * kPitch + j];
        }
        __syncthreads();
    }
    // scale columns:  L[r][j] = S[r][j] / sqrt(S[j][j]) (r > j), L[j][j] = sqrt(S[j][j]).
    // each thread first reads the pivots it needs, then everyone writes.
    {
        double piv[32];
#pragma unroll
        for (int u = 0; u < 32; u++) {
            int c = q + 4 * u;
            double d = S[c * kPitch + c];
            piv[u] = (d > 0.0) ? sqrt(d) : 1.0;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 32; u++) {
            int c = q + 4 * u;
            if (r > c) S[r * kPitch + c] = S[r * kPitch + c] / piv[u];
            else if (r == c) S[r * kPitch + c] = piv[u];
        }
    }
    __syncthreads();
    // write L (zero above the diagonal) back in the chunked layout
    for (int e = tid * 2; e < GPX_TILE_ELEMS; e += kThreadsP * 2) {
        int chunk = e >> 10, rr = (e >> 3) & 127, c = chunk * 8 + (e & 7);
        double2 v;
        v.x = (rr >= c) ? S[rr * kPitch + c] : 0.0;
        v.y = (rr >= c + 1) ? S[rr * kPitch + c + 1] : 0.0;
        *reinterpret_cast<double2*>(tile + e) = v;
    }
    // ---- inverse of L: U^T in the strict upper triangle, U[i][i] = 1 implied ----------------------------------
    // init: U[i][c] = 0 for c < i  ->  S[c][i] = 0 for c < i
    for (int e = tid; e < GPX_TILE * GPX_TILE; e += kThreadsP) {
        int a = e >> 7, b = e & 127;
        if (a < b) S[a * kPitch + b] = 0.0;
    }
    __syncthreads();
    // step k: for i > k, c <= k:   U[i][c] -= L[i][k] * (U[k][c] / L[k][k])      (U[k][k] = 1)
    // thread owns row i = r; columns c = q, q+4, ...
    for (int k = 0; k < GPX_TILE - 1; k++) {
        if (r > k) {
            double f = S[r * kPitch + k] / S[k * kPitch + k];
            for (int c = q; c <= k; c += 4) {
                double ukc = (c == k) ? 1.0 : S[c * kPitch + k];
                S[c * kPitch + r] -= f * ukc;
            }
        }
        __syncthreads();
    }
    // Linv[i][c] = U[i][c] / L[i][i]; write in the chunked layout with zeros above the diagonal
    for (int e = tid * 2; e < GPX_TILE_ELEMS; e += kThreadsP * 2) {
        int chunk = e >> 10, i = (e >> 3) & 127, c = chunk * 8 + (e & 7);
        double lii = S[i * kPitch + i];
        double2 v;
        v.x = (i > c) ? S[c * kPitch + i] / lii : ((i == c) ? 1.0 / lii : 0.0);
        v.y = (i > c + 1) ? S[(c + 1) * kPitch + i] / lii : ((i == c + 1) ? 1.0 / lii : 0.0);
        *reinterpret_cast<double2*>(inv_tile + e) = v;
    }
    gpx_publish_for_async_proxy();
}
}  // namespace

void gpx_potrf_init() {
    cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemP);
}

void gpx_launch_potrf_tile(double* tile, double* inv_tile, int* info, int col_index, cudaStream_t st) {
    potrf_tile_kernel<<<1, kThreadsP, kSmemP, st>>>(tile, inv_tile, info, col_index);
}
