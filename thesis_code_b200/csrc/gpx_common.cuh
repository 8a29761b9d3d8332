// Shared definitions for the exact-GP kernels (sm_100a).
//
// DATA LAYOUT IN HBM
//   Every matrix on the hot path (the factor L of Ky, the inverted diagonal tiles, the transposed cross-kernel /
//   triangular-solve workspace Tt) is stored as 128x128 fp64 TILES, each tile contiguous (128 KB) in a
//   "chunked K-major" order:   element (r, c) of a tile lives at   (c >> 3) * 1024 + r * 8 + (c & 7).
//   i.e. 16 chunks of 8 columns; inside a chunk the 128 rows are contiguous with 8 doubles (64 B) each.
//   Consequences the kernels rely on:
//     * a K-slice of 16 columns of an operand tile (rows x 16 k) is ONE contiguous 16 KB block -> a single
//       1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) stages it into shared memory, no tensor map needed;
//     * in shared memory the [row][8] order makes the 128-bit fragment loads of the FP64 MMA bank-conflict free
//       without padding (row pitch 64 B: a quarter warp reads 128 contiguous bytes);
//     * the 8x8 accumulator fragment of mma.m8n8k4 is a contiguous 512 B block of the output tile, so a warp's
//       fragment load/store is one fully coalesced 512 B transaction.
//   The lower factor keeps only tiles (I, J) with I >= J, tile column J contiguous (rows J..nt-1), found through a
//   per-column offset table (so a block-column-cyclic shard is just a different table).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define GPX_TILE 128
#define GPX_TILE_ELEMS (128 * 128)
#define GPX_TILE_BYTES (128 * 128 * 8)

__host__ __device__ __forceinline__ int gpx_tile_idx(int r, int c) { return ((c >> 3) << 10) + (r << 3) + (c & 7); }

// A matrix of tiles.  Address of tile (r, c) = base + coloff[c] + (r - (packed ? c : 0)) * GPX_TILE_ELEMS.
struct TileMat {
    double* base;
    const int64_t* coloff;  // device table, in elements
    int packed;             // 1: lower-packed (first stored row tile of column c is c); 0: rectangular
};

__device__ __forceinline__ double* tile_ptr(const TileMat& m, int r, int c) {
    return m.base + m.coloff[c] + (int64_t)(r - (m.packed ? c : 0)) * GPX_TILE_ELEMS;
}

// Tiles written with ordinary (generic-proxy) stores are later read by 1-D TMA bulk copies (async proxy) of the NEXT
// kernel on the stream.  The kernel boundary already orders them (tools/proxy_test.cu measured 0 stale reads in 2 GB x
// 30 rounds), so only the cheap writer-side cross-proxy fence is kept as insurance; a __threadfence() here cost
// ~1.5 us per output tile in the GEMM epilogue (ncu: MEMBAR/ERRBAR stalls) and was removed.
__device__ __forceinline__ void gpx_publish_for_async_proxy() {
    asm volatile("fence.proxy.async.global;" ::: "memory");
}

// scaled distance from the Gram-trick r^2: GPy takes the square root first and, for an isotropic kernel, divides by
// the lengthscale afterwards (rdiv; 1.0 for ARD where the inputs were scaled before the Gram product).  The division is
// done Markstein-style with the correctly rounded reciprocal: q = r * (1/l), then one FMA residual correction -- the
// correctly rounded quotient (IEEE r / l) in 3 FP64 instructions instead of the ~25 of the generic division sequence,
// which matters because the K-build is bound by the FP64 ALU, not by HBM.
__device__ __forceinline__ double gpx_scaled_r(double r2, double rdiv) {
    const double r = sqrt(r2);
    if (rdiv == 1.0) return r;            // uniform branch (kernel argument)
    const double rinv = 1.0 / rdiv;       // loop-invariant: hoisted by the compiler
    const double q = r * rinv;
    return fma(fma(-q, rdiv, r), rinv, q);
}

// stationary kernels of the scaled distance r (GPy K_of_r; src/kernels.py:4-7 of the reference)
__device__ __forceinline__ double gpx_k_of_r(int kernel_id, double variance, double r) {
    if (kernel_id == 0) {  // RBF: GPy squares the square root again
        return variance * exp(-0.5 * (r * r));
    } else if (kernel_id == 1) {  // Matern 5/2
        const double s5 = 2.23606797749978969641;
        return variance * (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * exp(-s5 * r);
    } else if (kernel_id == 2) {  // Matern 3/2
        const double s3 = 1.73205080756887729353;
        return variance * (1.0 + s3 * r) * exp(-s3 * r);
    } else {  // Exponential
        return variance * exp(-r);
    }
}

// the same, selected at compile time (the tile builders are instantiated per kernel: no per-element dispatch)
template <int KID>
__device__ __forceinline__ double gpx_k_of_r_t(double variance, double r) {
    if (KID == 0) return variance * exp(-0.5 * (r * r));
    if (KID == 1) { const double s5 = 2.23606797749978969641; return variance * (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * exp(-s5 * r); }
    if (KID == 2) { const double s3 = 1.73205080756887729353; return variance * (1.0 + s3 * r) * exp(-s3 * r); }
    return variance * exp(-r);
}

// ---- launchers implemented in the .cu files (all asynchronous on `st`) -------------------------------------
struct GemmJob {
    TileMat A, B, C;
    int tri;           // 1: output tiles (I,J) with I >= J (lower triangle of a packed matrix); 0: rectangle
    int i0, i1;        // output tile rows  [i0, i1)
    // output tile columns: ordinal m in [0, ncols) maps to  J = jbase + (m / jW) * jstride + (m % jW)
    // (contiguous range: jW = GPX_JW_CONTIG; block-column-cyclic shard: jW = panel width, jstride = jW * world)
    int ncols, jbase, jW, jstride;
    int k0, k1;        // contraction tile range [k0, k1):  C(I,J) (-)= sum_k A(I,k) * B(J,k)^T
    int bdiag;         // 1: B tile index is (k, k) regardless of J (TRSM with an inverted diagonal tile)
    int overwrite;     // 1: C = +A*B^T ; 0: C -= A*B^T
    int chunk;         // in: cap on output-tile slots per CTA (0 = default 16); the launcher overwrites it with its choice
    int strip;         // output tile columns per rasterisation strip (0 = default, see tile_walk.cuh)
    int interleave;    // set by the launcher: 0 = a CTA owns `chunk` consecutive slots; R > 0 = the R resident CTAs of a round
                       // interleave over a window of R * chunk slots (L2 locality); in: < 0 forces the contiguous mode
    int rowwise;       // 1: "panel solve" mode -- one strip of all ncols columns, each CTA owns ONE half tile row (64 rows) and
                       // walks its columns in DESCENDING order with the contraction cut at the diagonal, k in [k0, min(k1, J+1)):
                       //   C(I, J) = sum_{k <= J} A(I, k) B(J, k)^T  written IN PLACE over A (column J is only read by columns >= J,
                       // which the same CTA has already finished).  B = inverse of the panel's triangular diagonal block.
};
#define GPX_JW_CONTIG (1 << 28)
__host__ __device__ __forceinline__ int gpx_job_col(const GemmJob& j, int m) {
    return j.jbase + (m / j.jW) * j.jstride + (m % j.jW);
}
// returns the number of (tile, k-tile) steps the launch performs (each is 2*128^3 flop); 0 if nothing was launched
long long gpx_launch_gemm(const GemmJob& job, int num_sms, cudaStream_t st, int64_t* launch_counter);
void gpx_gemm_init();  // sets the dynamic shared memory attribute

void gpx_launch_potrf_tile(double* tile, double* inv_tile, int* info, int col_index, cudaStream_t st);
void gpx_potrf_init();

// input loader: optional z-score (shift/scale per column, null = off), ARD lengthscale division, |x|^2 (numpy order)
void gpx_launch_prep_x(const double* X, int64_t N, int64_t Np, int d, const double* ls, int ard, const double* shift,
                       const double* scale, double* Xs, double* xsq, cudaStream_t st);
// stats[0..d) = column means, stats[d..2d) = population standard deviations of the (N, d) row-major array A
void gpx_launch_col_stats(const double* A, int64_t N, int d, double* stats, cudaStream_t st);
// K(X*, X) tiles for `nb` test tiles (grid) laid out with `nbt_layout` test tiles per tile column (+ fused mean partials)
void gpx_launch_kcross(TileMat Tt, int nb, int nbt_layout, int nt, const double* Xt, const double* xtsq,
                       int64_t Nt_valid, const double* Xs, const double* xsq, int64_t N, int d, int kernel_id,
                       double variance, double rdiv, const double* alpha, int O, int64_t Np, double* mean_partial,
                       cudaStream_t st, double* vec_out = nullptr, int vec_rows = 0);
