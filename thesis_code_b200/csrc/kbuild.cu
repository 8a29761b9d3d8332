// Kernel-matrix builders: pairwise stationary kernels (RBF / Matern-5/2 / Matern-3/2 / Exponential, isotropic or
// ARD) from a Gram product plus a fused distance -> exp/Matern epilogue, written straight into 128x128 tiles of
// the factor storage (K(X,X) + diag_add*I, lower triangle only) or of the transposed cross-kernel workspace
// (K(X*,X), fused with the predictive-mean GEMV partial sums).
//
// Arithmetic follows GPy's Stationary._unscaled_dist / _scaled_dist (the library the reference's
// src/kernels.py:4-7 aliases) OPERATION BY OPERATION:
//   ARD:        inputs divided by the lengthscales first, r^2 = |a|^2 + |b|^2 - 2 a.b clipped at 0, r = sqrt(r^2);
//   isotropic:  the same Gram trick on the UNSCALED inputs, then r = sqrt(r^2) / lengthscale
// (exact 0 on the diagonal of K(X,X)), then K_of_r.  The order matters: with |x / l| ~ 2000 (BASELINE config C4) scaling
// first moves r^2 by ~1e-9 and the predictive mean by ~7e-7; with GPy's order r^2 reproduces numpy bit for bit at d = 1.
// |x|^2 is accumulated like numpy's pairwise np.sum(np.square(X), 1) (no FMA contraction, 8 partial sums for d >= 8).
//
// Roofline: the K(X,X) build writes 8 B per element once (HBM-write bound: 8*N(N+1)/2 + 8*N*d bytes) but at small
// d it is bound by the fp64 exp/sqrt ALU work (~60 DFMA-equivalents per element); both rates are reported.
#include "gpx_common.cuh"

namespace {

#ifndef GPX_KBUILD_MINBLOCKS
#define GPX_KBUILD_MINBLOCKS 3
#endif
constexpr int kMaxD = 64;
constexpr int kMaxFusedO = 4;   // outputs whose mean partial sums are fused into one K*-tile pass (more outputs: more passes)

// numpy's pairwise summation (pairwise_sum_DOUBLE) of the squares of a[0..n): sequential below 8 terms, 8 running sums
// combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) plus a sequential remainder up to 128 terms, recursive halving above.
// No FMA contraction (np.square then np.sum).  Verified term by term against numpy for d = 1..300.
__device__ double gpx_numpy_sumsq(const double* a, int n) {
    if (n < 8) {
        double s = -0.0;
        for (int i = 0; i < n; i++) s = __dadd_rn(s, __dmul_rn(a[i], a[i]));
        return s;
    }
    if (n <= 128) {
        double r[8];
#pragma unroll
        for (int u = 0; u < 8; u++) r[u] = __dmul_rn(a[u], a[u]);
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int u = 0; u < 8; u++) r[u] = __dadd_rn(r[u], __dmul_rn(a[i + u], a[i + u]));
        }
        double s = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                             __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
        for (; i < n; i++) s = __dadd_rn(s, __dmul_rn(a[i], a[i]));
        return s;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    return __dadd_rn(gpx_numpy_sumsq(a, n2), gpx_numpy_sumsq(a + n2, n - n2));
}

// Loader of the inputs (training or test):  v = X[i][k];  fused NormalizerModel z-score v = (v - shift_k) / scale_k
// (reference src/models/normalizer.py:36-48);  ARD: v /= lengthscale_k.  Rows i >= N are zero padding.
// xsq[i] = sum_k v^2 in numpy's summation order (np.sum(np.square(X), 1)) so that r^2 matches the reference's rounding.
__global__ void prep_x_kernel(const double* __restrict__ X, int64_t N, int64_t Np, int d,
                              const double* __restrict__ ls, int ard, const double* __restrict__ shift,
                              const double* __restrict__ scale, double* __restrict__ Xs, double* __restrict__ xsq) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Np) return;
    for (int k = 0; k < d; k++) {
        double v = 0.0;
        if (i < N) {
            v = X[i * d + k];
            if (shift) v = (v - shift[k]) / scale[k];
            if (ard) v = v / ls[k];
        }
        Xs[i * d + k] = v;
    }
    xsq[i] = gpx_numpy_sumsq(Xs + i * d, d);
}

// Column statistics for the fused normaliser: stats[k] = mean_k, stats[d + k] = population std_k (np.std, ddof 0) of the
// (N, d) row-major array A.  One CTA per column, two passes, fixed-order tree reductions (deterministic).
__global__ void __launch_bounds__(256) col_stats_kernel(const double* __restrict__ A, int64_t N, int d, double* __restrict__ stats) {
    __shared__ double red[256];
    __shared__ double s_mean;
    const int k = blockIdx.x, tid = threadIdx.x;
    double s = 0.0;
    for (int64_t i = tid; i < N; i += 256) s += A[i * d + k];
    red[tid] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) { if (tid < off) red[tid] += red[tid + off]; __syncthreads(); }
    if (tid == 0) s_mean = red[0] / (double)N;
    __syncthreads();
    const double m = s_mean;
    s = 0.0;
    for (int64_t i = tid; i < N; i += 256) { const double v = A[i * d + k] - m; s = fma(v, v, s); }
    __syncthreads();
    red[tid] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) { if (tid < off) red[tid] += red[tid + off]; __syncthreads(); }
    if (tid == 0) { stats[k] = m; stats[d + k] = sqrt(red[0] / (double)N); }
}

constexpr int kTmaMaxD = 16;   // input dimensions up to this are staged with TMA bulk copies (all five BASELINE configs)

__device__ __forceinline__ uint32_t kb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One CTA (256 threads) per output tile.  Rows come from (Xa, asq) tile ti, columns from (Xb, bsq) tile tj.
//  * staging: the 128 x d block of input rows of a tile is contiguous in the row-major inputs, so for d <= kTmaMaxD one
//    elected thread fetches each operand's block with ONE 1-D TMA bulk copy (cp.async.bulk, mbarrier complete_tx) and the
//    CTA re-lays it k-major in shared memory; wider inputs are gathered with coalesced loads in chunks of kMaxD columns.
//  * Gram product: k-major operands make the two columns a thread owns adjacent -> one 128-bit shared load feeds two FMAs,
//    the row operand is a conflict-free 64-bit load (8 distinct rows per warp).
//  * epilogue: r^2 -> sqrt -> (/ lengthscale) -> K_of_r, compiled per kernel (KID) so no per-element dispatch.
// Thread mapping = the DMMA accumulator fragment pattern: lane (g = lane>>2, t = lane&3) of warp w owns, for each
// of its 2 row blocks and 16 column chunks, the pair (r, c..c+1) -> one 16 B store; a warp stores 512 contiguous B.
template <bool kSymmetric, bool kFused, int KID>
__device__ __forceinline__ void build_tile(double* __restrict__ out, const double* __restrict__ Xa,
                                           const double* __restrict__ asq, int64_t Na, int ti,
                                           const double* __restrict__ Xb, const double* __restrict__ bsq,
                                           int64_t Nb, int tj, int d, double variance,
                                           double rdiv, double diag_add, double* sm, int pass,
                                           const double* __restrict__ alpha_tile, int O, int64_t alpha_stride,
                                           double* __restrict__ red, double* __restrict__ vec_row0, int64_t vec_stride,
                                           int vec_rows) {
    // the input dimension is processed in chunks of at most kMaxD columns (any d is supported)
    const int dcap = d < kMaxD ? d : kMaxD;
    const int nchunks = (d + kMaxD - 1) / kMaxD;
    double* sat = sm;                          // [dcap][128]  k-major row operand
    double* sbt = sat + GPX_TILE * dcap;       // [dcap][128]  k-major column operand
    double* sasq = sbt + GPX_TILE * dcap;      // [128]
    double* sbsq = sasq + GPX_TILE;            // [128]
    double* stage = sbsq + GPX_TILE;           // [2][128][d]  TMA landing zone (d <= kTmaMaxD only), then the mbarrier
    const int tid = threadIdx.x;
    const bool use_tma = d <= kTmaMaxD;
    if (use_tma) {
        uint64_t* bar = reinterpret_cast<uint64_t*>(stage + 2 * GPX_TILE * kTmaMaxD);
        const uint32_t bytes = (uint32_t)(GPX_TILE * d * sizeof(double));
        if (tid == 0) {
            if (pass == 0) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(kb_smem_u32(bar)) : "memory");
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(kb_smem_u32(bar)), "r"(2 * bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(kb_smem_u32(stage)),
                         "l"(Xa + (int64_t)ti * GPX_TILE * d), "r"(bytes), "r"(kb_smem_u32(bar)) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(kb_smem_u32(stage + GPX_TILE * d)),
                         "l"(Xb + (int64_t)tj * GPX_TILE * d), "r"(bytes), "r"(kb_smem_u32(bar)) : "memory");
        }
        if (tid < GPX_TILE) {
            sasq[tid] = asq[(int64_t)ti * GPX_TILE + tid];
            sbsq[tid] = bsq[(int64_t)tj * GPX_TILE + tid];
        }
        __syncthreads();   // the barrier is initialised before anybody polls it
        {
            uint32_t ok = 0;
            const uint32_t parity = (uint32_t)(pass & 1);
            while (!ok) {
                asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                             : "=r"(ok) : "r"(kb_smem_u32(bar)), "r"(parity) : "memory");
            }
        }
        for (int e = tid; e < GPX_TILE * d; e += blockDim.x) {   // [row][d] -> [k][row]
            const int rr = e / d, k = e - rr * d;
            sat[k * GPX_TILE + rr] = stage[e];
            sbt[k * GPX_TILE + rr] = stage[GPX_TILE * d + e];
        }
        __syncthreads();
    } else if (tid < GPX_TILE) {
        sasq[tid] = asq[(int64_t)ti * GPX_TILE + tid];
        sbsq[tid] = bsq[(int64_t)tj * GPX_TILE + tid];
    }
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const bool dtile = kSymmetric && ti == tj;
    const int64_t cv = Nb - (int64_t)tj * GPX_TILE;
    const int ncv = cv > GPX_TILE ? GPX_TILE : (int)(cv < 0 ? 0 : cv);   // valid columns of this tile
#pragma unroll
    for (int rb = 0; rb < 2; rb++) {
        const int r = warp * 16 + rb * 8 + g;
        const bool row_ok = (int64_t)ti * GPX_TILE + r < Na;
        double msum[kMaxFusedO];
#pragma unroll
        for (int o = 0; o < kMaxFusedO; o++) msum[o] = 0.0;
        // the 16 column chunks are taken in two halves of 8 (32 accumulators live instead of 64: no spills at 128 registers)
#pragma unroll 1
        for (int hf = 0; hf < 2; hf++) {
            double dot[8][2];
#pragma unroll
            for (int q = 0; q < 8; q++) { dot[q][0] = 0.0; dot[q][1] = 0.0; }
            for (int ci = 0; ci < nchunks; ci++) {
                const int d0 = ci * kMaxD, dc = (d - d0 < kMaxD) ? (d - d0) : kMaxD;
                if (!use_tma && ((rb == 0 && hf == 0) || nchunks > 1)) {   // a single chunk stays resident for all four passes
                    __syncthreads();
                    for (int e = tid; e < GPX_TILE * dc; e += blockDim.x) {
                        const int rr = e / dc, k = e - rr * dc;
                        sat[k * GPX_TILE + rr] = Xa[((int64_t)ti * GPX_TILE + rr) * d + d0 + k];
                        sbt[k * GPX_TILE + rr] = Xb[((int64_t)tj * GPX_TILE + rr) * d + d0 + k];
                    }
                    __syncthreads();
                }
                for (int k = 0; k < dc; k++) {
                    const double a = sat[k * GPX_TILE + r];
                    const double2* bp = reinterpret_cast<const double2*>(sbt + k * GPX_TILE + hf * 64 + 2 * t);
#pragma unroll
                    for (int q = 0; q < 8; q++) {
                        const double2 bv = bp[q * 4];
                        dot[q][0] = fma(a, bv.x, dot[q][0]);
                        dot[q][1] = fma(a, bv.y, dot[q][1]);
                    }
                }
            }
            const double ra = sasq[r];   // (written before the block-wide barriers of the staging above)
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const int kc = hf * 8 + q;
                double v[2];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int c = kc * 8 + 2 * t + h;
                    double r2 = -2.0 * dot[q][h] + (ra + sbsq[c]);
                    r2 = fmax(r2, 0.0);
                    const bool diag = dtile && r == c;
                    if (kSymmetric && diag) r2 = 0.0;
                    double kv = gpx_k_of_r_t<KID>(variance, gpx_scaled_r(r2, rdiv));
                    if (kSymmetric) {
                        if (diag) kv += diag_add;
                        if (!row_ok || c >= ncv) kv = diag ? 1.0 : 0.0;  // identity padding
                    } else {
                        if (!row_ok || c >= ncv) kv = 0.0;
                    }
                    v[h] = kv;
                }
                if (out) *reinterpret_cast<double2*>(out + kc * 1024 + r * 8 + 2 * t) = make_double2(v[0], v[1]);
                if (kFused && vec_row0 && r < vec_rows)   // latency path: k(x*_r, X) as right-hand-side vectors for the GEMV-style solve
                    *reinterpret_cast<double2*>(vec_row0 + (int64_t)r * vec_stride + kc * 8 + 2 * t) = make_double2(v[0], v[1]);
                if (kFused && alpha_tile) {               // fused predictive-mean partial sums over this thread's columns
                    const int c = kc * 8 + 2 * t;
                    for (int o = 0; o < O && o < kMaxFusedO; o++) {
                        const double* al = alpha_tile + (int64_t)o * alpha_stride;
                        msum[o] = fma(v[0], al[c], msum[o]);
                        msum[o] = fma(v[1], al[c + 1], msum[o]);
                    }
                }
            }
        }
        if (kFused && alpha_tile)
            for (int o = 0; o < O && o < kMaxFusedO; o++) red[(o * 4 + t) * GPX_TILE + r] = msum[o];
    }
    if (out) gpx_publish_for_async_proxy();
}

// K(X,X) + diag_add I into the packed lower factor storage: grid (nt, number of tile columns handled here).
template <int KID>
__global__ void __launch_bounds__(256, GPX_KBUILD_MINBLOCKS) kbuild_lower_kernel(TileMat L, int nt, int jfirst, int jstep, int i_first,
                                                           const double* __restrict__ Xs,
                                                           const double* __restrict__ xsq, int64_t N, int d,
                                                           double variance, double rdiv, double diag_add) {
    extern __shared__ __align__(16) double sm[];
    const int J = jfirst + blockIdx.y * jstep;
    const int I = i_first + blockIdx.x;
    if (J >= nt || I >= nt || I < J) return;
    build_tile<true, false, KID>(tile_ptr(L, I, J), Xs, xsq, N, I, Xs, xsq, N, J, d, variance, rdiv, diag_add, sm, 0, nullptr, 0, 0,
                                 nullptr, nullptr, 0, 0);
}

// K(X*, X) tiles (rows = test points, cols = training points) into Tt (if Tt.base != null) and the fused
// predictive-mean partial sums  mean_partial[J][n][o] = sum_{j in tile J} K(x*_n, x_j) alpha[o][j].
// grid (nbt, nt).
template <int KID>
__global__ void __launch_bounds__(256, 2) kcross_kernel(TileMat Tt, int nbt, int nt, const double* __restrict__ Xt,
                                                     const double* __restrict__ xtsq, int64_t Nt_valid,
                                                     const double* __restrict__ Xs, const double* __restrict__ xsq,
                                                     int64_t N, int d, double variance, double rdiv,
                                                     const double* __restrict__ alpha, int O, int64_t Np,
                                                     double* __restrict__ mean_partial, double* __restrict__ vec_out,
                                                     int vec_rows) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[kMaxFusedO * 4 * GPX_TILE];  // [o][t][row]
    const int n = blockIdx.x, J = blockIdx.y;
    const int tid = threadIdx.x;
    double* out = Tt.base ? tile_ptr(Tt, n, J) : nullptr;
    // The mean partial sums are accumulated inside the tile builder's epilogue (no per-thread copy of the tile: 239 -> ~128
    // registers, two CTAs per SM).  Outputs are handled kMaxFusedO at a time; passes after the first only recompute the sums.
    for (int o0 = 0; o0 < (mean_partial ? O : 1); o0 += kMaxFusedO) {
        const int Oc = mean_partial ? ((O - o0 < kMaxFusedO) ? (O - o0) : kMaxFusedO) : 0;
        const bool first = (o0 == 0);
        if (!first) __syncthreads();
        build_tile<false, true, KID>(first ? out : nullptr, Xt, xtsq, Nt_valid, n, Xs, xsq, N, J, d, variance, rdiv, 0.0, sm, o0 / kMaxFusedO,
                          mean_partial ? alpha + (int64_t)o0 * Np + (int64_t)J * GPX_TILE : nullptr, Oc, Np, red,
                          (first && vec_out && n == 0) ? vec_out + (int64_t)J * GPX_TILE : nullptr, Np, vec_rows);
        if (!mean_partial) return;
        __syncthreads();
        for (int e = tid; e < GPX_TILE * Oc; e += blockDim.x) {
            const int o = e / GPX_TILE, r = e - o * GPX_TILE;
            const double* rr = red + (o * 4) * GPX_TILE + r;
            const double s = (rr[0] + rr[GPX_TILE]) + (rr[2 * GPX_TILE] + rr[3 * GPX_TILE]);
            mean_partial[((int64_t)J * nbt + n) * GPX_TILE * O + (int64_t)r * O + o0 + o] = s;
        }
    }
}

// mean[n*128 + r][o] = mean_const + sum_J mean_partial[J][n][r][o]   (fixed summation order: deterministic)
// Fused output de-normalisation (reference normalizer.py:102-114): mean * std_o + mean_o.
__global__ void mean_reduce_kernel(const double* __restrict__ mean_partial, int nbt, int nt, int O, double mean_const,
                                   const double* __restrict__ ystats, double* __restrict__ mean, int64_t row0, int64_t Ns) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // index over nbt*128*O
    int64_t total = (int64_t)nbt * GPX_TILE * O;
    if (e >= total) return;
    double s = 0.0;
    for (int J = 0; J < nt; J++) s += mean_partial[(int64_t)J * total + e];
    int64_t row = row0 + e / O;
    const int o = (int)(e % O);
    double m = s + mean_const;
    if (ystats) m = m * ystats[O + o] + ystats[o];
    if (row < Ns) mean[row * O + o] = m;
}

// var[n] = variance + noise - sum_j Tt[n][j]^2 : one CTA per test tile, 256 threads = (row, half of the chunks)
// Row sums of squares of Tt, phase 1: grid (nb, S); CTA (n, s) covers the tile columns [s nt / S, (s+1) nt / S) of test
// tile n and writes 128 partial sums.  S is chosen so that nb * S CTAs fill the GPU (one CTA per test tile left most SMs
// idle: 650 GB/s at 20 test tiles, profiles/r02_hbm_kernels_ncu_summary.txt).
__global__ void __launch_bounds__(256) var_partial_kernel(TileMat Tt, int nt, int S, double* __restrict__ partial) {
    __shared__ double red[GPX_TILE];
    const int n = blockIdx.x, sidx = blockIdx.y, tid = threadIdx.x, r = tid & 127, half = tid >> 7;
    const int J0 = (int)((long long)nt * sidx / S), J1 = (int)((long long)nt * (sidx + 1) / S);
    double s0 = 0.0, s1 = 0.0;
    for (int J = J0; J < J1; J++) {
        const double* tp = tile_ptr(Tt, n, J);
#pragma unroll
        for (int kc = 0; kc < 8; kc++) {
            const double2* p = reinterpret_cast<const double2*>(tp + (half * 8 + kc) * 1024 + r * 8);
            double2 a = p[0], b = p[1], c = p[2], e = p[3];
            s0 = fma(a.x, a.x, s0); s1 = fma(a.y, a.y, s1);
            s0 = fma(b.x, b.x, s0); s1 = fma(b.y, b.y, s1);
            s0 = fma(c.x, c.x, s0); s1 = fma(c.y, c.y, s1);
            s0 = fma(e.x, e.x, s0); s1 = fma(e.y, e.y, s1);
        }
    }
    double s = s0 + s1;
    if (half == 1) red[r] = s;
    __syncthreads();
    if (half == 0) partial[((int64_t)n * S + sidx) * GPX_TILE + r] = s + red[r];
}

// phase 2: var[row][*] = prior_var - sum_s partial[n][s][r]  (fixed order: deterministic).
// var_cols = 1: one shared variance per test row (GPy); var_cols = O: (Ns, O), each column scaled by std_o^2 when the
// fused output normaliser is on (reference denormalize_variance) or plainly replicated when it is off.
__global__ void var_final_kernel(const double* __restrict__ partial, int nb, int S, double prior_var,
                                 const double* __restrict__ ystats, int var_cols, double* __restrict__ var, int64_t row0,
                                 int64_t Ns) {
    const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (int64_t)nb * GPX_TILE) return;
    const int n = (int)(e >> 7), r = (int)(e & 127);
    double s = 0.0;
    for (int k = 0; k < S; k++) s += partial[((int64_t)n * S + k) * GPX_TILE + r];
    const int64_t row = row0 + e;
    if (row < Ns) {
        const double v = prior_var - s;
        for (int o = 0; o < var_cols; o++) {
            const double sd = ystats ? ystats[var_cols + o] : 1.0;   // ystats = [O means | O stds]; var_cols == O here
            var[row * var_cols + o] = ystats ? v * (sd * sd) : v;
        }
    }
}

// dense row-major view of a tiled matrix (tests / export / small full covariance)
__global__ void detile_kernel(TileMat M, int lower, int64_t rows, int64_t cols, double* __restrict__ out) {
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r = blockIdx.y;
    if (c >= cols || r >= rows) return;
    // lower = 0: rectangular; 1: lower triangle, zeros above; 2: symmetric (mirror the stored lower triangle)
    double v = 0.0;
    int64_t rr = r, cc = c;
    if (lower == 2 && r < c) { rr = c; cc = r; }
    if (!lower || rr >= cc) v = tile_ptr(M, (int)(rr >> 7), (int)(cc >> 7))[gpx_tile_idx((int)(rr & 127), (int)(cc & 127))];
    out[r * cols + c] = v;
}

}  // namespace

static size_t kb_smem(int d) {
    const int dc = d < kMaxD ? d : kMaxD;
    size_t doubles = (size_t)2 * GPX_TILE * dc + 2 * GPX_TILE;                 // k-major operands + squared norms
    if (d <= kTmaMaxD) doubles += (size_t)2 * GPX_TILE * kTmaMaxD + 2;         // TMA landing zone + mbarrier
    return doubles * sizeof(double);
}

// ---- latency path, <= 8 test rows: loader + k(x*, X) + predictive mean in ONE kernel ----------------------------------------
// The tile builder evaluates 128 x 128 entries per CTA whatever the number of valid test rows, and loader / K* / reduce are
// three dependent launches; for the row-by-row callers (src/growth_model_GPR/ipopt_wrapper.py:55) this kernel does the same
// arithmetic (prep_x's z-score / ARD scaling / numpy-order |x|^2, the Gram-trick r^2 with the same FMA chain, K_of_r) for just
// the ns rows: grid nt CTAs x 128 threads, one training row per thread.  It writes k(x*_s, X) as right-hand-side vectors (for the
// variance) and per-tile partial sums of k . alpha; the CTA that arrives last (self-resetting ticket) adds them in tile order.
constexpr int kSmallRows = 8, kSmallMaxO = 8, kSmallMaxD = 256;
template <int KID>
__global__ void __launch_bounds__(128) small_kstar_kernel(const double* __restrict__ Xin, int ns, int d, const double* __restrict__ ls,
                                                          int ard, const double* __restrict__ shift, const double* __restrict__ scale,
                                                          const double* __restrict__ Xs, const double* __restrict__ xsq, int64_t N,
                                                          double variance, double rdiv, const double* __restrict__ alpha, int O,
                                                          int64_t Np, double mean_const, const double* __restrict__ ystats,
                                                          double* __restrict__ vec_out, double* __restrict__ partial,
                                                          unsigned int* __restrict__ ticket, double* __restrict__ mean) {
    extern __shared__ __align__(16) double sm[];
    double* xt = sm;                          // [ns][d]
    double* xtsq = xt + kSmallRows * d;       // [8]
    double* red = xtsq + kSmallRows;          // [4 warps][8 * 8]
    __shared__ bool last;
    const int J = blockIdx.x, nt = gridDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t i = (int64_t)J * GPX_TILE + tid;
    for (int e = tid; e < ns * d; e += GPX_TILE) {
        const int k = e % d;
        double v = Xin[e];
        if (shift) v = (v - shift[k]) / scale[k];
        if (ard) v = v / ls[k];
        xt[e] = v;
    }
    __syncthreads();
    if (tid < ns) xtsq[tid] = gpx_numpy_sumsq(xt + tid * d, d);
    __syncthreads();
    double dot[kSmallRows];
#pragma unroll
    for (int q = 0; q < kSmallRows; q++) dot[q] = 0.0;
    const double* xi = Xs + i * d;
    for (int k = 0; k < d; k++) {
        const double b = xi[k];
#pragma unroll
        for (int q = 0; q < kSmallRows; q++)
            if (q < ns) dot[q] = fma(xt[q * d + k], b, dot[q]);
    }
    const double bsq = xsq[i];
    double kv[kSmallRows];
#pragma unroll
    for (int q = 0; q < kSmallRows; q++) {
        kv[q] = 0.0;
        if (q < ns) {
            const double r2 = fmax(-2.0 * dot[q] + (xtsq[q] + bsq), 0.0);
            if (i < N) kv[q] = gpx_k_of_r_t<KID>(variance, gpx_scaled_r(r2, rdiv));
            if (vec_out) vec_out[(int64_t)q * Np + i] = kv[q];
        }
    }
    // per-tile partial sums of k . alpha: warp shuffle tree, then the four warps in order
    for (int o = 0; o < O; o++) {
        const double a = alpha[(int64_t)o * Np + i];
#pragma unroll
        for (int q = 0; q < kSmallRows; q++) {
            if (q < ns) {
                double v = kv[q] * a;
                for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                if (lane == 0) red[warp * 64 + q * kSmallMaxO + o] = v;
            }
        }
    }
    __syncthreads();
    if (tid < ns * O) {
        const int q = tid / O, o = tid - q * O, e = q * kSmallMaxO + o;
        partial[(int64_t)J * 64 + e] = (red[e] + red[64 + e]) + (red[128 + e] + red[192 + e]);
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) last = atomicAdd(ticket, 1u) == (unsigned)(nt - 1);
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (tid < ns * O) {
        const int q = tid / O, o = tid - q * O, e = q * kSmallMaxO + o;
        double m = 0.0;
        for (int j = 0; j < nt; j++) m += __ldcg(&partial[(int64_t)j * 64 + e]);
        m += mean_const;
        if (ystats) m = m * ystats[O + o] + ystats[o];
        mean[q * O + o] = m;
    }
    if (tid == 0) *ticket = 0u;
}

bool gpx_small_kstar_ok(int64_t ns, int d, int O) { return ns >= 1 && ns <= kSmallRows && O <= kSmallMaxO && d <= kSmallMaxD; }
// partial: nt * 64 doubles; ticket: one zero-initialised counter
void gpx_launch_small_kstar(const double* Xin, int ns, int d, const double* ls, int ard, const double* shift, const double* scale,
                            const double* Xs, const double* xsq, int64_t N, int nt, int kernel_id, double variance, double rdiv,
                            const double* alpha, int O, int64_t Np, double mean_const, const double* ystats, double* vec_out,
                            double* partial, unsigned int* ticket, double* mean, cudaStream_t st) {
    const size_t smem = (size_t)(kSmallRows * d + kSmallRows + 4 * 64) * sizeof(double);
    switch (kernel_id) {
#define GPX_SMALLK(K) case K: small_kstar_kernel<K><<<nt, GPX_TILE, smem, st>>>(Xin, ns, d, ls, ard, shift, scale, Xs, xsq, N, variance, rdiv, \
                                  alpha, O, Np, mean_const, ystats, vec_out, partial, ticket, mean); break;
        GPX_SMALLK(0) GPX_SMALLK(1) GPX_SMALLK(2) GPX_SMALLK(3)
#undef GPX_SMALLK
    }
}

int gpx_kbuild_max_d() { return 1 << 20; }   // any input dimension (processed in chunks of kMaxD)
int gpx_grad_max_d() { return kMaxD; }       // the gradient kernel keeps one partial sum per lengthscale in registers

static void gpx_grad_init_attr();   // defined after the gradient kernel below

template <int KID>
static void kb_set_attr(int bytes) {
    cudaFuncSetAttribute(kbuild_lower_kernel<KID>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(kcross_kernel<KID>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

// per device (function attributes are per device): called from gpx_create after cudaSetDevice
void gpx_kbuild_init() {
    gpx_grad_init_attr();
    size_t a = kb_smem(kMaxD), b = kb_smem(kTmaMaxD);
    const int s = (int)(a > b ? a : b);
    kb_set_attr<0>(s); kb_set_attr<1>(s); kb_set_attr<2>(s); kb_set_attr<3>(s);
}

static void kb_launch_lower(int kernel_id, dim3 grid, size_t smem, cudaStream_t st, TileMat L, int nt, int jfirst, int jstep, int i_first,
                            const double* Xs, const double* xsq, int64_t N, int d, double variance, double rdiv, double diag_add) {
    switch (kernel_id) {
#define GPX_KLOWER(K) case K: kbuild_lower_kernel<K><<<grid, 256, smem, st>>>(L, nt, jfirst, jstep, i_first, Xs, xsq, N, d, variance, rdiv, diag_add); break;
        GPX_KLOWER(0) GPX_KLOWER(1) GPX_KLOWER(2) GPX_KLOWER(3)
#undef GPX_KLOWER
    }
}

void gpx_launch_prep_x(const double* X, int64_t N, int64_t Np, int d, const double* ls, int ard, const double* shift,
                       const double* scale, double* Xs, double* xsq, cudaStream_t st) {
    int threads = 128;
    int blocks = (int)((Np + threads - 1) / threads);
    prep_x_kernel<<<blocks, threads, 0, st>>>(X, N, Np, d, ls, ard, shift, scale, Xs, xsq);
}

void gpx_launch_col_stats(const double* A, int64_t N, int d, double* stats, cudaStream_t st) {
    col_stats_kernel<<<d, 256, 0, st>>>(A, N, d, stats);
}

void gpx_launch_kbuild_lower(TileMat L, int nt, int jfirst, int jstep, int ncols, const double* Xs,
                             const double* xsq, int64_t N, int d, int kernel_id, double variance, double rdiv,
                             double diag_add, cudaStream_t st) {
    if (ncols <= 0) return;
    dim3 grid(nt, ncols);
    kb_launch_lower(kernel_id, grid, kb_smem(d), st, L, nt, jfirst, jstep, 0, Xs, xsq, N, d, variance, rdiv, diag_add);
}

// the tile rows [i_first, nt) of K(X,X) + diag_add I only (gpx_append: rows of new observations)
void gpx_launch_kbuild_rows(TileMat L, int nt, int i_first, const double* Xs, const double* xsq, int64_t N, int d,
                            int kernel_id, double variance, double rdiv, double diag_add, cudaStream_t st) {
    if (i_first >= nt) return;
    dim3 grid(nt - i_first, nt);
    kb_launch_lower(kernel_id, grid, kb_smem(d), st, L, nt, 0, 1, i_first, Xs, xsq, N, d, variance, rdiv, diag_add);
}

void gpx_launch_kcross(TileMat Tt, int nb, int nbt_layout, int nt, const double* Xt, const double* xtsq,
                       int64_t Nt_valid, const double* Xs, const double* xsq, int64_t N, int d, int kernel_id,
                       double variance, double rdiv, const double* alpha, int O, int64_t Np, double* mean_partial,
                       cudaStream_t st, double* vec_out, int vec_rows) {
    dim3 grid(nb, nt);
    switch (kernel_id) {
#define GPX_KCROSS(K) case K: kcross_kernel<K><<<grid, 256, kb_smem(d), st>>>(Tt, nbt_layout, nt, Xt, xtsq, Nt_valid, Xs, xsq, N, d, variance, \
                                                                          rdiv, alpha, O, Np, mean_partial, vec_out, vec_rows); break;
        GPX_KCROSS(0) GPX_KCROSS(1) GPX_KCROSS(2) GPX_KCROSS(3)
#undef GPX_KCROSS
    }
}

void gpx_launch_mean_reduce(const double* mean_partial, int nbt, int nt, int O, double mean_const, const double* ystats,
                            double* mean, int64_t row0, int64_t Ns, cudaStream_t st) {
    // nbt = layout stride (test tiles per tile column of mean_partial); rows past Ns are skipped in the kernel
    int64_t total = (int64_t)nbt * GPX_TILE * O;
    int threads = 128;
    mean_reduce_kernel<<<(int)((total + threads - 1) / threads), threads, 0, st>>>(mean_partial, nbt, nt, O,
                                                                                  mean_const, ystats, mean, row0, Ns);
}

int gpx_var_reduce_splits(int nbt, int nt, int num_sms) {
    int S = (4 * num_sms + nbt - 1) / nbt;     // ~4 CTAs per SM in flight
    if (S > nt) S = nt;
    if (S > 64) S = 64;
    return S < 1 ? 1 : S;
}

// `partial` holds nbt * S * 128 doubles (S = gpx_var_reduce_splits)
void gpx_launch_var_reduce(TileMat Tt, int nbt, int nt, int S, double* partial, double prior_var, const double* ystats,
                           int var_cols, double* var, int64_t row0, int64_t Ns, cudaStream_t st) {
    var_partial_kernel<<<dim3(nbt, S), 256, 0, st>>>(Tt, nt, S, partial);
    const int64_t total = (int64_t)nbt * GPX_TILE;
    var_final_kernel<<<(int)((total + 127) / 128), 128, 0, st>>>(partial, nbt, S, prior_var, ystats, var_cols, var, row0, Ns);
}

void gpx_launch_detile(TileMat M, int lower, int64_t rows, int64_t cols, double* out, cudaStream_t st) {
    dim3 grid((unsigned)((cols + 255) / 256), (unsigned)rows);
    detile_kernel<<<grid, 256, 0, st>>>(M, lower, rows, cols, out);
}

// =====================================================================================================================
// Gradient of the log marginal likelihood w.r.t. the kernel hyper-parameters (`do_optimize`).
//   dLML/dtheta = sum_ij W_ij dK_ij/dtheta,   W = 0.5 (alpha alpha^T - O Ky^-1)      (GPy: dL_dK)
// One CTA per lower tile (I >= J) of Ky^-1: the kernel entries and their derivatives are recomputed from the scaled
// inputs (nothing N x N is stored besides Ky^-1), multiplied with W and reduced to 2 + n_ls partial sums per tile in a
// fixed order (deterministic).  Off-diagonal elements count twice (symmetry).
//   d/d variance      : K_ij / variance
//   d/d lengthscale_q : f(r) (xs_iq - xs_jq)^2 / l_q     with xs = x / l and
//        f = K (RBF) | 5/3 v (1 + sqrt5 r) exp(-sqrt5 r) (Matern52) | 3 v exp(-sqrt3 r) (Matern32) | v exp(-r) / r (Exponential)
//        (isotropic: summed over q)
//   d/d noise         : trace(W)
// =====================================================================================================================
namespace {

__device__ __forceinline__ void gpx_k_and_f(int kernel_id, double variance, double r, double& k, double& f) {
    if (kernel_id == 0) { k = variance * exp(-0.5 * (r * r)); f = k; }
    else if (kernel_id == 1) {
        const double s5 = 2.23606797749978969641;
        const double e = exp(-s5 * r);
        k = variance * (1.0 + s5 * r + (5.0 / 3.0) * (r * r)) * e;
        f = (5.0 / 3.0) * variance * (1.0 + s5 * r) * e;
    } else if (kernel_id == 2) {
        const double s3 = 1.73205080756887729353;
        const double e = exp(-s3 * r);
        k = variance * (1.0 + s3 * r) * e;
        f = 3.0 * variance * e;
    } else {
        k = variance * exp(-r);
        f = r > 0.0 ? k / r : 0.0;
    }
}

constexpr int kMaxGradLs = kMaxD;

// grid (nt, nt); partial[(I * nt + J) * stride + {0: variance, 1: noise, 2..: lengthscales}]
__global__ void __launch_bounds__(256) lml_grad_tile_kernel(TileMat Kinv, int nt, const double* __restrict__ Xs,
                                                            const double* __restrict__ xsq, int64_t N, int d,
                                                            int kernel_id, double variance, double rdiv,
                                                            const double* __restrict__ ls, int n_ls,
                                                            const double* __restrict__ alpha, int O, int64_t Np,
                                                            double* __restrict__ partial, int stride, int rg, int world,
                                                            int rank) {
    extern __shared__ __align__(16) double sm[];
    const int I = blockIdx.x, J = blockIdx.y;
    const int tid = threadIdx.x;
    double* out = partial + ((int64_t)I * nt + J) * stride;
    if (I < J) return;
    if (world > 1 && (I / rg) % world != rank) return;   // sharded handle: the row groups of Ky^-1 are dealt out cyclically
    const int dp = d | 1;
    double* sa = sm;
    double* sb = sa + GPX_TILE * dp;
    double* sasq = sb + GPX_TILE * dp;
    double* sbsq = sasq + GPX_TILE;
    double* red = sbsq + GPX_TILE;        // [8 warps][stride]
    for (int e = tid; e < GPX_TILE * d; e += blockDim.x) {
        int rr = e / d, k = e - rr * d;
        sa[rr * dp + k] = Xs[((int64_t)I * GPX_TILE + rr) * d + k];
        sb[rr * dp + k] = Xs[((int64_t)J * GPX_TILE + rr) * d + k];
    }
    if (tid < GPX_TILE) {
        sasq[tid] = xsq[(int64_t)I * GPX_TILE + tid];
        sbsq[tid] = xsq[(int64_t)J * GPX_TILE + tid];
    }
    __syncthreads();
    const double* kt = tile_ptr(Kinv, I, J);
    const int warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    double gvar = 0.0, gnoise = 0.0;
    double gls[kMaxGradLs];
#pragma unroll 1
    for (int q = 0; q < n_ls; q++) gls[q] = 0.0;
    for (int rb = 0; rb < 2; rb++) {
        const int r = warp * 16 + rb * 8 + g;
        const int64_t gi = (int64_t)I * GPX_TILE + r;
        for (int kc = 0; kc < 16; kc++) {
            const double2 kinv2 = *reinterpret_cast<const double2*>(kt + kc * 1024 + r * 8 + 2 * t);
            for (int hh = 0; hh < 2; hh++) {
                const int c = kc * 8 + 2 * t + hh;
                const int64_t gj = (int64_t)J * GPX_TILE + c;
                if (gi >= N || gj >= N || gj > gi) continue;
                double dot = 0.0;
                for (int k = 0; k < d; k++) dot = fma(sa[r * dp + k], sb[c * dp + k], dot);
                double r2 = fmax(-2.0 * dot + (sasq[r] + sbsq[c]), 0.0);
                if (gi == gj) r2 = 0.0;
                double kv, f;
                gpx_k_and_f(kernel_id, variance, gpx_scaled_r(r2, rdiv), kv, f);
                double aa = 0.0;
                for (int o = 0; o < O; o++) aa = fma(alpha[(int64_t)o * Np + gi], alpha[(int64_t)o * Np + gj], aa);
                const double kinv = hh ? kinv2.y : kinv2.x;
                const double w = (gi == gj ? 0.5 : 1.0) * (aa - (double)O * kinv);   // off-diagonal counted twice
                gvar = fma(w, kv / variance, gvar);
                if (gi == gj) gnoise += w;
                const double wf = w * f;
                if (n_ls == 1) {
                    double s = 0.0;
                    for (int k = 0; k < d; k++) { double dx = sa[r * dp + k] - sb[c * dp + k]; s = fma(dx, dx, s); }
                    gls[0] = fma(wf, s / (rdiv * rdiv), gls[0]);   // isotropic: the staged inputs are unscaled
                } else {
                    for (int k = 0; k < d; k++) { double dx = sa[r * dp + k] - sb[c * dp + k]; gls[k] = fma(wf, dx * dx, gls[k]); }
                }
            }
        }
    }
    // fixed-order reduction: lanes (shuffle tree), then warps (shared memory)
    auto wsum = [](double v) { for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off); return v; };
    gvar = wsum(gvar);
    gnoise = wsum(gnoise);
    if (lane == 0) { red[warp * stride + 0] = gvar; red[warp * stride + 1] = gnoise; }
    for (int q = 0; q < n_ls; q++) {
        double v = wsum(gls[q]);
        if (lane == 0) red[warp * stride + 2 + q] = v / ls[q];
    }
    __syncthreads();
    if (tid < 2 + n_ls) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += red[w * stride + tid];
        out[tid] = s;
    }
}

// out[p] = sum over lower tiles of partial[.][p]  (fixed order)
__global__ void lml_grad_reduce_kernel(const double* __restrict__ partial, int nt, int stride, int nparam, double* __restrict__ out) {
    __shared__ double red[256];
    const int p = blockIdx.x;
    double s = 0.0;
    for (int64_t e = threadIdx.x; e < (int64_t)nt * nt; e += blockDim.x) {
        int I = (int)(e / nt), J = (int)(e % nt);
        if (I >= J) s += partial[e * stride + p];
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0 && p < nparam) out[p] = red[0];
}

// Tt <- identity (rectangular nt x nt tile matrix)
__global__ void identity_tiles_kernel(TileMat T, int nt) {
    const int n = blockIdx.x, J = blockIdx.y;
    double* tp = tile_ptr(T, n, J);
    for (int e = threadIdx.x; e < GPX_TILE_ELEMS; e += blockDim.x) {
        int r = (e >> 3) & 127, c = (e >> 10) * 8 + (e & 7);
        tp[e] = (n == J && r == c) ? 1.0 : 0.0;
    }
    gpx_publish_for_async_proxy();
}

// sum over all elements of alpha (gradient of the constant mean)
__global__ void sum_kernel(const double* __restrict__ a, int64_t n, double* __restrict__ out) {
    __shared__ double red[256];
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += a[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = red[0];
}

}  // namespace

static void gpx_grad_init_attr() {
    cudaFuncSetAttribute(lml_grad_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)((2 * GPX_TILE * (kMaxD | 1) + 2 * GPX_TILE + 8 * (2 + kMaxD)) * sizeof(double)));
}

namespace {
// T(i, j0 + j) <- (i == j) ? I : 0   for i, j < w   (rows are 0-based tile rows of a rectangular tile matrix)
__global__ void identity_block_kernel(TileMat T, int j0) {
    const int i = blockIdx.x, j = blockIdx.y;
    double* tp = tile_ptr(T, i, j0 + j);
    for (int e = threadIdx.x; e < GPX_TILE_ELEMS; e += blockDim.x) {
        int r = (e >> 3) & 127, c = (e >> 10) * 8 + (e & 7);
        tp[e] = (i == j && r == c) ? 1.0 : 0.0;
    }
    gpx_publish_for_async_proxy();
}

// M(p0 + j, p0 + k) <- Y(k, p0 + j)^T for k <= j < w.  grid (w (w + 1) / 2 tile pairs, 16 sub-blocks of 32 x 32).
__global__ void __launch_bounds__(256) transpose_block_kernel(TileMat Y, TileMat M, int p0, int w) {
    __shared__ double sm[32][33];
    int j = 0, rem = blockIdx.x;
    while (rem > j) { rem -= j + 1; j++; }       // pair index -> (j, k = rem), k <= j
    const int k = rem;
    const double* src = tile_ptr(Y, k, p0 + j);   // element (r, c) of Y's tile (k, j)
    double* dst = tile_ptr(M, p0 + j, p0 + k);    // element (c, r) of M's tile (j, k)
    const int br = (blockIdx.y >> 2) * 32, bc = (blockIdx.y & 3) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int rr = ty; rr < 32; rr += 8) sm[rr][tx] = src[gpx_tile_idx(br + rr, bc + tx)];
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) dst[gpx_tile_idx(bc + rr, br + tx)] = sm[tx][rr];
    gpx_publish_for_async_proxy();
}
}  // namespace

void gpx_launch_identity_block(TileMat T, int w, int j0, cudaStream_t st) {
    identity_block_kernel<<<dim3(w, w), 256, 0, st>>>(T, j0);
}

void gpx_launch_transpose_block(TileMat Y, TileMat M, int p0, int w, cudaStream_t st) {
    transpose_block_kernel<<<dim3(w * (w + 1) / 2, 16), 256, 0, st>>>(Y, M, p0, w);
}

void gpx_launch_identity_tiles(TileMat T, int nt, cudaStream_t st) {
    dim3 grid(nt, nt);
    identity_tiles_kernel<<<grid, 256, 0, st>>>(T, nt);
}

void gpx_launch_lml_grad(TileMat Kinv, int nt, const double* Xs, const double* xsq, int64_t N, int d, int kernel_id,
                         double variance, double rdiv, const double* ls, int n_ls, const double* alpha, int O, int64_t Np,
                         double* partial, double* out, cudaStream_t st, int rg, int world, int rank) {
    const int stride = 2 + n_ls;
    size_t smem = (size_t)(2 * GPX_TILE * (d | 1) + 2 * GPX_TILE + 8 * stride) * sizeof(double);
    dim3 grid(nt, nt);
    if (world > 1) cudaMemsetAsync(partial, 0, (size_t)nt * nt * stride * sizeof(double), st);   // tiles of the other ranks
    lml_grad_tile_kernel<<<grid, 256, smem, st>>>(Kinv, nt, Xs, xsq, N, d, kernel_id, variance, rdiv, ls, n_ls, alpha, O, Np,
                                                  partial, stride, rg, world, rank);
    lml_grad_reduce_kernel<<<stride, 256, 0, st>>>(partial, nt, stride, stride, out);
    // out[stride] = sum(alpha): gradient of the constant mean
    sum_kernel<<<1, 256, 0, st>>>(alpha, Np * O, out + stride);
}
