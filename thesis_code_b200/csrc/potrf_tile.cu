// Diagonal-tile kernel of the blocked Cholesky: factor one 128x128 fp64 tile (lower Cholesky, LAPACK dpotrf
// semantics) and invert the triangular factor, so that the panel TRSM, the alpha solves and the predictive-variance
// solve all become DMMA GEMMs / GEMVs against the inverted tile.  This kernel is the serial critical path of the
// factorisation (one launch per tile column), so it is latency-optimised:
//
//   * one CTA of 544 threads; thread t < 528 owns one 4x4 block (bi >= bj) of the 32x32 lower block grid IN REGISTERS;
//   * Cholesky, right-looking over 4-column block steps: the diagonal owner factors + inverts its 4x4 block,
//     the 31-kb owners below apply the inverse (register TRSM) and publish their 4 columns to shared memory,
//     every remaining block does a rank-4 register update.  Two __syncthreads per block step (64 in total, the
//     first version had 256 and was ~8x slower: profiles/r01_launches_c3_summary.txt);
//   * inverse, right-looking over block rows with the same ownership: X(k,:) = Dinv_k U(k,:), U(i,:) -= L(i,k) X(k,:);
//     one __syncthreads per step.
//   * results go straight to the two global tiles in the chunked K-major layout (32 B sector writes).
// info: first non-positive (or NaN) pivot, reported LAPACK-style as a 1-based global row index.
#include "gpx_common.cuh"

namespace {
constexpr int kNB = 32;                 // 4x4 blocks per dimension
constexpr int kOwners = kNB * (kNB + 1) / 2;   // 528
constexpr int kThreadsP = 544;
constexpr int kBS = 18;                 // doubles per published 4x4 block: 16 + 2 padding.  With a 128 B pitch every lane of
                                        // a warp (consecutive bj) hit the same banks: 32-way conflicts made this kernel
                                        // 129 us; 144 B spreads a warp's double2 loads over all banks (tools/gemm_bench.py)

struct PotrfSmem {
    double Lf[kNB * kNB * kBS];         // final L for the inverse phase, 4x4 blocks, block (bi, k) at (k*32 + bi)*kBS: the
                                        // writers (one block column, consecutive bi) and the readers (one block per warp,
                                        // broadcast) are both conflict-free
    double P[2][kNB * kBS];             // published 4-column panel, block-contiguous: [bi*kBS + r*4 + c]
    double Xr[2][kNB * kBS];            // published 4-row block row of the inverse: [bj*kBS + r*4 + c]
    double Dinv[kNB][16];               // inverted 4x4 diagonal blocks
};

__device__ __forceinline__ double* blk_ptr(double* tile, int bi, int bj, int r) {
    // element (4*bi + r, 4*bj) of a chunked-layout tile; the 4 doubles of the block row are contiguous (32 B)
    const int row = 4 * bi + r, c = 4 * bj;
    return tile + ((c >> 3) << 10) + (row << 3) + (c & 7);
}

__global__ void __launch_bounds__(kThreadsP, 1) potrf_tile_kernel(double* __restrict__ tile, double* __restrict__ inv_tile,
                                                                  int* __restrict__ info, int col_index) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PotrfSmem& S = *reinterpret_cast<PotrfSmem*>(smem_raw);
    const int t = threadIdx.x;
    const bool owner = t < kOwners;
    int bi = 0, bj = 0;
    if (owner) {
        bi = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
        while ((bi + 1) * (bi + 2) / 2 <= t) bi++;
        while (bi * (bi + 1) / 2 > t) bi--;
        bj = t - bi * (bi + 1) / 2;
    }
    double b[4][4];
    if (owner) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const double2* p = reinterpret_cast<const double2*>(blk_ptr(tile, bi, bj, r));
            double2 v0 = p[0], v1 = p[1];
            b[r][0] = v0.x; b[r][1] = v0.y; b[r][2] = v1.x; b[r][3] = v1.y;
        }
    }

    // ------------------------------------------------ Cholesky ------------------------------------------------
    for (int kb = 0; kb < kNB; kb++) {
        const int buf = kb & 1;
        if (owner && bj == kb && bi == kb) {
            // factor the 4x4 diagonal block in registers.  rsqrt (MUFU + Newton, ~1 ulp) replaces sqrt + two divisions
            // per column on this single-thread critical path: L_jj = d * rsqrt(d), 1 / L_jj = rsqrt(d).
            double rinv[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                double d = b[j][j];
                if (!(d > 0.0)) {
                    atomicCAS(info, 0, col_index * GPX_TILE + 4 * kb + j + 1);
                    d = 1.0;
                }
                const double inv = rsqrt(d);
                rinv[j] = inv;
                b[j][j] = d * inv;
#pragma unroll
                for (int i = j + 1; i < 4; i++) b[i][j] *= inv;
#pragma unroll
                for (int c = j + 1; c < 4; c++)
#pragma unroll
                    for (int i = c; i < 4; i++) b[i][c] -= b[i][j] * b[c][j];
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = r + 1; c < 4; c++) b[r][c] = 0.0;
            // invert the lower 4x4 (multiplications by the stored reciprocals only)
            double di[4][4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
#pragma unroll
                for (int j = 0; j < 4; j++) di[i][j] = 0.0;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                di[j][j] = rinv[j];
#pragma unroll
                for (int i = j + 1; i < 4; i++) {
                    double s = 0.0;
#pragma unroll
                    for (int k = j; k < i; k++) s += b[i][k] * di[k][j];
                    di[i][j] = -s * rinv[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) S.Dinv[kb][i * 4 + j] = di[i][j];
        }
        __syncthreads();
        if (owner && bj == kb) {
            if (bi > kb) {
                // register TRSM: X = B * Ld^-T,  x[r][c] = sum_{k<=c} b[r][k] * di[c][k]
                double di[16];
#pragma unroll
                for (int e = 0; e < 16; e++) di[e] = S.Dinv[kb][e];
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    double x[4];
#pragma unroll
                    for (int c = 0; c < 4; c++) {
                        double s = 0.0;
#pragma unroll
                        for (int k = 0; k <= c; k++) s += b[r][k] * di[c * 4 + k];
                        x[c] = s;
                    }
#pragma unroll
                    for (int c = 0; c < 4; c++) b[r][c] = x[c];
                }
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int c = 0; c < 4; c++) S.P[buf][bi * kBS + r * 4 + c] = b[r][c];
            }
            // this block of L is final: keep a shared copy for the inverse phase and write it to global memory
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int c = 0; c < 4; c++) S.Lf[(bj * kNB + bi) * kBS + r * 4 + c] = b[r][c];
                double2* p = reinterpret_cast<double2*>(blk_ptr(tile, bi, bj, r));
                p[0] = make_double2(b[r][0], b[r][1]);
                p[1] = make_double2(b[r][2], b[r][3]);
            }
        }
        __syncthreads();
        if (owner && bj > kb) {
            // rank-4 update: b[r][c] -= sum_k P[bi][r][k] * P[bj][c][k]
            double pa[4][4], pb[4][4];
            const double2* qa = reinterpret_cast<const double2*>(&S.P[buf][bi * kBS]);
            const double2* qb = reinterpret_cast<const double2*>(&S.P[buf][bj * kBS]);
#pragma unroll
            for (int r = 0; r < 4; r++) {
                double2 a0 = qa[2 * r], a1 = qa[2 * r + 1], c0 = qb[2 * r], c1 = qb[2 * r + 1];
                pa[r][0] = a0.x; pa[r][1] = a0.y; pa[r][2] = a1.x; pa[r][3] = a1.y;
                pb[r][0] = c0.x; pb[r][1] = c0.y; pb[r][2] = c1.x; pb[r][3] = c1.y;
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    double s = b[r][c];
#pragma unroll
                    for (int k = 0; k < 4; k++) s -= pa[r][k] * pb[c][k];
                    b[r][c] = s;
                }
        }
    }

    // ------------------------------------------------ inverse -------------------------------------------------
    // U(bi,bj) in registers: identity on the diagonal blocks, zero below.
    if (owner) {
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) b[r][c] = (bi == bj && r == c) ? 1.0 : 0.0;
    }
    __syncthreads();   // Lf complete
    for (int k = 0; k < kNB; k++) {
        const int buf = k & 1;
        if (owner && bi == k) {
            // X(k,bj) = Dinv_k * U(k,bj):  x[r][c] = sum_{m<=r} di[r][m] * u[m][c]
            double di[16];
#pragma unroll
            for (int e = 0; e < 16; e++) di[e] = S.Dinv[k][e];
            double x[4][4];
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    double s = 0.0;
#pragma unroll
                    for (int m = 0; m <= r; m++) s += di[r * 4 + m] * b[m][c];
                    x[r][c] = s;
                }
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int c = 0; c < 4; c++) S.Xr[buf][bj * kBS + r * 4 + c] = x[r][c];
                double2* p = reinterpret_cast<double2*>(blk_ptr(inv_tile, bi, bj, r));
                p[0] = make_double2(x[r][0], x[r][1]);
                p[1] = make_double2(x[r][2], x[r][3]);
            }
        }
        __syncthreads();
        if (owner && bi > k && bj <= k) {
            // U(bi,bj) -= L(bi,k) * X(k,bj)
            double l[4][4], x[4][4];
            const double2* qx = reinterpret_cast<const double2*>(&S.Xr[buf][bj * kBS]);
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const double2* ql = reinterpret_cast<const double2*>(&S.Lf[(k * kNB + bi) * kBS + r * 4]);
                double2 l0 = ql[0], l1 = ql[1], x0 = qx[2 * r], x1 = qx[2 * r + 1];
                l[r][0] = l0.x; l[r][1] = l0.y; l[r][2] = l1.x; l[r][3] = l1.y;
                x[r][0] = x0.x; x[r][1] = x0.y; x[r][2] = x1.x; x[r][3] = x1.y;
            }
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    double s = b[r][c];
#pragma unroll
                    for (int m = 0; m < 4; m++) s -= l[r][m] * x[m][c];
                    b[r][c] = s;
                }
        }
    }
    // zero the strictly upper blocks of both tiles (mirror of each strictly lower block)
    if (owner && bi > bj) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            double2* p = reinterpret_cast<double2*>(blk_ptr(tile, bj, bi, r));
            double2* q = reinterpret_cast<double2*>(blk_ptr(inv_tile, bj, bi, r));
            p[0] = make_double2(0.0, 0.0); p[1] = make_double2(0.0, 0.0);
            q[0] = make_double2(0.0, 0.0); q[1] = make_double2(0.0, 0.0);
        }
    }
    gpx_publish_for_async_proxy();
}
}  // namespace

void gpx_potrf_init() {
    cudaFuncSetAttribute(potrf_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PotrfSmem));
}

void gpx_launch_potrf_tile(double* tile, double* inv_tile, int* info, int col_index, cudaStream_t st) {
    potrf_tile_kernel<<<1, kThreadsP, sizeof(PotrfSmem), st>>>(tile, inv_tile, info, col_index);
}
