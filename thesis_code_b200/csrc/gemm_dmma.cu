// Tile GEMM on the FP64 tensor-core path (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4) for sm_100a.
//
//   C(I,J)  =  C(I,J) - sum_k A(I,k) * B(J,k)^T        (SYRK / GEMM trailing updates, triangular-solve updates)
//   C(I,J)  =           sum_k A(I,k) * B(k,k)^T        (TRSM against an inverted diagonal tile, overwrite)
//
// over 128x128 fp64 tiles in the chunked K-major layout of gpx_common.cuh.  Both operands are "row x k" with k
// fastest, which is exactly what mma.m8n8k4.row.col wants, so the same kernel serves the Cholesky trailing
// update (A = B = factored panel), the panel TRSM (B = inverted diagonal tile) and the predictive-variance
// triangular solve (A = transposed cross-kernel workspace, B = L).
//
// Structure (TWO co-resident CTAs per SM, 288 threads each, one 64x128 HALF tile (64 rows) per work slot):
//   * the last warp's lane 0 = producer: for every output half tile and every 16-wide K slice issues two 4 KB (A: the 64 rows
//     of this half, one per 8-column chunk) and one 16 KB (B: 128 rows) 1-D TMA bulk copies
//     (cp.async.bulk.shared::cluster.global, SASS UBLKCP) into a 4-stage shared-memory ring (96 KB per CTA), signalled
//     through mbarriers (full/empty);
//   * warps 0..7 = consumers, 2 (M) x 4 (N), each owning a 32x32 accumulator block = 4x4 DMMA 8x8 fragments held
//     in registers (32 doubles / thread).  Per 8-k chunk a warp issues 8 conflict-free LDS.128 and 32 DMMA.
//     (Round 1 used 4 warps of 64x32: four warps per SM sub-partition instead of two hide the LDS -> DMMA latency
//     better where the pipeline is short -- K = 1 tile 31.5 -> 32.6, K = 4 tiles 34.5 -> 35.1 TFLOP/s, K = 16 unchanged
//     at 35.9 (tools/gemm_bench.py, profiles/r02_gemm_warps.txt) -- which is what the sharded runs (K = 4) live on.)
//   * GemmJob::rowwise = "panel solve" mode: one tile row per CTA, columns in descending order, contraction cut at the
//     diagonal, in place -- the in-panel part of the variance solve against the inverted diagonal block (gpx_api.cu).
//   * accumulate mode initialises the accumulators with the old C half tile (coalesced 512 B fragment loads; the
//     producer has already pulled it into L2 with a bulk prefetch) and feeds NEGATED A fragments: store-only epilogue.
//   * Why two small CTAs instead of one 128x128 CTA (round-1 first version): the per-tile prologue/epilogue (C load
//     ~1.2 us, store drain + pipeline hand-over ~1.2 us, measured with tools/gemm_bench.py) left the DMMA pipe idle
//     3% of the time at K = 4 tiles and 12% at K = 1; with two independent CTAs per SM one keeps the pipe busy while
//     the other is between tiles (one warp per SM sub-partition already reaches 95% of the DMMA peak).
//     The split is along ROWS so that the in-place TRSM (C = A * Linv^T written over A) stays race-free: a half only
//     reads the A rows it overwrites.
//
// Roofline: FP64 tensor.  Algorithmic flops per output tile = 2 * 128^3 * (k1 - k0); measured DMMA peak on this
// pool's B200 is 36.96 TFLOP/s (profiles/r01_fp64_peak.txt).
#include <cstdio>
#include <cstdlib>
#include "gpx_common.cuh"
#include "tile_walk.cuh"

namespace {

constexpr int kStages = 4;
constexpr int kStageK = 16;                              // k columns per stage
constexpr int kOperBytes = GPX_TILE * kStageK * 8;       // B: 128 rows x 16 k = 16 KB per stage
constexpr int kHalf = 64;                                // output rows (= A rows) per work slot
constexpr int kAChunkBytes = kHalf * 8 * 8;              // A: 64 rows x one 8-column chunk = 4 KB
constexpr int kStageBytes = 2 * kAChunkBytes + kOperBytes;   // 24 KB: [A chunk 0][A chunk 1][B]
#ifndef GPX_GEMM_WARPS
#define GPX_GEMM_WARPS 8
#endif
constexpr int kConsumerWarps = GPX_GEMM_WARPS;           // 4: 1 (M) x 4 (N) warps of 64x32;  8: 2 (M) x 4 (N) warps of 32x32
constexpr int kWarpRows = kHalf / (kConsumerWarps / 4);  // accumulator rows per warp
constexpr int kFragI = kWarpRows / 8;                    // 8x8 accumulator fragments per warp along M
constexpr int kThreads = (kConsumerWarps + 1) * 32;
constexpr int kSmemBytes = kStages * kStageBytes + 2 * kStages * 8 + 128;
#ifndef GPX_WAR_FENCE_MODE
#define GPX_WAR_FENCE_MODE 1
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src_gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src_gmem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

template <bool kOverwrite>
__global__ void __launch_bounds__(kThreads, 2) gemm_dmma_kernel(const GemmJob job) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty = full + kStages;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < kStages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    TileWalk walk;
    walk.init(job);
    int stage = 0;
    uint32_t phase = 0;
    const int nslice = GPX_TILE / kStageK;  // 8 slices per operand tile
    // Slots of this CTA.  Contiguous mode (job.interleave == 0): [b * chunk, (b + 1) * chunk).  Interleaved mode: the R CTAs of
    // a "round" (R = the number of CTAs resident on the GPU, dispatched in blockIdx order) share a window of R * chunk
    // consecutive slots and each walks it with stride R, so at any moment the resident CTAs work on ~R NEIGHBOURING slots:
    // the operand panels they touch (a few tile rows of A, one strip of B) stay in L2, whereas with contiguous chunks the
    // resident set spans R * chunk slots -- at chunk 16 that is the whole A panel, re-fetched from HBM strip after strip.
    const long long R = gpx_cta_slot_stride(job.interleave);
    const long long t_begin = gpx_cta_first_slot(blockIdx.x, job.chunk, job.interleave);
    const long long t_end = t_begin + (long long)job.chunk * R;

    if (warp == kConsumerWarps) {
        // ---------------- producer ----------------
        if (lane == 0) {
            for (long long t = t_begin; t < t_end; t += R) {
                int I, J, half; bool valid;
                if (!walk.locate(job, t, I, J, half, valid)) break;
                if (!valid) continue;
                // the consumers will read-modify-write C(I,J) when they reach it: pull the tile into L2 now
                if (!kOverwrite && half == 0) bulk_prefetch_l2(tile_ptr(job.C, I, J), GPX_TILE_BYTES);
                const int k_hi = gpx_job_k1(job, J);
                for (int kb = job.k0; kb < k_hi; kb++) {
                    const double* a = tile_ptr(job.A, I, kb) + half * kHalf * 8;
                    const double* b = job.bdiag ? tile_ptr(job.B, kb, kb) : tile_ptr(job.B, J, kb);
                    for (int s = 0; s < nslice; s++) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        mbar_expect_tx(&full[stage], kStageBytes);
                        unsigned char* dst = smem + stage * kStageBytes;
                        // A rows [64*half, 64*half+64) of the two 8-column chunks of this K slice, then all of B's slice
                        bulk_g2s(dst, a + s * (GPX_TILE * kStageK), kAChunkBytes, &full[stage]);
                        bulk_g2s(dst + kAChunkBytes, a + s * (GPX_TILE * kStageK) + 1024, kAChunkBytes, &full[stage]);
                        bulk_g2s(dst + 2 * kAChunkBytes, b + s * (GPX_TILE * kStageK), kOperBytes, &full[stage]);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else {
        // ---------------- consumers ----------------
        const int wn = warp & 3, wm = warp >> 2;   // warps side by side along N (32 columns each), stacked along M
        const int g = lane >> 2, tq = lane & 3;
        // fragment offsets inside an 8-k chunk ([row][8] doubles)
        const int a_off = (wm * kWarpRows + g) * 8 + 2 * tq;
        const int b_off = (wn * 32 + g) * 8 + 2 * tq;
        // accumulator fragment (i,j) of row half h lives at tile offset  (wn*4 + j) * 1024 + (h*64 + wm*rows + i*8 + g) * 8 + 2*tq
        const int c_off0 = (wn * 4) * 1024 + (wm * kWarpRows + g) * 8 + 2 * tq;

        for (long long t = t_begin; t < t_end; t += R) {
            int I, J, half; bool valid;
            if (!walk.locate(job, t, I, J, half, valid)) break;
            if (!valid) continue;
            double* ctile = tile_ptr(job.C, I, J);
            const int c_off = c_off0 + half * kHalf * 8;
            double acc[kFragI][4][2];
            if (kOverwrite) {
#pragma unroll
                for (int i = 0; i < kFragI; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) { acc[i][j][0] = 0.0; acc[i][j][1] = 0.0; }
            } else {
#pragma unroll
                for (int i = 0; i < kFragI; i++)
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        double2 v = __ldcg(reinterpret_cast<const double2*>(ctile + c_off + j * 1024 + i * 64));
                        acc[i][j][0] = v.x; acc[i][j][1] = v.y;
                    }
            }
            const int k_hi = gpx_job_k1(job, J);
            for (int kb = job.k0; kb < k_hi; kb++) {
                for (int s = 0; s < nslice; s++) {
                    mbar_wait(&full[stage], phase);
                    const double* As = reinterpret_cast<const double*>(smem + stage * kStageBytes);
                    const double* Bs = As + 2 * kHalf * 8;        // A: [2 chunks][64 rows][8], then B: [2 chunks][128 rows][8]
#pragma unroll
                    for (int ch = 0; ch < kStageK / 8; ch++) {
                        double2 bf[4];
#pragma unroll
                        for (int j = 0; j < 4; j++)
                            bf[j] = *reinterpret_cast<const double2*>(Bs + ch * 1024 + b_off + j * 64);
#pragma unroll
                        for (int i = 0; i < kFragI; i++) {
                            double2 af = *reinterpret_cast<const double2*>(As + ch * (kHalf * 8) + a_off + i * 64);
                            if (!kOverwrite) { af.x = -af.x; af.y = -af.y; }
#pragma unroll
                            for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], af.x, bf[j].x);
#pragma unroll
                            for (int j = 0; j < 4; j++) dmma(acc[i][j][0], acc[i][j][1], af.y, bf[j].y);
                        }
                    }
                    // Cross-proxy WAR: the stage is re-filled by an async-proxy bulk copy as soon as the empty barrier
                    // completes, while it was read here through the generic proxy (LDS).  Without this proxy fence the
                    // refill was observed to overtake the reads on B200 (non-deterministic tiles at N >= ~5000;
                    // tools/gpu_debug2.py), with it 24/24 repeated factorizations are bit-identical.
#if GPX_WAR_FENCE_MODE == 1
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stage]);
#else
                    __syncwarp();
                    if (lane == 0) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        mbar_arrive(&empty[stage]);
                    }
#endif
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
#pragma unroll
            for (int i = 0; i < kFragI; i++)
#pragma unroll
                for (int j = 0; j < 4; j++)
                    *reinterpret_cast<double2*>(ctile + c_off + j * 1024 + i * 64) =
                        make_double2(acc[i][j][0], acc[i][j][1]);
            gpx_publish_for_async_proxy();
        }
    }
}

}  // namespace

void gpx_gemm_init() {
    cudaFuncSetAttribute(gemm_dmma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    cudaFuncSetAttribute(gemm_dmma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    // two CTAs per SM need 2 x 96 KB of shared memory: ask for the maximum carve-out
    cudaFuncSetAttribute(gemm_dmma_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(gemm_dmma_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

long long gpx_launch_gemm(const GemmJob& job_in, int num_sms, cudaStream_t st, int64_t* launch_counter) {
    GemmJob job = job_in;
    if (job.i1 <= job.i0 || job.ncols <= 0 || job.k1 <= job.k0) return 0;
    long long slots = 2 * count_slots(job);   // half-tile work slots
    if (slots <= 0) return 0;
    // Slots per CTA: a CTA owns `chunk` consecutive half-tile slots and retires; later CTAs are scheduled as SMs free
    // up (this balances ragged triangular jobs and lets the panel / NCCL kernels of the other streams interleave).
    // Two CTAs are resident per SM.  Pick the chunk that minimises  rounds * (chunk * slot_time + per-CTA overhead):
    // large chunks amortise the pipeline fill, small chunks shrink the idle tail of the last round.
    const int resident = 2 * num_sms;
    const int max_chunk = 2 * (job.chunk < 1 ? 16 : (job.chunk > 64 ? 64 : job.chunk));  // caller's cap is in full tiles
    long long chunk = 1;
    if (job.rowwise) {
        chunk = job.ncols;         // one half tile row per CTA (the in-place panel solve relies on it)
    } else if (slots > resident) {
        const double slot_us = 17.0 * (job.k1 - job.k0), ovh_us = 4.0;   // a half tile at half an SM's rate
        double best = 1e300;
        for (long long c = 1; c <= max_chunk; c++) {
            long long ctas = (slots + c - 1) / c;
            long long rounds = (ctas + resident - 1) / resident;
            double cost = (double)rounds * ((double)c * slot_us + ovh_us);
            if (cost < best - 1e-9) { best = cost; chunk = c; }
        }
    }
    job.chunk = (int)chunk;
    int grid = (int)((slots + chunk - 1) / chunk);
    // interleaved slot assignment (see the kernel) whenever a CTA owns more than one slot; the row-wise panel solve needs
    // its contiguous half tile row
    job.interleave = (job.rowwise || chunk == 1 || job.interleave < 0) ? 0 : resident;
    grid = (int)gpx_grid_for(slots, (int)chunk, job.interleave);
    if (job.overwrite)
        gemm_dmma_kernel<true><<<grid, kThreads, kSmemBytes, st>>>(job);
    else
        gemm_dmma_kernel<false><<<grid, kThreads, kSmemBytes, st>>>(job);
    if (launch_counter) (*launch_counter)++;
    return count_steps(job);
}
