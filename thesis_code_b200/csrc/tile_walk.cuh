// Work enumeration of the tile GEMM (csrc/gemm_dmma.cu), kept in a header so that the same code is exercised on the
// host by tests/test_tilewalk_cpu.py (every valid output tile of a job must be visited exactly once, whatever the grid).
#pragma once
#include "gpx_common.cuh"

#define GPX_GEMM_STRIP 16  // default output tile columns per rasterisation strip (GemmJob.strip = 0): 16 B-operand panels stay hot in
                           // L2 while the A row panels stream by once per strip (8 -> 16 measured +0.5 % at N = 60000, 32 slower)
__host__ __device__ __forceinline__ int gpx_job_strip(const GemmJob& j) {
    if (j.rowwise) return j.ncols;     // one strip of all columns
    return j.strip > 0 ? j.strip : GPX_GEMM_STRIP;
}
// contraction range of output tile column J
__host__ __device__ __forceinline__ int gpx_job_k1(const GemmJob& j, int J) { return (j.rowwise && J + 1 < j.k1) ? J + 1 : j.k1; }

// Output-tile enumeration shared by producer and consumers: strips of GPX_GEMM_STRIP column ordinals, row-major inside a
// strip, so CTAs running together share few operand panels (L2 reuse).  Invalid (above-diagonal) slots of a
// triangular job are skipped.  A CTA owns `chunk` consecutive slots; grids larger than the SM count are scheduled
// by the hardware as earlier CTAs retire, which balances the ragged triangular jobs and lets kernels of other
// streams (panel factorisation, NCCL) interleave.
struct TileWalk {
    int tri, i0, i1, ncols;
    int strip_m0, strip_w, strip_r0;
    long long strip_begin, strip_end;
    __host__ __device__ void set_strip(const GemmJob& j) {
        const int sw = gpx_job_strip(j);
        strip_w = (sw < ncols - strip_m0 ? sw : ncols - strip_m0);
        strip_r0 = tri ? (i0 > gpx_job_col(j, strip_m0) ? i0 : gpx_job_col(j, strip_m0)) : i0;
        if (strip_r0 > i1) strip_r0 = i1;
    }
    __host__ __device__ void init(const GemmJob& j) {
        tri = j.tri; i0 = j.i0; i1 = j.i1; ncols = j.ncols;
        strip_m0 = 0;
        set_strip(j);
        strip_begin = 0;
        strip_end = (long long)(i1 - strip_r0) * strip_w;
    }
    // t counts HALF-tile slots: tile slot t >> 1, row half t & 1.  Returns false when t is past the end.
    __host__ __device__ bool locate(const GemmJob& j, long long t2, int& I, int& J, int& half, bool& valid) {
        if (j.rowwise) {
            // panel-solve mode: slot order (row, half, column descending); a CTA's chunk of ncols slots is one half tile row
            const int w = ncols;
            if (t2 >= 2LL * (i1 - i0) * w) return false;
            const long long r = t2 / (2 * w);
            const int rem = (int)(t2 - r * 2 * w);
            half = rem / w;
            I = i0 + (int)r;
            J = gpx_job_col(j, w - 1 - (rem - half * w));
            valid = true;
            return true;
        }
        half = (int)(t2 & 1);
        const long long t = t2 >> 1;
        while (t >= strip_end) {
            strip_m0 += strip_w;
            if (strip_m0 >= ncols) return false;
            set_strip(j);
            strip_begin = strip_end;
            strip_end = strip_begin + (long long)(i1 - strip_r0) * strip_w;
        }
        long long u = t - strip_begin;
        I = strip_r0 + (int)(u / strip_w);
        J = gpx_job_col(j, strip_m0 + (int)(u % strip_w));
        valid = !tri || I >= J;
        return true;
    }
};


// ---- host-side counters used by the launcher ---------------------------------------------------------------------
inline long long count_valid_tiles(const GemmJob& j) {
    long long total = 0;
    for (int m = 0; m < j.ncols; m++) {
        int c = gpx_job_col(j, m);
        int r0 = j.tri ? (j.i0 > c ? j.i0 : c) : j.i0;
        if (r0 < j.i1) total += j.i1 - r0;
    }
    return total;
}

// (tile, k-tile) steps of a job: what the launcher reports as work (2 * 128^3 flop each)
inline long long count_steps(const GemmJob& j) {
    if (!j.rowwise) return count_valid_tiles(j) * (j.k1 - j.k0);
    long long per_row = 0;
    for (int m = 0; m < j.ncols; m++) { int J = gpx_job_col(j, m); int k1 = gpx_job_k1(j, J); if (k1 > j.k0) per_row += k1 - j.k0; }
    return per_row * (j.i1 - j.i0);
}

inline long long count_slots(const GemmJob& j) {
    long long total = 0;
    const int sw = gpx_job_strip(j);
    for (int m0 = 0; m0 < j.ncols; m0 += sw) {
        int w = (j.ncols - m0 < sw) ? (j.ncols - m0) : sw;
        int c0 = gpx_job_col(j, m0);
        int r0 = j.tri ? (j.i0 > c0 ? j.i0 : c0) : j.i0;
        if (r0 > j.i1) r0 = j.i1;
        total += (long long)(j.i1 - r0) * w;
    }
    return total;
}


// ---- slot -> CTA assignment ---------------------------------------------------------------------------------------
// Contiguous (interleave == 0): CTA b owns [b * chunk, (b + 1) * chunk).  Interleaved (interleave == R): the R CTAs of round
// g = b / R share the window [g R chunk, (g + 1) R chunk) and CTA b walks it from g R chunk + b % R with stride R.
__host__ __device__ __forceinline__ long long gpx_cta_first_slot(long long b, int chunk, int interleave) {
    return interleave > 0 ? (b / interleave) * (long long)interleave * chunk + (b % interleave) : b * chunk;
}
__host__ __device__ __forceinline__ long long gpx_cta_slot_stride(int interleave) { return interleave > 0 ? interleave : 1; }
// number of CTAs that covers `slots` half-tile slots
inline long long gpx_grid_for(long long slots, int chunk, int interleave) {
    if (interleave <= 0) return (slots + chunk - 1) / chunk;
    const long long window = (long long)interleave * chunk;
    const long long full = slots / window, rest = slots - full * window;
    return full * interleave + (rest > interleave ? interleave : rest);
}
