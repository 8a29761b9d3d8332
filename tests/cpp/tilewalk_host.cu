// Host-side check of the tile GEMM's work enumeration (thesis_code_b200/csrc/tile_walk.cuh): for many job shapes and
// chunk sizes, emulate every CTA's slot range and verify that each valid output half tile is produced exactly once and
// that nothing outside the job is touched.  Compiled with nvcc, runs on the CPU (no GPU needed).
#include <cstdio>
#include <map>
#include <utility>
#include <vector>
#include "../../thesis_code_b200/csrc/tile_walk.cuh"

static int check(const GemmJob& job, int chunk, const char* what, int interleave = 0) {
    const long long slots = 2 * count_slots(job);
    const long long grid = gpx_grid_for(slots, chunk, interleave);
    std::map<std::pair<int, int>, int> seen;   // (I, J) -> bitmask of halves
    long long visits = 0;
    for (long long b = 0; b < grid; b++) {
        TileWalk walk;
        walk.init(job);
        const long long t0 = gpx_cta_first_slot(b, chunk, interleave), st = gpx_cta_slot_stride(interleave);
        for (long long t = t0; t < t0 + (long long)chunk * st; t += st) {
            int I, J, half; bool valid;
            if (!walk.locate(job, t, I, J, half, valid)) break;
            if (!valid) continue;
            if (I < job.i0 || I >= job.i1) { printf("%s: row %d outside [%d,%d)\n", what, I, job.i0, job.i1); return 1; }
            int& m = seen[{I, J}];
            if (m & (1 << half)) { printf("%s: tile (%d,%d) half %d visited twice\n", what, I, J, half); return 1; }
            m |= 1 << half;
            visits++;
        }
    }
    // expected set
    long long expect = 0;
    for (int m = 0; m < job.ncols; m++) {
        const int J = gpx_job_col(job, m);
        for (int I = job.i0; I < job.i1; I++) {
            if (job.tri && I < J) continue;
            expect++;
            auto it = seen.find({I, J});
            if (it == seen.end() || it->second != 3) { printf("%s: tile (%d,%d) not fully covered\n", what, I, J); return 1; }
        }
    }
    if (expect != count_valid_tiles(job) || visits != 2 * expect || (long long)seen.size() != expect) {
        printf("%s: count mismatch expect %lld valid %lld visits %lld seen %zu\n", what, expect, count_valid_tiles(job), visits, seen.size());
        return 1;
    }
    return 0;
}

// "panel solve" jobs (GemmJob::rowwise): with chunk = ncols every CTA must own exactly one half tile row and visit its
// columns in descending order (the in-place update relies on both), and the contraction range must stop at the diagonal.
static int check_rowwise(int nb, int p0, int w) {
    GemmJob j{};
    j.tri = 0; j.i0 = 0; j.i1 = nb; j.ncols = w; j.jbase = p0; j.jW = GPX_JW_CONTIG; j.jstride = 0; j.k0 = p0; j.k1 = p0 + w; j.rowwise = 1;
    const int chunk = w;
    if (check(j, chunk, "rowwise")) return 1;
    const long long slots = 2 * count_slots(j);
    if (slots != 2LL * nb * w) { printf("rowwise: slot count\n"); return 1; }
    long long steps = 0;
    for (long long b = 0; b < slots / chunk; b++) {
        TileWalk walk; walk.init(j);
        int lastJ = 1 << 30, row = -1;
        for (long long t = b * chunk; t < (b + 1) * chunk; t++) {
            int I, J, half; bool valid;
            if (!walk.locate(j, t, I, J, half, valid) || !valid) { printf("rowwise: missing slot\n"); return 1; }
            if (row < 0) row = I;
            if (I != row || I != (int)(b / 2) || half != (int)(b & 1)) { printf("rowwise: CTA %lld touches row %d half %d\n", b, I, half); return 1; }
            if (J >= lastJ) { printf("rowwise: columns not strictly descending\n"); return 1; }
            lastJ = J;
            if (gpx_job_k1(j, J) != J + 1) { printf("rowwise: k range of column %d ends at %d\n", J, gpx_job_k1(j, J)); return 1; }
            if (half == 0) steps += gpx_job_k1(j, J) - j.k0;
        }
    }
    if (steps != count_steps(j) || steps != (long long)nb * w * (w + 1) / 2) { printf("rowwise: step count %lld vs %lld\n", steps, count_steps(j)); return 1; }
    return 0;
}

int main() {
    int bad = 0, cases = 0;
    for (int nb : {1, 2, 131, 148})
        for (int p0 : {0, 16, 37})
            for (int w : {1, 2, 4, 13, 16}) { bad += check_rowwise(nb, p0, w); cases++; }
    const int chunks[] = {1, 2, 3, 7, 8, 16, 32, 33};
    for (int chunk : chunks) {
        for (int nt : {1, 2, 5, 8, 9, 17, 40, 71}) {
            // triangular trailing update starting at column c0
            for (int c0 : {0, 1, 3, nt / 2, nt - 1}) {
                if (c0 < 0 || c0 >= nt) continue;
                GemmJob j{};
                j.tri = 1; j.i0 = c0; j.i1 = nt; j.ncols = nt - c0; j.jbase = c0; j.jW = GPX_JW_CONTIG; j.jstride = 0; j.k0 = 0; j.k1 = 1;
                bad += check(j, chunk, "tri-contig"); cases++;
                for (int R : {3, 8, 296}) { bad += check(j, chunk, "tri-interleaved", R); cases++; }
                // rows start below the first column (TRSM-like / inner update shapes)
                GemmJob r{};
                r.tri = 0; r.i0 = c0; r.i1 = nt; r.ncols = 1; r.jbase = c0; r.jW = GPX_JW_CONTIG; r.k0 = 0; r.k1 = 1;
                bad += check(r, chunk, "rect-1col"); cases++;
            }
            // rectangular (variance solve): nb test tiles x columns [p1, nt)
            for (int nb : {1, 3, 20, 148})
                for (int p1 : {0, 4, nt - 1}) {
                    if (p1 < 0 || p1 >= nt) continue;
                    GemmJob j{};
                    j.tri = 0; j.i0 = 0; j.i1 = nb; j.ncols = nt - p1; j.jbase = p1; j.jW = GPX_JW_CONTIG; j.k0 = 0; j.k1 = 4;
                    bad += check(j, chunk, "rect"); cases++;
                    bad += check(j, chunk, "rect-interleaved", 296); cases++;
                    bad += check(j, chunk, "rect-interleaved", 5); cases++;
                }
            // block-column-cyclic owned columns: panels of W columns, every G-th panel, starting at panel q
            for (int W : {1, 3, 4})
                for (int G : {2, 4, 8})
                    for (int q = 0; q < G && q * W < nt; q++) {
                        int ncols = 0;
                        for (int p = q; p * W < nt; p += G) ncols += ((p + 1) * W < nt ? (p + 1) * W : nt) - p * W;
                        GemmJob j{};
                        j.tri = 1; j.i0 = q * W; j.i1 = nt; j.ncols = ncols; j.jbase = q * W; j.jW = W; j.jstride = W * G; j.k0 = 0; j.k1 = W;
                        bad += check(j, chunk, "tri-cyclic"); cases++;
                        bad += check(j, chunk, "tri-cyclic-interleaved", 296); cases++;
                    }
        }
    }
    printf("%d cases, %d failures\n", cases, bad);
    return bad ? 1 : 0;
}
