// C ABI (include/gpx.h) and host-side orchestration of the exact-GP hot path.
//
// One handle = one process = one GPU.  The handle owns the scaled training inputs, its shard of the packed
// lower-triangular tile storage of Ky (factored in place into L), all inverted diagonal tiles, alpha, the predict
// workspaces, three CUDA streams and (world > 1) an NCCL communicator plus two panel receive buffers.
//
// Distribution: tile columns are grouped into outer panels of `outer_tiles` columns; panel p is
// owned by rank p % world (1-D block-column-cyclic).  world == 1 is the same code with every panel owned.
//
// Cholesky (right-looking, look-ahead depth 1), per panel p:
//   owner, stream F (high priority): POTRF(tile) -> TRSM (DMMA GEMM vs the inverted tile) -> rank-128 updates inside
//          the panel, column by column;
//   stream C (high priority): ncclBroadcast of the factored panel (+ its inverted diagonal tiles) from the owner;
//   every rank, stream U: (a) the update of the NEXT panel's columns if this rank owns it -- which releases stream F
//          to factor panel p+1 --, the forward-substitution steps of the alpha solve with the panel (fused: the panel
//          is read while it is hot), and (b) the SYRK/GEMM update of the rest of the owned columns, K = panel.
//   So panel p+1 is factored and broadcast while update (b) of panel p is still running.
// Backward substitution: row-oriented, z_J lives with the owner of column J; one 4 KB broadcast of alpha per panel.
// Predict: test tiles are sharded over ranks; the panels of L are re-broadcast (double buffered) while each rank runs
// its K*-build + blocked triangular solve + row sum-of-squares; results are combined with one small all-reduce.
#include "gpx_internal.cuh"

NcclApi& nccl() {
    static NcclApi api;
    if (api.tried) return api;
    api.tried = true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return api;
#define GPX_SYM(field, name) *(void**)(&api.field) = dlsym(api.lib, name)
    GPX_SYM(GetUniqueId, "ncclGetUniqueId");
    GPX_SYM(CommInitRank, "ncclCommInitRank");
    GPX_SYM(CommDestroy, "ncclCommDestroy");
    GPX_SYM(Broadcast, "ncclBroadcast");
    GPX_SYM(AllReduce, "ncclAllReduce");
    GPX_SYM(CommSplit, "ncclCommSplit");
    GPX_SYM(GroupStart, "ncclGroupStart");
    GPX_SYM(GroupEnd, "ncclGroupEnd");
    GPX_SYM(GetErrorString, "ncclGetErrorString");
#undef GPX_SYM
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Broadcast && api.AllReduce &&
             api.GroupStart && api.GroupEnd && api.GetErrorString;
    return api;
}


// every DMMA tile-GEMM launch goes through here (optionally bracketed by events for the roofline numbers)
void gemm(gpx_t* h, const GemmJob& job_in, cudaStream_t st) {
    GemmJob job = job_in;
    // Cap on tile slots per CTA.  With the interleaved slot assignment long chunks no longer buy L2 locality, and short CTAs let
    // the high-priority panel kernels of the look-ahead stream in as soon as they are ready: cap 16 -> 2 on one GPU is +1.8 % on
    // the C4 Cholesky, +4 % at N = 40000, +12 % at N = 20000 (profiles/r02_gemm_experiments.txt); 1 and 2 are equal within 0.1 %.
    job.chunk = h->gemm_max_chunk > 0 ? h->gemm_max_chunk : (h->world > 1 ? 4 : 2);
    job.strip = h->gemm_strip;
    job.interleave = h->gemm_interleave ? 0 : -1;
    // Only launches on the main (update) stream are timed and counted: the panel-factorisation GEMMs of stream F
    // overlap them, so adding their durations would count the same wall time twice.
    if (!h->gemm_timing || st != h->st) {
        long long steps = gpx_launch_gemm(job, h->num_sms, st, &h->launches);
        if (steps && st == h->st) { h->gemm_flops += 2.0 * 128 * 128 * 128 * (double)steps; h->gemm_launches++; }
        return;
    }
    if (h->gemm_ev_used + 2 > h->gemm_ev.size()) {
        size_t old = h->gemm_ev.size();
        h->gemm_ev.resize(old + 1024);
        for (size_t i = old; i < h->gemm_ev.size(); i++) cudaEventCreate(&h->gemm_ev[i]);
    }
    cudaEventRecord(h->gemm_ev[h->gemm_ev_used], st);
    long long steps = gpx_launch_gemm(job, h->num_sms, st, &h->launches);
    if (steps) {
        cudaEventRecord(h->gemm_ev[h->gemm_ev_used + 1], st);
        h->gemm_ev_used += 2;
        h->gemm_flops += 2.0 * 128 * 128 * 128 * (double)steps;
        h->gemm_launches++;
    }
}

// fold the recorded per-launch event pairs into gemm_ms (the main stream must be idle)
void gemm_collect(gpx_t* h) {
    for (size_t i = 0; i + 1 < h->gemm_ev_used; i += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, h->gemm_ev[i], h->gemm_ev[i + 1]) == cudaSuccess) h->gemm_ms += ms;
        else cudaGetLastError();
    }
    h->gemm_ev_used = 0;
}

// broadcast panel p (its packed columns and, optionally, its inverted diagonal tiles) from its owner, on stream C.
// During the factorisation (with_linv) the low-CTA communicator is used: an NCCL channel pins an SM for the whole
// transfer and the tile GEMM needs a whole SM's shared memory, so NCCL's default ~24 channels cost the overlapped
// trailing update ~20% on the ring's end ranks (profiles/r01_trace_4gpu.txt); 4 CTAs still move a panel in far less
// than one update step.  The predict sweep is bandwidth-bound and uses the default communicator.
int broadcast_panel(gpx_t* h, int p, bool with_linv) {
    ncclComm_t comm = with_linv ? h->comm_lo : h->comm;
    const int p0 = panel_begin(h, p), p1 = panel_end(h, p), root = panel_owner(h, p);
    const int64_t elems = panel_elems(h, p);
    double* buf = owns_panel(h, p) ? ptr<double>(h->L) + h->h_coloff[p0] : panel_recv_ptr(h, p);
    GPX_NCCL(h, nccl().GroupStart());
    ncclResult_t r1 = nccl().Broadcast(buf, buf, (size_t)elems, ncclDouble, root, comm, h->st_c);
    ncclResult_t r2 = ncclSuccess;
    if (with_linv) {
        double* li = ptr<double>(h->linv) + (int64_t)p0 * GPX_TILE_ELEMS;
        r2 = nccl().Broadcast(li, li, (size_t)(p1 - p0) * GPX_TILE_ELEMS, ncclDouble, root, comm, h->st_c);
    }
    GPX_NCCL(h, nccl().GroupEnd());
    GPX_NCCL(h, r1);
    GPX_NCCL(h, r2);
    return 0;
}

void trace_rec(gpx_t* h, int p, int which, cudaStream_t st) {
    if (!h->trace) return;
    cudaEventRecord(h->tev[1 + (size_t)6 * p + which], st);
}

// factor the tile columns [c0, c1) in place (rows up to nt) on stream st: POTRF + inverse of the diagonal tile, TRSM of
// the column as a DMMA GEMM against the inverted tile, rank-128 update of the remaining columns of the range
void factor_cols(gpx_t* h, int c0, int c1, int nt, cudaStream_t st) {
    TileMat L = lmat(h), Li = linvmat(h);
    int* info = ptr<int>(h->info);
    for (int k = c0; k < c1; k++) {
        double* diag = ptr<double>(h->L) + h->h_coloff[k];
        gpx_launch_potrf_tile(diag, ptr<double>(h->linv) + (int64_t)k * GPX_TILE_ELEMS, info, k, st);
        h->launches++;
        if (k + 1 < nt) {
            GemmJob t{};  // TRSM: L(I,k) <- L(I,k) * Linv_k^T for I > k
            t.A = L; t.B = Li; t.C = L;
            t.tri = 0; t.i0 = k + 1; t.i1 = nt; t.ncols = 1; t.jbase = k; t.jW = GPX_JW_CONTIG; t.jstride = 0;
            t.k0 = k; t.k1 = k + 1; t.bdiag = 1; t.overwrite = 1;
            gemm(h, t, st);
        }
        if (k + 1 < c1) {
            GemmJob u{};  // update the remaining columns of this range with column k
            u.A = L; u.B = L; u.C = L;
            u.tri = 1; u.i0 = k + 1; u.i1 = nt; u.ncols = c1 - (k + 1); u.jbase = k + 1; u.jW = GPX_JW_CONTIG; u.jstride = 0;
            u.k0 = k; u.k1 = k + 1; u.bdiag = 0; u.overwrite = 0;
            gemm(h, u, st);
        }
    }
}

// factor panel p in place on its owner (stream F)
void factor_panel(gpx_t* h, int p) { factor_cols(h, panel_begin(h, p), panel_end(h, p), h->nt, h->st_f); }

// Blocked right-looking Cholesky with look-ahead, grouped ("delayed") far updates and fused forward substitution.
// b (work1) is consumed, z -> work2.
//
// Panels (W tile columns, the ownership unit) are taken in groups of M.  For panel p of group g:
//   stream F (owner of p): wait for the far update of group g-1 on p's columns, apply the NEAR updates from the
//       earlier panels of the same group (K = W each), factor the panel;
//   stream C: broadcast the factored panel into this group's slot of the group buffer (world > 1);
//   stream U: forward-substitution steps with the panel; when the group is complete, the FAR update of all owned
//       columns of later groups with K = M*W (one DMMA GEMM with four times the depth of a per-panel update: fewer
//       passes over C, higher tile-GEMM efficiency) -- the owned panels of the NEXT group first, which is what stream
//       F of their owners waits for, then the bulk.
// M = 1 on one GPU (there W itself is made large); M = option "group_panels" when sharded: W stays 4 (fine-grained
// ownership, short critical path) while the bulk update can run with K = 4 M tiles.  Measured on 2 and 4 B200s the
// deeper update did not pay (the per-panel update is already >= 33 TFLOP/s), so the default is M = 1.
int cholesky(gpx_t* h) {
    const int nt = h->nt, np = h->npanels, G = h->world, W = h->W;
    const int M = (G > 1) ? (h->group_panels < 1 ? 1 : h->group_panels) : 1;
    const int max_ctas = 2 * h->num_sms;
    int rc = ensure_panel_events(h, np);
    if (rc) return rc;
    if (h->trace) {
        while (h->tev.size() < 1 + (size_t)6 * np) { cudaEvent_t e; cudaEventCreate(&e); h->tev.push_back(e); }
        cudaEventRecord(h->tev[0], h->st);
        for (int p = 0; p < np; p++) for (int w = 0; w < 6; w++) cudaEventRecord(h->tev[1 + (size_t)6 * p + w], h->st);
    }
    auto group_first = [&](int g) { return g * M; };
    auto group_last = [&](int g) { int e = (g + 1) * M; return (e < np ? e : np) - 1; };
    // operand view of group g / receive slot of panel p.  With M == 1 (default) a panel is used in place on its owner,
    // lives in its resident cache slot when cached, and otherwise in one of the two rotating buffers; with M > 1 the
    // whole group (own panels included) is assembled contiguously in a rotating group buffer.
    auto group_mat = [&](int g) -> TileMat {
        if (G == 1) return lmat(h);
        if (M == 1) {
            if (owns_panel(h, g)) return lmat(h);
            if (panel_cached(h, g)) return umat(h);
        }
        return TileMat{ptr<double>(h->gbuf[g & 1]) - h->h_vcoloff[panel_begin(h, group_first(g))], ptr<int64_t>(h->vcoloff), 1};
    };
    auto group_slot = [&](int p) -> double* {
        const int g = p / M;
        if (M == 1) {
            if (owns_panel(h, p)) return ptr<double>(h->L) + h->h_coloff[panel_begin(h, p)];
            if (panel_cached(h, p)) return ptr<double>(h->L) + h->h_ucoloff[panel_begin(h, p)];
        }
        return ptr<double>(h->gbuf[g & 1]) + (h->h_vcoloff[panel_begin(h, p)] - h->h_vcoloff[panel_begin(h, group_first(g))]);
    };
    // update the columns of owned panel q (rows >= its diagonal) with the tile columns [k0, k1) of operand P
    auto update_panel_cols = [&](TileMat P, int q, int k0, int k1, cudaStream_t st) {
        GemmJob a{};
        a.A = P; a.B = P; a.C = lmat(h);
        a.tri = 1; a.i0 = panel_begin(h, q); a.i1 = nt; a.ncols = panel_end(h, q) - panel_begin(h, q); a.jbase = panel_begin(h, q);
        a.jW = GPX_JW_CONTIG; a.jstride = 0; a.k0 = k0; a.k1 = k1; a.bdiag = 0; a.overwrite = 0;
        gemm(h, a, st);
    };
    cudaEvent_t ev_start;  // everything enqueued on the main stream so far (K-build, residual) is done
    GPX_CUDA(h, cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    GPX_CUDA(h, cudaEventRecord(ev_start, h->st));
    GPX_CUDA(h, cudaStreamWaitEvent(h->st_f, ev_start, 0));
    GPX_CUDA(h, cudaStreamWaitEvent(h->st_c, ev_start, 0));
    for (int p = 0; p < np; p++) {
        const int p0 = panel_begin(h, p), p1 = panel_end(h, p);
        const int g = p / M, gf = group_first(g), gl = group_last(g);
        const bool own = owns_panel(h, p);
        TileMat P = group_mat(g);
        if (own) {
            // ---- stream F: everything about my own panel inside the current group ----
            if (g > 0) GPX_CUDA(h, cudaStreamWaitEvent(h->st_f, pevent(h, group_first(g - 1), 1), 0));  // far-(a) of group g-1
            for (int pp = gf; pp < p; pp++) {  // near updates from the earlier panels of this group
                GPX_CUDA(h, cudaStreamWaitEvent(h->st_f, pevent(h, pp, 0), 0));
                update_panel_cols(P, p, panel_begin(h, pp), panel_end(h, pp), h->st_f);
            }
            trace_rec(h, p, 0, h->st_f);
            factor_panel(h, p);
            trace_rec(h, p, 1, h->st_f);
            GPX_CUDA(h, cudaEventRecord(pevent(h, p, 3), h->st_f));
        }
        if (G > 1) {
            // ---- stream C: panel p -> this group's slot of the group buffer on every rank ----
            if (own) GPX_CUDA(h, cudaStreamWaitEvent(h->st_c, pevent(h, p, 3), 0));
            if (p == gf && g >= 2) GPX_CUDA(h, cudaStreamWaitEvent(h->st_c, pevent(h, group_first(g - 2), 2), 0));  // buffer free
            const int root = panel_owner(h, p);
            double* slot = group_slot(p);
            const double* src = own ? ptr<double>(h->L) + h->h_coloff[p0] : slot;
            double* li = ptr<double>(h->linv) + (int64_t)p0 * GPX_TILE_ELEMS;
            // Early panels go over the low-CTA communicator (an NCCL channel pins an SM for the whole transfer and the overlapped
            // trailing update wants every SM).  Late in the factorisation the update of the owned columns is shorter than
            // factor + broadcast of the next panel and the update stream waits for the panel (profiles/r02_trace_8gpu.txt: ~3 ms
            // per panel around 75-85 % of the way at 8 GPUs, N = 200k): from there on the full-width communicator is used.
            ncclComm_t bc = (100LL * p >= (long long)h->nccl_full_from_pct * np) ? h->comm : h->comm_lo;
            GPX_NCCL(h, nccl().GroupStart());
            ncclResult_t r1 = nccl().Broadcast(src, slot, (size_t)panel_elems(h, p), ncclDouble, root, bc, h->st_c);
            ncclResult_t r2 = nccl().Broadcast(li, li, (size_t)(p1 - p0) * GPX_TILE_ELEMS, ncclDouble, root, bc, h->st_c);
            GPX_NCCL(h, nccl().GroupEnd());
            GPX_NCCL(h, r1);
            GPX_NCCL(h, r2);
            GPX_CUDA(h, cudaEventRecord(pevent(h, p, 0), h->st_c));
        } else {
            GPX_CUDA(h, cudaEventRecord(pevent(h, p, 0), h->st_f));
        }
        // ---- stream U: consume panel p ----
        GPX_CUDA(h, cudaStreamWaitEvent(h->st, pevent(h, p, 0), 0));
        trace_rec(h, p, 2, h->st);
        const int k0 = panel_begin(h, gf), k1 = panel_end(h, gl);   // the whole group as contraction range
        if (p == gl) {
            // far-(a): my panels of the NEXT group first -- their owners' stream F waits for this
            for (int q = next_owned_panel(h, (g + 1) * M); q < np && q <= group_last(g + 1); q += G)
                update_panel_cols(P, q, k0, k1, h->st);
            GPX_CUDA(h, cudaEventRecord(pevent(h, gf, 1), h->st));
        }
        trace_rec(h, p, 3, h->st);
        // fused forward substitution with the columns of this panel (replicated right-hand side).  On one GPU it runs on its own
        // stream beside the trailing update (one small launch per tile column: 3 % of the factorisation at N = 20000 when it
        // sat in the update stream); sharded, it stays in the update stream, which is what guards the rotating panel buffers.
        cudaStream_t fst = (G == 1) ? h->st_c : h->st;
        if (G == 1) GPX_CUDA(h, cudaStreamWaitEvent(fst, pevent(h, p, 0), 0));
        for (int J = p0; J < p1; J++) {
            gpx_launch_fwd_step(P, ptr<double>(h->linv), nt, J, ptr<double>(h->work1), ptr<double>(h->work2), h->Np, h->O,
                                max_ctas, fst);
            h->launches++;
        }
        trace_rec(h, p, 4, fst);
        if (p == gl) {
            // far-(b): the rest of the owned columns (groups >= g+2), K = the whole group
            const int qb = next_owned_panel(h, (g + 2) * M);
            if (qb < np) {
                GemmJob b{};
                b.A = P; b.B = P; b.C = lmat(h);
                b.tri = 1; b.i0 = panel_begin(h, qb); b.i1 = nt; b.ncols = owned_cols_from(h, qb); b.jbase = panel_begin(h, qb);
                b.jW = W; b.jstride = W * G; b.k0 = k0; b.k1 = k1; b.bdiag = 0; b.overwrite = 0;
                gemm(h, b, h->st);
            }
            GPX_CUDA(h, cudaEventRecord(pevent(h, gf, 2), h->st));
        }
        trace_rec(h, p, 5, h->st);
        if (!h->lookahead) {  // debugging aid: serialise the three streams
            cudaEvent_t e = pevent(h, p, 3);   // re-used as a scratch event: panel p's own "factored" role is over
            GPX_CUDA(h, cudaEventRecord(e, h->st));
            GPX_CUDA(h, cudaStreamWaitEvent(h->st_f, e, 0));
            GPX_CUDA(h, cudaStreamWaitEvent(h->st_c, e, 0));
        }
    }
    // join F and C into the main stream
    cudaEvent_t ev_f, ev_c;
    GPX_CUDA(h, cudaEventCreateWithFlags(&ev_f, cudaEventDisableTiming));
    GPX_CUDA(h, cudaEventCreateWithFlags(&ev_c, cudaEventDisableTiming));
    GPX_CUDA(h, cudaEventRecord(ev_f, h->st_f));
    GPX_CUDA(h, cudaEventRecord(ev_c, h->st_c));
    GPX_CUDA(h, cudaStreamWaitEvent(h->st, ev_f, 0));
    GPX_CUDA(h, cudaStreamWaitEvent(h->st, ev_c, 0));
    cudaEventDestroy(ev_start);
    cudaEventDestroy(ev_f);
    cudaEventDestroy(ev_c);
    return 0;
}

// backward substitution L^T alpha = z (z in work2, replicated after the fused forward pass)
int backward_solve(gpx_t* h) {
    const int np = h->npanels, G = h->world, W = h->W;
    const int max_ctas = 2 * h->num_sms;
    const int q0 = next_owned_panel(h, 0);
    const int ncols_owned = q0 < np ? owned_cols_from(h, q0) : 0;
    const int jbase = q0 < np ? panel_begin(h, q0) : 0;
    for (int p = np - 1; p >= 0; p--) {
        const int p0 = panel_begin(h, p), p1 = panel_end(h, p);
        const bool own = owns_panel(h, p);
        if (!own) {
            for (int o = 0; o < h->O; o++) {
                double* a = ptr<double>(h->alpha) + (int64_t)o * h->Np + (int64_t)p0 * GPX_TILE;
                GPX_NCCL(h, nccl().Broadcast(a, a, (size_t)(p1 - p0) * GPX_TILE, ncclDouble, panel_owner(h, p), h->comm_lo, h->st));
            }
        }
        for (int I = p1 - 1; I >= p0; I--) {
            gpx_launch_bwd_step(lmat(h), ptr<double>(h->linv), I, ptr<double>(h->work2), ptr<double>(h->alpha), h->Np, h->O,
                                own ? 1 : 0, ncols_owned, jbase, W, W * G, max_ctas, h->st);
            h->launches++;
        }
        if (own && G > 1) {
            for (int o = 0; o < h->O; o++) {
                double* a = ptr<double>(h->alpha) + (int64_t)o * h->Np + (int64_t)p0 * GPX_TILE;
                GPX_NCCL(h, nccl().Broadcast(a, a, (size_t)(p1 - p0) * GPX_TILE, ncclDouble, h->rank, h->comm_lo, h->st));
            }
        }
    }
    return 0;
}



extern "C" {

int gpx_version(void) { return 200; }

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
    if (prop.major != 10) {  // sm_100a-only library: refuse loudly rather than fall back
        delete h;
        return -5;
    }
    h->num_sms = prop.multiProcessorCount;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&h->st, cudaStreamNonBlocking, lo) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->st_f, cudaStreamNonBlocking, hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&h->st_c, cudaStreamNonBlocking, hi) != cudaSuccess) { delete h; return -2; }
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
    cudaStreamSynchronize(h->st_f);
    cudaStreamSynchronize(h->st_c);
    if (h->comm_lo && h->comm_lo != h->comm) nccl().CommDestroy(h->comm_lo);
    if (h->comm) nccl().CommDestroy(h->comm);
    DevBuf* all[] = {&h->Xs, &h->xsq, &h->ls, &h->L, &h->linv, &h->coloff, &h->vcoloff, &h->linv_coloff, &h->resid,
                     &h->work1, &h->work2, &h->alpha, &h->scal, &h->info, &h->stage, &h->gbuf[0], &h->gbuf[1], &h->cache,
                     &h->ucoloff, &h->Tt,
                     &h->tt_coloff, &h->Xt, &h->xtsq, &h->mean_partial, &h->dmean, &h->dvar, &h->sp, &h->xstats, &h->ystats, &h->var_partial, &h->minv, &h->minv_coloff,
                     &h->ywork, &h->ywork_coloff, &h->zkeep, &h->finv, &h->finv_coloff, &h->finv_partial};
    for (DevBuf* b : all) release(h, *b);
    release_scratch(h);
    for (auto& g : h->graphs) cudaGraphExecDestroy(g.exec);
    if (h->pin) cudaFreeHost(h->pin);
    if (h->ev_caller) cudaEventDestroy(h->ev_caller);
    for (auto& e : h->ev) cudaEventDestroy(e);
    for (auto& e : h->gemm_ev) cudaEventDestroy(e);
    for (auto& e : h->pev) cudaEventDestroy(e);
    for (auto& e : h->tev) cudaEventDestroy(e);
    cudaStreamDestroy(h->st);
    cudaStreamDestroy(h->st_f);
    cudaStreamDestroy(h->st_c);
    delete h;
}

const char* gpx_last_error(const gpx_t* h) { return h ? h->err.c_str() : "null handle"; }

int gpx_comm_unique_id(char* out128) {
    if (!out128) return -1;
    if (!nccl().ok) return -6;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return -7;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is expected to be 128 bytes");
    memcpy(out128, &id, 128);
    return 0;
}

int gpx_comm_init(gpx_t* h, int rank, int world, const char* uid128) {
    if (!h) return -1;
    if (world < 1 || rank < 0 || rank >= world) return fail(h, -1, "bad rank/world %d/%d", rank, world);
    if (h->comm_lo && h->comm_lo != h->comm) nccl().CommDestroy(h->comm_lo);
    h->comm_lo = nullptr;
    if (h->comm) { nccl().CommDestroy(h->comm); h->comm = nullptr; }
    h->fitted = false;
    h->rank = rank;
    h->world = world;
    if (world == 1) return 0;
    if (!uid128) return fail(h, -1, "null unique id");
    if (!nccl().ok) return fail(h, -6, "libnccl.so.2 could not be loaded (dlopen); multi-GPU path unavailable");
    GPX_CUDA(h, cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, uid128, 128);
    GPX_NCCL(h, nccl().CommInitRank(&h->comm, world, id, rank));
    h->comm_lo = h->comm;
    if (nccl().CommSplit && h->nccl_lo_ctas > 0) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.minCTAs = 1;
        cfg.maxCTAs = h->nccl_lo_ctas;
        ncclComm_t lo = nullptr;
        if (nccl().CommSplit(h->comm, 0, rank, &lo, &cfg) == ncclSuccess && lo) h->comm_lo = lo;
    }
    return 0;
}

int gpx_set_option(gpx_t* h, const char* name, int64_t value) {
    if (!h || !name) return -1;
    if (!strcmp(name, "outer_tiles")) { if (value < 0 || value > 64) return fail(h, -1, "outer_tiles out of range"); h->outer_tiles = (int)value; return 0; }
    if (!strcmp(name, "predict_batch_tiles")) { if (value < 0) return fail(h, -1, "bad predict_batch_tiles"); h->predict_batch_tiles = (int)value; return 0; }
    if (!strcmp(name, "sync_phases")) { h->sync_phases = value ? 1 : 0; return 0; }
    if (!strcmp(name, "gemm_timing")) { h->gemm_timing = value ? 1 : 0; return 0; }
    if (!strcmp(name, "lookahead")) { h->lookahead = value ? 1 : 0; return 0; }
    if (!strcmp(name, "cache_mb")) { h->cache_limit_mb = value; h->fitted = false; return 0; }
    if (!strcmp(name, "group_panels")) { if (value < 1 || value > 16) return fail(h, -1, "group_panels out of range"); h->group_panels = (int)value; h->fitted = false; return 0; }
    if (!strcmp(name, "nccl_lo_ctas")) { h->nccl_lo_ctas = (int)value; return 0; }  // takes effect at gpx_comm_init
    if (!strcmp(name, "nccl_full_from_pct")) { h->nccl_full_from_pct = (int)value; return 0; }  // 100 = low-CTA communicator throughout
    if (!strcmp(name, "trace")) { h->trace = value ? 1 : 0; return 0; }
    if (!strcmp(name, "gemm_max_chunk")) { h->gemm_max_chunk = (int)value; return 0; }
    if (!strcmp(name, "gemm_interleave")) { h->gemm_interleave = value ? 1 : 0; return 0; }
    if (!strcmp(name, "panel_inverse")) { h->panel_inverse = (int)value; return 0; }   // -1 auto, 0 off, 1 on
    if (!strcmp(name, "gemm_strip")) { if (value < 0 || value > 64) return fail(h, -1, "gemm_strip out of range"); h->gemm_strip = (int)value; return 0; }
    if (!strcmp(name, "reserve_rows")) { if (value < 0) return fail(h, -1, "bad reserve_rows"); h->reserve_rows = value; return 0; }
    if (!strcmp(name, "predict_graph")) { h->predict_graph = value ? 1 : 0; h->gen++; return 0; }
    if (!strcmp(name, "latency_inverse")) {   // -1 auto, 0 off (drops a built inverse), 1 build at the first small variance predict
        h->latency_inverse = (int)value;
        if (value == 0) {   // off: stop using a built inverse and give its memory back
            if (h->finv_valid) { h->finv_valid = false; h->gen++; }
            cudaSetDevice(h->device);
            cudaStreamSynchronize(h->st);
            release(h, h->finv); release(h, h->finv_partial);
            h->finv_layout_cap = -1;
        }
        h->finv_failed = false;
        return 0;
    }
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
    cudaStreamSynchronize(h->st_f);
    gemm_collect(h);
    if (total_ms) *total_ms = h->gemm_ms;
    if (issued_flops) *issued_flops = h->gemm_flops;
    if (launches) *launches = h->gemm_launches;
    h->gemm_ms = 0; h->gemm_flops = 0; h->gemm_launches = 0;
    return 0;
}
int64_t gpx_bytes_allocated(const gpx_t* h) { return h ? h->bytes_alloc : -1; }
int64_t gpx_trace(const gpx_t* h, double* out, int64_t capacity) {
    if (!h) return -1;
    int64_t n = (int64_t)h->trace_ms.size();
    if (out) for (int64_t i = 0; i < n && i < capacity; i++) out[i] = h->trace_ms[i];
    return n;
}

int gpx_fit(gpx_t* h, const double* X, int64_t N, int d, const double* Y, int O, int kernel_id, double variance,
            const double* lengthscale, int n_ls, int ard, double noise, double jitter, double mean_const, int mem) {
    if (!h) return -1;
    if (!X || !Y || !lengthscale) return fail(h, -1, "null argument");
    if (N < 1 || d < 1 || O < 1) return fail(h, -1, "bad shape N=%lld d=%d O=%d", (long long)N, d, O);
    if (d > gpx_kbuild_max_d()) return fail(h, -1, "input dimension %d exceeds the supported maximum %d", d, gpx_kbuild_max_d());
    if (kernel_id < 0 || kernel_id > 3) return fail(h, -1, "unknown kernel id %d", kernel_id);
    if (ard ? (n_ls != d) : (n_ls != 1))
        return fail(h, -1, "lengthscale must have d entries with ARD and 1 entry without (got %d, d=%d, ARD=%d)", n_ls, d, ard);
    if (!(variance > 0) || !(noise + jitter > 0)) return fail(h, -1, "variance and noise+jitter must be positive");
    for (int i = 0; i < n_ls; i++) if (!(lengthscale[i] > 0)) return fail(h, -1, "lengthscale must be positive");
    GPX_CUDA(h, cudaSetDevice(h->device));
    h->fitted = false;
    { int rcw = wait_for_caller(h); if (rcw) return rcw; }
    const int nt = (int)((N + GPX_TILE - 1) / GPX_TILE);
    // Capacity for gpx_append (single-GPU handles): the column-packed factor and the [o][Np] vectors are laid out for
    // cap_nt >= nt tile rows, so that rows can be appended in place until the reserve is used up (then gpx_append
    // re-lays the storage out).  Np is the padded vector length / stride; rows >= N hold zeros (identity in L).
    int cap_nt = nt;
    if (h->world == 1 && h->reserve_rows > 0) cap_nt = (int)((N + h->reserve_rows + GPX_TILE - 1) / GPX_TILE);
    const int64_t Np = (int64_t)cap_nt * GPX_TILE;
    h->cap_nt = cap_nt;
    h->gen++;   // invalidates captured predict graphs
    h->minv_valid = 0;
    finv_invalidate(h);
    h->N = N; h->Np = Np; h->nt = nt; h->d = d; h->O = O; h->kernel_id = kernel_id; h->n_ls = n_ls;
    h->variance = variance; h->noise = noise; h->jitter = jitter; h->mean_const = mean_const;
    h->ard = ard ? 1 : 0;
    h->rdiv = ard ? 1.0 : lengthscale[0];
    h->h_ls.assign(lengthscale, lengthscale + n_ls);
    h->fit_norm_x = h->norm_x; h->fit_norm_y = h->norm_y;
    if (h->outer_tiles > 0) h->W = h->outer_tiles;
    else if (h->world > 1) h->W = 4;
    else { h->W = nt / 24; if (h->W < 4) h->W = 4; if (h->W > 16) h->W = 16; }
    h->npanels = (nt + h->W - 1) / h->W;
    const int G = h->world;
    for (auto& m : h->phase_ms) m = 0.f;

    // packed lower-triangular tile storage: only the columns of the panels this rank owns
    h->h_coloff.assign(cap_nt, -1);
    h->h_vcoloff.resize(nt);
    std::vector<int64_t> lc(cap_nt);
    int64_t off = 0, voff = 0, max_panel = 0;
    for (int c = 0; c < cap_nt; c++) {
        if (c < nt) { h->h_vcoloff[c] = voff; voff += (int64_t)(nt - c) * GPX_TILE_ELEMS; }
        lc[c] = (int64_t)c * GPX_TILE_ELEMS;
        if (panel_owner(h, c / h->W) == h->rank) { h->h_coloff[c] = off; off += (int64_t)(cap_nt - c) * GPX_TILE_ELEMS; }
    }
    for (int p = 0; p < h->npanels; p++) { int64_t e = panel_elems(h, p); if (e > max_panel) max_panel = e; }
    int rc;
    if ((rc = ensure(h, h->L, (size_t)(off > 0 ? off : 1) * 8))) {
        // out of memory: give back everything that is only a cache or a workspace of earlier predicts / appends / larger
        // fits (they are rebuilt on demand) and try once more
        for (DevBuf* b : {&h->Tt, &h->mean_partial, &h->sp, &h->minv, &h->ywork, &h->var_partial, &h->dmean, &h->dvar, &h->Xt,
                          &h->xtsq, &h->zkeep, &h->cache, &h->gbuf[0], &h->gbuf[1], &h->stage, &h->finv, &h->finv_partial})
            release(h, *b);
        release_scratch(h);
        h->finv_layout_cap = -1;
        h->sp_key[0] = -1; h->minv_valid = 0; h->minv_layout_cap = -1; h->zkeep_valid = false;
        if ((rc = ensure(h, h->L, (size_t)(off > 0 ? off : 1) * 8))) return rc;
    }
    if ((rc = ensure(h, h->linv, (size_t)cap_nt * GPX_TILE_BYTES))) return rc;
    if ((rc = ensure(h, h->coloff, (size_t)cap_nt * 8))) return rc;
    if ((rc = ensure(h, h->vcoloff, (size_t)nt * 8))) return rc;
    if ((rc = ensure(h, h->linv_coloff, (size_t)cap_nt * 8))) return rc;
    if ((rc = ensure(h, h->Xs, (size_t)Np * d * 8))) return rc;
    if ((rc = ensure(h, h->xsq, (size_t)Np * 8))) return rc;
    if ((rc = ensure(h, h->ls, (size_t)n_ls * 8))) return rc;
    if ((rc = ensure(h, h->resid, (size_t)Np * O * 8))) return rc;
    if ((rc = ensure(h, h->work1, (size_t)Np * O * 8))) return rc;
    if ((rc = ensure(h, h->work2, (size_t)Np * O * 8))) return rc;
    if ((rc = ensure(h, h->alpha, (size_t)Np * O * 8))) return rc;
    if ((rc = ensure(h, h->scal, 64))) return rc;
    if ((rc = ensure(h, h->info, 64))) return rc;
    if (G > 1) {
        const int Mg = h->group_panels < 1 ? 1 : h->group_panels;
        int64_t group_elems = 0;   // the first group is the largest
        for (int p = 0; p < Mg && p < h->npanels; p++) group_elems += panel_elems(h, p);
        if (group_elems < max_panel) group_elems = max_panel;
        if ((rc = ensure(h, h->gbuf[0], (size_t)group_elems * 8))) return rc;
        if ((rc = ensure(h, h->gbuf[1], (size_t)group_elems * 8))) return rc;
    }
    // Panel cache (world > 1, M == 1): keep as many received panels [0, cache_panels) resident as fit beside a reserve
    // for the predict workspaces, so that predict sweeps do not re-broadcast them.  Every rank must take the same
    // decision: the budget is the minimum over ranks and the panel count is derived from sizes all ranks know.
    h->cache_panels = 0;
    h->h_ucoloff.assign(nt, -1);
    if (G > 1) {
        if ((rc = ensure(h, h->ucoloff, (size_t)nt * 8))) return rc;
        const int Mg = h->group_panels < 1 ? 1 : h->group_panels;
        double budget = 0.0;
        if (Mg == 1 && h->cache_limit_mb != 0) {
            size_t fr = 0, tot = 0;
            GPX_CUDA(h, cudaMemGetInfo(&fr, &tot));
            fr += h->cache.bytes;   // reusable
            // reserve for the predict workspaces (their batch size adapts to what is left), NCCL and the caller.  Round 1 kept
            // 35 % free, which left 10-30 % of L to be re-broadcast in every predict sweep at N = 200k; 10 % (>= 8 GB) lets every
            // rank keep a full replica of the factor there (160 GB of the 180 GB), so the sweep runs without any exchange.
            double reserve = 0.10 * (double)fr > 8e9 ? 0.10 * (double)fr : 8e9;
            budget = (double)fr > reserve ? (double)fr - reserve : 0.0;
            if (h->cache_limit_mb > 0 && budget > 1e6 * (double)h->cache_limit_mb) budget = 1e6 * (double)h->cache_limit_mb;
        }
        double* ds = ptr<double>(h->scal) + 5;
        GPX_CUDA(h, cudaMemcpyAsync(ds, &budget, 8, cudaMemcpyHostToDevice, h->st));
        GPX_NCCL(h, nccl().AllReduce(ds, ds, 1, ncclDouble, ncclMin, h->comm, h->st));
        GPX_CUDA(h, cudaMemcpyAsync(&budget, ds, 8, cudaMemcpyDeviceToHost, h->st));
        GPX_CUDA(h, cudaStreamSynchronize(h->st));
        // largest P such that every rank's non-owned panels below P fit the common budget
        std::vector<double> need(G, 0.0);
        int P = 0;
        for (; P < h->npanels; P++) {
            const double bytes = 8.0 * (double)panel_elems(h, P);
            bool ok = true;
            for (int r = 0; r < G; r++) if (r != panel_owner(h, P) && need[r] + bytes > budget) ok = false;
            if (!ok) break;
            for (int r = 0; r < G; r++) if (r != panel_owner(h, P)) need[r] += bytes;
        }
        h->cache_panels = P;
        int64_t coff = 0;
        for (int pp = 0; pp < P; pp++) if (!owns_panel(h, pp)) coff += panel_elems(h, pp);
        if ((rc = ensure(h, h->cache, (size_t)(coff > 0 ? coff : 1) * 8))) return rc;
        const int64_t base_delta = (int64_t)(ptr<double>(h->cache) - ptr<double>(h->L));   // in elements
        coff = 0;
        for (int c = 0; c < nt; c++) {
            const int pp = c / h->W;
            if (panel_owner(h, pp) == h->rank) h->h_ucoloff[c] = h->h_coloff[c];
            else if (pp < P) { h->h_ucoloff[c] = base_delta + coff; coff += (int64_t)(nt - c) * GPX_TILE_ELEMS; }
        }
        GPX_CUDA(h, cudaMemcpyAsync(h->ucoloff.p, h->h_ucoloff.data(), (size_t)nt * 8, cudaMemcpyHostToDevice, h->st));
    }
    GPX_CUDA(h, cudaMemcpyAsync(h->coloff.p, h->h_coloff.data(), (size_t)cap_nt * 8, cudaMemcpyHostToDevice, h->st));
    GPX_CUDA(h, cudaMemcpyAsync(h->vcoloff.p, h->h_vcoloff.data(), (size_t)nt * 8, cudaMemcpyHostToDevice, h->st));
    GPX_CUDA(h, cudaMemcpyAsync(h->linv_coloff.p, lc.data(), (size_t)cap_nt * 8, cudaMemcpyHostToDevice, h->st));
    GPX_CUDA(h, cudaMemcpyAsync(h->ls.p, lengthscale, (size_t)n_ls * 8, cudaMemcpyHostToDevice, h->st));
    GPX_CUDA(h, cudaMemsetAsync(h->info.p, 0, 64, h->st));
    GPX_CUDA(h, cudaStreamSynchronize(h->st));  // lc / lengthscale are host temporaries

    // inputs
    phase_begin(h, GPX_PHASE_H2D);
    const double *dX, *dY;
    DevBuf& stageY = h->stage_y;   // persistent: a per-fit cudaMalloc + cudaFree showed in small-N optimiser loops
    if ((rc = to_device(h, X, (size_t)N * d, mem, h->stage, &dX))) return rc;
    if ((rc = to_device(h, Y, (size_t)N * O, mem, stageY, &dY))) return rc;
    phase_end(h, GPX_PHASE_H2D);

    // K-build: only the owned panels' columns
    phase_begin(h, GPX_PHASE_KBUILD);
    // fused NormalizerModel: column means / population stds of the raw X and Y, computed here and applied by the loaders
    if (h->fit_norm_x) {
        if ((rc = ensure(h, h->xstats, (size_t)2 * d * 8))) return rc;
        gpx_launch_col_stats(dX, N, d, ptr<double>(h->xstats), h->st);
        h->launches++;
    }
    if (h->fit_norm_y) {
        if ((rc = ensure(h, h->ystats, (size_t)2 * O * 8))) return rc;
        gpx_launch_col_stats(dY, N, O, ptr<double>(h->ystats), h->st);
        h->launches++;
    }
    prep_inputs(h, dX, N, Np, ptr<double>(h->Xs), ptr<double>(h->xsq), h->st);
    gpx_launch_resid(dY, N, Np, O, mean_const, y_stats(h), ptr<double>(h->resid), h->st);
    h->launches++;
    if (G == 1) {
        gpx_launch_kbuild_lower(lmat(h), nt, 0, 1, nt, ptr<double>(h->Xs), ptr<double>(h->xsq), N, d, kernel_id, variance,
                                h->rdiv, noise + jitter, h->st);
        h->launches++;
    } else {
        for (int p = next_owned_panel(h, 0); p < h->npanels; p += G) {
            gpx_launch_kbuild_lower(lmat(h), nt, panel_begin(h, p), 1, panel_end(h, p) - panel_begin(h, p), ptr<double>(h->Xs),
                                    ptr<double>(h->xsq), N, d, kernel_id, variance, h->rdiv, noise + jitter, h->st);
            h->launches++;
        }
    }
    GPX_CUDA(h, cudaMemcpyAsync(h->work1.p, h->resid.p, (size_t)Np * O * 8, cudaMemcpyDeviceToDevice, h->st));
    if (cap_nt > nt) {   // rows of the reserve are never touched by the solves: keep them zero
        GPX_CUDA(h, cudaMemsetAsync(h->work2.p, 0, (size_t)Np * O * 8, h->st));
        GPX_CUDA(h, cudaMemsetAsync(h->alpha.p, 0, (size_t)Np * O * 8, h->st));
    }
    phase_end(h, GPX_PHASE_KBUILD);

    // Cholesky (+ fused forward substitution)
    phase_begin(h, GPX_PHASE_CHOLESKY);
    rc = cholesky(h);
    phase_end(h, GPX_PHASE_CHOLESKY);

    // alpha: backward substitution (it consumes z in work2: keep a copy for gpx_append's incremental forward pass)
    h->zkeep_valid = false;
    if (!rc && G == 1 && ensure(h, h->zkeep, (size_t)Np * O * 8) == 0) {
        GPX_CUDA(h, cudaMemcpyAsync(h->zkeep.p, h->work2.p, (size_t)Np * O * 8, cudaMemcpyDeviceToDevice, h->st));
        h->zkeep_valid = true;
    }
    if (!rc) {
        phase_begin(h, GPX_PHASE_SOLVE);
        rc = backward_solve(h);
        phase_end(h, GPX_PHASE_SOLVE);
    }

    // log-determinant and quadratic form
    if (!rc) {
        phase_begin(h, GPX_PHASE_LML);
        gpx_launch_lml(ptr<double>(h->linv), ptr<double>(h->alpha), ptr<double>(h->resid), (int64_t)nt * GPX_TILE, Np, O,
                       ptr<double>(h->scal), h->st);
        h->launches++;
        phase_end(h, GPX_PHASE_LML);
    }

    double hs[2] = {0, 0};
    int hinfo = 0;
    h->h_xstats.assign(h->fit_norm_x ? 2 * d : 0, 0.0);
    h->h_ystats.assign(h->fit_norm_y ? 2 * O : 0, 0.0);
    if (h->fit_norm_x) cudaMemcpyAsync(h->h_xstats.data(), h->xstats.p, (size_t)2 * d * 8, cudaMemcpyDeviceToHost, h->st);
    if (h->fit_norm_y) cudaMemcpyAsync(h->h_ystats.data(), h->ystats.p, (size_t)2 * O * 8, cudaMemcpyDeviceToHost, h->st);
    cudaMemcpyAsync(hs, h->scal.p, 16, cudaMemcpyDeviceToHost, h->st);
    cudaMemcpyAsync(&hinfo, h->info.p, 4, cudaMemcpyDeviceToHost, h->st);
    cudaError_t e = cudaStreamSynchronize(h->st);
    cudaError_t e2 = cudaStreamSynchronize(h->st_f);
    cudaError_t e3 = cudaStreamSynchronize(h->st_c);
    if (rc) return rc;
    if (e == cudaSuccess) e = e2;
    if (e == cudaSuccess) e = e3;
    if (e != cudaSuccess) return fail(h, -2, "CUDA error during fit: %s", cudaGetErrorString(e));
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, -2, "CUDA launch error during fit: %s", cudaGetErrorString(e));
    for (int ph : {GPX_PHASE_H2D, GPX_PHASE_KBUILD, GPX_PHASE_CHOLESKY, GPX_PHASE_SOLVE, GPX_PHASE_LML}) phase_collect(h, ph);
    gemm_collect(h);
    if (h->trace) {
        h->trace_ms.assign((size_t)6 * h->npanels, 0.f);
        for (size_t i = 0; i < h->trace_ms.size(); i++) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, h->tev[0], h->tev[1 + i]) == cudaSuccess) h->trace_ms[i] = ms;
            else cudaGetLastError();
        }
    }
    if (G > 1) {
        // a failed pivot is only seen by the owner of that column: agree on the smallest non-zero info
        int* dinfo = ptr<int>(h->info);
        int big = hinfo == 0 ? 0x7fffffff : hinfo;
        GPX_CUDA(h, cudaMemcpyAsync(dinfo + 1, &big, 4, cudaMemcpyHostToDevice, h->st));
        GPX_NCCL(h, nccl().AllReduce(dinfo + 1, dinfo + 1, 1, ncclInt32, ncclMin, h->comm, h->st));
        GPX_CUDA(h, cudaMemcpyAsync(&big, dinfo + 1, 4, cudaMemcpyDeviceToHost, h->st));
        GPX_CUDA(h, cudaStreamSynchronize(h->st));
        hinfo = big == 0x7fffffff ? 0 : big;
    }
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

// one panel step of the blocked variance solve  Tt <- Tt L^-T  for nb test tiles (rows of Tt), L operand = P
// (upper_rows: Tt is known to be upper triangular in tile terms -- rows >= p1 are still zero -- so they are skipped)
}  // extern "C"

// in-panel part of the blocked solve  X <- X L^-T  for the tile rows [0, nb) of X and the tile columns [p0, p1): per
// column a TRSM against the inverted diagonal tile, then a rank-128 update of the panel's remaining columns
void trsm_inner_stepwise(gpx_t* h, TileMat X, int nb, TileMat P, int p0, int p1, int i0) {
    TileMat Li = linvmat(h);
    for (int k = p0; k < p1; k++) {
        GemmJob t{};
        t.A = X; t.B = Li; t.C = X;
        t.tri = 0; t.i0 = i0; t.i1 = nb; t.ncols = 1; t.jbase = k; t.jW = GPX_JW_CONTIG; t.jstride = 0;
        t.k0 = k; t.k1 = k + 1; t.bdiag = 1; t.overwrite = 1;
        gemm(h, t, h->st);
        if (k + 1 < p1) {
            GemmJob u{};
            u.A = X; u.B = P; u.C = X;
            u.tri = 0; u.i0 = i0; u.i1 = nb; u.ncols = p1 - (k + 1); u.jbase = k + 1; u.jW = GPX_JW_CONTIG; u.jstride = 0;
            u.k0 = k; u.k1 = k + 1; u.bdiag = 0; u.overwrite = 0;
            gemm(h, u, h->st);
        }
    }
}

// Inverse of every panel's triangular diagonal block (W x W tiles), M_p = L(p,p)^-1, kept as a block-diagonal packed-lower
// tile matrix.  With it the in-panel part of the solve is ONE row-wise GEMM per panel and batch, X(:, panel) <- X(:, panel)
// M_p^T with K up to W tiles, instead of 2 W - 1 launches at K = 1 tile (which run at ~85 % of the DMMA rate at best and
// pay a launch ramp and tail each): at BASELINE config C4 the in-panel steps are 3.5 % of the variance solve's flops but
// cost ~6 % of its time.  Built lazily by the first large predict after a fit: Y = M_p^T from the step-wise solve applied
// to an identity block, then transposed tile by tile (the tile GEMM computes A B^T only).
// Built lazily and incrementally: panels [minv_valid, upto_panel) are (re)computed; gpx_append only invalidates the panels it
// touches.
int ensure_panel_inverse(gpx_t* h, int upto_panel) {
    const int W = h->W, cap = h->cap_nt;
    if (upto_panel > h->npanels) upto_panel = h->npanels;
    // layout by CAPACITY (cap_nt), so that gpx_append can keep the blocks of untouched panels: column c holds the tiles
    // (c .. end of its panel) and the offsets do not move when nt grows inside the capacity
    if (!h->minv.p || h->minv_layout_cap != cap || h->minv_layout_w != W) {
        std::vector<int64_t> moff(cap), yoff(cap);
        int64_t tiles = 0;
        for (int c = 0; c < cap; c++) {
            int p1 = (c / W + 1) * W;
            if (p1 > cap) p1 = cap;
            moff[c] = tiles * GPX_TILE_ELEMS;
            tiles += p1 - c;
            yoff[c] = (int64_t)(c % W) * W * GPX_TILE_ELEMS;
        }
        int rc;
        if ((rc = ensure(h, h->minv, (size_t)tiles * GPX_TILE_BYTES)) || (rc = ensure(h, h->minv_coloff, (size_t)cap * 8)) ||
            (rc = ensure(h, h->ywork, (size_t)W * W * GPX_TILE_BYTES)) || (rc = ensure(h, h->ywork_coloff, (size_t)cap * 8))) return rc;
        GPX_CUDA(h, cudaMemcpyAsync(h->minv_coloff.p, moff.data(), (size_t)cap * 8, cudaMemcpyHostToDevice, h->st));
        GPX_CUDA(h, cudaMemcpyAsync(h->ywork_coloff.p, yoff.data(), (size_t)cap * 8, cudaMemcpyHostToDevice, h->st));
        GPX_CUDA(h, cudaStreamSynchronize(h->st));   // moff / yoff are host temporaries
        h->minv_layout_cap = cap; h->minv_layout_w = W;
        h->minv_valid = 0;
    }
    TileMat Y{ptr<double>(h->ywork), ptr<int64_t>(h->ywork_coloff), 0};
    TileMat M{ptr<double>(h->minv), ptr<int64_t>(h->minv_coloff), 1};
    for (int p = h->minv_valid; p < upto_panel; p++) {
        const int p0 = panel_begin(h, p), p1 = panel_end(h, p);
        gpx_launch_identity_block(Y, p1 - p0, p0, h->st);
        trsm_inner_stepwise(h, Y, p1 - p0, lmat(h), p0, p1);
        gpx_launch_transpose_block(Y, M, p0, p1 - p0, h->st);
        h->launches += 2;
    }
    if (upto_panel > h->minv_valid) h->minv_valid = upto_panel;
    return 0;
}

// one panel step of the blocked variance solve  Tt <- Tt L^-T  for nb test tiles (rows of Tt), L operand = P
// (upper_rows: Tt is known to be upper triangular in tile terms -- rows >= p1 are still zero -- so they are skipped)
void var_trsm_panel(gpx_t* h, TileMat Tt, int nb_in, TileMat P, int p, bool upper_rows, int row0) {
    const int nt = h->nt, p0 = panel_begin(h, p), p1 = panel_end(h, p);
    const int nb = (upper_rows && p1 < nb_in) ? p1 : nb_in;
    if (row0 >= nb) return;
    if (h->minv_active && !upper_rows) {
        GemmJob t{};   // Tt(:, panel) <- Tt(:, panel) M_p^T, in place, one tile row per CTA (see GemmJob::rowwise)
        t.A = Tt; t.B = TileMat{ptr<double>(h->minv), ptr<int64_t>(h->minv_coloff), 1}; t.C = Tt;
        t.tri = 0; t.i0 = row0; t.i1 = nb; t.ncols = p1 - p0; t.jbase = p0; t.jW = GPX_JW_CONTIG; t.jstride = 0;
        t.k0 = p0; t.k1 = p1; t.bdiag = 0; t.overwrite = 1; t.rowwise = 1;
        gemm(h, t, h->st);
    } else {
        trsm_inner_stepwise(h, Tt, nb, P, p0, p1, row0);
    }
    if (p1 < nt) {
        GemmJob u{};
        u.A = Tt; u.B = P; u.C = Tt;
        u.tri = 0; u.i0 = row0; u.i1 = nb; u.ncols = nt - p1; u.jbase = p1; u.jW = GPX_JW_CONTIG; u.jstride = 0;
        u.k0 = p0; u.k1 = p1; u.bdiag = 0; u.overwrite = 0;
        gemm(h, u, h->st);
    }
}

// full blocked solve for one batch; world > 1 re-broadcasts the panels of L (double buffered, stream C)
int var_trsm(gpx_t* h, TileMat Tt, int nb) {
    const int np = h->npanels, G = h->world;
    if (G == 1) {
        if (nb > 0) for (int p = 0; p < np; p++) var_trsm_panel(h, Tt, nb, lmat(h), p);
        return 0;
    }
    int rc = ensure_panel_events(h, np);
    if (rc) return rc;
    cudaEvent_t ev_start;
    GPX_CUDA(h, cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    GPX_CUDA(h, cudaEventRecord(ev_start, h->st));
    GPX_CUDA(h, cudaStreamWaitEvent(h->st_c, ev_start, 0));  // previous users of the panel buffers are done
    for (int p = 0; p < np; p++) {
        if (p >= h->cache_panels) {   // panels below cache_panels are resident on every rank since the fit
            const int q = owns_panel(h, p) ? -1 : prev_same_buffer_panel(h, p);
            if (q >= 0) GPX_CUDA(h, cudaStreamWaitEvent(h->st_c, pevent(h, q, 2), 0));
            if ((rc = broadcast_panel(h, p, false))) return rc;
            GPX_CUDA(h, cudaEventRecord(pevent(h, p, 0), h->st_c));
            GPX_CUDA(h, cudaStreamWaitEvent(h->st, pevent(h, p, 0), 0));
        }
        if (nb > 0) var_trsm_panel(h, Tt, nb, panel_mat(h, p), p);
        GPX_CUDA(h, cudaEventRecord(pevent(h, p, 2), h->st));
    }
    cudaEventDestroy(ev_start);
    return 0;
}

extern "C" {

int gpx_predict(gpx_t* h, const double* Xs, int64_t Ns, double* mean, double* var, int var_cols, int include_noise,
                int mem) {
    if (!h) return -1;
    if (!h->fitted) return fail(h, -1, "gpx_predict called before a successful gpx_fit");
    if (!Xs || !mean) return fail(h, -1, "null argument");
    if (Ns < 1) return fail(h, -1, "bad Ns");
    if (var && var_cols != 1 && var_cols != h->O) return fail(h, -1, "var_cols must be 1 or O (%d)", h->O);
    if (var && h->fit_norm_y && h->O > 1 && var_cols == 1)
        return fail(h, -1, "with the output normaliser every output has its own variance scale: pass var_cols = O");
    GPX_CUDA(h, cudaSetDevice(h->device));
    { int rcw = wait_for_caller(h); if (rcw) return rcw; }
    if (h->world == 1 && Ns <= GPX_TILE) return gpx_predict_small(h, Xs, Ns, mean, var, var_cols, include_noise, mem);
    const int nt = h->nt, d = h->d, O = h->O, G = h->world;
    const int64_t Np = h->Np;
    for (int ph : {GPX_PHASE_PREDICT_KSTAR, GPX_PHASE_PREDICT_TRSM, GPX_PHASE_PREDICT_REDUCE, GPX_PHASE_H2D, GPX_PHASE_D2H}) h->phase_ms[ph] = 0.f;

    // test tiles are sharded contiguously over the ranks
    const int total_tiles = (int)((Ns + GPX_TILE - 1) / GPX_TILE);
    const int my_t0 = (int)((int64_t)total_tiles * h->rank / G), my_t1 = (int)((int64_t)total_tiles * (h->rank + 1) / G);
    int max_shard = 0;
    for (int r = 0; r < G; r++) {
        int s = (int)((int64_t)total_tiles * (r + 1) / G) - (int)((int64_t)total_tiles * r / G);
        if (s > max_shard) max_shard = s;
    }
    // batch size: a multiple of the SM count when possible, bounded by free memory when the variance is wanted.
    // All ranks must agree on the number of batches (the panel broadcasts are collective), so the memory bound
    // is the minimum over ranks.
    // default 3 test tiles per SM: larger batches mean fewer, longer GEMM launches per panel (148 -> 400 test tiles per batch:
    // variance solve 35.59 -> 35.71 TFLOP/s at C4) at the price of a bigger workspace, which the memory bound below caps
    int cap = h->predict_batch_tiles > 0 ? h->predict_batch_tiles : 3 * h->num_sms;
    if (var) {
        size_t fr = 0, tot = 0;
        GPX_CUDA(h, cudaMemGetInfo(&fr, &tot));
        fr += h->Tt.bytes;  // reusable
        if (G > 1) {
            double v = (double)fr;
            double* ds = ptr<double>(h->scal) + 4;
            GPX_CUDA(h, cudaMemcpyAsync(ds, &v, 8, cudaMemcpyHostToDevice, h->st));
            GPX_NCCL(h, nccl().AllReduce(ds, ds, 1, ncclDouble, ncclMin, h->comm, h->st));
            GPX_CUDA(h, cudaMemcpyAsync(&v, ds, 8, cudaMemcpyDeviceToHost, h->st));
            GPX_CUDA(h, cudaStreamSynchronize(h->st));
            fr = (size_t)v;
        }
        size_t per_tile_row = (size_t)nt * GPX_TILE_BYTES;
        size_t can = (size_t)(0.8 * (double)fr) / per_tile_row;
        if (can < 1) return fail(h, -3, "not enough device memory for the variance workspace");
        if ((size_t)cap > can) cap = (int)can;
    }
    if (cap > max_shard) cap = max_shard;
    if (cap < 1) cap = 1;
    const int nbatch = (max_shard + cap - 1) / cap;
    const int per = (max_shard + nbatch - 1) / nbatch;

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
        if (var) { if ((rc = ensure(h, h->dvar, (size_t)Ns * var_cols * 8))) return rc; dvar = ptr<double>(h->dvar); }
    }
    if (G > 1) {  // every rank fills its rows; one all-reduce assembles the result
        GPX_CUDA(h, cudaMemsetAsync(dmean, 0, (size_t)Ns * O * 8, h->st));
        if (var) GPX_CUDA(h, cudaMemsetAsync(dvar, 0, (size_t)Ns * var_cols * 8, h->st));
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
    // in-panel solve through the inverted diagonal blocks: pays when there are enough test tiles to amortise building them
    h->minv_active = 0;
    if (var && G == 1 && h->W >= 2 && (h->panel_inverse == 1 || (h->panel_inverse < 0 && total_tiles >= 64))) {
        if ((rc = ensure_panel_inverse(h, h->npanels))) return rc;
        h->minv_active = 1;
    }

    for (int b = 0; b < nbatch; b++) {
        const int t0 = my_t0 + b * per;
        int nb = my_t1 - t0;
        if (nb > per) nb = per;
        if (nb < 0) nb = 0;
        const int64_t row0 = (int64_t)t0 * GPX_TILE;
        const int64_t valid = nb > 0 ? ((Ns - row0 < (int64_t)nb * GPX_TILE) ? (Ns - row0) : (int64_t)nb * GPX_TILE) : 0;
        phase_begin(h, GPX_PHASE_PREDICT_KSTAR);
        if (nb > 0) {
            prep_inputs(h, dXs + row0 * d, valid, (int64_t)nb * GPX_TILE, ptr<double>(h->Xt), ptr<double>(h->xtsq), h->st);
            gpx_launch_kcross(Tt, nb, per, nt, ptr<double>(h->Xt), ptr<double>(h->xtsq), valid, ptr<double>(h->Xs),
                              ptr<double>(h->xsq), h->N, d, h->kernel_id, h->variance, h->rdiv, ptr<double>(h->alpha), O, Np,
                              ptr<double>(h->mean_partial), h->st);
            gpx_launch_mean_reduce(ptr<double>(h->mean_partial), per, nt, O, h->mean_const, y_stats(h), dmean, row0,
                                   row0 + valid, h->st);
            h->launches += 2;
        }
        phase_end(h, GPX_PHASE_PREDICT_KSTAR);
        if (var) {
            phase_begin(h, GPX_PHASE_PREDICT_TRSM);
            if ((rc = var_trsm(h, Tt, nb))) return rc;
            phase_end(h, GPX_PHASE_PREDICT_TRSM);
            phase_begin(h, GPX_PHASE_PREDICT_REDUCE);
            if (nb > 0) {
                if ((rc = var_reduce(h, Tt, nb, prior_var, var_cols, dvar, row0, row0 + valid))) return rc;
            }
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
    h->minv_active = 0;
    if (G > 1) {
        GPX_NCCL(h, nccl().AllReduce(dmean, dmean, (size_t)Ns * O, ncclDouble, ncclSum, h->comm, h->st));
        if (var) GPX_NCCL(h, nccl().AllReduce(dvar, dvar, (size_t)Ns * var_cols, ncclDouble, ncclSum, h->comm, h->st));
    }
    if (mem == GPX_MEM_HOST) {
        phase_begin(h, GPX_PHASE_D2H);
        GPX_CUDA(h, cudaMemcpyAsync(mean, dmean, (size_t)Ns * O * 8, cudaMemcpyDeviceToHost, h->st));
        if (var) GPX_CUDA(h, cudaMemcpyAsync(var, dvar, (size_t)Ns * var_cols * 8, cudaMemcpyDeviceToHost, h->st));
        phase_end(h, GPX_PHASE_D2H);
    }
    cudaError_t e = cudaStreamSynchronize(h->st);
    cudaError_t e2 = cudaStreamSynchronize(h->st_c);
    if (e == cudaSuccess) e = e2;
    if (e != cudaSuccess) return fail(h, -2, "CUDA error during predict: %s", cudaGetErrorString(e));
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, -2, "CUDA launch error during predict: %s", cudaGetErrorString(e));
    phase_collect(h, GPX_PHASE_H2D);
    if (mem == GPX_MEM_HOST) phase_collect(h, GPX_PHASE_D2H);
    gemm_collect(h);
    return 0;
}

int gpx_predict_cov(gpx_t* h, const double* Xs, int64_t Ns, double* mean, double* cov, int include_noise, int mem) {
    if (!h) return -1;
    if (!h->fitted) return fail(h, -1, "gpx_predict_cov called before a successful gpx_fit");
    if (!Xs || !mean || !cov) return fail(h, -1, "null argument");
    if (Ns < 1) return fail(h, -1, "bad Ns");
    GPX_CUDA(h, cudaSetDevice(h->device));
    const int nt = h->nt, d = h->d, O = h->O;
    const int64_t Np = h->Np;
    const int nbt = (int)((Ns + GPX_TILE - 1) / GPX_TILE);
    if (Ns > 65535) return fail(h, -1, "gpx_predict_cov: at most 65535 test points (the dense covariance would be %.0f GB); use full_cov=False",
                                (double)Ns * Ns * 8 / 1e9);
    // the whole transposed cross-kernel (nbt x nt tiles) and the packed covariance must be resident at once
    {
        size_t fr = 0, tot = 0;
        GPX_CUDA(h, cudaMemGetInfo(&fr, &tot));
        fr += h->Tt.bytes;
        double need = (double)nbt * nt * GPX_TILE_BYTES + (double)nbt * (nbt + 1) / 2 * GPX_TILE_BYTES + 3.0 * (double)Ns * Ns * 8;
        if (need > 0.8 * (double)fr)
            return fail(h, -3, "full predictive covariance for %lld test points needs %.1f GB of workspace; use full_cov=False",
                        (long long)Ns, need / 1e9);
    }
    int rc;
    const double* dXs;
    if ((rc = to_device(h, Xs, (size_t)Ns * d, mem, h->stage, &dXs))) return rc;
    DevBuf &dm = h->c_dm, &dc = h->c_dc, &covt = h->c_covt, &covoff = h->c_covoff;   // kept between calls
    auto cleanup = [&]() {};
    double* dmean = mean;
    double* dcov = cov;
    if (mem == GPX_MEM_HOST) {
        if ((rc = ensure(h, dm, (size_t)Ns * O * 8)) || (rc = ensure(h, dc, (size_t)Ns * Ns * 8))) { cleanup(); return rc; }
        dmean = ptr<double>(dm);
        dcov = ptr<double>(dc);
    }
    std::vector<int64_t> tc(nt), cc(nbt);
    for (int c = 0; c < nt; c++) tc[c] = (int64_t)c * nbt * GPX_TILE_ELEMS;
    int64_t o = 0;
    for (int c = 0; c < nbt; c++) { cc[c] = o; o += (int64_t)(nbt - c) * GPX_TILE_ELEMS; }
    if ((rc = ensure(h, h->Xt, (size_t)nbt * GPX_TILE * d * 8)) || (rc = ensure(h, h->xtsq, (size_t)nbt * GPX_TILE * 8)) ||
        (rc = ensure(h, h->mean_partial, (size_t)nt * nbt * GPX_TILE * O * 8)) ||
        (rc = ensure(h, h->Tt, (size_t)nbt * nt * GPX_TILE_BYTES)) || (rc = ensure(h, h->tt_coloff, (size_t)nt * 8)) ||
        (rc = ensure(h, covt, (size_t)o * 8)) || (rc = ensure(h, covoff, (size_t)nbt * 8))) { cleanup(); return rc; }
    cudaMemcpyAsync(h->tt_coloff.p, tc.data(), (size_t)nt * 8, cudaMemcpyHostToDevice, h->st);
    cudaMemcpyAsync(covoff.p, cc.data(), (size_t)nbt * 8, cudaMemcpyHostToDevice, h->st);
    cudaStreamSynchronize(h->st);
    TileMat Tt{ptr<double>(h->Tt), ptr<int64_t>(h->tt_coloff), 0};
    TileMat Cv{ptr<double>(covt), ptr<int64_t>(covoff), 1};
    // K(X*, X) tiles + mean;  every rank handles ALL test points here (small N*), panels of L are re-broadcast
    prep_inputs(h, dXs, Ns, (int64_t)nbt * GPX_TILE, ptr<double>(h->Xt), ptr<double>(h->xtsq), h->st);
    gpx_launch_kcross(Tt, nbt, nbt, nt, ptr<double>(h->Xt), ptr<double>(h->xtsq), Ns, ptr<double>(h->Xs), ptr<double>(h->xsq),
                      h->N, d, h->kernel_id, h->variance, h->rdiv, ptr<double>(h->alpha), O, Np, ptr<double>(h->mean_partial), h->st);
    gpx_launch_mean_reduce(ptr<double>(h->mean_partial), nbt, nt, O, h->mean_const, y_stats(h), dmean, 0, Ns, h->st);
    // K(X*, X*) + noise I (exact sigma_f^2 diagonal, like kern.K(Xnew)) into packed lower tiles
    gpx_launch_kbuild_lower(Cv, nbt, 0, 1, nbt, ptr<double>(h->Xt), ptr<double>(h->xtsq), Ns, d, h->kernel_id, h->variance,
                            h->rdiv, include_noise ? h->noise : 0.0, h->st);
    h->launches += 3;
    if ((rc = var_trsm(h, Tt, nbt))) { cleanup(); return rc; }
    // cov -= Tt Tt^T  (lower triangle, contraction over all training tiles)
    GemmJob g{};
    g.A = Tt; g.B = Tt; g.C = Cv;
    g.tri = 1; g.i0 = 0; g.i1 = nbt; g.ncols = nbt; g.jbase = 0; g.jW = GPX_JW_CONTIG; g.jstride = 0;
    g.k0 = 0; g.k1 = nt; g.bdiag = 0; g.overwrite = 0;
    gemm(h, g, h->st);
    gpx_launch_detile(Cv, 2, Ns, Ns, dcov, h->st);
    h->launches++;
    if (mem == GPX_MEM_HOST) {
        cudaMemcpyAsync(mean, dmean, (size_t)Ns * O * 8, cudaMemcpyDeviceToHost, h->st);
        cudaMemcpyAsync(cov, dcov, (size_t)Ns * Ns * 8, cudaMemcpyDeviceToHost, h->st);
    }
    cudaError_t e = cudaStreamSynchronize(h->st);
    cudaError_t e2 = cudaStreamSynchronize(h->st_c);
    if (e == cudaSuccess) e = e2;
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    gemm_collect(h);
    if (e != cudaSuccess) return fail(h, -2, "CUDA error in gpx_predict_cov: %s", cudaGetErrorString(e));
    return 0;
}

int gpx_lml_grad(gpx_t* h, double* grad, int capacity) {
    if (!h) return -1;
    if (!h->fitted) return fail(h, -1, "gpx_lml_grad called before a successful gpx_fit");
    const int G = h->world;
    // Sharded handle: every rank needs the whole factor (the panel cache of the fit holds it when memory allows -- the same
    // replica the predict sweep uses); the tile rows of Y = L^-T are then split over the ranks (each row is an independent
    // solve against the replica) and exchanged, the row groups of Ky^-1 = Y Y^T and of the gradient tiles are dealt out
    // cyclically and the 2 + n_ls sums meet in one all-reduce: 2 N^3 / (3 G) flop per rank + N^2 / 2 doubles exchanged.
    if (G > 1 && h->cache_panels < h->npanels)
        return fail(h, -6, "gpx_lml_grad on a sharded handle needs the whole factor resident on every rank, but the panel cache "
                    "holds %d of %d panels (N too large for a replica, or cache_mb too small): optimise at a smaller N or on one GPU",
                    h->cache_panels, h->npanels);
    const int nt = h->nt, n_ls = h->n_ls, nparam = n_ls + 3, stride = 2 + n_ls;
    if (!grad || capacity < nparam) return fail(h, -1, "gradient buffer too small: need %d", nparam);
    if (h->d > gpx_grad_max_d()) return fail(h, -1, "gpx_lml_grad supports input dimension <= %d", gpx_grad_max_d());
    GPX_CUDA(h, cudaSetDevice(h->device));
    // workspaces: Y = L^-T as an nt x nt tile matrix (in Tt), Ky^-1 packed lower, per-tile partial sums
    if (!h->finv_valid) { release(h, h->finv); release(h, h->finv_partial); h->finv_layout_cap = -1; }   // a stale latency-path inverse
    int rc = 0;
    DevBuf &kinv = h->g_kinv, &kinv_off = h->g_kinv_off, &partial = h->g_partial, &gout = h->g_out;   // kept between calls
    auto cleanup = [&]() {};
    // stage timing (GPX_GRAD_TRACE=1 prints it): development aid
    static const bool trace = getenv("GPX_GRAD_TRACE") != nullptr;
    cudaEvent_t tev[6] = {};
    auto mark = [&](int i) { if (trace) { if (!tev[i]) cudaEventCreate(&tev[i]); cudaEventRecord(tev[i], h->st); } };
    std::vector<int64_t> tc(nt), kc(nt);
    int64_t o = 0;
    for (int c = 0; c < nt; c++) { tc[c] = (int64_t)c * nt * GPX_TILE_ELEMS; kc[c] = o; o += (int64_t)(nt - c) * GPX_TILE_ELEMS; }
    {
        size_t fr = 0, tot = 0;
        GPX_CUDA(h, cudaMemGetInfo(&fr, &tot));
        fr += h->Tt.bytes + h->g_kinv.bytes;
        double need = ((double)nt * nt + (double)nt * (nt + 1) / 2) * GPX_TILE_BYTES;
        if (need > 0.9 * (double)fr && h->finv.p) {   // the latency path's inverse is only a cache: give it back
            release(h, h->finv); release(h, h->finv_partial);
            h->finv_layout_cap = -1;
            if (h->finv_valid) { h->finv_valid = false; h->gen++; }
            GPX_CUDA(h, cudaMemGetInfo(&fr, &tot));
            fr += h->Tt.bytes + h->g_kinv.bytes;
        }
        if (need > 0.9 * (double)fr) rc = fail(h, -3, "gpx_lml_grad needs %.1f GB of workspace (N too large for one GPU)", need / 1e9);
    }
    if (!rc) {
        (rc = ensure(h, h->Tt, (size_t)nt * nt * GPX_TILE_BYTES)) || (rc = ensure(h, h->tt_coloff, (size_t)nt * 8)) ||
        (rc = ensure(h, kinv, (size_t)o * 8)) || (rc = ensure(h, kinv_off, (size_t)nt * 8)) ||
        (rc = ensure(h, partial, (size_t)nt * nt * stride * 8)) || (rc = ensure(h, gout, (size_t)(stride + 1) * 8));
    }
    if (G > 1) {   // every rank must take the same decision: a rank that bailed out alone would leave the others in the all-reduce
        double ok = rc ? 0.0 : 1.0;
        double* ds = ptr<double>(h->scal) + 5;
        GPX_CUDA(h, cudaMemcpyAsync(ds, &ok, 8, cudaMemcpyHostToDevice, h->st));
        GPX_NCCL(h, nccl().AllReduce(ds, ds, 1, ncclDouble, ncclMin, h->comm, h->st));
        GPX_CUDA(h, cudaMemcpyAsync(&ok, ds, 8, cudaMemcpyDeviceToHost, h->st));
        GPX_CUDA(h, cudaStreamSynchronize(h->st));
        if (!rc && ok == 0.0) rc = fail(h, -3, "gpx_lml_grad: another rank ran out of workspace memory");
    }
    if (rc) { cleanup(); return rc; }
    cudaMemcpyAsync(h->tt_coloff.p, tc.data(), (size_t)nt * 8, cudaMemcpyHostToDevice, h->st);
    cudaMemcpyAsync(kinv_off.p, kc.data(), (size_t)nt * 8, cudaMemcpyHostToDevice, h->st);
    cudaStreamSynchronize(h->st);
    TileMat Y{ptr<double>(h->Tt), ptr<int64_t>(h->tt_coloff), 0};
    TileMat Ki{ptr<double>(kinv), ptr<int64_t>(kinv_off), 1};
    // 1. Y = L^-T : solve Y L^T = I with the blocked DMMA solve of the variance path (upper triangular rows only)
    mark(0);
    gpx_launch_identity_tiles(Y, nt, h->st);
    h->launches++;
    mark(1);
    // Sharded: the tile rows of Y are independent (row i is e_i^T L^-T), so each rank solves one contiguous block of them --
    // boundaries balance the work, ~ (nt - i)^2 per row -- against its replica of the factor, and the blocks are exchanged
    // with in-place broadcasts (column by column: a block's rows of one tile column are contiguous).
    std::vector<int> bnd(G + 1, nt);
    bnd[0] = 0;
    if (G > 1) {
        double total = 0.0, cum = 0.0;
        for (int i = 0; i < nt; i++) total += (double)(nt - i) * (nt - i) + 4.0 * nt;
        int r = 1;
        for (int i = 0; i < nt && r < G; i++) {
            cum += (double)(nt - i) * (nt - i) + 4.0 * nt;
            while (r < G && cum >= total * r / G) bnd[r++] = i + 1;
        }
    }
    const int my0 = bnd[h->rank], my1 = bnd[h->rank + 1];
    for (int p = 0; p < h->npanels; p++) {
        if (G == 1) var_trsm_panel(h, Y, nt, lmat(h), p, true);
        else if (my0 < my1) var_trsm_panel(h, Y, my1, panel_mat(h, p), p, true, my0);
    }
    if (G > 1) {
        for (int r = 0; r < G; r++) {
            const int a = bnd[r], b = bnd[r + 1];
            if (a >= b) continue;
            GPX_NCCL(h, nccl().GroupStart());
            for (int c = (a / h->W) * h->W; c < nt; c++) {
                double* blk = Y.base + tc[c] + (int64_t)a * GPX_TILE_ELEMS;
                nccl().Broadcast(blk, blk, (size_t)(b - a) * GPX_TILE_ELEMS, ncclDouble, r, h->comm, h->st);
            }
            GPX_NCCL(h, nccl().GroupEnd());
        }
    }
    mark(2);
    // 2. Ky^-1 = Y Y^T (lower tiles; Y(I,K) = 0 for K < I, so the contraction starts at the row group's first tile)
    const int RG = G > 1 ? 4 : 8;
    for (int I0 = 0; I0 < nt; I0 += RG) {
        if (G > 1 && (I0 / RG) % G != h->rank) continue;
        const int I1 = I0 + RG < nt ? I0 + RG : nt;
        GemmJob g{};
        g.A = Y; g.B = Y; g.C = Ki;
        g.tri = 1; g.i0 = I0; g.i1 = I1; g.ncols = I1; g.jbase = 0; g.jW = GPX_JW_CONTIG; g.jstride = 0;
        g.k0 = I0; g.k1 = nt; g.bdiag = 0; g.overwrite = 1;
        gemm(h, g, h->st);
    }
    mark(3);
    // 3. dLML/dtheta = sum W .* dK/dtheta
    gpx_launch_lml_grad(Ki, nt, ptr<double>(h->Xs), ptr<double>(h->xsq), h->N, h->d, h->kernel_id, h->variance, h->rdiv,
                        ptr<double>(h->ls), n_ls, ptr<double>(h->alpha), h->O, h->Np, ptr<double>(partial),
                        ptr<double>(gout), h->st, RG, G, h->rank);
    h->launches += 3;
    if (G > 1) GPX_NCCL(h, nccl().AllReduce(gout.p, gout.p, (size_t)stride, ncclDouble, ncclSum, h->comm, h->st));
    mark(4);
    if (trace) {
        cudaStreamSynchronize(h->st);
        float ms[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; i++) cudaEventElapsedTime(&ms[i], tev[i], tev[i + 1]);
        fprintf(stderr, "[gpx_lml_grad] N=%lld identity %.2f ms, Y = L^-T %.2f ms, Ky^-1 = Y Y^T %.2f ms, gradient tiles %.2f ms\n",
                (long long)h->N, ms[0], ms[1], ms[2], ms[3]);
        for (auto& e : tev) if (e) cudaEventDestroy(e);
    }
    std::vector<double> hg(stride + 1);
    cudaError_t e = cudaMemcpyAsync(hg.data(), gout.p, (size_t)(stride + 1) * 8, cudaMemcpyDeviceToHost, h->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cleanup();
    gemm_collect(h);
    if (e != cudaSuccess) return fail(h, -2, "CUDA error in gpx_lml_grad: %s", cudaGetErrorString(e));
    // order: [variance, lengthscale..., noise, mean_const].  The tile kernel weights diagonal elements with 1/2 and
    // strictly-lower ones with 1 (= 2 * 1/2, symmetry), i.e. it already sums W = 0.5 (alpha alpha^T - O Kinv).
    grad[0] = hg[0];
    for (int q = 0; q < n_ls; q++) grad[1 + q] = hg[2 + q];
    grad[1 + n_ls] = hg[1];
    grad[2 + n_ls] = hg[stride];
    return nparam;
}

#ifdef GPX_ENABLE_DEBUG_API
// Development aid (tools/gemm_bench.py): time the tile GEMM in isolation on a synthetic rectangular tile matrix.
// C (rows x cols tiles) (-)= A (rows x ktiles) * B (cols x ktiles)^T, `reps` launches; returns the mean ms per launch.
double gpx_debug_gemm(gpx_t* h, int rows, int cols, int ktiles, int overwrite, int max_chunk, int reps) {
    if (!h || rows < 1 || cols < 1 || ktiles < 1) return -1.0;
    cudaSetDevice(h->device);
    const int ncolsA = ktiles, ncolsC = cols;
    DevBuf A, B, C, offA, offB, offC;
    size_t tb = GPX_TILE_BYTES;
    if (ensure(h, A, (size_t)rows * ktiles * tb) || ensure(h, B, (size_t)cols * ktiles * tb) || ensure(h, C, (size_t)rows * cols * tb) ||
        ensure(h, offA, (size_t)ncolsA * 8) || ensure(h, offB, (size_t)ktiles * 8) || ensure(h, offC, (size_t)ncolsC * 8)) return -2.0;
    std::vector<int64_t> oa(ktiles), ob(ktiles), oc(cols);
    for (int k = 0; k < ktiles; k++) { oa[k] = (int64_t)k * rows * GPX_TILE_ELEMS; ob[k] = (int64_t)k * cols * GPX_TILE_ELEMS; }
    for (int c = 0; c < cols; c++) oc[c] = (int64_t)c * rows * GPX_TILE_ELEMS;
    cudaMemcpy(offA.p, oa.data(), (size_t)ktiles * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(offB.p, ob.data(), (size_t)ktiles * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(offC.p, oc.data(), (size_t)cols * 8, cudaMemcpyHostToDevice);
    cudaMemset(A.p, 0, A.bytes); cudaMemset(B.p, 0, B.bytes); cudaMemset(C.p, 0, C.bytes);
    GemmJob g{};
    g.A = TileMat{ptr<double>(A), ptr<int64_t>(offA), 0};
    g.B = TileMat{ptr<double>(B), ptr<int64_t>(offB), 0};
    g.C = TileMat{ptr<double>(C), ptr<int64_t>(offC), 0};
    g.tri = 0; g.i0 = 0; g.i1 = rows; g.ncols = cols; g.jbase = 0; g.jW = GPX_JW_CONTIG; g.jstride = 0;
    g.k0 = 0; g.k1 = ktiles; g.bdiag = 0; g.overwrite = overwrite; g.chunk = max_chunk; g.strip = h->gemm_strip; g.interleave = h->gemm_interleave ? 0 : -1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    gpx_launch_gemm(g, h->num_sms, h->st, nullptr);
    cudaEventRecord(e0, h->st);
    for (int r = 0; r < reps; r++) gpx_launch_gemm(g, h->num_sms, h->st, nullptr);
    cudaEventRecord(e1, h->st);
    cudaStreamSynchronize(h->st);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    for (DevBuf* b : {&A, &B, &C, &offA, &offB, &offC}) release(h, *b);
    return cudaGetLastError() == cudaSuccess ? ms / reps : -3.0;
}

// Development aid: mean device time (us) of the diagonal-tile POTRF+inverse kernel on a well-conditioned SPD tile.
double gpx_debug_potrf(gpx_t* h, int reps) {
    if (!h) return -1.0;
    cudaSetDevice(h->device);
    DevBuf src, work, inv, info;
    if (ensure(h, src, GPX_TILE_BYTES) || ensure(h, work, GPX_TILE_BYTES) || ensure(h, inv, GPX_TILE_BYTES) || ensure(h, info, 64)) return -2.0;
    std::vector<double> t(GPX_TILE_ELEMS);
    for (int r = 0; r < GPX_TILE; r++)
        for (int c = 0; c < GPX_TILE; c++) t[gpx_tile_idx(r, c)] = (r == c ? 4.0 : 0.0) + exp(-0.01 * (r - c) * (r - c));
    cudaMemcpy(src.p, t.data(), GPX_TILE_BYTES, cudaMemcpyHostToDevice);
    cudaMemset(info.p, 0, 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float total = 0.f;
    for (int r = 0; r < reps + 1; r++) {
        cudaMemcpyAsync(work.p, src.p, GPX_TILE_BYTES, cudaMemcpyDeviceToDevice, h->st);
        cudaEventRecord(e0, h->st);
        gpx_launch_potrf_tile(ptr<double>(work), ptr<double>(inv), ptr<int>(info), 0, h->st);
        cudaEventRecord(e1, h->st);
        cudaStreamSynchronize(h->st);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0) total += ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    for (DevBuf* b : {&src, &work, &inv, &info}) release(h, *b);
    return 1e3 * total / reps;
}

#endif  // GPX_ENABLE_DEBUG_API

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
        if (h->world > 1) return fail(h, -1, "gpx_export: the dense factor is only available on a single-GPU handle");
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
                      double variance, const double* lengthscale, int n_ls, int ard, double diag_add, double* K_out) {
    if (!h) return -1;
    if (!X || !lengthscale || !K_out) return fail(h, -1, "null argument");
    if (N < 1 || d < 1 || d > gpx_kbuild_max_d() || (ard ? n_ls != d : n_ls != 1)) return fail(h, -1, "bad shape");
    const double rdiv = ard ? 1.0 : lengthscale[0];
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
    gpx_launch_prep_x(ptr<double>(dX), N, (int64_t)ntr * GPX_TILE, d, ptr<double>(dls), ard, nullptr, nullptr, ptr<double>(dXs), ptr<double>(dsq), h->st);
    if (sym) {
        TileMat Lm{ptr<double>(tiles), ptr<int64_t>(off), 1};
        gpx_launch_kbuild_lower(Lm, ntr, 0, 1, ntr, ptr<double>(dXs), ptr<double>(dsq), N, d, kernel_id, variance, rdiv, diag_add, h->st);
        gpx_launch_detile(Lm, 1, N, N, ptr<double>(dense), h->st);
    } else {
        if ((rc = ensure(h, dX2, (size_t)M * d * 8)) || (rc = ensure(h, dX2s, (size_t)ntc * GPX_TILE * d * 8)) ||
            (rc = ensure(h, dsq2, (size_t)ntc * GPX_TILE * 8))) { cleanup(); return rc; }
        cudaMemcpyAsync(dX2.p, X2, (size_t)M * d * 8, cudaMemcpyHostToDevice, h->st);
        gpx_launch_prep_x(ptr<double>(dX2), M, (int64_t)ntc * GPX_TILE, d, ptr<double>(dls), ard, nullptr, nullptr, ptr<double>(dX2s), ptr<double>(dsq2), h->st);
        TileMat Tm{ptr<double>(tiles), ptr<int64_t>(off), 0};
        // rows = X (as "test"), cols = X2 (as "train")
        gpx_launch_kcross(Tm, ntr, ntr, ntc, ptr<double>(dXs), ptr<double>(dsq), N, ptr<double>(dX2s), ptr<double>(dsq2), M, d,
                          kernel_id, variance, rdiv, nullptr, 0, 0, nullptr, h->st);
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
