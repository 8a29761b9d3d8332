#!/usr/bin/env python
"""Benchmark of the exact-GP hot path (fit + predict) -- the driver's contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C1..C5] [--N n --Ns m]

A "step" = one full pass of the hot path over one synthetic data set: K-build + Cholesky + alpha + LML (fit) and
predictive mean + variance for every test point (predict).  Workloads are the BASELINE.json configs
(thesis_code_b200/synthetic.py).  N=1 runs C4 (N=60000, d=1, N*=100000), the largest configuration that BASELINE
lists for a single B200; N>1 runs C5 (N=200000, d=8, N*=2500) block-column-cyclically sharded over the ranks.

value        = algorithmic fp64 TFLOP/s of fit+predict (N^3/3 + N^2 N* flop / step time), inputs resident in HBM.
e2e          = the same through the public GPModel API with HOST numpy buffers (H2D / D2H copies inside the timing).
roofline     = the dominant kernel (DMMA tile GEMM, csrc/gemm_dmma.cu): algorithmic flops of the two tensor phases
               / summed CUDA-event time of its launches, against the measured FP64 DMMA peak of this pool's B200.
cpu_baseline = the numpy/scipy oracle (a port of the reference's GPy path) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) would land there too
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

FP64_DMMA_PEAK_TFLOPS = 36.96   # measured on this pool's B200 with tools/fp64_peak.cu (profiles/r01_fp64_peak.txt)
METRIC = "exact-GP fit+predict FP64 throughput (algorithmic N^3/3 + N^2*N* flop per fit+predict)"
# DRAM traffic of the dominant kernel from the latest `ncu --set full` capture (dram__bytes_read.sum + dram__bytes_write.sum of
# one launch; profiles/): refreshed whenever the kernel changes.
GEMM_TRAFFIC = {"dram_bytes_per_launch": 3.62e9,
                "note": "ncu --set full capture of one trailing-update launch of the shipped kernel "
                        "(profiles/r02_gemm_dmma_ncu_summary.txt): 3.62 GB DRAM for 2.9 GB algorithmic operand bytes = 1.25x "
                        "(6% of HBM peak: the kernel is DMMA-bound; DMMA pipe 96.5% of active cycles, L2 hit rate 86%)"}


def algorithmic_flops(N, Ns):
    return float(N) ** 3 / 3.0 + float(N) ** 2 * float(Ns)


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


def kernel_args(cfg):
    from thesis_code_b200 import _gpx
    kw = cfg["model"]["kwargs"]
    kk = kw["kernel"]["kwargs"]
    ls = np.atleast_1d(np.asarray(kk.get("lengthscale", 1.0), dtype=np.float64))
    if kk.get("ARD", False) and ls.size == 1:
        ls = np.full(cfg["d"], ls[0])
    return _gpx.KERNEL_IDS[kw["kernel"]["name"]], float(kk.get("variance", 1.0)), ls, float(kw["noise_prior"])


def cpu_oracle_step(cfg):
    """One fit+predict of the oracle on the host with every host core in the BLAS pool (torchrun exports
    OMP_NUM_THREADS=1, which would otherwise pin the baseline to one thread).  Returns (seconds, threads used)."""
    from oracle import gp_oracle as o
    kw = cfg["model"]["kwargs"]
    ncpu = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        ctx = threadpool_limits(limits=ncpu)
    except Exception:
        threadpool_info, ctx = None, None
    try:
        threads = max([p.get("num_threads", 1) for p in threadpool_info()] or [1]) if threadpool_info else 1
        t0 = time.perf_counter()
        m = o.OracleGPModel(kernel=kw["kernel"], noise_prior=kw["noise_prior"])
        m.init(cfg["X"], cfg["Y"])
        m.get_statistics(cfg["Xs"], full_cov=False)
        return time.perf_counter() - t0, threads
    finally:
        if ctx is not None:
            ctx.restore_original_limits()


WORKLOAD_TEXT = {"C4": "1-D natural-sound-shaped exact GP, N=60k, 100k test points",
                 "C5": "ARD-RBF exact GP d=8, N=200k, block-cyclic Cholesky",
                 "C3": "Matern-5/2 ARD kin40k-shaped d=8, N=40k",
                 "C2": "ARD-RBF Genz oscillatory d=10, N=20k",
                 "C1": "RBF Sinc 1-D, N=1000"}
FULL_SIZES = {"C1": (1000, 1, 1000), "C2": (20000, 10, 2500), "C3": (40000, 8, 4000), "C4": (60000, 1, 100000), "C5": (200000, 8, 2500)}
KERNEL_OF = {"C1": "GPyRBF", "C2": "GPyRBF", "C3": "GPyMatern52", "C4": "GPyRBF", "C5": "GPyRBF"}


def config_block(workload, N, d, Ns, world):
    """The `config` object of the JSON line -- identical for the CUDA arm and the reference arm of the same command."""
    return {"workload": "%s: %s" % (workload, WORKLOAD_TEXT[workload]), "N": N, "d": d, "N_test": Ns, "kernel": KERNEL_OF[workload],
            "parallelism": "1 GPU" if world == 1 else "block-column-cyclic x%d" % world,
            "l2_policy": "inputs larger than L2 (factor %.1f GB >> 126 MB)" % (N * (N + 128) / 2 * 8 / 1e9)}


def run_reference(args, workload, rank):
    """--impl reference: the reference's CPU implementation of the path.  GPy (where the reference's arithmetic lives)
    is not installable here, so this times the oracle port with all host BLAS threads, each step a BOUNDED SAMPLE of the
    workload (the same generator at N=12000, N*=2000 unless --N/--Ns say otherwise; full size would take hours on the
    host), reported in the metric's unit (algorithmic TFLOP/s of that sample)."""
    if rank != 0:
        return
    from thesis_code_b200 import synthetic
    t_start = time.perf_counter()
    budget = float(os.environ.get("GPX_BENCH_REF_BUDGET_S", "420"))
    sN, sNs = (args.N or 12000), (args.Ns or 2000)
    cfg = synthetic.make_config(workload, N=sN, Ns=sNs)
    warm = []
    for _ in range(max(1, args.warmup)):
        warm.append(cpu_oracle_step(cfg)[0])
        if time.perf_counter() - t_start > 0.25 * budget:
            break
    runs = []
    for _ in range(args.steps):
        runs.append(cpu_oracle_step(cfg))
        if time.perf_counter() - t_start + runs[-1][0] > budget:
            break
    t = float(np.mean([r[0] for r in runs]))
    threads = runs[0][1]
    val = algorithmic_flops(cfg["N"], cfg["Ns"]) / t / 1e12
    fN, fd, fNs = FULL_SIZES[workload]
    sample = ("oracle (numpy/scipy port of the GPy path) on the %s generator at N=%d, N*=%d per step (bounded sample; full size "
              "N=%d, N*=%d would take ~%.0f s per step at this rate)" % (workload, cfg["N"], cfg["Ns"], fN, fNs,
                                                                      algorithmic_flops(fN, fNs) / (val * 1e12)))
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": len(runs),
            "warmup": len(warm), "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(workload, fN, fd, fNs, args.gpus),
            "cpu_baseline": {"value": val, "unit": "TFLOP/s", "cores": threads, "kind": "port", "sample": sample,
                             "sample_N": cfg["N"], "sample_N_test": cfg["Ns"]},
            "e2e": {"value": val, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def golden_parity(h, workload, N, kid, variance, ls, noise, ard, torch):
    """Parity of THIS run's fitted state against the committed full-size golden vectors (tests/golden/
    exact_gp_golden_large.npz, generated by the oracle's arithmetic on the host): only for a workload / size it holds."""
    path = os.path.join(ROOT, "tests", "golden", "exact_gp_golden_large.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    if workload + "_N" not in z.files or int(z[workload + "_N"]) != N:
        return None
    Xs = np.ascontiguousarray(z[workload + "_Xs"])
    mean, var = h.predict(Xs, want_var=True, include_noise=True)
    nrm = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    lml = h.lml()[0]
    return {"vs": "tests/golden/exact_gp_golden_large.npz (oracle arithmetic at full size)", "test_points": int(Xs.shape[0]),
            "mean_normwise": nrm(mean, z[workload + "_mean"]), "var_normwise": nrm(var, z[workload + "_var_with_noise"]),
            "lml_rel": abs(lml - float(z[workload + "_lml"])) / abs(float(z[workload + "_lml"])), "gate": 1e-8}


def other_configs_and_latency(local_rank, torch):
    """Outside every timed region, single GPU: (i) one fit+predict of the parity configs C1-C3 at FULL size through the
    public GPModel API with host numpy buffers (second of two runs each), so that every BASELINE config that fits one GPU has
    a driver-run timing; (ii) the latency path of SURVEY.md section 8(f) row 3 on the fitted C2 model (N = 20000): one-row
    predictions and one-point appends, host wall-clock medians."""
    from thesis_code_b200 import GPModel, synthetic
    out = {"what": "fit+predict through GPModel.init + get_statistics(full_cov=False) on host numpy arrays, second of two runs; "
                   "latency = median host wall-clock per call on the fitted C2 model"}

    def med(fn, n):
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return float(np.median(ts))

    for wl in ("C1", "C2", "C3"):
        fN = FULL_SIZES[wl][0]
        c = synthetic.make_config(wl, N=fN + (8 if wl == "C2" else 0))
        X, Y = c["X"][:fN], c["Y"][:fN]
        m = GPModel.from_config(c["model"]["kwargs"])
        m.device = local_rank
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m.init(X, Y)
            ph = dict(m._get_handle().phase_ms())
            mean, var = m.get_statistics(c["Xs"], full_cov=False)
            t = time.perf_counter() - t0
        out[wl] = {"N": fN, "d": c["d"], "N_test": c["Ns"], "fit_predict_sec": t,
                   "tflops": algorithmic_flops(fN, c["Ns"]) / t / 1e12,
                   "cholesky_tflops": float(fN) ** 3 / 3 / (ph["cholesky"] * 1e-3) / 1e12 if ph.get("cholesky", 0) > 0 else None,
                   "var_min": float(var.min())}
        if wl == "C2":
            h2 = m._get_handle()
            # one optimiser evaluation of do_optimize=True: the fit above + the LML gradient (L^-T by the
            # DMMA triangular solve, Ky^-1 = L^-T L^-1, then the tile-wise reduction against dK/dtheta): N^3 flop in total
            for _ in range(2):
                t0 = time.perf_counter(); g = h2.lml_grad(m.kernel.lengthscale.size); tg = time.perf_counter() - t0
            fit_s = (ph["kbuild"] + ph["cholesky"] + ph["solve"] + ph["lml"]) * 1e-3
            out["optimizer_eval_C2"] = {"N": fN, "params": int(len(g)), "fit_ms": fit_s * 1e3, "lml_grad_ms": tg * 1e3,
                                        "tflops": float(fN) ** 3 / (fit_s + tg) / 1e12}
            row = c["Xs"][:1]
            lat = {"N": fN, "refit_ms": ph["kbuild"] + ph["cholesky"] + ph["solve"] + ph["lml"]}
            h2.set_option("latency_inverse", 0)
            m.get_statistics(row, full_cov=False); m.get_mean(row)
            lat["mean_only_1row_us"] = med(lambda: m.get_mean(row), 50) * 1e6
            lat["mean_var_1row_us_substitution"] = med(lambda: m.get_statistics(row, full_cov=False), 20) * 1e6
            h2.set_option("latency_inverse", 1)
            t0 = time.perf_counter(); m.get_statistics(row, full_cov=False); lat["inverse_build_ms"] = (time.perf_counter() - t0) * 1e3
            lat["mean_var_1row_us_inverse"] = med(lambda: m.get_statistics(row, full_cov=False), 100) * 1e6
            ts = []
            for i in range(4):
                t0 = time.perf_counter(); m.add_observations(c["X"][fN + i:fN + i + 1], c["Y"][fN + i:fN + i + 1]); ts.append(time.perf_counter() - t0)
            lat["append_1pt_ms_inverse"] = float(np.median(ts)) * 1e3
            h2.set_option("latency_inverse", 0)
            ts = []
            for i in range(4, 8):
                t0 = time.perf_counter(); m.add_observations(c["X"][fN + i:fN + i + 1], c["Y"][fN + i:fN + i + 1]); ts.append(time.perf_counter() - t0)
            lat["append_1pt_ms_substitution"] = float(np.median(ts)) * 1e3
            out["latency_C2"] = lat
        m._handle.close()
        m._handle = None
    return out


def main():
    t_proc = time.perf_counter()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--N", type=int, default=None)
    ap.add_argument("--Ns", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the untimed C1-C3 / latency-path block of the 1-GPU line")
    ap.add_argument("--trace", action="store_true", help="one extra (untimed) fit with the per-panel timeline of the Cholesky")
    ap.add_argument("--budget-s", type=float, default=float(os.environ.get("GPX_BENCH_BUDGET_S", "600")),
                    help="wall-clock budget of the whole run: the number of full-size timed steps is cut to fit it")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = args.workload or ("C4" if args.gpus == 1 else "C5")
    if args.steps is None:
        # defaults that finish within minutes: C4 on one GPU is ~12.5 s / step, C5 (N=200k) is ~42 / 21 / 11 s on 2 / 4 / 8 GPUs
        args.steps = 3 if args.gpus == 1 else 2

    if args.impl == "reference":
        run_reference(args, workload, rank)
        return

    import torch
    from thesis_code_b200 import GPModel, _gpx, synthetic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible; the exact-GP path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cfg = synthetic.make_config(workload, N=args.N, Ns=args.Ns)
    N, d, Ns = cfg["N"], cfg["d"], cfg["Ns"]
    kid, variance, ls, noise = kernel_args(cfg)
    ard = bool(cfg["model"]["kwargs"]["kernel"]["kwargs"].get("ARD", False))
    flops = algorithmic_flops(N, Ns)

    h = _gpx.Handle(local_rank)
    if world > 1:
        from thesis_code_b200 import distributed as gdist
        gdist.init_handle_comm(h, dist, rank, world)
    h.set_option("gemm_timing", 1)
    ext = torch.cuda.ExternalStream(h.stream_ptr(), device=torch.device("cuda", local_rank))

    # inputs resident in HBM for `value`
    dX = torch.from_numpy(cfg["X"]).cuda()
    dY = torch.from_numpy(cfg["Y"]).cuda()
    dXs = torch.from_numpy(cfg["Xs"]).cuda()
    dmean = torch.empty((Ns, 1), dtype=torch.float64, device="cuda")
    dvar = torch.empty((Ns,), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    from thesis_code_b200 import torch_bridge

    def device_step(n=N, ns=Ns):   # device tensors in / out through the thin torch bridge (thesis_code_b200/torch_bridge.py)
        torch_bridge.fit(h, dX[:n], dY[:n], kid, variance, ls, noise, ARD=ard)
        torch_bridge.predict(h, dXs[:ns], out=(dmean[:ns], dvar[:ns]))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def agree(x, op):   # every rank must take the same decisions about step counts
        if dist is None:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return float(t.item())

    # ---- warm-up: W-1 steps on a prefix of the same data (kernels, NCCL communicators, streams: < 1 s each), then ONE at
    # full size (allocation of the factor, page faults) whose wall time also sizes the timed region to the budget
    warm_small_n = min(N, 16384)
    n_small = max(0, args.warmup - 1)
    for _ in range(n_small):
        device_step(warm_small_n, min(Ns, 4096))
    barrier()
    t0 = time.perf_counter()
    device_step()
    barrier()
    t_full = agree(time.perf_counter() - t0, dist.ReduceOp.MAX if dist is not None else None)
    warm_done = n_small + 1
    e2e_steps = 1 if t_full > 20.0 else max(1, min(args.steps, 2))
    cpu_reserve = 0.0 if (args.no_cpu_baseline or rank != 0) else 25.0
    check_reserve = 0.0 if args.no_check else 0.15 * t_full + 10.0
    extras_reserve = 25.0 if (world == 1 and not args.no_extras) else 0.0
    left = args.budget_s - (time.perf_counter() - t_proc) - e2e_steps * 1.05 * t_full - cpu_reserve - check_reserve - extras_reserve - 20.0
    steps = int(agree(max(1, min(args.steps, int(left / t_full))), dist.ReduceOp.MIN if dist is not None else None))

    h.gemm_stats()
    launches0 = h.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase_acc = {}
    e0.record(ext)
    for i in range(steps):
        device_step()
        for k, v in h.phase_ms().items():
            phase_acc[k] = phase_acc.get(k, 0.0) + v
    e1.record(ext)
    barrier()
    sampler.stop_flag = True
    ms_total = e0.elapsed_time(e1)
    gemm_ms, gemm_issued, gemm_launches = h.gemm_stats()
    launches = h.launch_count() - launches0
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / steps
    value = flops / (ms_step * 1e-3) / 1e12
    phases = {k: v / steps for k, v in phase_acc.items()}

    # ---- self-check of the state the timed steps left behind (outside the timed region, on the device) --------------------
    check = None
    if not args.no_check:
        check = h.selfcheck()          # collective when sharded
        check["var_min"] = float(dvar.min().item())
        check["var_max"] = float(dvar.max().item())
        lml = h.lml()[0]
        check["lml"] = lml
        if dist is not None:
            t = torch.tensor([lml, -lml], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            check["lml_spread_over_ranks"] = float(t[0].item() + t[1].item())
        else:
            gp = golden_parity(h, workload, N, kid, variance, ls, noise, ard, torch)
            if gp is not None:
                check["golden_parity"] = gp
        check["what"] = ("solve_residual = ||Ky alpha - y|| / ||y||, factor_residual = ||L L^T v - Ky v|| / ||Ky v|| with Ky rebuilt "
                         "tile-wise by the K-build kernels; var_min / var_max over all test points of the last timed step")

    # ---- optional: per-panel timeline of one extra factorisation (outside every timed region) --------------------------
    trace_summary = None
    if args.trace:
        h.set_option("trace", 1)
        h.fit_device(dX.data_ptr(), N, d, dY.data_ptr(), 1, kid, variance, ls, noise, ARD=ard)
        tr = h.trace()            # (npanels, 6): factor start, factor end, panel ready on U, next-panel update done, fwd done, trailing done
        h.set_option("trace", 0)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        np.save(os.path.join(ROOT, "gpurun_out", "trace_%dgpu_rank%d.npy" % (world, rank)), tr)
        ready, bend = tr[:, 2], tr[:, 5]
        waits = np.maximum(0.0, ready[1:] - bend[:-1])          # stream U idle before panel p is ready
        own = tr[:, 1] > 0
        trace_summary = {"panels": int(tr.shape[0]), "cholesky_ms": float(bend[-1]), "update_stream_wait_ms": float(waits.sum()),
                         "update_stream_wait_frac": float(waits.sum() / bend[-1]), "panels_with_wait_over_50us": int((waits > 0.05).sum()),
                         "owner_factor_ms_mean": float((tr[own, 1] - tr[own, 0]).mean()) if own.any() else None,
                         "what": "wait = time the update stream sits idle between finishing the trailing update of panel p-1 and "
                                 "panel p being ready (factored + broadcast); rank 0's view"}
        barrier()

    # ---- e2e: the public GPModel API with host numpy buffers (H2D + D2H inside the timed region) ----------------
    model = GPModel.from_config(cfg["model"]["kwargs"])
    model.device = local_rank
    model._handle = h          # same handle: the NCCL communicator (world > 1) is already set up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        model.init(cfg["X"], cfg["Y"])
        mean, var = model.get_statistics(cfg["Xs"], full_cov=False)
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h.gemm_stats()
    e2e_val = flops / (e2e_ms * 1e-3) / 1e12
    h2d = int(cfg["X"].nbytes + cfg["Y"].nbytes + cfg["Xs"].nbytes)
    d2h = int(mean.nbytes // 1 + Ns * 8)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ------------------------------------------------------------------------
    # The timed launches are the tile GEMMs of the main stream on rank 0: the trailing updates of the Cholesky and the
    # whole variance solve (> 97% of the N^3/3 + N^2 N* flops; the panel-factorisation GEMMs overlap them on a second
    # stream and are excluded so no wall time is counted twice).  Their work is 2*128^3 flop per output tile and k-tile.
    achieved = gemm_issued / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else None
    chol_tflops = (float(N) ** 3 / 3) / (phases["cholesky"] * 1e-3) / 1e12
    trsm_tflops = (float(N) ** 2 * Ns) / (phases["predict_trsm"] * 1e-3) / 1e12 if phases.get("predict_trsm", 0) > 0 else None
    peaks = measured_peaks()
    roofline = {"bound": "tensor", "kernel": "gemm_dmma_kernel (FP64 DMMA tile GEMM)", "achieved": achieved,
                "peak": FP64_DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": (achieved / FP64_DMMA_PEAK_TFLOPS) if achieved else None,
                "traffic": GEMM_TRAFFIC["dram_bytes_per_launch"],
                "peak_source": "fallback: MEASURED_PEAKS.json has no FP64 figure; 36.96 TFLOP/s = register-resident "
                               "mma.sync.m8n8k4.f64 loop measured on this pool's B200 (profiles/r01_fp64_peak.txt)",
                "gemm_launches": gemm_launches, "gemm_ms_per_step": gemm_ms / steps,
                "timed_flops_share_of_step": gemm_issued / (flops * steps / max(world, 1)),
                "per_gpu": world > 1,
                "traffic_note": GEMM_TRAFFIC["note"],
                "kernel_share_of_step": gemm_ms / ms_total if ms_total > 0 else None}

    # ---- the other single-GPU configs and the latency path (untimed extras) ----------------------------------------------
    extras = None
    if world == 1 and not args.no_extras:
        try:
            extras = other_configs_and_latency(local_rank, torch)
        except Exception as e:     # never lose the headline line to an extra
            extras = {"error": repr(e)}

    # ---- CPU baseline: oracle port on a bounded sample ------------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline:
        sN, sNs = min(N, 12000), min(Ns, 2000)
        scfg = synthetic.make_config(workload, N=sN, Ns=sNs)
        t, threads = cpu_oracle_step(scfg)
        cpu = {"value": algorithmic_flops(sN, sNs) / t / 1e12, "unit": "TFLOP/s", "cores": threads, "kind": "port",
               "sample": "oracle (numpy/scipy port of the GPy path) on the %s generator at N=%d, N*=%d: %.1f s" % (workload, sN, sNs, t)}

    line = {"metric": METRIC, "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm_done,
            "steps_requested": args.steps, "warmup_requested": args.warmup,
            "warmup_detail": "%d on the first %d training rows, then 1 at full size" % (n_small, warm_small_n),
            "budget_s": args.budget_s, "wall_s": time.perf_counter() - t_proc,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_block(workload, N, d, Ns, world),
            "fit_predict_sec": ms_step * 1e-3, "cholesky_tflops": chol_tflops,
            "cholesky_frac_of_fp64_peak": chol_tflops / FP64_DMMA_PEAK_TFLOPS / max(world, 1),
            "variance_trsm_tflops": trsm_tflops, "phase_ms": phases,
            # HBM-side rates of the non-tensor passes (per GPU): K-build writes 8 B per lower-triangular element once,
            # the backward substitution reads L once (the forward pass is fused into the factorisation), the K*-build
            # evaluates N x N* kernel entries (ALU/exp bound)
            "kbuild_gbs": (8.0 * N * (N + 1) / 2 / max(world, 1) + 8.0 * N * d) / (phases["kbuild"] * 1e-3) / 1e9,
            "kbuild_gevals_per_s": (N * (N + 1) / 2.0 / max(world, 1)) / (phases["kbuild"] * 1e-3) / 1e9,
            "backward_solve_gbs": (8.0 * N * (N + 1) / 2 / max(world, 1)) / (phases["solve"] * 1e-3) / 1e9,
            "kstar_gevals_per_s": (float(N) * Ns / max(world, 1)) / (phases["predict_kstar"] * 1e-3) / 1e9 if phases.get("predict_kstar", 0) > 0 else None,
            "hbm_peak_gbs": peaks.get("hbm_gbs"),
            "roofline": roofline, "cpu_baseline": cpu, "check": check, "trace": trace_summary, "other_configs": extras,
            "e2e": {"value": e2e_val, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "steps": e2e_steps,
                    "api": "GPModel.init(X, Y) + get_statistics(Xtest, full_cov=False) on numpy arrays"},
            "gpu_launches": launches, "clocks": sampler.summary()}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
