"""SURVEY.md section 8(f) row 3 numbers: microseconds per 1-row predict and milliseconds per 1-point append, next to the
full refit, at N = 1k and 20k (C2-shaped data, d = 10, ARD RBF).  python tools/latency_bench.py > gpurun_out/latency.json"""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np
from thesis_code_b200 import GPModel, synthetic

out = {"what": "host wall-clock per call through the public GPModel API (numpy in / numpy out), median of the repeats", "cases": []}
for N in (1000, 20000):
    c = synthetic.make_config("C2", N=N + 96, Ns=512)
    kw = c["model"]["kwargs"]
    m = GPModel.from_config(kw)
    t0 = time.perf_counter(); m.init(c["X"][:N], c["Y"][:N]); t_fit_first = time.perf_counter() - t0
    t0 = time.perf_counter(); m.init(c["X"][:N], c["Y"][:N]); t_fit = time.perf_counter() - t0
    h = m._get_handle()
    row = c["Xs"][:1]
    res = {"N": N, "d": c["d"], "full_refit_ms": t_fit * 1e3}
    h.set_option("latency_inverse", 0)      # substitution route first; the explicit-inverse route is timed below
    for graph in (1, 0):
        h.set_option("predict_graph", graph)
        for name, fn in (("mean_only_1row", lambda: m.get_mean(row)),
                         ("mean_var_1row", lambda: m.get_statistics(row, full_cov=False)),
                         ("mean_var_8rows", lambda: m.get_statistics(c["Xs"][:8], full_cov=False)),
                         ("mean_var_128rows", lambda: m.get_statistics(c["Xs"][:128], full_cov=False))):
            for _ in range(5):
                fn()
            ts = []
            for _ in range(200 if N <= 1000 else 60):
                t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
            res["%s_us_graph%d" % (name, graph)] = float(np.median(ts) * 1e6)
    h.set_option("predict_graph", 1)
    # explicit inverse of the factor (option latency_inverse): the one-off build, then one triangular GEMV per call
    h.set_option("latency_inverse", 1)
    t0 = time.perf_counter(); m.get_statistics(row, full_cov=False); res["inverse_build_plus_first_call_ms"] = (time.perf_counter() - t0) * 1e3
    for name, fn in (("mean_var_1row", lambda: m.get_statistics(row, full_cov=False)),
                     ("mean_var_4rows", lambda: m.get_statistics(c["Xs"][:4], full_cov=False)),
                     ("mean_var_8rows", lambda: m.get_statistics(c["Xs"][:8], full_cov=False))):
        for _ in range(5):
            fn()
        ts = []
        for _ in range(200):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        res["%s_us_inverse" % name] = float(np.median(ts) * 1e6)
    # appends with the inverse resident: forward pass, new columns of the inverse and alpha are three GEMV sweeps
    ts = []
    for i in range(16):
        t0 = time.perf_counter(); m.add_observations(c["X"][N + 64 + i:N + 65 + i], c["Y"][N + 64 + i:N + 65 + i]); ts.append(time.perf_counter() - t0)
    res["append_1pt_ms_median_inverse"] = float(np.median(ts) * 1e3)
    ts = []
    for _ in range(50):
        t0 = time.perf_counter(); m.get_statistics(row, full_cov=False); ts.append(time.perf_counter() - t0)
    res["mean_var_1row_us_inverse_after_appends"] = float(np.median(ts) * 1e6)
    h.set_option("latency_inverse", 0)
    # raw C-ABI call (no GPModel shape plumbing)
    ts = []
    for _ in range(200):
        t0 = time.perf_counter(); h.predict(row, want_var=False); ts.append(time.perf_counter() - t0)
    res["cabi_mean_only_1row_us"] = float(np.median(ts) * 1e6)
    # appends: one point at a time (inside a tile, then across the tile boundary), and 16 at once
    ts = []
    for i in range(40):
        t0 = time.perf_counter(); m.add_observations(c["X"][N + i:N + i + 1], c["Y"][N + i:N + i + 1]); ts.append(time.perf_counter() - t0)
    res["append_1pt_ms_median"] = float(np.median(ts) * 1e3)
    res["append_1pt_ms_max"] = float(np.max(ts) * 1e3)
    t0 = time.perf_counter(); m.add_observations(c["X"][N + 40:N + 48], c["Y"][N + 40:N + 48]); res["append_8pts_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); m.add_observations(c["X"][N + 48:N + 64], c["Y"][N + 48:N + 64]); res["append_16pts_first_ms"] = (time.perf_counter() - t0) * 1e3
    c2 = synthetic.make_config("C2", N=32, Ns=8)
    t0 = time.perf_counter(); m.add_observations(c2["X"][:16] * 0.999, c2["Y"][:16]); res["append_16pts_ms"] = (time.perf_counter() - t0) * 1e3
    m2 = GPModel.from_config(kw); m2.incremental = False           # the reference's route: concatenate + refit from scratch
    m2.init(c["X"][:N], c["Y"][:N])
    ts = []
    for i in range(4):
        t0 = time.perf_counter(); m2.add_observations(c["X"][N + i:N + i + 1], c["Y"][N + i:N + i + 1]); ts.append(time.perf_counter() - t0)
    res["reference_style_refit_1pt_ms"] = float(np.median(ts[1:]) * 1e3)
    h.set_option("latency_inverse", -1)
    res["append_speedup_vs_refit"] = res["reference_style_refit_1pt_ms"] / res["append_1pt_ms_median"]
    out["cases"].append(res)
    print(json.dumps(res), file=sys.stderr, flush=True)
print(json.dumps(out))
