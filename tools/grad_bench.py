"""One optimiser evaluation of do_optimize=True at full size: gpx_fit + gpx_lml_grad on the C2 / C3
generators, device time by host wall-clock around the synchronous C-ABI calls.  python tools/grad_bench.py > gpurun_out/grad.json"""
import json, sys, time
sys.path.insert(0, '.')
import numpy as np
from thesis_code_b200 import _gpx, synthetic

out = []
h = _gpx.Handle(0)
for wl, N in (("C1", 1000), ("C2", 20000), ("C3", 40000)):
    c = synthetic.make_config(wl, N=N, Ns=8)
    kw = c["model"]["kwargs"]; kk = kw["kernel"]["kwargs"]
    ls = np.atleast_1d(np.asarray(kk.get("lengthscale", 1.0), dtype=np.float64))
    ard = bool(kk.get("ARD", False))
    if ard and ls.size == 1:
        ls = np.full(c["d"], ls[0])
    kid = _gpx.KERNEL_IDS[kw["kernel"]["name"]]
    res = {"workload": wl, "N": N, "d": c["d"], "params": int(ls.size + 3)}
    for rep in range(3):
        t0 = time.perf_counter(); h.fit(c["X"], c["Y"], kid, 1.0, ls, kw["noise_prior"], ARD=ard); t1 = time.perf_counter()
        g = h.lml_grad(ls.size); t2 = time.perf_counter()
    res["fit_ms"] = (t1 - t0) * 1e3
    res["grad_ms"] = (t2 - t1) * 1e3
    res["grad_tflops"] = (2.0 * N ** 3 / 3) / (t2 - t1) / 1e12
    res["eval_tflops"] = (1.0 * N ** 3) / (t2 - t0) / 1e12
    res["grad"] = [float(x) for x in g]
    out.append(res)
    print(json.dumps(res), file=sys.stderr, flush=True)
print(json.dumps(out))
