"""Where does the end-to-end (host numpy in/out) step spend time beyond the device-timed step?  python tools/e2e_probe.py [C4]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from thesis_code_b200 import GPModel, synthetic
name = sys.argv[1] if len(sys.argv) > 1 else "C4"
stag = int(sys.argv[2]) if len(sys.argv) > 2 else 1
c = synthetic.make_config(name)
m = GPModel.from_config(c["model"]["kwargs"])
m._get_handle().set_option("gemm_stagger", stag)
for rep in range(3):
    t0 = time.perf_counter(); m.init(c["X"], c["Y"]); t1 = time.perf_counter()
    mean, var = m.get_statistics(c["Xs"], full_cov=False); t2 = time.perf_counter()
    ph = m.phase_times_ms()
    print("rep %d stagger %d: init %.3f s  get_statistics %.3f s | device phases (ms): %s | sum %.1f" % (
        rep, stag, t1 - t0, t2 - t1, {k: round(v, 1) for k, v in ph.items()}, sum(ph.values())), flush=True)
