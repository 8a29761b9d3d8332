"""Small fit + predict for Nsight Compute captures (development aid): python tools/ncu_fit.py [N] [Ns] [opts k=v,...]
Every kernel of the hot path launches at least once (K-build, POTRF, tile GEMM, forward / backward steps, LML, K*-build,
variance reduce)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from thesis_code_b200 import _gpx, synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
Ns = int(sys.argv[2]) if len(sys.argv) > 2 else 2500
c = synthetic.make_config("C5", N=N, Ns=Ns)
kk = c['model']['kwargs']['kernel']['kwargs']
h = _gpx.Handle(0)
h.set_option("lookahead", 0)
for kv in filter(None, (sys.argv[3] if len(sys.argv) > 3 else "").split(",")):
    k, v = kv.split("="); h.set_option(k, int(v))
h.fit(c['X'], c['Y'], 0, 1.0, np.array(kk['lengthscale']), 1e-2, ARD=True)
mean, var = h.predict(c['Xs'])
print("lml", h.lml()[0], "phases ms", {k: round(v, 2) for k, v in h.phase_ms().items()}, "var min", var.min())
