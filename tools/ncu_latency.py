"""Latency-path kernels for Nsight Compute (development aid): explicit-inverse GEMV sweeps at N = 20000 (C2-shaped, d = 10).
python tools/ncu_latency.py [N]"""
import sys
import numpy as np
sys.path.insert(0, '.')
from thesis_code_b200 import _gpx, synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
c = synthetic.make_config("C2", N=N + 8, Ns=16)
kk = c['model']['kwargs']['kernel']['kwargs']
h = _gpx.Handle(0)
h.set_option("predict_graph", 0)
h.set_option("latency_inverse", 1)
h.fit(c['X'][:N], c['Y'][:N], 0, 1.0, np.full(10, kk['lengthscale']), 1e-2, ARD=True)
for rows in (1, 1, 4, 8):
    mean, var = h.predict(c['Xs'][:rows])
h.append(c['X'][N:N + 1], c['Y'][N:N + 1])
mean, var = h.predict(c['Xs'][:1])
print("var", var, "lml", h.lml()[0])
