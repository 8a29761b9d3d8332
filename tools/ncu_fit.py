"""Small fit-only run for Nsight Compute captures of the big trailing-update GEMM (development aid)."""
import sys
import numpy as np
sys.path.insert(0, '.')
from thesis_code_b200 import _gpx, synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
c = synthetic.make_config("C5", N=N, Ns=256)
kk = c['model']['kwargs']['kernel']['kwargs']
h = _gpx.Handle(0)
h.set_option("lookahead", 0)
h.fit(c['X'], c['Y'], 0, 1.0, np.array(kk['lengthscale']), 1e-2)
print("lml", h.lml()[0], "chol ms", h.phase_ms()['cholesky'])
