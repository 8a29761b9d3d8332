import sys; sys.path.insert(0, '.')
import numpy as np
from thesis_code_b200 import _gpx, synthetic
N = int(sys.argv[1]); W = int(sys.argv[2])
c = synthetic.make_config("C5", N=N, Ns=256)
kk = c['model']['kwargs']['kernel']['kwargs']
import os
h = _gpx.Handle(0); h.set_option("outer_tiles", W); h.set_option("trace", 1)
for kv in filter(None, os.environ.get("GPX_OPTS", "").split(",")):
    k, v = kv.split("="); h.set_option(k, int(v))
for rep in range(2):
    h.fit(c['X'], c['Y'], 0, 1.0, np.array(kk['lengthscale']), 1e-2, ARD=True)
tr = h.trace(); ph = h.phase_ms()
nt = (N + 127) // 128
if W == 0:   # auto panel width: recover it from the number of panels
    W = next(w for w in range(1, 65) if (nt + w - 1) // w == tr.shape[0])
print("N", N, "W", W, "chol ms", ph['cholesky'], "TF/s", N**3/3/ph['cholesky']/1e9)
fact = tr[:, 1] - tr[:, 0]; a = tr[:, 3] - tr[:, 2]; f = np.zeros_like(a); b = tr[:, 5] - tr[:, 3]   # the forward steps run on their own stream
gaps = tr[1:, 2] - tr[:-1, 5]
print("sum: factor %.1f (a) %.1f fwd %.1f (b) %.1f  waits-for-panel %.1f  end %.1f" % (fact.sum(), a.sum(), f.sum(), b.sum(), gaps[gaps > 0].sum(), tr[-1, 5]))
# efficiency of (b): flops of (b) at panel p
np_ = tr.shape[0]
tot_flops = 0; 
for p in range(0, np_, max(1, np_ // 12)):
    p1 = min((p + 1) * W, nt); p2 = min((p + 2) * W, nt)
    rows = nt - p2
    tiles = rows * (rows + 1) / 2
    fl = tiles * (p1 - p * W) * 2 * 128 ** 3
    print("p%3d: fact %.2f a %.2f fwd %.2f b %.2f ms -> b %.1f TF/s  wait %.2f" % (p, fact[p], a[p], f[p], b[p], fl / max(b[p], 1e-9) / 1e9, gaps[p - 1] if p > 0 else 0))
