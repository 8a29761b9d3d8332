"""Phase timing of fit + predict at a given config/size (development aid)."""
import sys, time, os
import numpy as np
sys.path.insert(0, '.')
from thesis_code_b200 import _gpx, synthetic
cfg = sys.argv[1] if len(sys.argv) > 1 else "C3"
N = int(sys.argv[2]) if len(sys.argv) > 2 else None
Ns = int(sys.argv[3]) if len(sys.argv) > 3 else None
W = int(sys.argv[4]) if len(sys.argv) > 4 else 4
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
c = synthetic.make_config(cfg, N=N, Ns=Ns)
kw = c['model']['kwargs']; kk = kw['kernel']['kwargs']
h = _gpx.Handle(0)
h.set_option("outer_tiles", W)
if os.environ.get("GPX_MAX_CHUNK"): h.set_option("gemm_max_chunk", int(os.environ["GPX_MAX_CHUNK"]))
for kv in filter(None, os.environ.get("GPX_OPTS", "").split(",")):      # e.g. GPX_OPTS=gemm_strip=16,predict_batch_tiles=296
    k, v = kv.split("="); h.set_option(k, int(v))
N, Ns = c['N'], c['Ns']
ls = np.atleast_1d(np.array(kk['lengthscale'], dtype=float))
kid = _gpx.KERNEL_IDS[kw['kernel']['name']]
for rep in range(reps):
    t0 = time.time()
    h.fit(c['X'], c['Y'], kid, kk.get('variance', 1.0), ls, kw['noise_prior'], ARD=kk.get('ARD', False))
    t1 = time.time()
    mean, var = h.predict(c['Xs'])
    t2 = time.time()
    ph = h.phase_ms()
    chol_tf = (N ** 3 / 3) / (ph['cholesky'] * 1e-3) / 1e12
    trsm_tf = (float(N) ** 2 * Ns) / (max(ph['predict_trsm'], 1e-9) * 1e-3) / 1e12
    print("%s N=%d Ns=%d W=%d opts=%s: fit %.3fs predict %.3fs | chol %.1f ms = %.2f TF/s | var-trsm %.1f ms = %.2f TF/s | %s | lml %.6e launches %d mem %.1f GB" % (
        cfg, N, Ns, W, os.environ.get("GPX_OPTS", "-"), t1 - t0, t2 - t1, ph['cholesky'], chol_tf, ph['predict_trsm'], trsm_tf,
        {k: round(v, 2) for k, v in ph.items()}, h.lml()[0], h.launch_count(), h.bytes_allocated() / 1e9), flush=True)
print("mean/var finite:", np.isfinite(mean).all(), np.isfinite(var).all(), "var min", var.min())
