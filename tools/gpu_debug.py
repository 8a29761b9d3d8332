import sys, time
import numpy as np
sys.path.insert(0, '.')
from oracle import gp_oracle as o
from thesis_code_b200 import _gpx, GPModel, synthetic
np.set_printoptions(linewidth=250, precision=1)
h = _gpx.Handle(0)
for N in [3500, 4000, 4096, 4200, 5000]:
    c = synthetic.make_config("C5", N=N, Ns=300)
    kw = c['model']['kwargs']['kernel']['kwargs']
    ls = np.array(kw['lengthscale'])
    K = h.kernel_matrix(c['X'], None, 0, 1.0, ls, 10.0)
    Ko = np.tril(o.kernel_matrix("GPyRBF", c['X'], None, 1.0, ls, True) + 10.0 * np.eye(N))
    print("N", N, "kbuild err", np.abs(K - Ko).max())
    for W in [1, 4]:
        h.set_option("outer_tiles", W)
        try:
            h.fit(c['X'], c['Y'], 0, 1.0, ls, 10.0)
        except Exception as e:
            print("  W", W, "fit failed:", e); continue
        alpha, L = h.export(N, 1, want_L=True)
        Lo = np.linalg.cholesky(Ko + np.tril(Ko, -1).T)
        nt = (N + 127) // 128
        E = np.zeros((nt, nt))
        for i in range(nt):
            for j in range(i + 1):
                E[i, j] = np.abs(L[i*128:(i+1)*128, j*128:(j+1)*128] - Lo[i*128:(i+1)*128, j*128:(j+1)*128]).max()
        bad = np.argwhere(E > 1e-9)
        print("  W", W, "max err", E.max(), "nbad", len(bad), "bad tiles (first 30):", bad[:30].tolist())
