import sys; sys.path.insert(0, '.')
import numpy as np
from oracle import gp_oracle as o
from thesis_code_b200 import _gpx
h = _gpx.Handle(0)
rng = np.random.RandomState(11)
for N in (100, 700):
    X = rng.rand(N, 3); Y = np.stack([np.sin(3 * X.sum(1)), np.cos(2 * X[:, 0])], 1) + 0.05 * rng.randn(N, 2)
    for name in ("GPyRBF", "GPyExponential"):
        for ls in (np.array([0.5, 0.8, 1.1]), np.array([0.7])):
            h.fit(X, Y, _gpx.KERNEL_IDS[name], 1.3, ls, 0.05, mean_const=0.2)
            g = h.lml_grad(ls.size)
            go = o.lml_grad(o.fit(X, Y, name, 1.3, ls, ls.size > 1, 0.05, mean_const=0.2))
            print(N, name, ls.size, "gpu", g, "\n      oracle", go, "relerr", np.abs(g - go).max() / np.abs(go).max())
