import sys, os
import numpy as np
sys.path.insert(0, '.')
from thesis_code_b200 import _gpx, synthetic
h = _gpx.Handle(0)
for N, reps in [(9000, 6)]:
    c = synthetic.make_config("C5", N=N, Ns=300)
    ls = np.array(c['model']['kwargs']['kernel']['kwargs']['lengthscale'])
    for W in [1, 4]:
        h.set_option("outer_tiles", W)
        Ls = []; As = []
        for rep in range(reps):
            h.fit(c['X'], c['Y'], 0, 1.0, ls, 10.0)
            alpha, L = h.export(N, 1, want_L=True)
            Ls.append(L); As.append(alpha)
        # majority reference: median elementwise
        dL = [float(np.abs(L - Ls[0]).max()) for L in Ls]
        dA = [float(np.abs(A - As[0]).max()) for A in As]
        print("N", N, "W", W, "sync", os.environ.get("GPX_SYNC_EACH"), "L diffs vs rep0", ["%.1e" % x for x in dL], "alpha diffs", ["%.1e" % x for x in dA])
