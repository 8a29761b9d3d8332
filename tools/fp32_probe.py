"""CPU probe behind DESIGN.md section 1 "fp32-with-jitter mode": the exact-GP arithmetic in numpy float32 with jitter 1e-4 against
float64, on the five BASELINE generators at N = 6000 (C1: 1000).  python tools/fp32_probe.py > profiles/r02_fp32_probe.txt"""
import sys, numpy as np, scipy.linalg as sl
sys.path.insert(0, '/root/repo')
from thesis_code_b200 import synthetic
def gp(X, Y, Xs, ls, var, noise, jitter, dt, kern):
    X = (X / ls).astype(dt); Xs = (Xs / ls).astype(dt); Y = Y.astype(dt)
    def K(A, B):
        r2 = np.maximum((A*A).sum(1)[:, None] + (B*B).sum(1)[None, :] - 2 * A @ B.T, 0)
        if kern == 'rbf': return (var * np.exp(-0.5 * r2)).astype(dt)
        r = np.sqrt(r2); s5 = np.sqrt(5.0)
        return (var * (1 + s5 * r + 5.0/3 * r2) * np.exp(-s5 * r)).astype(dt)
    Ky = K(X, X); np.fill_diagonal(Ky, var); Ky[np.diag_indices_from(Ky)] += dt(noise + jitter)
    L = sl.cholesky(Ky, lower=True)
    a = sl.cho_solve((L, True), Y)
    Ks = K(X, Xs)
    mean = Ks.T @ a
    T = sl.solve_triangular(L, Ks, lower=True)
    v = var + noise - (T*T).sum(0)
    lml = 0.5 * (-len(Y) * np.log(2*np.pi) - 2*np.log(np.diag(L).astype(np.float64)).sum() - float((a*Y).sum()))
    return mean.astype(np.float64).ravel(), v.astype(np.float64), lml
nrm = lambda a, b: float(np.abs(a-b).max()/np.abs(b).max())
for wl, N in (("C1", 1000), ("C2", 6000), ("C3", 6000), ("C4", 6000), ("C5", 6000)):
    c = synthetic.make_config(wl, N=N, Ns=500)
    kk = c["model"]["kwargs"]["kernel"]; kw = kk["kwargs"]
    ls = np.atleast_1d(np.asarray(kw.get("lengthscale", 1.0), float)); noise = c["model"]["kwargs"]["noise_prior"]
    kern = 'rbf' if kk["name"] == "GPyRBF" else 'm52'
    m64 = gp(c["X"], c["Y"], c["Xs"], ls, 1.0, noise, 1e-8, np.float64, kern)
    try:
        m32 = gp(c["X"], c["Y"], c["Xs"], ls, 1.0, noise, 1e-4, np.float32, kern)
        m64j = gp(c["X"], c["Y"], c["Xs"], ls, 1.0, noise, 1e-4, np.float64, kern)
        print(wl, N, "fp32 vs fp64 at the SAME jitter 1e-4: mean %.1e var %.1e lml %.1e" % (nrm(m32[0], m64j[0]), nrm(m32[1], m64j[1]), abs(m32[2]-m64j[2])/abs(m64j[2])))
        print(wl, N, "fp32+jitter1e-4 vs fp64: mean %.1e var %.1e lml %.1e | jitter alone (fp64, 1e-4 vs 1e-8): mean %.1e var %.1e lml %.1e" % (
            nrm(m32[0], m64[0]), nrm(m32[1], m64[1]), abs(m32[2]-m64[2])/abs(m64[2]),
            nrm(m64j[0], m64[0]), nrm(m64j[1], m64[1]), abs(m64j[2]-m64[2])/abs(m64[2])))
    except Exception as e:
        print(wl, N, "fp32 failed:", e)
