"""Generates tests/golden/exact_gp_golden_large.npz: the oracle's arithmetic run ONCE at the FULL size of BASELINE
configs C3 (N=40000, d=8, Matern-5/2 ARD) and C4 (N=60000, d=1, RBF -- the bench.py workload), so that the GPU tests can
check parity at the benchmark's own size without an O(N^3) CPU run on the GPU box.

Why not simply oracle.fit(): (i) oracle.kernel_matrix needs several N^2 temporaries (> 62 GB at N=60000) and (ii) scipy's
bundled OpenBLAS segfaults inside dpotrf at N >= 40000 on this host (dmesg: libscipy_openblas, write fault).  So the same
arithmetic is evaluated TILE-wise: Ky is built per 2048 x 2048 tile with the oracle's own functions (scaled_dist ->
k_of_r; exact-zero distance on the diagonal; + (noise+jitter) I), only the lower triangle is kept (14.4 GB at C4), and the
factorisation is the textbook right-looking blocked Cholesky on those tiles -- dpotrf on diagonal tiles, dtrsm, dsyrk/dgemm
-- which is also what LAPACK's dpotrf does internally.  `--selftest` checks the tiled path against oracle.fit / predict
at N=6000 (agreement ~1e-13, i.e. rounding-order level).

Stored per config: the first 512 test points of the config's own test set with their mean / var (noise included), LML,
logdet, the quadratic form and slices of alpha.  About 10 min for C4, 3 min for C3 on 8 cores.

Run:  python tests/golden/make_golden_large.py [--selftest] [C3 C4]     (no GPU, no /root/reference needed)
"""
import os
import sys
import time

import numpy as np
import scipy.linalg as sla
from scipy.linalg import blas

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import gp_oracle as o  # noqa: E402
from thesis_code_b200 import synthetic  # noqa: E402

NT = 512          # test points kept
NB = 2048         # tile size of the CPU factorisation


def tiled_fit_predict(c, n_test=NT, nb=NB):
    kw = c["model"]["kwargs"]
    kk = kw["kernel"]["kwargs"]
    kname = kw["kernel"]["name"]
    X, Y, Xs = c["X"], c["Y"], c["Xs"][:n_test]
    N = X.shape[0]
    ARD = bool(kk.get("ARD", False))
    ls = o._lengthscale_vector(kk.get("lengthscale"), X.shape[1], ARD)
    variance, noise, jitter = float(kk.get("variance", 1.0)), float(kw["noise_prior"]), o.DEFAULT_JITTER
    edges = list(range(0, N, nb)) + [N]
    B = len(edges) - 1
    sl = [slice(edges[i], edges[i + 1]) for i in range(B)]
    t0 = time.time()
    T = {}
    for i in range(B):
        for j in range(i + 1):
            if i == j:
                r = o.scaled_dist(X[sl[i]], None, ls, ARD)        # symmetric form: exact-zero diagonal
                blk = o.k_of_r(kname, r, variance)
                blk[np.diag_indices_from(blk)] += noise + jitter
            else:
                blk = o.k_of_r(kname, o.scaled_dist(X[sl[i]], X[sl[j]], ls, ARD), variance)
            T[i, j] = np.asfortranarray(blk)
    t1 = time.time()
    # right-looking blocked Cholesky, lower, in place on the tiles
    for k in range(B):
        Lkk, info = sla.lapack.dpotrf(T[k, k], lower=1, overwrite_a=1, clean=1)
        assert info == 0, (k, info)
        T[k, k] = Lkk
        for i in range(k + 1, B):   # L_ik = A_ik L_kk^-T
            T[i, k] = blas.dtrsm(1.0, Lkk, T[i, k], side=1, lower=1, trans_a=1, overwrite_b=1)
        for i in range(k + 1, B):
            T[i, i] = blas.dsyrk(-1.0, T[i, k], beta=1.0, c=T[i, i], lower=1, overwrite_c=1)
            for j in range(k + 1, i):
                T[i, j] = blas.dgemm(-1.0, T[i, k], T[j, k], beta=1.0, c=T[i, j], trans_b=1, overwrite_c=1)
    t2 = time.time()
    # alpha = Ky^-1 (Y - m): forward then backward substitution over the tiles
    resid = Y - 0.0
    z = resid.copy()
    for i in range(B):
        for j in range(i):
            z[sl[i]] -= T[i, j] @ z[sl[j]]
        z[sl[i]] = sla.solve_triangular(T[i, i], z[sl[i]], lower=True, check_finite=False)
    alpha = z.copy()
    for i in reversed(range(B)):
        for j in range(i + 1, B):
            alpha[sl[i]] -= T[j, i].T @ alpha[sl[j]]
        alpha[sl[i]] = sla.solve_triangular(T[i, i], alpha[sl[i]], lower=True, trans='T', check_finite=False)
    diagL = np.concatenate([np.diag(T[i, i]) for i in range(B)])
    logdet = 2.0 * np.sum(np.log(diagL))
    quad = float(np.sum(alpha * resid))
    lml = 0.5 * (-Y.size * o.LOG_2_PI - Y.shape[1] * logdet - quad)
    # predict: mean = K*^T alpha, var = sigma_f^2 - colsumsq(L^-1 K*) + sigma_n^2
    Kx = o.kernel_matrix(kname, X, Xs, variance, ls, ARD)
    mean = Kx.T @ alpha
    V = Kx
    for i in range(B):
        for j in range(i):
            V[sl[i]] -= T[i, j] @ V[sl[j]]
        V[sl[i]] = sla.solve_triangular(T[i, i], V[sl[i]], lower=True, check_finite=False)
    var = variance - np.sum(np.square(V), 0) + noise
    t3 = time.time()
    return dict(N=N, Xs=Xs, mean=mean, var=var, lml=lml, logdet=logdet, quad=quad, alpha=alpha, diagL=diagL,
                seconds=np.array([t1 - t0, t2 - t1, t3 - t2]))


def run(name):
    c = synthetic.make_config(name)
    r = tiled_fit_predict(c)
    print("%s: N=%d K-build %.0f s, Cholesky %.0f s, solves+predict %.0f s; lml=%.12e" % ((name, r["N"]) + tuple(r["seconds"]) + (r["lml"],)),
          flush=True)
    pre = name + "_"
    alpha = r["alpha"]
    return {pre + "N": np.array(r["N"]), pre + "Xs": r["Xs"], pre + "mean": r["mean"], pre + "var_with_noise": r["var"],
            pre + "lml": np.array(r["lml"]), pre + "logdet": np.array(r["logdet"]), pre + "quad": np.array(r["quad"]),
            pre + "alpha_head": alpha[:4096].copy(), pre + "alpha_tail": alpha[-4096:].copy(),
            pre + "alpha_absmax": np.array(np.abs(alpha).max()), pre + "diagL_head": r["diagL"][:4096].copy(),
            pre + "seconds": r["seconds"]}


def selftest():
    """the tiled path against the plain oracle at a size where both run"""
    for name in ("C3", "C4", "C2"):
        c = synthetic.make_config(name, N=6000, Ns=NT)
        r = tiled_fit_predict(c, nb=1024)
        kw = c["model"]["kwargs"]
        om = o.OracleGPModel(kernel=kw["kernel"], noise_prior=kw["noise_prior"])
        om.init(c["X"], c["Y"])
        m, v = om.get_statistics(c["Xs"], full_cov=False)
        nrm = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
        e = (nrm(r["mean"], m[0]), nrm(r["var"], v[0, :, 0]), abs(r["lml"] - om.state["lml"]) / abs(om.state["lml"]),
             nrm(r["alpha"], om.state["alpha"]))
        print("selftest %s N=6000: tiled vs oracle.fit  mean %.1e  var %.1e  lml %.1e  alpha %.1e" % ((name,) + e), flush=True)
        assert e[0] < 1e-10 and e[1] < 1e-10 and e[2] < 1e-12, e


def main():
    args = [a for a in sys.argv[1:]]
    if "--selftest" in args:
        selftest()
        args.remove("--selftest")
        if not args:
            return
    names = [a.upper() for a in args] or ["C3", "C4"]
    path = os.path.join(HERE, "exact_gp_golden_large.npz")
    out = {}
    if os.path.exists(path):
        z = np.load(path)
        out = {k: z[k] for k in z.files}
    for n in names:
        out.update(run(n))
        np.savez_compressed(path, **out)
    print("wrote", path, sorted(set(k.split("_")[0] for k in out)))


if __name__ == "__main__":
    main()
