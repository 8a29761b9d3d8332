"""Generates tests/golden/exact_gp_golden.npz.

The reference holds NO golden vectors for GPModel (parity unpinned by the reference's
own tests) and its arithmetic (GPy 1.9.6) is not importable here.  These fixtures are therefore produced by an
INDEPENDENT implementation the reference itself uses elsewhere -- scikit-learn's GaussianProcessRegressor
(/root/reference/src/models/lls_gp.py:4-6,61-65) -- with the kernel/noise conventions mapped onto GPy's:
    GPy RBF(variance v, lengthscale l)        == ConstantKernel(v) * RBF(l)
    GPy Matern52 / Matern32 / Exponential     == ConstantKernel(v) * Matern(l, nu=2.5 / 1.5 / 0.5)
    GPy Gaussian_noise.variance s2 (+1e-8)    == alpha = s2 + 1e-8
    GPy predict(include_likelihood=True) var  == sklearn std**2 + s2
    GPy log_likelihood()                      == sklearn log_marginal_likelihood(theta)
Run:  python tests/golden/make_golden.py      (needs scikit-learn; does not need a GPU or /root/reference)
"""
import os

import numpy as np
from sklearn.gaussian_process import GaussianProcessRegressor
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern

HERE = os.path.dirname(os.path.abspath(__file__))
NU = {"GPyMatern52": 2.5, "GPyMatern32": 1.5, "GPyExponential": 0.5}


def sk_kernel(name, variance, ls):
    ls = np.atleast_1d(ls)
    ls = ls if ls.size > 1 else float(ls[0])
    base = RBF(ls) if name == "GPyRBF" else Matern(ls, nu=NU[name])
    return ConstantKernel(variance) * base


def main():
    rng = np.random.RandomState(1234)
    cases = {}
    idx = 0
    for name in ["GPyRBF", "GPyMatern52", "GPyMatern32", "GPyExponential"]:
        for (N, d, ARD, noise, O) in [(40, 1, False, 1e-2, 1), (150, 3, True, 1e-1, 1), (131, 8, True, 1e-2, 2),
                                      (257, 2, False, 1e-3, 1)]:
            X = rng.uniform(-2, 2, (N, d))
            Y = np.sin(X.sum(1, keepdims=True) * np.arange(1, O + 1)[None, :]) + 0.1 * rng.randn(N, O)
            Xs = rng.uniform(-2.2, 2.2, (37, d))
            variance = float(rng.uniform(0.5, 2.0))
            ls = rng.uniform(0.4, 1.5, d if ARD else 1)
            gp = GaussianProcessRegressor(kernel=sk_kernel(name, variance, ls), alpha=noise + 1e-8, optimizer=None)
            gp.fit(X, Y)
            mean, std = gp.predict(Xs, return_std=True)
            mean = mean.reshape(37, O)
            std = np.asarray(std).reshape(37, -1)[:, 0]
            _, cov = gp.predict(Xs, return_cov=True)
            cov = np.asarray(cov)
            if cov.ndim == 3:
                cov = cov[:, :, 0]
            lml = float(gp.log_marginal_likelihood(gp.kernel_.theta))
            pre = "c%02d_" % idx
            cases[pre + "kernel"] = np.array(name)
            cases[pre + "X"] = X
            cases[pre + "Y"] = Y
            cases[pre + "Xs"] = Xs
            cases[pre + "variance"] = np.array(variance)
            cases[pre + "lengthscale"] = ls
            cases[pre + "ARD"] = np.array(ARD)
            cases[pre + "noise"] = np.array(noise)
            cases[pre + "mean"] = mean
            cases[pre + "var_with_noise"] = std ** 2 + noise
            cases[pre + "cov_with_noise"] = cov + noise * np.eye(37)
            cases[pre + "lml"] = np.array(lml)
            idx += 1
    cases["n_cases"] = np.array(idx)
    np.savez_compressed(os.path.join(HERE, "exact_gp_golden.npz"), **cases)
    print("wrote", idx, "cases")


if __name__ == "__main__":
    main()
