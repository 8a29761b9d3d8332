"""Generates tests/golden/reference_pinned.npz by EXECUTING REFERENCE CODE (unmodified, loaded from /root/reference by
oracle/ref_loader.py): the author's numpy/scipy exact GP ``GPVanillaModel`` (src/models/deprecated_models.py:8-48) and the
reference's ``NormalizerModel`` (src/models/normalizer.py:51-114) wrapped around it.  These vectors pin the oracle
(tests/test_oracle.py, CPU) and the CUDA path (tests/test_gpu_parity.py, GPU) to outputs of reference code that ran.

Convention map GPVanillaModel(lengthscale l_v, noise s) -> GPModel / oracle:
    K = exp(-|x-x'|^2 / l_v^2)          == GPyRBF(variance 1, lengthscale l_v / sqrt 2)
    Ky = K + s I  (no extra jitter)      == noise s, jitter 0
    mu = k Ky^-1 Y                        == mean, zero mean function
    diag(k** + s - k Ky^-1 k^T)           == variance with include_noise=True (GPy's include_likelihood)
Run (build container only; needs /root/reference):  python tests/golden/make_golden_reference.py
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from thesis_code_b200 import synthetic  # noqa: E402


def main():
    warnings.simplefilter("ignore", SyntaxWarning)
    ref = ref_loader.load()
    out = {}
    # ---- GPVanillaModel on the reference's own demo shape (C1: Sinc 1-D, N=1000) and on a 3-D problem ----------------
    cases = []
    c1 = synthetic.make_config("C1", N=1000, Ns=300)
    cases.append(("v0", c1["X"], c1["Y"], c1["Xs"], float(np.sqrt(2.0)), 1e-2))      # == C1's own RBF(l=1), noise 1e-2
    rng = np.random.RandomState(42)
    X = rng.uniform(-1, 1, (400, 3)); Xs = rng.uniform(-1, 1, (150, 3))
    Y = np.sin(3 * X[:, :1]) * np.cos(2 * X[:, 1:2]) + 0.5 * X[:, 2:3] + 0.05 * rng.randn(400, 1)
    cases.append(("v1", X, Y, Xs, 0.9, 5e-3))
    X = rng.uniform(0, 4, (257, 2)); Xs = rng.uniform(0, 4, (129, 2))
    Y = np.stack([np.sin(X.sum(1)), np.cos(X[:, 0] - X[:, 1])], 1)                    # two outputs
    cases.append(("v2", X, Y, Xs, 1.7, 1e-1))
    for tag, X, Y, Xs, lv, noise in cases:
        m = ref.GPVanillaModel(lengthscale=lv, noise=noise)
        m.init(X, Y)
        mu, var = m.get_statistics(Xs, full_cov=False)
        mu5, cov5 = m.get_statistics(Xs[:5], full_cov=True)
        out.update({tag + "_X": X, tag + "_Y": Y, tag + "_Xs": Xs, tag + "_lengthscale_v": np.array(lv), tag + "_noise": np.array(noise),
                    tag + "_mu": mu, tag + "_var": var, tag + "_cov5": cov5})
    # ---- reference NormalizerModel around the reference GPVanillaModel ---------------------------------------------------
    X = rng.uniform(0, 1, (500, 2)) * np.array([50.0, 0.2]) + np.array([-10.0, 3.0])
    Y = (np.sin(X[:, :1] / 7.0) * 20 + 100 + 40 * X[:, 1:2]) + 0.5 * rng.randn(500, 1)
    Xs = rng.uniform(0, 1, (140, 2)) * np.array([50.0, 0.2]) + np.array([-10.0, 3.0])
    for tag, nx, ny in (("n0", True, True), ("n1", False, True), ("n2", True, False)):
        nm = ref.NormalizerModel(model=ref.GPVanillaModel(lengthscale=0.8, noise=2e-2), normalize_input=nx, normalize_output=ny)
        nm.init(X, Y)
        mu, var = nm.get_statistics(Xs, full_cov=False)
        mu5, cov5 = nm.get_statistics(Xs[:5], full_cov=True)
        out.update({tag + "_X": X, tag + "_Y": Y, tag + "_Xs": Xs, tag + "_lengthscale_v": np.array(0.8), tag + "_noise": np.array(2e-2),
                    tag + "_normalize_input": np.array(nx), tag + "_normalize_output": np.array(ny), tag + "_mu": mu, tag + "_var": var,
                    tag + "_cov5": cov5})
        if nx:
            out[tag + "_x_mean"], out[tag + "_x_std"] = nm.X_normalizer._X_mean, nm.X_normalizer._X_std
        if ny:
            out[tag + "_y_mean"], out[tag + "_y_std"] = nm.Y_normalizer._X_mean, nm.Y_normalizer._X_std
    path = os.path.join(HERE, "reference_pinned.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "arrays")


if __name__ == "__main__":
    main()
