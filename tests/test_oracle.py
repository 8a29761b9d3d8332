"""CPU tests of the oracle (oracle/gp_oracle.py): golden vectors, closed forms, kernel identities.
The reference has no numerical test of GPModel, so these are the known-answer tests SURVEY.md section 8(c) asks for."""
import numpy as np
import pytest

from conftest import normwise
from oracle import gp_oracle as o

TOL = 1e-8  # north_star: 1e-8 relative (normwise) in fp64


def test_golden_vectors_pin_the_oracle(golden_cases):
    assert len(golden_cases) == 16
    for c in golden_cases:
        st = o.fit(c["X"], c["Y"], c["kernel"], c["variance"], c["lengthscale"], c["ARD"], c["noise"])
        mean, var = o.predict(st, c["Xs"], include_noise=True)
        assert normwise(mean, c["mean"]) < TOL
        assert normwise(var[:, 0], c["var_with_noise"]) < TOL
        assert abs(st["lml"] - c["lml"]) / abs(c["lml"]) < TOL
        mean2, cov = o.predict(st, c["Xs"], full_cov=True, include_noise=True)
        assert normwise(cov, c["cov_with_noise"]) < TOL
        assert normwise(mean2, c["mean"]) < TOL


def test_closed_form_n1():
    # N=1: mean = v k/(v+s+j) y ; var = v - v^2 k^2/(v+s+j) + s ; LML = -0.5 log 2pi - 0.5 log(v+s+j) - 0.5 y^2/(v+s+j)
    v, ls, s, j = 1.7, 0.8, 0.05, 1e-8
    X = np.array([[0.3]]); Y = np.array([[1.2]]); Xs = np.array([[0.9], [0.3], [5.0]])
    st = o.fit(X, Y, "GPyRBF", v, ls, False, s)
    mean, var = o.predict(st, Xs)
    k = np.exp(-0.5 * ((Xs[:, 0] - 0.3) / ls) ** 2)
    den = v + s + j
    np.testing.assert_allclose(mean[:, 0], v * k / den * 1.2, rtol=1e-13)
    np.testing.assert_allclose(var[:, 0], v - (v * k) ** 2 / den + s, rtol=1e-13)
    np.testing.assert_allclose(st["lml"], -0.5 * np.log(2 * np.pi) - 0.5 * np.log(den) - 0.5 * 1.2 ** 2 / den, rtol=1e-13)


def test_closed_form_n2():
    v, ls, s = 1.0, 1.0, 0.1
    X = np.array([[0.0], [1.0]]); Y = np.array([[1.0], [-0.5]])
    st = o.fit(X, Y, "GPyMatern32", v, ls, False, s)
    r = 1.0
    k01 = (1 + np.sqrt(3) * r) * np.exp(-np.sqrt(3) * r)
    a = v + s + 1e-8
    K = np.array([[a, k01], [k01, a]])
    alpha = np.linalg.solve(K, Y)
    np.testing.assert_allclose(st["alpha"], alpha, rtol=1e-12)
    lml = -0.5 * 2 * np.log(2 * np.pi) - 0.5 * np.log(a * a - k01 * k01) - 0.5 * float((Y.T @ alpha)[0, 0])
    np.testing.assert_allclose(st["lml"], lml, rtol=1e-12)


@pytest.mark.parametrize("name", o.KERNEL_NAMES)
def test_kernel_identities(name):
    rng = np.random.RandomState(0)
    X = rng.randn(50, 4)
    ls = np.array([0.5, 1.0, 2.0, 0.7])
    K = o.kernel_matrix(name, X, None, 2.5, ls, True)
    assert np.array_equal(K, K.T) or np.allclose(K, K.T, rtol=0, atol=1e-15)
    np.testing.assert_array_equal(np.diag(K), 2.5)          # exact sigma_f^2 diagonal (r forced to 0)
    assert np.linalg.eigvalsh(K + 1e-10 * np.eye(50)).min() > -1e-10
    # ARD == isotropic(1) on rescaled inputs
    K2 = o.kernel_matrix(name, X / ls, None, 2.5, 1.0, False)
    np.testing.assert_allclose(K, K2, rtol=1e-12, atol=1e-15)
    # cross kernel agrees with the symmetric one off the diagonal
    Kc = o.kernel_matrix(name, X, X[:7], 2.5, ls, True)
    off = ~np.eye(50, 7, dtype=bool)
    np.testing.assert_allclose(Kc[off], K[:, :7][off], rtol=1e-10, atol=1e-14)


def test_kernel_smoothness_ordering():
    # at a fixed moderate distance the smoother kernel is larger: RBF > Matern52 > Matern32 > Exponential
    r = np.array([[0.0], [0.7]])
    vals = [o.kernel_matrix(n, r, None, 1.0, 1.0, False)[0, 1] for n in ("GPyRBF", "GPyMatern52", "GPyMatern32", "GPyExponential")]
    assert vals[0] > vals[1] > vals[2] > vals[3]


def test_woodbury_form_matches_trsm_form():
    rng = np.random.RandomState(3)
    X = rng.rand(200, 3); Y = rng.randn(200, 1); Xs = rng.rand(40, 3)
    st = o.fit(X, Y, "GPyMatern52", 1.3, [0.3, 0.5, 0.9], True, 1e-2)
    m1, v1 = o.predict(st, Xs)
    m2, v2 = o.predict_woodbury(st, Xs)
    assert normwise(m1, m2) < 1e-12 and normwise(v1, v2) < 1e-9


def test_interpolation_at_small_noise_and_mean_const():
    rng = np.random.RandomState(4)
    X = np.sort(rng.rand(30, 1), axis=0); Y = np.sin(6 * X) + 3.0
    st = o.fit(X, Y, "GPyRBF", 1.0, 0.2, False, 1e-10, mean_const=float(Y.mean()))
    mean, var = o.predict(st, X, include_noise=False)
    assert np.abs(mean - Y).max() < 1e-4
    assert var.max() < 1e-4 and var.min() > -1e-8


def test_jitchol_retries_then_raises():
    A = np.ones((4, 4))          # rank one: dpotrf fails, a jitter of mean(diag)*1e-6 fixes it
    with pytest.warns(UserWarning):
        L = o.jitchol(A.copy())
    assert np.isfinite(L).all()
    with pytest.raises(np.linalg.LinAlgError):
        o.jitchol(-np.eye(3))


def test_oracle_model_api_shapes():
    rng = np.random.RandomState(5)
    X = rng.rand(64, 2); Y = rng.randn(64, 2); Xs = rng.rand(9, 2)
    m = o.OracleGPModel.from_config({"kernel": {"name": "GPyRBF", "kwargs": {"lengthscale": 0.6, "ARD": True}}, "noise_prior": 0.01})
    m.init(X, Y)
    mean, var = m.get_statistics(Xs, full_cov=False)
    assert mean.shape == (1, 9, 2) and var.shape == (1, 9, 2)
    mean, cov = m.get_statistics(Xs, full_cov=True)
    assert cov.shape == (1, 9, 9, 1)
    e = o.errors(mean[0], np.diagonal(cov[0, :, :, 0])[:, None], rng.randn(9, 2), Y.mean())
    assert set(e) == {"mae", "max_err", "rmse", "mnlp", "nmse"}


@pytest.mark.parametrize("name", o.KERNEL_NAMES)
def test_oracle_lml_gradient_vs_finite_differences_and_sklearn(name):
    rng = np.random.RandomState(0)
    N, d = 50, 3
    X = rng.rand(N, d); Y = np.sin(3 * X.sum(1))[:, None] + 0.05 * rng.randn(N, 1)
    ls = np.array([0.5, 0.8, 1.1])
    st = o.fit(X, Y, name, 1.3, ls, True, 0.05, mean_const=0.2)
    g = o.lml_grad(st)
    eps = 1e-6

    def lml(v, l, s, m):
        return o.fit(X, Y, name, v, l, True, s, mean_const=m)["lml"]
    fd = [(lml(1.3 + eps, ls, 0.05, 0.2) - lml(1.3 - eps, ls, 0.05, 0.2)) / (2 * eps)]
    for q in range(3):
        e = np.zeros(3); e[q] = eps
        fd.append((lml(1.3, ls + e, 0.05, 0.2) - lml(1.3, ls - e, 0.05, 0.2)) / (2 * eps))
    fd.append((lml(1.3, ls, 0.05 + eps, 0.2) - lml(1.3, ls, 0.05 - eps, 0.2)) / (2 * eps))
    fd.append((lml(1.3, ls, 0.05, 0.2 + eps) - lml(1.3, ls, 0.05, 0.2 - eps)) / (2 * eps))
    assert np.abs(g - np.array(fd)).max() / np.abs(g).max() < 1e-7
    # independent check: sklearn's analytic gradient w.r.t. log(variance), log(lengthscales)
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern
    nu = {"GPyMatern52": 2.5, "GPyMatern32": 1.5, "GPyExponential": 0.5}
    base = RBF(ls) if name == "GPyRBF" else Matern(ls, nu=nu[name])
    gp = GaussianProcessRegressor(kernel=ConstantKernel(1.3) * base, alpha=0.05 + 1e-8, optimizer=None).fit(X, Y - 0.2)
    _, gsk = gp.log_marginal_likelihood(gp.kernel_.theta, eval_gradient=True)
    np.testing.assert_allclose(gsk, np.concatenate([[g[0] * 1.3], g[1:4] * ls]), rtol=1e-7)


def test_optimizer_driver_reaches_sklearn_optimum():
    from thesis_code_b200 import optimize as gopt
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, WhiteKernel
    rng = np.random.RandomState(0); X = rng.rand(80, 2); Y = np.sin(4 * X[:, :1]) + 0.1 * rng.randn(80, 1)

    def obj(th):
        st = o.fit(X, Y, "GPyRBF", th[0], th[1:3], True, th[3]); g = o.lml_grad(st)
        return st["lml"], g[:4]
    th, lml, info = gopt.optimize_lbfgsb(obj, [1.0, 1.0, 1.0, 1.0], [True] * 4)
    gp = GaussianProcessRegressor(kernel=ConstantKernel(1.0, (1e-5, 1e5)) * RBF([1.0, 1.0], (1e-5, 1e5)) + WhiteKernel(1.0, (1e-8, 1e5)),
                                  alpha=1e-8, n_restarts_optimizer=0).fit(X, Y)
    assert abs(lml - gp.log_marginal_likelihood_value_) < 1e-5
    assert info["warnflag"] == 0
    # Logexp transform round trip and gradient factor (paramz conventions)
    x = np.array([-30.0, -1.0, 0.0, 2.0, 40.0])
    f = gopt.Logexp.f(x)
    np.testing.assert_allclose(gopt.Logexp.finv(f)[1:], x[1:], rtol=1e-10)
    np.testing.assert_allclose(gopt.Logexp.gradfactor(f, np.ones(5))[:4], 1 / (1 + np.exp(-x[:4])), rtol=1e-9)


def test_hmc_sampler_recipe():
    """GPy's HMC recipe (optimize.hmc_sample) on the oracle objective: reproducible with the global numpy seed, stays in
    the positive orthant, moves towards high marginal likelihood."""
    from thesis_code_b200 import optimize as gopt
    rng = np.random.RandomState(0); X = rng.rand(60, 1); Y = np.sin(6 * X) + 0.1 * rng.randn(60, 1)

    def obj(th):
        st = o.fit(X, Y, "GPyRBF", th[0], th[1:2], False, th[2]); g = o.lml_grad(st)
        return st["lml"], g[:3]
    np.random.seed(0)
    s1 = gopt.hmc_sample(obj, [1., 1., 1.], [True] * 3, num_samples=24, hmc_iters=5, stepsize=0.05)
    np.random.seed(0)
    s2 = gopt.hmc_sample(obj, [1., 1., 1.], [True] * 3, num_samples=24, hmc_iters=5, stepsize=0.05)
    assert s1.shape == (24, 3) and np.array_equal(s1, s2) and np.all(s1 > 0)
    assert obj(s1[-1])[0] > obj(s1[0])[0] + 20
    assert len(np.unique(s1[:, 0])) > 5                                            # proposals are being accepted


def test_oracle_vs_sklearn_random_cases():
    """Beyond the committed fixtures: 24 seeded random problems, oracle vs scikit-learn run live (mean, variance, LML)."""
    from sklearn.gaussian_process import GaussianProcessRegressor
    from sklearn.gaussian_process.kernels import RBF, ConstantKernel, Matern
    nu = {"GPyMatern52": 2.5, "GPyMatern32": 1.5, "GPyExponential": 0.5}
    rng = np.random.RandomState(99)
    for case in range(24):
        name = o.KERNEL_NAMES[case % 4]
        N, d = int(rng.randint(2, 120)), int(rng.randint(1, 6))
        ARD = bool(rng.randint(2))
        ls = rng.uniform(0.3, 2.0, d if ARD else 1)
        v, noise = float(rng.uniform(0.3, 3.0)), float(10 ** rng.uniform(-3, 0))
        X, Xs = rng.randn(N, d), rng.randn(17, d)
        Y = np.sin(X.sum(1, keepdims=True)) + 0.1 * rng.randn(N, 1)
        st = o.fit(X, Y, name, v, ls, ARD, noise)
        mean, var = o.predict(st, Xs, include_noise=False)
        lsk = ls if ARD else float(ls[0])
        base = RBF(lsk) if name == "GPyRBF" else Matern(lsk, nu=nu[name])
        gp = GaussianProcessRegressor(kernel=ConstantKernel(v) * base, alpha=noise + 1e-8, optimizer=None).fit(X, Y)
        ms, ss = gp.predict(Xs, return_std=True)
        assert normwise(mean[:, 0], ms.ravel()) < TOL, (case, name)
        assert np.abs(var[:, 0] - ss ** 2).max() < TOL * max(1.0, v), (case, name)
        lml = gp.log_marginal_likelihood(gp.kernel_.theta)
        assert abs(st["lml"] - lml) <= TOL * max(1.0, abs(lml)), (case, name)


# ---- the oracle pinned to REFERENCE CODE THAT RAN (tests/golden/reference_pinned.npz) -------------------------------------
def _vanilla_as_gpy(c):
    """GPVanillaModel(l_v, s) == RBF(variance 1, lengthscale l_v / sqrt 2), noise s, jitter 0 (make_golden_reference.py)."""
    return dict(kernel="GPyRBF", variance=1.0, lengthscale=float(c["lengthscale_v"]) / np.sqrt(2.0), ARD=False,
                noise=float(c["noise"]), jitter=0.0)


@pytest.mark.parametrize("tag", ["v0", "v1", "v2"])
def test_oracle_matches_reference_gpvanilla_fixtures(reference_pinned, tag):
    """src/models/deprecated_models.py:8-48 executed unmodified -> fixtures; the oracle must reproduce them."""
    c = reference_pinned[tag]
    st = o.fit(c["X"], c["Y"], **_vanilla_as_gpy(c))
    mean, var = o.predict(st, c["Xs"], include_noise=True)
    assert normwise(mean, c["mu"]) < TOL
    assert normwise(var, c["var"]) < TOL
    mean5, cov5 = o.predict(st, c["Xs"][:5], full_cov=True, include_noise=False)
    assert normwise(cov5 + float(c["noise"]), c["cov5"]) < TOL        # GPVanillaModel adds the noise to EVERY entry (:42)


@pytest.mark.parametrize("tag", ["n0", "n1", "n2"])
def test_oracle_with_hand_zscore_matches_reference_normalizer_fixtures(reference_pinned, tag):
    """reference NormalizerModel<GPVanillaModel> (normalizer.py:51-114) executed unmodified -> fixtures."""
    c = reference_pinned[tag]
    nx, ny = bool(c["normalize_input"]), bool(c["normalize_output"])
    X, Y, Xs = c["X"], c["Y"], c["Xs"]
    xm, xs = (X.mean(0), X.std(0)) if nx else (0.0, 1.0)
    ym, ys = (Y.mean(0), Y.std(0)) if ny else (0.0, 1.0)
    if nx:
        np.testing.assert_array_equal(xm, c["x_mean"]); np.testing.assert_array_equal(xs, c["x_std"])
    st = o.fit((X - xm) / xs, (Y - ym) / ys, **_vanilla_as_gpy(c))
    mean, var = o.predict(st, (Xs - xm) / xs, include_noise=True)
    assert normwise(mean * ys + ym, c["mu"]) < TOL
    assert normwise(var * ys ** 2, c["var"]) < TOL


def test_reference_code_runs_live_when_present():
    """In the build container the reference's own files are importable through oracle/ref_loader.py: run GPVanillaModel
    and NormalizerModel live on fresh data and compare with the oracle (skipped on the GPU box: no /root/reference)."""
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference is not present here")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ref = ref_loader.load()
    rng = np.random.RandomState(77)
    X = rng.rand(300, 2) * 5 + 2; Y = np.sin(X[:, :1]) * 3 + 10 + 0.1 * rng.randn(300, 1); Xs = rng.rand(40, 2) * 5 + 2
    m = ref.NormalizerModel(model=ref.GPVanillaModel(lengthscale=0.7, noise=1e-2))
    m.init(X, Y)
    mu, var = m.get_statistics(Xs, full_cov=False)
    st = o.fit((X - X.mean(0)) / X.std(0), (Y - Y.mean(0)) / Y.std(0), "GPyRBF", 1.0, 0.7 / np.sqrt(2), False, 1e-2, jitter=0.0)
    om, ov = o.predict(st, (Xs - X.mean(0)) / X.std(0))
    assert normwise(om * Y.std(0) + Y.mean(0), mu) < TOL and normwise(ov * Y.std(0) ** 2, var) < TOL


def test_large_golden_fixture_is_consistent(golden_large):
    """tests/golden/exact_gp_golden_large.npz (C3 / C4 at FULL size, make_golden_large.py): internal identities."""
    from thesis_code_b200 import synthetic
    for name, N in (("C3", 40000), ("C4", 60000)):
        g = golden_large[name]
        assert int(g["N"]) == N and g["mean"].shape == (512, 1) and g["var_with_noise"].shape == (512,)
        np.testing.assert_allclose(float(g["lml"]), 0.5 * (-N * o.LOG_2_PI - float(g["logdet"]) - float(g["quad"])), rtol=1e-14)
        assert g["var_with_noise"].min() > 0
    # C4's test draw does not depend on N: the stored test points are the first 512 of the config's own test set
    np.testing.assert_array_equal(golden_large["C4"]["Xs"], synthetic.make_config("C4", N=256, Ns=100000)["Xs"][:512])


def test_explicit_inverse_identities_behind_the_latency_path():
    """The algebra the latency path's explicit inverse Y = L^-T relies on (csrc/gpx_update.cu: finv_gemv_kernel, append_vec,
    refresh_factor_inverse), checked in numpy against the oracle's own factorisation:
      * var = prior - |Y^T k*|^2                                   (one triangular GEMV instead of a forward substitution)
      * appending k rows:  Y_new = [[Y, -U C^-T], [0, C^-T]],  Z = Y^T K(X, X_new),  C = chol(K_nn + s^2 I - Z^T Z),  U = Y Z,
        alpha_new = Y_new z_new with z_new = [z, C^-1 (r_new - Z^T z)]  (no backward substitution)
      * the leading tile columns of Y do not change when rows are appended (only the new columns are recomputed)."""
    import scipy.linalg as sla
    rng = np.random.RandomState(3)
    N, k, d = 300, 5, 3
    X = rng.rand(N + k, d); Y = np.sin(3 * X.sum(1, keepdims=True)) + 0.05 * rng.randn(N + k, 1)
    ls = np.array([0.4, 0.7, 1.1])
    kw = dict(kernel="GPyMatern52", variance=1.3, lengthscale=ls, ARD=True, noise=0.02, mean_const=0.1)
    st0 = o.fit(X[:N], Y[:N], **kw)
    st1 = o.fit(X, Y, **kw)
    L0, L1 = st0["L"], st1["L"]
    Yinv = sla.solve_triangular(L0, np.eye(N), lower=True).T                    # Y = L^-T (upper triangular)
    Xs = rng.rand(4, d)
    Ks = o.kernel_matrix("GPyMatern52", X[:N], Xs, 1.3, ls, True)
    mean, var = o.predict(st0, Xs)
    v = Yinv.T @ Ks
    assert np.allclose(1.3 + 0.02 - (v * v).sum(0), var[:, 0], rtol=1e-10, atol=0)
    # append
    Kn = o.kernel_matrix("GPyMatern52", X[:N], X[N:], 1.3, ls, True)
    Z = Yinv.T @ Kn
    S = o.kernel_matrix("GPyMatern52", X[N:], None, 1.3, ls, True) + (0.02 + st0["jitter"]) * np.eye(k) - Z.T @ Z
    C = np.linalg.cholesky(S)
    Cinv = sla.solve_triangular(C, np.eye(k), lower=True)
    assert np.allclose(L1[N:, :N], Z.T, rtol=1e-9, atol=1e-12) and np.allclose(L1[N:, N:], C, rtol=1e-9, atol=1e-12)
    U = Yinv @ Z
    Ynew = np.block([[Yinv, -U @ Cinv.T], [np.zeros((k, N)), Cinv.T]])
    Yref = sla.solve_triangular(L1, np.eye(N + k), lower=True).T
    assert np.allclose(Ynew, Yref, rtol=1e-8, atol=1e-10)
    assert np.array_equal(np.triu(Yref[:N, :N]) != 0, np.triu(Yinv) != 0) and np.allclose(Yref[:N, :N], Yinv, rtol=1e-9, atol=1e-12)
    r = Y - 0.1
    z = sla.solve_triangular(L0, r[:N], lower=True)
    znew = np.vstack([z, Cinv @ (r[N:] - Z.T @ z)])
    assert np.allclose(Ynew @ znew, st1["alpha"], rtol=1e-8, atol=1e-10)
