"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI (ctypes) and through the
GPModel API, against the numpy/scipy oracle and the committed golden vectors.  Gate (north_star): mean, variance and
log marginal likelihood within 1e-8 (normwise) in fp64."""
import os
import pickle

import numpy as np
import pytest

from conftest import normwise
from oracle import gp_oracle as o
from thesis_code_b200 import GPModel, _gpx, synthetic

pytestmark = pytest.mark.gpu
TOL = 1e-8


def _fit_both(c, h=None, **opts):
    kw = c["model"]["kwargs"]
    m = GPModel.from_config(kw)
    for k, v in opts.items():
        m._get_handle().set_option(k, v)
    m.init(c["X"], c["Y"])
    om = o.OracleGPModel(kernel=kw["kernel"], noise_prior=kw["noise_prior"])
    om.init(c["X"], c["Y"])
    return m, om


@pytest.mark.parametrize("name", ["GPyRBF", "GPyMatern52", "GPyMatern32", "GPyExponential"])
@pytest.mark.parametrize("N,M,d,ARD", [(1, 1, 1, False), (127, 3, 2, True), (128, 129, 1, False), (300, 77, 3, True), (1000, 500, 8, True), (515, 260, 33, True), (260, 131, 64, True), (200, 140, 150, True)])
def test_kernel_build_matches_oracle(gpu_handle, name, N, M, d, ARD):
    rng = np.random.RandomState(N + d)
    X = rng.rand(N, d) * 3
    X2 = rng.rand(M, d) * 3
    ls = (0.5 + rng.rand(d)) if ARD else np.array([0.7])
    K = gpu_handle.kernel_matrix(X, None, _gpx.KERNEL_IDS[name], 1.3, ls, 0.25)
    Ko = np.tril(o.kernel_matrix(name, X, None, 1.3, ls, ARD) + 0.25 * np.eye(N))
    Kc = gpu_handle.kernel_matrix(X, X2, _gpx.KERNEL_IDS[name], 1.3, ls)
    Kco = o.kernel_matrix(name, X, X2, 1.3, ls, ARD)
    # absolute tolerance: the Gram trick loses ~1e-16*|x|^2/l^2 in r^2; Exponential/Matern amplify it near r=0 via sqrt
    atol = 1e-13 if name == "GPyRBF" else 2e-7 if name == "GPyExponential" else 1e-11
    assert np.abs(K - Ko).max() < atol
    assert np.abs(Kc - Kco).max() < atol
    np.testing.assert_array_equal(np.diag(K), 1.3 + 0.25)      # exact diagonal, like GPy's zeroed r


def test_golden_vectors_through_c_abi(gpu_handle, golden_cases):
    for c in golden_cases:
        h = gpu_handle
        h.fit(c["X"], c["Y"], _gpx.KERNEL_IDS[c["kernel"]], c["variance"], c["lengthscale"], c["noise"])
        mean, var = h.predict(c["Xs"], want_var=True, include_noise=True, output_dim=c["Y"].shape[1])
        assert normwise(mean, c["mean"]) < TOL, c["kernel"]
        assert normwise(var, c["var_with_noise"]) < TOL, c["kernel"]
        assert abs(h.lml()[0] - c["lml"]) / abs(c["lml"]) < TOL
        mean_nv, none = h.predict(c["Xs"], want_var=False, output_dim=c["Y"].shape[1])
        assert none is None and np.array_equal(mean_nv, mean)
        mean2, var2 = h.predict(c["Xs"], want_var=True, include_noise=False, output_dim=c["Y"].shape[1])
        np.testing.assert_allclose(var2 + c["noise"], var, rtol=1e-12, atol=1e-14)


def test_full_covariance_matches_golden_and_oracle(gpu_handle, golden_cases):
    """full_cov=True (the reference's default for get_statistics, core_models.py:90-91,244-259)."""
    for c in golden_cases[::3]:
        gpu_handle.fit(c["X"], c["Y"], _gpx.KERNEL_IDS[c["kernel"]], c["variance"], c["lengthscale"], c["noise"])
        mean, cov = gpu_handle.predict_cov(c["Xs"], include_noise=True, output_dim=c["Y"].shape[1])
        assert normwise(mean, c["mean"]) < TOL and normwise(cov, c["cov_with_noise"]) < TOL
        assert np.array_equal(cov, cov.T)
    cfg = synthetic.make_config("C3", N=1500, Ns=300)
    m, om = _fit_both(cfg)
    mean, cov = m.get_statistics(cfg["Xs"])                      # default full_cov=True
    omean, ocov = om.get_statistics(cfg["Xs"], full_cov=True)
    assert mean.shape == (1, 300, 1) and cov.shape == (1, 300, 300, 1)
    assert normwise(mean, omean) < TOL and normwise(cov, ocov) < TOL
    mean2, var = m.get_statistics(cfg["Xs"], full_cov=False)
    np.testing.assert_allclose(np.diagonal(cov[0, :, :, 0]), var[0, :, 0], rtol=1e-10, atol=1e-12)
    assert np.array_equal(mean2, mean)


def test_closed_forms_n1_n2_n3(gpu_handle):
    v, ls, s, j = 1.7, 0.8, 0.05, 1e-8
    X = np.array([[0.3]]); Y = np.array([[1.2]]); Xs = np.array([[0.9], [0.3], [5.0]])
    gpu_handle.fit(X, Y, 0, v, [ls], s)
    mean, var = gpu_handle.predict(Xs)
    k = np.exp(-0.5 * ((Xs[:, 0] - 0.3) / ls) ** 2)
    den = v + s + j
    np.testing.assert_allclose(mean[:, 0], v * k / den * 1.2, rtol=1e-12)
    np.testing.assert_allclose(var, v - (v * k) ** 2 / den + s, rtol=1e-12)
    np.testing.assert_allclose(gpu_handle.lml()[0], -0.5 * np.log(2 * np.pi) - 0.5 * np.log(den) - 0.5 * 1.2 ** 2 / den, rtol=1e-12)
    for N in (2, 3):
        rng = np.random.RandomState(N)
        X = rng.rand(N, 2); Y = rng.randn(N, 1); Xs = rng.rand(5, 2)
        gpu_handle.fit(X, Y, 1, 0.9, [0.5, 1.5], 0.2)
        K = o.kernel_matrix("GPyMatern52", X, None, 0.9, [0.5, 1.5], True) + (0.2 + 1e-8) * np.eye(N)
        Kx = o.kernel_matrix("GPyMatern52", X, Xs, 0.9, [0.5, 1.5], True)
        alpha = np.linalg.solve(K, Y)
        mean, var = gpu_handle.predict(Xs)
        np.testing.assert_allclose(mean, Kx.T @ alpha, rtol=1e-11)
        np.testing.assert_allclose(var, 0.9 + 0.2 - np.einsum("ij,ij->j", Kx, np.linalg.solve(K, Kx)), rtol=1e-11)
        a_gpu, L_gpu = gpu_handle.export(N, 1, want_L=True)
        np.testing.assert_allclose(L_gpu, np.linalg.cholesky(K), rtol=1e-12, atol=1e-15)
        np.testing.assert_allclose(a_gpu, alpha, rtol=1e-11)


@pytest.mark.parametrize("cfg,N,Ns", [("C1", 1000, 1000), ("C2", 3000, 700), ("C3", 2500, 333), ("C4", 4000, 5000), ("C5", 5000, 1000)])
def test_baseline_configs_reduced_size_vs_oracle(cfg, N, Ns):
    c = synthetic.make_config(cfg, N=N, Ns=Ns)
    m, om = _fit_both(c)
    mean, var = m.get_statistics(c["Xs"], full_cov=False)
    omean, ovar = om.get_statistics(c["Xs"], full_cov=False)
    assert mean.shape == (1, Ns, 1) and var.shape == (1, Ns, 1)
    assert normwise(mean, omean) < TOL
    assert normwise(var, ovar) < TOL
    lml, olml = m.get_marginal_log_likelihood(), om.state["lml"]
    assert abs(lml - olml) / abs(olml) < TOL
    assert var.min() > 0                       # mnlp in src/utils.py:65-74 needs positive variances
    alpha, L = m._get_handle().export(N, 1, want_L=True)
    if cfg != "C4":                             # C4 (1-D, ~60 points per lengthscale) is too ill-conditioned for alpha itself
        assert normwise(L, om.state["L"]) < 1e-10
        assert normwise(alpha, om.state["alpha"]) < 1e-7
    e = o.errors(mean[0], var[0], om.get_statistics(c["Xs"], full_cov=False)[0][0], float(c["Y"].mean()))
    assert e["rmse"] < 1e-7


@pytest.mark.parametrize("outer,batch", [(1, 0), (3, 2), (8, 0), (16, 5)])
def test_blocking_options_do_not_change_results(outer, batch):
    c = synthetic.make_config("C5", N=2100, Ns=900)
    m, om = _fit_both(c, outer_tiles=outer, predict_batch_tiles=batch)
    mean, var = m.get_statistics(c["Xs"], full_cov=False)
    omean, ovar = om.get_statistics(c["Xs"], full_cov=False)
    assert normwise(mean, omean) < TOL and normwise(var, ovar) < TOL
    assert abs(m.get_marginal_log_likelihood() - om.state["lml"]) / abs(om.state["lml"]) < TOL


def test_multi_output_mean_prior_and_default_noise():
    rng = np.random.RandomState(7)
    X = rng.rand(700, 3); Y = np.stack([np.sin(4 * X[:, 0]) + 5.0, np.cos(3 * X[:, 1]) - 2.0], 1); Xs = rng.rand(200, 3)
    kern = {"name": "GPyMatern32", "kwargs": {"variance": 2.0, "lengthscale": [0.4, 0.6, 0.9], "ARD": True}}
    m = GPModel.from_config({"kernel": kern, "noise_prior": 0.05, "mean_prior": True})
    m.init(X, Y)
    st = o.fit(X, Y, "GPyMatern32", 2.0, [0.4, 0.6, 0.9], True, 0.05, mean_const=float(np.mean(Y)))
    omean, ovar = o.predict(st, Xs)
    mean, var = m.get_statistics(Xs, full_cov=False)
    assert mean.shape == (1, 200, 2) and var.shape == (1, 200, 2)
    assert normwise(mean[0], omean) < TOL and normwise(var[0, :, 0], ovar[:, 0]) < TOL
    assert np.array_equal(var[0, :, 0], var[0, :, 1])      # GPy shares one variance across outputs
    assert abs(m.get_marginal_log_likelihood() - st["lml"]) / abs(st["lml"]) < TOL
    # noise_prior=None -> GPy's default likelihood variance 1.0 (core_models.py:201-207)
    m1 = GPModel.from_config({"kernel": kern})
    m1.init(X, Y[:, :1])
    st1 = o.fit(X, Y[:, :1], "GPyMatern32", 2.0, [0.4, 0.6, 0.9], True, 1.0)
    assert m1.get_common_hyperparameters()["noise"][0] == 1.0
    assert normwise(m1.get_mean(Xs)[0], o.predict(st1, Xs)[0]) < TOL
    # noise_prior=0 is falsy: like the reference (`if self.noise_prior:`) it is ignored and the default 1.0 is used
    m0 = GPModel.from_config({"kernel": kern, "noise_prior": 0})
    m0.init(X, Y[:, :1])
    assert m0.get_common_hyperparameters()["noise"][0] == 1.0


def test_add_observations_refits_and_pickle_roundtrip(tmp_path):
    c = synthetic.make_config("C2", N=600, Ns=100)
    kw = c["model"]["kwargs"]
    m = GPModel.from_config(kw)
    m.init(c["X"][:500], c["Y"][:500])
    m.add_observations(c["X"][500:], c["Y"][500:])
    assert m.X.shape == (600, 10) and m.Y.shape == (600, 1)
    om = o.OracleGPModel(kernel=kw["kernel"], noise_prior=kw["noise_prior"]); om.init(c["X"], c["Y"])
    mean, var = m.get_statistics(c["Xs"], full_cov=False)
    omean, ovar = om.get_statistics(c["Xs"], full_cov=False)
    assert normwise(mean, omean) < TOL and normwise(var, ovar) < TOL
    xi, yi = m.get_incumbent()
    assert yi == c["Y"].max()
    m.save(str(tmp_path / "model"))
    m2 = GPModel.load(str(tmp_path / "model"))
    assert m2._handle is None
    mean2, var2 = m2.get_statistics(c["Xs"], full_cov=False)      # re-factorises lazily on the GPU
    assert np.array_equal(mean2, mean) and np.array_equal(var2, var)


def test_not_positive_definite_is_reported_and_retried(gpu_handle):
    X = np.zeros((300, 1)); X[:150] = 1.0                    # two distinct points repeated: rank 2
    Y = np.ones((300, 1))
    with pytest.raises(_gpx.NotPositiveDefiniteError) as ei:
        gpu_handle.fit(X, Y, 0, 1.0, [1.0], 1e-18, jitter=1e-300)
    assert ei.value.info > 0
    with pytest.raises(_gpx.GpxError):
        gpu_handle.lml()                                       # state is not "fitted" after a failed factorisation
    m = GPModel.from_config({"kernel": {"name": "GPyRBF"}, "noise_prior": 1e-18})
    m.jitter = 1e-300
    with pytest.warns(UserWarning):
        m.init(X, Y)                                           # jitchol-style retry with growing jitter succeeds
    assert np.isfinite(m.get_marginal_log_likelihood())


def test_bad_arguments_fail_loudly(gpu_handle):
    X = np.random.rand(10, 2); Y = np.random.rand(10, 1)
    with pytest.raises(_gpx.GpxError):
        gpu_handle.fit(X, Y, 7, 1.0, [1.0], 0.1)               # unknown kernel id
    with pytest.raises(_gpx.GpxError):
        gpu_handle.fit(X, Y, 0, 1.0, [1.0, 2.0, 3.0], 0.1)     # wrong lengthscale count
    with pytest.raises(_gpx.GpxError):
        gpu_handle.fit(X, Y, 0, -1.0, [1.0], 0.1)              # negative variance
    gpu_handle.fit(np.random.rand(10, 65), Y, 0, 1.0, [1.0], 0.1)       # any input dimension is accepted (chunked K-build)
    with pytest.raises(_gpx.GpxError):
        gpu_handle.lml_grad(1)                                         # ... but the gradient kernel is limited to d <= 64


def test_repeated_factorisations_are_bit_identical(gpu_handle):
    # regression test for the shared-memory cross-proxy WAR race found in round 1 (csrc/gemm_dmma.cu)
    N = 9000
    c = synthetic.make_config("C5", N=N, Ns=256)
    ls = np.array(c["model"]["kwargs"]["kernel"]["kwargs"]["lengthscale"])
    ref = None
    for W in (1, 4):
        gpu_handle.set_option("outer_tiles", W)
        outs = []
        for rep in range(4):
            gpu_handle.fit(c["X"], c["Y"], 0, 1.0, ls, 1e-2)
            alpha, _ = gpu_handle.export(N, 1)
            mean, var = gpu_handle.predict(c["Xs"])
            outs.append((alpha, mean, var, gpu_handle.lml()[0]))
        for a, mu, v, l in outs[1:]:
            assert np.array_equal(a, outs[0][0]) and np.array_equal(mu, outs[0][1]) and np.array_equal(v, outs[0][2]) and l == outs[0][3]
    gpu_handle.set_option("outer_tiles", 0)   # back to auto


def test_c2_full_size_vs_oracle():
    """BASELINE config C2 at its full size (N=20000, d=10, N*=2500): direct parity with the oracle."""
    c = synthetic.make_config("C2")
    m, om = _fit_both(c)
    mean, var = m.get_statistics(c["Xs"], full_cov=False)
    omean, ovar = om.get_statistics(c["Xs"], full_cov=False)
    assert normwise(mean, omean) < TOL and normwise(var, ovar) < TOL
    assert abs(m.get_marginal_log_likelihood() - om.state["lml"]) / abs(om.state["lml"]) < TOL


@pytest.mark.parametrize("cfg", ["C3"])
def test_full_size_properties(cfg):
    """BASELINE config C3 (N=40000, d=8, Matern-5/2 ARD) at FULL training size: size-independent properties instead of an
    O(N^3) CPU run.  (C4 at N=60000 was tried with this check: its inputs scaled by the lengthscale reach |x/l| ~ 2000, so
    the Gram-trick r^2 rebuilt by numpy on the host differs from the GPU's by ~1e-9 -- the SURVEY 7.3 #2 effect -- and
    with |alpha| ~ 1e3 the host-side residual is 2e-7: that measures the two K-builds against each other, not the solver.
    C4 is covered at reduced size against the oracle and at full size by bench.py.)  (i) Ky alpha = y - m (residual, Ky rebuilt blockwise on the host), (ii)
    predicting at training points with include_noise=False gives mean = y - (noise+jitter) alpha exactly (K alpha = y -
    (noise+jitter) alpha), (iii) variances in (0, sigma_f^2], (iv) LML equals -0.5(N log 2pi + logdet + y^T alpha) with
    the exported alpha."""
    c = synthetic.make_config(cfg, Ns=256)
    kw = c["model"]["kwargs"]; kk = kw["kernel"]["kwargs"]
    N = c["N"]
    m = GPModel.from_config(kw)
    m.init(c["X"], c["Y"])
    h = m._get_handle()
    alpha, _ = h.export(N, 1)
    noise = kw["noise_prior"]
    r = np.empty_like(alpha)
    for s in range(0, N, 4000):
        Kb = o.kernel_matrix(kw["kernel"]["name"], c["X"][s:s + 4000], c["X"], kk["variance"], kk["lengthscale"], kk.get("ARD", False))
        r[s:s + 4000] = Kb @ alpha
    r += (noise + 1e-8) * alpha
    assert np.linalg.norm(r - c["Y"]) / np.linalg.norm(c["Y"]) < 1e-9
    idx = np.arange(0, N, 13)[:3000]
    mean, var = h.predict(c["X"][idx], include_noise=False)
    np.testing.assert_allclose(mean, c["Y"][idx] - (noise + 1e-8) * alpha[idx], rtol=0, atol=1e-8 * max(1.0, np.abs(alpha).max() * noise))
    assert var.min() > 0 and var.max() <= kk["variance"] + 1e-12
    lml, logdet, quad = h.lml()
    np.testing.assert_allclose(quad, float((alpha * c["Y"]).sum()), rtol=1e-10)
    np.testing.assert_allclose(lml, -0.5 * (N * np.log(2 * np.pi) + logdet + quad), rtol=1e-13)


@pytest.mark.parametrize("name", ["GPyRBF", "GPyMatern52", "GPyMatern32", "GPyExponential"])
@pytest.mark.parametrize("ARD", [True, False])
def test_lml_gradient_matches_oracle(gpu_handle, name, ARD):
    """dLML/dtheta on the GPU (gpx_lml_grad: Ky^-1 from the tiled factor + fused dK/dtheta reduction) vs the dense numpy
    oracle (which is itself checked against finite differences in tests/test_oracle.py)."""
    rng = np.random.RandomState(11)
    N, d = 700, 3
    X = rng.rand(N, d)
    Y = np.stack([np.sin(3 * X.sum(1)), np.cos(2 * X[:, 0])], 1) + 0.05 * rng.randn(N, 2)
    ls = np.array([0.5, 0.8, 1.1]) if ARD else np.array([0.7])
    gpu_handle.set_option("outer_tiles", 2)
    gpu_handle.fit(X, Y, _gpx.KERNEL_IDS[name], 1.3, ls, 0.05, mean_const=0.2)
    g = gpu_handle.lml_grad(ls.size)
    st = o.fit(X, Y, name, 1.3, ls, ARD, 0.05, mean_const=0.2)
    go = o.lml_grad(st)
    assert g.shape == go.shape
    assert np.abs(g - go).max() / np.abs(go).max() < 1e-8, (g, go)
    gpu_handle.set_option("outer_tiles", 0)


def test_do_optimize_follows_gpy_recipe():
    """do_optimize=True (reference core_models.py:211-223): randomize + L-BFGS-B on -LML.  The same driver run with the
    oracle's objective from the same random start must land on the same hyper-parameters."""
    from thesis_code_b200 import optimize as gopt
    rng = np.random.RandomState(5)
    N = 500
    X = rng.rand(N, 2)
    Y = np.sin(4 * X[:, :1]) + 0.1 * rng.randn(N, 1)
    kern = {"name": "GPyRBF", "kwargs": {"ARD": True}}
    m = GPModel.from_config({"kernel": kern, "do_optimize": True})
    np.random.seed(123)
    m.init(X, Y)
    th = m._current_thetas[0]                       # [variance, ls1, ls2, noise]
    assert th.shape == (4,) and np.all(th > 0)

    def obj(theta):
        st = o.fit(X, Y, "GPyRBF", theta[0], theta[1:3], True, theta[3])
        g = o.lml_grad(st)
        return st["lml"], np.array([g[0], g[1], g[2], g[3]])
    np.random.seed(123)
    x0 = gopt.randomize(4)
    th_o, lml_o, _ = gopt.optimize_lbfgsb(obj, [1.0, 1.0, 1.0, 1.0], [True] * 4, x0=x0)
    assert abs(m.get_marginal_log_likelihood() - lml_o) / abs(lml_o) < 1e-6
    np.testing.assert_allclose(th[[0, 1, 3]], th_o[[0, 1, 3]], rtol=1e-3)        # ls2 is weakly identified (flat direction)
    # the optimum beats the default hyper-parameters and predictions use it
    m0 = GPModel.from_config({"kernel": kern}); m0.init(X, Y)
    assert m.get_marginal_log_likelihood() > m0.get_marginal_log_likelihood() + 100
    mean, var = m.get_statistics(X[:50], full_cov=False)
    assert np.sqrt(np.mean((mean[0] - Y[:50]) ** 2)) < 0.15
    hp = m.get_common_hyperparameters()
    np.testing.assert_allclose(hp["lengthscale"], th[1:3]); assert hp["noise"][0] == th[3]
    # fixed noise (numeric noise_prior) stays fixed; mean_prior adds a free constant
    m2 = GPModel.from_config({"kernel": {"name": "GPyMatern52"}, "do_optimize": True, "noise_prior": 0.01, "mean_prior": True})
    m2.randomize_before_optimize = False
    m2.init(X, Y + 3.0)
    assert m2.noise_variance == 0.01 and m2._current_thetas[0].shape == (4,)      # [mean, variance, ls, noise]
    mean2, _ = m2.get_statistics(X[:50], full_cov=False)
    assert np.isfinite(m2.mean_const) and np.sqrt(np.mean((mean2[0] - (Y[:50] + 3.0)) ** 2)) < 0.15


def test_normalizer_model_around_gpmodel(tmp_path):
    """The thesis' standard stack NormalizerModel<GPModel> (notebook_header.py:17-26) vs the oracle with hand z-scoring."""
    from thesis_code_b200 import NormalizerModel
    rng = np.random.RandomState(3)
    X = rng.rand(900, 4) * 50 - 10; Y = (np.sin(X[:, :1] / 7) * 20 + 100) + rng.randn(900, 1)
    Xs = rng.rand(120, 4) * 50 - 10
    inner = {"name": "GPModel", "kwargs": {"kernel": {"name": "GPyMatern52", "kwargs": {"lengthscale": 1.5}}, "noise_prior": 0.05}}
    m = NormalizerModel.from_config({"model": inner})
    m.init(X, Y)
    mean, var = m.get_statistics(Xs, full_cov=False)
    Xn = (X - X.mean(0)) / X.std(0); Yn = (Y - Y.mean(0)) / Y.std(0)
    st = o.fit(Xn, Yn, "GPyMatern52", 1.0, 1.5, False, 0.05)
    om, ov = o.predict(st, (Xs - X.mean(0)) / X.std(0))
    assert normwise(mean[0], om * Y.std(0) + Y.mean(0)) < TOL and normwise(var[0], ov * Y.std(0) ** 2) < TOL
    m.save(str(tmp_path / "nm"))
    m2 = NormalizerModel.load(str(tmp_path / "nm"))
    mean2, var2 = m2.get_statistics(Xs, full_cov=False)
    assert np.array_equal(mean2, mean) and np.array_equal(var2, var)


def test_hmc_hyper_posterior_gives_leading_sample_axis():
    """num_mcmc > 0 (reference core_models.py:212-219, 249-263): T retained HMC samples, predictions (T, N*, O), each slice
    equal to the oracle's prediction under that sample's hyper-parameters."""
    rng = np.random.RandomState(2)
    X = rng.rand(300, 1); Y = np.sin(6 * X) + 0.1 * rng.randn(300, 1); Xs = rng.rand(40, 1)
    m = GPModel.from_config({"kernel": {"name": "GPyRBF"}, "do_optimize": True, "num_mcmc": 3, "n_burnin": 4,
                             "subsample_interval": 2, "step_size": 0.03, "leapfrog_steps": 4})
    np.random.seed(7)
    m.init(X, Y)
    assert len(m._current_thetas) == 3 and m.has_mcmc_warmup
    mean, var = m.get_statistics(Xs, full_cov=False)
    assert mean.shape == (3, 40, 1) and var.shape == (3, 40, 1)
    for i, th in enumerate(m._current_thetas):
        st = o.fit(X, Y, "GPyRBF", th[0], th[1:2], False, th[2])
        om, ov = o.predict(st, Xs)
        assert normwise(mean[i], om) < TOL and normwise(var[i], ov) < TOL


def test_randomised_small_cases_vs_oracle(gpu_handle):
    """Seeded random sweep over sizes (incl. ragged N around the 128 tile edge), dimensions, kernels, ARD, noise levels,
    output counts and constant means: fit + predict + LML through the C ABI vs the oracle."""
    rng = np.random.RandomState(2024)
    names = list(_gpx.KERNEL_IDS)
    sizes = [1, 2, 5, 127, 128, 129, 255, 256, 257, 300, 383, 385, 511, 640]
    for case in range(40):
        N = int(sizes[case % len(sizes)] if case < 28 else rng.randint(1, 700))
        d = int(rng.randint(1, 13))
        O = int(rng.choice([1, 1, 1, 2, 3]))
        Ns = int(rng.choice([1, 7, 128, 129, 300]))
        name = names[rng.randint(4)]
        ARD = bool(rng.randint(2))
        ls = rng.uniform(0.3, 2.0, d if ARD else 1)
        variance = float(rng.uniform(0.2, 3.0))
        noise = float(10 ** rng.uniform(-4, 0))
        mean_const = float(rng.choice([0.0, rng.randn()]))
        X = rng.randn(N, d)
        Y = rng.randn(N, O) + mean_const
        Xs = rng.randn(Ns, d)
        gpu_handle.set_option("outer_tiles", int(rng.choice([0, 1, 2, 5])))
        gpu_handle.fit(X, Y, _gpx.KERNEL_IDS[name], variance, ls, noise, mean_const=mean_const)
        mean, var = gpu_handle.predict(Xs, output_dim=O)
        st = o.fit(X, Y, name, variance, ls, ARD, noise, mean_const=mean_const)
        om, ov = o.predict(st, Xs)
        tag = (case, N, d, O, Ns, name, ARD)
        assert normwise(mean, om) < TOL, tag
        assert normwise(var, ov[:, 0]) < TOL, tag
        assert abs(gpu_handle.lml()[0] - st["lml"]) <= TOL * max(1.0, abs(st["lml"])), tag
        alpha, _ = gpu_handle.export(N, O)
        assert normwise(alpha, st["alpha"]) < 1e-6, tag
    gpu_handle.set_option("outer_tiles", 0)
