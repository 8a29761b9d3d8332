"""CPU tests: the C-ABI library loads and exports every symbol include/gpx.h declares, the host-side model API
mirrors the reference's GPModel contract, and nothing silently falls back to the CPU."""
import ctypes
import os
import pickle
import re

import numpy as np
import pytest

import thesis_code_b200 as pkg
from thesis_code_b200 import _gpx, kernels, synthetic
from thesis_code_b200.config_helpers import LazyConstructor, construct_from_module, lazy_construct_from_module

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gpx.h")).read()
    text = re.sub(r"#ifdef GPX_ENABLE_DEBUG_API.*?#endif", "", text, flags=re.S)    # development aids: `make DEBUG_API=1` only
    return sorted(set(re.findall(r"GPX_API[^;(]*?\b(gpx_\w+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound():
    syms = _declared_symbols()
    assert len(syms) >= 15
    lib = ctypes.CDLL(_gpx.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libgpx.so does not export " + s
    assert sorted(_gpx.SIGNATURES) == syms, "ctypes prototypes out of sync with include/gpx.h"
    assert _gpx.load_library().gpx_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    with pytest.raises(_gpx.GpxError):
        _gpx.Handle(0)
    m = pkg.GPModel.from_config({"kernel": {"name": "GPyRBF"}, "noise_prior": 0.1})
    with pytest.raises(_gpx.GpxError):
        m.init(np.zeros((4, 1)), np.zeros((4, 1)))


def test_product_path_does_not_import_oracle():
    for root, _, files in os.walk(os.path.join(ROOT, "thesis_code_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(root, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
                assert "gp_oracle" not in src, f


def test_config_helpers_semantics():
    lazy = lazy_construct_from_module(kernels, {"name": "GPyMatern52", "kwargs": {"lengthscale": 0.6, "ARD": True}})
    assert isinstance(lazy, LazyConstructor)
    k = lazy(5)
    assert k.name == "GPyMatern52" and k.input_dim == 5 and k.ARD
    np.testing.assert_array_equal(k.lengthscale, np.full(5, 0.6))   # scalar broadcast under ARD
    np.testing.assert_array_equal(k.variance, [1.0])
    assert lazy_construct_from_module(kernels, None) is None
    k2 = lazy(3, variance=2.0)                                        # call-site kwargs win
    assert k2.variance[0] == 2.0
    m = construct_from_module(pkg, {"name": "GPModel", "kwargs": {"kernel": {"name": "GPyRBF"}, "noise_prior": 0.5}})
    assert isinstance(m, pkg.GPModel) and m.noise_prior == 0.5 and repr(m) == "ExactGP"
    assert construct_from_module(pkg, None) is None


def test_kernel_spec_validation():
    assert kernels.GPyRBF(3).lengthscale.shape == (1,)
    assert kernels.GPyRBF(3, ARD=True).lengthscale.shape == (3,)
    with pytest.raises(AssertionError):
        kernels.GPyRBF(3, lengthscale=[1.0, 2.0])            # non-ARD takes one lengthscale
    with pytest.raises(AssertionError):
        kernels.GPyRBF(3, lengthscale=[1.0, 2.0], ARD=True)  # wrong count
    assert {k: v for k, v in _gpx.KERNEL_IDS.items()} == {"GPyRBF": 0, "GPyMatern52": 1, "GPyMatern32": 2, "GPyExponential": 3}


def test_gpmodel_contract_without_gpu():
    m = pkg.GPModel.from_config({"kernel": {"name": "GPyRBF", "kwargs": {"lengthscale": 1.0}}, "noise_prior": 0})
    # ctor kwargs / defaults of the reference (core_models.py:139-148)
    for attr, val in dict(do_optimize=False, num_mcmc=0, n_burnin=100, subsample_interval=10, step_size=1e-1,
                          leapfrog_steps=20, mean_prior=None).items():
        assert getattr(m, attr) == val
    assert m.X is None and m.Y is None
    with pytest.raises(AssertionError):
        m.add_observations(np.zeros((1, 1)), np.zeros((1, 1)))        # "Call init first"
    with pytest.raises(AssertionError):
        m._fit(np.zeros((3, 1)), np.zeros((2, 1)))                     # X/Y size mismatch
    mo = pkg.GPModel.from_config({"kernel": {"name": "GPyRBF"}, "do_optimize": True, "noise_prior": object()})
    with pytest.raises(NotImplementedError):
        mo._fit(np.zeros((3, 1)), np.zeros((3, 1)))              # GPy prior objects on the noise are not supported
    # pickling never carries device state
    m2 = pickle.loads(pickle.dumps(m))
    assert repr(m2) == "ExactGP" and m2._handle is None
    assert issubclass(pkg.DerivativeGPModel, pkg.GPModel)


def test_synthetic_configs_shapes_and_determinism():
    sizes = {"C1": (1000, 1, 1000), "C2": (20000, 10, 2500), "C3": (40000, 8, 4000), "C4": (60000, 1, 100000), "C5": (200000, 8, 2500)}
    for name in ["C1", "C2"]:
        c = synthetic.make_config(name)
        assert (c["N"], c["d"], c["Ns"]) == sizes[name]
        assert c["X"].dtype == np.float64 and c["X"].flags.c_contiguous and c["Y"].shape == (c["N"], 1)
        c2 = synthetic.make_config(name)
        assert np.array_equal(c["X"], c2["X"]) and np.array_equal(c["Y"], c2["Y"]) and np.array_equal(c["Xs"], c2["Xs"])
    for name in ["C3", "C4", "C5"]:
        c = synthetic.make_config(name, N=500, Ns=64)
        assert c["X"].shape == (500, sizes[name][1]) and c["Xs"].shape == (64, sizes[name][1])
        assert c["model"]["name"] == "GPModel" and c["model"]["kwargs"]["do_optimize"] is False
    c = synthetic.make_config("C1")
    np.testing.assert_allclose(c["Y"], np.sin(c["X"]) / c["X"])          # Sinc (smooth.py:14-21)
    assert c["X"].min() >= -20 and c["X"].max() <= 20


def test_normalizer_model_contract_without_gpu(tmp_path):
    """NormalizerModel (reference src/models/normalizer.py:51-158): constructor / config contract, flags handed to the
    wrapped GPModel (the z-scores themselves run inside gpx_fit / gpx_predict: GPU tests), pickle layout."""
    inner = {"name": "GPModel", "kwargs": {"kernel": {"name": "GPyRBF"}, "noise_prior": 0.1}}
    nm = pkg.NormalizerModel.from_config({"model": inner, "normalize_output": False})
    assert isinstance(nm.model, pkg.GPModel) and nm._normalize_input and not nm._normalize_output
    assert nm.model._normalize_input and not nm.model._normalize_output          # pushed down: the fit will z-score X only
    nm2 = pkg.NormalizerModel.from_config({"model": inner})
    assert nm2.model._normalize_input and nm2.model._normalize_output and nm2.X is None
    with pytest.raises(AssertionError):
        nm2.add_observations(np.zeros((1, 1)), np.zeros((1, 1)))                 # "Call init first"
    with pytest.raises(TypeError):
        pkg.NormalizerModel(object())                                            # nothing to fuse into
    nm2.save(str(tmp_path / "nm"))                                               # model.pickle + model/model.pickle
    assert os.path.isfile(tmp_path / "nm" / "model.pickle") and os.path.isfile(tmp_path / "nm" / "model" / "model.pickle")
    nm3 = pkg.NormalizerModel.load(str(tmp_path / "nm"))
    assert isinstance(nm3.model, pkg.GPModel) and nm3.model._normalize_output and nm3._X is None
    n = pkg.Normalizer(np.array([1.0, 2.0]), np.array([2.0, 4.0]))               # reference helper semantics (:9-33)
    np.testing.assert_allclose(n.normalize(np.array([[3.0, 6.0]])), [[1.0, 1.0]])
    np.testing.assert_allclose(n.denormalize(np.array([[1.0, 1.0]])), [[3.0, 6.0]])
    np.testing.assert_allclose(n.denormalize_variance(np.array([[1.0, 1.0]])), [[4.0, 16.0]])
    assert n.denormalize_covariance(np.ones((1, 5, 5, 1))).shape == (1, 5, 5, 1, 2)   # the reference's trailing-axis quirk


def test_priors_restate_gpy_formulas():
    """thesis_code_b200.priors (GPy 1.9.6 priors.py restated) against scipy.stats, gradients against finite differences,
    and the Logexp log-Jacobian paramz adds for a prior on a transformed parameter."""
    import scipy.stats as st
    from thesis_code_b200 import optimize, priors
    x = np.array([0.3, 1.7, 4.0])
    np.testing.assert_allclose(priors.Gamma(2.5, 3.0).lnpdf(x), st.gamma.logpdf(x, 2.5, scale=1 / 3.0), rtol=1e-13)
    np.testing.assert_allclose(priors.LogGaussian(0.2, 0.7).lnpdf(x), st.lognorm.logpdf(x, 0.7, scale=np.exp(0.2)), rtol=1e-13)
    np.testing.assert_allclose(priors.Gaussian(1.0, 2.0).lnpdf(x), st.norm.logpdf(x, 1.0, 2.0), rtol=1e-13)
    g = priors.Gamma.from_EV(2.0, 0.5)
    assert abs(g.a / g.b - 2.0) < 1e-14 and abs(g.a / g.b ** 2 - 0.5) < 1e-14
    for p in (priors.Gamma(2.5, 3.0), priors.LogGaussian(0.2, 0.7), priors.Gaussian(1.0, 2.0)):
        fd = (p.lnpdf(x + 1e-6) - p.lnpdf(x - 1e-6)) / 2e-6
        np.testing.assert_allclose(p.lnpdf_grad(x), fd, rtol=1e-6)
        assert priors.is_prior(p) and np.all(np.isfinite(p.rvs(5)))
    assert not priors.is_prior(object()) and not priors.is_prior(0.1)
    f = np.array([0.5, 2.0, 40.0])
    fd = (optimize.Logexp.log_jacobian(f + 1e-6) - optimize.Logexp.log_jacobian(f - 1e-6)) / 2e-6
    np.testing.assert_allclose(optimize.Logexp.log_jacobian_grad(f)[:2], fd[:2], rtol=1e-6)
    m = pkg.GPModel.from_config({"kernel": {"name": "GPyRBF"}, "do_optimize": True, "noise_prior": priors.Gamma(2.0, 4.0)})
    assert m._noise_prior_object() is m.noise_prior            # accepted like Gaussian_noise.variance.set_prior(...)


def test_sharding_is_opt_in(monkeypatch):
    m = pkg.GPModel.from_config({"kernel": {"name": "GPyRBF"}, "noise_prior": 0.1})
    monkeypatch.setenv("WORLD_SIZE", "4")
    monkeypatch.delenv("GPX_DISTRIBUTED", raising=False)
    assert not m._wants_sharding()                      # a torchrun sweep of independent models must not meet in a collective
    monkeypatch.setenv("GPX_DISTRIBUTED", "1")
    assert m._wants_sharding()
    m.distributed = False
    assert not m._wants_sharding()
    from thesis_code_b200 import distributed as gd
    a = gd.problem_fingerprint(np.zeros((5, 2)), np.zeros((5, 1)), np.array([1.0, 2.0]))
    b = gd.problem_fingerprint(np.zeros((5, 2)), np.zeros((5, 1)), np.array([1.0, 2.5]))
    assert a.shape == (6,) and np.array_equal(a[:5], b[:5]) and a[5] != b[5]


def test_committed_bench_lines_follow_the_contract():
    """The bench lines kept under profiles/ (produced by bench.py on B200s) carry every key the driver's contract names."""
    import glob
    import json
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r0[12]_bench_c*gpu.json")))
    assert len(files) >= 6
    for f in files:
        line = json.loads(open(f).read().strip().splitlines()[-1])
        for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                    "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
            assert key in line, (f, key)
        assert line["dtype"] == "f64" and line["data"] == "synthetic" and line["vs_baseline"] is None
        assert "workload" in line["config"] and "model" not in line["config"]
        r = line["roofline"]
        assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
        assert 0.5 < r["frac"] < 1.05                                  # a DMMA-bound kernel close to, not above, the measured peak
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
        assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["gpu_launches"] > 0
        assert {"value", "unit", "cores", "kind", "sample"} <= set(line["cpu_baseline"]) and line["cpu_baseline"]["kind"] == "port"
        assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if os.path.basename(f).startswith("r02"):
            # round 2: every bench line carries a correctness check of the state it timed
            c = line["check"]
            assert c["solve_residual"] < 1e-9 and c["factor_residual"] < 1e-12 and c["var_min"] > 0
            assert line["steps"] <= line["steps_requested"] and line["wall_s"] < line["budget_s"]
            if line["n_gpus"] > 1:
                assert c["lml_spread_over_ranks"] == 0.0
            else:
                g = c["golden_parity"]
                assert max(g["mean_normwise"], g["var_normwise"], g["lml_rel"]) < g["gate"] == 1e-8


def test_reference_arm_of_bench_runs_on_cpu_and_only_on_rank_0():
    """`bench.py --impl reference` (the driver's CPU arm): rank 0 prints ONE JSON line with the contract's keys on the CUDA arm's
    config / metric / unit; the other ranks of a torchrun launch exit 0 without work."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1",
           "--N", "600", "--Ns", "100"]
    out0 = subprocess.run(cmd, env=dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0"), capture_output=True, text=True, timeout=300)
    assert out0.returncode == 0, out0.stderr[-2000:]
    lines = [l for l in out0.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "TFLOP/s" and d["higher_is_better"] is True and d["n_gpus"] == 2
    assert d["steps"] == 2 and d["warmup"] == 1 and d["steps_requested"] == 2
    assert d["config"]["N"] == 200000 and d["config"]["workload"].startswith("C5")        # the CUDA arm's config, not the sample's
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["sample_N"] == 600 and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    out1 = subprocess.run(cmd, env=dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1"), capture_output=True, text=True, timeout=300)
    assert out1.returncode == 0 and out1.stdout.strip() == ""
