"""CPU tests of the host-side model logic with a recording stand-in for the device handle: which C-ABI entry point GPModel
chooses (append vs refit, mean-only vs mean + variance, per-theta handles), what it passes (ARD flag, var_cols, raw vs
normalised arrays) and what it keeps across refits -- the decisions of thesis_code_b200/core_models.py and normalizer.py
that need no GPU to be checked.  The arithmetic itself is covered by the GPU parity tests."""
import numpy as np
import pytest

import thesis_code_b200 as pkg
from thesis_code_b200 import _gpx, core_models


class FakeHandle:
    """Records calls; predictions are zeros of the right shape."""
    instances = []

    def __init__(self, device=0):
        self.device = device
        self.calls = []
        self.N = 0
        self.norm = (False, False)
        self.fail_append = None
        self.closed = False
        FakeHandle.instances.append(self)

    def set_normalizer(self, nx, ny):
        self.norm = (bool(nx), bool(ny)); self.calls.append(("set_normalizer", nx, ny))

    def get_normalizer(self, d, O):
        return np.full(d, 2.0), np.full(d, 4.0), np.full(O, 10.0), np.full(O, 3.0)

    def set_option(self, name, value):
        self.calls.append(("set_option", name, value))

    def fit(self, X, Y, kernel_id, variance, lengthscale, noise, jitter=1e-8, mean_const=0.0, ARD=None):
        self.N = X.shape[0]
        self.calls.append(("fit", X.shape, Y.shape, kernel_id, float(variance), np.array(lengthscale, dtype=float).tolist(), float(noise),
                           float(jitter), float(mean_const), bool(ARD)))
        self.last_X = X

    def append(self, Xn, Yn):
        if self.fail_append is not None:
            raise self.fail_append
        self.N += Xn.shape[0]
        self.calls.append(("append", Xn.shape))

    def predict(self, Xs, want_var=True, include_noise=True, output_dim=1, var_cols=1):
        self.calls.append(("predict", Xs.shape, want_var, include_noise, output_dim, var_cols))
        var = None
        if want_var:
            var = np.ones(Xs.shape[0]) if var_cols == 1 else np.ones((Xs.shape[0], var_cols))
        return np.zeros((Xs.shape[0], output_dim)), var

    def predict_cov(self, Xs, include_noise=True, output_dim=1):
        self.calls.append(("predict_cov", Xs.shape))
        return np.zeros((Xs.shape[0], output_dim)), np.eye(Xs.shape[0])

    def lml(self):
        return -1.0, 0.0, 0.0

    def close(self):
        self.closed = True

    def names(self):
        return [c[0] for c in self.calls]


@pytest.fixture
def fake(monkeypatch):
    FakeHandle.instances = []
    monkeypatch.setattr(_gpx, "Handle", FakeHandle)
    monkeypatch.delenv("GPX_DISTRIBUTED", raising=False)
    return FakeHandle


def _model(**kw):
    cfg = {"kernel": {"name": "GPyMatern52", "kwargs": {"lengthscale": [0.5, 2.0], "ARD": True, "variance": 1.5}}, "noise_prior": 0.1}
    cfg.update(kw)
    return pkg.GPModel.from_config(cfg)


def test_fit_passes_gpy_conventions(fake):
    m = _model(mean_prior=True)
    X = np.random.RandomState(0).rand(20, 2); Y = X[:, :1] + 3.0
    m.init(X, Y)
    (call,) = [c for c in m._handle.calls if c[0] == "fit"]
    assert call[3] == _gpx.KERNEL_IDS["GPyMatern52"] and call[4] == 1.5 and call[5] == [0.5, 2.0] and call[6] == 0.1
    assert call[7] == 1e-8 and call[9] is True                                    # GPy's fixed jitter; the ARD flag travels
    assert call[8] == pytest.approx(float(np.mean(Y)))                            # Constant mean = mean(Y) at creation ...
    m.init(X, Y + 100.0)
    assert [c for c in m._handle.calls if c[0] == "fit"][-1][8] == pytest.approx(float(np.mean(Y)))   # ... and kept on set_XY
    iso = pkg.GPModel.from_config({"kernel": {"name": "GPyRBF"}})
    iso.init(X, Y)
    call = [c for c in iso._handle.calls if c[0] == "fit"][0]
    assert call[5] == [1.0] and call[6] == 1.0 and call[9] is False               # GPy defaults: lengthscale 1, noise variance 1


def test_add_observations_appends_only_when_nothing_but_the_data_changed(fake):
    X = np.random.RandomState(1).rand(30, 2); Y = X[:, :1]
    m = _model()
    m.init(X[:20], Y[:20])
    m.add_observations(X[20:25], Y[20:25])
    assert m._handle.names().count("fit") == 1 and m._handle.names().count("append") == 1
    assert m.X.shape == (25, 2) and m.Y.shape == (25, 1)
    assert ("set_option", "reserve_rows", 256) in m._handle.calls
    # a failed pivot or a handle that cannot append: the reference's route (concatenate + refit)
    m._handle.fail_append = _gpx.NotPositiveDefiniteError(7, "pivot")
    m.add_observations(X[25:27], Y[25:27])
    assert m._handle.names().count("fit") == 2 and m.X.shape == (27, 2)
    m._handle.fail_append = _gpx.AppendUnsupported("sharded")
    m.add_observations(X[27:28], Y[27:28])
    assert m._handle.names().count("fit") == 3 and m.X.shape == (28, 2)
    # the switch, Y_dir, the fused normaliser: always a refit
    m2 = _model(); m2.incremental = False
    m2.init(X[:20], Y[:20]); m2.add_observations(X[20:], Y[20:])
    assert m2._handle.names().count("fit") == 2 and "append" not in m2._handle.names()
    m3 = _model(); m3._normalize_input = True
    m3.init(X[:20], Y[:20]); m3.add_observations(X[20:], Y[20:])
    assert m3._handle.names().count("fit") == 2 and "append" not in m3._handle.names()


def test_predict_entry_points_and_shapes(fake):
    X = np.random.RandomState(2).rand(12, 2); Y = np.hstack([X[:, :1], X[:, 1:] * 2])
    m = _model(); m.init(X, Y)
    mean, var = m.get_statistics(X[:5], full_cov=False)
    assert mean.shape == (1, 5, 2) and var.shape == (1, 5, 2)
    assert m._handle.calls[-1] == ("predict", (5, 2), True, True, 2, 1)           # one shared variance, broadcast on the host
    mu = m.get_mean(X[:3])
    assert mu.shape == (1, 3, 2) and m._handle.calls[-1] == ("predict", (3, 2), False, True, 2, 1)   # no variance pass
    mean, cov = m.get_statistics(X[:4])                                           # the reference's default: full_cov=True
    assert cov.shape == (1, 4, 4, 2) and m._handle.calls[-1][0] == "predict_cov"
    n0 = len(m._handle.calls)
    mean, var = m.get_statistics(np.zeros((0, 2)), full_cov=False)                # nothing to predict: no device call
    assert mean.shape == (1, 0, 2) and var.shape == (1, 0, 2) and len(m._handle.calls) == n0


def test_fused_normalizer_hands_raw_data_down(fake):
    X = np.random.RandomState(3).rand(15, 2) * 50; Y = np.hstack([X[:, :1], X[:, 1:] * 2]) + 7
    nm = pkg.NormalizerModel.from_config({"model": {"name": "GPModel", "kwargs": {"kernel": {"name": "GPyRBF"}, "noise_prior": 0.1}}})
    nm.init(X, Y)
    h = nm.model._handle
    assert h.norm == (True, True) and h.last_X is X                              # raw array, z-scoring happens in gpx_fit
    mean, var = nm.get_statistics(X[:4], full_cov=False)
    assert h.calls[-1] == ("predict", (4, 2), True, True, 2, 2)                   # per-output variance scale: var_cols = O
    assert var.shape == (1, 4, 2)
    mean, cov = nm.get_statistics(X[:4], full_cov=True)
    assert cov.shape == (1, 4, 4, 2, 2)                                           # the reference's trailing-axis quirk, std^2 = 9
    assert np.allclose(cov[0, :, :, 0, 0], 9.0 * np.eye(4))
    np.testing.assert_array_equal(nm.X_normalizer._X_mean, [2.0, 2.0])           # views on the device's statistics
    nm.add_observations(X[:2], Y[:2])
    assert h.names().count("fit") == 2 and "append" not in h.names()              # the statistics changed: full refit


def test_hyper_samples_keep_one_handle_each(fake):
    X = np.random.RandomState(4).rand(10, 1); Y = X.copy()
    m = pkg.GPModel.from_config({"kernel": {"name": "GPyRBF"}})
    m.init(X, Y)
    m.max_cached_thetas = 2
    m._current_thetas = [np.array([1.0, 0.5, 0.1]), np.array([2.0, 0.7, 0.2]), np.array([3.0, 0.9, 0.3])]
    mean, var = m.get_statistics(X[:3], full_cov=False)
    assert mean.shape == (3, 3, 1)
    assert len(m._theta_handles) == 2 and sum(h.closed for h in fake.instances) == 1     # bounded: the oldest was closed
    fits = [h for h in fake.instances if "fit" in h.names()]
    assert len(fits) == 4                                                              # the model's own + one per sample
    m.init(X, Y)                                                                       # new data: the cache is dropped
    assert len(m._theta_handles) == 0


def test_sharded_do_optimize_shares_its_random_draws_first(fake, monkeypatch):
    """do_optimize on a sharded model (gpx_lml_grad works there when every rank holds the whole factor): before anything is
    drawn from the global numpy RNG, the ranks agree on one seed (distributed.sync_rng); a torchrun sweep of independent
    per-rank models (distributed = False) does not meet in that collective."""
    from thesis_code_b200 import distributed as gdist
    monkeypatch.setenv("WORLD_SIZE", "2")
    calls = []
    monkeypatch.setattr(gdist, "sync_rng", lambda handle, dist=None: calls.append(handle) or 1)
    monkeypatch.setattr(core_models.GPModel, "_optimize_hyperparameters", lambda self: calls.append("optimize"))
    monkeypatch.setattr(core_models.GPModel, "_new_handle", lambda self, shard: FakeHandle())
    for shard, expect in ((True, 2), (False, 1)):
        calls.clear()
        m = pkg.GPModel.from_config({"kernel": {"name": "GPyRBF"}, "do_optimize": True})
        m.distributed = shard
        m.init(np.zeros((4, 1)), np.zeros((4, 1)))
        assert len(calls) == expect and calls[-1] == "optimize", calls


def test_latency_inverse_attribute_reaches_the_handle(fake):
    """GPModel.latency_inverse (None = automatic) is handed to every handle the model creates as the gpx option of the same name."""
    X = np.random.RandomState(2).rand(12, 2); Y = X[:, :1]
    for setting, expect in ((None, []), (True, [1]), (False, [0])):
        m = _model()
        m.latency_inverse = setting
        m.init(X, Y)
        got = [c[2] for c in m._handle.calls if c[0] == "set_option" and c[1] == "latency_inverse"]
        assert got == expect, (setting, got)
