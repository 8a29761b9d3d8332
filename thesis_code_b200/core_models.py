"""Model layer: the reference's ``BaseModel`` / ``ProbModel`` / ``GPModel`` API, backed by the sm_100a CUDA library.

Drop-in for ``src/models/core_models.py`` of tmpethick/thesis-code on the exact-GP path:

* ``SaveMixin`` (:17-35), ``BaseModel`` (:38-123), ``ProbModel`` (:126-128) keep their names, arguments and behaviour;
* ``GPModel`` (:138-271) keeps its constructor kwargs, ``process_config``/``from_config`` (config dict with a lazy
  kernel constructor), ``init(X, Y, Y_dir=None)``, ``add_observations``, ``get_statistics(X, full_cov=True)``
  returning ``(mean (T, N*, O), var (T, N*, O) | cov (T, N*, N*, O))``, ``get_mean``,
  ``get_common_hyperparameters`` and ``__repr__() == "ExactGP"``.

Where the reference called GPy (``GPRegression`` :201, ``Gaussian_noise.fix`` :205, ``set_XY`` :209, ``predict``
:271), this class calls ``libgpx.so`` through ``_gpx.Handle``.  GPy behaviours that callers can observe are kept:
likelihood variance defaults to 1.0 when ``noise_prior`` is falsy, ``noise_prior == 0`` is rejected by an assert, the
constant mean is ``mean(Y)`` when ``mean_prior`` is given, the predictive variance includes the noise variance
(``include_likelihood=True``), a fixed 1e-8 jitter is added with the noise, and a failed factorisation is retried
with growing jitter like ``jitchol`` before raising ``LinAlgError``.

``do_optimize=True`` runs GPy's randomize + L-BFGS-B loop with the objective and its
gradient evaluated on the GPU (``optimize.py``, ``gpx_lml_grad``); ``num_mcmc > 0`` runs GPy's HMC recipe on the same
objective and predictions gain the leading hyper-sample axis.  Not built: GPy prior objects on the noise and the RBF
Jacobian/Hessian helpers (``NotImplementedError`` instead of silently doing something else).
"""
import os
import pathlib
import pickle
import warnings

import numpy as np

from . import _gpx
from .config_helpers import ConfigMixin, lazy_construct_from_module

GPY_JITTER = 1e-8  # GPy ExactGaussianInference adds variance + 1e-8 to the diagonal


class MarginalLogLikelihoodMixin:
    def get_marginal_log_likelihood(self, X, Y):
        raise NotImplementedError


class SaveMixin:
    """Pickle-to-directory persistence (reference core_models.py:17-35)."""

    def __getstate__(self):
        return self.__dict__.copy()

    def save(self, PATH):
        pathlib.Path(PATH).mkdir(parents=True, exist_ok=True)
        with open(os.path.join(PATH, 'model.pickle'), 'wb') as fd:
            pickle.dump(self, fd, pickle.HIGHEST_PROTOCOL)

    @classmethod
    def load(cls, PATH):
        with open(os.path.join(PATH, 'model.pickle'), 'rb') as fd:
            model = pickle.load(fd)
        return model.post_load_hook(PATH)

    def post_load_hook(self, PATH):
        return self


class BaseModel(SaveMixin):
    """The contract every model obeys (reference core_models.py:38-123)."""

    def __init__(self):
        self._X = None
        self._Y = None

    def __repr__(self):
        return "{}".format(type(self).__name__)

    def get_incumbent(self):
        i = np.argmax(self._Y)
        return self._X[i], self._Y[i]

    @property
    def X(self):
        return self._X

    @property
    def Y(self):
        return self._Y

    def init(self, X, Y, Y_dir=None):
        self._X = X
        self._Y = Y
        self.Y_dir = Y_dir
        self._fit(X, Y, Y_dir=self.Y_dir)

    def add_observations(self, X_new, Y_new, Y_dir_new=None):
        assert self._X is not None, "Call init first"
        X = np.concatenate([self._X, X_new])
        Y = np.concatenate([self._Y, Y_new])
        if self.Y_dir is not None:
            Y_dir = np.concatenate([self.Y_dir, Y_dir_new])
        else:
            Y_dir = None
        self.init(X, Y, Y_dir)

    def set_train_data(self, X, Y):
        raise NotImplementedError

    def _fit(self, X, Y, Y_dir=None):
        raise NotImplementedError

    def _get_statistics(self, X, full_cov=True):
        raise NotImplementedError

    def get_statistics(self, X, full_cov=True):
        return self._get_statistics(X, full_cov=full_cov)

    def get_mean(self, X):
        return self.get_statistics(X, full_cov=False)[0]


class ProbModel(BaseModel):
    def _get_statistics(self, X, full_cov=True):
        raise NotImplementedError


def _default_device():
    for key in ("GPX_DEVICE", "LOCAL_RANK"):
        if os.environ.get(key, "") != "":
            return int(os.environ[key])
    return 0


class GPModel(ConfigMixin, MarginalLogLikelihoodMixin, ProbModel):
    """Exact GP regression with a stationary kernel, fit and predicted on a B200 (reference core_models.py:138-271)."""

    def __init__(self,
                 kernel,
                 noise_prior=None,
                 do_optimize=False,
                 num_mcmc=0,
                 n_burnin=100,
                 subsample_interval=10,
                 step_size=1e-1,
                 leapfrog_steps=20,
                 mean_prior=None):
        super().__init__()
        self.kernel_constructor = kernel
        self.kernel = None
        self.noise_prior = noise_prior
        self.do_optimize = do_optimize
        self.num_mcmc = num_mcmc

        self.n_burnin = n_burnin
        self.subsample_interval = subsample_interval
        self.step_size = step_size
        self.leapfrog_steps = leapfrog_steps

        self.has_mcmc_warmup = False
        self.output_dim = None
        self.mean_prior = mean_prior

        # state that replaces the reference's `gpy_model`
        self.noise_variance = None
        self.mean_const = 0.0
        self.jitter = GPY_JITTER
        self._current_thetas = None
        self._handle = None
        self._fitted_theta = None
        self._noise_fixed = False
        self.device = None
        self.randomize_before_optimize = True   # GPModel calls gpy_model.randomize() before optimize()
        self.max_iters = 1000                    # paramz default

    @classmethod
    def process_config(cls, *, kernel=None, **kwargs):
        from . import kernels as kernels_module
        kernel = lazy_construct_from_module(kernels_module, kernel)
        return dict(kernel=kernel, **kwargs)

    def __repr__(self):
        return "ExactGP"

    # -- pickling: device handles never travel; the factor is rebuilt on first use after load -------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        state['_handle'] = None
        state['_fitted_theta'] = None
        return state

    def get_common_hyperparameters(self):
        return {
            'lengthscale': self.kernel.lengthscale,
            'noise': np.array([self.noise_variance]),
        }

    # -- parameter vector in GPy's order: [mean constant (if any), kern.variance, kern.lengthscale.., noise] --------
    def _param_array(self):
        parts = []
        if self.mean_prior is not None:
            parts.append(np.array([self.mean_const]))
        parts += [self.kernel.variance, self.kernel.lengthscale, np.array([self.noise_variance])]
        return np.concatenate(parts).astype(np.float64)

    def _set_params(self, theta):
        theta = np.asarray(theta, dtype=np.float64)
        i = 0
        if self.mean_prior is not None:
            self.mean_const = float(theta[0])
            i = 1
        nl = self.kernel.lengthscale.size
        self.kernel.variance = theta[i:i + 1].copy()
        self.kernel.lengthscale = theta[i + 1:i + 1 + nl].copy()
        self.noise_variance = float(theta[i + 1 + nl])

    def _get_handle(self):
        if self._handle is None:
            self._handle = _gpx.Handle(self.device if self.device is not None else _default_device())
            # one process per GPU: when torch.distributed is up (torchrun), shard the factor over the ranks
            if os.environ.get("GPX_DISTRIBUTED", "1") != "0" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                    from . import distributed as gdist
                    gdist.init_handle_comm(self._handle, dist)
        return self._handle

    def _factorize(self, retry=True):
        """K-build + Cholesky + alpha + LML on the GPU for the current parameters (jitchol-style retry)."""
        h = self._get_handle()
        X, Y = self._X, self._Y
        jitter = self.jitter
        extra = 0.0
        tries = 0
        while True:
            try:
                h.fit(X, Y, self.kernel.kernel_id, float(self.kernel.variance[0]), self.kernel.lengthscale,
                      self.noise_variance, jitter + extra, self.mean_const)
                break
            except _gpx.NotPositiveDefiniteError:
                if not retry:
                    raise
                # GPy util.linalg.jitchol: add diag.mean() * 1e-6 * 10^k for k < 5, then give up
                tries += 1
                if tries > 5:
                    raise np.linalg.LinAlgError("not positive definite, even with jitter.")
                diag_mean = float(self.kernel.variance[0]) + self.noise_variance + jitter
                extra = diag_mean * 1e-6 * 10 ** (tries - 1)
                warnings.warn("Added jitter of {:.10e}".format(extra))
        self._fitted_theta = self._param_array().copy()

    def _ensure_fitted(self, theta):
        if self._handle is None or self._fitted_theta is None or not np.array_equal(self._fitted_theta, theta):
            self._set_params(theta)
            self._factorize()

    def _fit(self, X, Y, Y_dir=None):
        assert X.shape[0] == Y.shape[0], \
            "X and Y has to match size. It was {} and {} respectively".format(X.shape[0], Y.shape[0])
        self.output_dim = Y.shape[-1]
        input_dim = X.shape[-1]
        first_fit = self.noise_variance is None
        if first_fit or self.kernel is None:
            self.kernel = self.kernel_constructor(input_dim)

        if self.mean_prior is not None:
            self.mean_const = float(np.mean(Y))
        else:
            self.mean_const = 0.0

        if first_fit:
            self.noise_variance = 1.0  # GPy GPRegression default likelihood variance
            self._noise_fixed = False
            if self.noise_prior:
                if isinstance(self.noise_prior, (float, int)):
                    assert self.noise_prior != 0, "Zero noise is ignored. Use e.g. 1e-10 instead."
                    self.noise_variance = float(self.noise_prior)
                    self._noise_fixed = True         # Gaussian_noise.fix(noise_prior)
                else:
                    raise NotImplementedError("GPy prior objects on the noise variance are not supported; "
                                              "pass a number (fixed noise) or None (free noise)")
        if self.do_optimize and self.num_mcmc > 0:
            self._sample_hyperparameters()           # sets _current_thetas (one per retained HMC sample)
            return
        if self.do_optimize:
            self._optimize_hyperparameters()
        else:
            self._factorize()
        self._current_thetas = [self._param_array()]

    def _free_parameter_plumbing(self):
        """(names, positive mask, unpack(theta), objective(theta) -> (lml, grad), theta0) for the free parameters in GPy's
        order [mean const, variance, lengthscales, noise (unless fixed)]."""
        nl = self.kernel.lengthscale.size
        has_mean = self.mean_prior is not None
        names = (["mean"] if has_mean else []) + ["variance"] + ["ls%d" % q for q in range(nl)] + ([] if self._noise_fixed else ["noise"])
        positive = np.array([n != "mean" for n in names])

        def unpack(theta):
            i = 0
            if has_mean:
                self.mean_const = float(theta[0]); i = 1
            self.kernel.variance = np.array([theta[i]])
            self.kernel.lengthscale = np.array(theta[i + 1:i + 1 + nl], dtype=np.float64)
            if not self._noise_fixed:
                self.noise_variance = float(theta[i + 1 + nl])

        def objective(theta):
            unpack(theta)
            self._factorize(retry=False)
            h = self._get_handle()
            g = h.lml_grad(nl)            # [variance, ls..., noise, mean]
            out = ([g[nl + 2]] if has_mean else []) + [g[0]] + list(g[1:1 + nl]) + ([] if self._noise_fixed else [g[nl + 1]])
            return h.lml()[0], np.array(out)

        theta0 = np.array(([self.mean_const] if has_mean else []) + [float(self.kernel.variance[0])] +
                          list(self.kernel.lengthscale) + ([] if self._noise_fixed else [self.noise_variance]))
        return names, positive, unpack, objective, theta0

    def _sample_hyperparameters(self):
        """``GPy.inference.mcmc.HMC(gpy_model, stepsize).sample(...)`` then burn-in / thinning (core_models.py:212-219).
        The chain state persists across refits (``has_mcmc_warmup``), like the reference's ``self.hmc`` object."""
        from . import optimize as gopt
        assert not self._noise_fixed, "HMC samples every parameter: leave the noise free (noise_prior=None)"
        names, positive, unpack, objective, theta0 = self._free_parameter_plumbing()
        ss = gopt.hmc_sample(objective, theta0, positive, num_samples=self.n_burnin + self.num_mcmc * self.subsample_interval,
                             hmc_iters=self.leapfrog_steps, stepsize=self.step_size)
        self.has_mcmc_warmup = True
        kept = ss[self.n_burnin::self.subsample_interval]
        self._current_thetas = [np.array(t) for t in kept]
        unpack(self._current_thetas[-1])
        self._factorize()

    def _optimize_hyperparameters(self):
        """``gpy_model.randomize(); gpy_model.optimize()`` (reference core_models.py:221-223): L-BFGS-B on -LML in the
        Logexp-transformed space, every evaluation = one GPU factorisation + one GPU gradient (gpx_lml_grad)."""
        from . import optimize as gopt
        names, positive, unpack, objective, theta0 = self._free_parameter_plumbing()
        x0 = None
        if getattr(self, "randomize_before_optimize", True):
            x0 = gopt.randomize(theta0.size)      # GPy randomizes the mean constant too (unconstrained N(0,1) draw)
        theta, lml, info = gopt.optimize_lbfgsb(objective, theta0, positive, x0=x0, max_iters=getattr(self, "max_iters", 1000))
        unpack(theta)
        self._factorize()
        self.optimization_info = info

    def _predict(self, X, full_cov=True):
        num_X = X.shape[0]
        num_obj = self.output_dim
        T = len(self._current_thetas)
        mean = np.zeros((T, num_X, num_obj))
        covar = np.zeros((T, num_X, num_X, num_obj)) if full_cov else None
        var = None if full_cov else np.zeros((T, num_X, num_obj))
        for i, theta in enumerate(self._current_thetas):
            self._ensure_fitted(theta)
            h = self._get_handle()
            if full_cov:
                mean_i, cov_i = h.predict_cov(X, include_noise=True, output_dim=num_obj)
                mean[i, :] = mean_i
                covar[i, :] = cov_i[:, :, None]
            else:
                mean_i, var_i = h.predict(X, want_var=True, include_noise=True, output_dim=num_obj)
                mean[i, :] = mean_i
                var[i, :] = var_i[:, None]
        if full_cov:
            return mean, covar
        return mean, var

    def _get_statistics(self, X, full_cov=True):
        assert self._current_thetas is not None, "Call init first"
        return self._predict(np.asarray(X), full_cov=full_cov)

    def get_marginal_log_likelihood(self, X=None, Y=None):
        """Un-normalised log marginal likelihood of the fitted data (GPy ``log_likelihood()``)."""
        if X is not None and Y is not None and (X is not self._X or Y is not self._Y):
            self.init(X, Y)
        self._ensure_fitted(self._current_thetas[0])
        return self._get_handle().lml()[0]

    def predict_jacobian(self, X, full_cov=True):
        """Reference core_models.py:313-333.  There the method reads ``self.normalizer``, an attribute GPModel never
        defines, so it cannot run as written; not built here."""
        raise NotImplementedError("predict_jacobian is not built (the reference version depends on an undefined "
                                  "`self.normalizer`); see SURVEY.md section 8(f)")

    def predict_hessian(self, X, full_cov=True):
        """Reference core_models.py:335-380 (RBF only, same undefined ``self.normalizer``); not built here."""
        raise NotImplementedError("predict_hessian is not built; see SURVEY.md section 8(f)")

    def phase_times_ms(self):
        """Per-phase CUDA-event times of the last fit / predict (the numbers Runner logs as time.training/pred)."""
        return self._get_handle().phase_ms()


class DerivativeGPModel(GPModel):
    """Fits the GP to the derivative observations (reference core_models.py:383-390)."""

    def _fit(self, X, Y, Y_dir=None):
        return super(DerivativeGPModel, self)._fit(X, Y_dir, Y_dir=None)
