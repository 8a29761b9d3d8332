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
objective and predictions gain the leading hyper-sample axis (one resident factorisation per retained sample).
``add_observations`` appends to the resident factor (``gpx_append``) when only the data changed; ``get_mean`` and small
batches take the latency path of ``gpx_predict``; ``predict_jacobian`` / ``predict_hessian`` evaluate the reference's
expressions from sums computed on the GPU.  A non-numeric ``noise_prior`` is a prior object with GPy's interface (``priors.py``).
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
        # additive knobs (not in the reference)
        self.distributed = None                  # shard the factor over the torch.distributed ranks: True / False, or None =
                                                 # only when the environment says GPX_DISTRIBUTED=1 (opt-in: ranks of a sweep
                                                 # that fit different models must not meet in a collective)
        self.incremental = True                  # add_observations appends to the resident factor when it can
        self.latency_inverse = None              # explicit inverse of the factor for row-by-row predictions / appends (gpx option
                                                 # "latency_inverse"): None = automatic (after 16 small predicts of a fit),
                                                 # True = at the first one, False = never
        self.max_cached_thetas = 16              # HMC: factorisations kept resident, one handle per hyper-sample
        self._normalize_input = False            # fused NormalizerModel (set by NormalizerModel, normalizer.py)
        self._normalize_output = False
        self._theta_handles = {}

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
        state['_theta_handles'] = {}
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        for k, v in (("distributed", None), ("incremental", True), ("latency_inverse", None), ("max_cached_thetas", 16), ("_normalize_input", False),
                     ("_normalize_output", False), ("_theta_handles", {})):
            self.__dict__.setdefault(k, v)

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

    # -- handles -------------------------------------------------------------------------------------------------
    def _wants_sharding(self):
        if self.distributed is not None:
            return bool(self.distributed)
        return os.environ.get("GPX_DISTRIBUTED", "0") == "1"

    def _new_handle(self, shard):
        h = _gpx.Handle(self.device if self.device is not None else _default_device())
        if shard and int(os.environ.get("WORLD_SIZE", "1")) > 1:
            import torch.distributed as dist
            if not (dist.is_available() and dist.is_initialized()):
                raise _gpx.GpxError("a sharded GPModel needs an initialised torch.distributed process group (torchrun)")
            if dist.get_world_size() > 1:
                from . import distributed as gdist
                gdist.init_handle_comm(h, dist)
        h.set_normalizer(self._normalize_input, self._normalize_output)
        if getattr(self, "latency_inverse", None) is not None:
            h.set_option("latency_inverse", 1 if self.latency_inverse else 0)
        return h

    def _get_handle(self):
        if self._handle is None:
            # one process per GPU: sharding is OPT-IN (model.distributed = True or GPX_DISTRIBUTED=1); every rank must
            # then call init / get_statistics with identical data and hyper-parameters (checked at fit time)
            self._handle = self._new_handle(self._wants_sharding())
        return self._handle

    def _sharded(self, h=None):
        h = h or self._handle
        return h is not None and getattr(h, "world", 1) > 1

    def _fit_handle(self, h, retry=True):
        """K-build + Cholesky + alpha + LML on handle ``h`` for the current parameters (jitchol-style retry)."""
        X, Y = self._X, self._Y
        if self._sharded(h):
            from . import distributed as gdist
            gdist.check_same_problem(h, X, Y, self._param_array())
        jitter = self.jitter
        extra = 0.0
        tries = 0
        while True:
            try:
                h.fit(X, Y, self.kernel.kernel_id, float(self.kernel.variance[0]), self.kernel.lengthscale,
                      self.noise_variance, jitter + extra, self.mean_const, ARD=self.kernel.ARD)
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

    def _factorize(self, retry=True):
        self._fit_handle(self._get_handle(), retry=retry)
        self._fitted_theta = self._param_array().copy()

    def _ensure_fitted(self, theta):
        if self._handle is None or self._fitted_theta is None or not np.array_equal(self._fitted_theta, theta):
            self._set_params(theta)
            self._factorize()

    def _handle_for_theta(self, theta):
        """The fitted handle for hyper-sample ``theta``.  One sample: the model's own handle.  HMC (T > 1): the reference
        re-factorises for every theta at every predict (``gpy_model[:] = theta``, core_models.py:249-251); here up to
        ``max_cached_thetas`` factorisations stay resident, one handle each, so repeated predictions cost O(T N^2 N*)
        instead of O(T N^3)."""
        if len(self._current_thetas) == 1 or self._sharded() or self._wants_sharding():
            self._ensure_fitted(theta)
            return self._get_handle()
        key = np.asarray(theta, dtype=np.float64).tobytes()
        h = self._theta_handles.get(key)
        if h is None:
            while len(self._theta_handles) >= max(1, self.max_cached_thetas):
                self._theta_handles.pop(next(iter(self._theta_handles))).close()
            h = self._new_handle(False)
            self._set_params(theta)
            self._fit_handle(h)
            self._theta_handles[key] = h
        return h

    def _drop_theta_handles(self):
        for h in self._theta_handles.values():
            h.close()
        self._theta_handles = {}

    # -- fitting -------------------------------------------------------------------------------------------------
    def _fit(self, X, Y, Y_dir=None):
        assert X.shape[0] == Y.shape[0], \
            "X and Y has to match size. It was {} and {} respectively".format(X.shape[0], Y.shape[0])
        self.output_dim = Y.shape[-1]
        input_dim = X.shape[-1]
        first_fit = self.noise_variance is None
        if first_fit or self.kernel is None:
            self.kernel = self.kernel_constructor(input_dim)
        self._drop_theta_handles()

        if first_fit:
            # The reference builds gpy_model once (core_models.py:190-207) and later only calls set_XY (:209): the Constant
            # mean mapping keeps the value it was created with (or the optimised one), it is NOT reset to the new mean(Y).
            if self.mean_prior is not None:
                Yn = Y
                if self._normalize_output:       # the inner model of a NormalizerModel sees z-scored outputs
                    Yn = (Y - np.mean(Y, axis=0)) / np.std(Y, axis=0)
                self.mean_const = float(np.mean(Yn))
            else:
                self.mean_const = 0.0
            self.noise_variance = 1.0  # GPy GPRegression default likelihood variance
            self._noise_fixed = False
            if self.noise_prior:
                if isinstance(self.noise_prior, (float, int)):
                    assert self.noise_prior != 0, "Zero noise is ignored. Use e.g. 1e-10 instead."
                    self.noise_variance = float(self.noise_prior)
                    self._noise_fixed = True         # Gaussian_noise.fix(noise_prior)
                else:
                    # Gaussian_noise.variance.set_prior(noise_prior) (reference core_models.py:206-207): any object with GPy's
                    # prior interface (lnpdf / lnpdf_grad / rvs) -- thesis_code_b200.priors restates GPy's Gamma / LogGaussian / Gaussian
                    from . import priors as _priors
                    if not _priors.is_prior(self.noise_prior):
                        raise NotImplementedError("noise_prior must be a number (fixed noise), None (free noise) or an object "
                                                  "with GPy's prior interface lnpdf / lnpdf_grad / rvs (see thesis_code_b200.priors)")
        if self.do_optimize and self._wants_sharding() and int(os.environ.get("WORLD_SIZE", "1")) > 1:
            # sharded optimisation: gpx_lml_grad works on a sharded handle whose ranks hold the whole factor (it says so
            # loudly otherwise); the ranks must draw the same random start point / momenta
            from . import distributed as gdist
            gdist.sync_rng(self._get_handle())
        if self.do_optimize and self.num_mcmc > 0:
            self._sample_hyperparameters()           # sets _current_thetas (one per retained HMC sample)
            return
        if self.do_optimize:
            self._optimize_hyperparameters()
        else:
            self._factorize()
        self._current_thetas = [self._param_array()]

    def add_observations(self, X_new, Y_new, Y_dir_new=None):
        """Reference ``BaseModel.add_observations`` (core_models.py:65-77) concatenates and refits from scratch, once
        per Bayesian-optimisation step (src/algorithms.py:120-143).  With fixed hyper-parameters the refit is replaced
        by ``gpx_append``: the new rows of K are solved against the resident factor (O(k N^2)); anything that changes
        more than the data (do_optimize, HMC, the normaliser's statistics, a sharded factor) takes the reference's route."""
        assert self._X is not None, "Call init first"
        can_append = (self.incremental and not self.do_optimize and self._handle is not None
                      and self._current_thetas is not None and len(self._current_thetas) == 1
                      and self._fitted_theta is not None and np.array_equal(self._fitted_theta, self._current_thetas[0])
                      and not self._sharded() and not (self._normalize_input or self._normalize_output)
                      and self.Y_dir is None)
        if not can_append:
            return super().add_observations(X_new, Y_new, Y_dir_new)
        X_new = np.asarray(X_new, dtype=np.float64)
        Y_new = np.asarray(Y_new, dtype=np.float64)
        assert X_new.shape[0] == Y_new.shape[0] and X_new.shape[1:] == self._X.shape[1:] and Y_new.shape[1:] == self._Y.shape[1:]
        h = self._handle
        h.set_option("reserve_rows", max(256, self._X.shape[0] // 8))   # later (re)layouts leave room for the next appends
        try:
            h.append(X_new, Y_new)
        except (_gpx.NotPositiveDefiniteError, _gpx.AppendUnsupported):
            self._fitted_theta = None
            return super().add_observations(X_new, Y_new, Y_dir_new)
        self._X = np.concatenate([self._X, X_new])
        self._Y = np.concatenate([self._Y, Y_new])

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

        prior = self._noise_prior_object()

        def objective(theta):
            unpack(theta)
            self._factorize(retry=True)   # GPy's inference runs jitchol inside every objective evaluation
            h = self._get_handle()
            g = h.lml_grad(nl)            # [variance, ls..., noise, mean]
            out = ([g[nl + 2]] if has_mean else []) + [g[0]] + list(g[1:1 + nl]) + ([] if self._noise_fixed else [g[nl + 1]])
            value, out = h.lml()[0], np.array(out)
            if prior is not None:
                # paramz Priorizable: objective = -(log likelihood + log prior), the log prior of a transformed parameter
                # includes the log Jacobian of the Logexp transform
                from .optimize import Logexp
                s2 = self.noise_variance
                value = value + float(np.sum(prior.lnpdf(np.array([s2])))) + float(Logexp.log_jacobian(s2))
                out[-1] += float(np.sum(prior.lnpdf_grad(np.array([s2])))) + float(Logexp.log_jacobian_grad(s2))
            return value, out

        theta0 = np.array(([self.mean_const] if has_mean else []) + [float(self.kernel.variance[0])] +
                          list(self.kernel.lengthscale) + ([] if self._noise_fixed else [self.noise_variance]))
        return names, positive, unpack, objective, theta0

    def _noise_prior_object(self):
        from . import priors as _priors
        if self.noise_prior and not isinstance(self.noise_prior, (float, int)) and not self._noise_fixed and _priors.is_prior(self.noise_prior):
            return self.noise_prior
        return None

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
            prior = self._noise_prior_object()
            if prior is not None:                 # ... and draws a priored parameter from its prior (paramz Priorizable.randomize)
                x0[-1] = float(gopt.Logexp.finv(np.asarray(prior.rvs(1), dtype=np.float64)[0]))
        theta, lml, info = gopt.optimize_lbfgsb(objective, theta0, positive, x0=x0, max_iters=getattr(self, "max_iters", 1000))
        unpack(theta)
        self._factorize()
        self.optimization_info = info

    # -- prediction ----------------------------------------------------------------------------------------------
    def _predict(self, X, full_cov=True, want_var=True):
        num_X = X.shape[0]
        num_obj = self.output_dim
        T = len(self._current_thetas)
        mean = np.zeros((T, num_X, num_obj))
        covar = np.zeros((T, num_X, num_X, num_obj)) if full_cov else None
        var = None if (full_cov or not want_var) else np.zeros((T, num_X, num_obj))
        fused_y = bool(self._normalize_output)
        if num_X == 0:                      # nothing to predict: empty arrays of the right shape, no device work
            return (mean, covar) if full_cov else (mean, var)
        for i, theta in enumerate(self._current_thetas):
            h = self._handle_for_theta(theta)
            if full_cov:
                mean_i, cov_i = h.predict_cov(X, include_noise=True, output_dim=num_obj)
                mean[i, :] = mean_i
                covar[i, :] = cov_i[:, :, None]
            elif not want_var:
                mean[i, :] = h.predict(X, want_var=False, output_dim=num_obj)[0]
            else:
                # with the fused output normaliser every output carries its own variance scale (var_cols = O)
                mean_i, var_i = h.predict(X, want_var=True, include_noise=True, output_dim=num_obj,
                                          var_cols=num_obj if (fused_y and num_obj > 1) else 1)
                mean[i, :] = mean_i
                var[i, :] = var_i if var_i.ndim == 2 else var_i[:, None]
        if full_cov:
            return mean, covar
        return mean, var

    def _get_statistics(self, X, full_cov=True):
        assert self._current_thetas is not None, "Call init first"
        return self._predict(np.asarray(X), full_cov=full_cov)

    def get_mean(self, X):
        """Reference: ``get_statistics(X, full_cov=False)[0]`` (core_models.py:93-94), i.e. GPy computes the variance too
        and throws it away.  Here the variance pass is skipped: a 1-row call (src/growth_model_GPR/ipopt_wrapper.py:55
        issues them by the million) is three kernels replayed as one CUDA graph."""
        assert self._current_thetas is not None, "Call init first"
        return self._predict(np.asarray(X), full_cov=False, want_var=False)[0]

    def get_marginal_log_likelihood(self, X=None, Y=None):
        """Un-normalised log marginal likelihood of the fitted data (GPy ``log_likelihood()``)."""
        if X is not None and Y is not None and (X is not self._X or Y is not self._Y):
            self.init(X, Y)
        self._ensure_fitted(self._current_thetas[0])
        return self._get_handle().lml()[0]

    # -- posterior-mean derivatives (reference core_models.py:313-380; consumer: CurvatureAcquisition,
    #    src/acquisition_functions.py:86,113) --------------------------------------------------------------------------
    def _raw_moments(self, X, order):
        """S0 (N*), S1 (N*, D), S2 (N*, D, D): sums over the training points of w, w dx_k, w dx_k dx_l with
        w = k(x_j, x*) alpha_j and dx = x* - x_j in INPUT coordinates (the device works in the kernel's coordinates: the
        ARD scaling is undone here)."""
        from .kernels import GPyRBF
        assert isinstance(self.kernel, GPyRBF), "Only support RBF for now"     # reference core_models.py:343
        assert not (self._normalize_input or self._normalize_output), \
            "derivatives are taken in the coordinates the GP sees; call them on the un-wrapped GPModel"
        self._ensure_fitted(self._current_thetas[0])
        X = np.asarray(X, dtype=np.float64)
        S0, S1, S2 = self._get_handle().predict_moments(X, order=order)
        if self.kernel.ARD:
            ls = self.kernel.lengthscale
            S1 = S1 * ls[None, :]
            if S2 is not None:
                S2 = S2 * ls[None, :, None] * ls[None, None, :]
        return S0, S1, S2

    def _lengthscale_per_dim(self):
        D = self._X.shape[1]
        ls = self.kernel.lengthscale
        return ls if ls.size == D else np.repeat(ls, D)

    def predict_jacobian(self, X, full_cov=True, exact=False):
        """Jacobian of the posterior mean, shape (N*, D), and None for its variance (reference core_models.py:313-333).

        The reference's expression is ``-Lambda^-1 sum_j (x* - x_j) k(x_j, x*) alpha_j`` with ``Lambda^-1 = diag(1 / l)``
        -- one power of the lengthscale short of the derivative of an RBF mean -- and its first line reads an attribute
        (``self.normalizer``) GPModel never defines.  Default: the reference's expression, evaluated on the GPU
        (``gpx_predict_moments``); ``exact=True``: the true gradient ``-S1 / l^2``."""
        _, S1, _ = self._raw_moments(X, order=1)
        l = self._lengthscale_per_dim()
        return (-S1 / (l ** 2 if exact else l)[None, :]), None

    def predict_hessian(self, X, full_cov=True, exact=False):
        """Hessian of the posterior mean, (N*, D, D), and None (reference core_models.py:335-380; RBF only).

        Reference-literal by default: ``H[i,k,l] = -(1/l_k) (S0[i] - S2[i,k,l] / l_l)`` -- the reference adds ``S0`` to
        every entry (its ``Ones`` einsum) and, like the Jacobian, uses 1/l where the calculus gives 1/l^2.
        ``exact=True``: ``S2[i,k,l] / (l_k^2 l_l^2) - delta_kl S0[i] / l_k^2``."""
        S0, _, S2 = self._raw_moments(X, order=2)
        l = self._lengthscale_per_dim()
        if exact:
            H = S2 / (l[None, :, None] ** 2 * l[None, None, :] ** 2)
            idx = np.arange(l.size)
            H[:, idx, idx] -= S0[:, None] / l[None, :] ** 2
            return H, None
        return -(S0[:, None, None] - S2 / l[None, None, :]) / l[None, :, None], None

    def phase_times_ms(self):
        """Per-phase CUDA-event times of the last fit / predict (the numbers Runner logs as time.training/pred)."""
        return self._get_handle().phase_ms()


class DerivativeGPModel(GPModel):
    """Fits the GP to the derivative observations (reference core_models.py:383-390)."""

    def _fit(self, X, Y, Y_dir=None):
        return super(DerivativeGPModel, self)._fit(X, Y_dir, Y_dir=None)
