"""Hyper-parameter optimisation in GPy's conventions (what ``gpy_model.randomize(); gpy_model.optimize()`` does when the
reference builds ``GPModel`` with ``do_optimize=True``, src/models/core_models.py:211-223):

* positive parameters (kernel variance, lengthscales, Gaussian noise variance) live in paramz' ``Logexp`` transformed
  space, ``theta = log(1 + exp(x))``; the constant mean (if any) is unconstrained;
* ``randomize()`` draws the transformed parameters from N(0, 1) with the global numpy RNG;
* ``optimize()`` minimises ``-LML`` with scipy's L-BFGS-B (``max_iters=1000``), gradients chained through the transform.

The objective is a callable ``theta -> (lml, dlml/dtheta)``; GPModel passes the GPU one (``gpx_fit`` + ``gpx_lml_grad``).
"""
import numpy as np
from scipy import optimize as _sopt

_LIM_VAL = 36.0
_LOG_LIM_VAL = np.log(np.finfo(np.float64).max)


class Logexp:
    """paramz.transformations.Logexp"""

    @staticmethod
    def f(x):
        x = np.asarray(x, dtype=np.float64)
        return np.where(x > _LIM_VAL, x, np.log1p(np.exp(np.clip(x, -_LOG_LIM_VAL, _LIM_VAL))))

    @staticmethod
    def finv(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f > _LIM_VAL, f, np.log(np.expm1(f)))

    @staticmethod
    def gradfactor(f, df):
        f = np.asarray(f, dtype=np.float64)
        return df * np.where(f > _LIM_VAL, 1.0, -np.expm1(-f))

    # a prior on a Logexp-transformed parameter is a density on the constrained value; the optimiser works on the
    # unconstrained one, so paramz adds the log Jacobian of the transform to the log prior (and its derivative)
    @staticmethod
    def log_jacobian(f):
        f = np.asarray(f, dtype=np.float64)
        return np.where(f > _LIM_VAL, f, np.log(np.expm1(f))) - f

    @staticmethod
    def log_jacobian_grad(f):
        f = np.asarray(f, dtype=np.float64)
        return 1.0 / np.expm1(f)


def randomize(n, rand_gen=None):
    """paramz ``randomize``: transformed parameters ~ N(0, 1) (global numpy RNG unless ``rand_gen`` is given)."""
    rand_gen = np.random.normal if rand_gen is None else rand_gen
    return rand_gen(size=n)


def optimize_lbfgsb(objective, theta0, positive, x0=None, max_iters=1000, messages=False):
    """Minimise -LML over the free parameters.

    objective(theta) -> (lml, grad) with grad = dLML/dtheta (same ordering as theta);
    positive[i] says whether theta[i] is Logexp-transformed.  ``x0``: start in transformed space (e.g. from
    :func:`randomize`); default: the transform of ``theta0``.  Returns (theta_opt, lml_opt, info).
    """
    theta0 = np.asarray(theta0, dtype=np.float64)
    positive = np.asarray(positive, dtype=bool)

    def to_theta(x):
        return np.where(positive, Logexp.f(x), x)

    if x0 is None:
        x0 = np.where(positive, Logexp.finv(np.where(positive, theta0, 1.0)), theta0)
    evals = {"n": 0, "fails": 0}

    def f_fp(x):
        theta = to_theta(x)
        evals["n"] += 1
        try:
            lml, g = objective(theta)
        except (np.linalg.LinAlgError, ZeroDivisionError, ValueError, RuntimeError):
            # paramz Model._objective_grads: swallow a few numerical failures and return a huge objective (RuntimeError covers
            # the library's own GpxError, e.g. a variance or lengthscale that the Logexp transform underflowed to 0)
            evals["fails"] += 1
            if evals["fails"] > 10:
                raise
            return np.inf, np.zeros_like(x)
        g = np.asarray(g, dtype=np.float64)
        gx = np.where(positive, Logexp.gradfactor(theta, g), g)
        return -float(lml), -gx

    res = _sopt.minimize(f_fp, np.asarray(x0, dtype=np.float64), jac=True, method="L-BFGS-B",
                         options={"maxfun": max_iters, "maxiter": max_iters})   # scipy defaults: ftol 2.2e-9, gtol 1e-5
    if messages:
        print(res.message)
    info = dict(warnflag=int(res.status), task=str(res.message), nit=int(res.nit), funcalls=int(res.nfev),
                evaluations=evals["n"], x=res.x)
    return to_theta(res.x), -float(res.fun), info


def hmc_sample(objective, theta0, positive, num_samples, hmc_iters=20, stepsize=1e-1):
    """Hybrid Monte Carlo over the hyper-parameters in the transformed space, unit mass matrix -- the algorithm of
    ``GPy.inference.mcmc.HMC(model, stepsize).sample(num_samples, hmc_iters)`` that GPModel runs when ``num_mcmc > 0``
    (reference src/models/core_models.py:212-219).  Potential energy = -LML (no priors); each of the ``hmc_iters``
    leapfrog steps takes two half kicks of the momentum around one full position step; Metropolis accept/reject on the
    Hamiltonian.  Like GPy, row i records the state *before* proposal i unless the proposal is accepted.
    Uses the global numpy RNG (np.random), as GPy does.  Returns (num_samples, n_params) constrained parameter values.
    """
    theta0 = np.asarray(theta0, dtype=np.float64)
    positive = np.asarray(positive, dtype=bool)
    n = theta0.size

    def to_theta(x):
        return np.where(positive, Logexp.f(x), x)

    def energy_and_grad(x, may_fail=True):
        theta = to_theta(x)
        try:
            lml, g = objective(theta)
        except (np.linalg.LinAlgError, ZeroDivisionError, ValueError, RuntimeError):
            if not may_fail:
                raise
            # a leapfrog step walked into parameters the model cannot be factorised at: infinite energy, so the
            # Metropolis test rejects the proposal instead of the whole chain aborting
            return np.inf, np.zeros(n)
        gx = np.where(positive, Logexp.gradfactor(theta, np.asarray(g, dtype=np.float64)), g)
        return -float(lml), -gx

    x = np.where(positive, Logexp.finv(np.where(positive, theta0, 1.0)), theta0)
    out = np.empty((num_samples, n))
    const = n * np.log(2 * np.pi) / 2.0
    for i in range(num_samples):
        p = np.random.multivariate_normal(np.zeros(n), np.eye(n))
        U, gU = energy_and_grad(x, may_fail=False)
        H_old = U + const + 0.5 * float(p @ p)
        x_old = x.copy()
        out[i] = to_theta(x)
        for _ in range(hmc_iters):
            p = p - 0.5 * stepsize * gU
            x = x + stepsize * p
            U, gU = energy_and_grad(x)
            p = p - 0.5 * stepsize * gU
        H_new = U + const + 0.5 * float(p @ p)
        if not np.isfinite(H_new):
            H_new = np.inf
        accept = 1.0 if H_old > H_new else float(np.exp(H_old - H_new))
        if np.random.rand() < accept:
            out[i] = to_theta(x)
        else:
            x = x_old
    return out
