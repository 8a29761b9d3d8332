"""Priors for the Gaussian noise variance (``GPModel(noise_prior=<prior object>)``).

The reference hands a non-numeric ``noise_prior`` to GPy: ``gpy_model.Gaussian_noise.variance.set_prior(noise_prior)``
(src/models/core_models.py:206-207).  GPy priors are duck-typed objects with ``lnpdf(x)``, ``lnpdf_grad(x)`` and ``rvs(n)``;
``GPModel`` accepts any object with those three methods -- a real ``GPy.priors.*`` instance where GPy is installed, or one of
the restatements below (GPy 1.9.6 ``GPy/core/parameterization/priors.py``: same parameterisation, same formulas) where it is
not.  How the optimiser uses them (paramz ``Priorizable.log_prior`` / ``_log_prior_gradients``): the objective is
``-LML - lnpdf(noise) - log_jacobian(noise)`` where the Jacobian term accounts for the Logexp transform of the positive
parameter; ``randomize()`` draws a priored parameter from its prior.
"""
import math

import numpy as np


class Prior:
    domain = "positive"

    def lnpdf(self, x):
        raise NotImplementedError

    def lnpdf_grad(self, x):
        raise NotImplementedError

    def rvs(self, n):
        raise NotImplementedError

    def pdf(self, x):
        return np.exp(self.lnpdf(x))


class Gaussian(Prior):
    """N(mu, sigma^2)."""
    domain = "real"

    def __init__(self, mu, sigma):
        self.mu, self.sigma = float(mu), float(sigma)
        self.sigma2 = self.sigma ** 2
        self.constant = -0.5 * math.log(2 * math.pi * self.sigma2)

    def lnpdf(self, x):
        return self.constant - 0.5 * np.square(np.asarray(x, dtype=np.float64) - self.mu) / self.sigma2

    def lnpdf_grad(self, x):
        return -(np.asarray(x, dtype=np.float64) - self.mu) / self.sigma2

    def rvs(self, n):
        return np.random.randn(n) * self.sigma + self.mu


class LogGaussian(Gaussian):
    """log x ~ N(mu, sigma^2)."""
    domain = "positive"

    def lnpdf(self, x):
        x = np.asarray(x, dtype=np.float64)
        return self.constant - 0.5 * np.square(np.log(x) - self.mu) / self.sigma2 - np.log(x)

    def lnpdf_grad(self, x):
        x = np.asarray(x, dtype=np.float64)
        return -((np.log(x) - self.mu) / self.sigma2 + 1.0) / x

    def rvs(self, n):
        return np.exp(np.random.randn(int(n)) * self.sigma + self.mu)


class Gamma(Prior):
    """Gamma(shape a, rate b): density b^a x^(a-1) exp(-b x) / Gamma(a)."""

    def __init__(self, a, b):
        self.a, self.b = float(a), float(b)
        self.constant = -math.lgamma(self.a) + self.a * math.log(self.b)

    @staticmethod
    def from_EV(E, V):
        """Gamma with expectation E and variance V (GPy ``Gamma.from_EV``)."""
        return Gamma(np.square(E) / V, E / V)

    def lnpdf(self, x):
        x = np.asarray(x, dtype=np.float64)
        return self.constant + (self.a - 1.0) * np.log(x) - self.b * x

    def lnpdf_grad(self, x):
        x = np.asarray(x, dtype=np.float64)
        return (self.a - 1.0) / x - self.b

    def rvs(self, n):
        return np.random.gamma(scale=1.0 / self.b, shape=self.a, size=n)


def is_prior(obj):
    """Duck-typed like GPy: anything with lnpdf / lnpdf_grad / rvs."""
    return all(callable(getattr(obj, m, None)) for m in ("lnpdf", "lnpdf_grad", "rvs"))
