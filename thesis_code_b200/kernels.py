"""Kernel specifications the experiment config dict may name.

The reference's ``src/kernels.py:4-7`` aliases four GPy kernels (``GPyExponential``, ``GPyMatern32``,
``GPyMatern52``, ``GPyRBF``) that ``GPModel`` constructs as ``kernel_constructor(input_dim)`` with the config kwargs
``variance``, ``lengthscale``, ``ARD`` (core_models.py:167-174,192).  Here they are plain parameter holders: the
arithmetic runs in the CUDA K-build kernels (csrc/kbuild.cu), selected by ``kernel_id``.
"""
import numpy as np

from ._gpx import KERNEL_IDS


class StationaryKernelSpec:
    """Same constructor contract as GPy's Stationary: (input_dim, variance=1., lengthscale=None, ARD=False)."""

    name = None

    def __init__(self, input_dim, variance=1., lengthscale=None, ARD=False):
        self.input_dim = int(input_dim)
        self.ARD = bool(ARD)
        self.variance = np.atleast_1d(np.asarray(variance, dtype=np.float64)).ravel()[:1].copy()
        assert self.variance.size == 1 and self.variance[0] > 0, "variance must be a positive scalar"
        if lengthscale is None:
            ls = np.ones(self.input_dim if self.ARD else 1)
        else:
            ls = np.atleast_1d(np.asarray(lengthscale, dtype=np.float64)).ravel().copy()
            if self.ARD:
                if ls.size == 1:
                    ls = np.ones(self.input_dim) * ls[0]
                assert ls.size == self.input_dim, "bad number of lengthscales"
            else:
                assert ls.size == 1, "Only 1 lengthscale needed for non-ARD kernel"
        assert np.all(ls > 0), "lengthscale must be positive"
        self.lengthscale = ls

    @property
    def kernel_id(self):
        return KERNEL_IDS[self.name]

    def __repr__(self):
        return "{}(input_dim={}, variance={}, lengthscale={}, ARD={})".format(
            self.name, self.input_dim, self.variance.tolist(), self.lengthscale.tolist(), self.ARD)


class GPyExponential(StationaryKernelSpec):
    name = "GPyExponential"


class GPyMatern32(StationaryKernelSpec):
    name = "GPyMatern32"


class GPyMatern52(StationaryKernelSpec):
    name = "GPyMatern52"


class GPyRBF(StationaryKernelSpec):
    name = "GPyRBF"
