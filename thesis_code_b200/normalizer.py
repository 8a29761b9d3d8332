"""Z-scoring wrapper around any model of this package, interface-compatible with the reference's
``src/models/normalizer.py`` (classes ``Normalizer`` / ``NormalizerModel``, lines 9-158): nearly every thesis config wraps
``GPModel`` in it (``notebook_header.py:17-26``).

The affine maps are O(N d) host work next to an O(N^3) GPU factorisation, so they stay in numpy.  Behaviour kept on
purpose: statistics are per column with population standard deviation (``np.std``), the variance is rescaled by
``std**2``, and a full covariance gets a trailing axis before rescaling (reference ``denormalize_covariance``).
"""
import os

import numpy as np

from .config_helpers import ConfigMixin, construct_from_module
from .core_models import BaseModel, SaveMixin


def zero_mean_unit_var_normalization(X, mean=None, std=None):
    """(X - mean) / std with column statistics taken from X unless given; returns (Z, mean, std)."""
    if mean is None:
        mean = X.mean(axis=0)
    if std is None:
        std = X.std(axis=0)
    return (X - mean) / std, mean, std


def zero_mean_unit_var_unnormalization(X_normalized, mean, std):
    return mean + std * X_normalized


class Normalizer:
    """Column-wise z-score fitted on the array it is constructed from."""

    def __init__(self, X_train):
        Z, mu, sigma = zero_mean_unit_var_normalization(np.asarray(X_train))
        self.X = Z
        self._X_mean = mu
        self._X_std = sigma

    # -- forward -----------------------------------------------------------------------------------------------
    def get_transformed(self):
        return self.X

    def normalize(self, X):
        return (X - self._X_mean) / self._X_std

    # -- backward ----------------------------------------------------------------------------------------------
    def denormalize(self, X):
        return zero_mean_unit_var_unnormalization(X, self._X_mean, self._X_std)

    def denormalize_variance(self, X_var):
        return np.square(self._X_std) * X_var

    def denormalize_covariance(self, covariance):
        return np.square(self._X_std) * covariance[..., None]


class NormalizerModel(ConfigMixin, BaseModel):
    """``NormalizerModel(model, normalize_input=True, normalize_output=True)``; the wrapped model only ever sees z-scored data."""

    _TRANSIENT = ("model", "_X", "_Y", "X_normalizer", "Y_normalizer")

    def __init__(self, model, normalize_input=True, normalize_output=True):
        super().__init__()
        self.model = model
        self._normalize_input = normalize_input
        self._normalize_output = normalize_output
        self.X_normalizer = self.Y_normalizer = None

    @classmethod
    def process_config(cls, *, model=None, **kwargs):
        import thesis_code_b200 as models_module
        kwargs["model"] = construct_from_module(models_module, model)
        return kwargs

    # -- fitting -----------------------------------------------------------------------------------------------
    def _normalize(self, X, Y):
        """(Re)build the normalizers from the raw data and return what the wrapped model should be given."""
        if self._normalize_input:
            self.X_normalizer = Normalizer(X)
            X = self.X_normalizer.get_transformed()
        if self._normalize_output:
            self.Y_normalizer = Normalizer(Y)
            Y = self.Y_normalizer.get_transformed()
        return X, Y

    def init(self, X, Y, Y_dir=None):
        self._X, self._Y, self.Y_dir = X, Y, Y_dir
        self.model.init(*self._normalize(X, Y), Y_dir)

    def add_observations(self, X_new, Y_new, Y_dir_new=None):
        assert self._X is not None, "Call init first"
        Y_dir = None if self.Y_dir is None else np.concatenate([self.Y_dir, Y_dir_new])
        self.init(np.concatenate([self._X, X_new]), np.concatenate([self._Y, Y_new]), Y_dir)

    def set_train_data(self, X, Y):
        self._X, self._Y = X, Y
        self.model.set_train_data(*self._normalize(X, Y))

    # -- prediction --------------------------------------------------------------------------------------------
    def get_statistics(self, X, full_cov=True):
        Xn = self.X_normalizer.normalize(X) if self._normalize_input else X
        mean, spread = self.model.get_statistics(Xn, full_cov=full_cov)
        if not self._normalize_output:
            return mean, spread
        scale_back = self.Y_normalizer.denormalize_covariance if full_cov else self.Y_normalizer.denormalize_variance
        return self.Y_normalizer.denormalize(mean), scale_back(spread)

    # -- persistence: the pickle holds only the flags; the wrapped model and the raw data sit beside it and the
    #    normalizers are rebuilt on load (same on-disk layout as the reference: model.pickle, X.npy, Y.npy, model/) --------
    def __getstate__(self):
        return {k: (None if k in self._TRANSIENT else v) for k, v in self.__dict__.items()}

    def save(self, PATH):
        super().save(PATH)
        if self._X is not None:
            for name, arr in (("X.npy", self._X), ("Y.npy", self._Y)):
                np.save(os.path.join(PATH, name), arr)
        self.model.save(os.path.join(PATH, "model"))

    def post_load_hook(self, PATH):
        self.model = SaveMixin.load(os.path.join(PATH, "model"))
        files = [os.path.join(PATH, n) for n in ("X.npy", "Y.npy")]
        if all(os.path.isfile(f) for f in files):
            self._X, self._Y = (np.load(f) for f in files)
            self._normalize(self._X, self._Y)
        else:
            self._X = self._Y = None
        return self
