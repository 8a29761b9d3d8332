"""Input/output z-scoring wrapper with the reference's interface (``src/models/normalizer.py:9-158``): nearly every thesis
config wraps GPModel in ``NormalizerModel`` (``notebook_header.py:17-26``).  The affine maps cost O(N d) on the host, which is
noise next to the O(N^3) factorisation, so they stay in numpy; the wrapped model runs on the GPU."""
import os

import numpy as np

from .config_helpers import ConfigMixin, construct_from_module
from .core_models import BaseModel, SaveMixin


def zero_mean_unit_var_normalization(X, mean=None, std=None):
    mean = np.mean(X, axis=0) if mean is None else mean
    std = np.std(X, axis=0) if std is None else std
    return (X - mean) / std, mean, std


def zero_mean_unit_var_unnormalization(X_normalized, mean, std):
    return X_normalized * std + mean


class Normalizer:
    """Holds the column means / standard deviations of the data it was built from."""

    def __init__(self, X_train):
        self.X, self._X_mean, self._X_std = zero_mean_unit_var_normalization(X_train)

    def get_transformed(self):
        return self.X

    def normalize(self, X):
        return zero_mean_unit_var_normalization(X, mean=self._X_mean, std=self._X_std)[0]

    def denormalize(self, X):
        return zero_mean_unit_var_unnormalization(X, self._X_mean, self._X_std)

    def denormalize_variance(self, X_var):
        return X_var * (self._X_std ** 2)

    def denormalize_covariance(self, covariance):
        # same arithmetic as the reference (normalizer.py:31-32): a trailing axis is appended before scaling
        return covariance[..., None] * (self._X_std ** 2)


class NormalizerModel(ConfigMixin, BaseModel):
    def __init__(self, model, normalize_input=True, normalize_output=True):
        super().__init__()
        self.model = model
        self.X_normalizer = None
        self.Y_normalizer = None
        self._normalize_input = normalize_input
        self._normalize_output = normalize_output

    @classmethod
    def process_config(cls, *, model=None, **kwargs):
        import thesis_code_b200 as models_module
        return dict(model=construct_from_module(models_module, model), **kwargs)

    def _normalize(self, X, Y):
        if self._normalize_input:
            self.X_normalizer = Normalizer(X)
            X = self.X_normalizer.get_transformed()
        if self._normalize_output:
            self.Y_normalizer = Normalizer(Y)
            Y = self.Y_normalizer.get_transformed()
        return X, Y

    def init(self, X, Y, Y_dir=None):
        self._X, self._Y, self.Y_dir = X, Y, Y_dir
        Xn, Yn = self._normalize(X, Y)
        self.model.init(Xn, Yn, Y_dir)

    def add_observations(self, X_new, Y_new, Y_dir_new=None):
        assert self._X is not None, "Call init first"
        X = np.concatenate([self._X, X_new])
        Y = np.concatenate([self._Y, Y_new])
        Y_dir = np.concatenate([self.Y_dir, Y_dir_new]) if self.Y_dir is not None else None
        self.init(X, Y, Y_dir)

    def get_statistics(self, X, full_cov=True):
        if self._normalize_input:
            X = self.X_normalizer.normalize(X)
        mean, covar = self.model.get_statistics(X, full_cov=full_cov)
        if self._normalize_output:
            mean = self.Y_normalizer.denormalize(mean)
            covar = (self.Y_normalizer.denormalize_covariance(covar) if full_cov
                     else self.Y_normalizer.denormalize_variance(covar))
        return mean, covar

    def set_train_data(self, X, Y):
        self._X, self._Y = X, Y
        Xn, Yn = self._normalize(X, Y)
        self.model.set_train_data(Xn, Yn)

    # -- persistence: the wrapped model and the raw data are stored beside the pickle, normalizers are rebuilt ----------
    def __getstate__(self):
        state = self.__dict__.copy()
        for k in ("model", "_X", "_Y", "X_normalizer", "Y_normalizer"):
            state[k] = None
        return state

    def save(self, PATH):
        super().save(PATH)
        if self._X is not None:
            np.save(os.path.join(PATH, 'X.npy'), self._X)
            np.save(os.path.join(PATH, 'Y.npy'), self._Y)
        self.model.save(os.path.join(PATH, 'model'))

    def post_load_hook(self, PATH):
        self.model = SaveMixin.load(os.path.join(PATH, 'model'))
        X_path, Y_path = os.path.join(PATH, 'X.npy'), os.path.join(PATH, 'Y.npy')
        if os.path.isfile(X_path):
            self._X, self._Y = np.load(X_path), np.load(Y_path)
            self._normalize(self._X, self._Y)
        else:
            self._X = self._Y = None
        return self
