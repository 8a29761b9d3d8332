"""``NormalizerModel`` with the z-scoring FUSED into the GPU path.

Interface of the reference's ``src/models/normalizer.py:51-158`` (nearly every thesis config wraps ``GPModel`` in it,
``notebook_header.py:17-26``): ``NormalizerModel(model, normalize_input=True, normalize_output=True)``, ``init``,
``add_observations``, ``get_statistics``, ``save`` / ``load``.  The reference materialises normalised copies of X and Y
with numpy and hands those to the wrapped model.  Here no normalised copy exists on the host:

* ``init`` hands the RAW arrays to the wrapped ``GPModel`` and switches the handle's normaliser on
  (``gpx_set_normalizer``): ``gpx_fit`` computes the column means / population standard deviations on the device and
  applies ``(x - mean) / std`` inside the K-build's input loader and ``(y - mean) / std`` inside the residual kernel;
* ``get_statistics`` hands the RAW test inputs over; the predict loader z-scores them and the reduce epilogues return
  ``mean * std_y + mean_y`` and ``var * std_y**2`` (reference ``denormalize`` / ``denormalize_variance``, :24-33);
* this class keeps the flags, the pickle layout (``model.pickle``, ``X.npy``, ``Y.npy``, ``model/``) and, for callers that
  ask, ``X_normalizer`` / ``Y_normalizer`` views built from the statistics read back from the device.

Kept on purpose (reference behaviour): population std (``np.std``), and a full covariance gets a TRAILING axis before
being scaled by ``std**2`` (``denormalize_covariance``, :32-33) -- that one O(N*^2) rescale stays on the host.
"""
import os

import numpy as np

from .config_helpers import ConfigMixin, construct_from_module
from .core_models import BaseModel, SaveMixin


class Normalizer:
    """The reference's helper object (normalizer.py:9-33) as a view on statistics that were computed on the device."""

    def __init__(self, mean, std):
        self._X_mean = np.asarray(mean, dtype=np.float64)
        self._X_std = np.asarray(std, dtype=np.float64)

    def normalize(self, X):
        return (X - self._X_mean) / self._X_std

    def denormalize(self, X):
        return X * self._X_std + self._X_mean

    def denormalize_variance(self, X_var):
        return X_var * (self._X_std ** 2)

    def denormalize_covariance(self, covariance):
        return covariance[..., None] * (self._X_std ** 2)


class NormalizerModel(ConfigMixin, BaseModel):
    def __init__(self, model, normalize_input=True, normalize_output=True):
        super().__init__()
        if not hasattr(model, "_normalize_input"):
            raise TypeError("NormalizerModel fuses the z-scoring into the GPU path of the wrapped model; "
                            "{} does not support that".format(type(model).__name__))
        self.model = model
        self._normalize_input = bool(normalize_input)
        self._normalize_output = bool(normalize_output)
        self._push_flags()

    @classmethod
    def process_config(cls, *, model=None, **kwargs):
        import thesis_code_b200 as models_module
        return dict(model=construct_from_module(models_module, model), **kwargs)

    def _push_flags(self):
        m = self.model
        if m is None:
            return
        changed = (m._normalize_input, m._normalize_output) != (self._normalize_input, self._normalize_output)
        m._normalize_input, m._normalize_output = self._normalize_input, self._normalize_output
        if m._handle is not None:
            m._handle.set_normalizer(self._normalize_input, self._normalize_output)
            if changed:
                m._fitted_theta = None

    # statistics of the last fit, read back from the device (d + O numbers)
    def _stats(self):
        h = self.model._get_handle()
        return h.get_normalizer(self._X.shape[1], self._Y.shape[1])

    @property
    def X_normalizer(self):
        if not self._normalize_input or self._X is None:
            return None
        xm, xs, _, _ = self._stats()
        return Normalizer(xm, xs)

    @property
    def Y_normalizer(self):
        if not self._normalize_output or self._Y is None:
            return None
        _, _, ym, ys = self._stats()
        return Normalizer(ym, ys)

    def init(self, X, Y, Y_dir=None):
        self._X, self._Y, self.Y_dir = X, Y, Y_dir
        self._push_flags()
        self.model.init(X, Y, Y_dir)          # raw data: the z-scores happen inside gpx_fit

    def add_observations(self, X_new, Y_new, Y_dir_new=None):
        assert self._X is not None, "Call init first"
        Y_dir = None if self.Y_dir is None else np.concatenate([self.Y_dir, Y_dir_new])
        # the statistics change with the data, so this is a full refit, as in the reference (normalizer.py:88-99)
        self.init(np.concatenate([self._X, X_new]), np.concatenate([self._Y, Y_new]), Y_dir)

    def get_statistics(self, X, full_cov=True):
        mean, spread = self.model.get_statistics(X, full_cov=full_cov)   # mean / variance leave the GPU de-normalised
        if full_cov and self._normalize_output:
            spread = self.Y_normalizer.denormalize_covariance(spread)   # predict_cov returns the normalised covariance
        return mean, spread

    def get_mean(self, X):
        return self.model.get_mean(X)

    def set_train_data(self, X, Y):
        self._X, self._Y = X, Y
        self.model.set_train_data(X, Y)

    # -- persistence: same on-disk layout as the reference (model.pickle, X.npy, Y.npy, model/) ---------------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        state["model"] = None
        state["_X"] = None
        state["_Y"] = None
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
        else:
            self._X = self._Y = None
        self._push_flags()
        return self
