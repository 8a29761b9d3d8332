"""Loads the reference's OWN exact-GP code, unmodified, from /root/reference so that it can be executed here.

THIS IS TEST INFRASTRUCTURE (oracle/).  It only works in the build container -- /root/reference does not exist on the GPU
box -- and is used by tests/golden/make_golden_reference.py to generate the committed fixtures that pin the oracle and the
CUDA path to reference code that actually ran (tests/golden/reference_pinned.npz).

What is loaded (the files are exec'd from where they lie; nothing is copied):
  * /root/reference/src/experiment/config_helpers.py      -> ``src.experiment.config_helpers`` (pure Python)
  * /root/reference/src/models/core_models.py             -> ``src.models`` (for ``SaveMixin`` / ``BaseModel`` / ``ProbModel``;
        its ``import GPy`` is satisfied by an empty stand-in module: GPy is not installable here, and none of the classes
        used below touches it)
  * /root/reference/src/models/deprecated_models.py:8-48   -> ``GPVanillaModel``, the author's numpy/scipy exact GP
  * /root/reference/src/models/normalizer.py               -> ``Normalizer`` / ``NormalizerModel``
The regular package import (``import src.models``) cannot be used: src/models/__init__.py pulls in gpytorch, GPy, sacred...
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("GPX_REFERENCE_ROOT", "/root/reference")


class _Anything(types.ModuleType):
    """Stand-in for an absent third-party module: any attribute is another stand-in (``GPy.kern.RBF`` ...)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Anything(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        raise RuntimeError("{} is a stand-in: the real library is not installed here".format(self.__name__))


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "models", "deprecated_models.py"))


def _exec_file(mod, relpath):
    path = os.path.join(REFERENCE_ROOT, relpath)
    with open(path) as f:
        code = compile(f.read(), path, "exec")
    mod.__file__ = path
    exec(code, mod.__dict__)
    return mod


def load():
    """Returns a namespace with the reference's GPVanillaModel, Normalizer, NormalizerModel, BaseModel, ProbModel, SaveMixin,
    ConfigMixin -- executed from the reference's own source files."""
    if not available():
        raise RuntimeError("reference sources not found under " + REFERENCE_ROOT)
    saved = {k: sys.modules.get(k) for k in ("GPy", "src", "src.experiment", "src.experiment.config_helpers", "src.kernels",
                                            "src.models", "src.models.deprecated_models", "src.models.normalizer")}
    try:
        sys.modules["GPy"] = _Anything("GPy")
        src = types.ModuleType("src"); src.__path__ = []
        exp = types.ModuleType("src.experiment"); exp.__path__ = []
        sys.modules["src"], sys.modules["src.experiment"] = src, exp
        ch = _exec_file(types.ModuleType("src.experiment.config_helpers"), "src/experiment/config_helpers.py")
        sys.modules["src.experiment.config_helpers"] = ch
        exp.config_helpers = ch
        kern = _exec_file(types.ModuleType("src.kernels"), "src/kernels.py")      # `GPyRBF = GPy.kern.RBF` -> stand-ins
        sys.modules["src.kernels"] = kern
        models = types.ModuleType("src.models"); models.__path__ = []
        sys.modules["src.models"] = models
        _exec_file(models, "src/models/core_models.py")                           # defines SaveMixin, BaseModel, ProbModel, GPModel...
        dep = _exec_file(types.ModuleType("src.models.deprecated_models"), "src/models/deprecated_models.py")
        nrm = _exec_file(types.ModuleType("src.models.normalizer"), "src/models/normalizer.py")
        ns = types.SimpleNamespace(GPVanillaModel=dep.GPVanillaModel, Normalizer=nrm.Normalizer, NormalizerModel=nrm.NormalizerModel,
                                   BaseModel=models.BaseModel, ProbModel=models.ProbModel, SaveMixin=models.SaveMixin,
                                   ConfigMixin=ch.ConfigMixin, zero_mean_unit_var_normalization=nrm.zero_mean_unit_var_normalization)
        return ns
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
