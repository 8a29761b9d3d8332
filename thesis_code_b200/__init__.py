"""thesis_code_b200: the exact-GP hot path of tmpethick/thesis-code's GPModel as sm_100a CUDA behind the reference's
own Python model API.  ``from thesis_code_b200 import GPModel`` is the drop-in for ``src.models.GPModel``."""
from .config_helpers import ConfigMixin, LazyConstructor, construct_from_module, lazy_construct_from_module  # noqa: F401
from .kernels import GPyExponential, GPyMatern32, GPyMatern52, GPyRBF  # noqa: F401
from .core_models import (BaseModel, DerivativeGPModel, GPModel, MarginalLogLikelihoodMixin, ProbModel,  # noqa: F401
                          SaveMixin)

from .normalizer import Normalizer, NormalizerModel  # noqa: F401,E402
from . import priors  # noqa: F401,E402

__all__ = ["Normalizer", "NormalizerModel", "GPModel", "DerivativeGPModel", "BaseModel", "ProbModel", "SaveMixin", "MarginalLogLikelihoodMixin",
           "GPyRBF", "GPyMatern52", "GPyMatern32", "GPyExponential", "ConfigMixin", "LazyConstructor",
           "construct_from_module", "lazy_construct_from_module"]
