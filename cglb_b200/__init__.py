"""cglb_b200 -- B200-native (sm_100a) implementation of the CGLB hot path behind the reference's
model/objective API (awav/CGLB, cglb/backend/pytorch).  See DESIGN.md and INTEGRATION.md.

Importing the package is cheap and works without a GPU; every compute entry point raises `CglbError` if
the shared library (cglb_b200/lib/libcglb_b200.so) or a CUDA device is missing -- there is no CPU path.
"""
from ._ffi import CglbError, LIB_PATH, load_library
from .conjugate_gradient import ConjugateGradient, ConjugateGradientStats, NystromPreconditioner
from .config import (CGLBConfig, InducingVariableConfig, Matern32Config, SquaredExponentialConfig,
                     INDUCING_VARIABLE_CONFIGS, KERNEL_CONFIGS, SGPR_CONFIGS)
from .distributed import Shard
from .gp import (ConstantMean, GaussianLikelihood, GreaterThan, InducingPointKernel, MaternKernel, RBFKernel,
                 ScaleKernel)
from .models import CGLB, GPR, SGPR, Bounds, CommonTerms, LowerBoundCG, PredictCG, gaussian, log_density
from .optimizer import Scipy
from .backend import BACKENDS, B200, Backend

__all__ = [n for n in dir() if not n.startswith("_")]
