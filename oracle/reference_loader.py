"""Load the reference's own hot-path files verbatim (TEST INFRASTRUCTURE; build container only).

/root/reference does not exist on the GPU box, so nothing that runs there may import this module;
it is used by make_golden.py and by the CPU-only tests that cross-check the restated oracle against
the reference code (skipped when /root/reference is absent).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CGLB_REFERENCE_ROOT", "/root/reference")
_PKG = "cglb.backend.pytorch"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "cglb/backend/pytorch/conjugate_gradient.py"))


def _fake_package(name: str, path: str):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = [path]          # a package, but its __init__.py (which imports interface.py) is NOT run
    sys.modules[name] = m
    return m


def _load(modname: str, relpath: str):
    full = f"{_PKG}.{modname}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


def load(jitter: float = 1e-6):
    """Returns (conjugate_gradient, models, optimizer) modules of the reference, executed from
    /root/reference on top of oracle.gpytorch_stub."""
    if not available():
        raise RuntimeError("reference sources not present at " + REFERENCE_ROOT)
    from . import gpytorch_stub
    gpytorch_stub.install(jitter)
    _fake_package("cglb", os.path.join(REFERENCE_ROOT, "cglb"))
    _fake_package("cglb.backend", os.path.join(REFERENCE_ROOT, "cglb/backend"))
    _fake_package(_PKG, os.path.join(REFERENCE_ROOT, "cglb/backend/pytorch"))
    cg = _load("conjugate_gradient", "cglb/backend/pytorch/conjugate_gradient.py")
    models = _load("models", "cglb/backend/pytorch/models.py")
    opt = _load("optimizer", "cglb/backend/pytorch/optimizer.py")
    return cg, models, opt


def build_reference_model(models, kind: str, x, y, z, noise, variance, lengthscale, mean_c=0.0):
    """What interface.py:263-323 does for CGLBConfig, with the stub classes (fp64, CPU)."""
    import torch
    from . import gpytorch_stub as gp
    d = x.shape[1]
    lik = gp.GaussianLikelihood(noise_constraint=gp.GreaterThan(1e-6)).double()
    lik.noise = noise
    base = (gp.MaternKernel(nu=1.5, ard_num_dims=d) if kind == "matern32" else gp.RBFKernel(ard_num_dims=d)).double()
    ls = torch.as_tensor(lengthscale, dtype=torch.float64).reshape(-1)
    base.lengthscale = ls if ls.numel() == d else ls.repeat(d)
    scale = gp.ScaleKernel(base).double()
    scale.outputscale = variance
    ipk = gp.InducingPointKernel(scale, torch.as_tensor(z, dtype=torch.float64), likelihood=lik)
    model = models.CGLB((x, y.reshape(-1)), lik, ipk)
    model.mean_module = model.mean_module.double()
    model.mean_module.constant.data.fill_(mean_c)
    return model
