"""Minimal stand-ins for the third-party symbols the reference imports (TEST INFRASTRUCTURE).

gpytorch and pykeops are un-pinned dependencies of the reference (requirements.txt:2,12) and are not
installed in this image.  To execute the reference's *own* files verbatim
(cglb/backend/pytorch/{conjugate_gradient,models,optimizer}.py) we provide just the API surface those
files touch: ``gpytorch.models.ExactGP``, ``gpytorch.means.ConstantMean``, ``gpytorch.kernels.{MaternKernel,
RBFKernel,ScaleKernel,InducingPointKernel}``, ``gpytorch.likelihoods.GaussianLikelihood``,
``gpytorch.constraints.GreaterThan``, ``gpytorch.mlls.ExactMarginalLogLikelihood``,
``gpytorch.settings.cholesky_jitter``, ``gpytorch.delazify``, ``gpytorch.lazy.LazyTensor``,
``pykeops.torch.LazyTensor``.

The kernel arithmetic is the published gpytorch definition restated from memory:
Matern nu=1.5: (1 + sqrt3 r) exp(-sqrt3 r), RBF: exp(-r^2/2), r = ||(x-x')/l|| with ARD lengthscales,
ScaleKernel multiplies by softplus(raw_outputscale).  The distance uses direct differences (the
gpytorch.kernels.keops variant the reference selects with --keops, interface.py:695-702).
"""
from __future__ import annotations

import math
import sys
import types

import torch
from torch import nn

_SQRT3 = math.sqrt(3.0)


def _inv_softplus(y):
    return y + torch.log(-torch.expm1(-y))


class _DenseLazy:
    """The slice of gpytorch.lazy.LazyTensor the reference uses: add_diag, detach, @, evaluate."""

    def __init__(self, dense: torch.Tensor):
        self._dense = dense

    def add_diag(self, value):
        n = self._dense.shape[-1]
        return _DenseLazy(self._dense + value * torch.eye(n, dtype=self._dense.dtype, device=self._dense.device))

    def detach(self):
        return _DenseLazy(self._dense.detach())

    def evaluate(self):
        return self._dense

    def __matmul__(self, other):
        return self._dense @ other

    @property
    def shape(self):
        return self._dense.shape


def delazify(obj):
    return obj.evaluate() if isinstance(obj, _DenseLazy) else obj


class GreaterThan:
    def __init__(self, lower_bound):
        self.lower_bound = float(lower_bound)

    def transform(self, raw):
        return torch.nn.functional.softplus(raw) + self.lower_bound

    def inverse_transform(self, value):
        return _inv_softplus(value - self.lower_bound)


class Positive(GreaterThan):
    def __init__(self):
        super().__init__(0.0)


class _Kernel(nn.Module):
    def __call__(self, x1, x2=None, diag=False, **kw):
        if x2 is None:
            x2 = x1
        if diag:
            return self.forward_diag(x1)
        return _DenseLazy(self.forward(x1, x2))


class _Stationary(_Kernel):
    def __init__(self, ard_num_dims=None, **kw):
        super().__init__()
        d = 1 if ard_num_dims is None else ard_num_dims
        self.raw_lengthscale = nn.Parameter(torch.zeros(1, d))
        self._c = Positive()

    @property
    def lengthscale(self):
        return self._c.transform(self.raw_lengthscale)

    @lengthscale.setter
    def lengthscale(self, value):
        value = torch.as_tensor(value, dtype=self.raw_lengthscale.dtype).reshape(1, -1)
        self.raw_lengthscale.data = self._c.inverse_transform(value).expand_as(self.raw_lengthscale).clone()

    def _sqdist(self, x1, x2):
        a = x1 / self.lengthscale
        b = x2 / self.lengthscale
        diff = a[:, None, :] - b[None, :, :]
        return (diff * diff).sum(-1)

    def forward_diag(self, x):
        return torch.ones(x.shape[0], dtype=x.dtype, device=x.device)


class MaternKernel(_Stationary):
    def __init__(self, nu=2.5, **kw):
        super().__init__(**kw)
        assert nu == 1.5, "the reference only builds nu=1.5 (interface.py:224)"

    def forward(self, x1, x2):
        r = torch.sqrt(self._sqdist(x1, x2).clamp_min(1e-30))
        s = _SQRT3 * r
        return (1.0 + s) * torch.exp(-s)


class RBFKernel(_Stationary):
    def forward(self, x1, x2):
        return torch.exp(-0.5 * self._sqdist(x1, x2))


class ScaleKernel(_Kernel):
    def __init__(self, base_kernel, **kw):
        super().__init__()
        self.base_kernel = base_kernel
        self.raw_outputscale = nn.Parameter(torch.zeros(()))
        self._c = Positive()

    @property
    def outputscale(self):
        return self._c.transform(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        value = torch.as_tensor(value, dtype=self.raw_outputscale.dtype).reshape(())
        self.raw_outputscale.data = self._c.inverse_transform(value)

    def forward(self, x1, x2):
        return self.outputscale * self.base_kernel.forward(x1, x2)

    def forward_diag(self, x):
        return self.outputscale * self.base_kernel.forward_diag(x)


class InducingPointKernel(_Kernel):
    def __init__(self, base_kernel, inducing_points, likelihood, **kw):
        super().__init__()
        self.base_kernel = base_kernel
        self.likelihood = likelihood
        if inducing_points.ndim == 1:
            inducing_points = inducing_points.unsqueeze(-1)
        self.inducing_points = nn.Parameter(inducing_points.clone())


class _HomoskedasticNoise(nn.Module):
    def __init__(self, constraint):
        super().__init__()
        self.raw_noise = nn.Parameter(torch.zeros(1))
        self._c = constraint

    @property
    def noise(self):
        return self._c.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        value = torch.as_tensor(value, dtype=self.raw_noise.dtype).reshape(1)
        self.raw_noise.data = self._c.inverse_transform(value)


class GaussianLikelihood(nn.Module):
    def __init__(self, noise_constraint=None, **kw):
        super().__init__()
        self.noise_covar = _HomoskedasticNoise(noise_constraint or GreaterThan(1e-4))

    @property
    def noise(self):
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value


class ConstantMean(nn.Module):
    def __init__(self):
        super().__init__()
        self.constant = nn.Parameter(torch.zeros(1))

    def forward(self, x):
        return self.constant.expand(x.shape[:-1])


class ExactGP(nn.Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        self.likelihood = likelihood
        self.train_inputs = (train_inputs,) if torch.is_tensor(train_inputs) else tuple(train_inputs)
        self.train_targets = train_targets


class ExactMarginalLogLikelihood(nn.Module):
    def __init__(self, likelihood, model):
        super().__init__()
        # the reference re-exposes these through properties (models.py:135-145); avoid nn.Module
        # attribute registration clashing with the read-only ``likelihood`` property
        object.__setattr__(self, "model", model)


class _Setting:
    def __init__(self, value):
        self._v = value

    def value(self):
        return self._v

    def _set_value(self, v):
        self._v = v


def install(jitter: float = 1e-6):
    """Register the stub modules in sys.modules (idempotent).  Returns the fake ``gpytorch``."""
    if "gpytorch" in sys.modules and getattr(sys.modules["gpytorch"], "__cglb_stub__", False):
        sys.modules["gpytorch"].settings.cholesky_jitter._set_value(jitter)
        return sys.modules["gpytorch"]
    g = types.ModuleType("gpytorch")
    g.__cglb_stub__ = True

    def sub(name, **attrs):
        m = types.ModuleType(f"gpytorch.{name}")
        for k, v in attrs.items():
            setattr(m, k, v)
        setattr(g, name, m)
        sys.modules[f"gpytorch.{name}"] = m
        return m

    sub("models", ExactGP=ExactGP)
    sub("means", ConstantMean=ConstantMean, Mean=ConstantMean)
    sub("kernels", MaternKernel=MaternKernel, RBFKernel=RBFKernel, ScaleKernel=ScaleKernel,
        InducingPointKernel=InducingPointKernel, Kernel=_Kernel)
    sub("likelihoods", GaussianLikelihood=GaussianLikelihood, Likelihood=GaussianLikelihood)
    sub("constraints", GreaterThan=GreaterThan, Positive=Positive)
    sub("mlls", ExactMarginalLogLikelihood=ExactMarginalLogLikelihood)
    sub("lazy", LazyTensor=_DenseLazy)
    sub("distributions", MultivariateNormal=object)
    sub("settings", cholesky_jitter=_Setting(jitter))
    g.delazify = delazify
    sys.modules["gpytorch"] = g

    pk = types.ModuleType("pykeops")
    pkt = types.ModuleType("pykeops.torch")
    pkt.LazyTensor = _DenseLazy
    pk.torch = pkt
    sys.modules.setdefault("pykeops", pk)
    sys.modules.setdefault("pykeops.torch", pkt)
    return g
