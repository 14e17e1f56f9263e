"""Model building blocks with the names and semantics the reference takes from GPyTorch.

The reference builds its model from third-party GPyTorch objects (cglb/backend/pytorch/interface.py:207-301,
models.py:38-47): `GaussianLikelihood(noise_constraint=GreaterThan(1e-6))`, `ConstantMean`,
`ScaleKernel(MaternKernel(nu=1.5, ard_num_dims=d) | RBFKernel(ard_num_dims=d))`, `InducingPointKernel`.
These classes keep the same attribute names, parameter shapes, softplus parameterisation and
`model.parameters()` order (raw noise, mean constant, inducing points, raw outputscale, raw
lengthscale), so the reference's optimiser wrapper and `model_parameters` keep working unchanged.
Kernel evaluation itself never happens in torch: `kernel(x1, x2)` returns a lazy operator whose
products run in the sm_100a kernels (operators.py).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

Tensor = torch.Tensor


def _inv_softplus(y: Tensor) -> Tensor:
    return y + torch.log(-torch.expm1(-y))


class GreaterThan:
    """gpytorch.constraints.GreaterThan: value = softplus(raw) + lower_bound."""

    def __init__(self, lower_bound: float):
        self.lower_bound = float(lower_bound)

    def transform(self, raw: Tensor) -> Tensor:
        return torch.nn.functional.softplus(raw) + self.lower_bound

    def inverse_transform(self, value: Tensor) -> Tensor:
        return _inv_softplus(value - self.lower_bound)


class Positive(GreaterThan):
    def __init__(self):
        super().__init__(0.0)


class Kernel(nn.Module):
    """Callable like a gpytorch kernel: kernel(x1, x2=None, diag=False) -> lazy operator / diagonal."""

    def __call__(self, x1: Tensor, x2: Optional[Tensor] = None, diag: bool = False, **kwargs):
        from .operators import KernelOperator
        if diag:
            return self.diag(x1)
        return KernelOperator(self, x1, x1 if x2 is None else x2, symmetric=x2 is None or x2 is x1)

    def diag(self, x: Tensor) -> Tensor:
        raise NotImplementedError


class _StationaryKernel(Kernel):
    kind: str = ""

    def __init__(self, ard_num_dims: Optional[int] = None, **kwargs):
        super().__init__()
        d = 1 if ard_num_dims is None else int(ard_num_dims)
        self.ard_num_dims = ard_num_dims
        self.register_parameter("raw_lengthscale", nn.Parameter(torch.zeros(1, d)))
        self.raw_lengthscale_constraint = Positive()

    @property
    def lengthscale(self) -> Tensor:
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale)

    @lengthscale.setter
    def lengthscale(self, value):
        value = torch.as_tensor(value, dtype=self.raw_lengthscale.dtype, device=self.raw_lengthscale.device).reshape(1, -1)
        value = value.expand_as(self.raw_lengthscale)
        self.raw_lengthscale.data = self.raw_lengthscale_constraint.inverse_transform(value).clone()

    def diag(self, x: Tensor) -> Tensor:
        return torch.ones(x.shape[0], dtype=x.dtype, device=x.device)


class MaternKernel(_StationaryKernel):
    """nu = 1.5 only (the reference never builds another one, interface.py:224)."""
    kind = "matern32"

    def __init__(self, nu: float = 1.5, **kwargs):
        if nu != 1.5:
            raise NotImplementedError("cglb_b200 implements the Matern kernel for nu=1.5 (Matern32) only")
        super().__init__(**kwargs)
        self.nu = nu


class RBFKernel(_StationaryKernel):
    kind = "rbf"


class ScaleKernel(Kernel):
    def __init__(self, base_kernel: Kernel, **kwargs):
        super().__init__()
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", nn.Parameter(torch.zeros(())))
        self.raw_outputscale_constraint = Positive()

    @property
    def outputscale(self) -> Tensor:
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    @outputscale.setter
    def outputscale(self, value):
        value = torch.as_tensor(value, dtype=self.raw_outputscale.dtype, device=self.raw_outputscale.device).reshape(())
        self.raw_outputscale.data = self.raw_outputscale_constraint.inverse_transform(value)

    @property
    def kind(self) -> str:
        return self.base_kernel.kind

    @property
    def lengthscale(self) -> Tensor:
        return self.base_kernel.lengthscale

    def diag(self, x: Tensor) -> Tensor:
        return self.outputscale * self.base_kernel.diag(x)


class InducingPointKernel(Kernel):
    """Parameter holder, as on the reference's path (models.py:139-145 only reads .base_kernel and
    .inducing_points)."""

    def __init__(self, base_kernel: Kernel, inducing_points: Tensor, likelihood: "GaussianLikelihood", **kwargs):
        super().__init__()
        self.base_kernel = base_kernel
        self.likelihood = likelihood
        if inducing_points.ndim == 1:
            inducing_points = inducing_points.unsqueeze(-1)
        self.register_parameter("inducing_points", nn.Parameter(inducing_points.detach().clone()))

    def diag(self, x: Tensor) -> Tensor:
        return self.base_kernel.diag(x)


class _HomoskedasticNoise(nn.Module):
    def __init__(self, noise_constraint: GreaterThan):
        super().__init__()
        self.register_parameter("raw_noise", nn.Parameter(torch.zeros(1)))
        self.raw_noise_constraint = noise_constraint

    @property
    def noise(self) -> Tensor:
        return self.raw_noise_constraint.transform(self.raw_noise)

    @noise.setter
    def noise(self, value):
        value = torch.as_tensor(value, dtype=self.raw_noise.dtype, device=self.raw_noise.device).reshape(1)
        self.raw_noise.data = self.raw_noise_constraint.inverse_transform(value)


class GaussianLikelihood(nn.Module):
    def __init__(self, noise_constraint: Optional[GreaterThan] = None, **kwargs):
        super().__init__()
        self.noise_covar = _HomoskedasticNoise(noise_constraint if noise_constraint is not None else GreaterThan(1e-4))

    @property
    def noise(self) -> Tensor:
        return self.noise_covar.noise

    @noise.setter
    def noise(self, value):
        self.noise_covar.noise = value


class ConstantMean(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_parameter("constant", nn.Parameter(torch.zeros(1)))

    def forward(self, x: Tensor) -> Tensor:
        return self.constant.expand(x.shape[:-1])
