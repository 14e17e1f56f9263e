"""CGLB model and objective with the reference's names, signatures and error behaviour
(cglb/backend/pytorch/models.py), computed by the sm_100a kernels through `BoundEvaluator`.

    model = CGLB((x, y), likelihood, kernel)
    loss = -LowerBoundCG(model)((x, y))
    grads = torch.autograd.grad(loss, model.parameters())        # works unchanged (optimizer.py:95-98)

The bound is a single `torch.autograd.Function` node over the constrained hyper-parameters; its backward
returns the closed-form gradients evaluated in the same device pass (bound.py), so the chain rule through
the softplus parameterisation is the only thing autograd does.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple, Union

import numpy as np
import torch
from torch import nn

from . import settings
from ._ffi import CglbError
from .bound import BoundEvaluator
from .conjugate_gradient import ConjugateGradient, ConjugateGradientStats, NystromPreconditioner
from .distributed import Shard
from .gp import ConstantMean, GaussianLikelihood, InducingPointKernel, Kernel
from .operators import KernelOperator, _kernel_pieces, delazify

Tensor = torch.Tensor
GenericTensor = Union[np.ndarray, Tensor]
Data = Tuple[GenericTensor, GenericTensor]


class GPR(nn.Module):
    """Stand-in for the reference's `GPR(gpytorch.models.ExactGP)` (models.py:38-47): a state holder with
    train_inputs / train_targets / likelihood / mean_module / covar_module."""

    def __init__(self, data: Data, likelihood: GaussianLikelihood, kernel: Kernel):
        super().__init__()
        self.likelihood = likelihood
        x, y = data
        self.train_inputs = (x,)
        self.train_targets = y
        self.mean_module = ConstantMean()
        self.covar_module = kernel

    def set_train_data(self, inputs=None, targets=None, strict: bool = True):
        if inputs is not None:
            self.train_inputs = (inputs,) if torch.is_tensor(inputs) else tuple(inputs)
        if targets is not None:
            self.train_targets = targets

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.train_inputs = tuple(fn(t) for t in self.train_inputs)
        self.train_targets = fn(self.train_targets)
        for name in ("_v_vec",):
            if hasattr(self, name):
                setattr(self, name, fn(getattr(self, name)))
        return out


class SGPR(GPR):
    ...


class CGLB(SGPR):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._v_vec = self._build_v_vec()

    def _build_v_vec(self) -> Tensor:                      # models.py:59-68
        num_data = self.train_targets.shape[0]
        return torch.zeros((int(num_data), 1), dtype=self.train_targets.dtype, device=self.train_targets.device,
                           requires_grad=False)

    @property
    def v_vec(self) -> Tensor:
        return self._v_vec

    @property
    def cg_stats(self) -> Optional[ConjugateGradientStats]:
        return getattr(self, "_cg_stats", None)

    @cg_stats.setter
    def cg_stats(self, value: ConjugateGradientStats):    # models.py:80-87
        steps, error = value.steps, value.residual_error
        if isinstance(steps, torch.Tensor):
            steps = steps.detach().cpu().numpy()
        if isinstance(error, torch.Tensor):
            error = error.detach().cpu().numpy()
        self._cg_stats = ConjugateGradientStats(steps, error)


@dataclass
class CommonTerms:              # models.py:90-95
    A: Tensor
    LB: Tensor
    AAt_diag_sum: Tensor
    L: Tensor


@dataclass
class Bounds:                   # models.py:98-101
    upper_bound: Tensor
    lower_bound: Tensor


class _BoundFunction(torch.autograd.Function):
    """bound(noise, mean constant, Z, outputscale, lengthscale): value and closed-form gradients from one
    device pass (replaces the autograd tape through KeOps / cuBLAS / cuSOLVER of the reference)."""

    @staticmethod
    def forward(ctx, noise, mean_c, Z, variance, lengthscale, objective, data):
        need_grad = any(ctx.needs_input_grad[:5])
        out = objective._evaluate(data, need_grad)
        dtype, device = noise.dtype, noise.device
        if need_grad:
            g = out.grads
            ctx.save_for_backward(g["noise"].reshape(noise.shape).to(dtype), g["mean_c"].reshape(mean_c.shape).to(dtype),
                                  g["Z"].reshape(Z.shape).to(dtype), g["variance"].reshape(variance.shape).to(dtype),
                                  g["lengthscale"].reshape(lengthscale.shape).to(dtype))
        objective.last_output = out
        return torch.tensor(out.bound, dtype=dtype, device=device)

    @staticmethod
    def backward(ctx, grad_out):
        g = ctx.saved_tensors
        return (*(grad_out * t for t in g), None, None)


class LowerBoundCG(nn.Module):
    """Reference: `LowerBoundCG(ExactMarginalLogLikelihood)`, models.py:104-286."""

    def __init__(self, model: SGPR, cg_opt: Optional[ConjugateGradient] = None, use_cache: bool = False,
                 cached_v_vec_initial: bool = False, shard: Optional[Shard] = None):
        if not isinstance(model, SGPR):
            raise ValueError(f"CGLB model expected in the constructor of the {self.__class__}")      # models.py:112-113
        super().__init__()
        object.__setattr__(self, "model", model)
        self.cg_opt = ConjugateGradient() if cg_opt is None else cg_opt
        self._cached_v_vec = cached_v_vec_initial
        self._use_cache = use_cache
        self._shard = shard
        self._evaluator: Optional[BoundEvaluator] = None
        self._evaluator_key = None
        self.last_output = None

    # ---- reference properties (models.py:122-149) ----------------------------------------------------
    @property
    def cached_v_vec(self) -> bool:
        return self._cached_v_vec

    @cached_v_vec.setter
    def cached_v_vec(self, value: bool):
        self._cached_v_vec = value

    @property
    def mean(self):
        return self.model.mean_module

    @property
    def likelihood(self):
        return self.model.likelihood

    @property
    def kernel(self):
        return self.model.covar_module.base_kernel

    @property
    def inducing_points(self) -> Tensor:
        return self.model.covar_module.inducing_points

    @property
    def noise(self) -> Tensor:
        return self.likelihood.noise.squeeze()

    # ---- evaluation -------------------------------------------------------------------------------------
    def evaluator(self, data) -> BoundEvaluator:
        x, y = data
        key = (x.data_ptr(), y.data_ptr(), tuple(x.shape))
        version = (x._version, y._version)
        if self._evaluator is None or self._evaluator_key != key:
            if not x.is_cuda:
                raise CglbError("LowerBoundCG needs CUDA tensors: cglb_b200 has no CPU fallback")
            # fp32 models (set_default_float("fp32")): kernel pairs in FP32, everything else promoted to FP64
            self._evaluator = BoundEvaluator(x.to(torch.float64), y.to(torch.float64), self._shard,
                                             pair_dtype="f32" if x.dtype == torch.float32 else "f64")
            self._evaluator_key, self._evaluator_version = key, version
        elif self._evaluator_version != version:
            # same buffers, new contents (e.g. a fresh host->device copy): keep the workspaces
            self._evaluator.refresh_data(x.to(torch.float64), y.to(torch.float64))
            self._evaluator_version = version
        return self._evaluator

    def _evaluate(self, data, need_grad: bool):
        ev = self.evaluator(data)
        kind, ls, var = _kernel_pieces(self.kernel)
        model = self.model
        use_cached = bool(self._use_cache and self.cached_v_vec)                         # models.py:263
        v64 = model.v_vec if model.v_vec.dtype == torch.float64 else model.v_vec.to(torch.float64)
        out = ev.evaluate(kind, self.inducing_points.detach().to(torch.float64), ls.detach().to(torch.float64),
                          float(var), float(self.noise), float(self.mean.constant.detach().reshape(-1)[0]), v64,
                          self.cg_opt, settings.cholesky_jitter.value(), use_cached_v=use_cached, need_grad=need_grad)
        if not use_cached:
            if v64 is not model.v_vec:
                model.v_vec.data.copy_(v64)                                              # models.py:274
            model.cg_stats = out.cg_stats                                                # :271
            self.cached_v_vec = self._use_cache                                          # :278
        return out

    def forward(self, data: Tuple[Tensor, Tensor], *params) -> Tensor:
        """bound = -upper + logdet - n/2 log 2 pi  (models.py:151-174), differentiable w.r.t. model.parameters()."""
        kind, ls, var = _kernel_pieces(self.kernel)
        return _BoundFunction.apply(self.likelihood.noise, self.mean.constant, self.inducing_points, var, ls, self, data)

    # ---- the reference's sub-steps, for callers that use them directly (not differentiable here) -------
    def logdet_and_quad_common_terms(self, data: Tuple) -> CommonTerms:                   # models.py:176-213
        ev = self.evaluator(data)
        kind, ls, var = _kernel_pieces(self.kernel)
        t = ev.common_terms(kind, self.inducing_points.detach().to(torch.float64),
                            ls.detach().reshape(-1).to(torch.float64).contiguous(), float(var),
                            float(self.noise), settings.cholesky_jitter.value())
        return CommonTerms(A=t.A[:, :t.ncols], LB=t.LB, AAt_diag_sum=t.AAt_diag_sum, L=t.L)

    def logdet_estimator(self, data: Tuple, terms: CommonTerms) -> Tensor:               # models.py:215-244
        x_data, y_data = data
        num_data = float(y_data.shape[0])
        sigma_sq = self.noise.detach()
        kdiag = self.kernel(x_data, diag=True).detach()
        trace = kdiag.sum() / sigma_sq - terms.AAt_diag_sum
        logdet = -terms.LB.diagonal().log().sum()
        logdet = logdet - 0.5 * num_data * torch.log(sigma_sq)
        logdet = logdet - 0.5 * num_data * torch.log(1.0 + trace / num_data)
        return logdet

    def quad_estimator(self, data, terms: Optional[CommonTerms] = None) -> Bounds:       # models.py:246-286
        out = self._evaluate(data, need_grad=False)
        dtype, device = self.noise.dtype, self.noise.device
        return Bounds(upper_bound=torch.tensor(out.upper, dtype=dtype, device=device),
                      lower_bound=torch.tensor(out.lower, dtype=dtype, device=device))


class PredictCG(LowerBoundCG):
    """Reference: models.py:289-354 (tight CG, max_error=1e-3)."""

    def __init__(self, model: SGPR, cg_opt: Optional[ConjugateGradient] = None, shard: Optional[Shard] = None):
        cg_opt = ConjugateGradient(max_error=1e-3) if cg_opt is None else cg_opt
        super().__init__(model, cg_opt, shard=shard)
        self._v_vec = model.v_vec.detach().clone()
        self.cached = False
        self.terms = None

    @property
    def v_vec(self):
        return self._v_vec

    def clear_cache(self):
        self.v_vec.copy_(self.model.v_vec.detach().clone())
        self.cached = False
        self.terms = None

    @torch.no_grad()
    def forward(self, xnew: Tensor, full_cov: bool = False, full_output_cov: bool = False) -> Tuple[Tensor, Tensor]:
        if full_cov:
            raise NotImplementedError("The predict_f method currently  supports only `full_cov=False` option")
        x, *_ = self.model.train_inputs
        out_dtype = xnew.dtype
        f64 = torch.float64
        # fp32 models (set_default_float("fp32")): every operand is promoted, as LowerBoundCG._evaluate does, and the
        # results are cast back (the kernels are fp64)
        y = self.model.train_targets.reshape(-1, 1).to(f64)
        ev = self.evaluator((x, self.model.train_targets))
        eng = ev.eng
        kind, ls, var = _kernel_pieces(self.kernel)
        ls = ls.detach().reshape(-1).to(f64).contiguous()
        var_f, noise = float(var), float(self.noise)
        mean_c = float(self.mean.constant.detach().reshape(-1)[0])
        err = (y - mean_c).contiguous()
        if self.cached:
            terms = self.terms
            ev.pack(kind, ls)
        else:
            terms = ev.common_terms(kind, self.inducing_points.detach().to(f64), ls, var_f, noise, settings.cholesky_jitter.value())
        cov = ev.operator(kind, var_f, noise)
        precon = ev.preconditioner(terms, noise)
        if self.cached:
            new_v = self.v_vec.to(f64)
        else:
            new_v, cg_stats = self.cg_opt(cov, err, self.v_vec.to(f64), precon)          # :330
            self.v_vec.data.copy_(new_v)
            self.terms = terms
            self.cached = True
        xnew = xnew.detach().to(f64).contiguous()
        nnew, d = xnew.shape
        m = terms.L.shape[0]
        xnew_p = eng.pack(kind, xnew, ls, ev.shift)
        cg_mean = eng.kmv_rect(kind, xnew_p, nnew, ev.xp, ev.n, d, new_v.reshape(-1).contiguous(), var_f).reshape(-1, 1)   # :334
        res = err - cov @ new_v                                                          # :335
        # a_res = A @ res : this rank's columns, all-reduced                             # :340
        a_res = eng.empty(m)
        eng.precond_project(terms.A, m, terms.ncols, res.reshape(-1)[ev.lo:ev.hi], a_res)
        ev.shard.all_reduce(a_res)
        sigma = noise ** 0.5
        c = torch.mv(terms.LBinv, a_res) / sigma                                         # :343  (LB^-1 through its inverse)
        ldn = nnew + (nnew & 1)
        kus = eng.zeros(m, ldn)
        eng.knm_build(kind, terms.zp, m, xnew_p, nnew, d, var_f, kus, ldn)               # :337
        tmp1 = eng.trsm_left_lower(terms.L, kus, nnew)                                   # :344
        tmp1v = tmp1[:, :nnew]
        tmp2 = tmp1.clone()
        eng.trsm_left_lower(terms.LB, tmp2, nnew)                                        # :345
        tmp2v = tmp2[:, :nnew]
        sgpr_mean = (tmp2v.t() @ c).reshape(-1, 1)                                       # :347
        f_mean = sgpr_mean + cg_mean + mean_c                                            # :348
        f_var = var_f + (tmp2v ** 2).sum(0) - (tmp1v ** 2).sum(0)                        # :350-351
        return f_mean.to(out_dtype), f_var.reshape(*f_mean.shape).to(out_dtype)


def log_density(m, y, f_mean, f_var) -> Tensor:          # models.py:370-372
    noise = m.likelihood.noise.squeeze()
    return gaussian(y, f_mean, f_var + noise).sum(axis=-1)


def gaussian(x, mu, var):                                # models.py:375-379
    pi = torch.tensor(np.pi, dtype=x.dtype, device=x.device)
    pi2 = torch.log(2 * pi)
    x = x.reshape(*mu.shape)
    return -0.5 * (pi2 + torch.log(var) + (mu - x) ** 2 / var)


def _output_dims(t: Tensor) -> Tuple[Tensor, Tensor]:    # models.py:382-385
    num_data = torch.tensor(t.size(0), dtype=t.dtype)
    output_dim = torch.tensor(t.size(1) if t.ndim == 2 else 1, dtype=t.dtype)
    return num_data, output_dim
