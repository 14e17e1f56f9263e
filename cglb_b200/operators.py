"""Lazy kernel operators: the `A @ x` protocol of the reference's solver (conjugate_gradient.py:57,66,72,
models.py:251-252,280) executed by the matrix-free sm_100a sweeps instead of KeOps / dense torch."""
from __future__ import annotations

from typing import Optional

import torch

from ._ffi import CglbError
from .engine import get_engine

Tensor = torch.Tensor


def _kernel_pieces(kernel):
    """(kind, lengthscale[d] tensor, outputscale 0-dim tensor) of ScaleKernel(base) or a bare base kernel."""
    from .gp import InducingPointKernel, ScaleKernel
    if isinstance(kernel, InducingPointKernel):
        kernel = kernel.base_kernel
    if isinstance(kernel, ScaleKernel):
        base = kernel.base_kernel
        return base.kind, base.lengthscale.reshape(-1), kernel.outputscale.reshape(())
    return kernel.kind, kernel.lengthscale.reshape(-1), torch.ones((), dtype=kernel.lengthscale.dtype,
                                                                   device=kernel.lengthscale.device)


class _KmvFunction(torch.autograd.Function):
    """y = (variance K(X,X) + diag I) v with gradients from the fused backward sweep (K2)."""

    @staticmethod
    def forward(ctx, lengthscale, variance, diag, v, op):
        eng = op.engine
        n, d = op.x1.shape
        ls = lengthscale.detach().to(torch.float64).contiguous()
        xp = op.packed(ls)
        vv = v.detach().reshape(-1).to(torch.float64).contiguous()
        y = eng.kmv_sym(op.kind, xp, n, d, vv, float(variance), float(diag))
        ctx.op, ctx.xp = op, xp
        ctx.save_for_backward(ls, variance.detach(), diag.detach(), vv)
        return y.reshape(v.shape).to(v.dtype)

    @staticmethod
    def backward(ctx, gy):
        ls, variance, diag, vv = ctx.saved_tensors
        op, xp, eng = ctx.op, ctx.xp, ctx.op.engine
        n, d = op.x1.shape
        u = gy.detach().reshape(-1).to(torch.float64).contiguous()
        g_ls = g_var = g_diag = g_v = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            out = eng.zeros(d + 1)
            eng.kmv_bwd_sym(op.kind, xp, n, d, u, vv, float(variance), ls, out)
            g_ls = out[:d].reshape(ls.shape).to(gy.dtype)
            g_var = out[d].reshape(variance.shape).to(gy.dtype)
        if ctx.needs_input_grad[2]:
            g_diag = (u * vv).sum().reshape(diag.shape).to(gy.dtype)
        if ctx.needs_input_grad[3]:
            g_v = eng.kmv_sym(op.kind, xp, n, d, u, float(variance), float(diag)).reshape(gy.shape).to(gy.dtype)
        return g_ls, g_var, g_diag, g_v, None


class KernelOperator:
    """kernel(x1, x2) as a lazy matrix: supports add_diag, detach, @ and evaluate (delazify)."""

    def __init__(self, kernel, x1: Tensor, x2: Tensor, symmetric: bool, diag_value: Optional[Tensor] = None,
                 detached: bool = False):
        if not (x1.is_cuda and x2.is_cuda):
            raise CglbError("kernel operators need CUDA inputs (cglb_b200 has no CPU fallback)")
        # the sm_100a kernels are fp64: fp32 models (set_default_float("fp32")) are promoted here and results
        # are cast back to the caller's dtype
        self.out_dtype = x1.dtype
        self.kernel = kernel
        self.x1 = x1.detach().to(torch.float64).contiguous()
        self.x2 = self.x1 if (symmetric and x2 is x1) else x2.detach().to(torch.float64).contiguous()
        self.symmetric = symmetric
        self.diag_value = diag_value
        self.detached = detached
        self.engine = get_engine(x1.device)
        self.kind = _kernel_pieces(kernel)[0]
        self._packed = None
        self._shift = self.x1.mean(0).contiguous()

    @property
    def shape(self):
        return torch.Size([self.x1.shape[0], self.x2.shape[0]])

    def add_diag(self, value) -> "KernelOperator":
        if not self.symmetric:
            raise CglbError("add_diag needs a square kernel operator")
        value = torch.as_tensor(value, dtype=self.out_dtype, device=self.x1.device).reshape(())
        total = value if self.diag_value is None else self.diag_value + value
        out = KernelOperator(self.kernel, self.x1, self.x2, True, total, self.detached)
        out._packed, out._shift = self._packed, self._shift
        return out

    def detach(self) -> "KernelOperator":
        out = KernelOperator(self.kernel, self.x1, self.x2, self.symmetric,
                             None if self.diag_value is None else self.diag_value.detach(), True)
        out._packed, out._shift = self._packed, self._shift
        return out

    def packed(self, lengthscale: Tensor):
        # re-pack only when the lengthscale VALUES changed (pointer identity is not reliable: the softplus
        # transform returns a fresh tensor on every access)
        if self._packed is None or not torch.equal(self._packed[0], lengthscale):
            xp1 = self.engine.pack(self.kind, self.x1, lengthscale, self._shift)
            xp2 = xp1 if self.symmetric else self.engine.pack(self.kind, self.x2, lengthscale, self._shift)
            self._packed = (lengthscale.clone(), xp1, xp2, lengthscale)
        return self._packed[1]

    def __matmul__(self, v: Tensor) -> Tensor:
        kind, ls, var = _kernel_pieces(self.kernel)
        if v.ndim == 2 and v.shape[1] != 1:
            # x: [N, t] (conjugate_gradient.py:57,66,72).  Without a tape the block goes through the multi-RHS sweep (every
            # kernel pair evaluated once for all t columns); with one, column by column through the differentiable node.
            diag0 = self.diag_value if self.diag_value is not None else torch.zeros((), dtype=v.dtype, device=v.device)
            taped = torch.is_grad_enabled() and not self.detached and (
                ls.requires_grad or var.requires_grad or diag0.requires_grad or v.requires_grad)
            n, d = self.x1.shape
            if self.symmetric and not taped and d <= 32:
                xp = self.packed(ls.detach().to(torch.float64).contiguous())
                y = self.engine.kmv_sym_multi(kind, xp, n, d, v.detach().to(torch.float64).contiguous(), float(var), float(diag0))
                return y.to(v.dtype)
            return torch.cat([self @ v[:, j:j + 1] for j in range(v.shape[1])], 1)
        if self.symmetric:
            diag = self.diag_value if self.diag_value is not None else torch.zeros((), dtype=v.dtype, device=v.device)
            needs_grad = torch.is_grad_enabled() and not self.detached and (
                ls.requires_grad or var.requires_grad or diag.requires_grad or v.requires_grad)
            if needs_grad:
                return _KmvFunction.apply(ls, var, diag, v, self)
            lsd = ls.detach().to(torch.float64).contiguous()
            xp = self.packed(lsd)
            n, d = self.x1.shape
            y = self.engine.kmv_sym(kind, xp, n, d, v.detach().reshape(-1).to(torch.float64).contiguous(), float(var), float(diag))
            return y.reshape(v.shape).to(v.dtype)
        lsd = ls.detach().to(torch.float64).contiguous()
        self.packed(lsd)
        _, xp1, xp2, _ = self._packed
        n1, d = self.x1.shape
        y = self.engine.kmv_rect(kind, xp1, n1, xp2, self.x2.shape[0], d, v.detach().reshape(-1).to(torch.float64).contiguous(), float(var))
        return y.reshape(n1, *v.shape[1:]).to(v.dtype)

    def evaluate(self) -> Tensor:
        """Dense matrix (delazify).  Only sensible for M-sized operators."""
        kind, ls, var = _kernel_pieces(self.kernel)
        lsd = ls.detach().to(torch.float64).contiguous()
        self.packed(lsd)
        _, xp1, xp2, _ = self._packed
        n1, d = self.x1.shape
        n2 = self.x2.shape[0]
        ld = n2 + (n2 & 1)
        out = self.engine.zeros(n1, ld)
        self.engine.knm_build(kind, xp1, n1, xp2, n2, d, float(var), out, ld)
        dense = out[:, :n2]
        if self.diag_value is not None:
            dense = dense + self.diag_value.detach().to(dense.dtype) * torch.eye(n1, dtype=dense.dtype, device=dense.device)
        return dense.to(self.out_dtype)


def delazify(obj):
    return obj.evaluate() if isinstance(obj, KernelOperator) else obj
