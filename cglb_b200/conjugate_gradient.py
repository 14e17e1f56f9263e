"""Preconditioned CG and the Nystrom/Woodbury preconditioner -- same names, signatures, defaults and stop
rule as the reference (cglb/backend/pytorch/conjugate_gradient.py), executed by fused sm_100a kernels.

Differences in mechanism (not in results): vector updates and dot products are fused device kernels with
device-resident scalars (K8); the preconditioner streams A twice with hand-written GEMVs (K3) and applies
B^-1 through the explicit inverse of its Cholesky factor (K4); `A` may be row-sharded (distributed.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, Optional, Tuple, Union

import torch

from ._ffi import CglbError
from .distributed import Shard
from .engine import get_engine

Tensor = torch.Tensor
Preconditioner = Callable[[Tensor], Tuple[Tensor, Tensor]]


@dataclass
class ConjugateGradientStats:                      # reference conjugate_gradient.py:25-28
    steps: Union[Tensor, float]
    residual_error: Union[Tensor, float]


@dataclass
class ConjugateGradientState:
    """What the loop holds when it stops: the residual of the returned v (from the recurrence r -= gamma A p, or
    recomputed as b - A v at the start and at every restart), z = P r and r^T z of the LAST preconditioner
    application, which was made with exactly this r.  Not part of the reference API (`solve` returns it)."""
    r: Tensor
    z: Tensor
    rz: Tensor


@dataclass
class ConjugateGradient:
    """CG stops if: 0.5 * r^T Q^-1 r < max_error || i > max_cg_iter   (reference :31-39)"""

    max_error: float = 1.0
    max_cg_iter: int = 100
    restart_cg_iter: int = 40

    def __call__(self, A: Any, b: Tensor, v: Tensor, precond: Preconditioner) -> Tuple[Tensor, ConjugateGradientStats]:
        """:param A: operator supporting `A @ x` for x of shape [N, 1]  (reference :41-86)
        :param b: [N, 1] right-hand side;  :param v: [N, 1] warm start;  :param precond: r -> (z, r^T z)"""
        v, stats, _ = self.solve(A, b, v, precond)
        return v, stats

    def solve(self, A: Any, b: Tensor, v: Tensor, precond: Preconditioner
              ) -> Tuple[Tensor, ConjugateGradientStats, ConjugateGradientState]:
        """`__call__` plus the final loop state, so that a caller that needs b - A v and P (b - A v) afterwards
        (models.py:280-282) does not have to pay another n^2 sweep for them."""
        if not b.is_cuda:
            raise CglbError("ConjugateGradient needs CUDA tensors (cglb_b200 has no CPU fallback)")
        eng = get_engine(b.device)
        n = b.numel()
        # the kernels are fp64 and take dense vectors: fp32 models (set_default_float("fp32")) and strided views are
        # promoted here, the results are cast back to the caller's dtype below
        out_dtype = b.dtype
        b = b.to(torch.float64).contiguous()
        v = v.to(torch.float64).clone().contiguous()                  # :55
        Av = (A @ v).to(torch.float64).contiguous()                   # :57
        r = torch.empty_like(b)
        eng.residual(n, b, Av, r)                                     # :58
        z, rz = precond(r)                                            # :59
        z = z.to(torch.float64).contiguous()
        p = z.clone()
        rz = rz.to(torch.float64).reshape(1).clone()
        pAp = torch.empty_like(rz)
        rz_host = float(rz.item())
        i = 0
        while (0.5 * rz_host > self.max_error) and (i < self.max_cg_iter):        # :65
            Ap = (A @ p).to(torch.float64).contiguous()               # :66
            eng.dot(p, Ap, pAp)
            restart = i % self.restart_cg_iter == self.restart_cg_iter - 1         # :70
            eng.cg_step(n, rz, pAp, p, Ap, v, r, restart)             # :67-68, :72 (non-restart branch)
            if restart:
                Av = (A @ v).to(torch.float64).contiguous()
                eng.residual(n, b, Av, r)                             # :72 (restart branch)
            z, new_rz = precond(r)                                    # :73
            new_rz = new_rz.to(torch.float64).reshape(1)
            z = z.to(torch.float64).contiguous()
            eng.cg_direction(n, z, p, new_rz, rz, restart)            # :75
            rz = new_rz.clone()
            rz_host = float(rz.item())                                # the loop test is evaluated on the host, as in :65
            i += 1
        stats = ConjugateGradientStats(i, torch.tensor(0.5 * rz_host, dtype=out_dtype))   # :83
        return v.to(out_dtype), stats, ConjugateGradientState(r=r, z=z, rz=rz)


@dataclass
class NystromPreconditioner:
    """z = (r - A^T B^-1 A r) / sigma^2, rz = r^T z with B = I + A A^T = LB LB^T   (reference :89-113).

    `A` is [M, N] (the reference's "[N, N]" comment is a typo).  When `shard` has
    world > 1, `A` holds only this rank's column block `cols = (lo, hi)` and r / z are full-length,
    replicated vectors; q = A r and (z, rz) are all-reduced."""
    A: Tensor
    LB: Tensor
    sigma_sq: Union[Tensor, float]
    shard: Optional[Shard] = None
    cols: Optional[Tuple[int, int]] = None
    lbinv: Optional[Tensor] = field(default=None, repr=False)

    def __post_init__(self):
        if not self.A.is_cuda:
            raise CglbError("NystromPreconditioner needs CUDA tensors (cglb_b200 has no CPU fallback)")
        self._out_dtype = self.A.dtype
        if self.A.dtype != torch.float64:         # fp32 models: the GEMV pair streams an fp64 copy
            self.A = self.A.to(torch.float64)
        if self.A.stride(-1) != 1:
            self.A = self.A.contiguous()          # e.g. a column-major result of a triangular solve
        self._eng = get_engine(self.A.device)
        if self.lbinv is None:
            self.lbinv = self._eng.tri_inverse(self.LB.detach().to(torch.float64).contiguous())
        self._m = self.A.shape[0]
        self._sigma_sq = float(self.sigma_sq)
        if self.shard is None:
            self.shard = Shard()
        self._q = self._eng.empty(self._m)
        self._w = self._eng.empty(self._m)

    def __call__(self, r: Tensor) -> Tuple[Tensor, Tensor]:
        eng, m = self._eng, self._m
        rf = r.detach().reshape(-1).to(torch.float64).contiguous()
        n = rf.numel()
        lo, hi = self.cols if self.cols is not None else (0, n)
        ncols = hi - lo
        # z buffer carries rz in its last slot so that one all-reduce moves both
        zbuf = eng.zeros(n + 1) if self.shard.world > 1 else eng.empty(n + 1)
        r_loc = rf[lo:hi]
        eng.precond_project(self.A, m, ncols, r_loc, self._q)                     # :105
        self.shard.all_reduce(self._q)
        eng.precond_finish(self.A, m, ncols, self.lbinv, self._q, r_loc, self._sigma_sq,
                           zbuf[lo:hi], self._w, zbuf[n:])                        # :106-113
        self.shard.all_reduce(zbuf)
        if self._out_dtype != torch.float64 and r.dtype != torch.float64:
            return zbuf[:n].reshape(r.shape).to(r.dtype), zbuf[n].to(r.dtype)
        return zbuf[:n].reshape(r.shape), zbuf[n]

    @property
    def w(self) -> Tensor:
        """B^-1 A r of the last application (used by the fused backward pass)."""
        return self._w
