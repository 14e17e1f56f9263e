"""fp64 CPU restatement of the CGLB hot path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every function cites the reference lines it follows (paths relative to /root/reference).
Kernel arithmetic follows the *direct-difference* form r = ||(x - x')/l|| that the reference's KeOps
path uses at the sizes that matter; the GPyTorch dense "expanded" form is kept as
``sqdist_expanded`` for cross-checks.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import numpy as np
import torch

Tensor = torch.Tensor

SQRT3 = math.sqrt(3.0)
NOISE_FLOOR = 1e-6  # GreaterThan(1e-6), cglb/backend/pytorch/interface.py:269


# ----------------------------------------------------------------------------------------------
# parameter transforms (gpytorch Positive / GreaterThan constraints = softplus)
# ----------------------------------------------------------------------------------------------
def softplus(x: Tensor) -> Tensor:
    return torch.nn.functional.softplus(x)


def inv_softplus(y) -> Tensor:
    y = torch.as_tensor(y, dtype=torch.float64)
    return y + torch.log(-torch.expm1(-y))


# ----------------------------------------------------------------------------------------------
# kernels (third-party gpytorch arithmetic restated; interface.py:207-230 builds
# ScaleKernel(MaternKernel(nu=1.5, ARD)) / ScaleKernel(RBFKernel(ARD)))
# ----------------------------------------------------------------------------------------------
def sqdist_direct(x1: Tensor, x2: Tensor, lengthscale: Tensor) -> Tensor:
    """sum_q ((x1_q - x2_q)/l_q)^2, direct differences (KeOps form)."""
    a = x1 / lengthscale
    b = x2 / lengthscale
    diff = a[:, None, :] - b[None, :, :]
    return (diff * diff).sum(-1)


def sqdist_expanded(x1: Tensor, x2: Tensor, lengthscale: Tensor, x1_eq_x2: bool = False) -> Tensor:
    """GPyTorch dense form: centre by mean(x1), ||a||^2 + ||b||^2 - 2ab, clamp >= 0, zero diagonal."""
    adj = x1.mean(-2, keepdim=True)
    a = (x1 - adj) / lengthscale
    b = (x2 - adj) / lengthscale
    res = (a * a).sum(-1, keepdim=True) + (b * b).sum(-1)[None, :] - 2.0 * a @ b.T
    if x1_eq_x2:
        res = res - torch.diag(torch.diagonal(res))
    return res.clamp_min(0.0)


def kernel_dense(kind: str, x1: Tensor, x2: Tensor, lengthscale: Tensor, variance: Tensor,
                 block: int = 0) -> Tensor:
    """sigma_f^2 kappa(r).  Matern32: (1+sqrt3 r) exp(-sqrt3 r); RBF: exp(-r^2/2)."""
    if block and x1.shape[0] > block:
        return torch.cat([kernel_dense(kind, x1[i:i + block], x2, lengthscale, variance)
                          for i in range(0, x1.shape[0], block)], 0)
    sq = sqdist_direct(x1, x2, lengthscale)
    if kind == "matern32":
        r = torch.sqrt(sq.clamp_min(1e-30))
        s = SQRT3 * r
        return variance * (1.0 + s) * torch.exp(-s)
    if kind == "rbf":
        return variance * torch.exp(-0.5 * sq)
    raise NotImplementedError(kind)


def kernel_diag(x: Tensor, variance: Tensor) -> Tensor:
    return variance.expand(x.shape[0]) if variance.ndim == 0 else variance.reshape(()).expand(x.shape[0])


class DenseOrBlockedKernelOperator:
    """K(X,X) + sigma^2 I as an object supporting ``A @ v`` (the operator protocol of
    conjugate_gradient.py:57,66,72 and models.py:251-252,280).  Dense when n is small, else row-blocked."""

    def __init__(self, kind, x, lengthscale, variance, sigma_sq, dense_max: int = 6000, block: int = 1024):
        self.kind, self.x, self.ls, self.var, self.sigma_sq = kind, x, lengthscale, variance, sigma_sq
        self.block = block
        self._dense = None
        if x.shape[0] <= dense_max:
            self._dense = kernel_dense(kind, x, x, lengthscale, variance)

    def detach(self):
        out = DenseOrBlockedKernelOperator.__new__(DenseOrBlockedKernelOperator)
        out.kind, out.x, out.ls, out.var = self.kind, self.x.detach(), self.ls.detach(), self.var.detach()
        out.sigma_sq, out.block = self.sigma_sq.detach(), self.block
        out._dense = None if self._dense is None else self._dense.detach()
        return out

    def __matmul__(self, v: Tensor) -> Tensor:
        if self._dense is not None:
            return self._dense @ v + self.sigma_sq * v
        rows = []
        for i in range(0, self.x.shape[0], self.block):
            kb = kernel_dense(self.kind, self.x[i:i + self.block], self.x, self.ls, self.var)
            rows.append(kb @ v)
        return torch.cat(rows, 0) + self.sigma_sq * v


# ----------------------------------------------------------------------------------------------
# solver (cglb/backend/pytorch/conjugate_gradient.py)
# ----------------------------------------------------------------------------------------------
@dataclass
class CGStats:
    steps: int
    residual_error: float


def nystrom_preconditioner(A: Tensor, LB: Tensor, sigma_sq: Tensor) -> Callable[[Tensor], Tuple[Tensor, Tensor]]:
    """conjugate_gradient.py:89-113: z = (r - A^T B^{-1} A r)/sigma^2, rz = r^T z."""

    def apply(r: Tensor):
        Ar = A @ r                                                            # :105
        t = torch.linalg.solve_triangular(LB, Ar, upper=False)                # :106
        t = torch.linalg.solve_triangular(LB.transpose(-1, -2), t, upper=True)  # :107
        p = t.transpose(-1, -2) @ A                                           # :110
        rp = r - p.transpose(-1, -2)                                          # :111
        rpr = (rp * r).sum()                                                  # :112
        return rp / sigma_sq, rpr / sigma_sq                                  # :113

    return apply


def conjugate_gradient(A, b: Tensor, v: Tensor, precond, max_error: float = 1.0,
                       max_cg_iter: int = 100, restart_cg_iter: int = 40) -> Tuple[Tensor, CGStats]:
    """conjugate_gradient.py:41-86 (defaults :37-39)."""
    v = v.clone()                                                             # :55
    Av = A @ v                                                                # :57
    r = b - Av
    z, rz = precond(r)
    p = z
    i = 0
    while (0.5 * rz > max_error) and (i < max_cg_iter):                       # :65
        Ap = A @ p
        gamma = rz / (p * Ap).sum()
        v = v + gamma * p
        restart = i % restart_cg_iter == restart_cg_iter - 1                  # :70
        r = (b - A @ v) if restart else (r - gamma * Ap)
        z, new_rz = precond(r)
        p = z if restart else (z + p * new_rz / rz)
        rz = new_rz
        i += 1
    return v, CGStats(i, float(0.5 * rz))


# ----------------------------------------------------------------------------------------------
# model / objective (cglb/backend/pytorch/models.py)
# ----------------------------------------------------------------------------------------------
@dataclass
class OracleParams:
    """Raw (unconstrained) parameters in the order ``model.parameters()`` yields them in the
    reference: raw noise [1], mean constant [1], Z [M,d], raw outputscale [], raw
    lengthscale [1,d]."""
    raw_noise: Tensor
    mean_constant: Tensor
    inducing_points: Tensor
    raw_outputscale: Tensor
    raw_lengthscale: Tensor

    @staticmethod
    def from_values(noise, mean_c, Z, variance, lengthscale) -> "OracleParams":
        f64 = torch.float64
        Z = torch.as_tensor(Z, dtype=f64).clone()
        d = Z.shape[1]
        ls = torch.as_tensor(lengthscale, dtype=f64).reshape(-1)
        if ls.numel() == 1:
            ls = ls.repeat(d)
        return OracleParams(
            raw_noise=inv_softplus(torch.as_tensor([noise - NOISE_FLOOR], dtype=f64)).requires_grad_(True),
            mean_constant=torch.as_tensor([mean_c], dtype=f64).requires_grad_(True),
            inducing_points=Z.requires_grad_(True),
            raw_outputscale=inv_softplus(torch.as_tensor(variance, dtype=f64)).reshape(()).requires_grad_(True),
            raw_lengthscale=inv_softplus(ls).reshape(1, d).clone().requires_grad_(True),
        )

    def tensors(self):
        return [self.raw_noise, self.mean_constant, self.inducing_points, self.raw_outputscale,
                self.raw_lengthscale]

    @property
    def noise(self):        # likelihood.noise.squeeze(), models.py:147-149
        return (softplus(self.raw_noise) + NOISE_FLOOR).squeeze()

    @property
    def variance(self):
        return softplus(self.raw_outputscale)

    @property
    def lengthscale(self):
        return softplus(self.raw_lengthscale)


@dataclass
class CommonTerms:      # models.py:90-95
    A: Tensor
    LB: Tensor
    AAt_diag_sum: Tensor
    L: Tensor


def common_terms(kind: str, p: OracleParams, x: Tensor, jitter: float) -> CommonTerms:
    """models.py:176-213."""
    sigma_sq = p.noise
    sigma = torch.sqrt(sigma_sq)
    Z = p.inducing_points
    kuf = kernel_dense(kind, Z, x, p.lengthscale, p.variance, block=256)          # :196-197
    kuu = kernel_dense(kind, Z, Z, p.lengthscale, p.variance)                    # :200
    kuu_jitter = kuu + jitter * torch.eye(Z.shape[0], dtype=kuu.dtype)           # :201
    L = torch.linalg.cholesky(kuu_jitter)                                        # :202
    A = torch.linalg.solve_triangular(L, kuf, upper=False) / sigma               # :206
    AAt = A @ A.transpose(-1, -2)                                                # :207
    B = AAt + torch.eye(Z.shape[0], dtype=AAt.dtype)                             # :208-209
    LB = torch.linalg.cholesky(B)                                                # :210
    return CommonTerms(A=A, LB=LB, AAt_diag_sum=AAt.diagonal().sum(), L=L)       # :211-213


def logdet_term(p: OracleParams, x: Tensor, terms: CommonTerms) -> Tensor:
    """models.py:215-244 (output_dim = 1)."""
    n = float(x.shape[0])
    sigma_sq = p.noise
    kdiag_sum = p.variance * n                                                   # :233 (stationary kernel)
    trace = kdiag_sum / sigma_sq - terms.AAt_diag_sum                            # :236
    logdet = -terms.LB.diagonal().log().sum()                                    # :239
    logdet = logdet - 0.5 * n * torch.log(sigma_sq)                              # :240
    logdet = logdet - 0.5 * n * torch.log(1.0 + trace / n)                       # :243
    return logdet


@dataclass
class BoundResult:
    bound: Tensor
    upper: Tensor          # Bounds.upper_bound (= -upper), models.py:286
    lower: Tensor
    v: Tensor
    cg: Optional[CGStats]
    logdet: Tensor


def lower_bound(kind: str, p: OracleParams, x: Tensor, y: Tensor, v0: Tensor, jitter: float = 1e-6,
                max_error: float = 1.0, max_cg_iter: int = 100, restart_cg_iter: int = 40,
                use_cached_v: bool = False, dense_max: int = 6000) -> BoundResult:
    """LowerBoundCG.forward, models.py:151-174 with quad_estimator :246-286.

    ``v0`` plays the role of ``model.v_vec`` (warm start, models.py:59-68, :274); the caller copies
    ``result.v`` back into it, as the reference does in place."""
    n = float(x.shape[0])
    terms = common_terms(kind, p, x, jitter)
    const_term = -0.5 * n * math.log(2.0 * math.pi)                              # :162-163
    logdet = logdet_term(p, x, terms)

    sigma_sq = p.noise
    cov = DenseOrBlockedKernelOperator(kind, x, p.lengthscale, p.variance, sigma_sq, dense_max=dense_max)
    err = y.reshape(-1, 1) - p.mean_constant.reshape(1, 1)                       # :253-254
    precon = nystrom_preconditioner(terms.A, terms.LB, sigma_sq)                 # :260
    cg_stats = None
    with torch.no_grad():                                                        # :262
        if use_cached_v:
            v = v0
        else:
            v, cg_stats = conjugate_gradient(cov.detach(), err.detach(), v0, precon,
                                             max_error, max_cg_iter, restart_cg_iter)
    cov_v = cov @ v                                                              # :280
    r = err - cov_v
    _, error_bound = precon(r)
    lower = (v * (r + 0.5 * cov_v)).sum()                                        # :283
    upper = lower + 0.5 * error_bound                                            # :284
    bound = -upper + logdet + const_term                                         # :168
    return BoundResult(bound=bound, upper=-upper, lower=-lower, v=v.detach(), cg=cg_stats, logdet=logdet)


def bound_and_grads(kind: str, p: OracleParams, x, y, v0, **kw):
    """loss = -bound; torch.autograd.grad(loss, params)  (pytorch/optimizer.py:95-98)."""
    res = lower_bound(kind, p, x, y, v0, **kw)
    loss = -res.bound
    grads = torch.autograd.grad(loss, p.tensors())
    return loss.detach(), [g.detach() for g in grads], res


def bound_and_grads_fixed_v_blocked(kind: str, p: OracleParams, x: Tensor, y: Tensor, v: Tensor, jitter: float = 1e-6,
                                    block: int = 256):
    """``bound_and_grads(..., use_cached_v=True)`` for n too large for a dense K with an autograd tape (the [n, n, d]
    difference tensor of ``sqdist_direct``): the same statements of models.py:151-286 with ``cov @ v`` (models.py:280)
    evaluated in row blocks.  Exact autograd in two stages: (1) ``Kv`` enters the bound as a leaf, one backward gives the
    direct parameter gradients and g = d(loss)/d(Kv); (2) sum_blocks g_blk^T (K_blk(theta) v) is differentiated block by
    block, so only one [block, n, d] tape is alive at a time.  Checked against the dense oracle in
    tests/test_oracle_golden.py."""
    n = x.shape[0]
    nf = float(n)
    v = v.detach().reshape(-1, 1)
    sigma_sq_d = p.noise.detach()
    with torch.no_grad():
        rows = [kernel_dense(kind, x[i:i + block], x, p.lengthscale, p.variance) @ v for i in range(0, n, block)]
        kv_val = torch.cat(rows, 0)
    kv_leaf = kv_val.clone().requires_grad_(True)                                # variance K(X,X) v
    terms = common_terms(kind, p, x, jitter)
    logdet = logdet_term(p, x, terms)
    sigma_sq = p.noise
    cov_v = kv_leaf + sigma_sq * v                                               # :251-252, :280
    err = y.reshape(-1, 1) - p.mean_constant.reshape(1, 1)                       # :253-254
    precon = nystrom_preconditioner(terms.A, terms.LB, sigma_sq)
    r = err - cov_v                                                              # :281
    _, error_bound = precon(r)                                                   # :282
    lower = (v * (r + 0.5 * cov_v)).sum()                                        # :283
    upper = lower + 0.5 * error_bound                                            # :284
    bound = -upper + logdet - 0.5 * nf * math.log(2.0 * math.pi)                 # :162-168
    loss = -bound
    params = p.tensors()
    grads = list(torch.autograd.grad(loss, params + [kv_leaf]))
    g_kv = grads.pop().detach()
    grads = [g.detach().clone() for g in grads]
    kparams = [p.raw_outputscale, p.raw_lengthscale]
    for i in range(0, n, block):
        f = (g_kv[i:i + block] * (kernel_dense(kind, x[i:i + block], x, p.lengthscale, p.variance) @ v)).sum()
        go, gl = torch.autograd.grad(f, kparams)
        grads[3] += go
        grads[4] += gl
    del sigma_sq_d
    return loss.detach(), grads


def predict(kind: str, p: OracleParams, x: Tensor, y: Tensor, xnew: Tensor, v0: Tensor,
            jitter: float = 1e-6, max_error: float = 1e-3):
    """PredictCG.forward, models.py:307-354 (tight CG, max_error=1e-3 at :291)."""
    with torch.no_grad():
        yv = y.reshape(-1, 1)
        err = yv - p.mean_constant.reshape(1, 1)
        sigma_sq = p.noise
        ksf = kernel_dense(kind, xnew, x, p.lengthscale, p.variance, block=512)
        cov = DenseOrBlockedKernelOperator(kind, x, p.lengthscale, p.variance, sigma_sq)
        terms = common_terms(kind, p, x, jitter)
        precon = nystrom_preconditioner(terms.A, terms.LB, sigma_sq)
        new_v, stats = conjugate_gradient(cov, err, v0, precon, max_error)
        cg_mean = ksf @ new_v                                                    # :334
        res = err - cov @ new_v
        kus = kernel_dense(kind, p.inducing_points, xnew, p.lengthscale, p.variance)
        sigma = torch.sqrt(sigma_sq)
        a_res = terms.A @ res                                                    # :340
        c = torch.linalg.solve_triangular(terms.LB, a_res, upper=False) / sigma
        tmp1 = torch.linalg.solve_triangular(terms.L, kus, upper=False)
        tmp2 = torch.linalg.solve_triangular(terms.LB, tmp1, upper=False)
        sgpr_mean = tmp2.transpose(-1, -2) @ c
        f_mean = sgpr_mean + cg_mean + p.mean_constant.reshape(1, 1)             # :348
        kss = p.variance.expand(xnew.shape[0])
        f_var = kss + (tmp2 ** 2).sum(0) - (tmp1 ** 2).sum(0)                    # :351
        return f_mean, f_var.reshape(*f_mean.shape), new_v, stats


# ----------------------------------------------------------------------------------------------
# synthetic data of the BASELINE.json shapes
# ----------------------------------------------------------------------------------------------
def synthetic_problem(n: int, d: int, M: int, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g, dtype=torch.float64)
    g1 = torch.Generator().manual_seed(seed + 1)
    w = torch.randn(d, generator=g1, dtype=torch.float64)
    w2 = torch.randn(d, generator=g1, dtype=torch.float64)
    f = torch.sin(2.0 * (x @ w) / math.sqrt(d)) + 0.5 * torch.cos((x @ w2) / math.sqrt(d))
    g2 = torch.Generator().manual_seed(seed + 2)
    yv = f + 0.1 * torch.randn(n, generator=g2, dtype=torch.float64)
    yv = (yv - yv.mean()) / yv.std()
    g3 = torch.Generator().manual_seed(seed + 3)
    perm = torch.randperm(n, generator=g3)
    z = x[perm[:M]].clone()
    return x, yv, z


def blocked_matvec_rows(kind: str, x: Tensor, v: Tensor, lengthscale: Tensor, variance: Tensor,
                        row0: int, nrows: int) -> Tensor:
    """Rows [row0, row0+nrows) of K(X,X) v -- the CPU baseline's bounded sample for large n."""
    kb = kernel_dense(kind, x[row0:row0 + nrows], x, lengthscale, variance)
    return kb @ v
