"""Device-side evaluation of the CGLB bound and of its hyper-parameter gradients.

`BoundEvaluator` is the B200 implementation of what `LowerBoundCG.forward` + `torch.autograd.grad` do in
the reference (cglb/backend/pytorch/models.py:151-286, optimizer.py:95-98):

  common terms (models.py:176-213)   K_uf -> A = L^-1 K_uf / sigma (in place), A A^T, LB          [K5-K7]
  CG solve     (models.py:262-274)   preconditioned CG with the symmetric matrix-free sweep        [K1,K3,K4,K8]
  bound        (models.py:280-284, 215-244, 162-168)
  gradients                          closed form: one fused backward sweep for the
                                     n x n part [K2], one (M x M)(M x n) product + fused K_nm backward for
                                     the Nystrom part -- no autograd tape, so only A (and one scratch of
                                     the same size) is resident instead of >= 3 saved M x n tensors.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import Dict, Optional

import torch

from ._ffi import CglbError
from .conjugate_gradient import ConjugateGradient, ConjugateGradientStats, NystromPreconditioner
from .distributed import Shard
from .engine import get_engine

Tensor = torch.Tensor


class ShardedKernelMatvec:
    """(variance K(X,X) + sigma^2 I) @ v over this rank's share of the symmetric work items + all-reduce."""

    def __init__(self, eng, kind, xp, n, d, variance, diag, shard: Shard, xpf=None):
        self.eng, self.kind, self.xp, self.n, self.d = eng, kind, xp, n, d
        self.variance, self.diag, self.shard = float(variance), float(diag), shard
        self.xpf = xpf          # fp32 packed inputs: kernel pairs evaluated in FP32 (fp32 models), else None
        self.count = 0

    def detach(self):
        return self

    def __matmul__(self, v: Tensor) -> Tensor:
        """`A @ x` for x of shape [n] / [n, 1] (one sweep) or [n, t] (conjugate_gradient.py:57,66,72: "[N, t]"): the block
        goes through the multi-RHS sweep, which evaluates every kernel pair once for all t columns."""
        vd = v.detach()
        if vd.dim() == 2 and vd.shape[1] > 1:
            if self.xpf is None and self.d <= 32:
                y = self.eng.kmv_sym_multi(self.kind, self.xp, self.n, self.d, vd.contiguous(), self.variance, self.diag,
                                           part=self.shard.rank, nparts=self.shard.world)
                self.shard.all_reduce(y)
                self.count += 1
                return y
            return torch.stack([(self @ vd[:, j]).reshape(-1) for j in range(vd.shape[1])], dim=1)      # wide / fp32-pair inputs
        sweep = self.eng.kmv_sym if self.xpf is None else self.eng.kmv_sym_f32
        y = sweep(self.kind, self.xp if self.xpf is None else self.xpf, self.n, self.d, vd.reshape(-1).contiguous(),
                  self.variance, self.diag, part=self.shard.rank, nparts=self.shard.world)
        self.shard.all_reduce(y)
        self.count += 1
        return y.reshape(v.shape)


@dataclass
class CommonTermsDev:
    """models.py:90-95 plus what the fused backward needs."""
    A: Tensor            # [M, ld] buffer; this rank's column block in [:, :ncols]
    LB: Tensor           # [M, M]
    AAt_diag_sum: Tensor
    L: Tensor            # [M, M]
    AAt: Tensor
    LBinv: Tensor
    zp: Tensor
    ncols: int


@dataclass
class BoundOutput:
    bound: float
    upper: float          # -upper_bound of models.py:286
    lower: float
    logdet: float
    cg_stats: Optional[ConjugateGradientStats]
    matvecs: int
    grads: Optional[Dict[str, Tensor]]


class BoundEvaluator:
    def __init__(self, x: Tensor, y: Tensor, shard: Optional[Shard] = None, pair_dtype: str = "f64"):
        """pair_dtype "f32": the n^2 kernel-pair evaluations of the K*v sweeps run in FP32 (the reference's fp32
        switch, interface.py:96-110); inputs, accumulation and everything M-sized stay FP64."""
        if pair_dtype not in ("f64", "f32"):
            raise CglbError(f"pair_dtype must be 'f64' or 'f32', got {pair_dtype!r}")
        if not x.is_cuda:
            raise CglbError("BoundEvaluator needs CUDA tensors (cglb_b200 has no CPU fallback)")
        if x.dtype != torch.float64:
            raise CglbError("the sm_100a kernels are fp64; fp32 models are promoted by the caller")
        self.eng = get_engine(x.device)
        self.x = x.detach().contiguous()
        self.y = y.detach().reshape(-1).contiguous()
        self.n, self.d = self.x.shape
        self.shard = shard if shard is not None else Shard.from_env()
        self.shift = self.x.mean(0).contiguous()
        self.lo, self.hi = self.shard.column_block(self.n)
        self.ncols = self.hi - self.lo
        self.ld = max(16, (self.ncols + 15) // 16 * 16)
        self.dp = self.eng.packed_width(self.d)
        self.xp = self.eng.empty(self.eng.padded_rows(self.n), self.dp)
        # fp32-pair mode exists for the register-resident dimensions only; wider inputs keep the FP64 DMMA sweep
        self.pair_dtype = pair_dtype if self.d <= 32 else "f64"
        self.xpf = None
        self._A = None
        self._T = None
        self.terms: Optional[CommonTermsDev] = None
        self._packed_key = None
        # True (default): K v, r and P r after the CG solve come from the final CG state (k + 1 + floor(k/restart) sweeps per
        # evaluation).  False (CGLB_RECOMPUTE_RESIDUAL=1): recompute them as models.py:280-282 does (one more n^2 sweep).
        # Both routes are compared with the reference's golden vectors at the north_star tolerance (1e-7); with the
        # fixed-order reductions they return the same v, and their gradients differ by ~1e-14 (profiles/grad_spread_r02.md).
        self.reuse_cg_state = os.environ.get("CGLB_RECOMPUTE_RESIDUAL", "0") in ("", "0")

    def refresh_data(self, x: Tensor, y: Tensor):
        """New contents for the same problem shape (workspaces are kept)."""
        self.x = x.detach().contiguous()
        self.y = y.detach().reshape(-1).contiguous()
        self.shift = self.x.mean(0).contiguous()

    # ------------------------------------------------------------------------------------------------
    def _buffers(self, m: int, need_t: bool):
        if self._A is None or self._A.shape[0] != m:
            self._A = self.eng.empty(m, self.ld)
            self._T = None
        if need_t and self._T is None:
            self._T = self.eng.empty(m, self.ld)

    def pack(self, kind: str, lengthscale: Tensor):
        self.eng.pack(kind, self.x, lengthscale, self.shift, out=self.xp)
        if self.pair_dtype == "f32":
            self.xpf = self.eng.pack_f32(kind, self.x, lengthscale, self.shift, out=self.xpf)

    def common_terms(self, kind: str, Z: Tensor, lengthscale: Tensor, variance: float, noise: float, jitter: float) -> CommonTermsDev:
        """models.py:176-213."""
        eng, m = self.eng, Z.shape[0]
        self._buffers(m, need_t=False)
        sigma = math.sqrt(noise)
        self.pack(kind, lengthscale)
        zp = eng.pack(kind, Z.detach().contiguous(), lengthscale, self.shift)
        A = self._A
        # K_uf for this rank's columns straight into the A buffer                       (:196-197)
        eng.knm_build(kind, zp, m, self.xp[self.lo:], self.ncols, self.d, variance, A, self.ld)
        kuu = eng.empty(m, m)
        eng.knm_build(kind, zp, m, zp, m, self.d, variance, kuu, kuu.stride(0))         # :200
        kuu.diagonal().add_(jitter)                                                     # :201
        L = eng.potrf(kuu, "K_uu + jitter")                                             # :202
        eng.trsm_left_lower(L, A, self.ncols, alpha=1.0 / sigma)                        # :206
        AAt = eng.empty(m, m)
        eng.syrk(A, m, self.ncols, AAt)                                                 # :207
        self.shard.all_reduce(AAt)
        B = AAt.clone()
        B.diagonal().add_(1.0)                                                          # :208-209
        LB = eng.potrf(B, "I + A A^T")                                                  # :210
        LBinv = eng.tri_inverse(LB)
        self.terms = CommonTermsDev(A=A, LB=LB, AAt_diag_sum=AAt.diagonal().sum(), L=L, AAt=AAt, LBinv=LBinv, zp=zp,
                                    ncols=self.ncols)
        return self.terms

    def preconditioner(self, terms: CommonTermsDev, noise: float) -> NystromPreconditioner:
        return NystromPreconditioner(terms.A, terms.LB, noise, shard=self.shard, cols=(self.lo, self.hi), lbinv=terms.LBinv)

    def operator(self, kind: str, variance: float, noise: float) -> ShardedKernelMatvec:
        return ShardedKernelMatvec(self.eng, kind, self.xp, self.n, self.d, variance, noise, self.shard,
                                   xpf=self.xpf if self.pair_dtype == "f32" else None)

    # ------------------------------------------------------------------------------------------------
    def evaluate(self, kind: str, Z: Tensor, lengthscale: Tensor, variance: float, noise: float, mean_c: float,
                 v_vec: Tensor, cg_opt: ConjugateGradient, jitter: float, use_cached_v: bool = False,
                 need_grad: bool = True) -> BoundOutput:
        """One bound (+ gradient) evaluation.  `v_vec` [n,1] is the warm start and is updated IN PLACE with
        the CG solution (models.py:274).  Gradients are those of the BOUND w.r.t. the constrained values."""
        eng, n, d = self.eng, self.n, self.d
        m = Z.shape[0]
        lengthscale = lengthscale.detach().reshape(-1).contiguous()
        terms = self.common_terms(kind, Z, lengthscale, variance, noise, jitter)
        op = self.operator(kind, variance, noise)
        precon = self.preconditioner(terms, noise)
        err = (self.y - mean_c).reshape(-1, 1)                                          # models.py:253-254
        cg_stats, state = None, None
        if use_cached_v:
            v = v_vec
        else:
            solve = getattr(cg_opt, "solve", None)
            if solve is not None and self.reuse_cg_state:
                v, cg_stats, state = solve(op, err, v_vec, precon)                      # :265-270
            else:
                v, cg_stats = cg_opt(op, err, v_vec, precon)
            v_vec.data.copy_(v)                                                         # :274
        v = v.reshape(-1, 1).contiguous()
        r = torch.empty_like(v)
        scal = eng.empty(1)
        if state is None:
            Kv = op @ v                                                                 # :280
            eng.quad_terms(n, err, Kv, v, r, scal)                                      # :281, :283
            z, eb = precon(r)                                                           # :282
        else:
            # The reference recomputes cov @ v, r and P r after the solve because autograd needs them on the tape.
            # The gradients here are closed-form, and the loop already holds all three for the returned v: its
            # residual (recomputed as b - K v at the start and at every restart, r -= gamma K p in between), and
            # z = P r, r^T z of its last preconditioner application (whose B^-1 A r the backward reads from
            # `precon.w`).  K v = err - r saves one n^2 sweep and two passes over A per evaluation; the residual
            # gap of the recurrence is O(eps k |K||v|), the size of the rounding error of one K v itself
            # (tests/test_gpu_parity.py compares both routes; CGLB_RECOMPUTE_RESIDUAL=1 selects the reference's).
            Kv = torch.sub(err, state.r.reshape(-1, 1))
            eng.quad_terms(n, err, Kv, v, r, scal)
            z, eb = state.z.reshape(-1, 1), state.rz
        lower = float(scal.item())
        eb = float(eb.item())
        upper = lower + 0.5 * eb                                                        # :284
        # log-det term, models.py:215-244
        tr_aat = float(terms.AAt_diag_sum.item())
        sum_log_diag = float(terms.LB.diagonal().log().sum().item())
        t = n * variance / noise - tr_aat                                               # :236 (kdiag = variance)
        logdet = -sum_log_diag - 0.5 * n * math.log(noise) - 0.5 * n * math.log(1.0 + t / n)
        const = -0.5 * n * math.log(2.0 * math.pi)                                      # :162-163
        bound = -upper + logdet + const                                                 # :168
        grads = None
        if need_grad:
            grads = self._gradients(kind, terms, precon, lengthscale, variance, noise, v, z, eb, t)
        if self.shard.world > 1:
            # replicated pieces (M x M algebra, K_uu backward) use atomics whose summation order differs
            # between ranks: publish rank 0's numbers so that every rank's optimiser sees identical values
            # and takes identical L-BFGS / CG branches
            keys = ["noise", "mean_c", "Z", "variance", "lengthscale"]
            flat = [torch.tensor([bound, upper, lower, logdet], dtype=torch.float64, device=self.x.device)]
            if grads is not None:
                flat += [grads[k].reshape(-1).to(torch.float64) for k in keys]
            buf = torch.cat(flat)
            self.shard.broadcast(buf, src=0)
            bound, upper, lower, logdet = (float(t_) for t_ in buf[:4].tolist())
            if grads is not None:
                off = 4
                for k in keys:
                    cnt = grads[k].numel()
                    grads[k] = buf[off:off + cnt].reshape(grads[k].shape)
                    off += cnt
        return BoundOutput(bound=bound, upper=-upper, lower=-lower, logdet=logdet, cg_stats=cg_stats,
                           matvecs=op.count, grads=grads)

    # ------------------------------------------------------------------------------------------------
    def _gradients(self, kind, terms, precon, lengthscale, variance, noise, v, z, eb, t) -> Dict[str, Tensor]:
        """Closed-form gradient of the bound for fixed v (checked against autograd of
        the oracle in tests/test_closed_form_backward.py)."""
        eng, n, d, shard = self.eng, self.n, self.d, self.shard
        m = terms.L.shape[0]
        sigma = math.sqrt(noise)
        a = 1.0 / (1.0 + t / n)
        self._buffers(m, need_t=True)
        vf, zf = v.reshape(-1), z.reshape(-1).contiguous()
        u = torch.add(zf, vf, alpha=0.5)                                                # u = v/2 + z
        # ---- sharded parts: [sweep(d+1) | ls(d) | var(1) | Z(m*d)] in one buffer, one all-reduce
        acc = eng.zeros(2 * d + 2 + m * d)
        sweep, o_ls, o_var, o_z = acc[:d + 1], acc[d + 1:2 * d + 1], acc[2 * d + 1:2 * d + 2], acc[2 * d + 2:]
        if self.pair_dtype == "f32":
            eng.kmv_bwd_sym_f32(kind, self.xpf, self.xp, n, d, u, vf.contiguous(), variance, lengthscale, sweep,
                                part=shard.rank, nparts=shard.world)                    # K2, FP32 kernel pairs
        else:
            eng.kmv_bwd_sym(kind, self.xp, n, d, u, vf.contiguous(), variance, lengthscale, sweep,
                            part=shard.rank, nparts=shard.world)                        # K2
        Linv = eng.tri_inverse(terms.L)
        LinvT = Linv.t().contiguous()
        LBinvT = terms.LBinv.t().contiguous()
        Binv = eng.empty(m, m)
        eng.gemm(LBinvT, LBinvT, Binv, m, m, m, transb=True)                            # B^-1 = LB^-T LB^-1
        aI_minus_Binv = -Binv
        aI_minus_Binv.diagonal().add_(a)
        H = eng.empty(m, m)
        eng.gemm(LinvT, aI_minus_Binv, H, m, m, m, alpha=1.0 / sigma)                   # H = L^-T (aI - B^-1) / sigma
        w = precon.w                                                                    # B^-1 A r
        wt = torch.mv(LinvT, w) / sigma
        if self.ncols > 0:
            eng.gemm(H, terms.A, self._T, m, self.ncols, m)                             # T = H A   (M x M)(M x n)
            eng.knm_backward(kind, terms.zp, m, self.xp[self.lo:], self.ncols, d, variance, lengthscale, self._T, self.ld,
                             wt.contiguous(), zf[self.lo:self.hi], o_ls, o_var, o_z)
        shard.all_reduce(acc)
        # ---- replicated M x M part: dS/dK_uu = -1/2 L^-T Mx L^-1
        Mx = a * terms.AAt + Binv + torch.outer(w, w) / noise
        Mx.diagonal().sub_(1.0)
        tmp = eng.empty(m, m)
        eng.gemm(LinvT, Mx, tmp, m, m, m)
        Gkuu = eng.empty(m, m)
        eng.gemm(tmp, Linv, Gkuu, m, m, m, alpha=-0.5)
        acc2 = eng.zeros(d + 1 + m * d)
        eng.knm_backward(kind, terms.zp, m, terms.zp, m, d, variance, lengthscale, Gkuu, Gkuu.stride(0), None, None,
                         acc2[:d], acc2[d:d + 1], acc2[d + 1:])
        g_ls = sweep[:d] + o_ls + acc2[:d]
        g_var = sweep[d] + o_var[0] + acc2[d] - 0.5 * a * n / noise
        g_Z = (o_z + 2.0 * acc2[d + 1:]).reshape(m, d)
        scal = eng.empty(1)
        eng.dot(u, vf.contiguous(), scal)
        g_noise = scal[0] - n / (2.0 * noise) + 0.5 * a * n * variance / noise ** 2 + 0.5 * eb / noise \
            - Mx.diagonal().sum() / (2.0 * noise)
        g_c = (vf + zf).sum()
        return dict(noise=g_noise, mean_c=g_c, Z=g_Z, variance=g_var, lengthscale=g_ls)
