"""Torch-tensor level wrappers over the C-ABI (include/cglb_b200.h).

`Engine` owns one `cglb_context` per CUDA device and exposes each entry point with torch tensors in
place of raw pointers.  torch is used for device memory, streams and (in distributed.py) NCCL only; all
arithmetic on n-sized or M x n-sized data runs in libcglb_b200.so.  There is no CPU path: tensors must
live on a CUDA device and the shared library must be built, otherwise `CglbError` is raised.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _ffi
from ._ffi import CglbError, KIND_IDS, check, ptr

Tensor = torch.Tensor

_ENGINES: Dict[int, "Engine"] = {}


def get_engine(device=None) -> "Engine":
    """One Engine (one cglb_context) per CUDA device."""
    if not torch.cuda.is_available():
        raise CglbError("cglb_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise CglbError(f"cglb_b200 tensors must live on a CUDA device, got {device}")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    eng = _ENGINES.get(idx)
    if eng is None:
        eng = Engine(idx)
        _ENGINES[idx] = eng
    return eng


def _req(t: Tensor, name: str, dtype=torch.float64):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise CglbError(f"{name}: expected a CUDA tensor (cglb_b200 has no CPU fallback)")
    if t.dtype != dtype:
        raise CglbError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise CglbError(f"{name}: expected a contiguous tensor")
    return t


def _req2d(t: Tensor, name: str):
    """a row-major matrix (or a column slice of one): fp64, CUDA, unit stride along the rows"""
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise CglbError(f"{name}: expected a CUDA tensor (cglb_b200 has no CPU fallback)")
    if t.dtype != torch.float64:
        raise CglbError(f"{name}: expected dtype torch.float64, got {t.dtype}")
    if t.dim() != 2 or t.stride(1) != 1:
        raise CglbError(f"{name}: expected a row-major 2-D tensor")
    return t


class Engine:
    def __init__(self, device_index: int):
        self.lib = _ffi.load_library()
        self.device = torch.device("cuda", device_index)
        handle = C.c_void_p()
        check(self.lib.cglb_create(C.byref(handle), device_index), "cglb_create")
        self.ctx = handle
        self._env_seen = {}
        # optional per-kernel CUDA-event timing (bench.py roofline): name -> list of (start, end) events
        self.timing = None

    # ---- optional timing of individual entry points (events on the launching stream) ---------------
    def enable_timing(self, on: bool = True):
        self.timing = {} if on else None

    def _timed(self, name, fn):
        if self.timing is None:
            return fn()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(self.device))
        out = fn()
        e1.record(torch.cuda.current_stream(self.device))
        self.timing.setdefault(name, []).append((e0, e1))
        return out

    def timing_summary(self):
        """name -> (launches, total ms).  Synchronises."""
        torch.cuda.synchronize(self.device)
        return {k: (len(v), float(sum(a.elapsed_time(b) for a, b in v))) for k, v in (self.timing or {}).items()}

    # ---- developer options ----------------------------------------------------------------------
    def set_option(self, name: str, value: int):
        """cglb_set_option: "dsweep" 0 / 1 / 2, "gemm_staging" 1 (cp.async) / 2 (TMA), "superrow" chunks per super-row (0 = auto)."""
        check(self.lib.cglb_set_option(self.ctx, name.encode(), int(value)), "cglb_set_option")

    _ENV_OPTIONS = {"CGLB_DSWEEP": ("dsweep", 1), "CGLB_GEMM_STAGING": ("gemm_staging", 1), "CGLB_SUPERROW": ("superrow", 0)}

    def _sync_env_options(self):
        """The library reads its switches from the environment once (cglb_create); tests and tools flip them at run time, so
        the Python wrapper forwards a changed value before the calls they affect (no getenv on the C side's dispatch path)."""
        import os
        for var, (name, default) in self._ENV_OPTIONS.items():
            val = os.environ.get(var)
            if self._env_seen.get(var, "__unset__") != val:
                self._env_seen[var] = val
                try:
                    self.set_option(name, int(val) if val not in (None, "") else default)
                except (ValueError, CglbError):
                    self.set_option(name, default)

    # ---- bookkeeping ---------------------------------------------------------------------------
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launch_count(self) -> int:
        return int(self.lib.cglb_launch_count(self.ctx))

    @property
    def num_sms(self) -> int:
        return int(self.lib.cglb_num_sms(self.ctx))

    def packed_width(self, d: int) -> int:
        return int(self.lib.cglb_packed_width(d))

    def padded_rows(self, n: int) -> int:
        return int(self.lib.cglb_padded_rows(n))

    def empty(self, *shape, dtype=torch.float64) -> Tensor:
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=torch.float64) -> Tensor:
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    # ---- K7 prep ---------------------------------------------------------------------------------
    def pack(self, kind: str, x: Tensor, lengthscale: Tensor, shift: Optional[Tensor], out: Optional[Tensor] = None) -> Tensor:
        _req(x, "x"); _req(lengthscale, "lengthscale")
        n, d = x.shape
        if lengthscale.numel() != d:
            raise CglbError("lengthscale must have one entry per input dimension (ARD)")
        if out is None:
            out = self.empty(self.padded_rows(n), self.packed_width(d))
        check(self.lib.cglb_pack_inputs(self.ctx, KIND_IDS[kind], ptr(x), n, d, ptr(lengthscale),
                                        ptr(shift) if shift is not None else None, ptr(out), self.stream()), "cglb_pack_inputs")
        return out

    # ---- K1 / K2 ------------------------------------------------------------------------------------
    KMV_VARIANTS = {0: "kmv_sweep_kernel (register-resident DFMA sweep)", 1: "dmma_sweep_kernel (DMMA-distance sweep, d <= 32)",
                    2: "wide_sweep_kernel (DMMA sweep, d > 32)"}

    def kmv_sym_variant(self, d: int, n: int, nparts: int = 1) -> int:
        """which kernel cglb_kmv_sym launches for this shape (include/cglb_b200.h)"""
        self._sync_env_options()
        return int(self.lib.cglb_kmv_sym_variant(self.ctx, int(d), int(n), int(nparts)))

    def kmv_sym(self, kind, xp, n, d, v, variance, diag, out=None, part=0, nparts=1) -> Tensor:
        _req(xp, "xp"); _req(v, "v")
        self._sync_env_options()
        if out is None:
            out = self.empty(n)
        self._timed("kmv_sym", lambda: check(self.lib.cglb_kmv_sym(
            self.ctx, KIND_IDS[kind], ptr(xp), n, d, ptr(v), ptr(out), float(variance), float(diag), int(part), int(nparts),
            self.stream()), "cglb_kmv_sym"))
        return out

    def kmv_sym_multi(self, kind, xp, n, d, v, variance, diag, out=None, part=0, nparts=1) -> Tensor:
        """Y = variance K(X,X) V + diag V for V of shape [n, t] (the reference's `A @ x`, x: [N, t])."""
        _req(xp, "xp"); _req(v, "v")
        if v.dim() != 2 or v.shape[0] != n:
            raise CglbError(f"v: expected shape [{n}, t], got {tuple(v.shape)}")
        t = int(v.shape[1])
        if out is None:
            out = self.empty(n, t)
        _req(out, "out")
        self._timed("kmv_sym_multi", lambda: check(self.lib.cglb_kmv_sym_multi(
            self.ctx, KIND_IDS[kind], ptr(xp), n, d, ptr(v), t, ptr(out), float(variance), float(diag), int(part), int(nparts),
            self.stream()), "cglb_kmv_sym_multi"))
        return out

    # ---- fp32-pair mode (models created under set_default_float("fp32")) --------------------------------
    def pack_f32(self, kind: str, x: Tensor, lengthscale: Tensor, shift: Optional[Tensor], out: Optional[Tensor] = None) -> Tensor:
        _req(x, "x"); _req(lengthscale, "lengthscale")
        n, d = x.shape
        if out is None:
            out = self.empty(self.padded_rows(n), int(self.lib.cglb_packed_width_f32(d)), dtype=torch.float32)
        check(self.lib.cglb_pack_inputs_f32(self.ctx, KIND_IDS[kind], ptr(x), n, d, ptr(lengthscale),
                                            ptr(shift) if shift is not None else None, ptr(out), self.stream()), "cglb_pack_inputs_f32")
        return out

    def kmv_sym_f32(self, kind, xpf, n, d, v, variance, diag, out=None, part=0, nparts=1) -> Tensor:
        _req(xpf, "xpf", torch.float32); _req(v, "v")
        if out is None:
            out = self.empty(n)
        self._timed("kmv_sym", lambda: check(self.lib.cglb_kmv_sym_f32(
            self.ctx, KIND_IDS[kind], ptr(xpf), n, d, ptr(v), ptr(out), float(variance), float(diag), int(part), int(nparts),
            self.stream()), "cglb_kmv_sym_f32"))
        return out

    def kmv_bwd_sym_f32(self, kind, xpf, xp, n, d, u, w, variance, lengthscale, out, part=0, nparts=1) -> Tensor:
        _req(xpf, "xpf", torch.float32); _req(xp, "xp"); _req(u, "u"); _req(w, "w"); _req(lengthscale, "lengthscale"); _req(out, "out")
        self._timed("kmv_bwd_sym", lambda: check(self.lib.cglb_kmv_bwd_sym_f32(
            self.ctx, KIND_IDS[kind], ptr(xpf), ptr(xp), n, d, ptr(u), ptr(w), float(variance), ptr(lengthscale), ptr(out),
            int(part), int(nparts), self.stream()), "cglb_kmv_bwd_sym_f32"))
        return out

    def kmv_rect(self, kind, xp_rows, nrows, xp_cols, ncols, d, v, variance, out=None) -> Tensor:
        _req(xp_rows, "xp_rows"); _req(xp_cols, "xp_cols"); _req(v, "v")
        if out is None:
            out = self.empty(nrows)
        check(self.lib.cglb_kmv_rect(self.ctx, KIND_IDS[kind], ptr(xp_rows), nrows, ptr(xp_cols), ncols, d, ptr(v), ptr(out),
                                     float(variance), self.stream()), "cglb_kmv_rect")
        return out

    def kmv_bwd_sym(self, kind, xp, n, d, u, w, variance, lengthscale, out, part=0, nparts=1) -> Tensor:
        _req(xp, "xp"); _req(u, "u"); _req(w, "w"); _req(lengthscale, "lengthscale"); _req(out, "out")
        self._sync_env_options()
        self._timed("kmv_bwd_sym", lambda: check(self.lib.cglb_kmv_bwd_sym(
            self.ctx, KIND_IDS[kind], ptr(xp), n, d, ptr(u), ptr(w), float(variance), ptr(lengthscale), ptr(out), int(part),
            int(nparts), self.stream()), "cglb_kmv_bwd_sym"))
        return out

    # ---- K7 -------------------------------------------------------------------------------------------
    def knm_build(self, kind, zp, m, xp, n, d, variance, out, ld) -> Tensor:
        _req(zp, "zp"); _req(xp, "xp"); _req(out, "out")
        check(self.lib.cglb_knm_build(self.ctx, KIND_IDS[kind], ptr(zp), m, ptr(xp), n, d, float(variance), ptr(out), ld,
                                      self.stream()), "cglb_knm_build")
        return out

    def knm_backward(self, kind, zp, m, xp, ncols, d, variance, lengthscale, t, ldt, wt, zvec, out_ls, out_var, out_z):
        check(self.lib.cglb_knm_backward(self.ctx, KIND_IDS[kind], ptr(zp), m, ptr(xp), ncols, d, float(variance),
                                         ptr(lengthscale), ptr(t), ldt, ptr(wt), ptr(zvec), ptr(out_ls), ptr(out_var),
                                         ptr(out_z), self.stream()), "cglb_knm_backward")

    # ---- K5 / K6 --------------------------------------------------------------------------------------
    def potrf(self, a: Tensor, what: str = "matrix") -> Tensor:
        """In-place lower Cholesky of a square matrix; raises like torch.cholesky on failure
        (the reference lets the RuntimeError propagate)."""
        _req(a, "a")
        m = a.shape[0]
        info = torch.zeros(1, dtype=torch.int32, device=self.device)
        check(self.lib.cglb_potrf(self.ctx, ptr(a), m, a.stride(0), ptr(info), self.stream()), "cglb_potrf")
        bad = int(info.item())
        if bad != 0:
            raise RuntimeError(f"cholesky: {what} is not positive definite (block {bad - 1})")
        return a

    def tri_inverse(self, l: Tensor, out: Optional[Tensor] = None) -> Tensor:
        _req(l, "l")
        m = l.shape[0]
        if out is None:
            out = self.empty(m, m)
        check(self.lib.cglb_tri_inverse(self.ctx, ptr(l), m, l.stride(0), ptr(out), out.stride(0), self.stream()), "cglb_tri_inverse")
        return out

    def trsm_left_lower(self, l: Tensor, b: Tensor, n: int, alpha: float = 1.0) -> Tensor:
        _req(l, "l"); _req(b, "b")
        self._timed("trsm", lambda: check(self.lib.cglb_trsm_left_lower(
            self.ctx, ptr(l), l.shape[0], l.stride(0), ptr(b), n, b.stride(0), float(alpha), self.stream()), "cglb_trsm_left_lower"))
        return b

    def syrk(self, a: Tensor, m: int, n: int, out: Tensor, accumulate: bool = False) -> Tensor:
        _req(a, "a"); _req(out, "out")
        self._timed("syrk", lambda: check(self.lib.cglb_syrk(
            self.ctx, ptr(a), m, n, a.stride(0), ptr(out), out.stride(0), int(accumulate), self.stream()), "cglb_syrk"))
        return out

    def gemm(self, a: Tensor, b: Tensor, out: Tensor, m: int, n: int, k: int, transb: bool = False, alpha: float = 1.0,
             beta: float = 0.0) -> Tensor:
        _req(a, "a"); _req(b, "b"); _req(out, "out")
        self._sync_env_options()
        self._timed("gemm", lambda: check(self.lib.cglb_gemm(
            self.ctx, int(transb), m, n, k, float(alpha), ptr(a), a.stride(0), ptr(b), b.stride(0), float(beta), ptr(out),
            out.stride(0), self.stream()), "cglb_gemm"))
        return out

    # ---- K3 / K4 --------------------------------------------------------------------------------------
    def precond_project(self, a: Tensor, m: int, ncols: int, r: Tensor, q: Tensor) -> Tensor:
        _req2d(a, "a"); _req(r, "r"); _req(q, "q")
        self._timed("precond_project", lambda: check(self.lib.cglb_precond_project(
            self.ctx, ptr(a), m, ncols, a.stride(0), ptr(r), ptr(q), self.stream()), "cglb_precond_project"))
        return q

    def precond_finish(self, a: Tensor, m: int, ncols: int, lbinv: Tensor, q: Tensor, r: Tensor, sigma_sq: float,
                       z: Tensor, w: Tensor, rz: Tensor):
        _req2d(a, "a"); _req(lbinv, "lbinv"); _req(q, "q"); _req(r, "r"); _req(z, "z"); _req(w, "w"); _req(rz, "rz")
        self._timed("precond_finish", lambda: check(self.lib.cglb_precond_finish(
            self.ctx, ptr(a), m, ncols, a.stride(0), ptr(lbinv), ptr(q), ptr(r), float(sigma_sq), ptr(z), ptr(w), ptr(rz),
            self.stream()), "cglb_precond_finish"))

    # ---- K8 -------------------------------------------------------------------------------------------
    def dot(self, x: Tensor, y: Tensor, out: Tensor) -> Tensor:
        _req(x, "x"); _req(y, "y"); _req(out, "out")
        if x.numel() != y.numel():
            raise CglbError("dot: operands differ in length")
        check(self.lib.cglb_dot(self.ctx, ptr(x), ptr(y), x.numel(), ptr(out), self.stream()), "cglb_dot")
        return out

    def cg_step(self, n, rz, pAp, p, Ap, v, r, restart: bool):
        for name, t in (("rz", rz), ("pAp", pAp), ("p", p), ("Ap", Ap), ("v", v), ("r", r)):
            _req(t, name)
        check(self.lib.cglb_cg_step(self.ctx, n, ptr(rz), ptr(pAp), ptr(p), ptr(Ap), ptr(v), ptr(r), int(restart), self.stream()), "cglb_cg_step")

    def residual(self, n, b, Av, r):
        _req(b, "b"); _req(Av, "Av"); _req(r, "r")
        check(self.lib.cglb_residual(self.ctx, n, ptr(b), ptr(Av), ptr(r), self.stream()), "cglb_residual")

    def cg_direction(self, n, z, p, rz_new, rz_old, restart: bool):
        _req(z, "z"); _req(p, "p"); _req(rz_new, "rz_new"); _req(rz_old, "rz_old")
        check(self.lib.cglb_cg_direction(self.ctx, n, ptr(z), ptr(p), ptr(rz_new), ptr(rz_old), int(restart), self.stream()), "cglb_cg_direction")

    def quad_terms(self, n, err, Kv, v, r, out):
        for name, t in (("err", err), ("Kv", Kv), ("v", v), ("r", r), ("out", out)):
            _req(t, name)
        check(self.lib.cglb_quad_terms(self.ctx, n, ptr(err), ptr(Kv), ptr(v), ptr(r), ptr(out), self.stream()), "cglb_quad_terms")
