"""GPU (-m gpu): parity of the sm_100a path, called through the C-ABI, against the CPU oracle and the golden
vectors produced by the reference's own code.  Tolerances are the north_star's: per-matvec relative error
<= 1e-10, bound and gradients <= 1e-7 relative, CG iteration count within +-1."""
import math
import os

import numpy as np
import pytest
import torch

import cglb_b200 as cb
from conftest import GOLDEN_CASES, GOLDEN_DIR, GRAD_NAMES
from helpers import make_model, rel_max
from oracle import cglb_oracle as o

pytestmark = pytest.mark.gpu
f64 = torch.float64
MATVEC_TOL = 1e-10
BOUND_TOL = 1e-7


@pytest.fixture(scope="module")
def eng():
    from cglb_b200.engine import get_engine
    return get_engine()


def _problem(n, d, seed, lsval=None):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g, dtype=f64)
    v = torch.randn(n, generator=g, dtype=f64)
    u = torch.randn(n, generator=g, dtype=f64)
    ls = (torch.rand(d, generator=g, dtype=f64) + 0.5) * (lsval if lsval else 0.5 * math.sqrt(d))
    return x, v, u, ls


# ---- K1: matvec ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,n,d", [("matern32", 1, 1), ("matern32", 2, 3), ("matern32", 300, 1), ("rbf", 777, 8),
                                      ("matern32", 2500, 11), ("rbf", 2049, 3), ("matern32", 1025, 5), ("rbf", 900, 16),
                                      ("matern32", 640, 20), ("rbf", 513, 32), ("matern32", 1300, 13),
                                      # d > 32: DMMA distance contraction (song-shaped config)
                                      ("matern32", 1, 40), ("matern32", 300, 33), ("rbf", 1500, 64), ("matern32", 2300, 90),
                                      ("rbf", 129, 100)])
def test_kmv_sym_matches_oracle(eng, kind, n, d):
    x, v, u, ls = _problem(n, d, seed=n + d)
    dev = eng.device
    xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    y = eng.kmv_sym(kind, xp, n, d, v.to(dev), 1.3, 0.07)
    K = o.kernel_dense(kind, x, x, ls, torch.tensor(1.3, dtype=f64))
    ref = K @ v + 0.07 * v
    assert float((y.cpu() - ref).norm() / ref.norm()) <= MATVEC_TOL
    # work-item partition (what each rank of a row-sharded run computes) sums to the full product
    parts = sum(eng.kmv_sym(kind, xp, n, d, v.to(dev), 1.3, 0.07, part=p, nparts=3) for p in range(3))
    assert float((parts.cpu() - ref).norm() / ref.norm()) <= MATVEC_TOL


# ---- K1 on DMMA (dsweep_impl.cuh): the path the n >= 24k, 10 <= d <= 32 sweeps take -------------------------
@pytest.mark.parametrize("kind,n,d,lsval", [("matern32", 1, 11, None), ("matern32", 257, 11, None), ("matern32", 2500, 11, None),
                                            ("rbf", 1500, 10, 2.0), ("matern32", 3333, 19, 2.0), ("rbf", 1111, 27, 3.0),
                                            ("rbf", 1300, 18, None), ("matern32", 900, 26, None), ("matern32", 1300, 2, 0.7),
                                            ("rbf", 2049, 3, None), ("matern32", 1500, 11, 0.05), ("matern32", 1500, 10, 30.0)])
def test_dmma_sweep_matches_oracle(eng, monkeypatch, kind, n, d, lsval):
    """CGLB_DSWEEP=2 forces the DMMA-distance sweep on shapes the size rule would give to the register kernel."""
    monkeypatch.setenv("CGLB_DSWEEP", "2")
    x, v, u, ls = _problem(n, d, seed=n + d, lsval=lsval)
    dev = eng.device
    xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    K = o.kernel_dense(kind, x, x, ls, torch.tensor(1.3, dtype=f64))
    ref = K @ v + 0.07 * v
    y = eng.kmv_sym(kind, xp, n, d, v.to(dev), 1.3, 0.07)
    assert float((y.cpu() - ref).norm() / ref.norm()) <= MATVEC_TOL
    parts = sum(eng.kmv_sym(kind, xp, n, d, v.to(dev), 1.3, 0.07, part=p, nparts=3) for p in range(3))
    assert float((parts.cpu() - ref).norm() / ref.norm()) <= MATVEC_TOL
    monkeypatch.setenv("CGLB_DSWEEP", "0")
    y0 = eng.kmv_sym(kind, xp, n, d, v.to(dev), 1.3, 0.07)
    # both are expanded-form kernels: at the tiny-lengthscale case their cancellation errors (|a|^2 ~ 1e5) differ
    assert float((y - y0).norm() / y0.norm()) <= (1e-12 if lsval is None else 1e-11)


@pytest.mark.parametrize("d", list(range(2, 33)))
def test_dmma_sweep_every_dimension(eng, monkeypatch, d):
    """every packed width: multiples of 4, widths whose last k-step reads into the next packed row (d = 12, 13,
    16, 17 ...), ragged n (last tile and last row block partly empty)."""
    monkeypatch.setenv("CGLB_DSWEEP", "2")
    kind = "matern32" if d % 2 else "rbf"
    n = 700 + 13 * d
    x, v, u, ls = _problem(n, d, seed=100 + d)
    dev = eng.device
    xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    ref = o.kernel_dense(kind, x, x, ls, torch.tensor(0.9, dtype=f64)) @ v + 0.3 * v
    y = eng.kmv_sym(kind, xp, n, d, v.to(dev), 0.9, 0.3)
    assert float((y.cpu() - ref).norm() / ref.norm()) <= MATVEC_TOL


def test_kmv_sym_variant_query(eng, monkeypatch):
    """cglb_kmv_sym_variant names the kernel cglb_kmv_sym launches: 0 register DFMA, 1 DMMA-distance, 2 wide DMMA."""
    monkeypatch.delenv("CGLB_DSWEEP", raising=False)
    assert eng.kmv_sym_variant(11, 2000000, 1) == 1 and eng.kmv_sym_variant(11, 2000000, 8) == 1
    assert eng.kmv_sym_variant(11, 5000, 1) == 0          # too few work items for the DMMA sweep
    assert eng.kmv_sym_variant(3, 434000, 1) == 0 and eng.kmv_sym_variant(8, 40000, 1) == 0 and eng.kmv_sym_variant(12, 500000, 1) == 0
    assert eng.kmv_sym_variant(90, 515000, 1) == 2
    monkeypatch.setenv("CGLB_DSWEEP", "0")
    assert eng.kmv_sym_variant(11, 2000000, 1) == 0


def test_dmma_sweep_duplicates_and_midsize(eng, monkeypatch):
    dev = eng.device
    # exact duplicates: zero and slightly negative expanded-form distances
    x = torch.randn(40, 11, dtype=f64, generator=torch.Generator().manual_seed(1)).repeat(8, 1)
    n, d = x.shape
    ls = torch.full((d,), 1.2, dtype=f64)
    v = torch.randn(n, dtype=f64, generator=torch.Generator().manual_seed(2))
    monkeypatch.setenv("CGLB_DSWEEP", "2")
    for kind in ("matern32", "rbf"):
        xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
        y = eng.kmv_sym(kind, xp, n, d, v.to(dev), 2.0, 0.5)
        ref = o.kernel_dense(kind, x, x, ls, torch.tensor(2.0, dtype=f64)) @ v + 0.5 * v
        assert float((y.cpu() - ref).norm() / ref.norm()) <= MATVEC_TOL
    # mid size, n not a multiple of anything: the size rule picks the DMMA sweep by itself; it must agree with the
    # register-resident sweep and its 8-way partition must sum to the full product
    n, d = 60001, 11
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, d, generator=g, dtype=f64, device=dev)
    v = torch.randn(n, generator=g, dtype=f64, device=dev)
    ls = torch.full((d,), 1.5, dtype=f64, device=dev)
    xp = eng.pack("matern32", x, ls, x.mean(0))
    monkeypatch.delenv("CGLB_DSWEEP")
    y1 = eng.kmv_sym("matern32", xp, n, d, v, 1.0, 0.01)
    parts = sum(eng.kmv_sym("matern32", xp, n, d, v, 1.0, 0.01, part=p, nparts=8) for p in range(8))
    monkeypatch.setenv("CGLB_DSWEEP", "0")
    y0 = eng.kmv_sym("matern32", xp, n, d, v, 1.0, 0.01)
    assert float((y1 - y0).norm() / y0.norm()) <= 1e-12
    assert float((parts - y0).norm() / y0.norm()) <= 1e-12


# ---- K1 in fp32-pair mode (f32sweep_impl.cuh): the fp32 switch of the reference ------------------------------------
F32_TOL = 5e-6      # kernel pairs in FP32 (expanded-form distances, MUFU rsqrt/ex2); accumulation in FP64


@pytest.mark.parametrize("kind,n,d", [("matern32", 1, 1), ("matern32", 2, 3), ("matern32", 300, 1), ("rbf", 777, 8),
                                      ("matern32", 2500, 11), ("rbf", 2049, 3), ("matern32", 1025, 5), ("rbf", 900, 16),
                                      ("matern32", 640, 20), ("rbf", 513, 32), ("matern32", 4100, 13)])
def test_kmv_sym_f32_matches_oracle(eng, kind, n, d):
    x, v, u, ls = _problem(n, d, seed=n + d)
    dev = eng.device
    xpf = eng.pack_f32(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    assert xpf.dtype == torch.float32 and xpf.shape[1] == (d + 4) // 4 * 4
    K = o.kernel_dense(kind, x, x, ls, torch.tensor(1.3, dtype=f64))
    ref = K @ v + 0.07 * v
    y = eng.kmv_sym_f32(kind, xpf, n, d, v.to(dev), 1.3, 0.07)
    assert y.dtype == f64
    assert float((y.cpu() - ref).norm() / ref.norm()) <= F32_TOL
    parts = sum(eng.kmv_sym_f32(kind, xpf, n, d, v.to(dev), 1.3, 0.07, part=p, nparts=3) for p in range(3))
    assert float((parts.cpu() - ref).norm() / ref.norm()) <= F32_TOL


@pytest.mark.parametrize("kind,n,d", [("matern32", 300, 1), ("rbf", 777, 8), ("matern32", 2500, 11), ("rbf", 2049, 3),
                                      ("matern32", 640, 20), ("rbf", 513, 32), ("matern32", 4100, 5)])
def test_kmv_bwd_f32_matches_autograd_of_oracle(eng, kind, n, d):
    """fp32-pair backward sweep: d(u^T K w)/d{lengthscale, variance} against autograd through the fp64 oracle.
    Tolerance 2e-4 of the gradient norm (FP32 kernel pairs and FP32 per-tile partial sums)."""
    x, v, u, ls = _problem(n, d, seed=n + 3 * d)
    dev = eng.device
    xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    xpf = eng.pack_f32(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    out = eng.zeros(d + 1)
    eng.kmv_bwd_sym_f32(kind, xpf, xp, n, d, u.to(dev), v.to(dev), 1.3, ls.to(dev), out)
    lsr = ls.clone().requires_grad_(True)
    varr = torch.tensor(1.3, dtype=f64, requires_grad=True)
    f = u @ (o.kernel_dense(kind, x, x, lsr, varr) @ v)
    gl, gv = torch.autograd.grad(f, [lsr, varr])
    ref = torch.cat([gl.reshape(-1), gv.reshape(1)])
    assert float((out.cpu() - ref).norm() / ref.norm()) <= 2e-4
    # partition of the work items sums to the full gradient
    parts = eng.zeros(d + 1)
    for p in range(3):
        eng.kmv_bwd_sym_f32(kind, xpf, xp, n, d, u.to(dev), v.to(dev), 1.3, ls.to(dev), parts, part=p, nparts=3)
    assert float((parts.cpu() - ref).norm() / ref.norm()) <= 2e-4


def test_kmv_sym_f32_midsize_and_errors(eng):
    dev = eng.device
    n, d = 150001, 11
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, d, generator=g, dtype=f64, device=dev)
    v = torch.randn(n, generator=g, dtype=f64, device=dev)
    ls = torch.full((d,), 1.5, dtype=f64, device=dev)
    xp, xpf = eng.pack("matern32", x, ls, x.mean(0)), eng.pack_f32("matern32", x, ls, x.mean(0))
    y64 = eng.kmv_sym("matern32", xp, n, d, v, 1.0, 0.01)
    y32 = eng.kmv_sym_f32("matern32", xpf, n, d, v, 1.0, 0.01)
    assert float((y32 - y64).norm() / y64.norm()) <= F32_TOL
    parts = sum(eng.kmv_sym_f32("matern32", xpf, n, d, v, 1.0, 0.01, part=p, nparts=8) for p in range(8))
    assert float((parts - y64).norm() / y64.norm()) <= F32_TOL
    with pytest.raises(cb.CglbError):       # fp64 packed array handed to the fp32 entry
        eng.kmv_sym_f32("matern32", xp, n, d, v, 1.0, 0.01)
    with pytest.raises(cb.CglbError):       # d > 32 has no fp32-pair sweep: fails loudly
        eng.kmv_sym_f32("matern32", torch.zeros(128, 44, dtype=torch.float32, device=dev), 100, 40, v[:100].contiguous(), 1.0, 0.0)


@pytest.mark.parametrize("kind,d,lsval", [("matern32", 3, 0.05), ("rbf", 3, 0.05), ("matern32", 8, 30.0)])
def test_kmv_extreme_lengthscales(eng, kind, d, lsval):
    """expanded-form distances must survive tiny lengthscales (huge |a|^2) and huge ones (K ~ all ones)."""
    n = 1500
    x, v, u, ls = _problem(n, d, seed=5, lsval=lsval)
    dev = eng.device
    xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    y = eng.kmv_sym(kind, xp, n, d, v.to(dev), 1.0, 0.0)
    ref = o.kernel_dense(kind, x, x, ls, torch.tensor(1.0, dtype=f64)) @ v
    assert float((y.cpu() - ref).norm() / ref.norm()) <= MATVEC_TOL


def test_kmv_duplicate_points_and_empty(eng):
    dev = eng.device
    x = torch.randn(40, 2, dtype=f64, generator=torch.Generator().manual_seed(1)).repeat(8, 1)   # exact duplicates
    n, d = x.shape
    ls = torch.tensor([0.9, 1.1], dtype=f64)
    v = torch.randn(n, dtype=f64, generator=torch.Generator().manual_seed(2))
    for kind in ("matern32", "rbf"):
        xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
        y = eng.kmv_sym(kind, xp, n, d, v.to(dev), 2.0, 0.5)
        ref = o.kernel_dense(kind, x, x, ls, torch.tensor(2.0, dtype=f64)) @ v + 0.5 * v
        assert float((y.cpu() - ref).norm() / ref.norm()) <= MATVEC_TOL
    y0 = eng.kmv_sym("rbf", eng.empty(0, 4), 0, 2, eng.empty(0), 1.0, 0.0, out=eng.empty(0))
    assert y0.numel() == 0


@pytest.mark.parametrize("kind,nr,nc,d", [("matern32", 100, 1000, 3), ("rbf", 1500, 333, 8), ("matern32", 64, 3000, 11),
                                          ("matern32", 7, 129, 20), ("matern32", 200, 1500, 90), ("rbf", 1100, 90, 48)])
def test_kmv_rect_matches_oracle(eng, kind, nr, nc, d):
    g = torch.Generator().manual_seed(nr + nc)
    xr, xc = torch.randn(nr, d, generator=g, dtype=f64), torch.randn(nc, d, generator=g, dtype=f64)
    v = torch.randn(nc, generator=g, dtype=f64)
    ls = torch.rand(d, generator=g, dtype=f64) + 0.8
    dev = eng.device
    shift = xc.mean(0).to(dev)
    y = eng.kmv_rect(kind, eng.pack(kind, xr.to(dev), ls.to(dev), shift), nr, eng.pack(kind, xc.to(dev), ls.to(dev), shift), nc, d,
                     v.to(dev), 0.9)
    ref = o.kernel_dense(kind, xr, xc, ls, torch.tensor(0.9, dtype=f64)) @ v
    assert float((y.cpu() - ref).norm() / ref.norm()) <= MATVEC_TOL


# ---- K2: backward sweep ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind,n,d", [("matern32", 300, 1), ("rbf", 777, 8), ("matern32", 2500, 11), ("rbf", 1030, 3),
                                      ("matern32", 520, 20), ("rbf", 300, 32), ("matern32", 1100, 40), ("rbf", 1030, 64),
                                      ("matern32", 1500, 90)])
def test_backward_sweep_matches_autograd(eng, kind, n, d):
    x, v, u, ls = _problem(n, d, seed=3 * n + d)
    dev = eng.device
    xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    out = eng.zeros(d + 1)
    eng.kmv_bwd_sym(kind, xp, n, d, u.to(dev), v.to(dev), 1.3, ls.to(dev), out)
    lsr, varr = ls.clone().requires_grad_(True), torch.tensor(1.3, dtype=f64, requires_grad=True)
    f = u @ (o.kernel_dense(kind, x, x, lsr, varr) @ v)
    gl, gv = torch.autograd.grad(f, [lsr, varr])
    ref = torch.cat([gl.reshape(-1), gv.reshape(1)])
    assert float((out.cpu() - ref).norm() / ref.norm()) <= 1e-9
    parts = eng.zeros(d + 1)
    for p in range(2):
        eng.kmv_bwd_sym(kind, xp, n, d, u.to(dev), v.to(dev), 1.3, ls.to(dev), parts, part=p, nparts=2)
    assert float((parts.cpu() - ref).norm() / ref.norm()) <= 1e-9


@pytest.mark.parametrize("d", [2, 3, 7, 8, 10, 11, 12, 13, 16, 17, 19, 24, 27, 32])
def test_dmma_backward_sweep_matches_autograd(eng, monkeypatch, d):
    """K2 on DMMA (dmma_bwd_kernel): forced onto small ragged shapes; the second DMMA product (cross term) reads
    coordinate slots past the packed row for d not a multiple of 8."""
    monkeypatch.setenv("CGLB_DSWEEP", "2")
    kind = "matern32" if d % 2 else "rbf"
    n = 500 + 37 * d
    x, v, u, ls = _problem(n, d, seed=11 * n + d)
    dev = eng.device
    xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    out = eng.zeros(d + 1)
    eng.kmv_bwd_sym(kind, xp, n, d, u.to(dev), v.to(dev), 1.3, ls.to(dev), out)
    lsr, varr = ls.clone().requires_grad_(True), torch.tensor(1.3, dtype=f64, requires_grad=True)
    f = u @ (o.kernel_dense(kind, x, x, lsr, varr) @ v)
    gl, gv = torch.autograd.grad(f, [lsr, varr])
    ref = torch.cat([gl.reshape(-1), gv.reshape(1)])
    assert float((out.cpu() - ref).abs().max() / ref.abs().max()) <= 1e-9
    parts = eng.zeros(d + 1)
    for p in range(3):
        eng.kmv_bwd_sym(kind, xp, n, d, u.to(dev), v.to(dev), 1.3, ls.to(dev), parts, part=p, nparts=3)
    assert float((parts.cpu() - ref).abs().max() / ref.abs().max()) <= 1e-9
    monkeypatch.setenv("CGLB_DSWEEP", "0")
    out0 = eng.zeros(d + 1)
    eng.kmv_bwd_sym(kind, xp, n, d, u.to(dev), v.to(dev), 1.3, ls.to(dev), out0)
    assert float((out - out0).abs().max() / out0.abs().max()) <= 1e-11


def test_operator_protocol_with_autograd(eng):
    """`kernel(x).add_diag(s2) @ v` (models.py:251-252,280) is differentiable w.r.t. the kernel parameters."""
    x, y, z = o.synthetic_problem(400, 3, 8, seed=2)
    model = make_model("matern32", x.numpy(), y.numpy(), z.numpy(), 0.2, 1.4, [0.7, 1.0, 1.2])
    kern = model.covar_module.base_kernel
    xd = model.train_inputs[0]
    v = torch.randn(400, 1, dtype=f64, device=xd.device, generator=torch.Generator(device=xd.device).manual_seed(0))
    cov = kern(xd).add_diag(model.likelihood.noise.squeeze())
    out = cov @ v
    g = torch.autograd.grad((out * out).sum(), [kern.raw_outputscale, kern.base_kernel.raw_lengthscale, model.likelihood.noise_covar.raw_noise])
    p = o.OracleParams.from_values(0.2, 0.0, z, 1.4, [0.7, 1.0, 1.2])
    K = o.kernel_dense("matern32", x, x, p.lengthscale, p.variance) + p.noise * torch.eye(400, dtype=f64)
    ref_out = K @ v.cpu()
    gref = torch.autograd.grad((ref_out * ref_out).sum(), [p.raw_outputscale, p.raw_lengthscale, p.raw_noise])
    assert rel_max(out.detach().cpu().numpy(), ref_out.detach().numpy()) <= MATVEC_TOL
    for a, b in zip(g, gref):
        assert rel_max(a.cpu().numpy(), b.numpy()) <= 1e-8
    assert rel_max((cov.detach() @ v).cpu().numpy(), ref_out.detach().numpy()) <= MATVEC_TOL
    dense = kern(model.covar_module.inducing_points.detach(), xd).evaluate()
    kref = o.kernel_dense("matern32", z, x, p.lengthscale.detach(), p.variance.detach())
    assert rel_max(dense.cpu().numpy(), kref.numpy()) <= 1e-12


# ---- K1 against a block of right-hand sides (x: [N, t], conjugate_gradient.py:57,66,72) --------------------------------
@pytest.mark.parametrize("kind,n,d,t", [("matern32", 700, 1, 2), ("rbf", 1300, 3, 3), ("matern32", 2500, 11, 4), ("rbf", 999, 8, 5),
                                        ("matern32", 1025, 20, 7), ("rbf", 640, 32, 2), ("matern32", 300, 11, 9), ("matern32", 1, 2, 3)])
def test_multi_rhs_matvec_matches_oracle(eng, kind, n, d, t):
    x, _, _, ls = _problem(n, d, seed=100 * n + t)
    V = torch.randn(n, t, dtype=f64, generator=torch.Generator().manual_seed(t))
    dev = eng.device
    xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    Y = eng.kmv_sym_multi(kind, xp, n, d, V.to(dev), 1.3, 0.07)
    ref = o.kernel_dense(kind, x, x, ls, torch.tensor(1.3, dtype=f64)) @ V + 0.07 * V
    assert Y.shape == (n, t)
    for j in range(t):
        assert float((Y[:, j].cpu() - ref[:, j]).norm() / ref[:, j].norm()) <= MATVEC_TOL, j
    # column j of the block product = the single-RHS sweep on column j (same arithmetic per pair, other summation order)
    y0 = eng.kmv_sym(kind, xp, n, d, V[:, 0].contiguous().to(dev), 1.3, 0.07)
    assert float((Y[:, 0] - y0).norm() / y0.norm()) <= 1e-13
    # work partition (row-sharded ranks): the parts add up to the whole, the diagonal term comes from part 0
    acc = torch.zeros_like(Y)
    for p in range(3):
        acc += eng.kmv_sym_multi(kind, xp, n, d, V.to(dev), 1.3, 0.07, part=p, nparts=3)
    assert float((acc - Y).norm() / Y.norm()) <= 1e-13
    # fixed summation order
    assert torch.equal(Y, eng.kmv_sym_multi(kind, xp, n, d, V.to(dev), 1.3, 0.07))


def test_operator_protocol_accepts_a_block_of_right_hand_sides(eng):
    """`kernel(x).add_diag(s2) @ V` and the bound's own operator for V: [N, t]."""
    from cglb_b200.bound import BoundEvaluator
    x, y, z = o.synthetic_problem(900, 4, 8, seed=3)
    model = make_model("matern32", x.numpy(), y.numpy(), z.numpy(), 0.3, 1.2, [0.7, 1.0, 1.2, 0.9])
    kern = model.covar_module.base_kernel
    xd = model.train_inputs[0]
    V = torch.randn(900, 3, dtype=f64, device=xd.device, generator=torch.Generator(device=xd.device).manual_seed(0))
    p = o.OracleParams.from_values(0.3, 0.0, z, 1.2, [0.7, 1.0, 1.2, 0.9])
    K = (o.kernel_dense("matern32", x, x, p.lengthscale, p.variance) + p.noise * torch.eye(900, dtype=f64)).detach()
    ref = K @ V.cpu()
    with torch.no_grad():
        out = kern(xd).add_diag(model.likelihood.noise.squeeze()) @ V
    assert out.shape == (900, 3) and rel_max(out.cpu().numpy(), ref.numpy()) <= MATVEC_TOL
    out_d = kern(xd).add_diag(model.likelihood.noise.squeeze()).detach() @ V
    assert rel_max(out_d.cpu().numpy(), ref.numpy()) <= MATVEC_TOL
    # with a tape: column by column through the differentiable node, same values
    out_t = kern(xd).add_diag(model.likelihood.noise.squeeze()) @ V
    assert out_t.requires_grad and rel_max(out_t.detach().cpu().numpy(), ref.numpy()) <= MATVEC_TOL
    ev = BoundEvaluator(xd, model.train_targets)
    ls = torch.tensor([0.7, 1.0, 1.2, 0.9], dtype=f64, device=xd.device)
    ev.pack("matern32", ls)
    op = ev.operator("matern32", 1.2, float(p.noise))
    assert rel_max((op @ V).cpu().numpy(), ref.numpy()) <= MATVEC_TOL


# ---- solver API --------------------------------------------------------------------------------------------------------
def test_conjugate_gradient_and_preconditioner_on_reference_golden_system():
    g = np.load(os.path.join(GOLDEN_DIR, "cg_dense_system.npz"))
    dev = torch.device("cuda")
    K, A, LB, b = (torch.from_numpy(g[k]).to(dev) for k in ("K", "A", "LB", "b"))
    pre = cb.NystromPreconditioner(A, LB, torch.tensor(float(g["sigma_sq"]), dtype=f64, device=dev))
    z, rz = pre(b)
    assert rel_max(z.cpu().numpy(), g["precond_z"]) < 1e-11
    assert abs(float(rz) - float(g["precond_rz"])) < 1e-11 * abs(float(g["precond_rz"]))
    for tag, kw in [("default", {}), ("tight", dict(max_error=1e-6)), ("restart", dict(max_error=1e-9, restart_cg_iter=5, max_cg_iter=23))]:
        v, st = cb.ConjugateGradient(**kw)(K, b, torch.zeros_like(b), pre)
        assert abs(int(st.steps) - int(g[f"steps_{tag}"])) <= 1
        assert rel_max(v.cpu().numpy(), g[f"v_{tag}"]) < 1e-4
        if tag == "default" and int(st.steps) == int(g[f"steps_{tag}"]):
            # (near machine-precision convergence, the "tight" case, the final residual is rounding noise)
            assert abs(float(st.residual_error) - float(g[f"err_{tag}"])) <= 1e-3 * abs(float(g[f"err_{tag}"]))


# ---- objective: golden vectors of the reference's own LowerBoundCG -------------------------------------------------------
def _golden_trajectory(name, reuse, grad_tol):
    g = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    kind = str(g["kind"])
    model = make_model(kind, g["x"], g["y"], g["z"], float(g["noise"]), float(g["variance"]), g["lengthscale"], float(g["mean_c"]))
    cg = cb.ConjugateGradient(max_error=float(g["cg_max_error"]), max_cg_iter=int(g["cg_max_iter"]), restart_cg_iter=int(g["cg_restart"]))
    lb = cb.LowerBoundCG(model, cg_opt=cg)
    data = (model.train_inputs[0], model.train_targets)
    lb.evaluator(data).reuse_cg_state = reuse
    params = list(model.parameters())
    for e, mult in enumerate(g["ls_mults"]):
        model.covar_module.base_kernel.base_kernel.lengthscale = torch.as_tensor(g["lengthscale"] * mult)
        loss = -lb(data)
        grads = torch.autograd.grad(loss, params)                                  # optimizer.py:95-98
        ref = float(g[f"loss_{e}"])
        assert abs(float(loss) - ref) <= BOUND_TOL * abs(ref)
        k, kg = int(model.cg_stats.steps), int(g[f"cg_steps_{e}"])
        assert abs(k - kg) <= 1
        assert lb.last_output.matvecs == k + (1 if reuse else 2) + k // cg.restart_cg_iter
        if k == kg:
            assert rel_max(model.v_vec.cpu().numpy(), g[f"v_{e}"]) < 1e-4
            for nm, gr in zip(GRAD_NAMES, grads):
                refg = g[f"grad_{nm}_{e}"]
                assert np.abs(gr.cpu().numpy() - refg).max() <= grad_tol * np.abs(refg).max() + 1e-9, (name, e, nm)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_bound_and_gradients_match_reference_golden(name):
    """The product's default route (K v, r, P r taken from the final CG state) against the golden vectors of the
    reference's own LowerBoundCG along its CG trajectories: bound and every gradient <= 1e-7 (north_star), CG iterations
    +-1.  Every reduction on this path has a fixed summation order, so the numbers are the same on every run; measured
    worst case 7.2e-8 (d/d mean constant of kin_like_rbf, where sum(v + z) cancels to 1/2000 of its terms; the CPU oracle
    itself is 4.8e-8 from the same golden value, another summation order 5.6e-8: profiles/grad_spread_r02.md)."""
    _golden_trajectory(name, reuse=True, grad_tol=1e-7)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_recomputed_residual_route_on_the_golden_trajectories(name):
    """The reference's own route (CGLB_RECOMPUTE_RESIDUAL=1 / evaluator.reuse_cg_state = False): K v, r and P r recomputed
    after the solve as models.py:280-282 does, one more n^2 sweep; same tolerance."""
    _golden_trajectory(name, reuse=False, grad_tol=1e-7)


def test_repeated_evaluations_are_bitwise_identical(eng):
    """Fixed summation order (SURVEY.md section 7 risk 5): per-CTA copies + ordered second stage in the sweeps, ordered
    partials in the preconditioner GEMVs and the split-K GEMMs.  Two runs of the same CG trajectory give the same bits."""
    for case in ("kin_like_rbf", "song_like_wide"):          # d = 8 (register / DMMA kernels) and d > 32 (wide kernels, GEMM epilogues)
        g = np.load(os.path.join(GOLDEN_DIR, f"{case}.npz"))
        outs = []
        for rep in range(2):
            model = make_model(str(g["kind"]), g["x"], g["y"], g["z"], float(g["noise"]), float(g["variance"]), g["lengthscale"], float(g["mean_c"]))
            lb = cb.LowerBoundCG(model)
            loss = -lb((model.train_inputs[0], model.train_targets))
            grads = torch.autograd.grad(loss, list(model.parameters()))
            outs.append((float(loss), int(model.cg_stats.steps), model.v_vec.detach().cpu().clone(), [g_.cpu().clone() for g_ in grads]))
        assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1], case
        assert torch.equal(outs[0][2], outs[1][2]), case
        for ga, gb in zip(outs[0][3], outs[1][3]):          # the K_nm backward and the split-K GEMMs are ordered too
            assert torch.equal(ga, gb), case
    # the sweeps themselves, on the three kernels (register-resident, DMMA distance, wide)
    for n, d, mode in ((3000, 3, "0"), (3000, 11, "2"), (1500, 40, "1")):
        os.environ["CGLB_DSWEEP"] = mode
        try:
            x, v, u, ls = _problem(n, d, seed=5)
            dev = eng.device
            xp = eng.pack("matern32", x.to(dev), ls.to(dev), x.mean(0).to(dev))
            ys = [eng.kmv_sym("matern32", xp, n, d, v.to(dev), 1.3, 0.1).clone() for _ in range(3)]
            assert torch.equal(ys[0], ys[1]) and torch.equal(ys[0], ys[2]), (n, d)
            gs = []
            for _ in range(2):
                out = eng.zeros(d + 1)
                eng.kmv_bwd_sym("matern32", xp, n, d, u.to(dev), v.to(dev), 1.3, ls.to(dev), out)
                gs.append(out.clone())
            assert torch.equal(gs[0], gs[1]), (n, d)
        finally:
            os.environ.pop("CGLB_DSWEEP", None)


@pytest.mark.parametrize("name", ["kin_like_rbf_fp32", "house_like_matern_fp32"])
def test_fp32_mode_matches_reference_fp32_golden(name):
    """The reference's own LowerBoundCG run on float32 tensors (oracle/make_golden.py: FP32_CASES, jitter 1e-5)
    against an fp32 model here (FP32 kernel pairs, FP64 accumulation and linear algebra).  The reference's fp32 run
    is itself ~3e-6 away from its fp64 run on the bound and ~1e-4 on the gradients: tolerances 2e-5 / 1e-3."""
    from cglb_b200 import settings
    g = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    kind = str(g["kind"])
    old = settings.cholesky_jitter.value()
    settings.cholesky_jitter._set_value(float(g["jitter"]))
    try:
        model = make_model(kind, g["x"], g["y"], g["z"], float(g["noise"]), float(g["variance"]), g["lengthscale"],
                           float(g["mean_c"]), dtype=torch.float32)
        assert model.train_inputs[0].dtype == torch.float32
        loss = -cb.LowerBoundCG(model)((model.train_inputs[0], model.train_targets))
        grads = torch.autograd.grad(loss, list(model.parameters()))
        ref = float(g["loss_0"])
        assert loss.dtype == torch.float32 and abs(float(loss) - ref) <= 2e-5 * abs(ref)
        assert abs(int(model.cg_stats.steps) - int(g["cg_steps_0"])) <= 1
        for nm, gr in zip(GRAD_NAMES, grads):
            refg = g[f"grad_{nm}_0"]
            assert np.abs(gr.cpu().numpy() - refg).max() <= 1e-3 * np.abs(refg).max() + 1e-6, (name, nm)
    finally:
        settings.cholesky_jitter._set_value(old)


@pytest.mark.parametrize("kind,n,d,M,noise", [("matern32", 900, 3, 40, 0.05), ("rbf", 700, 8, 33, 0.3), ("matern32", 513, 11, 64, 0.01),
                                              ("matern32", 400, 20, 24, 0.1), ("matern32", 1200, 90, 48, 0.05), ("rbf", 700, 40, 17, 0.2)])
def test_bound_and_gradients_fixed_v_match_oracle(kind, n, d, M, noise):
    """With CG disabled (use_cache, as the reference's metrics path interface.py:621-625) every term is a
    deterministic function of v: compare at tight tolerance."""
    x, y, z = o.synthetic_problem(n, d, M, seed=n)
    ls = np.linspace(0.8, 1.6, d) * 0.5 * math.sqrt(d)
    model = make_model(kind, x.numpy(), y.numpy(), z.numpy(), noise, 1.2, ls, 0.05)
    v = 0.1 * torch.randn(n, 1, dtype=f64, generator=torch.Generator().manual_seed(1))
    model.v_vec.data.copy_(v.cuda())
    lb = cb.LowerBoundCG(model, use_cache=True, cached_v_vec_initial=True)
    loss = -lb((model.train_inputs[0], model.train_targets))
    grads = torch.autograd.grad(loss, list(model.parameters()))
    p = o.OracleParams.from_values(noise, 0.05, z, 1.2, ls)
    res = o.lower_bound(kind, p, x, y, v, use_cached_v=True)
    ref_grads = torch.autograd.grad(-res.bound, p.tensors())
    assert abs(float(loss) + float(res.bound)) <= 1e-10 * abs(float(res.bound))
    for nm, a, b in zip(GRAD_NAMES, grads, ref_grads):
        assert np.abs(a.cpu().numpy() - b.numpy()).max() <= 1e-8 * np.abs(b.numpy()).max() + 1e-10, nm


def test_warm_start_sequence_matches_oracle_iteration_counts():
    """>= 3 consecutive evaluations with changing hyper-parameters (risk 5)."""
    n, d, M = 1200, 3, 48
    x, y, z = o.synthetic_problem(n, d, M, seed=8)
    model = make_model("matern32", x.numpy(), y.numpy(), z.numpy(), 0.02, 1.0, 0.9)
    lb = cb.LowerBoundCG(model)
    for mult in (1.0, 1.01, 0.99, 1.05):
        # the oracle is warm-started from the SAME vector as the device path: CG amplifies rounding
        # differences of earlier solves, which is a property of the algorithm, not of the implementation
        v = model.v_vec.detach().cpu().clone()
        model.covar_module.base_kernel.base_kernel.lengthscale = torch.full((d,), 0.9 * mult, dtype=f64)
        loss = -lb((model.train_inputs[0], model.train_targets))
        p = o.OracleParams.from_values(0.02, 0.0, z, 1.0, 0.9 * mult)
        res = o.lower_bound("matern32", p, x, y, v)
        assert abs(int(model.cg_stats.steps) - res.cg.steps) <= 1
        if int(model.cg_stats.steps) == res.cg.steps:
            assert abs(float(loss) + float(res.bound)) <= BOUND_TOL * abs(float(res.bound))
    assert float(model.v_vec.abs().max()) > 0


@pytest.mark.parametrize("M,noise,max_error,expect_restart", [(24, 0.3, 1.0, False), (64, 0.02, 1e-3, True)])
def test_cg_state_reuse_matches_the_recomputed_route(M, noise, max_error, expect_restart):
    """models.py:280-282 recomputes K v, r and P r after the solve; the device path takes them from the final CG state
    (bound.py: reuse_cg_state, one n^2 sweep less per evaluation).  Both routes at the SAME v: the solve with reuse,
    then the reference's cached-v route (models.py:263-264), which recomputes all three; then the solve again with
    reuse switched off.  k + 1 + floor(k/40) against k + 2 + floor(k/40) sweeps, bound and gradients equal far inside
    the tolerance (the CPU side of the claim: tests/test_cg_residual_reuse.py)."""
    n, d = 1400, 5
    x, y, z = o.synthetic_problem(n, d, M, seed=21)
    model = make_model("matern32", x.numpy(), y.numpy(), z.numpy(), noise, 1.1, 1.3, 0.02)
    data = (model.train_inputs[0], model.train_targets)
    params = list(model.parameters())
    lb = cb.LowerBoundCG(model, cg_opt=cb.ConjugateGradient(max_error=max_error))
    assert lb.evaluator(data).reuse_cg_state
    loss = -lb(data)
    grads = [g_.cpu().numpy() for g_ in torch.autograd.grad(loss, params)]
    k = int(model.cg_stats.steps)
    assert (k >= 40) == expect_restart and k < 100
    assert lb.last_output.matvecs == k + 1 + k // 40
    # the recomputed route at the same v
    lb_c = cb.LowerBoundCG(model, use_cache=True, cached_v_vec_initial=True)
    loss_c = -lb_c(data)
    grads_c = [g_.cpu().numpy() for g_ in torch.autograd.grad(loss_c, params)]
    assert lb_c.last_output.matvecs == 1
    assert abs(float(loss) - float(loss_c)) <= 1e-11 * abs(float(loss_c))
    for nm, a, b in zip(GRAD_NAMES, grads, grads_c):
        assert np.abs(a - b).max() <= 1e-8 * np.abs(b).max() + 1e-10, nm
    # the reference's count with reuse switched off (what CGLB_RECOMPUTE_RESIDUAL=1 selects)
    model.v_vec.data.zero_()
    lb.evaluator(data).reuse_cg_state = False
    loss_r = -lb(data)
    k_r = int(model.cg_stats.steps)
    assert abs(k_r - k) <= 1 and lb.last_output.matvecs == k_r + 2 + k_r // 40
    assert abs(float(loss) - float(loss_r)) <= BOUND_TOL * abs(float(loss_r))


@pytest.mark.parametrize("name", ["road_like_trained", "house_like_warmstart", "snelson_like_init", "kin_like_m256", "house_like_m256"])
def test_predict_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    last = len(g["ls_mults"]) - 1
    model = make_model(str(g["kind"]), g["x"], g["y"], g["z"], float(g["noise"]), float(g["variance"]),
                       g["lengthscale"] * g["ls_mults"][last], float(g["mean_c"]))
    model.v_vec.data.copy_(torch.from_numpy(g[f"v_{last}"]).cuda())
    pred = cb.PredictCG(model)
    mean, var = pred(torch.from_numpy(g["xnew"]).cuda())
    assert rel_max(mean.cpu().numpy(), g["f_mean"]) < 1e-4        # v is only converged to max_error = 1e-3
    assert rel_max(var.cpu().numpy(), g["f_var"]) < 1e-9
    mean2, var2 = pred(torch.from_numpy(g["xnew"]).cuda())       # cached path (models.py:323-325)
    assert torch.allclose(mean, mean2, rtol=1e-12, atol=1e-13) and torch.allclose(var, var2, rtol=1e-12, atol=1e-13)
    with pytest.raises(NotImplementedError):
        pred(torch.from_numpy(g["xnew"]).cuda(), full_cov=True)


# ---- error behaviour -----------------------------------------------------------------------------------------
def test_error_behaviour(eng):
    with pytest.raises(ValueError):
        cb.LowerBoundCG(torch.nn.Linear(1, 1))
    bad = -torch.eye(130, dtype=f64, device=eng.device)
    with pytest.raises(RuntimeError):
        eng.potrf(bad)
    with pytest.raises(cb.CglbError):
        eng.pack("matern32", torch.zeros(4, 2, dtype=f64), torch.ones(2, dtype=f64, device=eng.device), None)
    with pytest.raises(cb.CglbError):                                # d beyond what the shared-memory tiles hold
        eng.kmv_sym("rbf", eng.zeros(128, eng.packed_width(300)), 10, 300, eng.zeros(10), 1.0, 0.0)


# ---- dense kernels ------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("m", [17, 128, 200, 1024])
def test_dense_factorisations(eng, m):
    g = torch.Generator().manual_seed(m)
    xx = torch.randn(m, m + 20, generator=g, dtype=f64).to(eng.device)
    spd = xx @ xx.t() / m + torch.eye(m, dtype=f64, device=eng.device)
    l = eng.potrf(spd.clone())
    lref = torch.linalg.cholesky(spd)
    assert rel_max(l.cpu().numpy(), lref.cpu().numpy()) < 1e-12
    linv = eng.tri_inverse(l)
    assert rel_max((linv @ l).cpu().numpy(), np.eye(m)) < 1e-11
    n = 333
    b = torch.randn(m, n + 1, generator=g, dtype=f64).to(eng.device)
    ref = 0.7 * torch.linalg.solve_triangular(lref, b[:, :n], upper=False)
    eng.trsm_left_lower(l, b, n, alpha=0.7)
    assert rel_max(b[:, :n].cpu().numpy(), ref.cpu().numpy()) < 1e-11
    c = eng.empty(m, m)
    eng.syrk(b, m, n, c)
    assert rel_max(c.cpu().numpy(), (b[:, :n] @ b[:, :n].t()).cpu().numpy()) < 1e-12


@pytest.mark.parametrize("m,n,k,transb", [(128, 128, 16, False), (130, 257, 100, False), (300, 200, 333, True), (64, 1000, 2048, True),
                                          (2048, 160, 4000, False), (1, 1, 5, True), (257, 129, 1, False)])
def test_gemm_both_staging_paths(eng, m, n, k, transb):
    """cglb_gemm with both operand-staging kernels -- the cp.async ring (default) and the TMA bulk-copy ring
    (cglb_set_option "gemm_staging" 2; 16-byte aligned operand rows only, odd leading dimensions fall back) -- on ragged edge
    tiles, K tails, the split-K route for few tiles and a long K, beta != 0."""
    g = torch.Generator().manual_seed(m * 7 + n)
    dev = eng.device
    try:
        for staging in (1, 2):
            eng.set_option("gemm_staging", staging)
            for pad in (0, 1):                 # pad = 1: odd leading dimensions -> rows are not 16-byte aligned
                lda = k + (k & 1) + pad
                a_buf = torch.randn(m, lda, generator=g, dtype=f64).to(dev)
                a = a_buf[:, :k]
                if transb:
                    ldb = k + (k & 1) + pad
                    b_buf = torch.randn(n, ldb, generator=g, dtype=f64).to(dev)
                    ref = a @ b_buf[:, :k].t()
                else:
                    ldb = n + (n & 1) + pad
                    b_buf = torch.randn(k, ldb, generator=g, dtype=f64).to(dev)
                    ref = a @ b_buf[:, :n]
                ldc = n + (n & 1)
                c_buf = torch.randn(m, ldc, generator=g, dtype=f64).to(dev)
                c0 = c_buf[:, :n].clone()
                rc = eng.lib.cglb_gemm(eng.ctx, int(transb), m, n, k, 0.7, a_buf.data_ptr(), lda, b_buf.data_ptr(), ldb, -0.3,
                                       c_buf.data_ptr(), ldc, eng.stream())
                assert rc == 0
                want = 0.7 * ref - 0.3 * c0
                err = float((c_buf[:, :n] - want).abs().max() / (want.abs().max() + 1e-300))
                assert err <= 1e-12, (staging, pad, err)
    finally:
        eng.set_option("gemm_staging", 1)


def test_dense_factorisations_with_tma_staging(eng):
    """potrf / tri_inverse / trsm / syrk through the TMA-staged GEMM (in-place panel and block-row products included)."""
    try:
        eng.set_option("gemm_staging", 2)
        test_dense_factorisations(eng, 200)
        test_dense_factorisations(eng, 1024)
    finally:
        eng.set_option("gemm_staging", 1)


@pytest.mark.parametrize("q", [1, 2, 3])
def test_dmma_sweeps_with_several_super_rows(eng, monkeypatch, q):
    """The L2-blocked item order of the DMMA sweeps (DCursor::decode): with `superrow` = q chunks a problem of 5000 rows spans
    several super-rows (a chunk is 1024 rows); forward and backward sweeps against the oracle, and the 3-way work partition."""
    monkeypatch.setenv("CGLB_DSWEEP", "2")
    n, d, kind = 5000, 11, "matern32"
    x, v, u, ls = _problem(n, d, seed=77)
    dev = eng.device
    xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    try:
        eng.set_option("superrow", q)
        y = eng.kmv_sym(kind, xp, n, d, v.to(dev), 1.3, 0.07)
        ref = o.kernel_dense(kind, x, x, ls, torch.tensor(1.3, dtype=f64), block=1000) @ v + 0.07 * v
        assert float((y.cpu() - ref).norm() / ref.norm()) <= MATVEC_TOL
        parts = sum(eng.kmv_sym(kind, xp, n, d, v.to(dev), 1.3, 0.07, part=p, nparts=3) for p in range(3))
        assert float((parts - y).norm() / y.norm()) <= 1e-13
        out = eng.zeros(d + 1)
        eng.kmv_bwd_sym(kind, xp, n, d, u.to(dev), v.to(dev), 1.3, ls.to(dev), out)
        eng.set_option("superrow", 0)
        out0 = eng.zeros(d + 1)
        eng.kmv_bwd_sym(kind, xp, n, d, u.to(dev), v.to(dev), 1.3, ls.to(dev), out0)
        assert float((out - out0).abs().max() / out0.abs().max()) <= 1e-11
        y0 = eng.kmv_sym(kind, xp, n, d, v.to(dev), 1.3, 0.07)
        assert float((y - y0).norm() / y0.norm()) <= 1e-13
    finally:
        eng.set_option("superrow", 0)
