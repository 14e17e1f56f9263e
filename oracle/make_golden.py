"""Generate tests/golden/*.npz by running the REFERENCE'S OWN code (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):   python -m oracle.make_golden

For every case the verbatim ``LowerBoundCG`` / ``ConjugateGradient`` / ``NystromPreconditioner`` /
``PredictCG`` from /root/reference/cglb/backend/pytorch/{models,conjugate_gradient}.py are executed on
CPU in fp64 (third-party gpytorch symbols supplied by oracle/gpytorch_stub.py), the loss is
differentiated exactly as pytorch/optimizer.py:95-98 does, and inputs + outputs are stored.  A case
with several ``evals`` re-evaluates the bound after perturbing the lengthscales, which exercises the
warm start ``model.v_vec`` (models.py:274).
"""
from __future__ import annotations

import os
import warnings

import numpy as np
import torch

from . import cglb_oracle as o
from . import reference_loader as rl

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name, kind, n, d, M, noise, variance, lengthscale(s), mean_c, ls multipliers for consecutive evals, seed
CASES = [
    ("snelson_like_init", "matern32", 300, 1, 24, 1.0, 1.0, 1.0, 0.0, [1.0], 10),
    ("road_like_trained", "matern32", 500, 3, 32, 0.05, 1.3, [0.7, 1.1, 0.9], 0.1, [1.0], 11),
    ("kin_like_rbf", "rbf", 400, 8, 48, 0.02, 0.8, 2.0, -0.05, [1.0], 12),
    ("house_like_warmstart", "matern32", 600, 11, 64, 0.01, 1.0, 1.66, 0.0, [1.0, 1.01, 0.99], 13),
    ("ragged_rbf_init", "rbf", 257, 2, 17, 1.0, 1.0, 1.0, 0.0, [1.0, 1.01], 14),
    ("wide_d_matern", "matern32", 300, 20, 32, 0.05, 1.0, 3.0, 0.0, [1.0], 15),
    ("song_like_wide", "matern32", 350, 90, 40, 0.05, 1.0, 4.7, 0.0, [1.0, 1.01], 17),
    # M = 256 > the 128-wide panels of the blocked factorisations (round 2): the reference's own code pins the blocked
    # potrf / TRSM / SYRK path through the bound, with a warm-started second evaluation
    ("kin_like_m256", "rbf", 2000, 8, 256, 0.05, 1.0, 2.0, 0.0, [1.0, 1.01], 18),
    ("house_like_m256", "matern32", 1500, 11, 256, 0.02, 1.0, 1.66, 0.0, [1.0, 1.01], 19),
    # restart branch of conjugate_gradient.py:70-75 exercised through LowerBoundCG(model, cg_opt=...)
    ("restart_path", "matern32", 400, 2, 6, 0.02, 1.5, 0.5, 0.0, [1.0], 16,
     dict(max_error=1e-3, restart_cg_iter=4, max_cg_iter=100)),
]


def run_case(cg, models, name, kind, n, d, M, noise, variance, ls, mean_c, mults, seed, cg_kw=None, n_new=64):
    x, y, z = o.synthetic_problem(n, d, M, seed=seed)
    ls_vec = np.broadcast_to(np.asarray(ls, dtype=np.float64).reshape(-1), (d,)).copy()
    model = rl.build_reference_model(models, kind, x, y, z, noise, variance, ls_vec, mean_c)
    params = list(model.parameters())
    cg_kw = cg_kw or {}
    lower_bound = models.LowerBoundCG(model, cg_opt=cg.ConjugateGradient(**cg_kw) if cg_kw else None)
    out = dict(kind=kind, x=x.numpy(), y=y.numpy(), z=z.numpy(), noise=noise, variance=variance,
               lengthscale=ls_vec, mean_c=mean_c, ls_mults=np.asarray(mults), jitter=1e-6,
               cg_max_error=cg_kw.get("max_error", 1.0), cg_restart=cg_kw.get("restart_cg_iter", 40),
               cg_max_iter=cg_kw.get("max_cg_iter", 100))
    base_raw = model.covar_module.base_kernel.base_kernel.raw_lengthscale.data.clone()
    for e, mult in enumerate(mults):
        kern = model.covar_module.base_kernel.base_kernel
        kern.lengthscale = torch.as_tensor(ls_vec * mult)
        loss = -lower_bound((x, y))
        grads = torch.autograd.grad(loss, params)
        out[f"loss_{e}"] = loss.detach().numpy()
        for gname, g in zip(["raw_noise", "mean_constant", "inducing_points", "raw_outputscale", "raw_lengthscale"], grads):
            out[f"grad_{gname}_{e}"] = g.detach().numpy()
        out[f"cg_steps_{e}"] = int(model.cg_stats.steps)
        out[f"cg_error_{e}"] = float(model.cg_stats.residual_error)
        out[f"v_{e}"] = model.v_vec.detach().numpy().copy()
    # prediction with the reference's PredictCG (models.py:289-354) at the last hyper-parameters
    g = torch.Generator().manual_seed(seed + 100)
    xnew = torch.randn(n_new, d, generator=g, dtype=torch.float64)
    with torch.no_grad():
        pred = models.PredictCG(model)
        f_mean, f_var = pred(xnew)
    out["xnew"] = xnew.numpy()
    out["f_mean"] = f_mean.numpy()
    out["f_var"] = f_var.numpy()
    out["predict_v"] = pred.v_vec.detach().numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), **out)
    print(f"{name}: loss={[float(out[f'loss_{e}']) for e in range(len(mults))]} "
          f"cg_steps={[out[f'cg_steps_{e}'] for e in range(len(mults))]}")


# fp32 mode of the reference (interface.py:94-104, jitter 1e-5 per backend.py:76-79): the same verbatim code on
# float32 tensors.  name, kind, n, d, M, noise, variance, lengthscale, mean_c, seed
FP32_CASES = [
    ("kin_like_rbf_fp32", "rbf", 400, 8, 48, 0.05, 0.9, 2.0, 0.0, 21),
    ("house_like_matern_fp32", "matern32", 600, 11, 64, 0.05, 1.0, 1.66, 0.1, 22),
]


def run_case_fp32(models, name, kind, n, d, M, noise, variance, ls, mean_c, seed):
    from . import gpytorch_stub as gp
    x, y, z = o.synthetic_problem(n, d, M, seed=seed)
    x32, y32, z32 = x.float(), y.float(), z.float()
    lik = gp.GaussianLikelihood(noise_constraint=gp.GreaterThan(1e-6)).float()
    lik.noise = noise
    base = (gp.MaternKernel(nu=1.5, ard_num_dims=d) if kind == "matern32" else gp.RBFKernel(ard_num_dims=d)).float()
    base.lengthscale = torch.full((d,), float(ls), dtype=torch.float32)
    scale = gp.ScaleKernel(base).float()
    scale.outputscale = variance
    model = models.CGLB((x32, y32.reshape(-1)), lik, gp.InducingPointKernel(scale, z32, likelihood=lik))
    model.mean_module = model.mean_module.float()
    model.mean_module.constant.data.fill_(mean_c)
    params = list(model.parameters())
    loss = -models.LowerBoundCG(model)((x32, y32))
    grads = torch.autograd.grad(loss, params)
    out = dict(kind=kind, x=x32.numpy(), y=y32.numpy(), z=z32.numpy(), noise=noise, variance=variance,
               lengthscale=np.full(d, ls, dtype=np.float32), mean_c=mean_c, jitter=1e-5,
               loss_0=loss.detach().numpy(), cg_steps_0=int(model.cg_stats.steps))
    for gname, g in zip(["raw_noise", "mean_constant", "inducing_points", "raw_outputscale", "raw_lengthscale"], grads):
        out[f"grad_{gname}_0"] = g.detach().numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, f"{name}.npz"), **out)
    print(f"{name}: loss={float(loss)} cg_steps={int(model.cg_stats.steps)} dtype={loss.dtype}")


def run_cg_case(cg):
    """Verbatim ConjugateGradient + NystromPreconditioner on a dense SPD system with explicit A, LB."""
    g = torch.Generator().manual_seed(77)
    n, m = 200, 16
    x = torch.randn(n, 2, generator=g, dtype=torch.float64)
    z = x[:m].clone()
    ls, var, s2 = torch.tensor([[0.8, 1.2]], dtype=torch.float64), torch.tensor(1.1, dtype=torch.float64), torch.tensor(0.03, dtype=torch.float64)
    K = o.kernel_dense("matern32", x, x, ls, var) + s2 * torch.eye(n, dtype=torch.float64)
    kuf = o.kernel_dense("matern32", z, x, ls, var)
    kuu = o.kernel_dense("matern32", z, z, ls, var) + 1e-6 * torch.eye(m, dtype=torch.float64)
    L = torch.linalg.cholesky(kuu)
    A = torch.linalg.solve_triangular(L, kuf, upper=False) / torch.sqrt(s2)
    LB = torch.linalg.cholesky(A @ A.T + torch.eye(m, dtype=torch.float64))
    b = torch.randn(n, 1, generator=g, dtype=torch.float64)
    precond = cg.NystromPreconditioner(A, LB, s2)
    z0, rz0 = precond(b)
    out = dict(K=K.numpy(), A=A.numpy(), LB=LB.numpy(), sigma_sq=float(s2), b=b.numpy(),
               precond_z=z0.numpy(), precond_rz=float(rz0))
    for tag, kw in [("default", {}), ("tight", dict(max_error=1e-6)), ("restart", dict(max_error=1e-9, restart_cg_iter=5, max_cg_iter=23))]:
        v, stats = cg.ConjugateGradient(**kw)(K, b, torch.zeros(n, 1, dtype=torch.float64), precond)
        out[f"v_{tag}"] = v.numpy()
        out[f"steps_{tag}"] = int(stats.steps)
        out[f"err_{tag}"] = float(stats.residual_error)
        print("cg", tag, stats.steps, float(stats.residual_error))
    np.savez_compressed(os.path.join(GOLDEN_DIR, "cg_dense_system.npz"), **out)


def main():
    warnings.filterwarnings("ignore")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    cg, models, _ = rl.load(jitter=1e-6)
    run_cg_case(cg)
    for case in CASES:
        run_case(cg, models, *case)
    rl.load(jitter=1e-5)             # the reference's fp32 jitter
    for case in FP32_CASES:
        run_case_fp32(models, *case)
    rl.load(jitter=1e-6)


if __name__ == "__main__":
    main()
