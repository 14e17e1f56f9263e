"""CPU: the restated oracle (oracle/cglb_oracle.py) against the golden vectors produced by the
reference's own LowerBoundCG / ConjugateGradient / NystromPreconditioner / PredictCG
(oracle/make_golden.py), plus the mathematical identities of SURVEY.md section 4."""
import math
import os

import numpy as np
import pytest
import torch

from oracle import cglb_oracle as o
from conftest import GOLDEN_CASES, GOLDEN_DIR, GOLDEN_FP32_CASES, GRAD_NAMES

f64 = torch.float64


def _rel(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def test_cg_and_preconditioner_match_reference_golden():
    g = np.load(os.path.join(GOLDEN_DIR, "cg_dense_system.npz"))
    K, A, LB, b = (torch.from_numpy(g[k]) for k in ("K", "A", "LB", "b"))
    s2 = torch.tensor(float(g["sigma_sq"]), dtype=f64)
    pre = o.nystrom_preconditioner(A, LB, s2)
    z, rz = pre(b)
    assert _rel(z.numpy(), g["precond_z"]) < 1e-12
    assert abs(float(rz) - float(g["precond_rz"])) < 1e-12 * abs(float(g["precond_rz"]))
    for tag, kw in [("default", {}), ("tight", dict(max_error=1e-6)),
                    ("restart", dict(max_error=1e-9, restart_cg_iter=5, max_cg_iter=23))]:
        v, st = o.conjugate_gradient(K, b, torch.zeros_like(b), pre, **kw)
        assert abs(st.steps - int(g[f"steps_{tag}"])) <= 1          # north_star: CG iteration count within +-1
        assert _rel(v.numpy(), g[f"v_{tag}"]) < 1e-4   # CG amplifies rounding differences; v is only converged to max_error
        if st.steps == int(g[f"steps_{tag}"]):
            assert abs(st.residual_error - float(g[f"err_{tag}"])) <= 1e-3 * abs(float(g[f"err_{tag}"]))


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_bound_and_grads_match_reference_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    kind = str(g["kind"])
    x, y, z = (torch.from_numpy(g[k]) for k in ("x", "y", "z"))
    v = torch.zeros(x.shape[0], 1, dtype=f64)
    for e, mult in enumerate(g["ls_mults"]):
        p = o.OracleParams.from_values(float(g["noise"]), float(g["mean_c"]), z, float(g["variance"]),
                                       g["lengthscale"] * mult)
        loss, grads, res = o.bound_and_grads(kind, p, x, y, v, jitter=float(g["jitter"]),
                                             max_error=float(g["cg_max_error"]), max_cg_iter=int(g["cg_max_iter"]),
                                             restart_cg_iter=int(g["cg_restart"]))
        v = res.v
        assert abs(float(loss) - float(g[f"loss_{e}"])) <= 1e-9 * abs(float(g[f"loss_{e}"]))
        assert res.cg.steps == int(g[f"cg_steps_{e}"])
        assert _rel(res.v.numpy(), g[f"v_{e}"]) < 1e-4
        for gname, gr in zip(GRAD_NAMES, grads):
            ref = g[f"grad_{gname}_{e}"]
            assert np.abs(gr.numpy() - ref).max() <= 1e-7 * np.abs(ref).max() + 1e-9, gname


@pytest.mark.parametrize("name", GOLDEN_FP32_CASES)
def test_fp32_golden_is_consistent_with_the_fp64_oracle(name):
    """The fp32 golden vectors (the reference's own code on float32 tensors, jitter 1e-5) against the fp64 oracle on
    the same (float32-rounded) inputs: the reference's fp32 arithmetic is ~3e-6 away on the bound and ~2e-4 on the
    gradients, and takes the same number of CG iterations (+-1)."""
    g = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    assert g["x"].dtype == np.float32 and g["loss_0"].dtype == np.float32
    kind = str(g["kind"])
    x, y, z = (torch.from_numpy(g[k]).double() for k in ("x", "y", "z"))
    p = o.OracleParams.from_values(float(g["noise"]), float(g["mean_c"]), z, float(g["variance"]), g["lengthscale"].astype(np.float64))
    loss, grads, res = o.bound_and_grads(kind, p, x, y, torch.zeros(x.shape[0], 1, dtype=f64), jitter=float(g["jitter"]))
    assert abs(float(loss) - float(g["loss_0"])) <= 2e-5 * abs(float(loss))
    assert abs(res.cg.steps - int(g["cg_steps_0"])) <= 1
    for gname, gr in zip(GRAD_NAMES, grads):
        ref = g[f"grad_{gname}_0"]
        assert np.abs(gr.numpy() - ref).max() <= 1e-3 * np.abs(gr.numpy()).max() + 1e-6, gname


@pytest.mark.parametrize("name", ["road_like_trained", "kin_like_rbf", "kin_like_m256", "house_like_m256"])
def test_predict_matches_reference_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    kind = str(g["kind"])
    x, y, z, xnew = (torch.from_numpy(g[k]) for k in ("x", "y", "z", "xnew"))
    last = len(g["ls_mults"]) - 1
    p = o.OracleParams.from_values(float(g["noise"]), float(g["mean_c"]), z, float(g["variance"]),
                                   g["lengthscale"] * g["ls_mults"][last])
    v0 = torch.from_numpy(g[f"v_{last}"])
    mean, var, _, _ = o.predict(kind, p, x, y, xnew, v0)
    assert _rel(mean.numpy(), g["f_mean"]) < 1e-7
    assert _rel(var.numpy(), g["f_var"]) < 1e-7


# ---- identities -----------------------------------------------------------
def _small_problem(kind="matern32", n=120, d=2, M=15, noise=0.1):
    x, y, z = o.synthetic_problem(n, d, M, seed=5)
    p = o.OracleParams.from_values(noise, 0.0, z, 1.2, 0.9)
    return x, y, z, p


def test_tight_cg_recovers_exact_quadratic_and_bounds_bracket():
    x, y, z, p = _small_problem()
    n = x.shape[0]
    K = o.kernel_dense("matern32", x, x, p.lengthscale, p.variance) + p.noise * torch.eye(n, dtype=f64)
    exact = -0.5 * (y[None, :] @ torch.linalg.solve(K, y[:, None])).item()
    v0 = torch.zeros(n, 1, dtype=f64)
    loose = o.lower_bound("matern32", p, x, y, v0, max_error=1.0)
    tight = o.lower_bound("matern32", p, x, y, v0, max_error=1e-12, max_cg_iter=500)
    # Bounds(upper_bound=-upper, lower_bound=-lower): -upper <= exact <= -lower
    assert float(loose.upper) <= exact + 1e-9 <= float(loose.lower) + 2e-9
    assert abs(float(tight.upper) - exact) < 1e-8 * abs(exact)
    assert abs(float(tight.lower) - exact) < 1e-8 * abs(exact)


def test_v_zero_gives_sgpr_quadratic_term():
    x, y, z, p = _small_problem()
    n = x.shape[0]
    terms = o.common_terms("matern32", p, x, 1e-6)
    pre = o.nystrom_preconditioner(terms.A, terms.LB, p.noise)
    err = y.reshape(-1, 1)
    _, ePe = pre(err)
    res = o.lower_bound("matern32", p, x, y, torch.zeros(n, 1, dtype=f64), use_cached_v=True)
    assert abs(float(res.upper) + 0.5 * float(ePe)) < 1e-10 * abs(float(ePe))


def test_full_inducing_set_recovers_exact_log_marginal_likelihood():
    x, y, _, _ = _small_problem()
    n = x.shape[0]
    p = o.OracleParams.from_values(0.1, 0.0, x.clone(), 1.2, 0.9)   # Z = X, M = n
    K = o.kernel_dense("matern32", x, x, p.lengthscale, p.variance) + p.noise * torch.eye(n, dtype=f64)
    exact = (-0.5 * (y[None, :] @ torch.linalg.solve(K, y[:, None])) - 0.5 * torch.logdet(K)
             - 0.5 * n * math.log(2 * math.pi)).item()
    res = o.lower_bound("matern32", p, x, y, torch.zeros(n, 1, dtype=f64), max_error=1e-12, max_cg_iter=500)
    assert abs(float(res.bound) - exact) < 1e-4 * abs(exact)   # jitter 1e-6 on Kuu limits the match


def test_preconditioner_inverts_nystrom_plus_noise():
    x, y, z, p = _small_problem()
    terms = o.common_terms("matern32", p, x, 1e-6)
    pre = o.nystrom_preconditioner(terms.A, terms.LB, p.noise)
    q = p.noise * (terms.A.T @ terms.A) + p.noise * torch.eye(x.shape[0], dtype=f64)
    r = torch.randn(x.shape[0], 1, dtype=f64, generator=torch.Generator().manual_seed(3))
    zz, _ = pre(q @ r)
    assert _rel(zz.detach().numpy(), r.numpy()) < 1e-9


def test_expanded_distance_matches_direct_form():
    x, _, z, p = _small_problem(d=2)
    a = o.sqdist_direct(x, x, p.lengthscale.detach())
    b = o.sqdist_expanded(x, x, p.lengthscale.detach(), x1_eq_x2=True)
    assert float((a - b).abs().max()) < 1e-12


def test_blocked_fixed_v_oracle_matches_the_dense_one():
    """oracle.bound_and_grads_fixed_v_blocked (used by the GPU tests at Nystrom sizes M = 1024 / 2048, n = 16k) against the
    dense oracle on a problem small enough for both."""
    n, d, M = 700, 5, 48
    x, y, z = o.synthetic_problem(n, d, M, seed=17)
    v = 0.1 * torch.randn(n, 1, dtype=f64, generator=torch.Generator().manual_seed(3))
    for kind in ("matern32", "rbf"):
        p = o.OracleParams.from_values(0.05, 0.1, z, 1.3, np.linspace(0.8, 1.4, d))
        loss_d, grads_d, _ = o.bound_and_grads(kind, p, x, y, v, use_cached_v=True)
        p2 = o.OracleParams.from_values(0.05, 0.1, z, 1.3, np.linspace(0.8, 1.4, d))
        loss_b, grads_b = o.bound_and_grads_fixed_v_blocked(kind, p2, x, y, v, block=96)
        assert abs(float(loss_d) - float(loss_b)) <= 1e-13 * abs(float(loss_d))
        for a, b in zip(grads_d, grads_b):
            assert _rel(b.numpy(), a.numpy()) <= 1e-11
