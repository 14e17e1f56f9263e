"""CPU: the numerical claim behind `BoundEvaluator.reuse_cg_state` (cglb_b200/bound.py).

The reference recomputes K v, r = err - K v and P r after the CG solve (models.py:280-282) because autograd needs them
on its tape; the device path takes all three from the final state of the CG loop (conjugate_gradient.py:66-75: r is
recomputed at the start and at every restart and updated by r -= gamma K p in between).  Here the reference's loop is
replayed on the CPU oracle with a dense K, and the quadratic-form bounds from both routes are compared on problems from
well to badly conditioned, with and without restarts, up to the 100-iteration cap."""
import math

import pytest
import torch

from oracle import cglb_oracle as o

f64 = torch.float64


def _replay(kind, n, d, M, noise, ls, max_error, max_iter, restart, seed):
    x, y, z = o.synthetic_problem(n, d, M, seed=seed)
    p = o.OracleParams.from_values(noise, 0.0, z, 1.0, ls)
    with torch.no_grad():
        terms = o.common_terms(kind, p, x, 1e-6)
        K = o.kernel_dense(kind, x, x, p.lengthscale, p.variance) + p.noise * torch.eye(n, dtype=f64)
        precon = o.nystrom_preconditioner(terms.A, terms.LB, p.noise)
        b = y.reshape(-1, 1)
        # conjugate_gradient.py:55-86 with the loop state kept
        v = torch.zeros(n, 1, dtype=f64)
        r = b - K @ v
        z_, rz = precon(r)
        pd = z_
        i = 0
        while 0.5 * rz > max_error and i < max_iter:
            Ap = K @ pd
            gamma = rz / (pd * Ap).sum()
            v = v + gamma * pd
            rs = i % restart == restart - 1
            r = (b - K @ v) if rs else (r - gamma * Ap)
            z_, new_rz = precon(r)
            pd = z_ if rs else (z_ + pd * new_rz / rz)
            rz = new_rz
            i += 1
        # models.py:280-284, the reference's route
        Kv = K @ v
        r_true = b - Kv
        _, eb_true = precon(r_true)
        lower_true = (v * (r_true + 0.5 * Kv)).sum()
        upper_true = lower_true + 0.5 * eb_true
        # the device path's route
        Kv_rec = b - r
        lower_rec = (v * (r + 0.5 * Kv_rec)).sum()
        upper_rec = lower_rec + 0.5 * rz
        logdet = o.logdet_term(p, x, terms)
        bound = -upper_true + logdet - 0.5 * n * math.log(2 * math.pi)
    return dict(steps=i, gap=float((r - r_true).norm() / b.norm()), r_rel=float(r_true.norm() / b.norm()),
                d_upper=float(abs(upper_rec - upper_true)), d_lower=float(abs(lower_rec - lower_true)),
                bound=float(abs(bound)), z_gap=float((z_ - precon(r_true)[0]).norm() / (z_.norm() + 1e-300)))


@pytest.mark.parametrize("kind,n,d,M,noise,ls,max_error,max_iter,restart", [
    ("matern32", 1500, 3, 64, 1.0, 1.0, 1.0, 100, 40),          # reference initial values: a handful of iterations
    ("matern32", 1500, 3, 32, 1e-2, 0.9, 1.0, 100, 40),         # trained-like
    ("matern32", 1200, 11, 16, 1e-3, 1.5, 1e-6, 100, 40),       # badly conditioned, runs into restarts / the cap
    ("rbf", 1000, 8, 16, 1e-4, 2.0, 1e-9, 100, 40),             # RBF spectrum decays fastest: cond ~ 1e7
    ("rbf", 900, 2, 8, 1e-3, 0.7, 1e-9, 39, 1000),              # 39 iterations of pure recurrence, no restart
])
def test_recurrence_residual_gives_the_same_bounds(kind, n, d, M, noise, ls, max_error, max_iter, restart):
    s = _replay(kind, n, d, M, noise, ls, max_error, max_iter, restart, seed=n + d)
    # north_star: bound within 1e-7 relative of the reference -- the two routes differ by < 1e-11 of the bound
    assert s["d_upper"] <= 1e-11 * s["bound"], s
    assert s["d_lower"] <= 1e-11 * s["bound"], s
    # the residual gap itself is rounding-sized relative to the right-hand side
    assert s["gap"] <= 1e-11, s


def test_the_hard_cases_really_iterate():
    s = _replay("matern32", 1200, 11, 16, 1e-3, 1.5, 1e-6, 100, 40, seed=1211)
    assert s["steps"] >= 41          # crosses a restart


if __name__ == "__main__":
    for args in [("matern32", 1500, 3, 64, 1.0, 1.0, 1.0, 100, 40), ("matern32", 1500, 3, 32, 1e-2, 0.9, 1.0, 100, 40),
                 ("matern32", 1200, 11, 16, 1e-3, 1.5, 1e-6, 100, 40), ("rbf", 1000, 8, 16, 1e-4, 2.0, 1e-9, 100, 40),
                 ("rbf", 900, 2, 8, 1e-3, 0.7, 1e-9, 39, 1000)]:
        print(args, _replay(*args, seed=args[1] + args[2]))
