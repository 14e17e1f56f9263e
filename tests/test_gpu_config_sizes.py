"""GPU (-m gpu): bound + gradients at the Nystrom sizes of BASELINE.json's configs (M = 1024 / 2048), against the oracle.

The goldens stop at M = 64; here the blocked potrf (8-16 panels), the 128-block TRSM, the split-K SYRK, `LB^-1` applied as
an explicit inverse (the reference uses two triangular solves, conjugate_gradient.py:106-107) and the fused K_nm backward run
at the real M through the whole bound (models.py:176-213, 246-286), at sigma^2 = 0.01 where cond(B) is large.
  #1 snelson-shaped   n = 2000 d = 1  Matern32 M = 1024: FULL size, CG trajectory + gradients (dense oracle)
  #2 kin40k-shaped    n = 6000 d = 8  RBF      M = 1024: sub-sampled n, fixed v
  #3 3droad-shaped    n = 12000 d = 3 Matern32 M = 2048, sigma^2 = 0.01: sub-sampled n, fixed v
  #5 houseelectric-   n = 12000 d = 11 Matern32 M = 2048, sigma^2 = 0.01: sub-sampled n, fixed v
Fixed v = the reference's cached-v route (models.py:263-264): every term is then a deterministic function of v, so the
comparison carries tight tolerances (bound 1e-9, gradients 1e-7 -- the north_star's; measured values in the asserts'
messages).  The oracle evaluates `cov @ v` in row blocks with exact two-stage autograd
(oracle.bound_and_grads_fixed_v_blocked, checked against the dense oracle on CPU)."""
import math

import numpy as np
import pytest
import torch

import cglb_b200 as cb
from oracle import cglb_oracle as o
from conftest import GRAD_NAMES
from helpers import make_model

pytestmark = pytest.mark.gpu
f64 = torch.float64


def _compare(loss, grads, ref_loss, ref_grads, bound_tol, grad_tol, tag):
    rel = abs(float(loss) - float(ref_loss)) / abs(float(ref_loss))
    assert rel <= bound_tol, (tag, "bound", rel)
    for nm, a, b in zip(GRAD_NAMES, grads, ref_grads):
        err = float(np.abs(a.detach().cpu().numpy() - b.numpy()).max() / (np.abs(b.numpy()).max() + 1e-300))
        assert err <= grad_tol, (tag, nm, err)


@pytest.mark.parametrize("theta", ["init", "trained"])
def test_config1_snelson_shaped_full_size(theta):
    """n = 2000, d = 1, Matern32, M = 1024 (BASELINE.json configs[0]) through the CG solve, at the reference's initial
    hyper-parameters (config.py:76,105) and at a trained-like point."""
    n, d, M = 2000, 1, 1024
    x, y, z = o.synthetic_problem(n, d, M, seed=0)
    noise, var, ls = (1.0, 1.0, 1.0) if theta == "init" else (0.01, 1.0, 0.5)
    model = make_model("matern32", x.numpy(), y.numpy(), z.numpy(), noise, var, ls, 0.0)
    lb = cb.LowerBoundCG(model)
    loss = -lb((model.train_inputs[0], model.train_targets))
    grads = torch.autograd.grad(loss, list(model.parameters()))
    p = o.OracleParams.from_values(noise, 0.0, z, var, ls)
    ref_loss, ref_grads, res = o.bound_and_grads("matern32", p, x, y, torch.zeros(n, 1, dtype=f64))
    assert abs(int(model.cg_stats.steps) - res.cg.steps) <= 1
    if int(model.cg_stats.steps) == res.cg.steps:
        _compare(loss, grads, ref_loss, ref_grads, 1e-7, 1e-7, f"snelson-{theta}")


@pytest.mark.parametrize("kind,n,d,M,noise,ls_scale", [("rbf", 6000, 8, 1024, 0.05, 0.5),
                                                      ("matern32", 12000, 3, 2048, 0.01, 0.5),
                                                      ("matern32", 12000, 11, 2048, 0.01, 0.5)])
def test_bound_and_gradients_at_nystrom_sizes_fixed_v(kind, n, d, M, noise, ls_scale):
    x, y, z = o.synthetic_problem(n, d, M, seed=d)
    ls = np.linspace(0.9, 1.1, d) * ls_scale * math.sqrt(d)
    model = make_model(kind, x.numpy(), y.numpy(), z.numpy(), noise, 1.2, ls, 0.05)
    v = 0.1 * torch.randn(n, 1, dtype=f64, generator=torch.Generator().manual_seed(1))
    model.v_vec.data.copy_(v.cuda())
    lb = cb.LowerBoundCG(model, use_cache=True, cached_v_vec_initial=True)
    loss = -lb((model.train_inputs[0], model.train_targets))
    grads = torch.autograd.grad(loss, list(model.parameters()))
    p = o.OracleParams.from_values(noise, 0.05, z, 1.2, ls)
    ref_loss, ref_grads = o.bound_and_grads_fixed_v_blocked(kind, p, x, y, v, block=256)
    _compare(loss, grads, ref_loss, ref_grads, 1e-9, 1e-7, f"{kind}-n{n}-d{d}-M{M}")
    # the preconditioner with the explicit LB^-1 against two triangular solves (conjugate_gradient.py:106-107) at this M
    ev = lb.evaluator((model.train_inputs[0], model.train_targets))
    terms = ev.terms
    pre = ev.preconditioner(terms, noise)
    r = torch.randn(n, 1, dtype=f64, generator=torch.Generator().manual_seed(2)).cuda()
    zg, rzg = pre(r)
    A, LB = terms.A[:, :n].cpu(), terms.LB.cpu()
    zr, rzr = o.nystrom_preconditioner(A, LB, torch.tensor(noise, dtype=f64))(r.cpu())
    assert float((zg.cpu() - zr).abs().max() / zr.abs().max()) <= 1e-9
    assert abs(float(rzg) - float(rzr)) <= 1e-9 * abs(float(rzr))
