"""Dense torch statement of the hand-derived backward pass that cglb_b200/engine.py implements with the
CUDA kernels.  Test infrastructure: it is checked against autograd of the oracle
on CPU (test_closed_form_backward.py), and the GPU tests check the engine against the oracle directly."""
import math

import torch

from oracle import cglb_oracle as o


def kernel_and_derivs(kind, z, x, ls, var):
    """k(z_m, x_i) and the per-pair factors: dk/dl_q = var*cfac*ew*delta_q^2/l_q, dk/dz_mq = -var*cfac*ew*delta_q*cs/l_q
    with delta = cs (z - x)/l."""
    cs = math.sqrt(3.0) if kind == "matern32" else math.sqrt(0.5)
    cfac = 1.0 if kind == "matern32" else 2.0
    delta = cs * (z[:, None, :] - x[None, :, :]) / ls.reshape(1, 1, -1)
    q = (delta * delta).sum(-1)
    if kind == "matern32":
        s = torch.sqrt(q.clamp_min(1e-300))
        ew = torch.exp(-s)
        kap = (1 + s) * ew
    else:
        ew = torch.exp(-q)
        kap = ew
    return kap, ew, delta, cs, cfac


def knm_backward(kind, z, x, ls, var, G):
    kap, ew, delta, cs, cfac = kernel_and_derivs(kind, z, x, ls, var)
    gp = G * ew * var * cfac
    g_ls = (gp[:, :, None] * delta * delta).sum((0, 1)) / ls.reshape(-1)
    g_var = (G * kap).sum()
    g_z = -(gp[:, :, None] * delta).sum(1) * cs / ls.reshape(1, -1)
    return g_ls, g_var, g_z


def closed_form_grads(kind, x, y, Z, ls, var, noise, c, v, jitter=1e-6):
    """Gradients of the BOUND (not the loss) w.r.t. (noise, c, Z, var, ls) for a fixed CG solution v."""
    n, M = x.shape[0], Z.shape[0]
    ls = ls.reshape(-1)
    sigma = math.sqrt(noise)
    eye = torch.eye(M, dtype=x.dtype)
    Kuf = o.kernel_dense(kind, Z, x, ls, torch.tensor(var, dtype=x.dtype))
    Kuu = o.kernel_dense(kind, Z, Z, ls, torch.tensor(var, dtype=x.dtype)) + jitter * eye
    L = torch.linalg.cholesky(Kuu)
    A = torch.linalg.solve_triangular(L, Kuf, upper=False) / sigma
    AAt = A @ A.T
    B = AAt + eye
    LB = torch.linalg.cholesky(B)
    Kxx = o.kernel_dense(kind, x, x, ls, torch.tensor(var, dtype=x.dtype))
    err = y.reshape(-1, 1) - c
    Kv = Kxx @ v + noise * v
    r = err - Kv
    qv = A @ r
    LBinv = torch.linalg.inv(LB)
    Binv = LBinv.T @ LBinv
    w = Binv @ qv
    zv = (r - A.T @ w) / noise
    eb = float((r * zv).sum())
    t = n * var / noise - float(torch.trace(AAt))
    a = 1.0 / (1.0 + t / n)
    u = 0.5 * v + zv
    # (1) K_xx part: u^T dK v
    kap, ew, delta, cs, cfac = kernel_and_derivs(kind, x, x, ls, var)
    uv = u @ v.T
    g_ls = (uv[:, :, None] * (ew * var * cfac)[:, :, None] * delta * delta).sum((0, 1)) / ls
    g_var = (uv * kap).sum()
    # (3) S part through A
    Linv = torch.linalg.inv(L)
    Mx = a * AAt - eye + Binv + (w @ w.T) / noise
    H = Linv.T @ (a * eye - Binv) / sigma
    wt = Linv.T @ w / sigma
    G_kuf = H @ A + wt @ zv.T
    G_kuu = -0.5 * Linv.T @ Mx @ Linv
    l1, v1, z1 = knm_backward(kind, Z, x, ls, var, G_kuf)
    l2, v2, z2 = knm_backward(kind, Z, Z, ls, var, G_kuu)
    g_ls = g_ls + l1 + l2
    g_var = g_var + v1 + v2 - 0.5 * a * n / noise
    g_Z = z1 + 2.0 * z2
    g_noise = float((u * v).sum()) - n / (2 * noise) + 0.5 * a * n * var / noise ** 2 + 0.5 * eb / noise \
        - float(torch.trace(Mx)) / (2 * noise)
    g_c = float((v + zv).sum())
    return dict(noise=g_noise, c=g_c, Z=g_Z, var=float(g_var), ls=g_ls)
