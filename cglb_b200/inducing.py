"""Greedy conditional-variance inducing-point initialisation.

The reference calls `robustgp.ConditionalVariance(sample=False)(X, M, kernel_fn)` (config.py:62-65), a
third-party routine (markvdw/RobustGP, un-pinned git dependency, absent here).  Its published algorithm
(Burt et al. 2020, greedy pivoted-Cholesky / MAP of the M-DPP) is restated: start from the point of
largest prior variance, then repeatedly add the point whose variance conditioned on the chosen set is
largest.  `kernel_fn(x1, x2, full_cov)` follows the reference callback (interface.py:278-288):
diagonal when full_cov=False, dense cross-covariance otherwise.  Runs on the host in numpy, outside the
timed step, exactly where the reference runs it.
"""
from __future__ import annotations

import numpy as np


class ConditionalVariance:
    def __init__(self, sample: bool = False, threshold: float = 0.0, seed: int = 0):
        self.sample, self.threshold, self.seed = sample, threshold, seed

    def __call__(self, training_inputs: np.ndarray, M: int, kernel):
        X = np.asarray(training_inputs)
        N = X.shape[0]
        rng = np.random.RandomState(self.seed)
        perm = rng.permutation(N)              # robustgp permutes first so that ties are broken at random
        X = X[perm]
        M = min(M, N)
        indices = np.zeros(M, dtype=int)
        di = np.asarray(kernel(X, None, full_cov=False), dtype=np.float64).reshape(-1) + 1e-12
        if self.sample:
            indices[0] = rng.choice(N, p=di / di.sum())
        else:
            indices[0] = int(np.argmax(di))
        ci = np.zeros((M - 1, N)) if M > 1 else np.zeros((0, N))
        for m in range(M - 1):
            j = int(indices[m])
            new_Z = X[j:j + 1]
            dj = np.sqrt(di[j])
            cj = ci[:m, j]
            Lraw = np.asarray(kernel(X, new_Z, full_cov=True), dtype=np.float64).reshape(-1)
            Lraw[j] += 1e-12
            ei = (Lraw - cj @ ci[:m]) / dj
            ci[m] = ei
            di = np.clip(di - ei ** 2, 0.0, None)
            if self.sample:
                indices[m + 1] = rng.choice(N, p=di / di.sum())
            else:
                indices[m + 1] = int(np.argmax(di))
            if di.sum() < self.threshold:
                indices = indices[:m + 2]
                break
        Z = X[indices]
        return Z, perm[indices]


def conditional_variance_gpu(training_inputs: np.ndarray, M: int, kernel_module, seed: int = 0):
    """Same greedy selection with X, the running diagonal and the pivoted-Cholesky rows resident on the GPU
    (SURVEY.md section 8f-2): kernel columns k(X, x_j) come from the sm_100a `cglb_knm_build` kernel, the rank
    updates are device GEMVs; only the argmax index crosses to the host each iteration.  Memory: (M-1) x n
    doubles, so this is for the sizes the reference itself initialises this way (n up to ~1e5 at M = 1024)."""
    import torch
    from .engine import get_engine
    from .operators import _kernel_pieces
    X = np.asarray(training_inputs)
    N = X.shape[0]
    rng = np.random.RandomState(seed)
    perm = rng.permutation(N)
    M = min(M, N)
    eng = get_engine()
    dev = eng.device
    x = torch.as_tensor(X[perm], dtype=torch.float64, device=dev).contiguous()
    kind, ls, var = _kernel_pieces(kernel_module)
    ls = ls.detach().to(torch.float64).reshape(-1).contiguous()
    var = float(var)
    d = x.shape[1]
    shift = x.mean(0).contiguous()
    xp = eng.pack(kind, x, ls, shift)
    di = torch.full((N,), var + 1e-12, dtype=torch.float64, device=dev)
    ci = torch.zeros(max(M - 1, 0), N, dtype=torch.float64, device=dev)
    indices = np.zeros(M, dtype=int)
    indices[0] = int(torch.argmax(di).item())
    ld = N + (N & 1)
    col = eng.zeros(1, ld)
    for m in range(M - 1):
        j = int(indices[m])
        zp = eng.pack(kind, x[j:j + 1].contiguous(), ls, shift)
        eng.knm_build(kind, zp, 1, xp, N, d, var, col, ld)
        lraw = col[0, :N].clone()
        lraw[j] += 1e-12
        dj = torch.sqrt(di[j])
        ei = (lraw - torch.mv(ci[:m].t(), ci[:m, j])) / dj if m > 0 else lraw / dj
        ci[m] = ei
        di = torch.clamp(di - ei * ei, min=0.0)
        indices[m + 1] = int(torch.argmax(di).item())
    Z = X[perm][indices]
    return Z, perm[indices]
