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
