"""CPU measurement behind the gradient tolerance of the CG-trajectory tests (tests/test_gpu_parity.py).

Question: along a warm-started CG trajectory stopped at 0.5 r^T P r <= max_error (conjugate_gradient.py:65), how far do
the gradients of two EQUALLY VALID fp64 evaluations of the same algorithm drift apart?  The oracle's loop is replayed on
every golden case with K v evaluated (a) as the golden vectors were (dense K, direct-difference distances), (b) with the
columns summed in a permuted order (K[:, perm] @ v[perm]: the same numbers, another summation order -- what any tiled or
parallel K v does), (c) with the expanded-form distances |a|^2 + |b|^2 - 2 a.b of GPyTorch's dense path
(oracle.sqdist_expanded).  Only the CG solve is perturbed; bound and gradients are then evaluated by the unmodified
oracle at the v each variant returns.  Output: one row per (case, evaluation) with the largest relative gradient
difference to the reference's golden gradients.  Test infrastructure: imports oracle/, never used by the product.

    python tools/grad_spread_cpu.py > profiles/grad_spread_cpu_r02.md
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import cglb_oracle as o          # noqa: E402
from conftest import GOLDEN_CASES, GOLDEN_DIR, GRAD_NAMES      # noqa: E402

f64 = torch.float64


class PermutedSum:
    def __init__(self, K, s2, seed):
        n = K.shape[0]
        self.perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed))
        self.Kp, self.s2 = K[:, self.perm].contiguous(), s2

    def __matmul__(self, v):
        return self.Kp @ v[self.perm] + self.s2 * v


class Plain:
    def __init__(self, K, s2):
        self.K, self.s2 = K, s2

    def __matmul__(self, v):
        return self.K @ v + self.s2 * v


def run_case(name, variant):
    g = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    kind = str(g["kind"])
    x, y, z = (torch.from_numpy(g[k]) for k in ("x", "y", "z"))
    v = torch.zeros(x.shape[0], 1, dtype=f64)
    rows = []
    for e, mult in enumerate(g["ls_mults"]):
        p = o.OracleParams.from_values(float(g["noise"]), float(g["mean_c"]), z, float(g["variance"]), g["lengthscale"] * mult)
        with torch.no_grad():
            terms = o.common_terms(kind, p, x, float(g["jitter"]))
            s2 = p.noise
            if variant == "expanded":
                sq = o.sqdist_expanded(x, x, p.lengthscale, x1_eq_x2=True)
                if kind == "matern32":
                    s = o.SQRT3 * torch.sqrt(sq.clamp_min(1e-30))
                    K = p.variance * (1.0 + s) * torch.exp(-s)
                else:
                    K = p.variance * torch.exp(-0.5 * sq)
                A = Plain(K, s2)
            else:
                K = o.kernel_dense(kind, x, x, p.lengthscale, p.variance)
                A = Plain(K, s2) if variant == "golden" else PermutedSum(K, s2, seed=e + 1)
            err = y.reshape(-1, 1) - p.mean_constant.reshape(1, 1)
            precon = o.nystrom_preconditioner(terms.A, terms.LB, s2)
            v, st = o.conjugate_gradient(A, err, v, precon, float(g["cg_max_error"]), int(g["cg_max_iter"]), int(g["cg_restart"]))
        loss, grads, _ = o.bound_and_grads(kind, p, x, y, v, jitter=float(g["jitter"]), use_cached_v=True)
        worst, which = 0.0, ""
        for nm, gr in zip(GRAD_NAMES, grads):
            ref = g[f"grad_{nm}_{e}"]
            rel = float(np.abs(gr.numpy() - ref).max() / (np.abs(ref).max() + 1e-300))
            if rel > worst:
                worst, which = rel, nm
        rows.append((e, st.steps, int(g[f"cg_steps_{e}"]), abs(float(loss) - float(g[f"loss_{e}"])) / abs(float(g[f"loss_{e}"])),
                     float(np.abs(v.numpy() - g[f"v_{e}"]).max() / np.abs(g[f"v_{e}"]).max()), worst, which))
    return rows


def main():
    print("# Gradient spread of equally valid fp64 evaluations along the golden CG trajectories (CPU, oracle)\n")
    print("`python tools/grad_spread_cpu.py` -- see the docstring.  rel = max |g - g_golden| / max |g_golden| over the entries of one parameter.\n")
    print("| case | eval | K v in the CG solve | CG its (golden) | rel. bound diff | rel. v diff | worst rel. gradient diff | parameter |")
    print("|---|---|---|---|---|---|---|---|")
    overall = {}
    for name in GOLDEN_CASES:
        for variant in ("golden", "permuted", "expanded"):
            for e, k, kg, dl, dv, worst, which in run_case(name, variant):
                print(f"| {name} | {e} | {variant} | {k} ({kg}) | {dl:.1e} | {dv:.1e} | {worst:.1e} | {which} |")
                overall[variant] = max(overall.get(variant, 0.0), worst)
    print()
    for k, val in overall.items():
        print(f"* worst gradient difference, K v evaluated as `{k}`: **{val:.1e}**")


if __name__ == "__main__":
    main()
