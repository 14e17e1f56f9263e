"""GPU counterpart of tools/grad_spread_cpu.py: bound / gradient differences to the reference's golden vectors along the
golden CG trajectories, for the two routes of bound.py (CG-state reuse, recomputed residual) and run-to-run.
Test infrastructure (reads tests/golden, never used by the product).

    python tools/grad_spread_gpu.py > gpurun_out/grad_spread_gpu.md
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cglb_b200 as cb                                           # noqa: E402
from conftest import GOLDEN_CASES, GOLDEN_DIR, GRAD_NAMES      # noqa: E402
from helpers import make_model                                  # noqa: E402


def run(name, reuse):
    g = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    kind = str(g["kind"])
    model = make_model(kind, g["x"], g["y"], g["z"], float(g["noise"]), float(g["variance"]), g["lengthscale"], float(g["mean_c"]))
    cg = cb.ConjugateGradient(max_error=float(g["cg_max_error"]), max_cg_iter=int(g["cg_max_iter"]), restart_cg_iter=int(g["cg_restart"]))
    lb = cb.LowerBoundCG(model, cg_opt=cg)
    data = (model.train_inputs[0], model.train_targets)
    lb.evaluator(data).reuse_cg_state = reuse
    params = list(model.parameters())
    out = []
    for e, mult in enumerate(g["ls_mults"]):
        model.covar_module.base_kernel.base_kernel.lengthscale = torch.as_tensor(g["lengthscale"] * mult)
        loss = -lb(data)
        grads = [t.cpu().numpy() for t in torch.autograd.grad(loss, params)]
        rels = {nm: float(np.abs(gr - g[f"grad_{nm}_{e}"]).max() / (np.abs(g[f"grad_{nm}_{e}"]).max() + 1e-300)) for nm, gr in zip(GRAD_NAMES, grads)}
        out.append(dict(e=e, k=int(model.cg_stats.steps), kg=int(g[f"cg_steps_{e}"]),
                        dl=abs(float(loss) - float(g[f"loss_{e}"])) / abs(float(g[f"loss_{e}"])),
                        dv=float(np.abs(model.v_vec.cpu().numpy() - g[f"v_{e}"]).max() / np.abs(g[f"v_{e}"]).max()),
                        rels=rels, grads=grads, loss=float(loss)))
    return out


def main():
    print("# GPU: differences to the reference's golden vectors along the golden CG trajectories\n")
    print("| case | eval | route | CG its (golden) | rel. bound diff | rel. v diff | worst rel. gradient diff | parameter | run-to-run worst gradient diff |")
    print("|---|---|---|---|---|---|---|---|---|")
    worst_all = {}
    for name in GOLDEN_CASES:
        for reuse in (True, False):
            a, b = run(name, reuse), run(name, reuse)
            for ra, rb in zip(a, b):
                which = max(ra["rels"], key=ra["rels"].get)
                rr = max(float(np.abs(x - y).max() / (np.abs(x).max() + 1e-300)) for x, y in zip(ra["grads"], rb["grads"]))
                route = "reuse" if reuse else "recompute"
                print(f"| {name} | {ra['e']} | {route} | {ra['k']} ({ra['kg']}) | {ra['dl']:.1e} | {ra['dv']:.1e} | {ra['rels'][which]:.1e} | {which} | {rr:.1e} |")
                worst_all[route] = max(worst_all.get(route, 0.0), ra["rels"][which])
    print()
    for k, v in worst_all.items():
        print(f"* worst gradient difference, route `{k}`: **{v:.1e}**")


if __name__ == "__main__":
    main()
