"""Small run through every kernel for compute-sanitizer (one tool per gpurun call)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cglb_b200 as cb
from helpers import make_model
from oracle import cglb_oracle as o
for kind, n, d, M in [("matern32", 700, 3, 40), ("rbf", 300, 11, 17), ("matern32", 390, 40, 24), ("matern32", 1100, 90, 33), ("rbf", 257, 20, 16)]:
    x, y, z = o.synthetic_problem(n, d, M, seed=n)
    model = make_model(kind, x.numpy(), y.numpy(), z.numpy(), 0.05, 1.2, 0.5 * d ** 0.5, 0.1)
    lb = cb.LowerBoundCG(model)
    loss = -lb((model.train_inputs[0], model.train_targets))
    grads = torch.autograd.grad(loss, list(model.parameters()))
    pred = cb.PredictCG(model)
    fm, fv = pred(torch.randn(65, d, dtype=torch.float64, device="cuda"))
    torch.cuda.synchronize()
    print(kind, n, d, M, float(loss), int(model.cg_stats.steps), float(fm.sum()), flush=True)
print("done")
