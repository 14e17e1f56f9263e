"""Developer check (GPU): model-level bound + grads vs golden."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cglb_b200 as cb
from helpers import make_model, rel_max
names = sys.argv[1:] or ["snelson_like_init", "road_like_trained", "kin_like_rbf", "house_like_warmstart", "ragged_rbf_init", "restart_path"]
G = ["raw_noise", "mean_constant", "inducing_points", "raw_outputscale", "raw_lengthscale"]
for name in names:
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    kind = str(g["kind"])
    model = make_model(kind, g["x"], g["y"], g["z"], float(g["noise"]), float(g["variance"]), g["lengthscale"], float(g["mean_c"]))
    cg = cb.ConjugateGradient(max_error=float(g["cg_max_error"]), max_cg_iter=int(g["cg_max_iter"]), restart_cg_iter=int(g["cg_restart"]))
    lb = cb.LowerBoundCG(model, cg_opt=cg)
    data = (model.train_inputs[0], model.train_targets)
    params = list(model.parameters())
    for e, mult in enumerate(g["ls_mults"]):
        model.covar_module.base_kernel.base_kernel.lengthscale = torch.as_tensor(g["lengthscale"] * mult)
        loss = -lb(data)
        grads = torch.autograd.grad(loss, params)
        print(name, e, "loss", float(loss), float(g[f"loss_{e}"]), "rel", abs(float(loss) - float(g[f"loss_{e}"])) / abs(float(g[f"loss_{e}"])),
              "cg", model.cg_stats.steps, int(g[f"cg_steps_{e}"]))
        for nm, gr in zip(G, grads):
            print("   grad", nm, rel_max(gr.cpu().numpy(), g[f"grad_{nm}_{e}"]))
    pred = cb.PredictCG(model)
    fm, fv = pred(torch.as_tensor(g["xnew"]).cuda())
    print("   predict mean", rel_max(fm.cpu().numpy(), g["f_mean"]), "var", rel_max(fv.cpu().numpy(), g["f_var"]))
