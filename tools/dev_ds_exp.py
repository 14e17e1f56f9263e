"""Developer experiment (GPU): timing decomposition of the DMMA sweep (needs a -DCGLB_DS_EXPERIMENT build)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cglb_b200.engine import get_engine
eng = get_engine(); dev = eng.device
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
kind, n, d = "matern32", 200000, 11
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(n, d, generator=g, dtype=torch.float64, device=dev)
v = torch.randn(n, generator=g, dtype=torch.float64, device=dev)
ls = torch.full((d,), 0.5 * d ** 0.5, dtype=torch.float64, device=dev)
xp = eng.pack(kind, x, ls, x.mean(0)); y = eng.empty(n)
os.environ["CGLB_DSWEEP"] = "0"
ms0 = timeit(lambda: eng.kmv_sym(kind, xp, n, d, v, 1.0, 0.01, out=y)); ref = y.clone()
print(f"register sweep: {ms0:8.3f} ms  {n*n/ms0/1e6:8.1f} Gpairs/s", flush=True)
os.environ["CGLB_DSWEEP"] = "2"
os.environ.pop("CGLB_DS_EXP", None)
eng.kmv_sym(kind, xp, n, d, v, 1.0, 0.01, out=y)
print("dsweep relerr vs register sweep:", float((y - ref).norm() / ref.norm()), flush=True)
res = {}
for e in sys.argv[1:] or ["0", "1", "2", "3", "4", "8", "16", "32", "28", "60", "61"]:
    os.environ["CGLB_DS_EXP"] = e
    ms = timeit(lambda: eng.kmv_sym(kind, xp, n, d, v, 1.0, 0.01, out=y))
    print(f"EXP={e}: {ms:8.3f} ms  {n*n/ms/1e6:8.1f} Gpairs/s", flush=True)
    res[e] = ms
json.dump(res, open("gpurun_out/dev_ds_exp.json", "w"), indent=1)
