"""Developer timing (GPU): forward (register / DMMA) and backward sweeps at the bench shapes."""
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
res = {}
shapes = [("matern32", 300000, 3), ("rbf", 100000, 4), ("matern32", 100000, 5), ("rbf", 100000, 6), ("matern32", 100000, 7), ("matern32", 100000, 9)] if len(sys.argv) > 2 else [("matern32", 200000, 11), ("rbf", 100000, 8), ("matern32", 100000, 10), ("matern32", 100000, 12), ("matern32", 100000, 13), ("rbf", 100000, 16), ("matern32", 100000, 19), ("matern32", 100000, 24), ("matern32", 100000, 32)] if len(sys.argv) > 1 else [("matern32", 200000, 11), ("matern32", 300000, 3), ("rbf", 40000, 8), ("rbf", 200000, 8), ("matern32", 100000, 19), ("matern32", 100000, 20)]
for kind, n, d in shapes:
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, d, generator=g, dtype=torch.float64, device=dev)
    v = torch.randn(n, generator=g, dtype=torch.float64, device=dev); u = torch.randn(n, generator=g, dtype=torch.float64, device=dev)
    ls = torch.full((d,), 0.5 * d ** 0.5, dtype=torch.float64, device=dev)
    xp = eng.pack(kind, x, ls, x.mean(0)); y = eng.empty(n); out = eng.zeros(d + 1)
    for mode in (["0", "2"] if len(sys.argv) > 1 else ["0", "1"]):
        os.environ["CGLB_DSWEEP"] = mode
        ms = timeit(lambda: eng.kmv_sym(kind, xp, n, d, v, 1.0, 0.01, out=y))
        print(f"fwd {kind} n={n} d={d} dsweep={mode}: {ms:8.3f} ms  {n*n/ms/1e6:8.1f} Gpairs/s", flush=True)
        res[f"fwd_{kind}_{n}_{d}_{mode}"] = n * n / ms / 1e6
    for mode in (["0", "2"] if len(sys.argv) > 1 else ["0", "1"]):
        os.environ["CGLB_DSWEEP"] = mode
        ms = timeit(lambda: eng.kmv_bwd_sym(kind, xp, n, d, u, v, 1.0, ls, out))
        print(f"bwd {kind} n={n} d={d} dsweep={mode}: {ms:8.3f} ms  {n*n/ms/1e6:8.1f} Gpairs/s", flush=True)
        res[f"bwd_{kind}_{n}_{d}_{mode}"] = n * n / ms / 1e6
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/dev_time_sweeps.json", "w"), indent=1)
