"""Developer experiment (GPU): time the (WARPS, TI) variants of the sweeps (needs a CGLB_KMV_EXPERIMENT build)."""
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
for kind, n, d in [("matern32", 200000, 11), ("matern32", 300000, 3), ("rbf", 200000, 8)]:
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, d, generator=g, dtype=torch.float64, device=dev)
    v = torch.randn(n, generator=g, dtype=torch.float64, device=dev); u = torch.randn(n, generator=g, dtype=torch.float64, device=dev)
    ls = torch.full((d,), 0.5 * d ** 0.5, dtype=torch.float64, device=dev)
    xp = eng.pack(kind, x, ls, x.mean(0)); y = eng.empty(n); out = eng.zeros(d + 1)
    ref = None
    for var in ["841", "842", "832", "824", "1222", "1224", "1232", "1622", "1614"]:
        os.environ["CGLB_KMV_VARIANT"] = var
        ms = timeit(lambda: eng.kmv_sym(kind, xp, n, d, v, 1.0, 0.01, out=y))
        if ref is None: ref = y.clone()
        err = float((y - ref).norm() / ref.norm())
        print(f"fwd {kind} n={n} d={d} variant {var}: {ms:8.3f} ms  {n*n/ms/1e6:8.1f} Gpairs/s  relerr_vs_first {err:.1e}", flush=True)
        res[f"fwd_{kind}_{n}_{d}_{var}"] = n * n / ms / 1e6
    for var in []:
        os.environ["CGLB_BWD_VARIANT"] = var
        ms = timeit(lambda: eng.kmv_bwd_sym(kind, xp, n, d, u, v, 1.0, ls, out))
        print(f"bwd {kind} n={n} d={d} variant {var}: {ms:8.3f} ms  {n*n/ms/1e6:8.1f} Gpairs/s", flush=True)
        res[f"bwd_{kind}_{n}_{d}_{var}"] = n * n / ms / 1e6
json.dump(res, open("gpurun_out/dev_variants.json", "w"), indent=1)
