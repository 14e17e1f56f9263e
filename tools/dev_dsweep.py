"""Developer check (GPU): DMMA-distance sweep (dsweep.cu) vs the register-resident sweep and the oracle + timing."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cglb_b200.engine import get_engine
from oracle import cglb_oracle as o
eng = get_engine(); dev = eng.device
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {}
# correctness on small odd shapes against the oracle (forced) and the old kernel
for kind, n, d, lsv in [("matern32", 1, 11, 1.0), ("matern32", 300, 3, 1.0), ("matern32", 2500, 11, 1.0), ("rbf", 2049, 3, 1.0), ("rbf", 1500, 10, 2.0),
                        ("matern32", 1500, 3, 0.05), ("matern32", 3333, 19, 2.0), ("rbf", 1111, 27, 3.0), ("matern32", 1300, 2, 0.7)]:
    g = torch.Generator().manual_seed(n + d)
    x = torch.randn(n, d, generator=g, dtype=torch.float64); v = torch.randn(n, generator=g, dtype=torch.float64)
    ls = (torch.rand(d, generator=g, dtype=torch.float64) + 0.5) * lsv
    K = o.kernel_dense(kind, x, x, ls, torch.tensor(1.3, dtype=torch.float64))
    yref = K @ v + 0.1 * v
    xd, vd, lsd = x.to(dev), v.to(dev), ls.to(dev)
    xp = eng.pack(kind, xd, lsd, xd.mean(0))
    out = {}
    for mode in ["0", "2"]:
        os.environ["CGLB_DSWEEP"] = mode
        y = eng.kmv_sym(kind, xp, n, d, vd, 1.3, 0.1)
        out[mode] = float((y.cpu() - yref).norm() / yref.norm())
        ysum = torch.zeros_like(y)
        for part in range(3):
            ysum += eng.kmv_sym(kind, xp, n, d, vd, 1.3, 0.1, part=part, nparts=3)
        out[mode + "_parts"] = float((ysum.cpu() - yref).norm() / yref.norm())
    print(f"check {kind} n={n} d={d} ls*{lsv}: relerr old {out['0']:.2e} dmma {out['2']:.2e} (3 parts: {out['0_parts']:.2e} / {out['2_parts']:.2e})", flush=True)
    res[f"check_{kind}_{n}_{d}"] = out
for kind, n, d in [("matern32", 200000, 11), ("matern32", 434000, 3), ("rbf", 200000, 10), ("matern32", 100000, 19), ("matern32", 50000, 11)]:
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, d, generator=g, dtype=torch.float64, device=dev)
    v = torch.randn(n, generator=g, dtype=torch.float64, device=dev)
    ls = torch.full((d,), 0.5 * d ** 0.5, dtype=torch.float64, device=dev)
    xp = eng.pack(kind, x, ls, x.mean(0)); y = eng.empty(n)
    ref = None
    for mode in ["0", "2"]:
        os.environ["CGLB_DSWEEP"] = mode
        ms = timeit(lambda: eng.kmv_sym(kind, xp, n, d, v, 1.0, 0.01, out=y))
        if ref is None: ref = y.clone()
        err = float((y - ref).norm() / ref.norm())
        print(f"fwd {kind} n={n} d={d} dsweep={mode}: {ms:8.3f} ms  {n*n/ms/1e6:8.1f} Gpairs/s  relerr_vs_old {err:.1e}", flush=True)
        res[f"fwd_{kind}_{n}_{d}_{mode}"] = n * n / ms / 1e6
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/dev_dsweep.json", "w"), indent=1)
