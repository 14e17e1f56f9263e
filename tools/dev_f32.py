"""Developer check (GPU): fp32-pair K*v sweep vs the fp64 sweep and the oracle + timing."""
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
for kind, n, d, lsv in [("matern32", 1, 3, 1.0), ("matern32", 300, 1, 1.0), ("matern32", 2500, 11, 1.0), ("rbf", 2049, 3, 1.0), ("rbf", 1500, 8, 2.0),
                        ("matern32", 1500, 3, 0.05), ("matern32", 3333, 19, 2.0), ("rbf", 1111, 32, 3.0), ("matern32", 5000, 16, 1.0)]:
    g = torch.Generator().manual_seed(n + d)
    x = torch.randn(n, d, generator=g, dtype=torch.float64); v = torch.randn(n, generator=g, dtype=torch.float64)
    ls = (torch.rand(d, generator=g, dtype=torch.float64) + 0.5) * lsv * (0.5 * d ** 0.5)
    K = o.kernel_dense(kind, x, x, ls, torch.tensor(1.3, dtype=torch.float64))
    yref = K @ v + 0.1 * v
    xd, vd, lsd = x.to(dev), v.to(dev), ls.to(dev)
    xpf = eng.pack_f32(kind, xd, lsd, xd.mean(0))
    y = eng.kmv_sym_f32(kind, xpf, n, d, vd, 1.3, 0.1)
    e1 = float((y.cpu() - yref).norm() / yref.norm())
    ysum = torch.zeros_like(y)
    for part in range(3):
        ysum += eng.kmv_sym_f32(kind, xpf, n, d, vd, 1.3, 0.1, part=part, nparts=3)
    e2 = float((ysum.cpu() - yref).norm() / yref.norm())
    print(f"check {kind} n={n} d={d} ls*{lsv}: relerr f32 {e1:.2e} (3 parts {e2:.2e})", flush=True)
    res[f"check_{kind}_{n}_{d}"] = [e1, e2]
for kind, n, d in [("matern32", 200000, 11), ("matern32", 434000, 3), ("rbf", 200000, 8), ("matern32", 100000, 19), ("rbf", 40000, 8)]:
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, d, generator=g, dtype=torch.float64, device=dev)
    v = torch.randn(n, generator=g, dtype=torch.float64, device=dev)
    ls = torch.full((d,), 0.5 * d ** 0.5, dtype=torch.float64, device=dev)
    xp = eng.pack(kind, x, ls, x.mean(0)); xpf = eng.pack_f32(kind, x, ls, x.mean(0)); y = eng.empty(n); y2 = eng.empty(n)
    ms64 = timeit(lambda: eng.kmv_sym(kind, xp, n, d, v, 1.0, 0.01, out=y))
    ms32 = timeit(lambda: eng.kmv_sym_f32(kind, xpf, n, d, v, 1.0, 0.01, out=y2))
    err = float((y2 - y).norm() / y.norm())
    print(f"fwd {kind} n={n} d={d}: fp64 {ms64:8.3f} ms {n*n/ms64/1e6:8.1f} Gpairs/s | fp32-pair {ms32:8.3f} ms {n*n/ms32/1e6:8.1f} Gpairs/s  relerr {err:.1e}", flush=True)
    res[f"fwd_{kind}_{n}_{d}"] = [n * n / ms64 / 1e6, n * n / ms32 / 1e6, err]
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/dev_f32.json", "w"), indent=1)
