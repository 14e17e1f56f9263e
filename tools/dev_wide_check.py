"""Developer check (GPU): wide-d (d > 32) kernels vs the oracle."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cglb_b200.engine import get_engine
from oracle import cglb_oracle as o
eng = get_engine(); dev = eng.device; f64 = torch.float64
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for kind, n, d in [("matern32", 300, 40), ("rbf", 1500, 64), ("matern32", 2300, 90), ("matern32", 1025, 33), ("rbf", 129, 100)]:
    g = torch.Generator().manual_seed(n)
    x = torch.randn(n, d, generator=g, dtype=f64); v = torch.randn(n, generator=g, dtype=f64)
    ls = (torch.rand(d, generator=g, dtype=f64) + 0.5) * 0.5 * d ** 0.5
    xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    y = eng.kmv_sym(kind, xp, n, d, v.to(dev), 1.3, 0.07)
    K = o.kernel_dense(kind, x, x, ls, torch.tensor(1.3, dtype=f64), block=256)
    ref = K @ v + 0.07 * v
    e1 = float((y.cpu() - ref).norm() / ref.norm())
    parts = sum(eng.kmv_sym(kind, xp, n, d, v.to(dev), 1.3, 0.07, part=p, nparts=3) for p in range(3))
    e2 = float((parts.cpu() - ref).norm() / ref.norm())
    nr = max(1, n // 3)
    xr = eng.pack(kind, x[:nr].contiguous().to(dev), ls.to(dev), x.mean(0).to(dev))
    yr = eng.kmv_rect(kind, xr, nr, xp, n, d, v.to(dev), 1.3)
    e3 = float((yr.cpu() - K[:nr] @ v).norm() / (K[:nr] @ v).norm())
    print(f"{kind} n={n} d={d}: sym {e1:.2e} parts {e2:.2e} rect {e3:.2e}", flush=True)
for kind, n, d in [("matern32", 100000, 90), ("matern32", 515000, 90)]:
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, d, generator=g, dtype=f64, device=dev); v = torch.randn(n, generator=g, dtype=f64, device=dev)
    ls = torch.full((d,), 0.5 * d ** 0.5, dtype=f64, device=dev)
    xp = eng.pack(kind, x, ls, x.mean(0)); y = eng.empty(n)
    ms = timeit(lambda: eng.kmv_sym(kind, xp, n, d, v, 1.0, 0.01, out=y), 2)
    print(f"{kind} n={n} d={d}: {ms:.1f} ms {n*n/ms/1e6:.1f} Gpairs/s  alg TFLOP/s {(2*d+10)*n*n/ms/1e9:.1f}", flush=True)
# ---- backward sweep + K_nm build/backward (wide) -----------------------------------------------------------
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from closed_form_reference import knm_backward as ref_knm_backward
for kind, n, d in [("matern32", 300, 40), ("rbf", 1100, 64), ("matern32", 1500, 90), ("matern32", 130, 33)]:
    g = torch.Generator().manual_seed(n + 1)
    x = torch.randn(n, d, generator=g, dtype=f64); v = torch.randn(n, generator=g, dtype=f64); u = torch.randn(n, generator=g, dtype=f64)
    ls = (torch.rand(d, generator=g, dtype=f64) + 0.5) * 0.5 * d ** 0.5
    xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    out = eng.zeros(d + 1)
    eng.kmv_bwd_sym(kind, xp, n, d, u.to(dev), v.to(dev), 1.3, ls.to(dev), out)
    lsr, varr = ls.clone().requires_grad_(True), torch.tensor(1.3, dtype=f64, requires_grad=True)
    f = u @ (o.kernel_dense(kind, x, x, lsr, varr, block=128) @ v)
    gl, gv = torch.autograd.grad(f, [lsr, varr]); ref = torch.cat([gl.reshape(-1), gv.reshape(1)])
    parts = eng.zeros(d + 1)
    for p in range(2): eng.kmv_bwd_sym(kind, xp, n, d, u.to(dev), v.to(dev), 1.3, ls.to(dev), parts, part=p, nparts=2)
    print(f"bwd {kind} n={n} d={d}: {float((out.cpu()-ref).norm()/ref.norm()):.2e} parts {float((parts.cpu()-ref).norm()/ref.norm()):.2e}", flush=True)
    m = 70
    zz = torch.randn(m, d, generator=g, dtype=f64)
    zp = eng.pack(kind, zz.to(dev), ls.to(dev), x.mean(0).to(dev))
    ld = n + (n % 2); K = eng.zeros(m, ld)
    eng.knm_build(kind, zp, m, xp, n, d, 1.7, K, ld)
    kref = o.kernel_dense(kind, zz, x, ls, torch.tensor(1.7, dtype=f64))
    G = torch.randn(m, n, generator=g, dtype=f64); wt = torch.randn(m, generator=g, dtype=f64); zv = torch.randn(n, generator=g, dtype=f64)
    Gd = torch.zeros(m, ld, dtype=f64); Gd[:, :n] = G
    o_ls = eng.zeros(d); o_var = eng.zeros(1); o_z = eng.zeros(m, d)
    eng.knm_backward(kind, zp, m, xp, n, d, 1.7, ls.to(dev), Gd.to(dev), ld, wt.to(dev), zv.to(dev), o_ls, o_var, o_z)
    rl, rv, rz = ref_knm_backward(kind, zz, x, ls, 1.7, G + wt[:, None] * zv[None, :])
    print(f"knm {kind} n={n} d={d}: build {float((K[:, :n].cpu()-kref).norm()/kref.norm()):.2e} ls {float((o_ls.cpu()-rl).norm()/rl.norm()):.2e} "
          f"var {abs(float(o_var)-float(rv))/abs(float(rv)):.2e} z {float((o_z.cpu()-rz).norm()/rz.norm()):.2e}", flush=True)
kind, n, d = "matern32", 200000, 90
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(n, d, generator=g, dtype=f64, device=dev); v = torch.randn(n, generator=g, dtype=f64, device=dev); u = torch.randn(n, generator=g, dtype=f64, device=dev)
ls = torch.full((d,), 0.5 * d ** 0.5, dtype=f64, device=dev); xp = eng.pack(kind, x, ls, x.mean(0)); out = eng.zeros(d + 1)
ms = timeit(lambda: eng.kmv_bwd_sym(kind, xp, n, d, u, v, 1.0, ls, out), 2)
print(f"bwd {kind} n={n} d={d}: {ms:.1f} ms {n*n/ms/1e6:.1f} Gpairs/s", flush=True)
