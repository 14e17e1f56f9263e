"""Developer check (GPU): dense / vector entry points vs torch fp64.  Not part of the test-suite."""
import os, sys, json, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cglb_b200.engine import get_engine
from oracle import cglb_oracle as o
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from closed_form_reference import knm_backward as ref_knm_backward

eng = get_engine()
dev = eng.device
f64 = torch.float64
def rel(a, b): return float((a - b).norm() / (b.norm() + 1e-300))
def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
res = {}
g = torch.Generator(device="cpu").manual_seed(0)
for (m, n, k) in [(128, 256, 64), (200, 130, 70), (1024, 1000, 512), (130, 2050, 1030)]:
    a = torch.randn(m, k, generator=g, dtype=f64).to(dev); b = torch.randn(k, n, generator=g, dtype=f64).to(dev)
    bt = b.t().contiguous(); c0 = torch.randn(m, n, generator=g, dtype=f64).to(dev)
    c = c0.clone(); eng.gemm(a, b, c, m, n, k, transb=False, alpha=1.5, beta=0.5)
    res[f"gemm_nn_{m}_{n}_{k}"] = rel(c, 1.5 * a @ b + 0.5 * c0)
    c = c0.clone(); eng.gemm(a, bt, c, m, n, k, transb=True, alpha=-1.0, beta=0.0)
    res[f"gemm_nt_{m}_{n}_{k}"] = rel(c, -a @ b)
for m in [64, 128, 200, 1024, 2048]:
    x = torch.randn(m, m + 50, generator=g, dtype=f64).to(dev)
    spd = x @ x.t() / m + torch.eye(m, dtype=f64, device=dev)
    l = spd.clone(); eng.potrf(l)
    lref = torch.linalg.cholesky(spd)
    res[f"potrf_{m}"] = rel(l, lref)
    linv = eng.tri_inverse(l)
    res[f"trinv_{m}"] = rel(linv @ l, torch.eye(m, dtype=f64, device=dev))
    n = 1000 if m < 1024 else 4096
    b = torch.randn(m, n, generator=g, dtype=f64).to(dev)
    bs = b.clone(); eng.trsm_left_lower(l, bs, n, alpha=0.7)
    res[f"trsm_{m}"] = rel(bs, 0.7 * torch.linalg.solve_triangular(lref, b, upper=False))
    cc = eng.empty(m, m); eng.syrk(bs, m, n, cc)
    res[f"syrk_{m}"] = rel(cc, bs @ bs.t())
    if m >= 1024:
        res[f"potrf_ms_{m}"] = timeit(lambda: (l.copy_(spd), eng.potrf(l)))
        res[f"trinv_ms_{m}"] = timeit(lambda: eng.tri_inverse(l, linv))
# big TRSM / SYRK timing
m, n = 2048, 200000
l = (torch.tril(torch.randn(m, m, generator=g, dtype=f64)) * 0.01 + torch.eye(m, dtype=f64)).to(dev)
b = torch.randn(m, n, dtype=f64, device=dev)
ms = timeit(lambda: eng.trsm_left_lower(l, b, n, 1.0), 2); res["trsm_2048x200k_ms"] = ms; res["trsm_tflops"] = m * m * n / ms / 1e9
b = torch.randn(m, n, dtype=f64, device=dev)
cc = eng.empty(m, m)
ms = timeit(lambda: eng.syrk(b, m, n, cc), 2); res["syrk_2048x200k_ms"] = ms; res["syrk_tflops_full"] = 2 * m * m * n / ms / 1e9
res["syrk_big_rel"] = rel(cc, b @ b.t())
w = torch.randn(m, m, dtype=f64, device=dev); t = eng.empty(m, n)
ms = timeit(lambda: eng.gemm(w, b, t, m, n, m), 2); res["gemm_2048x200kx2048_ms"] = ms; res["gemm_tflops"] = 2 * m * m * n / ms / 1e9
ms = timeit(lambda: torch.mm(w, b, out=t), 2); res["cublas_same_gemm_tflops"] = 2 * m * m * n / ms / 1e9
del b, t
# preconditioner
m, n = 1024, 30001
A = (torch.randn(m, n + 1, generator=g, dtype=f64) * 0.05).to(dev)    # lda = n+1 (even)
Av = A[:, :n]
LB = torch.linalg.cholesky(Av @ Av.t() + torch.eye(m, dtype=f64, device=dev))
lbinv = eng.tri_inverse(LB.contiguous())
r = torch.randn(n, generator=g, dtype=f64).to(dev)
q = eng.empty(m); z = eng.empty(n); wv = eng.empty(m); rz = eng.empty(1)
eng.precond_project(A, m, n, r, q)
res["precond_q"] = rel(q, Av @ r)
eng.precond_finish(A, m, n, lbinv, q, r, 0.3, z, wv, rz)
wref = torch.cholesky_solve((Av @ r)[:, None], LB)[:, 0]
zref = (r - Av.t() @ wref) / 0.3
res["precond_w"] = rel(wv, wref); res["precond_z"] = rel(z, zref); res["precond_rz"] = abs(float(rz) - float(zref @ r)) / abs(float(zref @ r))
res["precond_ms_1024x30001"] = timeit(lambda: (eng.precond_project(A, m, n, r, q), eng.precond_finish(A, m, n, lbinv, q, r, 0.3, z, wv, rz)))
m, n = 2048, 400000
A = torch.randn(m, n, dtype=f64, device=dev) * 0.01
lb = torch.eye(m, dtype=f64, device=dev); r = torch.randn(n, dtype=f64, device=dev)
q = eng.empty(m); z = eng.empty(n); wv = eng.empty(m); rz = eng.empty(1)
ms1 = timeit(lambda: eng.precond_project(A, m, n, r, q)); ms2 = timeit(lambda: eng.precond_finish(A, m, n, lb, q, r, 0.3, z, wv, rz))
res["gemv_rows_GBs"] = m * n * 8 / ms1 / 1e6; res["precond_finish_GBs"] = m * n * 8 / ms2 / 1e6
del A
# vector ops
n = 100003
x = torch.randn(n, generator=g, dtype=f64).to(dev); y = torch.randn(n, generator=g, dtype=f64).to(dev)
out = eng.empty(1); eng.dot(x, y, out); res["dot"] = abs(float(out) - float(x @ y)) / abs(float(x @ y))
# knm build / backward
for kind, d in [("matern32", 3), ("rbf", 8), ("matern32", 11), ("matern32", 1)]:
    n, m = 1500, 70
    xx = torch.randn(n, d, generator=g, dtype=f64); zz = torch.randn(m, d, generator=g, dtype=f64)
    ls = torch.rand(d, generator=g, dtype=f64) + 0.7
    shift = xx.mean(0)
    xp = eng.pack(kind, xx.to(dev), ls.to(dev), shift.to(dev)); zp = eng.pack(kind, zz.to(dev), ls.to(dev), shift.to(dev))
    ld = n + (n % 2)
    out = eng.zeros(m, ld)
    eng.knm_build(kind, zp, m, xp, n, d, 1.7, out, ld)
    kref = o.kernel_dense(kind, zz, xx, ls, torch.tensor(1.7, dtype=f64))
    res[f"knm_build_{kind}_{d}"] = rel(out[:, :n].cpu(), kref)
    G = torch.randn(m, n, generator=g, dtype=f64); wt = torch.randn(m, generator=g, dtype=f64); zv = torch.randn(n, generator=g, dtype=f64)
    Gd = torch.zeros(m, ld, dtype=f64); Gd[:, :n] = G
    o_ls = eng.zeros(d); o_var = eng.zeros(1); o_z = eng.zeros(m, d)
    eng.knm_backward(kind, zp, m, xp, n, d, 1.7, ls.to(dev), Gd.to(dev), ld, wt.to(dev), zv.to(dev), o_ls, o_var, o_z)
    rl, rv, rzg = ref_knm_backward(kind, zz, xx, ls, 1.7, G + wt[:, None] * zv[None, :])
    res[f"knm_bwd_ls_{kind}_{d}"] = rel(o_ls.cpu(), rl); res[f"knm_bwd_var_{kind}_{d}"] = abs(float(o_var) - float(rv)) / abs(float(rv))
    res[f"knm_bwd_z_{kind}_{d}"] = rel(o_z.cpu(), rzg)
for k, v in res.items(): print(f"{k:40s} {v:.4g}")
json.dump(res, open("gpurun_out/dev_dense_check.json", "w"), indent=1)
