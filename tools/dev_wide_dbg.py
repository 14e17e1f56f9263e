import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cglb_b200.engine import get_engine
from oracle import cglb_oracle as o
eng = get_engine(); dev = eng.device; f64 = torch.float64
for kind, n, d in [("matern32", 1000, 64), ("matern32", 1100, 40), ("matern32", 1100, 64), ("matern32", 2100, 40), ("matern32", 1024, 90), ("matern32", 1030, 90)]:
    g = torch.Generator().manual_seed(n + 1)
    x = torch.randn(n, d, generator=g, dtype=f64); v = torch.randn(n, generator=g, dtype=f64); u = torch.randn(n, generator=g, dtype=f64)
    ls = (torch.rand(d, generator=g, dtype=f64) + 0.5) * 0.5 * d ** 0.5
    xp = eng.pack(kind, x.to(dev), ls.to(dev), x.mean(0).to(dev))
    out = eng.zeros(d + 1)
    eng.kmv_bwd_sym(kind, xp, n, d, u.to(dev), v.to(dev), 1.3, ls.to(dev), out)
    lsr, varr = ls.clone().requires_grad_(True), torch.tensor(1.3, dtype=f64, requires_grad=True)
    f = u @ (o.kernel_dense(kind, x, x, lsr, varr, block=128) @ v)
    gl, gv = torch.autograd.grad(f, [lsr, varr])
    oc = out.cpu()
    print(f"n={n} d={d}: ls relerr {float((oc[:d]-gl.reshape(-1)).norm()/gl.norm()):.2e} var relerr {abs(float(oc[d])-float(gv))/abs(float(gv)):.2e}", flush=True)
