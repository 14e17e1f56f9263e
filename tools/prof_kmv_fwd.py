"""Short program for a metrics-only ncu pass: ONE launch of the symmetric K*v sweep at a given shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cglb_b200.engine import get_engine
kind, n, d = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
eng = get_engine(); dev = eng.device
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(n, d, generator=g, dtype=torch.float64, device=dev)
v = torch.randn(n, generator=g, dtype=torch.float64, device=dev)
ls = torch.full((d,), 0.5 * d ** 0.5, dtype=torch.float64, device=dev)
xp = eng.pack(kind, x, ls, x.mean(0))
y = eng.empty(n)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
eng.kmv_sym(kind, xp, n, d, v, 1.0, 0.01, out=y)
torch.cuda.synchronize()
e0.record(); eng.kmv_sym(kind, xp, n, d, v, 1.0, 0.01, out=y); e1.record(); torch.cuda.synchronize()
print("ok", float(y.sum()), "ms", e0.elapsed_time(e1), "Gpairs/s", n * n / e0.elapsed_time(e1) / 1e6)
