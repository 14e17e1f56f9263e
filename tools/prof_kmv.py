"""Short program for ncu: a few launches of the K*v sweep (and the backward sweep) at a given shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cglb_b200.engine import get_engine
kind, n, d = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
eng = get_engine(); dev = eng.device
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(n, d, generator=g, dtype=torch.float64, device=dev)
v = torch.randn(n, generator=g, dtype=torch.float64, device=dev)
u = torch.randn(n, generator=g, dtype=torch.float64, device=dev)
ls = torch.full((d,), 0.5 * d ** 0.5, dtype=torch.float64, device=dev)
xp = eng.pack(kind, x, ls, x.mean(0))
y = eng.empty(n); out = eng.zeros(d + 1)
for _ in range(reps):
    eng.kmv_sym(kind, xp, n, d, v, 1.0, 0.01, out=y)
for _ in range(max(1, reps - 1)):
    eng.kmv_bwd_sym(kind, xp, n, d, u, v, 1.0, ls, out)
torch.cuda.synchronize()
print("ok", float(y.sum()))
