"""Short program for ncu: a few launches of the fp32-pair K*v sweep."""
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
xpf = eng.pack_f32(kind, x, ls, x.mean(0)); y = eng.empty(n)
for _ in range(3):
    eng.kmv_sym_f32(kind, xpf, n, d, v, 1.0, 0.01, out=y)
torch.cuda.synchronize(); print("ok", float(y.sum()))
