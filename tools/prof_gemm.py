"""Short program for ncu: the big NN GEMM of the backward pass (T = H A, (M x M)(M x n)) and the SYRK A A^T."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cglb_b200.engine import get_engine
m, n = int(sys.argv[1]), int(sys.argv[2])
eng = get_engine(); dev = eng.device
g = torch.Generator(device=dev).manual_seed(0)
ld = (n + 15) // 16 * 16
A = torch.randn(m, ld, generator=g, dtype=torch.float64, device=dev) * (1.0 / m ** 0.5)
H = torch.randn(m, m, generator=g, dtype=torch.float64, device=dev)
T = eng.empty(m, ld)
C = eng.empty(m, m)
for _ in range(2):
    eng.gemm(H, A, T, m, n, m)
    eng.syrk(A, m, n, C)
torch.cuda.synchronize()
print("ok", float(T[:, :n].sum()), float(C.trace()))
