"""Short program for ncu: the preconditioner GEMV pair (K3/K4) and the dense FP64 kernels (K5/K6) at a given (M, n)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cglb_b200.engine import get_engine
m, n = int(sys.argv[1]), int(sys.argv[2])
eng = get_engine(); dev = eng.device
g = torch.Generator(device=dev).manual_seed(0)
ld = (n + 15) // 16 * 16
A = torch.randn(m, ld, generator=g, dtype=torch.float64, device=dev) * (1.0 / m ** 0.5)
r = torch.randn(n, generator=g, dtype=torch.float64, device=dev)
S = torch.randn(m, m, generator=g, dtype=torch.float64, device=dev)
spd = S @ S.t() / m + torch.eye(m, dtype=torch.float64, device=dev)
L = eng.potrf(spd.clone())
LBinv = eng.tri_inverse(L)
q, w, z, rz = eng.empty(m), eng.empty(m), eng.empty(n + 1), None
for _ in range(3):
    eng.precond_project(A, m, n, r, q)
    eng.precond_finish(A, m, n, LBinv, q, r, 0.5, z[:n], w, z[n:])
B = A.clone()
eng.trsm_left_lower(L, B, n, alpha=1.0)
C = eng.empty(m, m)
eng.syrk(B, m, n, C)
H = torch.randn(m, m, generator=g, dtype=torch.float64, device=dev)
T = eng.empty(m, ld)
eng.gemm(H, A, T, m, n, m)
torch.cuda.synchronize()
print("ok", float(q.sum()), float(C.trace()), float(T[:, :n].sum()))
