import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cglb_b200.engine import get_engine
eng = get_engine(); dev = eng.device
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x = torch.randn(m, m + 50, dtype=torch.float64, device=dev)
spd = x @ x.t() / m + torch.eye(m, dtype=torch.float64, device=dev)
for _ in range(2):
    l = spd.clone(); eng.potrf(l); linv = eng.tri_inverse(l)
torch.cuda.synchronize(); print("ok")
