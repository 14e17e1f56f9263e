import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cglb_b200.engine import get_engine
eng = get_engine(); dev = eng.device
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for kind, n, d in [("matern32", 200000, 11), ("matern32", 300000, 3)]:
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, d, generator=g, dtype=torch.float64, device=dev)
    v = torch.randn(n, generator=g, dtype=torch.float64, device=dev)
    ls = torch.full((d,), 0.5 * d ** 0.5, dtype=torch.float64, device=dev)
    xpf = eng.pack_f32(kind, x, ls, x.mean(0)); y = eng.empty(n)
    for var in sys.argv[1:]:
        os.environ["CGLB_F32_VARIANT"] = var
        ms = timeit(lambda: eng.kmv_sym_f32(kind, xpf, n, d, v, 1.0, 0.01, out=y))
        print(f"f32 {kind} n={n} d={d} variant {var}: {ms:8.3f} ms {n*n/ms/1e6:8.1f} Gpairs/s", flush=True)
