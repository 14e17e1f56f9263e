"""Developer timing (GPU): the dense FP64 kernels at the configs' shapes, TFLOP/s of the work they do (CUDA events), with the
torch (cuBLAS / cuSOLVER) figure beside each where one exists."""
import os, sys, json
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
res = {}
g = torch.Generator(device=dev).manual_seed(0)
for m, n in ((2048, 200000), (1024, 40000), (2048, 54250)):
    ld = (n + 15) // 16 * 16
    A = torch.randn(m, ld, generator=g, dtype=torch.float64, device=dev) * (1.0 / m ** 0.5)
    H = torch.randn(m, m, generator=g, dtype=torch.float64, device=dev)
    T = eng.empty(m, ld)
    S = torch.randn(m, m, generator=g, dtype=torch.float64, device=dev)
    spd = S @ S.t() / m + torch.eye(m, dtype=torch.float64, device=dev)
    L = eng.potrf(spd.clone())
    C = eng.empty(m, m)
    ms = timeit(lambda: eng.gemm(H, A, T, m, n, m)); res[f"gemm_{m}x{n}x{m}"] = 2.0 * m * m * n / ms / 1e9
    ms_t = timeit(lambda: torch.matmul(H, A[:, :n])); res[f"cublas_gemm_{m}x{n}x{m}"] = 2.0 * m * m * n / ms_t / 1e9
    B = A.clone()
    ms = timeit(lambda: eng.trsm_left_lower(L, B, n, alpha=1.0)); res[f"trsm_{m}x{n}"] = 1.0 * m * m * n / ms / 1e9
    ms_t = timeit(lambda: torch.linalg.solve_triangular(L, A[:, :n], upper=False)); res[f"cublas_trsm_{m}x{n}"] = 1.0 * m * m * n / ms_t / 1e9
    ms = timeit(lambda: eng.syrk(A, m, n, C)); res[f"syrk_{m}x{n}"] = 1.0 * m * m * n / ms / 1e9
    ms_t = timeit(lambda: torch.matmul(A[:, :n], A[:, :n].t())); res[f"cublas_gemm_as_syrk_{m}x{n}"] = 1.0 * m * m * n / ms_t / 1e9
    ms = timeit(lambda: eng.potrf(spd.clone())); res[f"potrf_{m}_ms"] = ms
    ms_t = timeit(lambda: torch.linalg.cholesky(spd)); res[f"cusolver_potrf_{m}_ms"] = ms_t
    ms = timeit(lambda: eng.tri_inverse(L)); res[f"tri_inverse_{m}_ms"] = ms
    del A, T, B
for k, v in res.items():
    print(f"{k:36s} {v:10.3f} {'ms' if k.endswith('_ms') else 'TFLOP/s (of the real work)'}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/dev_dense_time.json", "w"), indent=1)
