"""Developer check (GPU): sweeps vs the CPU oracle + timing.  Not part of the test-suite."""
import ctypes as C, os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cglb_b200 import _ffi
from oracle import cglb_oracle as o

lib = C.CDLL(_ffi.LIB_PATH)
for name, (res, args) in _ffi.SIGNATURES.items():
    if hasattr(lib, name):
        f = getattr(lib, name); f.restype = res; f.argtypes = args
ctx = C.c_void_p()
assert lib.cglb_create(C.byref(ctx), 0) == 0, lib.cglb_last_error()
dev = torch.device("cuda:0")
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
P = lambda t: C.c_void_p(t.data_ptr())
def ck(rc):
    assert rc == 0, lib.cglb_last_error()

def pack(kind, x, ls):
    n, d = x.shape
    dp = lib.cglb_packed_width(d); npad = lib.cglb_padded_rows(n)
    xp = torch.empty(npad, dp, dtype=torch.float64, device=dev)
    shift = x.mean(0).contiguous()
    ck(lib.cglb_pack_inputs(ctx, kind, P(x), n, d, P(ls), P(shift), P(xp), st()))
    return xp, shift

def run(kindname, n, d, seed=0, var=1.3, noise=0.1, lsval=None, check=True):
    kind = _ffi.KIND_IDS[kindname]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g, dtype=torch.float64)
    v = torch.randn(n, generator=g, dtype=torch.float64)
    u = torch.randn(n, generator=g, dtype=torch.float64)
    ls = (torch.rand(d, generator=g, dtype=torch.float64) + 0.5) * (lsval if lsval else (0.5 * d ** 0.5))
    xd, vd, ud, lsd = x.to(dev), v.to(dev), u.to(dev), ls.to(dev)
    xp, shift = pack(kind, xd, lsd)
    y = torch.empty(n, dtype=torch.float64, device=dev)
    ck(lib.cglb_kmv_sym(ctx, kind, P(xp), n, d, P(vd), P(y), var, noise, 0, 1, st()))
    torch.cuda.synchronize()
    out = {"kind": kindname, "n": n, "d": d}
    if check:
        lsr = ls.clone().requires_grad_(True); varr = torch.tensor(var, dtype=torch.float64, requires_grad=True)
        K = o.kernel_dense(kindname, x, x, lsr, varr)
        yref = (K @ v + noise * v).detach()
        out["kmv_sym_relerr"] = float((y.cpu() - yref).norm() / yref.norm())
        # partitioned
        ysum = torch.zeros_like(y)
        for part in range(3):
            yp = torch.empty_like(y)
            ck(lib.cglb_kmv_sym(ctx, kind, P(xp), n, d, P(vd), P(yp), var, noise, part, 3, st()))
            ysum += yp
        out["kmv_sym_parts_relerr"] = float((ysum.cpu() - yref).norm() / yref.norm())
        # rect: rows = first n//3 points vs all cols
        nr = max(1, n // 3)
        xr = xd[:nr].contiguous()
        dp = lib.cglb_packed_width(d)
        xpr = torch.empty(lib.cglb_padded_rows(nr), dp, dtype=torch.float64, device=dev)
        ck(lib.cglb_pack_inputs(ctx, kind, P(xr), nr, d, P(lsd), P(shift), P(xpr), st()))
        yr = torch.empty(nr, dtype=torch.float64, device=dev)
        ck(lib.cglb_kmv_rect(ctx, kind, P(xpr), nr, P(xp), n, d, P(vd), P(yr), var, st()))
        yrref = (K[:nr] @ v).detach()
        out["kmv_rect_relerr"] = float((yr.cpu() - yrref).norm() / yrref.norm())
        # backward
        gout = torch.zeros(d + 1, dtype=torch.float64, device=dev)
        ck(lib.cglb_kmv_bwd_sym(ctx, kind, P(xp), n, d, P(ud), P(vd), var, P(lsd), P(gout), 0, 1, st()))
        f = (u @ (K @ v))
        gl, gv = torch.autograd.grad(f, [lsr, varr])
        gref = torch.cat([gl.reshape(-1), gv.reshape(1)])
        out["bwd_relerr_max"] = float(((gout.cpu() - gref).abs() / (gref.abs() + 1e-300)).max())
        out["bwd_relerr_norm"] = float((gout.cpu() - gref).norm() / gref.norm())
    # timing
    def timeit(fn, reps=5):
        fn(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    gout = torch.zeros(d + 1, dtype=torch.float64, device=dev)
    reps = 3 if n > 100000 else 10
    ms = timeit(lambda: ck(lib.cglb_kmv_sym(ctx, kind, P(xp), n, d, P(vd), P(y), var, noise, 0, 1, st())), reps)
    out["kmv_sym_ms"] = ms; out["kmv_sym_Gpairs_s"] = n * n / ms / 1e6
    msb = timeit(lambda: ck(lib.cglb_kmv_bwd_sym(ctx, kind, P(xp), n, d, P(ud), P(vd), var, P(lsd), P(gout), 0, 1, st())), reps)
    out["bwd_ms"] = msb; out["bwd_Gpairs_s"] = n * n / msb / 1e6
    print(json.dumps(out), flush=True)
    return out

if __name__ == "__main__":
    res = []
    for args in [("matern32", 300, 1), ("matern32", 1000, 3), ("rbf", 777, 8), ("matern32", 2500, 11), ("rbf", 2049, 3),
                 ("matern32", 5000, 8)]:
        res.append(run(*args))
    # small-lengthscale stress for the expanded distance form
    res.append(run("matern32", 1500, 3, lsval=0.05))
    res.append(run("rbf", 1500, 3, lsval=0.05))
    for args in [("rbf", 40000, 8), ("matern32", 100000, 11), ("matern32", 200000, 3), ("matern32", 434000, 3)]:
        res.append(run(*args, check=False))
    json.dump(res, open("gpurun_out/dev_kmv_check.json", "w"), indent=1)
