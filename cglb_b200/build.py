"""Build libcglb_b200.so (sm_100a only) in-tree with nvcc.

    python -m cglb_b200.build            # all kernels, d = 1..32
    CGLB_KMV_DIMS=3,8,11 python -m cglb_b200.build   # quick developer build

Objects go to cglb_b200/csrc/build/, the shared library to cglb_b200/lib/libcglb_b200.so (git-ignored,
shipped to the GPU box by gpurun).  nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libcglb_b200.so")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr"]
PLAIN_UNITS = ["context.cu", "kmv_api.cu", "dense.cu", "vecops.cu", "widek.cu", "dsweep.cu"]
ALL_DIMS = list(range(1, 33))


def _dims():
    env = os.environ.get("CGLB_KMV_DIMS")
    if not env:
        return ALL_DIMS
    return sorted({int(t) for t in env.split(",") if t.strip()})


def _stamp(paths, extra=""):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _headers():
    hs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(HERE), "include", "cglb_b200.h"))
    return hs


def _compile(src, obj, defs):
    cmd = [NVCC, *ARCH, *COMMON, *defs, "-c", src, "-o", obj]
    stamp_file = obj + ".stamp"
    stamp = _stamp([src] + _headers(), " ".join(cmd))
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, "cached"
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return obj, "built"


def build(verbose: bool = True) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    dims = _dims()
    dims_list = " ".join(f"X({d})" for d in dims)
    jobs = []
    for u in PLAIN_UNITS:
        src = os.path.join(CSRC, u)
        if not os.path.exists(src):
            continue
        defs = [f"-DCGLB_KMV_DIMS_LIST={dims_list}"] if u == "kmv_api.cu" else []
        if u == "dsweep.cu":
            defs = os.environ.get("CGLB_EXTRA_DEFS", "").split()
        jobs.append((src, os.path.join(OBJ, u.replace(".cu", ".o")), defs))
    extra = os.environ.get("CGLB_EXTRA_DEFS", "").split()      # e.g. -DCGLB_KMV_EXPERIMENT (developer builds)
    for d in dims:
        jobs.append((os.path.join(CSRC, "kmv_inst.cu"), os.path.join(OBJ, f"kmv_d{d}.o"), [f"-DCGLB_KMV_D={d}", *extra]))
    objs = []
    workers = max(1, min(len(jobs), os.cpu_count() or 4))
    with cf.ThreadPoolExecutor(workers) as ex:
        for obj, how in ex.map(lambda j: _compile(*j), jobs):
            objs.append(obj)
            if verbose:
                print(f"[cglb_b200.build] {how:6s} {os.path.relpath(obj, HERE)}", flush=True)
    link = [NVCC, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"[cglb_b200.build] linked {LIB} (dims {dims})", flush=True)
    return LIB


if __name__ == "__main__":
    build()
    sys.exit(0)
