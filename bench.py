#!/usr/bin/env python
"""Headline benchmark: one CGLB bound + gradient step (BASELINE.json `metric`) on synthetic data.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--theta init|trained]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's algorithm on the host cores (oracle port)

A "step" is what the reference's optimiser evaluates once per L-BFGS function evaluation
(cglb/backend/pytorch/optimizer.py:41-46, 95-98): set the hyper-parameters from a flat host vector, compute
loss = -LowerBoundCG(model)(data) -- common terms, warm-started preconditioned CG, bound -- and its
gradients w.r.t. all model parameters, return both to the host.  Between steps the lengthscales are
perturbed by +-1 % so that the warm-started CG has real work to do.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np
import torch

# name -> (description, kind, n, d, M)      BASELINE.json configs
WORKLOADS = {
    "snelson1d": ("snelson1d-shaped synthetic 1-D n=2000 Matern32 M=1024 fp64", "matern32", 2000, 1, 1024),
    "kin40k": ("kin40k-shaped synthetic n=40k d=8 RBF M=1024 fp64", "rbf", 40000, 8, 1024),
    "3droad": ("3droad-shaped synthetic n=434k d=3 Matern32 M=2048 fp64", "matern32", 434000, 3, 2048),
    "song": ("song-shaped synthetic n=515k d=90 Matern32 M=2048 fp64 (DMMA distance contraction)", "matern32", 515000, 90, 2048),
    "houseelectric": ("houseelectric-shaped synthetic n=2M d=11 Matern32 M=2048 fp64", "matern32", 2000000, 11, 2048),
}
DEFAULT_WORKLOAD = "houseelectric"          # the configuration BASELINE.json's metric is quoted on (n = 2M); fits one B200
THETAS = {      # reference initial values (config.py:76,105) / a trained-like operating point
    "init": dict(ls=lambda d: 1.0, variance=1.0, noise=1.0),
    "trained": dict(ls=lambda d: 0.5 * math.sqrt(d), variance=1.0, noise=0.01),
}
LS_MULTS = (1.0, 1.01, 0.99)


def algorithmic_flops_per_pair(kind: str, d: int) -> int:
    """Matern32 direct form 3d+7, RBF 3d+4 (t = 1 right-hand side; sqrt, exp = 1 FLOP each);
    expanded/DMMA form 2d+10 for the wide (d > 32) path."""
    if d > 32:
        return 2 * d + 10 if kind == "matern32" else 2 * d + 7
    return 3 * d + 7 if kind == "matern32" else 3 * d + 4


def synthetic(n, d, M, seed=0):
    """X ~ N(0, I), smooth f + 0.1 noise, z-scored y, Z = M random rows (oracle.synthetic_problem's recipe,
    restated here so that the product arm does not import the oracle)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, d, generator=g, dtype=torch.float64)
    g1 = torch.Generator().manual_seed(seed + 1)
    w = torch.randn(d, generator=g1, dtype=torch.float64)
    w2 = torch.randn(d, generator=g1, dtype=torch.float64)
    f = torch.sin(2.0 * (x @ w) / math.sqrt(d)) + 0.5 * torch.cos((x @ w2) / math.sqrt(d))
    g2 = torch.Generator().manual_seed(seed + 2)
    y = f + 0.1 * torch.randn(n, generator=g2, dtype=torch.float64)
    y = (y - y.mean()) / y.std()
    g3 = torch.Generator().manual_seed(seed + 3)
    perm = torch.randperm(n, generator=g3)
    return x, y, x[perm[:M]].clone()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def load_json(path):
    try:
        with open(path) as f:
            return json.load(f)
    except Exception:
        return None


# ================================================================================================
# reference arm: the reference's algorithm (oracle port) on the host cores, bounded sample
# ================================================================================================
def run_reference(args, rank):
    if rank != 0:
        return
    from oracle import cglb_oracle as o
    desc, kind, n, d, M = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    th = THETAS[args.theta]
    threads = torch.get_num_threads()
    ls = torch.full((1, d), th["ls"](d), dtype=torch.float64)
    var = torch.tensor(th["variance"], dtype=torch.float64)
    x, y, z = synthetic(n, d, M)
    v = torch.randn(n, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    counts = load_json(os.path.join(ROOT, "profiles", "bench_counts.json")) or {}
    key = f"{args.workload}:{args.theta}"
    sweeps = counts.get(key, {}).get("reference_sweeps_per_step", 8.0)
    # bounded sample: `rows` rows of the row-blocked K v (what the reference's CPU backend does for every
    # matvec) and `cols` columns of the M x n Nystrom algebra, timed and extrapolated to the full step.
    rows = max(8, min(n, int(2.0e8 // max(n, 1)), int(6.0e9 // (8.0 * n * d))))   # ~2e8 kernel pairs per timed step, <= 6 GB of differences
    cols = max(64, min(n, int(2.0e7 // M)))

    def step():
        t0 = time.perf_counter()
        o.blocked_matvec_rows(kind, x, v, ls, var, 0, rows)
        t_mv = time.perf_counter() - t0
        t0 = time.perf_counter()
        kuf = o.kernel_dense(kind, z, x[:cols], ls, var)
        kuu = o.kernel_dense(kind, z, z, ls, var) + 1e-6 * torch.eye(M, dtype=torch.float64)
        L = torch.linalg.cholesky(kuu)
        A = torch.linalg.solve_triangular(L, kuf, upper=False)
        AAt = A @ A.T
        t_nm = time.perf_counter() - t0
        per_matvec = t_mv * (n / rows)
        per_setup = t_nm * (n / cols)
        # forward (k+2 matvecs) + autograd backward of cov@v (~2 more n^2 d sweeps) + the
        # Nystrom terms forward and (x~2) backward
        return sweeps * per_matvec + 3.0 * per_setup, per_matvec

    for _ in range(args.warmup):
        step()
    vals, mvs = [], []
    for _ in range(args.steps):
        s, mv = step()
        vals.append(s)
        mvs.append(mv)
    val = float(np.mean(vals))
    sample = (f"{rows} rows x {n} cols of the row-blocked K v and {cols} columns of the M x n Nystrom algebra per step, "
              f"extrapolated to n^2 pairs x {sweeps:g} sweeps/step (forward k+2 + autograd backward) + 3 x Nystrom terms")
    line = {"impl": "reference", "metric": "cglb_bound_grad_step_time", "value": val, "unit": "s/step", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": val * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc if not args.n else f"{desc} (n overridden to {n})", "theta": args.theta,
                       "extrapolated": True, "per_matvec_s_extrapolated": float(np.mean(mvs))},
            "cpu_baseline": {"value": val, "unit": "s/step", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "s/step", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ================================================================================================
# own arm
# ================================================================================================
def run_b200(args, rank, world, local_rank):
    import torch.distributed as dist
    import cglb_b200 as cb
    from cglb_b200.engine import get_engine

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    shard = cb.Shard(rank, world, None)
    eng = get_engine(dev)

    desc, kind, n, d, M = WORKLOADS[args.workload]
    if args.n:
        n = args.n
    th = THETAS[args.theta]
    x_h, y_h, z_h = synthetic(n, d, M)
    fdt = torch.float32 if args.float_type == "fp32" else torch.float64      # exploration only: BASELINE.json's metric is fp64
    x_h, y_h, z_h = x_h.to(fdt), y_h.to(fdt), z_h.to(fdt)
    x_pin, y_pin = x_h.pin_memory(), y_h.pin_memory()

    # model exactly as interface.py:263-323 builds it (Z = M random rows: stand-in for ConditionalVariance,
    # which runs on the host outside the timed step)
    lik = cb.GaussianLikelihood(noise_constraint=cb.GreaterThan(1e-6)).to(fdt)
    lik.noise = th["noise"]
    base = (cb.MaternKernel(nu=1.5, ard_num_dims=d) if kind == "matern32" else cb.RBFKernel(ard_num_dims=d)).to(fdt)
    base.lengthscale = torch.full((d,), th["ls"](d), dtype=fdt)
    scale = cb.ScaleKernel(base).to(fdt)
    scale.outputscale = th["variance"]
    ipk = cb.InducingPointKernel(scale, z_h, likelihood=lik)
    x_dev = torch.empty(n, d, dtype=fdt, device=dev)
    y_dev = torch.empty(n, dtype=fdt, device=dev)
    x_dev.copy_(x_pin, non_blocking=True)
    y_dev.copy_(y_pin, non_blocking=True)
    model = cb.CGLB((x_dev, y_dev), lik, ipk).to(fdt).to(dev)
    data = (x_dev, y_dev)
    lower_bound = cb.LowerBoundCG(model, shard=shard)
    params = list(model.parameters())
    closure = lambda: -lower_bound(data)
    eval_func = cb.Scipy.eval_func(closure, params)             # the reference's bound+gradient call (a10)
    x0 = cb.Scipy.to_numpy(cb.Scipy.pack(params)).astype(np.float64)
    ls_slice = slice(x0.size - d, x0.size)                      # raw lengthscale is the last parameter
    raw_ls0 = x0[ls_slice].copy()
    base_ls = th["ls"](d)

    def theta_vector(step_idx):
        mult = LS_MULTS[step_idx % len(LS_MULTS)]
        xv = x0.copy()
        lsv = base_ls * mult
        xv[ls_slice] = lsv + np.log(-np.expm1(-lsv))            # inverse softplus
        return xv

    h2d_bytes = x_pin.numel() * x_pin.element_size() + y_pin.numel() * y_pin.element_size() + x0.size * 8
    d2h_bytes = (x0.size + 1) * 8
    stats = []

    def one_step(step_idx, timed):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)
        ev[0].record()
        x_dev.copy_(x_pin, non_blocking=True)                    # host -> device copy of the step's inputs (e2e only)
        y_dev.copy_(y_pin, non_blocking=True)
        ev[1].record()                                           # inputs resident: device-timed region starts
        loss, grad = eval_func(theta_vector(step_idx))           # params H2D, bound + grads, loss/grads D2H
        ev[2].record()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        t_dev, t_e2e = ev[1].elapsed_time(ev[2]), ev[0].elapsed_time(ev[2])
        if timed:
            out = lower_bound.last_output
            stats.append(dict(ms=t_dev, ms_e2e=t_e2e, cg=int(out.cg_stats.steps), matvecs=out.matvecs, loss=float(loss)))
        return t_dev, t_e2e

    def note(msg):
        if rank == 0:
            print(f"[bench] {msg}", file=sys.stderr, flush=True)

    note(f"{desc}, theta={args.theta}, {world} GPU(s): {args.warmup} warm-up + {args.steps} timed steps")
    for i in range(args.warmup):
        td, _ = one_step(i, False)
        note(f"warm-up step {i + 1}/{args.warmup}: {td / 1e3:.2f} s")
    eng.enable_timing(True)
    launches0 = eng.launch_count
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    times = []
    for i in range(args.steps):
        times.append(one_step(args.warmup + i, True))
        note(f"timed step {i + 1}/{args.steps}: {times[-1][0] / 1e3:.2f} s, CG iterations {stats[-1]['cg']}")
    clocks = sampler.stop() if sampler else None
    launches = eng.launch_count - launches0
    ksum = eng.timing_summary()
    eng.enable_timing(False)

    t_dev = torch.tensor([sum(t[0] for t in times), sum(t[1] for t in times)], dtype=torch.float64, device=dev)
    lt = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    kt = torch.tensor([ksum.get("kmv_sym", (0, 0.0))[1], ksum.get("kmv_bwd_sym", (0, 0.0))[1],
                       ksum.get("precond_project", (0, 0.0))[1] + ksum.get("precond_finish", (0, 0.0))[1],
                       ksum.get("trsm", (0, 0.0))[1] + ksum.get("syrk", (0, 0.0))[1] + ksum.get("gemm", (0, 0.0))[1]],
                      dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)             # max over ranks
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_step = float(t_dev[0]) / args.steps
    ms_e2e = float(t_dev[1]) / args.steps
    n_kmv = ksum.get("kmv_sym", (0, 0.0))[0]
    fpp = algorithmic_flops_per_pair(kind, d)
    peaks = load_json(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")) or {}
    measured = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json")) or {}
    fp64_peak = float(peaks.get("fp64_peak_tflops_used_as_denominator", 37.1))
    hbm_peak = float(measured.get("hbm_gbs", 6650.0))
    # dominant kernel: the symmetric K v sweep.  Algorithmic FLOPs per launch on this rank = (3d+7) n^2 / world.
    kmv_ms = float(kt[0]) / max(n_kmv, 1)
    achieved = fpp * float(n) * n / world / (kmv_ms * 1e-3) / 1e12 if n_kmv else None
    ncu = load_json(os.path.join(ROOT, "profiles", "kmv_ncu_summary.json")) or {}
    # DRAM bytes per launch (ncu, per rank): every rank streams the whole packed input array once, whatever its share of
    # the work items, so the single-GPU figure holds for every N
    traffic = ncu.get(f"{args.workload}", {}).get("dram_bytes_per_launch")
    variant = eng.kmv_sym_variant(d, n, world)
    diag_block = {0: 1024.0, 1: 256.0, 2: 128.0}[variant]      # rows of the diagonal blocks, evaluated as full squares
    eval_frac = 0.5 + 0.5 * min(1.0, diag_block / n)
    n_pre = ksum.get("precond_project", (0, 0.0))[0]
    pre_bytes = 2.0 * M * (n / world) * 8 + 4.0 * (n / world) * 8
    pre_ms = float(kt[2]) / max(n_pre, 1)
    line = {
        "metric": "cglb_bound_grad_step_time", "value": ms_step * 1e-3, "unit": "s/step", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64" if args.float_type == "fp64" else "f32 kernel pairs, f64 accumulation and linear algebra",
        "data": "synthetic",
        "config": {"workload": (desc if not args.n else f"{desc} (n overridden to {n})") +
                               ("" if args.float_type == "fp64" else " -- run in the fp32 mode of the API, NOT the fp64 metric of BASELINE.json"),
                   "theta": args.theta, "kernel": kind, "n": n, "d": d, "M": M, "parallelism": f"row-sharded x{world}",
                   "cg": "reference defaults (max_error=1, max_cg_iter=100, restart=40), warm start carried across steps, "
                         "lengthscales perturbed by (1, 1.01, 0.99) per step; K v, r and P r after the solve come from the final "
                         "CG state (k + 1 + floor(k/40) sweeps per step; the reference recomputes them for its autograd tape: "
                         "k + 2 + floor(k/40); CGLB_RECOMPUTE_RESIDUAL=1 selects that)",
                   "cg_steps": [s["cg"] for s in stats], "kv_sweeps_per_step": [s["matvecs"] for s in stats],
                   "loss": [s["loss"] for s in stats],
                   "l2": "inputs larger than L2 (X packed %.0f MB, A %.1f GB per rank)" % (n * (d + 2) * 8 / 1e6, M * n / world * 8 / 1e9)},
        "kv_gpairs_per_s": (float(n) * n / (kmv_ms * 1e-3) / 1e9) if n_kmv else None,
        "e2e": {"value": ms_e2e * 1e-3, "unit": "s/step", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
        "gpu_launches": int(lt.item()),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "K1 symmetric matrix-free K*v: " + eng.KMV_VARIANTS[variant],
                     "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": (achieved / fp64_peak) if achieved else None,
                     "traffic": traffic,
                     "note": "algorithmic-FLOP fraction: (3d+7) n^2 FLOP per K*v (SURVEY.md 8d) / CUDA-event time per launch; "
                             "FP64 has no tcgen05 path, the denominator is the measured DMMA.8x8x4 rate "
                             "(profiles/fp64_peaks_r01.json, 'of measured'); the symmetric sweep evaluates each unordered pair once, so `achieved` "
                             "(nominal n^2 pairs of the reference's K*v) can exceed the peak; `achieved_evaluated` counts only the pairs "
                             "actually evaluated (n^2/2 + the diagonal blocks)",
                     "achieved_evaluated": (achieved * eval_frac) if achieved else None,
                     "frac_evaluated": (achieved * eval_frac / fp64_peak) if achieved else None,
                     "fp64_pipe_utilisation_ncu": (ncu.get("fp64_pipe_active_pct_by_variant") or {}).get(str(variant), ncu.get("fp64_pipe_active_pct")),
                     "launches": n_kmv, "ms_per_launch": kmv_ms,
                     "share_of_step": float(kt[0]) / float(t_dev[0]) if float(t_dev[0]) else None},
        "roofline_other": {
            "precond_gemv_pair": {"bound": "hbm", "achieved": (pre_bytes / (pre_ms * 1e-3) / 1e9) if n_pre else None,
                                  "peak": hbm_peak, "unit": "GB/s",
                                  "frac": (pre_bytes / (pre_ms * 1e-3) / 1e9 / hbm_peak) if n_pre else None,
                                  "share_of_step": float(kt[2]) / float(t_dev[0]) if float(t_dev[0]) else None,
                                  "note": "two read-only streaming passes over A per apply; MEASURED_PEAKS.json's hbm_gbs is a "
                                          "read+write copy, which a pure read stream can exceed (frac > 1 at the largest sizes)"},
            "backward_sweep": {"share_of_step": float(kt[1]) / float(t_dev[0]) if float(t_dev[0]) else None,
                               "ms_per_launch": float(kt[1]) / max(ksum.get("kmv_bwd_sym", (1, 0))[0], 1)},
            "dense_trsm_syrk_gemm": {"share_of_step": float(kt[3]) / float(t_dev[0]) if float(t_dev[0]) else None}},
    }
    if args.float_type == "fp32" and d <= 32:
        # exploration: the sweep that ran is f32_sweep_kernel (FP32 FMA + MUFU pipes); the FP64 roofline does not apply
        line["roofline"].update({"kernel": "K1 symmetric matrix-free K*v: f32_sweep_kernel (FP32 kernel pairs, FP64 accumulation)",
                                 "frac": None, "achieved_evaluated": None, "frac_evaluated": None, "fp64_pipe_utilisation_ncu": None,
                                 "note": "fp32 mode of the API: 2 MUFU + ~17 FP32 + ~3 other instructions per pair, issue-bound; "
                                         "`achieved` is still the nominal (3d+7) n^2 figure, not comparable with the FP64 peak"})
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(kind, n, d, M, th, stats)
    print(json.dumps(line), flush=True)
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        tag = f"_n{n}" if args.n else ""
        with open(os.path.join(ROOT, "gpurun_out", f"bench_{args.workload}{tag}_{args.theta}_n{world}.json"), "w") as f:
            json.dump(line, f, indent=1)
    except Exception:
        pass
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(kind, n, d, M, th, stats):
    """The oracle (CPU port of the reference path) on a bounded sample of the same workload, extrapolated."""
    from oracle import cglb_oracle as o
    threads = torch.get_num_threads()
    x, y, z = synthetic(n, d, M)
    ls = torch.full((1, d), th["ls"](d), dtype=torch.float64)
    var = torch.tensor(th["variance"], dtype=torch.float64)
    v = torch.randn(n, 1, dtype=torch.float64, generator=torch.Generator().manual_seed(5))
    rows = max(8, min(n, int(2.0e8 // n), int(6.0e9 // (8.0 * n * d))))      # ~2e8 pairs, <= 6 GB for the [rows, n, d] differences
    o.blocked_matvec_rows(kind, x, v, ls, var, 0, min(rows, 8))            # warm-up
    t0 = time.perf_counter()
    reps = 0
    while time.perf_counter() - t0 < 12.0 and reps < 50:
        o.blocked_matvec_rows(kind, x, v, ls, var, (reps * rows) % max(1, n - rows), rows)
        reps += 1
    per_matvec = (time.perf_counter() - t0) / reps * (n / rows)
    # the reference's own count: k + 2 + floor(k/40) forward sweeps (conjugate_gradient.py:57,66,72; models.py:280)
    # + the autograd backward of cov@v (~2 more n^2 d sweeps) -- not this implementation's count, which reuses the
    # CG residual instead of the sweep of models.py:280
    sweeps = float(np.mean([s["cg"] + 2 + s["cg"] // 40 for s in stats])) + 2.0 if stats else 8.0
    return {"value": sweeps * per_matvec, "unit": "s/step", "cores": threads, "kind": "port",
            "sample": f"{reps} x {rows} rows x {n} cols of the oracle's row-blocked K v (torch fp64, {threads} threads), "
                      f"extrapolated to n^2 pairs x {sweeps:g} sweeps/step; Nystrom terms not included",
            "per_matvec_s_extrapolated": per_matvec, "gpairs_per_s": float(n) * n / per_matvec / 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1)      # n = 2M on one GPU is ~90 s per step (15-17 CG iterations)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--theta", default="init", choices=sorted(THETAS))
    ap.add_argument("--n", "--n-rows", dest="n", type=int, default=0,
                    help="override n (exploration only; the line says so); spell it --n-rows under torch.distributed.run, whose "
                         "own parser claims the abbreviation --n")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--float-type", default="fp64", choices=["fp64", "fp32"],
                    help="fp32: the API's fp32 switch (FP32 kernel pairs); exploration only, the headline metric is fp64")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; cglb_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
