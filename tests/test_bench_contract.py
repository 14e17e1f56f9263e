"""CPU: the reference arm of bench.py (the oracle port timed on the host cores) prints ONE JSON line with the
keys the bench contract names; the own arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=600)


def test_reference_arm_line():
    r = _run("--impl", "reference", "--workload", "snelson1d", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "cglb_bound_grad_step_time" and j["unit"] == "s/step"
    assert j["higher_is_better"] is False and j["dtype"] == "f64" and j["data"] == "synthetic"
    assert j["steps"] == 1 and j["warmup"] == 1 and j["value"] > 0 and abs(j["ms_per_step"] - 1e3 * j["value"]) < 1e-6 * j["ms_per_step"]
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    assert j["e2e"] == {"value": j["value"], "unit": "s/step", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "snelson1d" in j["config"]["workload"]


def test_own_arm_has_no_cpu_fallback():
    if torch.cuda.is_available():
        return
    r = _run("--workload", "snelson1d", "--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_budget_projection():
    """The wall-clock budget that keeps `--steps 20 --warmup 5` inside the driver's limits (round 1 was killed)."""
    sys.path.insert(0, ROOT)
    import time
    import bench
    b = bench.Budget(seconds=(time.time() - bench.T_PROCESS_START) + 400.0, reserve=35.0)
    assert 360.0 < b.left() <= 365.0
    assert b.fits(80.0, 3)              # 3 x 80 s x 1.25 = 300 s
    assert not b.fits(80.0, 4)          # 400 s
    assert not b.fits(300.0)


def test_committed_bench_lines_respect_the_roofline_convention():
    """roofline.frac counts the pairs a launch EVALUATES (the nominal n^2 figure read as 104-150 % of peak)."""
    import glob
    lines = sorted(glob.glob(os.path.join(ROOT, "profiles", "bench_*_r02*.json")))
    for path in lines:
        with open(path) as f:
            j = json.load(f)
        if j.get("impl") == "reference":
            continue
        frac = j["roofline"]["frac"]
        assert frac is None or 0.0 < frac <= 1.0, (path, frac)
        assert "achieved_nominal_n2" in j["roofline"]
        assert j["steps"] >= 1 and j["config"]["steps_requested"] >= j["steps"]


def test_budget_estimate_ignores_the_cold_step_and_trusts_a_full_cycle():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.Budget.estimate([185.0]) == (185.0, bench.Budget.SAFETY)                 # only the cold step so far
    assert bench.Budget.estimate([185.0, 75.0]) == (75.0, bench.Budget.SAFETY)            # CG from v = 0 is left out
    est, safety = bench.Budget.estimate([185.0, 75.0, 83.0, 92.0])
    assert est == 92.0 and safety == 1.05                                                  # a whole lengthscale cycle seen
    est, safety = bench.Budget.estimate([185.0, 75.0, 83.0, 92.0, 75.0])
    assert est == 92.0 and safety == 1.05
