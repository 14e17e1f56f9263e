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
