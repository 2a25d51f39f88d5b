"""GPU: the bench.py line keeps its contract (one JSON line on stdout with the keys the driver reads), at a size that
finishes in seconds."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "64", "--warmup", "3", "--no-cpu-baseline",
                        "--no-also", *extra], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    return json.loads(lines[0])


def test_default_line_has_the_contract_keys():
    d = _run()
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["steps"] == 64 and d["warmup"] == 3 and d["n_gpus"] == 1
    assert d["gpu_launches"] % 64 == 0 and d["gpu_launches"] >= 64  # the 64 steps are repeated until the region is >= 20 ms
    assert d["unit"] == "Gpix/s" and d["value"] > 1.0 and d["vs_baseline"] is None
    rf = d["roofline"]
    for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert k in rf, k
    assert rf["bound"] == "hbm" and 0.0 < rf["frac"] < 1.0 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    fr = rf["fractions"]
    assert fr["survey_bytes_at_step_time"] == rf["frac"] and 0 < fr["physical_bytes_at_kernel_duration"] < rf["frac"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


def test_separate_launch_mode_runs():
    d = _run("--separate")
    assert d["gpu_launches"] % 128 == 0 and d["value"] > 1.0
