"""The bench.py contract the driver depends on: one JSON line on stdout with the agreed keys, for both arms."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
             "config", "e2e", "gpu_launches"}


def _run(args, timeout):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, f"bench.py must print exactly one line on stdout, found {len(lines)}"
    return json.loads(lines[0])


def test_reference_arm_line():
    """--impl reference: the CPU restatement of the reference's unfused chain on the host cores (no GPU needed)"""
    j = _run(["--impl", "reference", "--steps", "1", "--warmup", "1"], 600)
    assert BASE_KEYS <= set(j) and j["impl"] == "reference"
    assert j["metric"].startswith("2160p50 v210 4-layer composite") and j["unit"] == "frames/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["e2e"] == {"value": j["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "sample" in cb
    assert "workload" in j["config"] and "model" not in j["config"]


@pytest.mark.gpu
def test_our_arm_line():
    j = _run(["--steps", "1", "--warmup", "3", "--frames-per-step", "6"], 900)
    assert BASE_KEYS | {"roofline", "cpu_baseline", "clocks"} <= set(j) and "impl" not in j
    assert j["n_gpus"] == 1 and j["warmup"] >= 3 and j["dtype"] == "f32" and j["data"] == "synthetic" and j["scaling"] == "weak" and j["vs_baseline"] is None
    r = j["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["peak"] > 1000
    assert r["bytes_moved_per_launch"] <= r["algorithmic_bytes_per_launch"] == 132710400
    e = j["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < j["value"]
    assert j["gpu_launches"] == 6 and j["config"]["launches_per_frame"] == 1 and "workload" in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] > 0
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(j["clocks"])
