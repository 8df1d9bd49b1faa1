"""bench.py's output contract (the driver parses these lines).  The reference arm runs on the CPU, so its line is
checked here; the device arm's keys are checked on the GPU box."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
          "dtype", "data", "config", "cpu_baseline", "e2e"}


def run_bench(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, "bench.py must print exactly ONE JSON line"
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--steps", "1", "--warmup", "0", "--workload", "C2")
    assert COMMON <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "svo_build_leaf_voxels_per_s" and d["unit"] == "leaf voxels/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_device_arm_line():
    d = run_bench("--steps", "3", "--warmup", "3", "--workload", "C2", "--no-cpu-baseline")
    assert (COMMON - {"cpu_baseline"}) <= set(d)
    assert d.get("impl") != "reference"
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3 and d["gpu_launches"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and 0 < rf["frac"] < 1 and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] <= d["value"] * 1.001
    assert "sm_mhz" in d["clocks"] and "reasons" in d["clocks"]


@pytest.mark.gpu
def test_device_arm_parity_check():
    """The bench compares the oracle tree it computes for cpu_baseline with a CUDA build of the same voxel window."""
    d = run_bench("--steps", "3", "--warmup", "3", "--workload", "C2")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    pc = d["parity_check"]
    assert pc["result"] == "ok" and pc["fragments"] > 100_000 and pc["leaves"] > 0
    assert d["stitch_check"] is None  # single GPU: nothing is stitched
