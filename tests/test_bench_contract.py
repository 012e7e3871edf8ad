"""bench.py prints ONE JSON line with the contract's keys (both arms)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def _run(args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "env-steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert "workload" in d["config"] and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


@pytest.mark.gpu
def test_our_arm_line(cuda_device):
    d = _run(["--steps", "4", "--warmup", "3", "--envs-per-gpu", "512"])
    assert BASE_KEYS <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] >= 3 and d["gpu_launches"] >= 4
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["scaling"] == "weak"
    assert d["e2e"]["h2d_bytes_per_step"] == 512 * 80 * 4 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] > 0 and d["e2e"]["value"] != d["value"]
    rf = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in rf
    assert 0 < rf["frac"] < 1.5
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0
