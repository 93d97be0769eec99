"""bench.py contract: exactly one JSON line on stdout with the keys the driver reads (CPU: the reference arm on a small
problem; GPU: our arm on a small problem)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "e2e",
             "gpu_launches", "cpu_baseline"}


def _run(args, timeout=600):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--log2n", "13", "--steps", "2", "--warmup", "1"])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "fp64_pair_interactions_per_s" and d["unit"] == "pairs/s" and d["higher_is_better"] is True
    assert d["value"] > 1e7 and d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and "sample" in d["cpu_baseline"]
    assert d["config"]["workload"].startswith("Stokes3D_FxU") and d["vs_baseline"] is None and d["gpu_launches"] == 0


@pytest.mark.gpu
def test_our_arm_line_small():
    d = _run(["--log2n", "14", "--steps", "2", "--warmup", "3", "--cpu-targets", "2048"])
    assert BASE_KEYS | {"roofline", "clocks", "parity"} <= set(d)
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 3 and d["dtype"] == "f64" and d["data"] == "synthetic"
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and 0.05 < r["frac"] < 1.0 and r["unit"] == "TFLOP/s"
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] > 0
    assert d["gpu_launches"] >= 2 * d["steps"]
    assert d["parity"]["max_abs_err_over_vinf"] <= 1e-12
    km = r["kernel_matrix_view"]  # the path's HBM-bound kernel in the contract's hbm form
    assert km["bound"] == "hbm" and km["unit"] == "GB/s" and {"achieved", "peak", "frac", "traffic"} <= set(km)
    assert km["algorithmic_bytes_per_launch"] == 8192 * 3 * 8192 * 3 * 8 and 0.5 < km["frac"] < 1.3
    assert d["cpu_baseline"]["value"] > 0 and d["value"] > d["cpu_baseline"]["value"]


@pytest.mark.gpu
def test_our_arm_surface_points_small():
    """--points torus: surface nodes, targets == sources (every target has one r2 == 0 pair that must contribute 0)."""
    d = _run(["--log2n", "14", "--points", "torus", "--steps", "2", "--warmup", "3", "--cpu-targets", "2048", "--no-e2e"])
    assert "torus surface nodes" in d["config"]["workload"] and "targets == sources" in d["config"]["workload"]
    assert d["parity"]["max_abs_err_over_vinf"] <= 1e-12 and d["value"] > 0


def test_reference_arm_uses_all_host_threads_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must still use every host thread it may run on."""
    env = dict(os.environ, OMP_NUM_THREADS="1", TORCHELASTIC_RUN_ID="test", RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--log2n", "13", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    d = json.loads(r.stdout.strip())
    want = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == want and d["n_gpus"] == 2
    # the other ranks exit 0 without work and print nothing
    env["RANK"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--log2n", "13"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
