"""bench.py prints one JSON line with the keys the driver depends on."""
import json
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "gpu_launches"}


def run(*args):
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True,
                         cwd=str(ROOT), timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run("--impl", "reference", "--steps", "1", "--warmup", "0", "--no-reference-python")
    assert COMMON <= set(d) and d["impl"] == "reference"
    assert d["unit"] == "molecules/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["value"] > 1e5 and d["dtype"] == "f64" and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # both arms print the same `config` dict (the driver compares them)
    sys.path.insert(0, str(ROOT))
    import bench

    assert d["config"] == bench.shared_config() and d["sample_molecules_per_step"] > 0


def test_reference_python_leg():
    """The unmodified reference (baseline/_ref) timed in a subprocess, when it has been installed."""
    sys.path.insert(0, str(ROOT))
    import bench

    r = bench.reference_python(2)
    if "unavailable" in r:
        assert not (ROOT / "baseline" / "_ref" / "trajectories").exists()
        pytest.skip(r["unavailable"])
    assert r["all_cores"]["cores"] == 2 and r["all_cores"]["n"] == 16000 and 1e2 < r["all_cores"]["value"] < 1e6
    assert r["one_core"]["cores"] == 1 and r["one_core"]["n"] == 20000 and 1e2 < r["one_core"]["value"] < 1e5


@pytest.mark.gpu
def test_gpu_arm_line():
    d = run("--steps", "3", "--warmup", "3", "--no-cpu", "--molecules", "2e6")
    assert COMMON <= set(d) and "impl" not in d
    assert d["n_gpus"] == 1 and d["steps"] == 3 and d["warmup"] == 3
    assert d["gpu_launches"] == 3 * 6          # per step: walk + 4 lens segments of 150 RK steps + tail
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and d["dtype"] == "f64"
    for key in ("e2e", "e2e_philox", "e2e_api"):
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d[key]) and d[key]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 48 * 2_000_000 and d["e2e"]["counters_match_device_run"] is True
    assert d["roofline"]["regime"].startswith("overlapped") and d["roofline"]["avg_launch_ms"] <= d["ms_per_step"] * 1.0001
    assert d["roofline_one_stream"]["regime"].startswith("one stream") and 0 < d["roofline"]["fp64_pipe_busy"] < 1
    assert d["e2e"]["roofline"]["bound"] == "pcie" and 0 < d["e2e"]["roofline"]["frac"] <= 1.05
    for key in ("roofline", "roofline_walk"):
        r = d[key]
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
        assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and 0 < r["frac"] < 1.2
    assert d["roofline"]["bound"] == "fp64" and d["roofline_walk"]["bound"] == "hbm"
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert sum(d["counters"].values()) == 2_000_000
    assert d["contracted_math"]["value"] > d["value_one_stream"]
