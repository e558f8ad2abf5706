"""CPU-only: the reference arm of bench.py (the one leg that runs without a GPU) prints one JSON line with the
keys the driver reads, on a bounded sample."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


@pytest.mark.parametrize("workload", ["c2", "c4"])
def test_reference_arm_json(workload):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload,
                          "--steps", "1", "--warmup", "0", "--ref-seconds", "0.3", "--sweeps", "4"],
                         capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert KEYS <= set(line), KEYS - set(line)
    assert line["impl"] == "reference" and line["metric"] == "spin-updates/sec" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["vs_baseline"] is None
    assert "workload" in line["config"]


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_tensor_roofline_denominator_follows_the_timed_region():
    """MEASURED_PEAKS.json holds a burst and a sustained cuBLAS bf16 rate; bench.py divides by the burst rate for a short
    timed region and by the sustained rate for regions of 2 s and more, and always reports both fractions."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    pk = {"tc_burst": 1617.1, "tc_sustained": 1355.9, "src": "test"}
    p_short, s_short = b.tensor_peak(pk, 0.4)
    p_long, s_long = b.tensor_peak(pk, 2.5)
    assert p_short == 1617.1 and "BURST" in s_short
    assert p_long == 1355.9 and "SUSTAINED" in s_long
