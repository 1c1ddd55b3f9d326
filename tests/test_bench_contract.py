"""bench.py's output contract, checked without a GPU: the reference arm (the C/OpenMP oracle port on host cores) runs here on a
tiny slice and must print ONE JSON line with the contract's keys; the GPU arm's lines committed under profiles/ are checked for
the same keys plus `roofline`, `cpu_baseline`, `e2e`, `gpu_launches` and `clocks` (they are what the judge reads)."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config")


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--samples", "40", "--regions", "2000", "--cpu-samples", "40", "--cpu-chroms", "3"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference"
    for k in BASE_KEYS:
        assert k in d, k
    assert d["unit"] == "input regions/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["config"]["workload"].startswith("C5")
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # both arms word the workload identically (the driver compares `config`), and the probe for a real GMQL-Spark is recorded
    import bench
    import argparse
    a = argparse.Namespace(samples=40, regions=2000)
    assert d["config"] == bench.c5_config(a, 1) and d["warmup"] == 0
    assert d["gmql_spark_probe"]["conclusion"]


def test_reference_arm_uses_every_core_under_torchrun_env():
    """torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm must not drop to one core because of it."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--samples", "20", "--regions", "1000", "--cpu-samples", "20", "--cpu-chroms", "2"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads([ln for ln in p.stdout.splitlines() if ln.strip()][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(ROOT, "profiles", "r01[kl]_bench_c5_*gpu.json"))))
def test_committed_gpu_lines_follow_the_contract(path):
    d = json.load(open(path))
    for k in BASE_KEYS:
        assert k in d, k
    n = d["n_gpus"]
    assert d["unit"] == "input regions/s" and d["scaling"] == "strong" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["steps"] >= 1 and d["warmup"] >= 3
    assert abs(d["value"] - d["config"]["input_regions"] / (d["ms_per_step"] * 1e-3)) / d["value"] < 1e-6     # whole-job aggregate
    assert d["config"]["workload"].startswith("C5") and "l2" in d["config"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and 0 < e["value"] < d["value"]
    assert d["gpu_launches"] > 0
    assert d["clocks"]["sm_mhz"] > 0 and not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert (r["traffic"] is not None) == (n == 1)
    if n == 1 and "cpu_baseline" in d:
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    # the same job at every N: identical result tables
    assert d["config"]["cover_regions"] == 10197 and d["config"]["map_count_checksum"] == 100009343
