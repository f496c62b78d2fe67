"""The bench.py JSON contract on CPU: the reference arm (the oracle port on the host cores) prints ONE line with the
keys the driver reads, on the same metric / unit / config as the GPU arm."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "line-gridpoint evals/s" and d["unit"] == "pairs/s"
    assert d["higher_is_better"] is True and d["value"] > 1e6 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "cfg2" in d["config"]["workload"]


def test_gpu_arm_declares_the_same_metric():
    import bench
    assert bench.METRIC == "line-gridpoint evals/s" and bench.UNIT == "pairs/s"
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert bench.METRIC.split()[0] in base["metric"]


def test_real_reference_timing_fixture_describes_cfg1():
    """tests/golden/reference_cfg1_timing.json (the real, unmodified reference timed on cfg1 in the build container,
    scripts/time_reference_cfg1.py) is on the workload bench.py's cfg1 object runs: same lines, same accumulate count, and the
    timed run itself agreed with the oracle."""
    import numpy as np
    from oracle import physics as ph
    from pyrad_b200 import workloads
    d = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_cfg1_timing.json")))
    w = workloads.cfg1()
    n = ph.grid_len(w["range_min"], w["range_max"], w["res"])
    ln = w["per_group_lines"][0]
    lo, hi = ph.effective_range(w["range_min"], w["range_max"], w["cutoff"])
    nu = ln["nu"][(ln["nu"] > lo) & (ln["nu"] < hi)]
    assert d["pairs"] == ph.pair_count(ph.line_index(nu, w["range_min"], w["res"]), n, ph.window_len(w["cutoff"], w["res"]))
    assert d["cores"] == 1 and d["get_transmittance_s"] > 1.0 and d["max_abs_T_diff_oracle_vs_reference"] <= 1e-13


def test_rank_affinity_helper_never_fails_and_never_widens():
    """bench.pin_rank_to_gpu_numa binds a rank to its GPU's CPUs when the box tells which those are; without nvidia-smi or
    sysfs (here) it reports why nothing changed and leaves the affinity alone."""
    import bench
    before = os.sched_getaffinity(0)
    msg = bench.pin_rank_to_gpu_numa(0)
    assert isinstance(msg, str) and msg
    assert os.sched_getaffinity(0) <= before
    os.sched_setaffinity(0, before)
