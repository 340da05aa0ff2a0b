"""CPU: the bench.py contract that can be checked without a GPU -- the reference arm runs the oracle port on
host cores and prints exactly one JSON line with the agreed keys; the GPU arm refuses to run without a GPU
(no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args, timeout=240):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_one_json_line():
    p = run("--impl", "reference", "--workload", "c2", "--steps", "1", "--warmup", "0")
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "pair-interactions/s" and d["unit"] == "pair-interactions/s"
    assert d["higher_is_better"] is True and d["value"] > 1e7
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("c2:") and d["config"]["n_bodies"] == 65536


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=60, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = run("--steps", "1", timeout=120)
    assert p.returncode != 0
    assert "no CPU path" in (p.stderr + p.stdout) or "no CUDA device" in (p.stderr + p.stdout)


def test_clock_sampler_parses_nvidia_smi_rows():
    """bench.py's clocks record: median SM clock under load, max clock, active throttle reasons in the timed window."""
    import importlib.util
    import time

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    cs = b.ClockSampler(0)

    class P:
        def terminate(self):
            pass

    cs.proc = P()
    t = time.perf_counter()
    cs.rows = [
        (t - 5.0, "120, 1965, 140.0, Not Active, Not Active, Not Active, Not Active"),     # before the window: ignored
        (t + 0.1, "1965, 1965, 640.5, Not Active, Not Active, Not Active, Not Active"),
        (t + 0.2, "1950, 1965, 655.0, Not Active, Not Active, Not Active, Active"),
        (t + 0.3, "1965, 1965, 650.0, Not Active, Not Active, Not Active, Not Active"),
    ]
    out = cs.stop(t, t + 0.4)
    assert out["sm_mhz"] == 1965.0 and out["sm_max_mhz"] == 1965.0 and out["samples"] == 3
    assert out["reasons"] == ["sw_power_cap"] and abs(out["power_w_max"] - 655.0) < 1e-6


def test_workload_table_matches_baseline_configs():
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    w = b.WORKLOADS
    assert w["c2"]["n"] == 65536 and w["c3"]["n"] == 1048576 and w["c4"]["n"] == 262144 and w["c5"]["n"] == 4194304
    assert w["c4"]["theta"] == 0.5 and w["c5"]["theta"] == 0.75 and b.FLOP_PER_PAIR == 12
    cfg = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]
    assert "65,536" in cfg[1] and "1,048,576" in cfg[2] and "262,144" in cfg[3] and "4,194,304" in cfg[4]


@pytest.mark.parametrize("name,gpus", [("r02_bench_g1.json", 1), ("r02_bench_g8.json", 8)])
def test_committed_bench_lines_carry_the_contract_keys(name, gpus):
    """The bench lines committed under profiles/ (measured on B200 by tools/evidence_run.sh / final_multi.sh) are whole:
    headline metric, e2e with its copy bytes, roofline, clocks without thermal throttling, kernel launches, the Barnes-Hut
    sub-records and a passing parity check."""
    path = os.path.join(ROOT, "profiles", name)
    if not os.path.exists(path):
        pytest.skip(f"{name} not committed")
    d = json.load(open(path))
    assert d["metric"] == d["unit"] == "pair-interactions/s" and d["n_gpus"] == gpus and d["higher_is_better"] is True
    assert d["config"]["workload"].startswith("c3:") and d["config"]["n_bodies"] == 1048576 and d["data"] == "synthetic"
    assert d["value"] > 4e12 * gpus * 0.9 and d["gpu_launches"] > 0 and d["warmup"] >= 3
    assert abs(d["value"] - 1048576 * 1048575 * d["steps"] / (d["ms_per_step"] * d["steps"] * 1e-3)) / d["value"] < 1e-6
    e = d["e2e"]
    assert e["value"] > 0 and e["value"] <= d["value"] * 1.001 and e["h2d_bytes_per_step"] == 1048576 * 20 == e["d2h_bytes_per_step"]
    r = d["roofline"]
    assert r["bound"] == "fp32" and 0.6 < r["frac"] < 0.75 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    c = d["clocks"]
    assert c["sm_mhz"] >= 0.95 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    subs = d["bh"]
    assert set(subs) == ({"c4", "c5"} if gpus == 1 else {"c5"})
    for k, b in subs.items():
        assert b["steps_per_s"] > 0 and abs(b["steps_per_s"] - 1e3 / b["ms_per_step"]) / b["steps_per_s"] < 1e-6
        assert 0.3 < b["lane_efficiency"] < 1.0 and sum(b["lanes_per_pop_histogram"]) == b["pops_per_step"]
        assert b["interactions_per_step"] > 0 and b["e2e"]["steps_per_s"] > 0 and set(b["phases_ms"]) >= {"force", "sort", "build"}
    if gpus == 1:
        assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 1
        assert subs["c4"]["cpu_baseline"]["steps_per_s"] > 0 and subs["c5"]["cpu_baseline"]["steps_per_s"] > 0
    else:
        assert len(subs["c5"]["per_rank"]) == gpus and subs["c5"]["ordering_point_ms"] > 0
    p = d["parity_check"]
    assert p["pass"] is True and p["exact_bitwise"] is True and p["exact_bh_bitwise"] is True and p["ranks_agree"] is True
    assert p["fast_rel_pos_err"] <= 1e-4 and p["fast_bh_rel_pos_err"] <= 1e-4
