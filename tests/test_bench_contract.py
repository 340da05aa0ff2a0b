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
