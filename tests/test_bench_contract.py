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
