"""GPU (>= 2 devices): the sharded path on real GPUs -- runs tools/dist_check.py under torchrun."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_parity_and_transports():
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tools", "dist_check.py")]
    env = dict(os.environ, NB_DIST_BIG="65536", NB_DIST_BH="131072")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    checks = [json.loads(l) for l in p.stdout.splitlines() if l.startswith("{")]
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert checks and all(c["pass"] for c in checks)
