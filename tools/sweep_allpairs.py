"""GPU: sweep the all-pairs kernel's tuning knobs (bodies/thread, CTAs/SM, waves) and print pairs/s.
Usage: python tools/sweep_allpairs.py [n_bodies]   -> JSON lines (gpurun_out/sweep.jsonl when redirected)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_exp_b200 as pkg  # noqa: E402
from rust_exp_b200 import ic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
lib = pkg.load()
lib.init(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
lib.set_stream(st.cuda_stream)
lib.set_particles(ic.plummer_2d(n, seed=3))
shares = (0, 1, 2) if os.environ.get("NB_SWEEP_SHARE") else (0,)
wave_list = tuple(int(v) for v in os.environ.get("NB_SWEEP_WAVES", "64").split(","))
for share in shares:
  os.environ["NB_SHARE_RCP"] = str(share)
  for bpt, cps in [(1, 5), (2, 3), (2, 4), (4, 2), (4, 3)]:
    for waves in wave_list:
        lib.tune(bpt, waves, cps)
        for _ in range(2):
            lib.step_brute_force(0.01)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        reps = 5
        e0.record(st)
        for _ in range(reps):
            lib.step_brute_force(0.01)
        e1.record(st)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"n": n, "share_rcp": share, "bodies_per_thread": bpt, "ctas_per_sm": cps, "waves": waves, "ms": ms,
                          "pairs_per_s": n * (n - 1) / (ms * 1e-3), "frac_fp32": n * (n - 1) * 12 / (ms * 1e-3) / 74.45e12}), flush=True)
