"""GPU box: per-phase times and counters of the single-GPU Barnes-Hut step for one workload (c4 / c5), optionally with a
forced sort depth (NB_SORT_LEVELS).  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import rust_exp_b200 as pkg  # noqa: E402
from rust_exp_b200 import ic  # noqa: E402


def main():
    n = int(os.environ.get("N", 262144))
    theta = float(os.environ.get("THETA", 0.5))
    gen = os.environ.get("GEN", "disk")
    steps = int(os.environ.get("STEPS", 20))
    lib = pkg.load()
    lib.init(0)
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    lib.set_stream(st.cuda_stream)
    lib.set_async(True)
    s = ic.random_disk(n, seed=4) if gen == "disk" else ic.plummer_2d(n, seed=4)
    lib.set_particles(s)
    for _ in range(6):
        lib.step_barnes_hut(theta, 0.01, 1)
    lib.synchronize()
    c0 = lib.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        lib.step_barnes_hut(theta, 0.01, 1)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    lib.phase_timing(True)
    for _ in range(steps):
        lib.step_barnes_hut(theta, 0.01, 1)
    ph = lib.phase_ms()
    lib.phase_timing(False)
    lib.synchronize()
    c = lib.counters()
    print(json.dumps({"n": n, "theta": theta, "gen": gen, "forced_levels": os.environ.get("NB_SORT_LEVELS"), "ms_per_step_back_to_back": ms,
                      "phases_ms": ph, "sort_levels": c["bh_sort_levels"], "sort_levels_before": c0["bh_sort_levels"]}))


if __name__ == "__main__":
    main()
