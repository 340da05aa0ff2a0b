"""GPU: step time of the EXACT (bit-exact) and FAST modes at a few sizes, including the reference's default scene
(10,000 stable orbits, theta=0.85, dt=0.01 -- the README screenshot shows ~32.5 steps/s on a 2016 laptop core)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rust_exp_b200 as pkg  # noqa: E402
from rust_exp_b200 import binding, ic  # noqa: E402

lib = pkg.load()
lib.init(0)
lib.set_async(True)   # back-to-back steps: device time per step (the default synchronous calls add a host round trip each)


def timeit(fn, reps):
    fn(); lib.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    lib.synchronize()
    return (time.perf_counter() - t) / reps * 1e3


for name, s, theta in (("default_scene_10k_orbits", ic.stable_orbits(10000, 0.5, 30.0, seed=1), 0.85),
                       ("disk_65536", ic.random_disk(65536, seed=2), 0.5),
                       ("disk_262144", ic.random_disk(262144, seed=4), 0.5)):
    for mode, mname in ((binding.MODE_FAST, "fast"), (binding.MODE_EXACT, "exact")):
        lib.set_mode(mode)
        lib.set_particles(s)
        reps = 20 if mode == binding.MODE_FAST else 3
        bh = timeit(lambda: lib.step_barnes_hut(theta, 0.01, 1), reps)
        lib.set_particles(s)
        ap = timeit(lambda: lib.step_brute_force(0.01), reps) if s.shape[0] <= 65536 or mode == binding.MODE_FAST else None
        print(json.dumps({"scene": name, "mode": mname, "n": int(s.shape[0]), "bh_ms_per_step": bh, "bh_steps_per_s": 1e3 / bh,
                          "allpairs_ms_per_step": ap}), flush=True)
lib.set_mode(binding.MODE_FAST)
