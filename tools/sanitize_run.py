"""Small-N pass through every kernel of libnbody_b200, meant to be run under compute-sanitizer on the GPU box:
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
    compute-sanitizer --tool synccheck python tools/sanitize_run.py
(the reference has no race detection; this is the build's equivalent.)  Sizes are small because the
tools slow kernels down 10-100x.  Prints one line per covered path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_exp_b200 as pkg  # noqa: E402
from rust_exp_b200 import binding, ic  # noqa: E402


def main():
    os.environ["NB_BH_PARTS_MIN_N"] = "0"
    lib = pkg.load()
    lib.init(0)
    scale = int(os.environ.get("SAN_SCALE", "1"))

    def done(name):
        lib.synchronize()
        print(f"sanitize_run: {name} ok", flush=True)

    s = ic.random_disk(3000 * scale, seed=1)
    lib.set_particles(s)
    for _ in range(2):
        lib.step_brute_force(0.01)
    lib.get_particles()
    done("all-pairs FAST (TMA ring, packed FP32) + integrate")
    lib.tune(2, 0, 0); lib.step_brute_force(0.01); lib.tune(1, 0, 0); lib.step_brute_force(0.01); lib.tune(0, 0, 0)
    done("all-pairs FAST, 2 and 1 bodies per thread variants")
    lib.set_mode(binding.MODE_EXACT)
    lib.set_particles(s[:1500])
    lib.step_brute_force(0.01)
    lib.step_barnes_hut(0.5, 0.01, 1)
    done("EXACT all-pairs + EXACT Barnes-Hut (parallel exact build or serial build)")
    m = s[:1500].copy(); m[1, :2] = m[0, :2] + np.float32(3e-5)
    lib.set_particles(m)
    lib.step_barnes_hut(0.5, 0.01, 1)
    done("EXACT Barnes-Hut with a too-close pair (serial Node::insert restatement)")
    lib.set_mode(binding.MODE_FAST)
    b = ic.random_disk(20000 * scale, seed=2)
    lib.set_particles(b)
    for _ in range(4):
        lib.step_barnes_hut(0.5, 0.01, 1)          # graph replay, partial sort + fix-up from the 3rd step on
    lib.bh_count_interactions(True)
    lib.step_barnes_hut(0.75, 0.01, 1)              # counting build of the walk, no graph
    lib.bh_count_interactions(False)
    lib.bh_accelerations(0.5)
    lib.bh_flatten()
    done("FAST Barnes-Hut single tree (aabb, keys, sort, fix-up, gather, scans, build, walk, integrate; graph + direct)")
    for parts in (2, 4, 8):
        lib.bh_partition(parts)
        lib.set_particles(b)
        for _ in range(3):
            lib.step_barnes_hut(0.5, 0.01, 1)
        lib.get_particles()
    lib.bh_partition(0)
    done("FAST Barnes-Hut domain-partitioned, 2/4/8 virtual ranks (boxes, send, merge, per-part build, top tree, walk)")
    lib.set_integrator(binding.INTEGRATOR_LEAPFROG_KDK); lib.set_square_aabb(True)
    lib.set_particles(b[:5000])
    lib.step_barnes_hut(0.5, 0.01, 1); lib.step_brute_force(0.01); lib.get_particles()
    lib.set_integrator(binding.INTEGRATOR_EULER); lib.set_square_aabb(False)
    done("opt-ins: leapfrog KDK closing kick, squared root box")
    lib.random_disk(5000); lib.stable_orbits(5000, 0.5, 30.0); lib.plummer(5000, 5.0, 1e-2)
    lib.draw(160, 120)
    done("device generators + nb_draw scatter")
    lib.configure3(binding.LAW3_NEWTON, 1e-4)
    lib.set_particles3(ic.plummer_3d(2048 * scale, seed=3))
    lib.step3(0.01); lib.accelerations3()
    lib.configure3(binding.LAW3_REF, 1e-4); lib.step3(0.01)
    done("nbx3 all-pairs (NEWTON, REF)")
    lib.shutdown()
    print("sanitize_run: complete", flush=True)


if __name__ == "__main__":
    main()
