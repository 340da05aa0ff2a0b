"""Generate tests/golden/*.npz.

PROVENANCE -- read this before trusting the files: the reference (rs-src/nbody.rs) cannot be built or run
in this image (no Rust toolchain) and ships no test vectors, so these are NOT reference outputs.  They are
outputs of the two independent restatements in oracle/ (C and numpy-float32), written only after both
agree bit for bit, plus the hand-derived known-answer vectors of SURVEY.md section 8c.  They pin the oracle against
regressions and travel to the GPU box, where /root/reference does not exist.

Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import oracle_np as onp  # noqa: E402
from rust_exp_b200 import ic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
o = oracle.get()


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def both_brute(aos, dt, steps):
    o.set_particles(aos)
    s = aos.copy()
    for _ in range(steps):
        o.step_brute_force(dt)
        s = onp.step_brute_force(s, dt)
    c = o.get_particles()
    assert np.array_equal(bits(c), bits(s)), "C and numpy restatements disagree (brute force)"
    return c


def both_bh(aos, theta, dt, steps, nthreads=3):
    o.set_particles(aos)
    s = aos.copy()
    for _ in range(steps):
        o.step_barnes_hut(theta, dt, nthreads)
        s = onp.step_barnes_hut(s, theta, dt)
    c = o.get_particles()
    assert np.array_equal(bits(c), bits(s)), "C and numpy restatements disagree (Barnes-Hut)"
    return c


def main():
    # KAT-1: two bodies, one brute-force step, hand-computed bit patterns
    kat1_in = np.array([[0, 0, 0, 0, 1], [1, 0, 0, 0, 2]], dtype=np.float32)
    kat1_out = both_brute(kat1_in, 0.01, 1)
    assert bits(kat1_out)[0, 0] == 0x3951B1B8 and bits(kat1_out)[0, 2] == 0x3CA3D2D8
    assert bits(kat1_out)[1, 0] == 0x3F7FF972 and bits(kat1_out)[1, 2] == 0xBC23D2D8

    disk = ic.random_disk(192, seed=11)
    orbits = ic.stable_orbits(160, 0.5, 30.0, seed=12)
    brute_disk = both_brute(disk, 0.01, 5)
    brute_orbits = both_brute(orbits, 0.01, 5)
    bh_disk_05 = both_bh(disk, 0.5, 0.01, 5)
    bh_disk_085 = both_bh(disk, 0.85, 0.01, 5)
    bh_orbits = both_bh(orbits, 0.85, 0.01, 5)

    # tree of the disk (DFS pre-order records) from both restatements
    o.set_particles(disk)
    o.bh_build()
    tree_c = o.bh_flatten()
    tree_np = onp.flatten(onp.build_tree(disk))
    assert np.array_equal(bits(tree_c), bits(tree_np))

    # a merge case (KAT-9): two bodies 5e-5 apart plus two far ones
    merge = np.array(
        [[1.0, 1.0, 0, 0, 1.0], [1.00005, 1.00002, 0, 0, 2.0], [-3.0, 2.0, 0, 0, 1.5], [4.0, -2.5, 0, 0, 0.5]],
        dtype=np.float32,
    )
    o.set_particles(merge)
    o.bh_build()
    tree_merge = o.bh_flatten()
    assert np.array_equal(bits(tree_merge), bits(onp.flatten(onp.build_tree(merge))))
    bh_merge = both_bh(merge, 0.5, 0.01, 3)

    # draw
    o.set_particles(disk)
    fb = o.draw(96, 64)

    np.savez_compressed(
        os.path.join(HERE, "nbody_golden.npz"),
        kat1_in=kat1_in, kat1_out=kat1_out,
        disk=disk, orbits=orbits,
        brute_disk_dt001_k5=brute_disk, brute_orbits_dt001_k5=brute_orbits,
        bh_disk_t05_dt001_k5=bh_disk_05, bh_disk_t085_dt001_k5=bh_disk_085, bh_orbits_t085_dt001_k5=bh_orbits,
        tree_disk=tree_c, merge=merge, tree_merge=tree_merge, bh_merge_t05_dt001_k3=bh_merge,
        draw_disk_96x64=fb,
    )
    print("wrote", os.path.join(HERE, "nbody_golden.npz"))


if __name__ == "__main__":
    main()
