#!/usr/bin/env python
"""Writes the input states for tools/ref_vectors/nbody_vectors.rs (little-endian f32 AoS {px,py,vx,vy,m}, 20 B/body)
and the manifest tests/test_reference_vectors.py reads.  Deterministic (seeded)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from rust_exp_b200 import ic  # noqa: E402

f32 = np.float32


def inputs():
    out = {}
    out["kat1_two_bodies"] = np.array([[0, 0, 0, 0, 1.0], [1.0, 0, 0, 0, 2.0]], dtype=f32)
    out["disk_1024"] = ic.random_disk(1024, seed=101)
    out["orbits_1024"] = ic.stable_orbits(1024, 0.5, 30.0, seed=102)
    out["plummer_8192"] = ic.plummer_2d(8192, seed=103)
    out["orbits_10000"] = ic.stable_orbits(10000, 0.5, 30.0, seed=104)                           # the reference's default scene
    out["disk_16384"] = ic.random_disk(16384, seed=105)
    out["plummer_16384"] = ic.plummer_2d(16384, seed=106)
    m = ic.random_disk(500, seed=107)
    m[1, :2] = m[0, :2] + f32(3e-5)      # too close: merged leaf (rs-src/nbody.rs:249-260)
    m[3, :2] = m[2, :2]                  # identical positions
    m[5, 0] = m[4, 0] + f32(5e-5)        # close in x only: not merged
    out["merge_500"] = m
    out["kill_4"] = np.array([[56.0, 0, 1.0, 1.0, 1.0], [54.0, 0, 1.0, 1.0, 1.0], [0, 0, 0, 0, 1.0], [0, 55.5, 0.5, 0, 1.0]], dtype=f32)
    out["disk_65536"] = ic.random_disk(65536, seed=108)
    rng = np.random.default_rng(109)
    k = np.zeros((512, 5), dtype=f32)
    k[:, :2] = (rng.normal(0, 1, (512, 2)) * 10.0 ** rng.uniform(-4, 2, (512, 1))).astype(f32)
    k[:, 4] = (10.0 ** rng.uniform(-2, 3, 512)).astype(f32)
    k[1::2, :2] = np.where(rng.random((256, 1)) < 0.2, k[0::2, :2] + f32(1e-5), k[1::2, :2])     # near-coincident pairs
    out["force_kat_pairs"] = k
    return out


# (input, output, theta (< 0: brute force), dt, steps, nthreads) -- the CASES table of nbody_vectors.rs
CASES = [
    ("kat1_two_bodies", "kat1_two_bodies_brute_dt001_k1", -1.0, 0.01, 1, 1),
    ("disk_1024", "disk_1024_brute_dt001_k100", -1.0, 0.01, 100, 1),
    ("orbits_1024", "orbits_1024_brute_dt001_k100", -1.0, 0.01, 100, 1),
    ("plummer_8192", "plummer_8192_brute_dt001_k10", -1.0, 0.01, 10, 1),
    ("disk_1024", "disk_1024_bh_t05_dt001_k20", 0.5, 0.01, 20, 1),
    ("disk_1024", "disk_1024_bh_t0_dt001_k3", 0.0, 0.01, 3, 3),
    ("orbits_10000", "orbits_10000_bh_t085_dt001_k10", 0.85, 0.01, 10, 4),
    ("disk_16384", "disk_16384_bh_t05_dt001_k10", 0.5, 0.01, 10, 7),
    ("plummer_16384", "plummer_16384_bh_t075_dt001_k5", 0.75, 0.01, 5, 2),
    ("merge_500", "merge_500_bh_t05_dt001_k4", 0.5, 0.01, 4, 1),
    ("kill_4", "kill_4_bh_t05_dt0001_k1", 0.5, 0.001, 1, 1),
    ("disk_65536", "disk_65536_bh_t05_dt001_k2", 0.5, 0.01, 2, 8),
]


def main():
    d = os.path.join(HERE, "inputs")
    os.makedirs(d, exist_ok=True)
    for name, a in inputs().items():
        np.ascontiguousarray(a, dtype="<f4").tofile(os.path.join(d, name + ".bin"))
    json.dump({"cases": CASES}, open(os.path.join(HERE, "manifest.json"), "w"), indent=1)
    print(f"wrote {len(inputs())} input files to {d}")


if __name__ == "__main__":
    main()
