"""Diagnostic (GPU box): how far apart are the FAST Barnes-Hut step, the oracle (reference f32 semantics) and an
f64 evaluation at configs[4] size (4,194,304 bodies, theta = 0.75)?  Prints one JSON line per measurement."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (test tool)
import rust_exp_b200 as pkg  # noqa: E402
from rust_exp_b200 import binding, ic  # noqa: E402


def q(x):
    return {k: float(v) for k, v in zip(("median", "q99", "q999", "max"), (np.median(x), np.quantile(x, 0.99), np.quantile(x, 0.999), x.max()))}


def main():
    n = int(os.environ.get("N", 1 << 22))
    theta = float(os.environ.get("THETA", 0.75))
    gen = os.environ.get("GEN", "disk")
    lib = pkg.load()
    lib.init(0)
    o = oracle.get()
    nc = os.cpu_count() or 1
    s = ic.random_disk(n, seed=5) if gen == "disk" else ic.plummer_2d(n, seed=5)
    o.set_particles(s)
    t = time.time(); o.bh_build(); tb = time.time() - t
    t = time.time(); f = o.bh_forces_rows(theta, 0, n).astype(np.float64) / s[:, 4:5]; tf = time.time() - t
    lib.set_particles(s)
    a = lib.bh_accelerations(theta).astype(np.float64)
    scale = np.abs(f).max()
    err = np.abs(a - f).max(1) / scale
    rel = np.sqrt(((a - f) ** 2).sum(1)) / np.maximum(np.sqrt((f ** 2).sum(1)), 1e-30)
    print(json.dumps({"what": "accel gpu FAST vs oracle", "n": n, "gen": gen, "theta": theta, "oracle_build_s": tb, "oracle_forces_s": tf,
                      "err_over_max": q(err), "rel_per_body": q(rel), "amax": float(scale), "amedian": float(np.median(np.sqrt((f ** 2).sum(1))))}), flush=True)
    rows = np.random.default_rng(2).choice(n, 512, replace=False).astype(np.int32)
    t = time.time(); a64 = o.accel_f64_rows(rows); t64 = time.time() - t
    n64 = np.maximum(np.sqrt((a64 ** 2).sum(1)), 1e-30)
    e_gpu = np.sqrt(((a[rows] - a64) ** 2).sum(1)) / n64
    e_ref = np.sqrt(((f[rows] - a64) ** 2).sum(1)) / n64
    print(json.dumps({"what": "sampled rows vs f64 brute force", "rows": 512, "f64_s": t64, "gpu_rel": q(e_gpu), "ref_rel": q(e_ref),
                      "gpu_mean": float(e_gpu.mean()), "ref_mean": float(e_ref.mean())}), flush=True)
    for dt in (0.01, 0.001):
        o.set_particles(s); o.step_barnes_hut(theta, dt, nc); r = o.get_particles()
        lib.set_particles(s); lib.step_barnes_hut(theta, dt, 1); g = lib.get_particles()
        ext, vext = np.abs(r[:, :2]).max(), np.abs(r[:, 2:4]).max()
        ep = np.abs(g[:, :2].astype(np.float64) - r[:, :2]).max(1) / ext
        ev = np.abs(g[:, 2:4].astype(np.float64) - r[:, 2:4]).max(1) / vext
        disp = np.abs(r[:, :2].astype(np.float64) - s[:, :2]).max(1)
        print(json.dumps({"what": "one step gpu FAST vs oracle", "dt": dt, "ext": float(ext), "vext": float(vext), "pos": q(ep), "vel": q(ev),
                          "displacement_median": float(np.median(disp)), "frac_pos_le_1e-4": float((ep <= 1e-4).mean())}), flush=True)
    if os.environ.get("EXACT", "1") == "1":
        lib.set_mode(binding.MODE_EXACT)
        lib.set_particles(s)
        t = time.time(); lib.step_barnes_hut(theta, 0.01, 1); g = lib.get_particles(); te = time.time() - t
        o.set_particles(s); o.step_barnes_hut(theta, 0.01, nc); r = o.get_particles()
        print(json.dumps({"what": "EXACT one step vs oracle", "bitwise": bool(np.array_equal(g.view(np.uint32), r.view(np.uint32))), "gpu_s": te}), flush=True)
        lib.set_mode(binding.MODE_FAST)


if __name__ == "__main__":
    main()
