"""Multi-GPU soak: many back-to-back collective steps mixing every sharded code path (partitioned Barnes-Hut,
replicated Barnes-Hut, the three all-pairs transports, set/get/draw epochs) to flush out rare ordering bugs.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29555 tools/soak.py
Rank 0 prints one JSON line per phase; non-zero exit on any failure (a hung peer traps after ~4 s instead of hanging)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rust_exp_b200 as pkg  # noqa: E402
from rust_exp_b200 import binding, ic  # noqa: E402
from rust_exp_b200 import dist as nbdist  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    lib = pkg.load()
    lib.init(lr)
    nbh = int(os.environ.get("NB_SOAK_BH", 1 << 22))
    nap = int(os.environ.get("NB_SOAK_AP", 1 << 19))
    nbdist.wire(lib, max(nbh, nap), binding.TRANSPORT_NCCL)
    ok = True

    def report(name, passed, **kw):
        nonlocal ok
        ok = ok and bool(passed)
        if rank == 0:
            print(json.dumps({"phase": name, "pass": bool(passed), **kw}), flush=True)

    # ---- Barnes-Hut, partitioned, long run with interleaved reads -------------------------------------
    s = ic.random_disk(nbh, seed=5)
    s[:, 2:4] *= 0.1
    s[:, 4] *= 1e-3          # light bodies: a gentle, long-lived disk
    lib.dist_set_transport(binding.TRANSPORT_P2P_GATHER)
    lib.set_particles(s)
    steps = int(os.environ.get("NB_SOAK_STEPS", 300))
    t0 = time.perf_counter()
    for k in range(steps):
        lib.step_barnes_hut(0.75, 0.01, 1)
        if k % 97 == 50:
            fb = lib.draw(256, 256)
            assert fb[128, 128] == 0x00FF00FF
        if k % 113 == 60:
            g = lib.get_particles()
            assert np.isfinite(g).all()
    lib.synchronize()
    el = time.perf_counter() - t0
    g = lib.get_particles()
    allg = [None] * world
    dist.all_gather_object(allg, int(np.ascontiguousarray(g).view(np.uint32).sum(dtype=np.uint64)))
    report("bh_partitioned_soak", np.isfinite(g).all() and all(a == allg[0] for a in allg) and np.array_equal(g[:, 4], s[:, 4]),
           n=nbh, steps=steps, ms_per_step_incl_reads=el / steps * 1e3)

    # ---- alternate partitioned / replicated / brute-force(theta=0 path at a smaller size) ------------------
    s2 = ic.plummer_2d(nap, seed=6)
    lib.set_particles(s2)
    for k in range(30):
        lib.bh_partition(1 if k % 3 == 0 else 0)
        lib.dist_set_transport(k % 3)
        lib.step_barnes_hut(0.5, 0.01, 1)
        if k % 10 == 9:
            lib.step_barnes_hut(0.0, 0.01, 1)     # all-pairs step through the Barnes-Hut entry point
    lib.bh_partition(0)
    g2 = lib.get_particles()
    report("mixed_paths", np.isfinite(g2).all(), n=nap)

    # ---- all-pairs transports back to back ------------------------------------------------------------------
    for tr in (0, 1, 2, 0, 2, 1):
        lib.dist_set_transport(tr)
        for _ in range(5):
            lib.step_brute_force(0.01)
    g3 = lib.get_particles()
    m = g3[:, 4:5].astype(np.float64)
    p = (m * g3[:, 2:4]).sum(0)
    report("allpairs_transports", np.isfinite(g3).all(), momentum=[float(p[0]), float(p[1])])
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
