"""Multi-GPU parity + transport comparison, one rank per GPU.  Launch:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
Checks (rank 0 prints one JSON line per check, exit code != 0 on failure):
  * EXACT mode, sharded over N GPUs == single-process oracle, bit for bit (brute force and Barnes-Hut)
  * FAST mode: the three transports (P2P_DIRECT, P2P_GATHER, NCCL) give bitwise identical results
  * FAST sharded vs oracle within the 1e-4 tolerance; ragged sizes; nb_draw / nb_get on every rank
  * per-transport step time at a larger size
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (this is a test tool)
import rust_exp_b200 as pkg  # noqa: E402
from rust_exp_b200 import binding, ic  # noqa: E402
from rust_exp_b200 import dist as nbdist  # noqa: E402


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    lib = pkg.load()
    lib.init(lr)
    big = int(os.environ.get("NB_DIST_BIG", 262144))
    nbdist.wire(lib, max(big, 131072), binding.TRANSPORT_NCCL)  # NCCL comm + IPC arenas; transports switch below
    o = oracle.get()
    ok = True

    def report(name, passed, **kw):
        nonlocal ok
        ok = ok and bool(passed)
        if rank == 0:
            print(json.dumps({"check": name, "pass": bool(passed), **kw}), flush=True)

    # ---- EXACT sharded == oracle ----------------------------------------------------------------
    for n, steps in ((3000, 5), (4096, 3), (1, 2)):
        s = ic.random_disk(n, seed=41)
        lib.set_mode(binding.MODE_EXACT)
        for tr in (binding.TRANSPORT_P2P_DIRECT, binding.TRANSPORT_P2P_GATHER, binding.TRANSPORT_NCCL):
            lib.dist_set_transport(tr)
            lib.set_particles(s)
            for _ in range(steps):
                lib.step_brute_force(0.01)
            g = lib.get_particles()
            o.set_particles(s)
            for _ in range(steps):
                o.step_brute_force(0.01)
            report(f"exact_brute_n{n}_transport{tr}", np.array_equal(bits(g), bits(o.get_particles())))
        lib.dist_set_transport(binding.TRANSPORT_P2P_GATHER)
        lib.set_particles(s)
        for _ in range(steps):
            lib.step_barnes_hut(0.5, 0.01, 1)
        g = lib.get_particles()
        o.set_particles(s)
        for _ in range(steps):
            o.step_barnes_hut(0.5, 0.01, 2)
        report(f"exact_bh_n{n}", np.array_equal(bits(g), bits(o.get_particles())))
    lib.set_mode(binding.MODE_FAST)

    # ---- FAST: transports agree bitwise; within tolerance of the oracle ----------------------------
    n, steps = 20000, 10
    s = ic.plummer_2d(n, seed=42)
    outs = {}
    for tr in (binding.TRANSPORT_P2P_DIRECT, binding.TRANSPORT_P2P_GATHER, binding.TRANSPORT_NCCL):
        lib.dist_set_transport(tr)
        lib.set_particles(s)
        for _ in range(steps):
            lib.step_brute_force(0.01)
        outs[tr] = lib.get_particles()
    report("fast_transports_bitwise_equal", np.array_equal(bits(outs[0]), bits(outs[1])) and np.array_equal(bits(outs[0]), bits(outs[2])))
    o.set_particles(s)
    for _ in range(steps):
        o.step_brute_force(0.01)
    r = o.get_particles()
    err = float(np.abs(outs[0][:, :2].astype(np.float64) - r[:, :2]).max() / np.abs(r[:, :2]).max())
    report("fast_brute_vs_oracle", err <= 1e-4, rel_pos_err=err)
    # every rank must hold the same gathered state
    allg = [None] * world
    dist.all_gather_object(allg, bits(outs[0]).tobytes())
    report("get_particles_identical_on_all_ranks", all(a == allg[0] for a in allg))

    lib.dist_set_transport(binding.TRANSPORT_P2P_GATHER)
    lib.set_particles(s)
    for _ in range(5):
        lib.step_barnes_hut(0.5, 0.01, 1)
    g = lib.get_particles()
    o.set_particles(s)
    for _ in range(5):
        o.step_barnes_hut(0.5, 0.01, 2)
    r = o.get_particles()
    err = float(np.abs(g[:, :2].astype(np.float64) - r[:, :2]).max() / np.abs(r[:, :2]).max())
    report("fast_bh_vs_oracle", err <= 1e-4, rel_pos_err=err)
    # ---- domain-partitioned Barnes-Hut (one part per GPU, remote subtrees walked in place) vs replicated tree
    os.environ["NB_BH_PARTS_MIN_N"] = "0"
    nb_ = 120000
    sp = ic.random_disk(nb_, seed=44)
    lib.bh_count_interactions(True)
    res = {}
    for mode, name in ((1, "replicated"), (0, "partitioned")):
        lib.bh_partition(mode)
        lib.set_particles(sp)
        lib.reset_counters()
        acc = lib.bh_accelerations(0.5)           # each rank fills its own index shard's rows
        t = torch.from_numpy(acc).cuda()
        dist.all_reduce(t)                        # rows are disjoint -> sum assembles the full array
        cnt = lib.counters()
        ct = torch.tensor([cnt["bh_nodes_built"], cnt["bh_interactions"], cnt["bh_nodes_visited"]], dtype=torch.int64, device="cuda")
        dist.all_reduce(ct)
        res[name] = (t.cpu().numpy().astype(np.float64), [int(v) for v in ct.tolist()])
    o.set_particles(sp)
    o.bh_build()
    a_rep, c_rep = res["replicated"]
    a_par, c_par = res["partitioned"]
    report("bh_partitioned_node_count_equals_reference", c_par[0] == o.bh_node_count(), nodes=c_par[0], ref=o.bh_node_count())
    report("bh_partitioned_interaction_lists", abs(c_par[1] * world - c_rep[1]) <= 1e-5 * c_rep[1] or abs(c_par[1] - c_rep[1] // 1) <= 1e-5 * c_rep[1],
           part=c_par[1:], repl=c_rep[1:])
    err = np.abs(a_par - a_rep).max(1) / np.abs(a_rep).max()
    report("bh_partitioned_forces_equal_replicated", float(np.quantile(err, 0.999)) <= 1e-6 and float(err.max()) <= 2e-3,
           q999=float(np.quantile(err, 0.999)), max=float(err.max()))
    lib.bh_partition(0)
    lib.set_particles(sp)
    for _ in range(5):
        lib.step_barnes_hut(0.5, 0.01, 1)
    g = lib.get_particles()
    o.set_particles(sp)
    for _ in range(5):
        o.step_barnes_hut(0.5, 0.01, 8)
    r = o.get_particles()
    err = float(np.abs(g[:, :2].astype(np.float64) - r[:, :2]).max() / np.abs(r[:, :2]).max())
    report("bh_partitioned_5_steps_vs_oracle", err <= 1e-4, rel_pos_err=err)
    lib.bh_count_interactions(False)

    # ---- nbx3 extension sharded by rows (ncclAllGather of the positions) == every rank computing every row, bit for bit ----
    for n3, law in ((5000, binding.LAW3_NEWTON), (20000, binding.LAW3_REF)):
        s3 = ic.plummer_3d(n3, seed=46)
        lib.configure3(law, 1e-4)
        res = {}
        for sharded in (False, True):
            lib.set_sharded3(sharded)
            lib.set_particles3(s3)
            a3 = lib.accelerations3()
            for _ in range(4):
                lib.step3(0.01)
            res[sharded] = (a3, lib.get_particles3())
        same = np.array_equal(bits(res[True][0]), bits(res[False][0])) and np.array_equal(bits(res[True][1]), bits(res[False][1]))
        every = [None] * world
        dist.all_gather_object(every, bits(res[True][1]).tobytes())
        # and against an f64 evaluation of the same law
        p = s3[:, :3].astype(np.float64); mm = s3[:, 6].astype(np.float64)
        ref = np.zeros((n3, 3))
        for i0 in range(0, n3, 1000):
            d = p[None, :, :] - p[i0:i0 + 1000, None, :]
            d2 = (d * d).sum(-1) + 1e-4
            wgt = mm[None, :] / (d2 ** 1.5 if law == binding.LAW3_NEWTON else d2)
            ref[i0:i0 + 1000] = (wgt[:, :, None] * d).sum(1)
        err = float(np.abs(res[True][0] - ref).max() / np.abs(ref).max())
        report(f"nbx3_sharded_n{n3}_law{law}", same and all(b == every[0] for b in every) and err <= 2e-5 and np.isfinite(res[True][1]).all(),
               acc_err_vs_f64=err)
    lib.set_sharded3(True)

    # ---- device-side generators when sharded: counter-based RNG keyed by the GLOBAL body index ----------
    lib.seed(77)
    lib.stable_orbits(50000, 0.5, 30.0)
    ga = lib.get_particles()
    gathered = [None] * world
    dist.all_gather_object(gathered, bits(ga).tobytes())
    rr = np.hypot(ga[1:, 0], ga[1:, 1])
    report("generators_sharded", all(a == gathered[0] for a in gathered) and tuple(ga[0]) == (0, 0, 0, 0, 1000.0)
           and rr.min() >= 0.5 - 1e-3 and rr.max() <= 30.0 + 1e-3 and lib.num_particles() == 50000)
    lib.step_barnes_hut(0.85, 0.01, 1)   # the reference's default frame: step, then draw
    fb0 = lib.draw(160, 120)
    report("default_scene_frame_sharded", fb0[60, 80] == 0x00FF00FF and (fb0 != 0).sum() > 1000)
    lib.set_particles(sp[:20000])

    fb = lib.draw(128, 96)
    g = lib.get_particles()
    o.set_particles(g)
    report("draw_sharded", np.array_equal(fb, o.draw(128, 96)))

    # ---- timing per transport -----------------------------------------------------------------------
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    lib.set_stream(st.cuda_stream)
    sb = ic.plummer_2d(big, seed=43)
    for tr, name in ((0, "p2p_direct"), (1, "p2p_gather"), (2, "nccl")):
        lib.dist_set_transport(tr)
        lib.set_particles(sb)
        for _ in range(3):
            lib.step_brute_force(0.01)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        reps = 10
        for _ in range(reps):
            lib.step_brute_force(0.01)
        e1.record(st)
        torch.cuda.synchronize(); dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        report(f"time_{name}", True, n=big, world=world, ms_per_step=float(t.item()),
               pairs_per_s=big * (big - 1) / (float(t.item()) * 1e-3))
    # Barnes-Hut step time, replicated vs partitioned
    nbh = int(os.environ.get("NB_DIST_BH", 1 << 20))
    if nbh <= max(big, 131072):
        sb2 = ic.random_disk(nbh, seed=45)
        for mode, name in ((1, "bh_replicated"), (0, "bh_partitioned")):
            lib.bh_partition(mode)
            lib.dist_set_transport(1)
            lib.set_particles(sb2)
            for _ in range(3):
                lib.step_barnes_hut(0.75, 0.01, 1)
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            reps = 10
            for _ in range(reps):
                lib.step_barnes_hut(0.75, 0.01, 1)
            e1.record(st)
            torch.cuda.synchronize(); dist.barrier()
            t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            report(f"time_{name}", True, n=nbh, world=world, ms_per_step=float(t.item()))
        lib.bh_partition(0)
    lib.set_stream(None)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
