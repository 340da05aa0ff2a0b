"""CPU, world_size 2, gloo: the N>1 host path.

What runs here is the HOST side of the sharded step (rust_exp_b200/dist.py): shard arithmetic, the
handle exchange / unique-id broadcast sequence, and the claim the design rests on -- that sharding rows i
by index and giving every rank all j (in global index order) reproduces the single-process result bit
for bit.  The per-rank force evaluation is delegated to the oracle (this is a test, the only place that
may), standing in for the GPU kernel; the GPU version of the same check is tests/test_gpu_dist.py.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rust_exp_b200 import ic
from rust_exp_b200.dist import (ShardLayout, all_gather_bytes, broadcast_bytes, merge_by_rank, run_boundaries, sfc_partition,
                                wire)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class FakeLib:
    """Stands in for libnbody_b200 on a GPU-less host: records the wiring calls."""

    def __init__(self, rank):
        self.rank, self.log, self.n = rank, [], 0

    def dist_init(self, rank, world, maxp):
        self.log.append(("init", rank, world, maxp)); self.world, self.maxp = world, maxp

    def dist_export(self):
        return bytes([self.rank]) * 128

    def dist_import(self, handles, world):
        self.log.append(("import", handles, world))

    def dist_nccl_unique_id(self):
        return b"U" * 128

    def dist_nccl_init(self, uid):
        self.log.append(("nccl", uid))

    def dist_set_transport(self, t):
        self.log.append(("transport", t))

    def num_particles(self):
        return self.n

    def dist_local_range(self):
        return ShardLayout.for_capacity(self.maxp, self.world).local_range(self.rank, self.n)


def _worker(rank, world, port, n, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        o = oracle.get()
        # --- plumbing ---
        got = all_gather_bytes(bytes([rank]) * 128)
        assert got == [bytes([r]) * 128 for r in range(world)]
        assert broadcast_bytes(b"id" if rank == 0 else None) == b"id"
        fl = FakeLib(rank)
        lay = wire(fl, n, transport=2)
        assert fl.log[0] == ("init", rank, world, n)
        assert fl.log[1] == ("import", b"".join(bytes([r]) * 128 for r in range(world)), world)
        assert fl.log[2] == ("nccl", b"U" * 128) and fl.log[3] == ("transport", 2)

        # --- sharded step: rows by index, all j in global order, gather of positions each step ---
        full = ic.random_disk(n, seed=31)
        b, c = lay.local_range(rank, n)
        mine = full[b:b + c].copy()
        dt = np.float32(0.01)
        for _ in range(steps):
            # all-gather the shards' positions (padded to L like the arena), rebuild the global order
            pad = np.zeros((lay.shard_len, 5), dtype=np.float32)
            pad[:c] = mine
            bufs = [torch.zeros(lay.shard_len, 5) for _ in range(world)]
            dist.all_gather(bufs, torch.from_numpy(pad))
            glob = np.concatenate([bufs[g].numpy()[: lay.local_range(g, n)[1]] for g in range(world)])
            o.set_particles(glob)
            f = o.brute_forces_rows(b, b + c)  # local rows against ALL j, ascending global j
            m = mine[:, 4]
            mine[:, 2] = mine[:, 2] + (dt * f[:, 0]) / m
            mine[:, 3] = mine[:, 3] + (dt * f[:, 1]) / m
            mine[:, 0] = mine[:, 0] + dt * mine[:, 2]
            mine[:, 1] = mine[:, 1] + dt * mine[:, 3]
        outs = [None] * world
        dist.all_gather_object(outs, mine)
        if rank == 0:
            q.put(np.concatenate(outs))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1500, 2048])  # ragged (rank 1 short) and exact
def test_sharded_step_equals_single_process(oracle, n):
    world, steps = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, PORT[0], n, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    oracle.set_particles(ic.random_disk(n, seed=31))
    for _ in range(steps):
        oracle.step_brute_force(0.01)
    ref = oracle.get_particles()
    assert np.array_equal(res.view(np.uint32), ref.view(np.uint32))


PORT = [0]


@pytest.fixture(autouse=True)
def _port():
    PORT[0] = _free_port()
    yield


# ---- the Barnes-Hut exchange of the domain-partitioned step, host-side model over gloo -----------------------------
def _bh_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        # 48-bit Morton-like keys with many ties in the top bits AND some fully identical keys
        keys_all = (rng.integers(0, 1 << 14, n).astype(np.uint64) << np.uint64(34)) | rng.integers(0, 4, n).astype(np.uint64)
        lay = ShardLayout.for_capacity(n, world).for_set(n)
        b, c = lay.local_range(rank, n)
        mine = keys_all[b:b + c]
        order = np.argsort(mine, kind="stable")                   # the rank's own radix sort (stable)
        ks = mine[order]
        cells = (ks >> np.uint64(38)).astype(np.int64)            # cut-level cell = top 10 of 48 bits
        # every rank derives the same partition from the same (all-reduced) cell histogram
        hist = torch.from_numpy(np.bincount((keys_all[b:b + c] >> np.uint64(38)).astype(np.int64), minlength=1024))
        dist.all_reduce(hist)
        cut = sfc_partition(hist.tolist(), world)
        lo = run_boundaries(cells.tolist(), cut)
        runs_out = [[(int(ks[i]), b + int(order[i])) for i in range(lo[p], lo[p + 1])] for p in range(world)]
        # the all-to-all-v: run p goes to rank p (NVLink peer stores on the device; object all-gather here)
        box = [None] * world
        dist.all_gather_object(box, runs_out)
        inbox = [box[s][rank] for s in range(world)]              # what every source sent to ME, in source order
        keys, origin = merge_by_rank([[k for k, _ in r] for r in inbox])
        gidx = [inbox[s][i][1] for s, i in origin]
        out = [None] * world
        dist.all_gather_object(out, (keys, gidx, cut))
        if rank == 0:
            q.put((out, keys_all))
    finally:
        dist.destroy_process_group()


def test_partitioned_exchange_equals_global_stable_sort():
    """Shard-local sort -> runs by destination part -> exchange -> merge by ranking gives, part after part, exactly the
    global stable sort by (key, global index): what one GPU's radix sort over all bodies would have produced."""
    world, n = 2, 6000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bh_worker, args=(r, world, PORT[0], n, q)) for r in range(world)]
    for p in procs:
        p.start()
    out, keys_all = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out[0][2] == out[1][2]                                 # same cuts on every rank
    merged_keys = [k for part in out for k in part[0]]
    merged_idx = [g for part in out for g in part[1]]
    ref = np.argsort(keys_all, kind="stable")
    assert merged_keys == [int(k) for k in keys_all[ref]]
    assert merged_idx == [int(i) for i in ref]
    sizes = [len(part[0]) for part in out]
    assert sum(sizes) == n and max(sizes) - min(sizes) <= n // 8   # cell-granular balance


# ---- the sharded nbx3 step (row shards + all-gather of the positions), host-side model over gloo -----------------------
def _x3_accel(pos, m, rows, eps2=np.float32(1e-4)):
    """f32 model of the nbx3 Newtonian law for the given rows against ALL bodies in ascending j (row-independent sums)."""
    p = pos[rows]
    dx = pos[None, :, 0] - p[:, None, 0]
    dy = pos[None, :, 1] - p[:, None, 1]
    dz = pos[None, :, 2] - p[:, None, 2]
    d2 = dx * dx + dy * dy + dz * dz + eps2
    w = (m[None, :] / (d2 * np.sqrt(d2))).astype(np.float32)
    return np.stack([(w * dx).sum(1, dtype=np.float32), (w * dy).sum(1, dtype=np.float32), (w * dz).sum(1, dtype=np.float32)], 1)


def _x3_worker(rank, world, port, n, steps, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rust_exp_b200.dist import nbx3_local_rows, nbx3_shard_len

        fl = FakeLib(rank)
        wire(fl, n, transport=0, nccl=True)                     # the sharded nbx3 path needs the communicator with ANY transport
        assert ("nccl", b"U" * 128) in fl.log and fl.log[-1] == ("transport", 0)
        s = ic.plummer_3d(n, seed=21)
        shard = nbx3_shard_len(n, world)
        b, c = nbx3_local_rows(n, rank, world)
        pos = np.zeros((shard * world, 3), dtype=np.float32)    # every rank keeps ALL positions (the all-gather's receive buffer)
        pos[:n] = s[:, :3]
        vel = s[b:b + c, 3:6].copy()                            # velocities live with their rows
        m = s[:, 6].copy()
        dt = np.float32(0.01)
        rows = np.arange(b, b + c)
        for _ in range(steps):
            a = _x3_accel(pos[:n], m, rows)
            vel += dt * a
            pos[b:b + c] += dt * vel
            bufs = [torch.zeros(shard, 3) for _ in range(world)]
            dist.all_gather(bufs, torch.from_numpy(pos[rank * shard:(rank + 1) * shard].copy()))   # in place on the device
            pos = np.concatenate([t.numpy() for t in bufs])
        outs = [None] * world
        dist.all_gather_object(outs, (pos[:n].copy(), vel))
        if rank == 0:
            q.put(outs)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [3000, 4096])
def test_sharded_nbx3_step_equals_single_process(n):
    """Rows sharded by index + one all-gather of the new positions per step == every row computed in one process, bit for
    bit; every rank ends up with the same complete position set (the GPU version: tools/dist_check.py nbx3_sharded_*)."""
    world, steps = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_x3_worker, args=(r, world, PORT[0], n, steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    s = ic.plummer_3d(n, seed=21)
    pos, vel, m, dt = s[:, :3].copy(), s[:, 3:6].copy(), s[:, 6].copy(), np.float32(0.01)
    for _ in range(steps):
        a = _x3_accel(pos, m, np.arange(n))
        vel += dt * a
        pos += dt * vel
    assert np.array_equal(outs[0][0].view(np.uint32), outs[1][0].view(np.uint32))          # all ranks agree
    assert np.array_equal(outs[0][0].view(np.uint32), pos.view(np.uint32))
    assert np.array_equal(np.concatenate([o[1] for o in outs]).view(np.uint32), vel.view(np.uint32))
