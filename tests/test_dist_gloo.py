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
