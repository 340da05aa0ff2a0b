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
from rust_exp_b200.dist import ShardLayout, all_gather_bytes, broadcast_bytes, wire


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
