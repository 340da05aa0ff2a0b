"""torch.distributed plumbing for the index-sharded particle set (one process per GPU).

torch.distributed is used ONLY to wire the processes together (exchange the 128 bytes of CUDA IPC handles of
the symmetric arenas, broadcast the NCCL unique id, barrier / max-reduce timings).  The data path of a
step never goes through it: positions move GPU-to-GPU inside the library's own kernels (bulk-TMA loads
from peer HBM over NVLink) or, for the baseline transport, the library's own ncclAllGather.

The shard arithmetic here mirrors nb_dist.cu::dist_init / nb_engine.cu::local_begin so that it can be
tested on CPU (gloo) without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass

SHARD_ALIGN = 1024  # nb_engine.h kShardAlign
MAX_RANKS = 8  # nb_engine.h kMaxRanks


@dataclass(frozen=True)
class ShardLayout:
    """Rank g owns global bodies [g*L, min((g+1)*L, n)); slots beyond n are zero-mass padding.
    `for_capacity` gives the layout the arenas are allocated for; `for_set(n)` the balanced one a set of n uses."""

    world: int
    shard_len: int  # L

    @staticmethod
    def for_capacity(max_particles: int, world: int) -> "ShardLayout":
        if not (1 <= world <= MAX_RANKS):
            raise ValueError(f"world must be 1..{MAX_RANKS}")
        if max_particles < 1:
            raise ValueError("max_particles must be >= 1")
        per = (max_particles + world - 1) // world
        L = (per + SHARD_ALIGN - 1) // SHARD_ALIGN * SHARD_ALIGN
        return ShardLayout(world, L)

    def for_set(self, n: int) -> "ShardLayout":
        """Layout actually used for a set of n bodies: shards are balanced -- L = ceil(n / world) rounded up to
        the shard granularity, capped by the capacity the arena was allocated for (nb_engine.cu::ensure_capacity)."""
        per = (max(n, 0) + self.world - 1) // self.world
        L = max(SHARD_ALIGN, (per + SHARD_ALIGN - 1) // SHARD_ALIGN * SHARD_ALIGN)
        return ShardLayout(self.world, min(L, self.shard_len))

    def local_range(self, rank: int, n: int) -> tuple[int, int]:
        b = min(rank * self.shard_len, n)
        c = max(0, min(self.shard_len, n - rank * self.shard_len))
        return b, c

    def owner(self, i: int) -> int:
        return i // self.shard_len

    def segment_order(self, rank: int) -> list[int]:
        """Order in which a rank's all-pairs kernel consumes j segments: its own first (needs no
        cross-rank wait), then ring order -- nb_allpairs.cu `g = (my_rank + q) % nseg`."""
        return [(rank + q) % self.world for q in range(self.world)]


def nbx3_shard_len(n: int, world: int) -> int:
    """Rows per rank of the sharded nbx3 set (mirror of nb_3d.cu::x3_set): the set is padded to a multiple of 1,024 and
    the 1,024-row tiles are dealt out evenly; world * shard rows is also the length of every rank's arrays (they are the
    receive buffers of the in-place ncclAllGather)."""
    n_pad = (max(n, 0) + 1023) // 1024 * 1024 + (1024 if n <= 0 else 0)
    return (n_pad // 1024 + world - 1) // world * 1024


def nbx3_local_rows(n: int, rank: int, world: int) -> tuple[int, int]:
    """(begin, count) of the rows rank `rank` evaluates and integrates (mirror of nb_3d.cu::x3_step)."""
    shard = nbx3_shard_len(n, world)
    b = min(n, rank * shard)
    return b, min(n, b + shard) - b


def warp_first_true(lo: int, hi: int, pred, width: int = 32) -> tuple[int, int]:
    """Mirror of nb_bh.cu::warp_first_true, the 32-ary search of the send / merge / cell-table kernels: first index in
    [lo, hi) at which the monotone predicate (False..False True..True) holds, hi if none.  Every "lane" probes one
    position per round and the first true lane narrows the range `width`-fold.  Returns (index, rounds)."""
    rounds = 0
    while hi - lo > width:
        rounds += 1
        step = (hi - lo + width - 1) // width
        probes = [min(hi - 1, lo + (lane + 1) * step - 1) for lane in range(width)]
        hits = [pred(p) for p in probes]
        if not any(hits):
            return hi, rounds
        f = hits.index(True)
        hi = probes[f] + 1
        lo = lo + f * step
    rounds += 1
    for p in range(lo, min(hi, lo + width)):
        if pred(p):
            return p, rounds
    return hi, rounds


BODY_WEIGHT = 8  # nb_bh.cu kBodyWeight: build cost of one body in units of one walk pop


def sfc_partition(cell_weights, nparts: int) -> list[int]:
    """Cut points of the Barnes-Hut domain split (mirror of nb_bh.cu::bh_top_build_kernel).

    `cell_weights[c]` = weight of cut-level quadtree cell c, cells in Morton order -- on the device
    BODY_WEIGHT * bodies(c) + walk cost measured for c on the previous step (body counts alone on the first step).
    Part g owns the contiguous cell range [cut[g], cut[g+1]); a cut is placed at the first cell whose exclusive
    prefix weight reaches g*W/nparts, so every part gets ~W/nparts of the work up to the granularity of one cell,
    every cell has exactly one owner, and all ranks compute the same cuts from the same table."""
    total = int(sum(int(k) for k in cell_weights))
    cut = [0]
    cum = 0
    g = 1
    for c, k in enumerate(cell_weights):
        while g < nparts and cum * nparts >= g * total:
            cut.append(c)
            g += 1
        cum += int(k)
    while g <= nparts:
        cut.append(len(cell_weights))
        g += 1
    return cut


def run_boundaries(sorted_cells, cut) -> list[int]:
    """Mirror of nb_bh.cu::bhp_send_kernel: a shard sorted by key falls apart into len(cut)-1 contiguous runs, run p =
    the bodies whose cut-level cell lies in [cut[p], cut[p+1]).  Returns the run starts (+ the end)."""
    import bisect

    return [bisect.bisect_left(sorted_cells, c) for c in cut]


def merge_by_rank(runs):
    """Mirror of nb_bh.cu::bhp_merge_kernel: G key-sorted runs -> one sorted sequence without any global sort.  Element
    i of run s goes to position i + sum over the other runs r of (elements of r that are < key, or <= key when r < s):
    ties are ordered by source rank, then by position, so the result is deterministic.  Returns (keys, (source, index))."""
    import bisect

    n = sum(len(r) for r in runs)
    keys = [None] * n
    origin = [None] * n
    for s, run in enumerate(runs):
        for i, k in enumerate(run):
            pos = i
            for r, other in enumerate(runs):
                if r == s:
                    continue
                pos += bisect.bisect_right(other, k) if r < s else bisect.bisect_left(other, k)
            keys[pos] = k
            origin[pos] = (s, i)
    return keys, origin


def all_gather_bytes(payload: bytes, group=None) -> list[bytes]:
    import torch.distributed as dist

    out: list = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, payload, group=group)
    return out


def broadcast_bytes(payload: bytes | None, src: int = 0, group=None) -> bytes:
    import torch.distributed as dist

    box = [payload]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def wire(lib, max_particles: int, transport: int = 0, group=None, nccl: bool = False) -> ShardLayout:
    """Collective: allocate + exchange the symmetric arenas of all ranks and (optionally) set up NCCL.

    nccl=True creates the library's NCCL communicator whatever the transport (the sharded nbx3_* path needs it)."""
    import torch.distributed as dist

    from .binding import TRANSPORT_NCCL

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lib.dist_init(rank, world, max_particles)
    handles = all_gather_bytes(lib.dist_export(), group)
    lib.dist_import(b"".join(handles), world)
    if transport == TRANSPORT_NCCL or nccl:
        uid = lib.dist_nccl_unique_id() if rank == 0 else None
        uid = broadcast_bytes(uid, 0, group)
        lib.dist_nccl_init(uid)
    lib.dist_set_transport(transport)
    dist.barrier(group)
    lay = ShardLayout.for_capacity(max_particles, world)
    b, c = lib.dist_local_range()
    assert (b, c) == lay.local_range(rank, lib.num_particles()), "host/library shard arithmetic disagree"
    return lay
