"""CPU: host-side logic -- the experiment plug-in mirror, the IC generators, the shard arithmetic."""
import numpy as np
import pytest

from rust_exp_b200 import ic
from rust_exp_b200.dist import ShardLayout, sfc_partition
from rust_exp_b200.experiment import RustNBodyExperiment


class FakeLib:
    """Records the C calls the plug-in makes (the boundary contract of hs-src/RustNBodyExperiment.hs)."""

    def __init__(self):
        self.calls = []
        self.n = 0

    def stable_orbits(self, n, a, b):
        self.calls.append(("nb_stable_orbits", n, a, b)); self.n = n

    def random_disk(self, n):
        self.calls.append(("nb_random_disk", n)); self.n = n

    def step_barnes_hut(self, theta, dt, nt):
        self.calls.append(("nb_step_barnes_hut", theta, dt, nt))

    def synchronize(self):
        pass

    def draw(self, w, h, fb):
        self.calls.append(("nb_draw", w, h))

    def num_particles(self):
        return self.n


def test_experiment_defaults_and_call_order():
    L = FakeLib()
    e = RustNBodyExperiment(L)
    assert L.calls == [("nb_stable_orbits", 10000, 0.5, 30.0)]  # hs-src/RustNBodyExperiment.hs:42
    assert (e.time_step, e.theta, e.num_threads, e.num_steps) == (0.01, 0.85, 1, 0)
    fb = np.zeros((48, 64), dtype=np.uint32)
    e.draw(fb)
    # theta FIRST (rs-src/nbody.rs:187), step before draw (hs-src/RustNBodyExperiment.hs:55-60)
    assert L.calls[1] == ("nb_step_barnes_hut", 0.85, 0.01, 1)
    assert L.calls[2] == ("nb_draw", 64, 48)
    assert e.num_steps == 1 and len(e.times) == 1


def test_experiment_keys_and_clamps():
    L = FakeLib()
    e = RustNBodyExperiment(L)
    e.key("W"); assert L.calls[-1] == ("nb_random_disk", 10000)
    e.key("E"); assert L.calls[-1] == ("nb_stable_orbits", 5, 5.0, 40.0)
    e.key("Q"); assert L.calls[-1] == ("nb_stable_orbits", 10000, 0.5, 30.0)
    e.key("X"); assert e.time_step == 0.02
    e.key("X", shift=True); e.key("X", shift=True); assert e.time_step == 0.005
    for _ in range(10):
        e.key("A")
    assert e.theta == 0.95  # clamp, hs-src/RustNBodyExperiment.hs:90-93
    for _ in range(40):
        e.key("A", shift=True)
    assert e.theta == 0.0  # theta can reach exactly 0 -> brute-force path (rs-src/nbody.rs:197)
    for _ in range(40):
        e.key("P")
    assert e.num_threads == 16
    for _ in range(40):
        e.key("P", shift=True)
    assert e.num_threads == 1


def test_experiment_status_string_format():
    L = FakeLib()
    e = RustNBodyExperiment(L)
    s = e.status_string()
    assert s.startswith("0 Steps, 1.0SPS/1000.00ms | 10K Bodies\n")
    assert "Time Step [X][x]: 0.0100 | Theta [A][a]: 0.85 | Threads [P][p]: 1" in s
    e.key("E")
    assert "| 5 Bodies" in e.status_string()
    e.times.extend([0.01, 0.03, 0.02])
    assert "50.0SPS/20.00ms" in e.status_string()  # median, hs-src/RustNBodyExperiment.hs:65


def test_ic_generators_match_reference_distributions():
    a = ic.stable_orbits(1024, 0.5, 30.0, seed=1)
    assert a.dtype == np.float32 and a.shape == (1024, 5)
    assert tuple(a[0]) == (0, 0, 0, 0, 1000.0)
    r = np.hypot(a[1:, 0], a[1:, 1]); v = np.hypot(a[1:, 2], a[1:, 3])
    assert r.min() >= 0.5 - 1e-4 and r.max() <= 30.0 + 1e-4
    assert np.allclose(v, np.sqrt(1000.0), rtol=1e-5)
    assert np.allclose(a[1:, 0] * a[1:, 2] + a[1:, 1] * a[1:, 3], 0, atol=1e-2)  # v perpendicular to r
    d = ic.random_disk(4096, seed=2)
    assert np.hypot(d[:, 0], d[:, 1]).max() <= 23.0 + 1e-4
    assert d[:, 4].min() >= 0.1 and d[:, 4].max() < 1.5 and np.abs(d[:, 2:4]).max() <= 3.5
    p = ic.plummer_2d(4096, seed=3)
    assert np.abs(p[:, :2]).max() < 55.0  # fits the kill box
    assert np.array_equal(ic.plummer_2d(4096, seed=3), p)  # reproducible


def test_shard_layout_matches_library_arithmetic():
    lay = ShardLayout.for_capacity(1 << 20, 8)
    assert lay.shard_len == 131072
    assert [lay.local_range(r, 1 << 20) for r in (0, 7)] == [(0, 131072), (917504, 131072)]
    lay = ShardLayout.for_capacity(1000, 2)  # ragged: L = 1024, rank 1 owns nothing
    assert lay.shard_len == 1024
    assert lay.local_range(0, 1000) == (0, 1000) and lay.local_range(1, 1000) == (1000, 0)
    lay = ShardLayout.for_capacity(5000, 4)  # L = 2048
    assert [lay.local_range(r, 5000) for r in range(4)] == [(0, 2048), (2048, 2048), (4096, 904), (5000, 0)]
    assert sum(c for _, c in (lay.local_range(r, 5000) for r in range(4))) == 5000
    assert lay.segment_order(2) == [2, 3, 0, 1]
    # a set smaller than the capacity is re-balanced over all ranks (shards follow n, not max_particles)
    cap = ShardLayout.for_capacity(1 << 22, 8)
    assert cap.shard_len == 524288 and cap.for_set(1 << 20).shard_len == 131072 and cap.for_set(1 << 22).shard_len == 524288
    assert cap.for_set(3000).shard_len == 1024 and cap.for_set(0).shard_len == 1024
    assert [cap.for_set(3000).local_range(r, 3000) for r in range(4)] == [(0, 1024), (1024, 1024), (2048, 952), (3000, 0)]
    assert ShardLayout.for_capacity(1000, 2).for_set(5000).shard_len == 1024   # never beyond the capacity
    with pytest.raises(ValueError):
        ShardLayout.for_capacity(10, 9)


def test_sfc_partition_properties():
    rng = np.random.default_rng(3)
    for nparts in (1, 2, 3, 8):
        for hist in (rng.integers(0, 5000, 1024), np.r_[np.zeros(1000, int), rng.integers(1, 9, 24)],
                     np.eye(1, 1024, 517, dtype=int)[0] * 100000, np.zeros(1024, int)):
            cut = sfc_partition(hist, nparts)
            assert len(cut) == nparts + 1 and cut[0] == 0 and cut[-1] == 1024
            assert all(a <= b for a, b in zip(cut, cut[1:]))            # contiguous, every cell owned once
            n = int(hist.sum())
            sizes = [int(hist[a:b].sum()) for a, b in zip(cut, cut[1:])]
            assert sum(sizes) == n
            if n:
                # never more than the ideal share plus one cell's worth of bodies
                assert max(sizes) <= n / nparts + hist.max()


# ---- round 2: mirrors of the 32-ary search and of the nbx3 row sharding ---------------------------------------------
def test_warp_first_true_equals_bisection_and_needs_few_rounds():
    import bisect
    import random

    from rust_exp_b200.dist import warp_first_true

    rnd = random.Random(5)
    for n in (0, 1, 31, 32, 33, 1000, 1024, 1025, 65536, 524288):
        for _ in range(12):
            keys = sorted(rnd.randrange(0, max(1, n // 3 + 1)) for _ in range(n))       # many ties
            k = rnd.randrange(-1, max(1, n // 3 + 2))
            for strict in (False, True):
                pred = (lambda p: keys[p] > k) if strict else (lambda p: keys[p] >= k)
                want = (bisect.bisect_right if strict else bisect.bisect_left)(keys, k)
                got, rounds = warp_first_true(0, n, pred)
                assert got == want
                assert rounds <= 4 if n <= 32 ** 3 * 32 else True                          # 524,288 < 32^4: at most 4 rounds
            lo = rnd.randrange(0, n + 1)
            got, _ = warp_first_true(lo, n, lambda p: keys[p] >= k)                       # sub-range (the merge's second search)
            assert got == max(lo, bisect.bisect_left(keys, k))


def test_nbx3_row_shards_cover_the_set_in_whole_tiles():
    from rust_exp_b200.dist import nbx3_local_rows, nbx3_shard_len

    for n in (0, 1, 777, 1024, 5000, 20000, 65536, 1 << 20, (1 << 20) + 1):
        for world in (1, 2, 3, 4, 8):
            shard = nbx3_shard_len(n, world)
            assert shard % 1024 == 0 and shard * world >= n
            assert shard * world - max(n, 1) < 1024 * world + 1024                        # no more than one spare tile per rank
            rows = [nbx3_local_rows(n, r, world) for r in range(world)]
            assert rows[0][0] == 0 and sum(c for _, c in rows) == n
            for r in range(1, world):
                assert rows[r][0] == min(n, rows[r - 1][0] + rows[r - 1][1]) or rows[r][1] == 0
                assert rows[r][0] % 1024 == 0 or rows[r][1] == 0                          # shards start on tile boundaries
