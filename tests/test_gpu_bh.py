"""GPU: parity of nb_step_barnes_hut through the C ABI against the oracle.

EXACT mode (serial device restatement of Node::insert + per-body nested-sum walk) must equal the oracle
bit for bit, merges and kill hack included.  FAST mode (Morton build + warp walk) must reproduce the
reference's tree TOPOLOGY (node count) and per-body interaction lists (counts), and forces / positions
within the stated FP32 tolerance.
"""
import os

import numpy as np
import pytest

from rust_exp_b200 import binding, ic

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "nbody_golden.npz"))
f32 = np.float32


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def run_gpu(lib, s, theta, dt, steps, nthreads=1):
    lib.set_particles(s)
    for _ in range(steps):
        lib.step_barnes_hut(theta, dt, nthreads)
    return lib.get_particles()


def run_ora(oracle, s, theta, dt, steps, nthreads=2):
    oracle.set_particles(s)
    for _ in range(steps):
        oracle.step_barnes_hut(theta, dt, nthreads)
    return oracle.get_particles()


# ------------------------------------------------------------------ EXACT: bit parity ---------------
@pytest.mark.parametrize("key,src,theta,k", [
    ("bh_disk_t05_dt001_k5", "disk", 0.5, 5), ("bh_disk_t085_dt001_k5", "disk", 0.85, 5),
    ("bh_orbits_t085_dt001_k5", "orbits", 0.85, 5), ("bh_merge_t05_dt001_k3", "merge", 0.5, 3)])
def test_exact_golden(fresh, key, src, theta, k):
    fresh.set_mode(binding.MODE_EXACT)
    assert np.array_equal(bits(run_gpu(fresh, GOLD[src], theta, 0.01, k)), bits(GOLD[key]))


@pytest.mark.parametrize("n,steps,theta,gen", [
    (1, 2, 0.5, "disk"), (2, 3, 0.5, "disk"), (5, 5, 0.85, "orbits"), (2000, 20, 0.5, "disk"),
    (10000, 5, 0.85, "orbits"), (16384, 10, 0.5, "disk"), (16384, 5, 1e-6, "plummer")])
def test_exact_matches_oracle_bitwise(fresh, oracle, n, steps, theta, gen):
    s = ic.stable_orbits(n, 0.5 if n > 5 else 5.0, 30.0 if n > 5 else 40.0, seed=3) if gen == "orbits" else \
        (ic.random_disk(n, seed=3) if gen == "disk" else ic.plummer_2d(n, seed=3))
    fresh.set_mode(binding.MODE_EXACT)
    g = run_gpu(fresh, s, theta, 0.01, steps)
    r = run_ora(oracle, s, theta, 0.01, steps)
    assert np.array_equal(bits(g), bits(r))


def test_exact_c4_size_262144_one_step_bitwise(fresh, oracle):
    """configs[3] size in EXACT mode: the serial device build + nested-sum walk reproduce the oracle's step for
    all 262,144 bodies bit for bit (tree of ~770k nodes, merges included)."""
    n = 262144
    s = ic.random_disk(n, seed=4)
    fresh.set_mode(binding.MODE_EXACT)
    g = run_gpu(fresh, s, 0.5, 0.01, 1)
    r = run_ora(oracle, s, 0.5, 0.01, 1, nthreads=os.cpu_count() or 1)
    assert np.array_equal(bits(g), bits(r))


@pytest.mark.parametrize("n,gen", [(20000, "disk"), (65536, "orbits"), (8192, "plummer")])
def test_exact_parallel_build_is_the_reference_tree_bit_for_bit(fresh, oracle, n, gen, monkeypatch):
    """EXACT mode on a merge-free set takes the parallel build (sort-based topology + per-node f32 folds in particle
    index order).  Its tree, dumped in the oracle's format, must equal the reference tree in EVERY bit -- boxes, COMs,
    masses, flags -- and stepping must agree bitwise with the serial-build path and with the oracle."""
    s = ic.stable_orbits(n, 0.5, 30.0, seed=6) if gen == "orbits" else \
        (ic.random_disk(n, seed=6) if gen == "disk" else ic.plummer_2d(n, seed=6))
    fresh.set_mode(binding.MODE_EXACT)
    fresh.set_particles(s)
    fresh.bh_accelerations(0.5)
    t = fresh.bh_flatten()
    oracle.set_particles(s)
    oracle.bh_build()
    r = oracle.bh_flatten()
    if t.shape[0] == 0:
        pytest.skip("this set contains a too-close pair: the serial build ran (covered by the other EXACT tests)")
    assert t.shape == r.shape and np.array_equal(bits(t), bits(r))
    g_par = run_gpu(fresh, s, 0.5, 0.01, 3)
    monkeypatch.setenv("NB_EXACT_PARALLEL", "0")
    g_ser = run_gpu(fresh, s, 0.5, 0.01, 3)
    ref = run_ora(oracle, s, 0.5, 0.01, 3, nthreads=4)
    assert np.array_equal(bits(g_par), bits(ref)) and np.array_equal(bits(g_ser), bits(ref))


def test_exact_merge_and_coincident_bodies(fresh, oracle):
    s = ic.random_disk(500, seed=8)
    s[1, :2] = s[0, :2] + f32(3e-5)   # too close: merged leaf (rs-src/nbody.rs:249-260)
    s[3, :2] = s[2, :2]               # identical positions
    s[5, 0] = s[4, 0] + f32(5e-5)     # close in x only: not merged
    fresh.set_mode(binding.MODE_EXACT)
    g = run_gpu(fresh, s, 0.5, 0.01, 4)
    r = run_ora(oracle, s, 0.5, 0.01, 4)
    assert np.array_equal(bits(g), bits(r))


def test_kill_hack_threshold_and_thread_hint(fresh, oracle):
    s = np.array([[56.0, 0, 1.0, 1.0, 1.0], [54.0, 0, 1.0, 1.0, 1.0], [0, 0, 0, 0, 1.0], [0, 55.5, 0.5, 0, 1.0]], dtype=f32)
    for mode in (binding.MODE_EXACT, binding.MODE_FAST):
        fresh.set_mode(mode)
        q = run_gpu(fresh, s, 0.5, 1e-3, 1)
        assert q[0, 2] == 0 and q[0, 3] == 0 and q[3, 2] == 0 and q[3, 3] == 0   # KAT-5
        assert q[1, 2] != 0 and q[1, 3] != 0
    fresh.set_mode(binding.MODE_EXACT)
    d = ic.random_disk(700, seed=5)
    outs = [bits(run_gpu(fresh, d, 0.85, 0.01, 3, nthreads=nt)) for nt in (1, 2, 3, 7, 16)]   # KAT-6
    assert all(np.array_equal(outs[0], o) for o in outs[1:])
    assert np.array_equal(outs[0], bits(run_ora(oracle, d, 0.85, 0.01, 3, nthreads=7)))


# ------------------------------------------------------------------ FAST ----------------------------
@pytest.mark.parametrize("n,theta,gen", [(4096, 0.5, "disk"), (16384, 0.85, "orbits"), (65536, 0.5, "plummer"),
                                         (50000, 0.75, "disk")])
def test_fast_topology_and_interaction_lists_match_reference(fresh, oracle, n, theta, gen):
    s = ic.stable_orbits(n, 0.5, 30.0, seed=4) if gen == "orbits" else \
        (ic.random_disk(n, seed=4) if gen == "disk" else ic.plummer_2d(n, seed=4))
    fresh.bh_count_interactions(True)
    fresh.set_particles(s)
    fresh.reset_counters()
    a = fresh.bh_accelerations(theta).astype(np.float64)
    c = fresh.counters()
    oracle.set_particles(s)
    oracle.bh_build()
    inter, visited = oracle.bh_count(theta)
    # same quadtree: the reference allocates 1 + 4 * (number of splits) nodes
    assert c["bh_nodes_built"] == oracle.bh_node_count()
    # same per-body walks, up to opening tests that flip on COM rounding (razor-edge cases only)
    assert abs(c["bh_interactions"] - inter) <= 2e-4 * inter
    assert abs(c["bh_nodes_visited"] - visited) <= 2e-4 * visited
    f = oracle.bh_forces_rows(theta, 0, n).astype(np.float64) / s[:, 4:5]
    err = np.abs(a - f).max(1) / np.abs(f).max()
    # rounding-level agreement for (almost) every body; the few bodies whose walk differs by a flipped
    # razor-edge opening test move by no more than the Barnes-Hut approximation error itself
    assert np.quantile(err, 0.999) <= 5e-5
    assert err.max() <= 2e-3


@pytest.mark.parametrize("n,steps,theta,gen", [(16384, 30, 0.5, "plummer"), (10000, 20, 0.85, "plummer"),
                                               (3000, 10, 0.5, "disk")])
def test_fast_positions_within_tolerance(fresh, oracle, n, steps, theta, gen):
    s = ic.plummer_2d(n, seed=6) if gen == "plummer" else ic.random_disk(n, seed=6)
    g = run_gpu(fresh, s, theta, 0.01, steps)
    r = run_ora(oracle, s, theta, 0.01, steps)
    ext = np.abs(r[:, :2]).max()
    assert np.abs(g[:, :2].astype(np.float64) - r[:, :2]).max() / ext <= 1e-4
    assert np.abs(g[:, 2:4].astype(np.float64) - r[:, 2:4]).max() / np.abs(r[:, 2:4]).max() <= 1e-3


def test_fast_force_error_vs_brute_force_not_worse_than_reference(fresh, oracle):
    n, theta = 65536, 0.5
    s = ic.random_disk(n, seed=7)
    fresh.set_particles(s)
    a_bh = fresh.bh_accelerations(theta).astype(np.float64)
    rows = np.random.default_rng(2).choice(n, 1024, replace=False).astype(np.int32)
    oracle.set_particles(s)
    a64 = oracle.accel_f64_rows(rows)
    oracle.bh_build()
    f_ref = oracle.bh_forces_rows(theta, 0, n).astype(np.float64) / s[:, 4:5]
    scale = np.abs(a64).max()
    e_gpu = np.sqrt(((a_bh[rows] - a64) ** 2).sum(1)).mean() / scale
    e_ref = np.sqrt(((f_ref[rows] - a64) ** 2).sum(1)).mean() / scale
    assert e_gpu <= 1.1 * e_ref


def test_c4_size_262144_theta05_one_step(fresh, oracle):
    """configs[3]: 262,144 bodies, theta=0.5.  One full step against the oracle (the reference's own tree)."""
    n = 262144
    s = ic.random_disk(n, seed=4)
    g = run_gpu(fresh, s, 0.5, 0.01, 1)
    r = run_ora(oracle, s, 0.5, 0.01, 1, nthreads=os.cpu_count() or 1)
    ext = np.abs(r[:, :2]).max()
    ep = np.abs(g[:, :2].astype(np.float64) - r[:, :2]).max(1) / ext
    ev = np.abs(g[:, 2:4].astype(np.float64) - r[:, 2:4]).max(1) / np.abs(r[:, 2:4]).max()
    assert ep.max() <= 1e-4 and ev.max() <= 1e-3          # the stated FP32 tolerance, every body
    assert np.quantile(ep, 0.999) <= 1e-6 and np.quantile(ev, 0.999) <= 1e-5   # rounding level for all but flipped walks


def test_c4_size_262144_ten_steps_fast(fresh, oracle):
    """configs[3], K=10 through the graph-replayed FAST step: every body within the stated tolerance of the oracle."""
    n = 262144
    s = ic.plummer_2d(n, seed=4)
    g = run_gpu(fresh, s, 0.5, 0.01, 10)
    r = run_ora(oracle, s, 0.5, 0.01, 10, nthreads=os.cpu_count() or 1)
    ext = np.abs(r[:, :2]).max()
    assert np.abs(g[:, :2].astype(np.float64) - r[:, :2]).max() / ext <= 1e-4
    assert np.abs(g[:, 2:4].astype(np.float64) - r[:, 2:4]).max() / np.abs(r[:, 2:4]).max() <= 1e-3


@pytest.mark.parametrize("gen", ["disk", "plummer"])
def test_c5_size_4m_theta075_vs_oracle(fresh, oracle, gen):
    """configs[4] (4,194,304 bodies, theta=0.75) against the ORACLE, not against itself, for the single tree and for the
    domain-partitioned path (8 virtual parts -- the multi-GPU code path); uniform disk (the benchmark set) and a
    Plummer model (clustered: deep tree, thousands of EPS-merged pairs).

    At this size the reference's own f32 arithmetic is the noise floor: its node masses / COMs are f32 running sums over
    up to 4M insertions, so ITS forces sit ~3e-5 (median) from what the FAST tree (f64 moment sums, rounded once) gives,
    and with dt = 0.01 one step displaces a body by dt^2|a| ~ 9 units -- the size of the system -- so a relative force
    difference appears undiminished as a relative position difference.  Hence: (1) same tree (node count) and same
    per-body interaction lists (counts); (2) accelerations within the f32 noise of the reference for every body whose
    walk did not flip a razor-edge opening test, and within the Barnes-Hut error for those; (3) against an f64
    brute-force evaluation FAST is at least as accurate as the reference; (4) the stated 1e-4 position tolerance holds
    for EVERY body at dt = 0.001 and for > 90 % at dt = 0.01.  (EXACT mode reproduces the oracle's step at this size
    bit for bit: test_exact_c5_size_4m_one_step_bitwise.)"""
    n, theta = 1 << 22, 0.75
    s = ic.random_disk(n, seed=5) if gen == "disk" else ic.plummer_2d(n, seed=5)
    nc = os.cpu_count() or 1
    oracle.set_particles(s)
    oracle.bh_build()
    nodes = oracle.bh_node_count()
    inter, vis = oracle.bh_count(theta)
    f = oracle.bh_forces_rows(theta, 0, n, nthreads=nc).astype(np.float64) / s[:, 4:5]
    fn = np.maximum(np.sqrt((f ** 2).sum(1)), 1e-30)
    rows = np.random.default_rng(2).choice(n, 192, replace=False).astype(np.int32)
    a64 = oracle.accel_f64_rows(rows)
    ref_step = {}
    for dt in (0.01, 0.001):
        oracle.set_particles(s)
        oracle.step_barnes_hut(theta, dt, nc)
        ref_step[dt] = oracle.get_particles()
    fresh.bh_count_interactions(True)
    accs = []
    for parts in (1, 8):
        fresh.bh_partition(parts)
        fresh.set_particles(s)
        fresh.reset_counters()
        a = fresh.bh_accelerations(theta).astype(np.float64)
        c = fresh.counters()
        if gen == "disk":
            assert c["bh_nodes_built"] == nodes
        else:   # merged leaves follow the reference's insertion-order rule only approximately
            assert abs(c["bh_nodes_built"] - nodes) <= 1e-4 * nodes
        assert abs(c["bh_interactions"] - inter) <= 2e-4 * inter
        assert abs(c["bh_nodes_visited"] - vis) <= 2e-4 * vis
        assert 0.05 < c["bh_pop_lanes"] / (32.0 * c["bh_pops"]) <= 1.0     # lane-efficiency counters are live
        rel = np.sqrt(((a - f) ** 2).sum(1)) / fn
        # disk (unequal masses): the reference's f32 tree sums sit ~3e-5 from the f64-summed ones.  Plummer (4M EQUAL
        # masses of 0.01): its f32 running sums stall -- 41,943 + 0.01 rounds in steps of 0.0039 -- and its forces are
        # ~4e-3 from the f64 brute force (measured, profiles/r02_c5_parity_probe.jsonl), FAST ~1.2e-3: the distance
        # between the two is the reference's own error, and check (3) below is what holds FAST to the truth.
        med, q999, mx = (1e-4, 1e-3, 5e-2) if gen == "disk" else (5e-3, 3e-2, 1e-1)
        assert np.median(rel) <= med and np.quantile(rel, 0.999) <= q999 and rel.max() <= mx
        e_gpu = (np.sqrt(((a[rows] - a64) ** 2).sum(1)) / np.sqrt((a64 ** 2).sum(1))).mean()
        e_ref = (np.sqrt(((f[rows] - a64) ** 2).sum(1)) / np.sqrt((a64 ** 2).sum(1))).mean()
        assert e_gpu <= 1.02 * e_ref
        accs.append(a)
        for dt in (0.01, 0.001):
            fresh.set_particles(s)
            fresh.step_barnes_hut(theta, dt, 1)
            g, r = fresh.get_particles(), ref_step[dt]
            ep = np.abs(g[:, :2].astype(np.float64) - r[:, :2]).max(1) / np.abs(r[:, :2]).max()
            ev = np.abs(g[:, 2:4].astype(np.float64) - r[:, 2:4]).max(1) / np.abs(r[:, 2:4]).max()
            if dt == 0.001 or gen == "plummer":
                assert ep.max() <= 1e-4 and np.quantile(ev, 0.999) <= 1e-3
            else:
                assert (ep <= 1e-4).mean() >= 0.9 and np.quantile(ep, 0.999) <= 1e-3 and np.quantile(ev, 0.999) <= 1e-3
    # the partitioned walk gives the single-tree forces to rounding
    err = np.abs(accs[1] - accs[0]).max(1) / np.abs(accs[0]).max()
    assert np.quantile(err, 0.999) <= 1e-6 and err.max() <= 5e-3
    # run-to-run determinism of the (graph-replayed) production step, and counting build == production build
    fresh.bh_count_interactions(False)
    fresh.bh_partition(0)
    fresh.set_particles(s)
    b1 = fresh.bh_accelerations(theta)
    b2 = fresh.bh_accelerations(theta)
    assert np.array_equal(bits(b1), bits(b2)) and np.array_equal(bits(b1), bits(accs[0].astype(np.float32)))


def test_exact_c5_size_4m_one_step_bitwise(fresh, oracle):
    """configs[4] size in EXACT mode: one full step of 4,194,304 bodies at theta = 0.75 equals the oracle's bit for bit
    (serial device restatement of Node::insert: ~20 s on the GPU -- the semantics pin, not a throughput path)."""
    n = 1 << 22
    s = ic.random_disk(n, seed=5)
    fresh.set_mode(binding.MODE_EXACT)
    g = run_gpu(fresh, s, 0.75, 0.01, 1)
    r = run_ora(oracle, s, 0.75, 0.01, 1, nthreads=os.cpu_count() or 1)
    assert np.array_equal(bits(g), bits(r))


def test_nthreads_zero_or_negative_moves_nothing(fresh, oracle):
    """rs-src/nbody.rs:424-428: the per-thread closure (which holds the division by nthreads) is mapped over the
    empty range 0..nthreads, so the reference builds the tree and returns: no body moves, nothing aborts."""
    s = ic.random_disk(3000, seed=12)
    for nt in (0, -3):
        fresh.set_particles(s)
        fresh.step_barnes_hut(0.5, 0.01, nt)
        assert np.array_equal(bits(fresh.get_particles()), bits(s))
        oracle.set_particles(s)
        oracle.step_barnes_hut(0.5, 0.01, nt)
        assert np.array_equal(bits(oracle.get_particles()), bits(s))
    fresh.set_particles(s)
    fresh.step_barnes_hut(0.0, 0.01, 0)          # theta == 0 returns through brute force before nthreads is looked at
    oracle.set_particles(s)
    oracle.step_barnes_hut(0.0, 0.01, 0)
    r = oracle.get_particles()
    assert np.abs(fresh.get_particles()[:, :2].astype(np.float64) - r[:, :2]).max() / np.abs(r[:, :2]).max() <= 1e-4


def test_graph_replay_survives_reallocation_and_parameter_changes(fresh, oracle):
    """The captured Barnes-Hut step must never replay with stale pointers or stale theta / dt: grow the set inside the
    same arena (workspace reallocated), shrink it back, change theta and dt between steps (the reference UI does, by
    key press, hs-src/RustNBodyExperiment.hs:88-93) -- every state must match a fresh oracle run."""
    def check(s, plan):
        fresh.set_particles(s)
        oracle.set_particles(s)
        for theta, dt in plan:
            fresh.step_barnes_hut(theta, dt, 1)
            oracle.step_barnes_hut(theta, dt, 2)
        g, r = fresh.get_particles(), oracle.get_particles()
        assert np.abs(g[:, :2].astype(np.float64) - r[:, :2]).max() / np.abs(r[:, :2]).max() <= 1e-4

    a, b = ic.random_disk(1000, seed=1), ic.random_disk(1020, seed=2)
    check(a, [(0.5, 0.01)] * 2)            # both graph slots captured at n = 1000
    check(b, [(0.5, 0.01)])                # same 1024-slot arena, workspace reallocated, slot 0 re-captured
    check(a, [(0.5, 0.01)] * 3)            # slot 1 must not replay the freed buffers
    check(a, [(0.5, 0.01), (0.85, 0.01), (0.85, 0.005), (0.3, 0.02), (0.5, 0.01)] * 5)   # 25 key presses: no kill switch
    check(ic.random_disk(70000, seed=3), [(0.6, 0.01)] * 3)
    check(a, [(0.5, 0.01)] * 3)


def test_steps_are_synchronous_by_default_and_async_on_request(fresh):
    """The reference's host times the step call with a wall clock (hs-src/RustNBodyExperiment.hs:55-57): by default
    the call returns when the GPU is done; nbx_set_async(1) returns after enqueueing."""
    import time

    s = ic.random_disk(262144, seed=3)
    fresh.set_particles(s)
    fresh.step_brute_force(0.01)
    fresh.synchronize()
    t = time.perf_counter(); fresh.step_brute_force(0.01); t_sync = time.perf_counter() - t
    fresh.set_async(True)
    t = time.perf_counter(); fresh.step_brute_force(0.01); t_async = time.perf_counter() - t
    fresh.synchronize()
    fresh.set_async(False)
    assert t_sync > 5e-3          # a 262,144-body all-pairs step takes ~17 ms on a B200
    assert t_async < 0.5 * t_sync


def test_get_particles_local_single_gpu_equals_get_particles(fresh):
    s = ic.random_disk(5000, seed=9)
    fresh.set_particles(s)
    fresh.step_barnes_hut(0.5, 0.01, 1)
    out = np.full((5000, 5), np.nan, dtype=np.float32)
    fresh.get_particles_local(out)
    assert np.array_equal(bits(out), bits(fresh.get_particles()))


@pytest.mark.parametrize("n,gen,parts", [(4096, "disk", 1), (65536, "plummer", 1), (50000, "orbits", 1), (65536, "disk", 4),
                                         (100000, "plummer", 8)])
def test_fast_tree_equals_reference_tree_node_by_node(fresh, oracle, n, gen, parts, monkeypatch):
    """The FAST (sort-based) tree dumped in the oracle's DFS format: same nodes in the same order, cell boxes bit
    for bit, leaves bit for bit, interior mass/COM to rounding (f64 prefix sums vs the f32 running mean)."""
    monkeypatch.setenv("NB_BH_PARTS_MIN_N", "0")
    s = ic.stable_orbits(n, 0.5, 30.0, seed=4) if gen == "orbits" else \
        (ic.random_disk(n, seed=4) if gen == "disk" else ic.plummer_2d(n, seed=4))
    fresh.bh_partition(parts)
    fresh.set_particles(s)
    fresh.bh_accelerations(0.5)
    t = fresh.bh_flatten()
    oracle.set_particles(s)
    oracle.bh_build()
    r = oracle.bh_flatten()
    assert t.shape == r.shape
    assert np.array_equal(t[:, 7:9], r[:, 7:9])                       # same topology, same DFS order
    assert np.array_equal(bits(t[:, 0:4]), bits(r[:, 0:4]))           # cell boxes bit for bit
    leaf = r[:, 7] == 0
    single = leaf & (r[:, 6] > 0)
    # leaves holding one body are exact copies; merged leaves (rare) agree to rounding
    exact = np.all(bits(t[single][:, 4:7]) == bits(r[single][:, 4:7]), axis=1)
    assert exact.mean() >= 0.999
    # Interior nodes: the reference's mass / COM are f32 running sums over up to n insertions.  For equal masses
    # the rounding of every `m += dm` has the same sign within a binade, so the error grows LINEARLY (measured:
    # ~1e-3 relative at the root of a 65k-body equal-mass set); the FAST tree's are f64 prefix sums rounded once.
    ext = np.abs(s[:, :2]).max()
    assert np.abs(t[:, 4:6].astype(np.float64) - r[:, 4:6]).max() / ext <= 2e-3
    assert np.abs(t[:, 6].astype(np.float64) - r[:, 6]).max() / r[:, 6].max() <= 5e-3
    # ...and against an f64 evaluation the FAST root is the more accurate of the two
    m64, x64 = s[:, 4].astype(np.float64), s[:, 0].astype(np.float64)
    com_x, mass = (m64 * x64).sum() / m64.sum(), m64.sum()
    assert abs(t[0, 4] - com_x) <= abs(r[0, 4] - com_x) + 1e-6 * ext
    assert abs(t[0, 6] - mass) <= abs(r[0, 6] - mass) + 1e-6 * mass
    assert abs(t[0, 6] - mass) <= 1e-6 * mass


# ------------------------------------------------------------------ domain-partitioned build -----------
@pytest.mark.parametrize("n,theta,gen,parts", [(65536, 0.5, "disk", 2), (65536, 0.75, "plummer", 8), (100000, 0.5, "orbits", 4),
                                               (262144, 0.5, "disk", 8)])
def test_partitioned_tree_equals_single_tree(fresh, oracle, n, theta, gen, parts, monkeypatch):
    """The multi-GPU code path (cell-aligned Morton partition, per-part subtree forests, shared top tree,
    part-tagged block indices) run with `parts` virtual ranks on one GPU: same quadtree as the reference
    (node count), same interaction lists (counts), same forces as the single-tree walk to rounding."""
    monkeypatch.setenv("NB_BH_PARTS_MIN_N", "0")
    s = ic.stable_orbits(n, 0.5, 30.0, seed=4) if gen == "orbits" else \
        (ic.random_disk(n, seed=4) if gen == "disk" else ic.plummer_2d(n, seed=4))
    fresh.bh_count_interactions(True)
    fresh.set_particles(s)
    fresh.bh_partition(1)
    fresh.reset_counters()
    a1 = fresh.bh_accelerations(theta).astype(np.float64)
    c1 = fresh.counters()
    fresh.bh_partition(parts)
    fresh.reset_counters()
    a2 = fresh.bh_accelerations(theta).astype(np.float64)
    c2 = fresh.counters()
    oracle.set_particles(s)
    oracle.bh_build()
    assert c1["bh_nodes_built"] == oracle.bh_node_count()
    assert c2["bh_nodes_built"] == oracle.bh_node_count()
    assert abs(c2["bh_interactions"] - c1["bh_interactions"]) <= 1e-5 * c1["bh_interactions"]
    assert abs(c2["bh_nodes_visited"] - c1["bh_nodes_visited"]) <= 1e-5 * c1["bh_nodes_visited"]
    err = np.abs(a2 - a1).max(1) / np.abs(a1).max()
    assert np.quantile(err, 0.999) <= 1e-6 and err.max() <= 2e-3
    # and a few steps through the partitioned path stay within tolerance of the oracle
    g = run_gpu(fresh, s, theta, 0.01, 3)
    r = run_ora(oracle, s, theta, 0.01, 3, nthreads=os.cpu_count() or 1)
    assert np.abs(g[:, :2].astype(np.float64) - r[:, :2]).max() / np.abs(r[:, :2]).max() <= 1e-4


@pytest.mark.parametrize("case", ["coincident", "collinear_x", "collinear_y", "two_clusters", "tiny", "huge_extent", "heavy_sun"])
def test_fast_degenerate_inputs_match_reference_tree(fresh, oracle, case):
    """Edge cases of the tree build (rs-src/nbody.rs:249-301): merged leaves, zero-height boxes, deep trees."""
    rng = np.random.default_rng(17)
    n = 3000
    s = ic.random_disk(n, seed=17)
    if case == "coincident":
        s[:, 0] = 1.25; s[:, 1] = -3.5                      # everything merges into the root leaf
    elif case == "collinear_x":
        s[:, 1] = 2.0                                       # zero-height bounding box
    elif case == "collinear_y":
        s[:, 0] = -7.0
    elif case == "two_clusters":
        s[: n // 2, :2] = rng.normal(0, 1e-3, (n // 2, 2)).astype(f32) + f32(10.0)
        s[n // 2:, :2] = rng.normal(0, 1e-3, (n - n // 2, 2)).astype(f32) - f32(10.0)
    elif case == "tiny":
        s = s[:3].copy()
    elif case == "huge_extent":
        s[0, :2] = (4000.0, -4000.0)                        # one escaper: root cell / 2^24 > EPS
    elif case == "heavy_sun":
        s = ic.stable_orbits(n, 0.5, 30.0, seed=17)
    fresh.bh_count_interactions(True)
    fresh.set_particles(s)
    fresh.reset_counters()
    a = fresh.bh_accelerations(0.5).astype(np.float64)
    c = fresh.counters()
    oracle.set_particles(s)
    oracle.bh_build()
    f = oracle.bh_forces_rows(0.5, 0, s.shape[0]).astype(np.float64) / s[:, 4:5]
    assert np.isfinite(a).all()
    # Sets with many bodies closer than EPS (a line of 3000 bodies, two 1e-3-wide clusters) exercise the
    # reference's insertion-order-dependent merge rule (a merged blob travels as a unit, rs-src/nbody.rs:261-282),
    # which a sort-based build can only approximate: there the trees may differ by a few nodes (documented in
    # DESIGN.md 4.3) and forces are compared more loosely.
    merge_heavy = case in ("collinear_x", "collinear_y", "two_clusters")
    if merge_heavy:
        assert abs(c["bh_nodes_built"] - oracle.bh_node_count()) <= 0.005 * oracle.bh_node_count() + 8
    else:
        assert c["bh_nodes_built"] == oracle.bh_node_count()
    scale = max(np.abs(f).max(), 1e-30)
    err = np.abs(a - f).max(1) / scale
    assert np.quantile(err, 0.99) <= (2e-3 if merge_heavy else 1e-4) and err.max() <= 2e-2


def test_theta_zero_goes_brute_force(fresh, oracle):
    s = ic.random_disk(1500, seed=9)
    fresh.set_mode(binding.MODE_EXACT)
    fresh.set_particles(s)
    fresh.step_barnes_hut(0.0, 0.01, 5)
    oracle.set_particles(s)
    oracle.step_brute_force(0.01)
    assert np.array_equal(bits(fresh.get_particles()), bits(oracle.get_particles()))


def test_experiment_plugin_drives_the_library(fresh):
    """The reference's plug-in sequence (hs-src/RustNBodyExperiment.hs): stable orbits 10000 -> step BH -> draw."""
    from rust_exp_b200 import RustNBodyExperiment

    e = RustNBodyExperiment(fresh)
    fb = np.zeros((240, 320), dtype=np.uint32)
    for _ in range(3):
        e.draw(fb)
    assert e.num_steps == 3 and fresh.num_particles() == 10000
    assert (fb != 0).sum() > 1000 and fb[120, 160] == 0x00FF00FF
    assert "10K Bodies" in e.status_string()
