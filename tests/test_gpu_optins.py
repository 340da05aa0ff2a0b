"""GPU: the two opt-ins taken from the reference author's own TODO list --
leapfrog kick-drift-kick instead of semi-implicit Euler (rs-src/nbody.rs:449) and a squared-up root box for the
quadtree (rs-src/nbody.rs:400-407).  Both change results, so both are off by default; they are checked against an f64
evaluation of the same scheme / against the oracle running the reference's commented-out variant."""
import numpy as np
import pytest

from rust_exp_b200 import binding, ic

pytestmark = pytest.mark.gpu
EPS = 1e-4


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def acc64(p, m):
    d = p[None, :, :] - p[:, None, :]
    r2 = (d ** 2).sum(-1) + EPS
    return (m[None, :, None] * d / r2[:, :, None]).sum(1)       # self term is exactly 0 (d = 0)


def kdk64(s, dt, steps):
    p, v, m = s[:, :2].astype(np.float64), s[:, 2:4].astype(np.float64), s[:, 4].astype(np.float64)
    a = acc64(p, m)
    for _ in range(steps):
        v = v + 0.5 * dt * a
        p = p + dt * v
        a = acc64(p, m)
        v = v + 0.5 * dt * a
    return p, v


def energy(p, v, m):
    d = p[None, :, :] - p[:, None, :]
    r2 = (d ** 2).sum(-1) + EPS
    u = 0.5 * np.triu(m[:, None] * m[None, :] * np.log(r2), 1).sum()    # F = -grad U for the reference's 2-D law
    return 0.5 * (m * (v ** 2).sum(1)).sum() + u


@pytest.fixture()
def kdk(fresh):
    fresh.set_integrator(binding.INTEGRATOR_LEAPFROG_KDK)
    yield fresh
    fresh.set_integrator(binding.INTEGRATOR_EULER)
    fresh.set_square_aabb(False)


def test_leapfrog_matches_f64_kick_drift_kick(kdk):
    s = ic.stable_orbits(600, 2.0, 30.0, seed=3)
    dt, steps = np.float32(0.002), 200
    kdk.set_particles(s)
    for k in range(steps):
        kdk.step_brute_force(dt)
        if k == 77:
            kdk.get_particles()      # a read-back in the middle applies the owed half kick: must not change the trajectory
    g = kdk.get_particles()
    p, v = kdk64(s, float(dt), steps)
    assert np.abs(g[:, :2] - p).max() / np.abs(p).max() <= 1e-4
    assert np.abs(g[:, 2:4] - v).max() / np.abs(v).max() <= 1e-3


def test_leapfrog_is_time_reversible_and_conserves_energy_better_than_euler(fresh):
    s = ic.stable_orbits(400, 3.0, 30.0, seed=5)
    dt, steps = np.float32(0.004), 150
    m = s[:, 4].astype(np.float64)
    e0 = energy(s[:, :2].astype(np.float64), s[:, 2:4].astype(np.float64), m)
    drift = {}
    for name, integ in (("euler", binding.INTEGRATOR_EULER), ("kdk", binding.INTEGRATOR_LEAPFROG_KDK)):
        fresh.set_integrator(integ)
        fresh.set_particles(s)
        for _ in range(steps):
            fresh.step_brute_force(dt)
        g = fresh.get_particles()
        drift[name] = abs(energy(g[:, :2].astype(np.float64), g[:, 2:4].astype(np.float64), m) - e0) / abs(e0)
        if name == "kdk":       # reverse the velocities and integrate back: leapfrog returns to the start, Euler does not
            g[:, 2:4] *= -1
            fresh.set_particles(g)
            for _ in range(steps):
                fresh.step_brute_force(dt)
            back = fresh.get_particles()
            assert np.abs(back[:, :2] - s[:, :2]).max() / np.abs(s[:, :2]).max() <= 2e-4
    fresh.set_integrator(binding.INTEGRATOR_EULER)
    assert drift["kdk"] < 0.25 * drift["euler"]


def test_leapfrog_through_barnes_hut_and_graph_replay(kdk, oracle):
    """KDK with the tree forces: theta tiny => the tree sum is the all-pairs sum, so the f64 KDK applies; the closing
    half kick at read-back uses the tree too.  Also exercises the first-step / later-step kick sizes in the graph."""
    s = ic.random_disk(700, seed=8)
    s[:, 2:4] *= 0.2
    dt, steps = np.float32(0.002), 40
    kdk.set_particles(s)
    for _ in range(steps):
        kdk.step_barnes_hut(1e-6, dt, 1)
    g = kdk.get_particles()
    p, v = kdk64(s, float(dt), steps)
    assert np.abs(g[:, :2] - p).max() / np.abs(p).max() <= 1e-4
    assert np.abs(g[:, 2:4] - v).max() / np.abs(v).max() <= 1e-3


def test_euler_default_is_untouched_by_the_optin_code(fresh, oracle):
    s = ic.random_disk(2000, seed=4)
    fresh.set_integrator(binding.INTEGRATOR_LEAPFROG_KDK)
    fresh.set_integrator(binding.INTEGRATOR_EULER)
    fresh.set_mode(binding.MODE_EXACT)
    fresh.set_particles(s)
    oracle.set_particles(s)
    for _ in range(5):
        fresh.step_brute_force(0.01)
        oracle.step_brute_force(0.01)
    assert np.array_equal(bits(fresh.get_particles()), bits(oracle.get_particles()))


@pytest.mark.parametrize("n,gen,parts", [(20000, "disk", 1), (65536, "plummer", 1), (65536, "disk", 4)])
def test_square_aabb_tree_equals_the_oracles_squared_tree(fresh, oracle, n, gen, parts, monkeypatch):
    """nbx_set_square_aabb(1): same cells as the oracle with the reference's commented-out squaring enabled, node by
    node; EXACT stepping bit-identical to it; and the root box really is square."""
    monkeypatch.setenv("NB_BH_PARTS_MIN_N", "0")
    s = ic.random_disk(n, seed=4) if gen == "disk" else ic.plummer_2d(n, seed=4)
    s[:, 1] *= 0.5                                   # a clearly non-square set
    oracle.set_square_aabb(True)
    try:
        fresh.set_square_aabb(True)
        fresh.bh_partition(parts)
        fresh.set_particles(s)
        fresh.bh_accelerations(0.5)
        t = fresh.bh_flatten()
        oracle.set_particles(s)
        oracle.bh_build()
        r = oracle.bh_flatten()
        assert t.shape == r.shape
        assert np.array_equal(t[:, 7:9], r[:, 7:9]) and np.array_equal(bits(t[:, 0:4]), bits(r[:, 0:4]))
        assert abs((t[0, 2] - t[0, 0]) - (t[0, 3] - t[0, 1])) <= 1e-5 * (t[0, 2] - t[0, 0])
        if parts == 1:
            fresh.set_mode(binding.MODE_EXACT)
            fresh.set_particles(s)
            oracle.set_particles(s)
            for _ in range(3):
                fresh.step_barnes_hut(0.5, 0.01, 1)
                oracle.step_barnes_hut(0.5, 0.01, 4)
            assert np.array_equal(bits(fresh.get_particles()), bits(oracle.get_particles()))
            fresh.set_mode(binding.MODE_FAST)
        # and the squared tree is a different tree from the default one
        oracle.set_square_aabb(False)
        oracle.set_particles(s)
        oracle.bh_build()
        assert oracle.bh_flatten().shape != r.shape or not np.array_equal(bits(oracle.bh_flatten()[:, :4]), bits(r[:, :4]))
    finally:
        oracle.set_square_aabb(False)
        fresh.set_square_aabb(False)
