"""GPU: the 3-D extension surface nbx3_* (LAW_NEWTON: rsqrt-based; LAW_REF: the reference law in 3-D).

No reference counterpart exists for 3-D (rs-src/nbody.rs is 2-D), so: LAW_REF with z == 0 must reproduce the
2-D FAST kernel bit for bit (which is itself parity-tested against the oracle), and both laws are checked
against an f64 evaluation of the same law (oracle/nbody_oracle.c: ora3_*)."""
import numpy as np
import pytest

from rust_exp_b200 import binding, ic

pytestmark = pytest.mark.gpu
f32 = np.float32


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("n", [1000, 4096, 20000])
def test_ref_law_with_z0_equals_2d_fast_kernel_bitwise(fresh, n):
    s2 = ic.random_disk(n, seed=5)
    s3 = np.zeros((n, 7), dtype=f32)
    s3[:, 0:2] = s2[:, 0:2]
    s3[:, 3:5] = s2[:, 2:4]
    s3[:, 6] = s2[:, 4]
    fresh.configure3(binding.LAW3_REF, 1e-4)
    fresh.set_particles3(s3)
    fresh.set_particles(s2)
    a3 = fresh.accelerations3()
    a2 = fresh.accelerations()
    assert np.array_equal(bits(a3[:, :2]), bits(a2)) and (a3[:, 2] == 0).all()
    for _ in range(5):
        fresh.step3(0.01)
        fresh.step_brute_force(0.01)
    g3, g2 = fresh.get_particles3(), fresh.get_particles()
    assert np.array_equal(bits(g3[:, [0, 1, 3, 4, 6]]), bits(g2)) and (g3[:, 2] == 0).all() and (g3[:, 5] == 0).all()


@pytest.mark.parametrize("law", [binding.LAW3_NEWTON, binding.LAW3_REF])
@pytest.mark.parametrize("n", [1, 2, 777, 8192])
def test_accelerations_vs_f64(fresh, oracle, law, n):
    s = ic.plummer_3d(n, seed=7)
    fresh.configure3(law, 1e-4)
    fresh.set_particles3(s)
    a = fresh.accelerations3().astype(np.float64)
    ref = oracle.accel3_f64(s, law, float(f32(1e-4)))
    scale = max(np.abs(ref).max(), 1e-30)
    assert np.abs(a - ref).max() / scale <= 2e-5
    m = s[:, 6:7].astype(np.float64)
    assert np.abs((m * a).sum(0)).max() <= 1e-5 * max(np.abs(m * a).sum(), 1e-30)   # antisymmetry


def test_newton_steps_vs_f64(fresh, oracle):
    n, steps = 4096, 20
    s = ic.plummer_3d(n, seed=8)
    fresh.configure3(binding.LAW3_NEWTON, 1e-4)
    fresh.set_particles3(s)
    t = s.astype(np.float64)
    for _ in range(steps):
        fresh.step3(0.01)
        t = oracle.step3_f64(t, binding.LAW3_NEWTON, float(f32(1e-4)), float(f32(0.01)))
    g = fresh.get_particles3()
    ext = np.abs(t[:, :3]).max()
    assert np.abs(g[:, :3] - t[:, :3]).max() / ext <= 1e-4
    assert np.abs(g[:, 3:6] - t[:, 3:6]).max() / np.abs(t[:, 3:6]).max() <= 1e-3
    assert np.array_equal(g[:, 6], s[:, 6])


def test_newton_65536_sampled_rows(fresh, oracle):
    n = 65536
    s = ic.plummer_3d(n, seed=2)
    fresh.configure3(binding.LAW3_NEWTON, 1e-4)
    fresh.set_particles3(s)
    a = fresh.accelerations3().astype(np.float64)
    rows = np.random.default_rng(3).choice(n, 512, replace=False).astype(np.int32)
    ref = oracle.accel3_f64(s, 0, float(f32(1e-4)), rows)
    assert np.abs(a[rows] - ref).max() / np.abs(ref).max() <= 2e-5


def test_empty_and_roundtrip(fresh):
    fresh.set_particles3(np.zeros((0, 7), dtype=f32))
    fresh.step3(0.01)
    assert fresh.get_particles3().shape == (0, 7)
    s = ic.plummer_3d(3000, seed=9)
    fresh.set_particles3(s)
    assert np.array_equal(bits(fresh.get_particles3()), bits(s))
