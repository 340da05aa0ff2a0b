"""GPU: seeded randomized sweep of sizes / distributions / parameters -- EXACT mode must equal the oracle bit for
bit for both paths, FAST mode must stay within tolerance; sizes straddle every tile boundary of the kernels."""
import numpy as np
import pytest

from rust_exp_b200 import binding, ic

pytestmark = pytest.mark.gpu
f32 = np.float32
SIZES = [3, 31, 32, 33, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 1535, 2047, 2048, 2049, 3000, 4095, 4097]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def make(rng, n):
    kind = rng.integers(0, 4)
    if kind == 0:
        s = ic.random_disk(n, seed=int(rng.integers(1 << 30)))
    elif kind == 1:
        s = ic.stable_orbits(max(n, 2), 0.5, 30.0, seed=int(rng.integers(1 << 30)))[:n]
    elif kind == 2:
        s = ic.plummer_2d(n, seed=int(rng.integers(1 << 30)))
    else:  # clumpy: a few tight groups, some bodies closer than EPS
        s = ic.random_disk(n, seed=int(rng.integers(1 << 30)))
        k = max(1, n // 10)
        s[:k, :2] = s[0, :2] + rng.normal(0, 2e-4, (k, 2)).astype(f32)
    return s


@pytest.mark.parametrize("case", range(len(SIZES)))
def test_exact_mode_bitwise_random_cases(fresh, oracle, case):
    rng = np.random.default_rng(1000 + case)
    n = SIZES[case]
    s = make(rng, n)
    dt = float(rng.choice([0.01, 0.005, 0.02]))
    theta = float(rng.choice([0.3, 0.5, 0.85, 0.95]))
    steps = int(rng.integers(2, 5))
    fresh.set_mode(binding.MODE_EXACT)
    fresh.set_particles(s)
    oracle.set_particles(s)
    for k in range(steps):   # alternate the two paths on the same evolving state, like a user toggling theta
        if (k + case) % 2 == 0:
            fresh.step_brute_force(dt); oracle.step_brute_force(dt)
        else:
            fresh.step_barnes_hut(theta, dt, 1 + k); oracle.step_barnes_hut(theta, dt, 1 + k)
    assert np.array_equal(bits(fresh.get_particles()), bits(oracle.get_particles()))


@pytest.mark.parametrize("case", range(0, len(SIZES), 2))
def test_fast_mode_tolerance_random_cases(fresh, oracle, case):
    rng = np.random.default_rng(2000 + case)
    n = SIZES[case]
    s = ic.plummer_2d(n, seed=int(rng.integers(1 << 30)))   # smooth field: rounding is not chaotically amplified
    theta = float(rng.choice([0.3, 0.5, 0.85]))
    fresh.set_particles(s)
    oracle.set_particles(s)
    for k in range(6):
        if k % 2 == 0:
            fresh.step_brute_force(0.01); oracle.step_brute_force(0.01)
        else:
            fresh.step_barnes_hut(theta, 0.01, 1); oracle.step_barnes_hut(theta, 0.01, 2)
    g, r = fresh.get_particles(), oracle.get_particles()
    ext = max(np.abs(r[:, :2]).max(), 1e-30)
    assert np.abs(g[:, :2].astype(np.float64) - r[:, :2]).max() / ext <= 1e-4
    assert np.abs(g[:, 2:4].astype(np.float64) - r[:, 2:4]).max() / max(np.abs(r[:, 2:4]).max(), 1e-30) <= 1e-3
