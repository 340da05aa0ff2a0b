"""GPU: nb_draw (device scatter + one contiguous copy) is pixel-identical to the oracle's restatement of
rs-src/nbody.rs:482-617."""
import os

import numpy as np
import pytest

from rust_exp_b200 import ic

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "nbody_golden.npz"))


def test_draw_golden(fresh):
    fresh.set_particles(GOLD["disk"])
    assert np.array_equal(fresh.draw(96, 64), GOLD["draw_disk_96x64"])


@pytest.mark.parametrize("n,w,h", [(0, 64, 48), (1, 100, 100), (10000, 512, 512), (50000, 320, 200), (200000, 256, 256)])
def test_draw_matches_oracle(fresh, oracle, n, w, h):
    s = ic.random_disk(n, seed=3) if n != 1 else np.array([[10.0, 0, 1.0, 0, 1.0]], dtype=np.float32)
    if n > 1:
        s[:50, :2] *= 3.0  # some bodies outside the viewport -> bounds check
    fresh.set_particles(s)
    oracle.set_particles(s)
    a, b = fresh.draw(w, h), oracle.draw(w, h)
    assert np.array_equal(a, b)


def test_draw_saturates(fresh, oracle):
    s = np.zeros((4000, 5), dtype=np.float32)
    s[:, 0] = 3.0
    s[:, 1] = 4.0
    s[:, 2] = 1.0
    s[:, 4] = 1.0
    fresh.set_particles(s)
    oracle.set_particles(s)
    a = fresh.draw(128, 128)
    assert np.array_equal(a, oracle.draw(128, 128))
    assert (a == 0x00A6FFFF).any() or (a & 0xFF).max() == 255


def test_draw_tail_octants_and_boundaries(fresh, oracle):
    """Every octant plus velocities exactly on the octant boundaries (multiples of 45 degrees) and zero velocity."""
    v = [(1, 0), (1, 1), (0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1), (1, -1), (0, 0), (1e-30, -1e-30), (3, 3.0000002),
         (1.0, 0.1), (-0.1, 1.0), (-1.0, -0.1), (0.1, -1.0), (2.5, -2.5), (-7, 7)]
    s = np.zeros((len(v), 5), dtype=np.float32)
    s[:, 0] = np.linspace(-40, 40, len(v))
    s[:, 1] = np.linspace(-30, 30, len(v))
    s[:, 2:4] = np.array(v, dtype=np.float32)
    s[:, 4] = 1.0
    fresh.set_particles(s)
    oracle.set_particles(s)
    assert np.array_equal(fresh.draw(200, 150), oracle.draw(200, 150))
