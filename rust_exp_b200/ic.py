"""Seeded initial conditions for the benchmark configurations.

The reference seeds from an unseeded thread_rng (rs-src/nbody.rs:46,90), so "identical initial
conditions" can only mean: generate once here, inject the same bytes into the oracle and the GPU library
through nb_set_particles.  All generators return an (N,5) float32 AoS array {px,py,vx,vy,m}.
"""
from __future__ import annotations

import numpy as np


def stable_orbits(n: int, rmin: float = 0.5, rmax: float = 30.0, seed: int = 1) -> np.ndarray:
    """Restatement of nb_stable_orbits (rs-src/nbody.rs:73-104): sun + n-1 planets, speed sqrt(1000)."""
    rng = np.random.default_rng(seed)
    a = np.zeros((n, 5), dtype=np.float32)
    a[0] = (0, 0, 0, 0, 1000.0)
    r = (rmax - rmin) * rng.random(n - 1) + rmin
    th = 2 * np.pi * rng.random(n - 1)
    speed = np.sqrt(1000.0)
    a[1:, 0] = r * np.cos(th)
    a[1:, 1] = r * np.sin(th)
    a[1:, 2] = -speed * np.sin(th)
    a[1:, 3] = speed * np.cos(th)
    a[1:, 4] = 1.0
    return a


def random_disk(n: int, seed: int = 1) -> np.ndarray:
    """Restatement of nb_random_disk (rs-src/nbody.rs:39-71): uniform disk r<=23, v in [-3.5,3.5), m in [0.1,1.5)."""
    rng = np.random.default_rng(seed)
    a = np.zeros((n, 5), dtype=np.float32)
    r = 23.0 * np.sqrt(rng.random(n))
    th = 2 * np.pi * rng.random(n)
    a[:, 0] = r * np.cos(th)
    a[:, 1] = r * np.sin(th)
    a[:, 2] = rng.uniform(-3.5, 3.5, n)
    a[:, 3] = rng.uniform(-3.5, 3.5, n)
    a[:, 4] = rng.uniform(0.1, 1.5, n)
    return a


def plummer_2d(n: int, seed: int = 2, a_scale: float = 5.0, total_mass_per_body: float = 1e-2) -> np.ndarray:
    """Plummer sphere (scale a=5, truncated at 10a) projected to z=0, equal masses.

    Velocities are isotropic with the local Plummer dispersion scaled for this 2-D pair law's units; they
    only need to be a fixed, reproducible smooth field for throughput and parity purposes.
    """
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 5), dtype=np.float32)
    # radius from the cumulative mass profile M(r)/M = r^3/(r^2+a^2)^(3/2), truncated at 10a
    xmax = (10.0**3) / (101.0**1.5)
    u = rng.random(n) * xmax
    r = a_scale / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    cz = rng.uniform(-1, 1, n)
    ph = rng.uniform(0, 2 * np.pi, n)
    s = np.sqrt(1 - cz * cz)
    out[:, 0] = r * s * np.cos(ph)
    out[:, 1] = r * s * np.sin(ph)
    m = total_mass_per_body
    sigma = np.sqrt(n * m / 6.0) * (1 + (r / a_scale) ** 2) ** (-0.25)
    out[:, 2] = rng.normal(0, 1, n) * sigma
    out[:, 3] = rng.normal(0, 1, n) * sigma
    out[:, 4] = m
    return out


def plummer_3d(n: int, seed: int = 2, a_scale: float = 5.0, mass_per_body: float = 1e-2) -> np.ndarray:
    """Plummer sphere for the 3-D extension: (N,7) float32 {px,py,pz,vx,vy,vz,m}, truncated at 10a."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n, 7), dtype=np.float32)
    xmax = (10.0**3) / (101.0**1.5)
    u = rng.random(n) * xmax
    r = a_scale / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    cz = rng.uniform(-1, 1, n)
    ph = rng.uniform(0, 2 * np.pi, n)
    s = np.sqrt(1 - cz * cz)
    out[:, 0] = r * s * np.cos(ph)
    out[:, 1] = r * s * np.sin(ph)
    out[:, 2] = r * cz
    sigma = np.sqrt(n * mass_per_body / (6.0 * a_scale)) * (1 + (r / a_scale) ** 2) ** (-0.25)
    out[:, 3:6] = rng.normal(0, 1, (n, 3)) * sigma[:, None]
    out[:, 6] = mass_per_body
    return out
