"""ctypes loader for the CPU oracle (oracle/nbody_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (rust_exp_b200) never imports this.

Parity status: unpinned by reference vectors (the reference has none and cannot be built here);
see the header of nbody_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nbody_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "liboracle.so"])
    return _SO


class Oracle:
    """Thin typed wrapper over liboracle.so; one process-global particle set, like the reference."""

    def __init__(self) -> None:
        build()
        L = C.CDLL(_SO)
        f, i32, vp = C.c_float, C.c_int32, C.c_void_p
        L.ora_num_particles.restype = i32
        L.ora_set_particles.argtypes = [vp, i32]
        L.ora_get_particles.argtypes = [vp, i32]
        L.ora_seed.argtypes = [C.c_uint64]
        L.ora_random_disk.argtypes = [i32]
        L.ora_stable_orbits.argtypes = [i32, f, f]
        L.ora_force.argtypes = [f, f, f, f, f, f, vp]
        L.ora_brute_forces_rows.argtypes = [i32, i32, vp]
        L.ora_brute_forces_rows_mt.argtypes = [i32, i32, vp, i32]
        L.ora_step_brute_force.argtypes = [f]
        L.ora_step_brute_force_mt.argtypes = [f, i32]
        L.ora_step_barnes_hut.argtypes = [f, f, i32]
        L.ora_bh_build.argtypes = []
        L.ora_set_square_aabb.argtypes = [i32]
        L.ora_bh_node_count.restype = i32
        L.ora_bh_max_depth.restype = i32
        L.ora_bh_flatten.argtypes = [vp, i32]
        L.ora_bh_flatten.restype = i32
        L.ora_bh_forces_rows.argtypes = [f, i32, i32, vp]
        L.ora_bh_forces_rows_mt.argtypes = [f, i32, i32, vp, i32]
        L.ora_bh_count.argtypes = [f, vp, vp]
        L.ora_bh_count_mt.argtypes = [f, vp, vp, i32]
        L.ora_accel_f64_rows.argtypes = [vp, i32, vp]
        L.ora3_accel_f64.argtypes = [vp, i32, i32, C.c_double, vp, i32, vp]
        L.ora3_step_f64.argtypes = [vp, i32, i32, C.c_double, C.c_double]
        L.ora_draw.argtypes = [i32, i32, vp]
        L.ora_rgb_to_abgr32.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8, f]
        L.ora_rgb_to_abgr32.restype = C.c_uint32
        L.ora_add_abgr32.argtypes = [C.c_uint32, C.c_uint32]
        L.ora_add_abgr32.restype = C.c_uint32
        self.L = L

    # -- state ------------------------------------------------------------------------------
    def num_particles(self) -> int:
        return int(self.L.ora_num_particles())

    def set_particles(self, aos: np.ndarray) -> None:
        a = np.ascontiguousarray(aos, dtype=np.float32).reshape(-1, 5)
        self.L.ora_set_particles(a.ctypes.data, a.shape[0])

    def get_particles(self) -> np.ndarray:
        n = self.num_particles()
        out = np.empty((n, 5), dtype=np.float32)
        self.L.ora_get_particles(out.ctypes.data, n)
        return out

    def seed(self, s: int) -> None:
        self.L.ora_seed(s)

    def random_disk(self, n: int) -> None:
        self.L.ora_random_disk(n)

    def stable_orbits(self, n: int, rmin: float, rmax: float) -> None:
        self.L.ora_stable_orbits(n, rmin, rmax)

    # -- all pairs --------------------------------------------------------------------------
    def force(self, p1, m1, p2, m2) -> np.ndarray:
        out = np.empty(2, dtype=np.float32)
        self.L.ora_force(p1[0], p1[1], m1, p2[0], p2[1], m2, out.ctypes.data)
        return out

    def brute_forces_rows(self, i0: int, i1: int, nthreads: int = 1) -> np.ndarray:
        out = np.empty((i1 - i0, 2), dtype=np.float32)
        if nthreads == 1:
            self.L.ora_brute_forces_rows(i0, i1, out.ctypes.data)
        else:
            self.L.ora_brute_forces_rows_mt(i0, i1, out.ctypes.data, nthreads)
        return out

    def step_brute_force(self, dt: float, nthreads: int = 1) -> None:
        """nthreads > 1 splits the (independent) force rows over host threads: bit-identical result."""
        if nthreads <= 1:
            self.L.ora_step_brute_force(dt)
        else:
            self.L.ora_step_brute_force_mt(dt, nthreads)

    # -- Barnes-Hut -------------------------------------------------------------------------
    def step_barnes_hut(self, theta: float, dt: float, nthreads: int = 1) -> None:
        self.L.ora_step_barnes_hut(theta, dt, nthreads)

    def set_square_aabb(self, on: bool) -> None:
        """Opt-in, NOT reference behaviour: the commented-out squared-up root box of rs-src/nbody.rs:400-407."""
        self.L.ora_set_square_aabb(1 if on else 0)

    def bh_build(self) -> None:
        self.L.ora_bh_build()

    def bh_node_count(self) -> int:
        return int(self.L.ora_bh_node_count())

    def bh_max_depth(self) -> int:
        return int(self.L.ora_bh_max_depth())

    def bh_flatten(self) -> np.ndarray:
        n = self.bh_node_count()
        out = np.empty((n, 9), dtype=np.float32)
        got = self.L.ora_bh_flatten(out.ctypes.data, n)
        assert got == n, (got, n)
        return out

    def bh_forces_rows(self, theta: float, i0: int, i1: int, nthreads: int = 1) -> np.ndarray:
        out = np.empty((i1 - i0, 2), dtype=np.float32)
        if nthreads <= 1:
            self.L.ora_bh_forces_rows(theta, i0, i1, out.ctypes.data)
        else:
            self.L.ora_bh_forces_rows_mt(theta, i0, i1, out.ctypes.data, nthreads)
        return out

    def bh_count(self, theta: float, nthreads: int = 0) -> tuple[int, int]:
        a, b = C.c_uint64(0), C.c_uint64(0)
        self.L.ora_bh_count_mt(theta, C.byref(a), C.byref(b), nthreads or (os.cpu_count() or 1))
        return int(a.value), int(b.value)

    # -- f64 helper -------------------------------------------------------------------------
    def accel_f64_rows(self, rows: np.ndarray) -> np.ndarray:
        r = np.ascontiguousarray(rows, dtype=np.int32)
        out = np.empty((r.shape[0], 2), dtype=np.float64)
        self.L.ora_accel_f64_rows(r.ctypes.data, r.shape[0], out.ctypes.data)
        return out

    # -- 3-D extension checker (f64) ---------------------------------------------------------
    def accel3_f64(self, state7: np.ndarray, law: int, eps2: float, rows: np.ndarray | None = None) -> np.ndarray:
        s = np.ascontiguousarray(state7, dtype=np.float64).reshape(-1, 7)
        if rows is None:
            out = np.empty((s.shape[0], 3), dtype=np.float64)
            self.L.ora3_accel_f64(s.ctypes.data, s.shape[0], law, eps2, None, s.shape[0], out.ctypes.data)
        else:
            r = np.ascontiguousarray(rows, dtype=np.int32)
            out = np.empty((r.shape[0], 3), dtype=np.float64)
            self.L.ora3_accel_f64(s.ctypes.data, s.shape[0], law, eps2, r.ctypes.data, r.shape[0], out.ctypes.data)
        return out

    def step3_f64(self, state7: np.ndarray, law: int, eps2: float, dt: float) -> np.ndarray:
        s = np.array(state7, dtype=np.float64, copy=True).reshape(-1, 7)
        self.L.ora3_step_f64(s.ctypes.data, s.shape[0], law, eps2, dt)
        return s

    # -- draw -------------------------------------------------------------------------------
    def draw(self, w: int, h: int) -> np.ndarray:
        fb = np.empty((h, w), dtype=np.uint32)
        self.L.ora_draw(w, h, fb.ctypes.data)
        return fb


_singleton: Oracle | None = None


def get() -> Oracle:
    global _singleton
    if _singleton is None:
        _singleton = Oracle()
    return _singleton
