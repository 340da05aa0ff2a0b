"""ctypes binding of libnbody_b200.so -- one Python method per symbol of include/nbody_b200.h.

Signatures are the reference's (rs-src/nbody.rs:34-35,39-40,73-74,106-107,186-187,482-483 as imported
by hs-src/RustNBodyExperiment.hs:101-106) plus the documented additions.  Plain pointers and sizes only.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

f32, i32, u64, vp = C.c_float, C.c_int32, C.c_uint64, C.c_void_p

# name -> (restype, argtypes).  tests/test_abi_symbols.py checks this table against the header.
SYMBOLS = {
    # reference surface
    "nb_num_particles": (i32, []),
    "nb_random_disk": (None, [i32]),
    "nb_stable_orbits": (None, [i32, f32, f32]),
    "nb_step_brute_force": (None, [f32]),
    "nb_step_barnes_hut": (None, [f32, f32, i32]),
    "nb_draw": (None, [i32, i32, vp]),
    # required additions
    "nb_set_particles": (None, [vp, i32]),
    "nb_get_particles": (None, [vp, i32]),
    # prefixed aliases (the Rust shim of INTEGRATION.md option A forwards to these)
    "b200_nb_num_particles": (i32, []),
    "b200_nb_random_disk": (None, [i32]),
    "b200_nb_stable_orbits": (None, [i32, f32, f32]),
    "b200_nb_step_brute_force": (None, [f32]),
    "b200_nb_step_barnes_hut": (None, [f32, f32, i32]),
    "b200_nb_draw": (None, [i32, i32, vp]),
    "b200_nb_set_particles": (None, [vp, i32]),
    "b200_nb_get_particles": (None, [vp, i32]),
    # extension surface
    "nbx_init": (i32, [i32]),
    "nbx_shutdown": (None, []),
    "nbx_last_error": (C.c_char_p, []),
    "nbx_version": (C.c_char_p, []),
    "nbx_set_mode": (i32, [i32]),
    "nbx_get_mode": (i32, []),
    "nbx_set_stream": (i32, [vp]),
    "nbx_synchronize": (i32, []),
    "nbx_set_async": (i32, [i32]),
    "nbx_set_integrator": (i32, [i32]),
    "nbx_set_square_aabb": (i32, [i32]),
    "nbx_set_peer_timeout_ms": (i32, [i32]),
    "nbx_seed": (None, [u64]),
    "nbx_plummer": (None, [i32, f32, f32]),
    "nbx_tune": (i32, [i32, i32, i32]),
    "nbx_get_counters": (None, [vp]),
    "nbx_reset_counters": (None, []),
    "nbx_bh_count_interactions": (i32, [i32]),
    "nbx_bh_pop_histogram": (i32, [vp, i32]),
    "nbx_bh_flatten": (i32, [vp, i32]),
    "nbx_bh_partition": (i32, [i32]),
    "nbx_phase_timing": (i32, [i32]),
    "nbx_get_phase_ms": (i32, [vp]),
    "nbx_get_phase_sub_ms": (i32, [i32, vp]),
    "nbx_accelerations": (i32, [vp, i32]),
    "nbx_bh_accelerations": (i32, [f32, vp, i32]),
    "nbx3_num_particles": (i32, []),
    "nbx3_set_particles": (i32, [vp, i32]),
    "nbx3_get_particles": (i32, [vp, i32]),
    "nbx3_configure": (i32, [i32, f32]),
    "nbx3_set_sharded": (i32, [i32]),
    "nbx3_step_all_pairs": (i32, [f32]),
    "nbx3_accelerations": (i32, [vp, i32]),
    "nbx_dist_init": (i32, [i32, i32, i32]),
    "nbx_dist_handle_bytes": (i32, []),
    "nbx_dist_export": (i32, [vp]),
    "nbx_dist_import": (i32, [vp, i32]),
    "nbx_dist_nccl_unique_id": (i32, [vp]),
    "nbx_dist_nccl_init": (i32, [vp]),
    "nbx_dist_set_transport": (i32, [i32]),
    "nbx_dist_local_range": (i32, [vp, vp]),
    "nbx_dist_sync_test": (f32, [i32]),
    "nbx_get_particles_local": (i32, [vp, i32]),
}

MODE_FAST, MODE_EXACT = 0, 1
INTEGRATOR_EULER, INTEGRATOR_LEAPFROG_KDK = 0, 1
TRANSPORT_P2P_DIRECT, TRANSPORT_P2P_GATHER, TRANSPORT_NCCL = 0, 1, 2
LAW3_NEWTON, LAW3_REF = 0, 1
PHASES = ("force", "integrate", "aabb", "keys", "sort", "build", "com", "xrank")


class Counters(C.Structure):
    _fields_ = [
        ("kernel_launches", u64),
        ("allpairs_pairs", u64),
        ("bh_interactions", u64),
        ("bh_nodes_visited", u64),
        ("bh_nodes_built", u64),
        ("steps", u64),
        ("bh_pops", u64),
        ("bh_pop_lanes", u64),
        ("bh_part_bodies", u64),
        ("bh_sort_levels", u64),
    ]


def lib_path() -> str:
    return os.environ.get("NBODY_B200_LIB", os.path.join(_HERE, "libnbody_b200.so"))


class NBodyLib:
    """Typed handle on the shared library.  Loading never touches the GPU; computing requires one."""

    def __init__(self, path: str | None = None) -> None:
        path = path or lib_path()
        if not os.path.exists(path):
            raise FileNotFoundError(
                f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)"
            )
        self.path = path
        self.L = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(self.L, name)  # AttributeError if the symbol is missing: fail loudly
            fn.restype = res
            fn.argtypes = args

    # ---- reference surface ---------------------------------------------------------------------
    def num_particles(self) -> int:
        return int(self.L.nb_num_particles())

    def random_disk(self, n: int) -> None:
        self.L.nb_random_disk(n)

    def stable_orbits(self, n: int, rmin: float, rmax: float) -> None:
        self.L.nb_stable_orbits(n, rmin, rmax)

    def step_brute_force(self, dt: float) -> None:
        self.L.nb_step_brute_force(dt)

    def step_barnes_hut(self, theta: float, dt: float, nthreads: int = 1) -> None:
        self.L.nb_step_barnes_hut(theta, dt, nthreads)

    def draw(self, w: int, h: int, fb: np.ndarray | None = None) -> np.ndarray:
        if fb is None:
            fb = np.empty((h, w), dtype=np.uint32)
        assert fb.dtype == np.uint32 and fb.size == w * h and fb.flags.c_contiguous
        self.L.nb_draw(w, h, fb.ctypes.data)
        return fb

    # ---- additions -------------------------------------------------------------------------------
    def set_particles(self, aos: np.ndarray) -> None:
        a = np.ascontiguousarray(aos, dtype=np.float32).reshape(-1, 5)
        self.L.nb_set_particles(a.ctypes.data, a.shape[0])

    def get_particles(self, out: np.ndarray | None = None) -> np.ndarray:
        n = self.num_particles()
        if out is None:
            out = np.empty((n, 5), dtype=np.float32)
        self.L.nb_get_particles(out.ctypes.data, n)
        return out

    def get_particles_local(self, out: np.ndarray) -> np.ndarray:
        """Owner-only read-back into the caller's full (n,5) array: only this rank's rows are written."""
        n = self.num_particles()
        assert out.dtype == np.float32 and out.size >= 5 * n and out.flags.c_contiguous
        self._chk(self.L.nbx_get_particles_local(out.ctypes.data, n), "nbx_get_particles_local")
        return out

    # ---- extensions ------------------------------------------------------------------------------
    def _chk(self, rc: int, what: str) -> None:
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.last_error()}")

    def init(self, device: int = 0) -> None:
        self._chk(self.L.nbx_init(device), "nbx_init")

    def shutdown(self) -> None:
        self.L.nbx_shutdown()

    def last_error(self) -> str:
        return (self.L.nbx_last_error() or b"").decode()

    def version(self) -> str:
        return self.L.nbx_version().decode()

    def set_mode(self, mode: int) -> None:
        self._chk(self.L.nbx_set_mode(mode), "nbx_set_mode")

    def get_mode(self) -> int:
        return int(self.L.nbx_get_mode())

    def set_stream(self, cuda_stream: int | None) -> None:
        self._chk(self.L.nbx_set_stream(C.c_void_p(cuda_stream or 0)), "nbx_set_stream")

    def synchronize(self) -> None:
        self._chk(self.L.nbx_synchronize(), "nbx_synchronize")

    def set_integrator(self, integrator: int) -> None:
        self._chk(self.L.nbx_set_integrator(integrator), "nbx_set_integrator")

    def set_square_aabb(self, on: bool) -> None:
        self._chk(self.L.nbx_set_square_aabb(1 if on else 0), "nbx_set_square_aabb")

    def set_async(self, on: bool) -> None:
        self._chk(self.L.nbx_set_async(1 if on else 0), "nbx_set_async")

    def set_peer_timeout_ms(self, ms: int) -> None:
        self._chk(self.L.nbx_set_peer_timeout_ms(ms), "nbx_set_peer_timeout_ms")

    def seed(self, s: int) -> None:
        self.L.nbx_seed(s)

    def plummer(self, n: int, a_scale: float = 5.0, mass_per_body: float = 1e-2) -> None:
        """Device-side Plummer set (projected to z = 0), the counterpart of ic.plummer_2d."""
        self.L.nbx_plummer(n, a_scale, mass_per_body)

    def tune(self, bodies_per_thread: int = 0, target_waves: int = 0, ctas_per_sm: int = 0) -> None:
        self._chk(self.L.nbx_tune(bodies_per_thread, target_waves, ctas_per_sm), "nbx_tune")

    def counters(self) -> dict:
        c = Counters()
        self.L.nbx_get_counters(C.byref(c))
        return {k: int(getattr(c, k)) for k, _ in Counters._fields_}

    def reset_counters(self) -> None:
        self.L.nbx_reset_counters()

    def bh_count_interactions(self, on: bool) -> None:
        self.L.nbx_bh_count_interactions(1 if on else 0)

    def bh_pop_histogram(self, reset: bool = True) -> np.ndarray:
        out = np.zeros(33, dtype=np.uint64)
        self._chk(self.L.nbx_bh_pop_histogram(out.ctypes.data, 1 if reset else 0), "nbx_bh_pop_histogram")
        return out

    def bh_flatten(self) -> np.ndarray:
        n = int(self.L.nbx_bh_flatten(None, 0))
        if n < 0:
            raise RuntimeError(f"nbx_bh_flatten failed ({n})")
        out = np.empty((n, 9), dtype=np.float32)
        got = int(self.L.nbx_bh_flatten(out.ctypes.data, n))
        assert got == n, (got, n)
        return out

    def bh_partition(self, parts: int) -> None:
        self._chk(self.L.nbx_bh_partition(parts), "nbx_bh_partition")

    def phase_timing(self, on: bool) -> None:
        self.L.nbx_phase_timing(1 if on else 0)

    def phase_ms(self) -> dict:
        out = (C.c_float * 8)()
        self._chk(self.L.nbx_get_phase_ms(out), "nbx_get_phase_ms")
        return dict(zip(PHASES, [float(v) for v in out]))

    def phase_sub_ms(self, phase: str) -> list:
        out = (C.c_float * 4)()
        self._chk(self.L.nbx_get_phase_sub_ms(PHASES.index(phase), out), "nbx_get_phase_sub_ms")
        return [float(v) for v in out]

    def accelerations(self, n: int | None = None) -> np.ndarray:
        n = self.num_particles() if n is None else n
        out = np.zeros((n, 2), dtype=np.float32)
        self._chk(self.L.nbx_accelerations(out.ctypes.data, n), "nbx_accelerations")
        return out

    def bh_accelerations(self, theta: float, n: int | None = None) -> np.ndarray:
        n = self.num_particles() if n is None else n
        out = np.zeros((n, 2), dtype=np.float32)
        self._chk(self.L.nbx_bh_accelerations(theta, out.ctypes.data, n), "nbx_bh_accelerations")
        return out

    # ---- 3-D extension ----------------------------------------------------------------------------
    def set_particles3(self, aos7: np.ndarray) -> None:
        a = np.ascontiguousarray(aos7, dtype=np.float32).reshape(-1, 7)
        self._chk(self.L.nbx3_set_particles(a.ctypes.data, a.shape[0]), "nbx3_set_particles")

    def get_particles3(self) -> np.ndarray:
        n = int(self.L.nbx3_num_particles())
        out = np.empty((n, 7), dtype=np.float32)
        self._chk(self.L.nbx3_get_particles(out.ctypes.data, n), "nbx3_get_particles")
        return out

    def configure3(self, law: int, eps2: float = 1e-4) -> None:
        self._chk(self.L.nbx3_configure(law, eps2), "nbx3_configure")

    def set_sharded3(self, on: bool) -> None:
        self._chk(self.L.nbx3_set_sharded(1 if on else 0), "nbx3_set_sharded")

    def step3(self, dt: float) -> None:
        self._chk(self.L.nbx3_step_all_pairs(dt), "nbx3_step_all_pairs")

    def accelerations3(self) -> np.ndarray:
        n = int(self.L.nbx3_num_particles())
        out = np.zeros((n, 3), dtype=np.float32)
        self._chk(self.L.nbx3_accelerations(out.ctypes.data, n), "nbx3_accelerations")
        return out

    # ---- multi-GPU -------------------------------------------------------------------------------
    def dist_init(self, rank: int, world: int, max_particles: int) -> None:
        self._chk(self.L.nbx_dist_init(rank, world, max_particles), "nbx_dist_init")

    def dist_export(self) -> bytes:
        buf = C.create_string_buffer(int(self.L.nbx_dist_handle_bytes()))
        self._chk(self.L.nbx_dist_export(buf), "nbx_dist_export")
        return buf.raw

    def dist_import(self, handles: bytes, world: int) -> None:
        buf = C.create_string_buffer(handles, len(handles))
        self._chk(self.L.nbx_dist_import(buf, world), "nbx_dist_import")

    def dist_nccl_unique_id(self) -> bytes:
        buf = C.create_string_buffer(128)
        self._chk(self.L.nbx_dist_nccl_unique_id(buf), "nbx_dist_nccl_unique_id")
        return buf.raw

    def dist_nccl_init(self, uid: bytes) -> None:
        buf = C.create_string_buffer(uid, 128)
        self._chk(self.L.nbx_dist_nccl_init(buf), "nbx_dist_nccl_init")

    def dist_set_transport(self, t: int) -> None:
        self._chk(self.L.nbx_dist_set_transport(t), "nbx_dist_set_transport")

    def dist_sync_test(self, iters: int = 200) -> float:
        return float(self.L.nbx_dist_sync_test(iters))

    def dist_local_range(self) -> tuple[int, int]:
        b, c = i32(0), i32(0)
        self._chk(self.L.nbx_dist_local_range(C.byref(b), C.byref(c)), "nbx_dist_local_range")
        return int(b.value), int(c.value)


_lib: NBodyLib | None = None


def load() -> NBodyLib:
    """Process-wide handle (the library's state is process-global, like the reference's)."""
    global _lib
    if _lib is None:
        _lib = NBodyLib()
    return _lib
