"""rust_exp_b200 -- B200-native drop-in for the N-body hot path of blitzcode/rust-exp.

The product is the C-ABI shared library ``libnbody_b200.so`` (include/nbody_b200.h) built from
``csrc/``; this package is the thin host side above it:

* ``binding``     ctypes binding of every symbol in include/nbody_b200.h (no torch types).
* ``experiment``  host-side mirror of hs-src/RustNBodyExperiment.hs (the reference's plug-in for this path).
* ``dist``        torch.distributed plumbing that wires the per-GPU processes together (handle exchange).
* ``ic``          seeded initial-condition generators for the benchmark configurations.

(The directory is ``rust_exp_b200`` rather than ``rust-exp_b200`` because a Python package name cannot
contain a hyphen.)

There is no CPU fallback anywhere in this package: without a CUDA device every compute entry point
fails loudly.
"""
from .binding import NBodyLib, load, lib_path, SYMBOLS  # noqa: F401
from .experiment import RustNBodyExperiment  # noqa: F401

__all__ = ["NBodyLib", "load", "lib_path", "SYMBOLS", "RustNBodyExperiment"]
