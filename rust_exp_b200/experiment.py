"""Host-side mirror of the reference's plug-in for this path, hs-src/RustNBodyExperiment.hs.

Same fields, defaults, key bindings, clamps and status line as the Haskell `Experiment` instance, driving
the same six C symbols in the same order (step, then draw) -- so a test written against the reference
plug-in's behaviour reads the same here.  It exists to exercise the drop-in boundary exactly the way
the reference host does; it contains no numerics of its own.
"""
from __future__ import annotations

import statistics
import time
from collections import deque

import numpy as np

from .binding import NBodyLib, load


class RustNBodyExperiment:
    name = "RustNBody"  # hs-src/RustNBodyExperiment.hs:49

    def __init__(self, lib: NBodyLib | None = None) -> None:
        self.lib = lib or load()
        # withExperiment, hs-src/RustNBodyExperiment.hs:42-48
        self.lib.stable_orbits(10000, 0.5, 30.0)
        self.num_steps = 0
        self.times: deque[float] = deque(maxlen=30)
        self.time_step = 0.01
        self.theta = 0.85
        self.num_threads = 1

    # experimentDraw, hs-src/RustNBodyExperiment.hs:50-62: simulate first, then draw
    def draw(self, fb: np.ndarray) -> None:
        h, w = fb.shape
        t0 = time.perf_counter()
        self.lib.step_barnes_hut(self.theta, self.time_step, self.num_threads)
        self.lib.synchronize()  # the reference call is synchronous; keep `timeIt` meaningful
        self.times.append(time.perf_counter() - t0)
        self.lib.draw(w, h, fb)
        self.num_steps += 1

    # experimentStatusString, hs-src/RustNBodyExperiment.hs:63-80
    def status_string(self) -> str:
        avg = statistics.median(self.times) if self.times else 1.0
        n = self.lib.num_particles()
        bodies = f"{n // 1000}K" if n > 999 else str(n)
        return (
            f"{self.num_steps} Steps, {1 / avg:.1f}SPS/{avg * 1000:.2f}ms | {bodies} Bodies\n"
            f"[QWE] Scene | Time Step [X][x]: {self.time_step:.4f} | Theta [A][a]: {self.theta:.2f} | "
            f"Threads [P][p]: {self.num_threads}"
        )

    # experimentGLFWEvent, hs-src/RustNBodyExperiment.hs:81-99
    def key(self, k: str, shift: bool = False) -> None:
        k = k.upper()
        if k == "Q":
            self.lib.stable_orbits(10000, 0.5, 30.0)
        elif k == "W":
            self.lib.random_disk(10000)
        elif k == "E":
            self.lib.stable_orbits(5, 5.0, 40.0)
        elif k == "X":
            self.time_step = self.time_step / 2 if shift else self.time_step * 2
        elif k == "A":
            self.theta = max(0.0, min(0.95, self.theta - 0.05 if shift else self.theta + 0.05))
        elif k == "P":
            self.num_threads = max(1, min(16, self.num_threads - 1 if shift else self.num_threads + 1))
