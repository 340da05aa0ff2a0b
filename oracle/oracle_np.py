"""Independent numpy-float32 restatement of rs-src/nbody.rs, used ONLY to cross-check the C oracle.

TEST INFRASTRUCTURE.  Written separately from nbody_oracle.c (vectorised over i for the all-pairs
step, plain Python recursion with np.float32 scalars for the tree) so that a slip in one restatement
shows up as a bit difference against the other.  Every float operation is a numpy float32 operation,
i.e. one IEEE binary32 rounding, no contraction.

Parity status: unpinned by reference vectors (see nbody_oracle.c).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
EPS = f32(0.0001)  # rs-src/nbody.rs:17
VP_WDH = f32(100.0)  # rs-src/nbody.rs:13


def step_brute_force(aos: np.ndarray, dt: float) -> np.ndarray:
    """rs-src/nbody.rs:106-162.  aos is (N,5) float32 {px,py,vx,vy,m}; returns the new state."""
    p = np.array(aos, dtype=f32, copy=True)
    n = p.shape[0]
    px, py, m = p[:, 0].copy(), p[:, 1].copy(), p[:, 4].copy()
    fx = np.zeros(n, dtype=f32)
    fy = np.zeros(n, dtype=f32)
    idx = np.arange(n)
    for j in range(n):  # ascending j: the accumulation order of rs-src/nbody.rs:135-143
        dx = px[j] - px  # rs-src/nbody.rs:174  (p2 - p1, p1 = body i)
        dy = py[j] - py
        d2 = dx * dx + dy * dy
        f = (m * m[j]) / (d2 + EPS)  # rs-src/nbody.rs:180  m1*m2 with m1 = body i
        ax, ay = f * dx, f * dy
        keep = idx != j  # rs-src/nbody.rs:136
        fx = np.where(keep, fx + ax, fx)
        fy = np.where(keep, fy + ay, fy)
    dt = f32(dt)
    p[:, 2] = p[:, 2] + (dt * fx) / m  # rs-src/nbody.rs:155
    p[:, 3] = p[:, 3] + (dt * fy) / m
    p[:, 0] = p[:, 0] + dt * p[:, 2]  # rs-src/nbody.rs:158
    p[:, 1] = p[:, 1] + dt * p[:, 3]
    return p


def _force(px1, py1, m1, px2, py2, m2):
    """rs-src/nbody.rs:164-184"""
    dx = px2 - px1
    dy = py2 - py1
    d2 = dx * dx + dy * dy
    f = m1 * m2 / (d2 + EPS)
    return f * dx, f * dy


class _Node:
    """rs-src/nbody.rs:206-214"""

    __slots__ = ("x1", "y1", "x2", "y2", "px", "py", "m", "ch")

    def __init__(self, x1, y1, x2, y2):
        self.x1, self.y1, self.x2, self.y2 = x1, y1, x2, y2
        self.px = self.py = self.m = f32(0)
        self.ch = None

    def add_mass(self, px, py, m):  # rs-src/nbody.rs:303-320
        assert m > 0
        if self.m == 0:
            self.px, self.py, self.m = px, py, m
        else:
            inv = f32(1.0) / (self.m + m)
            self.px = (self.px * self.m + px * m) * inv
            self.py = (self.py * self.m + py * m) * inv
            self.m = self.m + m

    def quadrant(self, x, y):  # rs-src/nbody.rs:322-331 ; UL=0 UR=1 LL=2 LR=3
        cx = (self.x1 + self.x2) * f32(0.5)
        cy = (self.y1 + self.y2) * f32(0.5)
        if y < cy:
            return 2 if x < cx else 3
        return 0 if x < cx else 1

    def create_children(self):  # rs-src/nbody.rs:286-301
        cx = (self.x1 + self.x2) * f32(0.5)
        cy = (self.y1 + self.y2) * f32(0.5)
        self.ch = [
            _Node(self.x1, cy, cx, self.y2),
            _Node(cx, cy, self.x2, self.y2),
            _Node(self.x1, self.y1, cx, cy),
            _Node(cx, self.y1, self.x2, cy),
        ]

    def insert(self, px, py, m, depth):  # rs-src/nbody.rs:226-284
        if depth > 50:
            raise RecursionError("Node::insert depth > 50")
        if self.ch is not None:
            self.add_mass(px, py, m)
            self.ch[self.quadrant(px, py)].insert(px, py, m, depth + 1)
        else:
            too_close = abs(self.px - px) < EPS and abs(self.py - py) < EPS
            if self.m == 0 or too_close:
                self.add_mass(px, py, m)
            else:
                po, qo, mo = self.px, self.py, self.m
                self.px = self.py = self.m = f32(0)
                self.create_children()
                self.insert(po, qo, mo, depth + 1)
                self.insert(px, py, m, depth + 1)

    def compute_force(self, px, py, m, theta):  # rs-src/nbody.rs:333-377
        fx = fy = f32(0)
        if self.ch is not None:
            s = self.x2 - self.x1
            dx = self.px - px
            dy = self.py - py
            d = np.sqrt(dx * dx + dy * dy)
            with np.errstate(divide="ignore", invalid="ignore"):
                accept = s / d < theta
            if accept:
                fx, fy = _force(px, py, m, self.px, self.py, self.m)
            else:
                for c in self.ch:
                    ax, ay = c.compute_force(px, py, m, theta)
                    fx = fx + ax
                    fy = fy + ay
        else:
            if self.px == px and self.py == py:
                return f32(0), f32(0)
            if self.m == 0:
                return f32(0), f32(0)
            fx, fy = _force(px, py, m, self.px, self.py, self.m)
        return fx, fy


def build_tree(aos: np.ndarray) -> _Node:
    """rs-src/nbody.rs:388-417"""
    p = np.asarray(aos, dtype=f32)
    x1 = y1 = f32(np.finfo(f32).max)
    x2 = y2 = f32(np.finfo(f32).min)
    for q in p:
        x1 = q[0] if q[0] < x1 else x1
        y1 = q[1] if q[1] < y1 else y1
        x2 = q[0] if q[0] > x2 else x2
        y2 = q[1] if q[1] > y2 else y2
    root = _Node(x1, y1, x2, y2)
    for q in p:
        root.insert(q[0], q[1], q[4], 0)
    return root


def flatten(root: _Node) -> np.ndarray:
    """DFS pre-order records x1,y1,x2,y2,px,py,m,has_children,depth (same as ora_bh_flatten)."""
    out = []

    def rec(n, d):
        out.append([n.x1, n.y1, n.x2, n.y2, n.px, n.py, n.m, 1.0 if n.ch else 0.0, float(d)])
        if n.ch:
            for c in n.ch:
                rec(c, d + 1)

    rec(root, 0)
    return np.array(out, dtype=f32)


def step_barnes_hut(aos: np.ndarray, theta: float, dt: float) -> np.ndarray:
    """rs-src/nbody.rs:186-480 (thread partition omitted: it is result-neutral)."""
    if f32(theta) == 0:
        return step_brute_force(aos, dt)  # rs-src/nbody.rs:197-200
    p = np.array(aos, dtype=f32, copy=True)
    root = build_tree(p)
    theta, dt = f32(theta), f32(dt)
    lim = VP_WDH * f32(0.55)
    for q in p:
        fx, fy = root.compute_force(q[0], q[1], q[4], theta)
        q[2] = q[2] + dt * fx / q[4]  # rs-src/nbody.rs:453-454
        q[3] = q[3] + dt * fy / q[4]
        q[0] = q[0] + dt * q[2]  # rs-src/nbody.rs:457-458
        q[1] = q[1] + dt * q[3]
        if abs(f32(0) - q[0]) > lim or abs(f32(0) - q[1]) > lim:  # rs-src/nbody.rs:466-471
            q[2] = 0
            q[3] = 0
    return p
