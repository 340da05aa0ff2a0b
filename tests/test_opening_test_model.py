"""CPU: the FAST walk's folded opening test against the reference's, decision by decision.

The reference opens an interior node unless `s / d < theta` with s = x2 - x1 and d = sqrt(dx^2 + dy^2) in f32
(rs-src/nbody.rs:341-345).  The walk kernel compares `q < dy*dy + (dx*dx + EPS)` where the build stored
q = (s*s) / theta^2 + EPS (nb_bh.cu qrec(), bh_traverse_fast_kernel) -- algebraically the same test, rounded
differently.  This model replays the reference's traversal on the oracle's own tree for a sample of bodies and counts
the (body, node) decisions on which the two f32 evaluations disagree: they must be a vanishing fraction (the GPU
parity tests allow interaction-count differences of 2e-4)."""
import numpy as np
import pytest

from rust_exp_b200 import ic

f32 = np.float32
EPS = f32(1e-4)


def fma32(a, b, c):
    """round_f32(a*b + c) for f32 inputs (a*b is exact in f64; the f64 sum's own rounding is far below an f32 ulp)."""
    return f32(np.float64(a) * np.float64(b) + np.float64(c))


def children_of(flat):
    """flat = oracle DFS pre-order rows (x1,y1,x2,y2,px,py,m,has_children,depth) -> per node the 4 child row indices."""
    kids = [None] * len(flat)
    stack = []   # (node, children found so far)
    for i, r in enumerate(flat):
        while stack and len(kids[stack[-1]]) == 4:
            stack.pop()
        if stack:
            kids[stack[-1]].append(i)
        if r[7] != 0:
            kids[i] = []
            stack.append(i)
    return kids


@pytest.mark.parametrize("gen,theta", [("disk", 0.5), ("plummer", 0.75), ("disk", 1e-3)])
def test_folded_opening_test_agrees_with_the_reference_decision_by_decision(oracle, gen, theta):
    n = 20000
    s = ic.random_disk(n, seed=11) if gen == "disk" else ic.plummer_2d(n, seed=11)
    oracle.set_particles(s)
    oracle.bh_build()
    flat = oracle.bh_flatten()
    kids = children_of(flat)
    th = f32(theta)
    inv_t2 = f32(1.0) / f32(th * th)          # host: inv_theta2()
    rng = np.random.default_rng(3)
    decisions = differ = visited_ref = visited_fold = 0
    for b in rng.choice(n, 60, replace=False):
        px, py = s[b, 0], s[b, 1]
        for use_folded in (False, True):
            stack, count = [0], 0
            while stack:
                i = stack.pop()
                r = flat[i]
                count += 1
                if r[7] == 0:
                    continue                                                           # leaf: always an interaction
                sw = f32(f32(r[2]) - f32(r[0]))
                dx, dy = f32(f32(r[4]) - px), f32(f32(r[5]) - py)
                d = np.sqrt(f32(f32(dx * dx) + f32(dy * dy)))
                with np.errstate(divide="ignore"):
                    ref_accept = bool(f32(sw / d) < th)                                  # rs-src/nbody.rs:345
                q = fma32(f32(sw * sw), inv_t2, EPS)                                   # qrec()
                e = fma32(dy, dy, fma32(dx, dx, EPS))                                  # the walk's d^2 + EPS
                fold_accept = bool(q < e)
                if not use_folded:
                    decisions += 1
                    differ += ref_accept != fold_accept
                if not (fold_accept if use_folded else ref_accept):
                    stack.extend(reversed(kids[i]))
            if use_folded:
                visited_fold += count
            else:
                visited_ref += count
    assert decisions > 5000
    assert differ <= max(1, decisions * 2e-5), (differ, decisions)
    assert abs(visited_fold - visited_ref) <= max(8, visited_ref * 5e-5), (visited_fold, visited_ref)
