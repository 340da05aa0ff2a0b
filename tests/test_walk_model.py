"""CPU: the warp walk of nb_bh.cu (a shared stack of (node block, lane mask) entries, four children per pop, only
children that some lane must open are pushed) gives every body exactly the interaction list of the reference's per-body
recursion (rs-src/nbody.rs:333-377) -- checked on tools/walk_model.py, the numpy model of the kernel's control flow that
also produced profiles/r02_walk_model_groupings.json."""
import importlib.util
import os

import numpy as np

from rust_exp_b200 import ic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


def test_warp_walk_visits_exactly_the_reference_interaction_lists(oracle):
    spec = importlib.util.spec_from_file_location("walk_model", os.path.join(ROOT, "tools", "walk_model.py"))
    wm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(wm)
    n, theta = 3000, f32(0.6)
    s = ic.random_disk(n, seed=23)
    oracle.set_particles(s)
    oracle.bh_build()
    flat = oracle.bh_flatten()
    kids = wm.children_of(flat)
    K = np.full((len(flat), 4), -1, np.int64)
    for i, k in enumerate(kids):
        if k:
            K[i, :len(k)] = k
    X, Y = flat[:, 4].astype(f32), flat[:, 5].astype(f32)
    S = (flat[:, 2] - flat[:, 0]).astype(f32)
    interior = flat[:, 7] != 0

    def reference_list(px, py):   # the recursion of compute_force: accepted interior nodes and every leaf reached
        out, stack = [], [0]
        while stack:
            i = stack.pop()
            if not interior[i]:
                out.append(i)
                continue
            dx, dy = X[i] - px, Y[i] - py
            d = np.sqrt(f32(f32(dx * dx) + f32(dy * dy)))
            with np.errstate(divide="ignore"):
                if f32(S[i] / d) < theta:
                    out.append(i)
                else:
                    stack.extend(reversed(kids[i]))
        return out

    leaves = np.nonzero((~interior) & (flat[:, 6] != 0))[0]
    total_pops = 0
    for g in (0, 7, 40, len(leaves) // 32 - 1):
        sel = leaves[32 * g:32 * g + 32]
        rec = [[] for _ in sel]
        pops, lane_pops = wm.walk_group(X[sel], Y[sel], theta, K, X, Y, S, interior, record=rec)
        total_pops += pops
        assert 32 <= lane_pops <= 32 * pops
        for lane, body in enumerate(sel):
            ref = reference_list(X[body], Y[body])
            assert sorted(rec[lane]) == sorted(ref)            # the same nodes, each exactly once
            assert len(set(rec[lane])) == len(rec[lane])
    assert total_pops > 100
