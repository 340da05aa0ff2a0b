"""CPU model of the FAST Barnes-Hut walk (nb_bh.cu::bh_traverse_fast_kernel) on the oracle's tree: counts the stack pops a
warp makes for a group of 32 bodies -- the quantity the kernel's time is proportional to (72.5 instructions per pop,
issue-bound) -- for alternative ways of forming the groups.  No GPU needed; used to rank the "next" ideas of DESIGN.md
section 8 before spending GPU time on them.

    python tools/walk_model.py [n] [theta] > profiles/r02_walk_model_groupings.json

Schemes:  morton32   32 consecutive bodies in Morton (key) order per warp            -- what the kernel does
          hilbert32  32 consecutive bodies in Hilbert order per warp
          half16 / quarter8 / eighth4   the same groups, but 2 / 4 / 8 independent sub-stacks of 16 / 8 / 4 lanes per
                     warp: iterations = max over the sub-stacks (each load then touches 2 / 4 / 8 block addresses)
A pop evaluates the four children of one opened node for the lanes of its mask; a lane that fails the reference's test
s/d < theta on an interior child asks for it to be opened (rs-src/nbody.rs:333-377)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402  (a test tool)
from rust_exp_b200 import ic  # noqa: E402

f32 = np.float32


def children_of(flat):
    kids = [None] * len(flat)
    stack = []
    for i, r in enumerate(flat):
        while stack and len(kids[stack[-1]]) == 4:
            stack.pop()
        if stack:
            kids[stack[-1]].append(i)
        if r[7] != 0:
            kids[i] = []
            stack.append(i)
    return kids


def hilbert_index(ix, iy, bits):
    """Hilbert curve index of integer grid points (vectorised classic xy2d)."""
    d = np.zeros(len(ix), np.int64)
    x, y = ix.astype(np.int64).copy(), iy.astype(np.int64).copy()
    s = 1 << (bits - 1)
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        flip = (ry == 0) & (rx == 1)
        x = np.where(flip, s - 1 - x, x)
        y = np.where(flip, s - 1 - y, y)
        swap = ry == 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        s >>= 1
    return d


def walk_group(px, py, theta, K, X, Y, S, INTERIOR, record=None, on_pop=None):
    """pops and lane-pops of one warp walking for the bodies (px, py); K[node] = its 4 children (array, -1 = none).
    record: optional list of per-lane lists that receive the nodes each lane interacts with, in evaluation order."""
    lanes = len(px)
    stack = [(0, np.ones(lanes, bool))]          # (node whose children are evaluated, lane mask); root block = {root}
    pops = lane_pops = 0
    first = True                                 # the root itself is "block 0": one pop that tests the root for every lane
    while stack:
        node, mask = stack.pop()
        pops += 1
        lane_pops += int(mask.sum())
        if on_pop is not None and not first:
            on_pop(node)                                      # the block popped holds the children of `node`
        ch = np.array([0, -1, -1, -1]) if first else K[node]
        first = False
        opens = []
        for c in ch:
            if c < 0:
                continue
            if not INTERIOR[c]:                               # leaf: always accepted (an empty one contributes nothing)
                if record is not None:
                    for lane in np.nonzero(mask)[0]:
                        record[lane].append(int(c))
                continue
            dx, dy = X[c] - px, Y[c] - py
            d = np.sqrt((dx * dx + dy * dy).astype(f32))
            with np.errstate(divide="ignore"):
                accept = (S[c] / d).astype(f32) < theta
            if record is not None:
                for lane in np.nonzero(mask & accept)[0]:
                    record[lane].append(int(c))
            need = mask & ~accept
            if need.any():
                opens.append((int(c), need))
        for c, need in reversed(opens):                       # child 0 ends up on top
            stack.append((c, need))
    return pops, lane_pops


def remote_pops_study(n, theta, nparts, cut_level=5):
    """Domain-partitioned walk: which share of a rank's pops reads node blocks that live in ANOTHER
    rank's memory, by depth below the cut -- i.e. what replicating the first k levels below the cut locally would save.
    Parts = contiguous runs of cut-level cells in Morton order with equal body counts."""
    s = ic.random_disk(n, seed=5)
    o = oracle.get()
    o.set_particles(s)
    o.bh_build()
    flat = o.bh_flatten()
    kids = children_of(flat)
    N = len(flat)
    K = np.full((N, 4), -1, np.int64)
    for i, k in enumerate(kids):
        if k:
            K[i, :len(k)] = k
    X, Y = flat[:, 4].astype(f32), flat[:, 5].astype(f32)
    S = (flat[:, 2] - flat[:, 0]).astype(f32)
    INTERIOR = flat[:, 7] != 0
    depth = flat[:, 8].astype(np.int64)
    cell = np.full(N, -1, np.int64)          # DFS index of the node's ancestor at the cut level (-1: above the cut)
    cur = -1
    for i in range(N):                       # DFS pre-order: a cut-level node starts a new cell, deeper nodes inherit it
        if depth[i] == cut_level:
            cur = i
        if depth[i] >= cut_level:
            cell[i] = cur
    body_leaf = (~INTERIOR) & (flat[:, 6] != 0)
    leaves = np.nonzero(body_leaf)[0]
    cells = np.unique(cell[cell >= 0])
    count = {c: 0 for c in cells}
    for l in leaves:
        if cell[l] >= 0:
            count[cell[l]] += 1
    total = sum(count.values())
    part_of_cell, acc = {}, 0
    for c in cells:                          # Morton order = DFS order
        part_of_cell[c] = min(nparts - 1, acc * nparts // max(total, 1))
        acc += count[c]
    part = np.array([part_of_cell[c] if c >= 0 else -1 for c in cell])
    hist = {}                                # depth below the cut -> [local pops, remote pops]
    top = [0]
    ngroups = len(leaves) // 32
    for g in range(0, ngroups, 4):           # every 4th group is plenty
        sel = leaves[32 * g:32 * g + 32]
        mine = part[sel[0]]

        def on_pop(node, mine=mine):
            if cell[node] < 0:
                top[0] += 1
                return
            d = int(depth[node] - cut_level)
            h = hist.setdefault(d, [0, 0])
            h[1 if part[node] != mine else 0] += 1

        walk_group(X[sel], Y[sel], theta, K, X, Y, S, INTERIOR, on_pop=on_pop)
    allp = top[0] + sum(a + b for a, b in hist.values())
    remote = sum(b for _, b in hist.values())
    out = {"n_bodies": int(len(leaves)), "theta": float(theta), "parts": nparts, "cut_level": cut_level,
           "pops_in_the_shared_top_tree": top[0] / allp, "remote_pops": remote / allp, "by_depth_below_cut": {}}
    cum = 0
    for d in sorted(hist):
        cum += hist[d][1]
        out["by_depth_below_cut"][d] = {"local": hist[d][0] / allp, "remote": hist[d][1] / allp,
                                        "remote_pops_removed_by_replicating_down_to_here": cum / max(remote, 1)}
    print(json.dumps(out, indent=1))


def main():
    if len(sys.argv) > 3 and sys.argv[3].startswith("parts="):
        return remote_pops_study(int(sys.argv[1]), f32(float(sys.argv[2])), int(sys.argv[3][6:]))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    theta = f32(float(sys.argv[2]) if len(sys.argv) > 2 else 0.5)
    s = ic.random_disk(n, seed=4)
    o = oracle.get()
    o.set_particles(s)
    o.bh_build()
    flat = o.bh_flatten()
    kids = children_of(flat)
    N = len(flat)
    K = np.full((N, 4), -1, np.int64)
    for i, k in enumerate(kids):
        if k:
            K[i, :len(k)] = k
    X, Y = flat[:, 4].astype(f32), flat[:, 5].astype(f32)
    S = (flat[:, 2] - flat[:, 0]).astype(f32)
    INTERIOR = flat[:, 7] != 0
    leaves = np.nonzero((~INTERIOR) & (flat[:, 6] != 0))[0]  # non-empty leaves in DFS order = bodies in Morton order
    bx, by = X[leaves], Y[leaves]
    x1, y1, x2, y2 = flat[0, 0], flat[0, 1], flat[0, 2], flat[0, 3]
    bits = 16
    ix = np.clip(((bx - x1) / (x2 - x1) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    iy = np.clip(((by - y1) / (y2 - y1) * (1 << bits)).astype(np.int64), 0, (1 << bits) - 1)
    orders = {"morton": np.arange(len(leaves)), "hilbert": np.argsort(hilbert_index(ix, iy, bits), kind="stable")}
    out = {"n_bodies": int(len(leaves)), "theta": float(theta), "ic": "uniform disk (seed 4)", "tree_nodes": int(N), "schemes": {}}
    for oname, order in orders.items():
        gx, gy = bx[order], by[order]
        pops32 = lanes32 = 0
        sub = {16: [0, 0], 8: [0, 0], 4: [0, 0]}               # lanes per sub-stack -> [iterations, sub-pops]
        ngroups = (len(gx) + 31) // 32
        for g in range(ngroups):
            qx, qy = gx[32 * g:32 * g + 32], gy[32 * g:32 * g + 32]
            p, lp = walk_group(qx, qy, theta, K, X, Y, S, INTERIOR)
            pops32 += p
            lanes32 += lp
            for w in sub:
                qp = [walk_group(qx[w * k:w * k + w], qy[w * k:w * k + w], theta, K, X, Y, S, INTERIOR)[0]
                      for k in range(32 // w) if len(qx) > w * k]
                sub[w][0] += max(qp)
                sub[w][1] += sum(qp)
        out["schemes"][f"{oname}32"] = {"pops_per_group": pops32 / ngroups, "lane_efficiency": lanes32 / (32.0 * pops32)}
        for w, (it, sp) in sub.items():
            name = {16: "half16", 8: "quarter8", 4: "eighth4"}[w]
            out["schemes"][f"{oname}_{name}"] = {"iterations_per_group": it / ngroups, "sub_pops_per_group": sp / ngroups,
                                                 "balance_max_over_mean": it / (sp / (32.0 / w)),
                                                 "distinct_block_addresses_per_load": 32 // w}
    base = out["schemes"]["morton32"]["pops_per_group"]
    for k, v in out["schemes"].items():
        v["iterations_relative_to_morton32"] = (v.get("pops_per_group") or v.get("iterations_per_group")) / base
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
