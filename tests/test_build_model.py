"""CPU: the single-pass tree build of nb_bh.cu as a numpy model, against the oracle's tree.

nb_bh.cu never inserts bodies one by one: it sorts them by the quadrant path of the reference's own f32 midpoint
recursion (bh_keys_kernel), takes delta(i) = number of levels sorted bodies i and i+1 share (bh_delta_kernel), caps it
for chains of bodies closer than EPS (bh_cap_kernel: the reference's merge rule, rs-src/nbody.rs:249-260) and declares
that an interior node at level l whose first body is i exists exactly for l in (dcap(i-1), dcap(i)].  This model
restates those three kernels and checks the claim against the tree the oracle builds by insertion
(rs-src/nbody.rs:226-301): the same set of interior cells, level by level, box by box (bitwise), with the same body
counts.  The GPU twin is tests/test_gpu_bh.py::test_fast_tree_equals_reference_tree_node_by_node."""
import numpy as np
import pytest

from rust_exp_b200 import ic

f32 = np.float32
LEVELS = 24
EPS = f32(1e-4)


def keys_of(x, y, box):
    """bh_keys_kernel: 2 bits per level, UL=0 UR=1 LL=2 LR=3, cell bounds by cx = (x1+x2)*0.5 in f32."""
    n = len(x)
    x1 = np.full(n, box[0], f32); y1 = np.full(n, box[1], f32); x2 = np.full(n, box[2], f32); y2 = np.full(n, box[3], f32)
    key = np.zeros(n, np.uint64)
    for _ in range(LEVELS):
        cx = (x1 + x2) * f32(0.5)
        cy = (y1 + y2) * f32(0.5)
        lower, left = y < cy, x < cx
        q = np.where(lower, 2, 0) | np.where(left, 0, 1)
        key = (key << np.uint64(2)) | q.astype(np.uint64)
        x2 = np.where(left, cx, x2); x1 = np.where(left, x1, cx)
        y2 = np.where(lower, cy, y2); y1 = np.where(lower, y1, cy)
    return key


def shared_levels(k):
    """bh_delta_kernel: levels that sorted neighbours share (LEVELS for identical keys)."""
    xor = k[:-1] ^ k[1:]
    bl = np.array([int(v).bit_length() for v in xor], dtype=np.int64)
    return np.where(xor == 0, LEVELS, (2 * LEVELS - bl) >> 1)


def capped(delta, close, n):
    """bh_cap_kernel: dcap(j) for j in [0, n-1); pairs inside a chain of close bodies are capped below the first level at
    which the chain is alone in its cell."""
    dcap = np.minimum(delta, LEVELS - 1).astype(np.int64)
    for j in np.nonzero(close)[0]:
        p, q = j, j + 1
        while p > 0 and close[p - 1]:
            p -= 1
        while q < n - 1 and close[q]:
            q += 1
        dl = delta[p - 1] if p > 0 else -1
        dr = delta[q] if q < n - 1 else -1
        dcap[j] = min(dcap[j], max(dl, dr))
    return dcap


def box_of(key, level, box):
    x1, y1, x2, y2 = (f32(v) for v in box)
    for t in range(level):
        q = (int(key) >> (2 * (LEVELS - 1 - t))) & 3
        cx, cy = (x1 + x2) * f32(0.5), (y1 + y2) * f32(0.5)
        if q & 1: x1 = cx
        else: x2 = cx
        if q & 2: y2 = cy
        else: y1 = cy
    return (x1, y1, x2, y2)


@pytest.mark.parametrize("case", ["disk", "plummer", "merges"])
def test_delta_rule_gives_exactly_the_reference_tree(oracle, case):
    n = 3000
    s = ic.random_disk(n, seed=17) if case != "plummer" else ic.plummer_2d(n, seed=17)
    if case == "merges":   # too-close pairs and a chain of three (each closer than EPS to the next)
        s[10, :2] = s[11, :2] + f32(3e-5)
        s[500, :2] = s[501, :2] + f32(4e-5)
        s[502, :2] = s[500, :2] + f32(4e-5)
        s[2000, :2] = s[2001, :2]           # coincident
    oracle.set_particles(s)
    oracle.bh_build()
    flat = oracle.bh_flatten()
    ref_cells = {(int(r[8]), f32(r[0]).tobytes(), f32(r[1]).tobytes(), f32(r[2]).tobytes(), f32(r[3]).tobytes()): float(r[6])
                 for r in flat if r[7] != 0}
    x, y, m = s[:, 0], s[:, 1], s[:, 4]
    box = (x.min(), y.min(), x.max(), y.max())                       # tight, non-square (rs-src/nbody.rs:388-398)
    key = keys_of(x, y, box)
    order = np.argsort(key, kind="stable")
    k, sx, sy, sm = key[order], x[order], y[order], m[order]
    delta = shared_levels(k)
    close = (np.abs(sx[:-1] - sx[1:]) < EPS) & (np.abs(sy[:-1] - sy[1:]) < EPS)
    dcap = np.concatenate([capped(delta, close, n), [-1]])           # dcap(n-1) = -1
    mine = {}
    pm = np.concatenate([[0.0], np.cumsum(sm.astype(np.float64))])
    for i in range(n):
        dlo = dcap[i - 1] if i > 0 else -1
        for l in range(dlo + 1, dcap[i] + 1):
            end = n
            if l > 0:
                j = i + 1
                while j < n and delta[j - 1] >= l:
                    j += 1
                end = j
            b = box_of(k[i], l, box)
            mine[(l, b[0].tobytes(), b[1].tobytes(), b[2].tobytes(), b[3].tobytes())] = pm[end] - pm[i]
    assert len(mine) == len(ref_cells) and set(mine) == set(ref_cells)             # same interior cells, bitwise boxes
    worst = max(abs(mine[c] - ref_cells[c]) / ref_cells[c] for c in mine)
    assert worst < 2e-5                                                            # same bodies inside (mass: f64 sum vs f32 running sum)
