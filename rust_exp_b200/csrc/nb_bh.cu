// nb_bh.cu -- Barnes-Hut step on the device (rs-src/nbody.rs:186-480).
//
// Pipeline of one step (all on the library stream, no host round trip):
//   aabb      tight, non-square bounding box (rs-src/nbody.rs:388-398); exact (min/max do not round)
//   keys      per body, the quadrant path of the reference's own f32 midpoint recursion
//             cx=(x1+x2)*0.5 (rs-src/nbody.rs:289-290,322-331) -> 2 bits per level, child order
//             UL=0 UR=1 LL=2 LR=3 (:295-300).  Cell boundaries therefore match the reference bit for bit.
//   sort      CUB radix sort of (key, index); stable, so equal keys keep particle-index order
//   scan      f64 prefix sums of m, m*x, m*y over the sorted bodies: any node's mass and centre of mass
//             is a difference of two prefix entries, so the build is purely top-down
//   build     level-synchronous: a node with range [first, first+count) in sorted order splits into
//             four children located by binary search on the 2 key bits of its level.  Splitting stops at
//             count<=1, at a too-close pair (|dx|,|dy| < EPS: the reference's merge rule, :249-260), or at
//             the last key level (cells there are smaller than EPS, so their bodies would merge too).
//   traverse  FAST: one warp per 32 Morton-consecutive bodies, shared-memory node stack; the warp walks
//             the union of its lanes' DFS paths, but every lane applies the reference's own per-body
//             opening test and is masked off inside subtrees it has accepted -- so each body evaluates
//             exactly the interaction list the reference would (:333-377), summed in DFS order.
//             EXACT: one thread per body, explicit frame stack reproducing the nested summation order.
//   integrate Euler + velocity kill (:453-471).
//
// EXACT mode builds the tree with a literal single-thread restatement of Node::insert (:226-284) so that
// insertion-order effects (running-mean COM rounding, merges) are reproduced bit for bit; it is the
// semantics pin for the tree, not a throughput path.
#include <cub/device/device_radix_sort.cuh>
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>

#include <algorithm>
#include <cstring>

#include "nb_engine.h"

namespace nb {

constexpr int kLevels = 24;            // key levels (2 bits each); root cell / 2^24 < EPS for extents < 1677
constexpr int kKeyBits = 2 * kLevels;
constexpr int kNone = 0x7fffffff;
constexpr int kStackPerWarp = 4 * kLevels + 8;
constexpr int kTravWarps = 8;

struct BhStatus {
    int node_count;
    int overflow;
    int ticket;
    int depth_error;          // exact build: recursion depth > 50 (the reference panics)
    int lvl_begin[kLevels + 3];
    int aabb_enc[4];          // ordered-int encodings of x1,y1 (min) and x2,y2 (max)
    int n_mine;
    int n_interior;
    unsigned long long interactions;
    unsigned long long visited;
};

struct BhWork {
    int cap_n = 0;
    unsigned long long *keys = nullptr, *keys_sorted = nullptr;
    int *idx = nullptr, *idx_sorted = nullptr, *mine = nullptr;
    float *sx = nullptr, *sy = nullptr, *sm = nullptr;
    double *w3 = nullptr;      // [3][n+1] weights then prefix sums (m, m*x, m*y)
    double *p3 = nullptr;
    double *tile_sums = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    int cap_nodes = 0;
    float4* ndata = nullptr;   // com.x, com.y, mass, width (s = x2-x1; < 0 marks a leaf)
    float4* nbounds = nullptr; // x1,y1,x2,y2
    int* nchild = nullptr;     // first of 4 children, -1 for a leaf   (ndata/nbounds/nchild: EXACT mode, AoS)
    float4* nblk = nullptr;    // FAST mode: block-SoA records
    int4* ncblk = nullptr;
    signed char *delta = nullptr, *dcap = nullptr;
    unsigned char* close = nullptr;
    int *count = nullptr, *base = nullptr, *owner = nullptr;
    BhStatus* status = nullptr;
    BhStatus* status_host = nullptr;
    float2* acc = nullptr;     // per local body: acceleration (FAST) or force (EXACT)
    int cap_acc = 0;
    bool warned = false;
    bool status_pending = false;
    cudaEvent_t status_ev = nullptr;
};

static BhWork& work(Engine& e) {
    if (!e.bh) e.bh = new BhWork();
    return *static_cast<BhWork*>(e.bh);
}

__device__ __forceinline__ int f2ord(float f) {
    const int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

// ---- aabb -----------------------------------------------------------------------------------------
__global__ void bh_reset_kernel(BhStatus* st) {
    if (threadIdx.x == 0) {
        st->node_count = 1;
        st->overflow = 0;
        st->ticket = 0;
        st->depth_error = 0;
        for (int l = 0; l < kLevels + 3; l++) st->lvl_begin[l] = (l == 0) ? 0 : 1;
        st->aabb_enc[0] = st->aabb_enc[1] = f2ord(3.40282347e+38f);
        st->aabb_enc[2] = st->aabb_enc[3] = f2ord(-3.40282347e+38f);
        st->interactions = 0;
        st->visited = 0;
    }
}

__global__ void bh_aabb_kernel(const float* __restrict__ x, const float* __restrict__ y, int n, BhStatus* st) {
    int mnx = f2ord(3.40282347e+38f), mny = mnx, mxx = f2ord(-3.40282347e+38f), mxy = mxx;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int ex = f2ord(x[i]), ey = f2ord(y[i]);
        mnx = min(mnx, ex); mxx = max(mxx, ex);
        mny = min(mny, ey); mxy = max(mxy, ey);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
        mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    __shared__ int red[4][8];
    const int wid = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { red[0][wid] = mnx; red[1][wid] = mny; red[2][wid] = mxx; red[3][wid] = mxy; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int k = 1; k < nw; k++) {
            mnx = min(mnx, red[0][k]); mny = min(mny, red[1][k]);
            mxx = max(mxx, red[2][k]); mxy = max(mxy, red[3][k]);
        }
        atomicMin(&st->aabb_enc[0], mnx);
        atomicMin(&st->aabb_enc[1], mny);
        atomicMax(&st->aabb_enc[2], mxx);
        atomicMax(&st->aabb_enc[3], mxy);
    }
}

// ---- keys: the reference's midpoint recursion, kLevels deep ------------------------------------------
__global__ void bh_keys_kernel(const float* __restrict__ x, const float* __restrict__ y, int n, const BhStatus* st,
                               unsigned long long* __restrict__ keys, int* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x1 = ord2f(st->aabb_enc[0]), y1 = ord2f(st->aabb_enc[1]);
    float x2 = ord2f(st->aabb_enc[2]), y2 = ord2f(st->aabb_enc[3]);
    const float px = x[i], py = y[i];
    unsigned long long key = 0;
#pragma unroll 4
    for (int l = 0; l < kLevels; l++) {
        const float cx = __fmul_rn(__fadd_rn(x1, x2), 0.5f);  // rs-src/nbody.rs:289-290 / :324-325
        const float cy = __fmul_rn(__fadd_rn(y1, y2), 0.5f);
        const bool lower = py < cy, left = px < cx;           // rs-src/nbody.rs:326-330
        const unsigned q = (lower ? 2u : 0u) | (left ? 0u : 1u);
        key = (key << 2) | q;
        // child bounds, rs-src/nbody.rs:295-300
        if (left) x2 = cx; else x1 = cx;
        if (lower) y2 = cy; else y1 = cy;
    }
    keys[i] = key;
    idx[i] = i;
}

__global__ void bh_gather_sorted_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                        const float* __restrict__ m, const int* __restrict__ idx_sorted, int n,
                                        float* __restrict__ sx, float* __restrict__ sy, float* __restrict__ sm,
                                        double* __restrict__ w3) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    double wm = 0.0, wx = 0.0, wy = 0.0;
    if (i < n) {
        const int j = idx_sorted[i];
        const float xx = x[j], yy = y[j], mm = m[j];
        sx[i] = xx; sy[i] = yy; sm[i] = mm;
        wm = mm; wx = static_cast<double>(mm) * xx; wy = static_cast<double>(mm) * yy;
    }
    const size_t stride = static_cast<size_t>(n) + 1;
    w3[i] = wm; w3[stride + i] = wx; w3[2 * stride + i] = wy;
}

// ---- deterministic f64 exclusive scan ---------------------------------------------------------------------
// cub::DeviceScan's decoupled look-back combines tile prefixes in a timing-dependent order, which for
// floating point makes the low bits (and, rarely, a rounded COM and an opening test) vary from run to
// run.  This reduce-then-scan has a fixed association: tile sums in a fixed tree, a single block scanning
// the tile sums sequentially, then a fixed in-tile scan.  blockIdx.y selects one of the 3 arrays.
constexpr int kScanThreads = 256, kScanItems = 8, kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const double* __restrict__ in, int len, size_t stride,
                                                                      double* __restrict__ tile_sums, int ntiles) {
    using BR = cub::BlockReduce<double, kScanThreads>;
    __shared__ typename BR::TempStorage tmp;
    const double* a = in + blockIdx.y * stride;
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) v += (base + k < len) ? a[base + k] : 0.0;
    const double t = BR(tmp).Sum(v);
    if (threadIdx.x == 0) tile_sums[blockIdx.y * ntiles + blockIdx.x] = t;
}
__global__ void __launch_bounds__(kScanThreads) scan_tile_offsets_kernel(double* tile_sums, int ntiles) {
    // one block per array: chunks of kScanThreads tiles, fixed block-scan tree + a sequential carry
    using BS = cub::BlockScan<double, kScanThreads>;
    __shared__ typename BS::TempStorage tmp;
    __shared__ double carry;
    double* t = tile_sums + blockIdx.x * ntiles;
    if (threadIdx.x == 0) carry = 0.0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const double v = i < ntiles ? t[i] : 0.0;
        double ex, total;
        BS(tmp).ExclusiveSum(v, ex, total);
        const double c = carry;
        if (i < ntiles) t[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const double* __restrict__ in, double* __restrict__ out, int len,
                                                                  size_t stride, const double* __restrict__ tile_offs, int ntiles) {
    using BS = cub::BlockScan<double, kScanThreads>;
    __shared__ typename BS::TempStorage tmp;
    const double* a = in + blockIdx.y * stride;
    double* o = out + blockIdx.y * stride;
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    double v[kScanItems];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) v[k] = (base + k < len) ? a[base + k] : 0.0;
    BS(tmp).ExclusiveSum(v, v);
    const double off = tile_offs[blockIdx.y * ntiles + blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) if (base + k < len) o[base + k] = off + v[k];
}

// ---- build: single pass over the sorted keys, no level synchronisation ---------------------------------
//
// delta(i) = number of quadtree levels on which sorted bodies i and i+1 share their cell (common key
// prefix / 2 bits).  A cell with >= 2 bodies is an interior node unless all its bodies merge (reference rule
// rs-src/nbody.rs:249-260: closer than EPS in both axes).  An interior node at level l whose first body is i
// exists exactly for l in (dcap(i-1), dcap(i)], where dcap is delta capped (a) at kLevels-1 and (b), for
// pairs inside a chain of mutually close bodies, below the first level at which the chain is alone in
// its cell -- that cell becomes a merged leaf.  A prefix sum over count(i) = max(0, dcap(i)-dcap(i-1)) numbers
// the interior nodes; node k owns the 4-slot child block [4+4k, 4+4k+4) (64-byte aligned float4 records),
// and every node's record lives in its parent's block (the root in slot 0).  One thread per body then emits
// the nodes that start at it: range end and child boundaries by binary search on the keys, mass/COM from the
// f64 prefix sums, cell bounds by replaying the key bits through the reference's midpoint recursion.
__device__ __forceinline__ int lower_bound_quadrant(const unsigned long long* __restrict__ keys, int lo, int hi,
                                                    int shift, unsigned q) {
    // first position in [lo,hi) whose 2 key bits at `shift` are >= q (keys are sorted)
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (((keys[mid] >> shift) & 3ull) < q) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// rs-src/nbody.rs:303-320 with the reference's rounding (used for merged leaves)
__device__ __forceinline__ void add_mass_ref(float& px, float& py, float& m, float x, float y, float mm) {
    if (m == 0.0f) {
        px = x; py = y; m = mm;
    } else {
        const float inv = __fdiv_rn(1.0f, __fadd_rn(m, mm));
        px = __fmul_rn(__fadd_rn(__fmul_rn(px, m), __fmul_rn(x, mm)), inv);
        py = __fmul_rn(__fadd_rn(__fmul_rn(py, m), __fmul_rn(y, mm)), inv);
        m = __fadd_rn(m, mm);
    }
}

// delta[i] for the pair (i, i+1), i in [0, n-1); bit 7 = "too close" flag.  delta[n-1] = sentinel -1.
__global__ void bh_delta_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ sx,
                                const float* __restrict__ sy, int n, signed char* __restrict__ delta,
                                unsigned char* __restrict__ close) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (i == n - 1) { delta[i] = -1; close[i] = 0; return; }
    const unsigned long long x = keys[i] ^ keys[i + 1];
    const int d = x ? (__clzll(static_cast<long long>(x)) - (64 - kKeyBits)) >> 1 : kLevels;
    delta[i] = static_cast<signed char>(d);
    close[i] = (fabsf(sx[i] - sx[i + 1]) < kEps && fabsf(sy[i] - sy[i + 1]) < kEps) ? 1 : 0;
}

__global__ void bh_cap_kernel(const signed char* __restrict__ delta, const unsigned char* __restrict__ close, int n,
                              signed char* __restrict__ dcap, int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    auto cap_of = [&](int j) -> int {   // capped delta of pair j (j in [-1, n-1])
        if (j < 0 || j >= n - 1) return -1;
        int d = min(static_cast<int>(delta[j]), kLevels - 1);
        if (close[j]) {
            int p = j, q = j + 1;            // chain of close pairs containing pair j: bodies p..q
            for (int g = 0; g < 4096 && p > 0 && close[p - 1]; g++) p--;
            for (int g = 0; g < 4096 && q < n - 1 && close[q]; g++) q++;
            const int dl = p > 0 ? static_cast<int>(delta[p - 1]) : -1;
            const int dr = q < n - 1 ? static_cast<int>(delta[q]) : -1;
            d = min(d, max(dl, dr));         // l* - 1, l* = 1 + max(dl, dr)
        }
        return d;
    };
    const int c = cap_of(i), cl = cap_of(i - 1);
    dcap[i] = static_cast<signed char>(c);
    count[i] = max(0, c - cl);
}

struct BuildArgs {
    const unsigned long long* keys;
    const float *sx, *sy, *sm;
    const double* p3;
    const signed char* dcap;
    const int* base;      // exclusive prefix of count[]
    float4* nblk;     // block-SoA node records: block b = {x[4], y[4], m[4], s[4]} at nblk[4b .. 4b+3]
    int4* ncblk;      // child block index of each of the 4 nodes of block b (-1 = leaf)
    int n, cap_interior;
};

__device__ __forceinline__ int interior_id(const BuildArgs& a, int first, int level) {
    // id of the interior node that starts at body `first` on `level`
    const int dl = first > 0 ? static_cast<int>(a.dcap[first - 1]) : -1;
    return a.base[first] + (level - dl - 1);
}

// owner[id] = first body of interior node id (one thread per body; bodies start 0.76 nodes on average)
__global__ void bh_owner_kernel(const BuildArgs a, int* __restrict__ owner, BhStatus* st) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int dhi = a.dcap[i];
    const int dlo = i > 0 ? static_cast<int>(a.dcap[i - 1]) : -1;
    const int b = a.base[i];
    for (int k = 0; k < dhi - dlo && b + k < a.cap_interior; k++) owner[b + k] = i;
    if (i == 0) {
        const int total = a.base[a.n - 1];   // count[n-1] == 0
        st->node_count = 1 + 4 * min(total, a.cap_interior);
        st->n_interior = min(total, a.cap_interior);
        if (total > a.cap_interior) st->overflow = 1;
        if (dhi < 0) {
            // no interior node at all: the root is a leaf (one body, or everything merges)
            float cx = 0.f, cy = 0.f, mm = 0.f;
            for (int k = 0; k < a.n; k++) add_mass_ref(cx, cy, mm, a.sx[k], a.sy[k], a.sm[k]);
            a.nblk[0] = make_float4(cx, 0.f, 0.f, 0.f);
            a.nblk[1] = make_float4(cy, 0.f, 0.f, 0.f);
            a.nblk[2] = make_float4(mm, 0.f, 0.f, 0.f);
            a.nblk[3] = make_float4(-1.f, -1.f, -1.f, -1.f);
            a.ncblk[0] = make_int4(-1, -1, -1, -1);
        }
    }
}

// one thread per interior node
__global__ void bh_emit_kernel(const BuildArgs a, const int* __restrict__ owner, const BhStatus* st) {
    const int T = st->n_interior;
    const size_t stride = static_cast<size_t>(a.n) + 1;
    const float rx1 = ord2f(st->aabb_enc[0]), ry1 = ord2f(st->aabb_enc[1]);
    const float rx2 = ord2f(st->aabb_enc[2]), ry2 = ord2f(st->aabb_enc[3]);
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < T; id += gridDim.x * blockDim.x) {
        const int i = owner[id];
        const int dlo = i > 0 ? static_cast<int>(a.dcap[i - 1]) : -1;
        const int l = dlo + 1 + (id - a.base[i]);
        const unsigned long long key = a.keys[i];
        // cell bounds at level l: replay the key bits through rs-src/nbody.rs:286-301
        float x1 = rx1, y1 = ry1, x2 = rx2, y2 = ry2;
        for (int t = 0; t < l; t++) {
            const unsigned q = static_cast<unsigned>(key >> (2 * (kLevels - 1 - t))) & 3u;
            const float cx = __fmul_rn(__fadd_rn(x1, x2), 0.5f), cy = __fmul_rn(__fadd_rn(y1, y2), 0.5f);
            if (q & 1u) x1 = cx; else x2 = cx;      // UR/LR: right half
            if (q & 2u) y2 = cy; else y1 = cy;      // LL/LR: lower half
        }
        const int shift = 2 * (kLevels - 1 - l);
        // end of this cell's range: first body whose level-l prefix differs (galloping, then bisection)
        int end;
        if (l == 0) {
            end = a.n;
        } else {
            const unsigned long long pre = key >> (shift + 2);
            int lo = i + 1, step = 2;
            int hi = min(a.n, lo + step);
            while (hi < a.n && (a.keys[hi - 1] >> (shift + 2)) == pre) { lo = hi; step <<= 1; hi = min(a.n, lo + step); }
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if ((a.keys[mid] >> (shift + 2)) == pre) lo = mid + 1; else hi = mid;
            }
            end = lo;
        }
        const int e1 = lower_bound_quadrant(a.keys, i, end, shift, 1u);
        const int e2 = lower_bound_quadrant(a.keys, e1, end, shift, 2u);
        const int e3 = lower_bound_quadrant(a.keys, e2, end, shift, 3u);
        const int f[5] = {i, e1, e2, e3, end};
        const float cx = __fmul_rn(__fadd_rn(x1, x2), 0.5f);
        const int blk = 1 + id;   // block 0 holds the root (slot 0; slots 1..3 are empty leaves)
        if (l == 0) {   // the root's own record
            const double M = a.p3[a.n] - a.p3[0];
            a.nblk[0] = make_float4(static_cast<float>((a.p3[stride + a.n] - a.p3[stride]) / M), 0.f, 0.f, 0.f);
            a.nblk[1] = make_float4(static_cast<float>((a.p3[2 * stride + a.n] - a.p3[2 * stride]) / M), 0.f, 0.f, 0.f);
            a.nblk[2] = make_float4(static_cast<float>(M), 0.f, 0.f, 0.f);
            a.nblk[3] = make_float4(__fsub_rn(x2, x1), -1.f, -1.f, -1.f);
            a.ncblk[0] = make_int4(blk, -1, -1, -1);
        }
        float4 rec[4];
        int4 ch = make_int4(-1, -1, -1, -1);
        int* chp = &ch.x;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int first = f[q], cnt = f[q + 1] - f[q];
            rec[q] = make_float4(0.f, 0.f, 0.f, -1.0f);
            if (cnt == 1) {
                rec[q] = make_float4(a.sx[first], a.sy[first], a.sm[first], -1.0f);   // exact copy (:305-311)
            } else if (cnt >= 2) {
                int cid = -1;
                if (static_cast<int>(a.dcap[first]) >= l + 1) {
                    cid = interior_id(a, first, l + 1);
                    if (cid >= a.cap_interior) cid = -1;    // node pool exhausted: degrade to a merged leaf
                }
                if (cid >= 0) {
                    const double M = a.p3[first + cnt] - a.p3[first];
                    const double MX = a.p3[stride + first + cnt] - a.p3[stride + first];
                    const double MY = a.p3[2 * stride + first + cnt] - a.p3[2 * stride + first];
                    // child cell width s = x2 - x1 (rs-src/nbody.rs:341), children per :295-300
                    const float s = (q & 1) ? __fsub_rn(x2, cx) : __fsub_rn(cx, x1);
                    rec[q] = make_float4(static_cast<float>(MX / M), static_cast<float>(MY / M), static_cast<float>(M), s);
                    chp[q] = 1 + cid;
                } else {
                    float bx = 0.f, by = 0.f, bm = 0.f;
                    for (int k = 0; k < cnt; k++) add_mass_ref(bx, by, bm, a.sx[first + k], a.sy[first + k], a.sm[first + k]);
                    rec[q] = make_float4(bx, by, bm, -1.0f);
                }
            }
        }
        a.nblk[4 * blk + 0] = make_float4(rec[0].x, rec[1].x, rec[2].x, rec[3].x);
        a.nblk[4 * blk + 1] = make_float4(rec[0].y, rec[1].y, rec[2].y, rec[3].y);
        a.nblk[4 * blk + 2] = make_float4(rec[0].z, rec[1].z, rec[2].z, rec[3].z);
        a.nblk[4 * blk + 3] = make_float4(rec[0].w, rec[1].w, rec[2].w, rec[3].w);
        a.ncblk[blk] = ch;
    }
}

// ---- EXACT build: literal serial restatement of rs-src/nbody.rs:226-320, one thread ---------------------
__global__ void bh_build_serial_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                       const float* __restrict__ m, int n, int cap_nodes, float4* ndata,
                                       float4* nbounds, int* nchild, BhStatus* st) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    int node_count = 1;
    nbounds[0] = make_float4(ord2f(st->aabb_enc[0]), ord2f(st->aabb_enc[1]), ord2f(st->aabb_enc[2]), ord2f(st->aabb_enc[3]));
    ndata[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    nchild[0] = -1;
    for (int i = 0; i < n; i++) {
        const float px = x[i], py = y[i], pm = m[i];
        int node = 0;
        unsigned depth = 0;
        while (true) {
            if (depth > 50) { st->depth_error = 1; break; }       // :230-232 (the reference panics)
            float4 d = ndata[node];
            const int ch = nchild[node];
            if (ch >= 0) {                                          // :234-247 interior
                add_mass_ref(d.x, d.y, d.z, px, py, pm);
                ndata[node] = d;
                const float4 b = nbounds[node];
                const float cx = __fmul_rn(__fadd_rn(b.x, b.z), 0.5f);
                const float cy = __fmul_rn(__fadd_rn(b.y, b.w), 0.5f);
                const int q = (py < cy) ? ((px < cx) ? 2 : 3) : ((px < cx) ? 0 : 1);
                node = ch + q;
                depth++;
                continue;
            }
            const bool too_close = fabsf(__fsub_rn(d.x, px)) < kEps && fabsf(__fsub_rn(d.y, py)) < kEps;  // :249
            if (d.z == 0.0f || too_close) {                         // :250-260
                add_mass_ref(d.x, d.y, d.z, px, py, pm);
                ndata[node] = d;
                break;
            }
            // :261-282 split: clear, create children, re-insert the original, then keep inserting
            if (node_count + 4 > cap_nodes) { st->overflow = 1; break; }
            const float4 b = nbounds[node];
            const float cx = __fmul_rn(__fadd_rn(b.x, b.z), 0.5f);
            const float cy = __fmul_rn(__fadd_rn(b.y, b.w), 0.5f);
            const int c0 = node_count;
            node_count += 4;
            nbounds[c0 + 0] = make_float4(b.x, cy, cx, b.w);
            nbounds[c0 + 1] = make_float4(cx, cy, b.z, b.w);
            nbounds[c0 + 2] = make_float4(b.x, b.y, cx, cy);
            nbounds[c0 + 3] = make_float4(cx, b.y, b.z, cy);
            for (int k = 0; k < 4; k++) { ndata[c0 + k] = make_float4(0.f, 0.f, 0.f, 0.f); nchild[c0 + k] = -1; }
            nchild[node] = c0;
            // self.insert(original, depth+1): interior now, mass 0 -> exact copy, then the empty child takes it
            if (depth + 1 > 50 || depth + 2 > 50) { st->depth_error = 1; break; }
            const int qo = (d.y < cy) ? ((d.x < cx) ? 2 : 3) : ((d.x < cx) ? 0 : 1);
            ndata[c0 + qo] = make_float4(d.x, d.y, d.z, 0.f);
            ndata[node] = make_float4(d.x, d.y, d.z, 0.f);
            depth++;                                                // self.insert(px,py,m, depth+1) on this node
        }
        if (st->depth_error || st->overflow) break;
    }
    st->node_count = node_count;
}

__global__ void bh_finalize_exact_kernel(int cap_nodes, float4* ndata, const float4* __restrict__ nbounds,
                                         const int* __restrict__ nchild, const BhStatus* st) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= st->node_count || i >= cap_nodes) return;
    float4 d = ndata[i];
    const float4 b = nbounds[i];
    d.w = nchild[i] >= 0 ? __fsub_rn(b.z, b.x) : -1.0f;
    ndata[i] = d;
}

// ---- traversal, FAST ---------------------------------------------------------------------------------------
// One warp per 32 Morton-consecutive bodies.  Stack entries are (node block, mask of lanes that must look at
// it).  A block is the four children of an opened node in SoA form (x[4] y[4] m[4] s[4], 64 bytes): the warp
// evaluates all four branch-free with packed FP32 -- a lane in the mask interacts with a child if it is a
// leaf or passes the reference's per-body opening test (rs-src/nbody.rs:345), otherwise it asks for the
// child to be opened; only children that some lane must open are pushed, with that lane mask.  Every body
// therefore evaluates exactly the reference's interaction list (:333-377).  Empty leaves (m = 0, :367) and
// the body's own leaf (d = 0, :365) need no test: their contribution is an exact zero because EPS > 0.
template <bool COUNT>
__global__ void __launch_bounds__(kTravWarps * 32) bh_traverse_fast_kernel(
    const float4* __restrict__ nblk, const int4* __restrict__ ncblk, const float* __restrict__ sx,
    const float* __restrict__ sy, const int* __restrict__ idx_sorted, const int* __restrict__ mine, int n_list,
    int i_begin, float theta2, float2* __restrict__ acc, BhStatus* st) {
    __shared__ uint2 stk[kTravWarps][kStackPerWarp];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * kTravWarps + warp;
    const int li = w * 32 + lane;
    if (w * 32 >= n_list) return;
    const bool live = li < n_list;
    const int pos = live ? (mine ? mine[li] : li) : 0;  // position in sorted order
    const float px = sx[pos], py = sy[pos];
    const float2 npx = make_float2(-px, -px), npy = make_float2(-py, -py);
    const float2 eps2 = make_float2(kEps, kEps), th2 = make_float2(theta2, theta2);
    float2 ax = make_float2(0.f, 0.f), ay = make_float2(0.f, 0.f);
    uint2* s = stk[warp];
    int sp = 0;
    unsigned long long n_int = 0, n_vis = 0;
    {
        const unsigned m0 = __ballot_sync(0xffffffffu, live);
        if (lane == 0) s[0] = make_uint2(0u, m0);   // block 0 = {root, empty, empty, empty}
        sp = 1;
        __syncwarp();
        if (COUNT) n_vis -= live ? 3 : 0;           // the three padding slots of block 0 are not nodes
    }
    while (sp > 0) {
        --sp;
        const uint2 e = s[sp];
        __syncwarp();
        const bool act = (e.y >> lane) & 1u;
        const float4 X = __ldg(&nblk[4 * e.x + 0]), Y = __ldg(&nblk[4 * e.x + 1]);
        const float4 M = __ldg(&nblk[4 * e.x + 2]), S = __ldg(&nblk[4 * e.x + 3]);
        const float2 dx01 = __fadd2_rn(make_float2(X.x, X.y), npx), dx23 = __fadd2_rn(make_float2(X.z, X.w), npx);
        const float2 dy01 = __fadd2_rn(make_float2(Y.x, Y.y), npy), dy23 = __fadd2_rn(make_float2(Y.z, Y.w), npy);
        const float2 d01 = __ffma2_rn(dy01, dy01, __fmul2_rn(dx01, dx01));
        const float2 d23 = __ffma2_rn(dy23, dy23, __fmul2_rn(dx23, dx23));
        const float2 e01 = __fadd2_rn(d01, eps2), e23 = __fadd2_rn(d23, eps2);
        // opening test s/d < theta  <=>  s^2 < theta^2 d^2   (s < 0 marks a leaf)
        const float2 t01 = __fmul2_rn(th2, d01), t23 = __fmul2_rn(th2, d23);
        const float2 q01 = __fmul2_rn(make_float2(S.x, S.y), make_float2(S.x, S.y));
        const float2 q23 = __fmul2_rn(make_float2(S.z, S.w), make_float2(S.z, S.w));
        const bool l0 = S.x < 0.f, l1 = S.y < 0.f, l2 = S.z < 0.f, l3 = S.w < 0.f;
        const bool a0 = q01.x < t01.x, a1 = q01.y < t01.y, a2 = q23.x < t23.x, a3 = q23.y < t23.y;
        const bool u0 = act && (l0 || a0), u1 = act && (l1 || a1), u2 = act && (l2 || a2), u3 = act && (l3 || a3);
        float2 c01 = __fmul2_rn(make_float2(M.x, M.y), make_float2(rcp_approx(e01.x), rcp_approx(e01.y)));
        float2 c23 = __fmul2_rn(make_float2(M.z, M.w), make_float2(rcp_approx(e23.x), rcp_approx(e23.y)));
        c01.x = u0 ? c01.x : 0.f; c01.y = u1 ? c01.y : 0.f;
        c23.x = u2 ? c23.x : 0.f; c23.y = u3 ? c23.y : 0.f;
        ax = __ffma2_rn(c01, dx01, ax); ay = __ffma2_rn(c01, dy01, ay);
        ax = __ffma2_rn(c23, dx23, ax); ay = __ffma2_rn(c23, dy23, ay);
        const unsigned ob = (act && !l0 && !a0 ? 1u : 0u) | (act && !l1 && !a1 ? 2u : 0u) |
                            (act && !l2 && !a2 ? 4u : 0u) | (act && !l3 && !a3 ? 8u : 0u);
        if (COUNT) {
            n_vis += act ? 4 : 0;
            n_int += (u0 && M.x != 0.f && !(X.x == px && Y.x == py)) + (u1 && M.y != 0.f && !(X.y == px && Y.y == py)) +
                     (u2 && M.z != 0.f && !(X.z == px && Y.z == py)) + (u3 && M.w != 0.f && !(X.w == px && Y.w == py));
        }
        const unsigned any = __reduce_or_sync(0xffffffffu, ob);
        if (any) {
            const int4 C = __ldg(&ncblk[e.x]);
            // push order 3..0 so that child 0 is opened first (DFS-like order)
            if (any & 8u) { const unsigned om = __ballot_sync(0xffffffffu, ob & 8u); if (lane == 0) s[sp] = make_uint2(C.w, om); sp++; }
            if (any & 4u) { const unsigned om = __ballot_sync(0xffffffffu, ob & 4u); if (lane == 0) s[sp] = make_uint2(C.z, om); sp++; }
            if (any & 2u) { const unsigned om = __ballot_sync(0xffffffffu, ob & 2u); if (lane == 0) s[sp] = make_uint2(C.y, om); sp++; }
            if (any & 1u) { const unsigned om = __ballot_sync(0xffffffffu, ob & 1u); if (lane == 0) s[sp] = make_uint2(C.x, om); sp++; }
        }
        __syncwarp();
    }
    if (live) acc[idx_sorted[pos] - i_begin] = make_float2(ax.x + ax.y, ay.x + ay.y);
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            n_int += __shfl_xor_sync(0xffffffffu, n_int, o);
            n_vis += __shfl_xor_sync(0xffffffffu, n_vis, o);
        }
        if (lane == 0) { atomicAdd(&st->interactions, n_int); atomicAdd(&st->visited, n_vis); }
    }
}

// ---- traversal, EXACT: rs-src/nbody.rs:333-377 with its nested summation order ------------------------------
struct Frame { int node; int next; float fx, fy; };

__device__ __forceinline__ void force_ref(float px1, float py1, float m1, float px2, float py2, float m2, float& fx, float& fy) {
    const float dx = __fsub_rn(px2, px1), dy = __fsub_rn(py2, py1);              // rs-src/nbody.rs:174-175
    const float dist_sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
    const float f = __fdiv_rn(__fmul_rn(m1, m2), __fadd_rn(dist_sq, kEps));      // :180
    fx = __fmul_rn(f, dx); fy = __fmul_rn(f, dy);
}

__global__ void bh_traverse_exact_kernel(const float4* __restrict__ ndata, const int* __restrict__ nchild,
                                         const float* __restrict__ x, const float* __restrict__ y,
                                         const float* __restrict__ m, int i_begin, int n_local, float theta,
                                         float2* __restrict__ force_out) {
    const int il = blockIdx.x * blockDim.x + threadIdx.x;
    if (il >= n_local) return;
    const float px = x[i_begin + il], py = y[i_begin + il], pm = m[i_begin + il];
    Frame st[64];
    int sp = 0;
    st[0] = Frame{0, -1, 0.f, 0.f};
    float rx = 0.f, ry = 0.f;  // value returned by the frame that just finished
    while (sp >= 0) {
        Frame& f = st[sp];
        if (f.next < 0) {
            // first visit: evaluate the node
            const float4 nd = ndata[f.node];
            if (nd.w >= 0.f) {  // interior
                const float dx = __fsub_rn(nd.x, px), dy = __fsub_rn(nd.y, py);
                const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));  // :344
                if (__fdiv_rn(nd.w, d) < theta) {                                           // :345
                    force_ref(px, py, pm, nd.x, nd.y, nd.z, rx, ry);
                    sp--;
                } else {
                    f.next = 0; f.fx = 0.f; f.fy = 0.f;
                    const int c = nchild[f.node];
                    st[sp + 1] = Frame{c, -1, 0.f, 0.f};
                    f.next = 1;
                    sp++;
                    continue;
                }
            } else {  // leaf :363-374
                if ((nd.x == px && nd.y == py) || nd.z == 0.0f) { rx = 0.f; ry = 0.f; }
                else force_ref(px, py, pm, nd.x, nd.y, nd.z, rx, ry);
                sp--;
            }
        } else {
            // a child returned (rx,ry): fx += fx_add (:358-359)
            f.fx = __fadd_rn(f.fx, rx);
            f.fy = __fadd_rn(f.fy, ry);
            if (f.next < 4) {
                const int c = nchild[f.node] + f.next;
                f.next++;
                st[sp + 1] = Frame{c, -1, 0.f, 0.f};
                sp++;
                continue;
            }
            rx = f.fx; ry = f.fy;
            sp--;
        }
        if (sp < 0) break;
    }
    force_out[il] = make_float2(rx, ry);
}

// mine[] = sorted positions whose body index lies in [b, b+c)
struct InRange {
    const int* idx_sorted;
    int b, e;
    __device__ bool operator()(int pos) const { const int i = idx_sorted[pos]; return i >= b && i < e; }
};
using Iota = thrust::counting_iterator<int>;

static void check_status(Engine& e, struct BhWork& w, bool sync_now);

// ---- host orchestration ----------------------------------------------------------------------------------------
static void ensure_work(Engine& e, BhWork& w, int n) {
    if (n > w.cap_n) {
        NB_CUDA(cudaStreamSynchronize(e.stream));
        auto fr = [](void* p) { if (p) cudaFree(p); };
        fr(w.keys); fr(w.keys_sorted); fr(w.idx); fr(w.idx_sorted); fr(w.mine); fr(w.sx); fr(w.sy); fr(w.sm);
        fr(w.w3); fr(w.p3); fr(w.tile_sums); fr(w.ndata); fr(w.nbounds); fr(w.nchild); fr(w.nblk); fr(w.ncblk); fr(w.delta); fr(w.dcap); fr(w.close); fr(w.count); fr(w.base); fr(w.owner); fr(w.cub_tmp);
        const size_t N = static_cast<size_t>(n);
        NB_CUDA(cudaMalloc(&w.keys, N * 8)); NB_CUDA(cudaMalloc(&w.keys_sorted, N * 8));
        NB_CUDA(cudaMalloc(&w.idx, N * 4)); NB_CUDA(cudaMalloc(&w.idx_sorted, N * 4)); NB_CUDA(cudaMalloc(&w.mine, N * 4));
        NB_CUDA(cudaMalloc(&w.sx, N * 4)); NB_CUDA(cudaMalloc(&w.sy, N * 4)); NB_CUDA(cudaMalloc(&w.sm, N * 4));
        NB_CUDA(cudaMalloc(&w.w3, 3 * (N + 1) * 8)); NB_CUDA(cudaMalloc(&w.p3, 3 * (N + 1) * 8));
        w.cap_nodes = static_cast<int>(std::min<size_t>(8 * N + 4096, 0x3fffff0));
        NB_CUDA(cudaMalloc(&w.ndata, sizeof(float4) * w.cap_nodes));
        NB_CUDA(cudaMalloc(&w.nbounds, sizeof(float4) * w.cap_nodes));
        NB_CUDA(cudaMalloc(&w.nchild, sizeof(int) * w.cap_nodes));
        NB_CUDA(cudaMalloc(&w.nblk, sizeof(float4) * w.cap_nodes));
        NB_CUDA(cudaMalloc(&w.ncblk, sizeof(int4) * (w.cap_nodes / 4)));
        NB_CUDA(cudaMalloc(&w.delta, N)); NB_CUDA(cudaMalloc(&w.dcap, N)); NB_CUDA(cudaMalloc(&w.close, N));
        NB_CUDA(cudaMalloc(&w.count, N * 4)); NB_CUDA(cudaMalloc(&w.base, N * 4));
        NB_CUDA(cudaMalloc(&w.owner, sizeof(int) * (w.cap_nodes / 4)));
        NB_CUDA(cudaMalloc(&w.tile_sums, 3 * ((N + 1 + kScanTile - 1) / kScanTile) * 8));
        size_t b1 = 0, b2 = 0, b3 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, b1, w.keys, w.keys_sorted, w.idx, w.idx_sorted, n, 0, kKeyBits, e.stream);
        cub::DeviceScan::ExclusiveSum(nullptr, b2, w.count, w.base, n, e.stream);
        cub::DeviceSelect::If(nullptr, b3, Iota(0), w.mine, &w.status->n_mine, n, InRange{w.idx_sorted, 0, n}, e.stream);
        w.cub_bytes = std::max(b1, std::max(b2, b3)) + 256;
        NB_CUDA(cudaMalloc(&w.cub_tmp, w.cub_bytes));
        w.cap_n = n;
    }
}

static void ensure_status(BhWork& w) {
    if (!w.status) {
        NB_CUDA(cudaMalloc(&w.status, sizeof(BhStatus)));
        NB_CUDA(cudaMallocHost(&w.status_host, sizeof(BhStatus)));
        memset(w.status_host, 0, sizeof(BhStatus));
        NB_CUDA(cudaEventCreateWithFlags(&w.status_ev, cudaEventDisableTiming));
    }
}

// positions of ALL bodies for the tree: the arena on one GPU, the gathered mirror when sharded
struct GlobalPos { const float *x, *y, *m; };
static GlobalPos global_positions(Engine& e) {
    if (!e.dist || e.world == 1) return GlobalPos{e.arena.x(e.lay, e.cur), e.arena.y(e.lay, e.cur), e.arena.m(e.lay)};
    PhaseScope ps(e, 7);
    dist_gather_mirror(e, e.cur);
    const size_t GL = static_cast<size_t>(e.world) * e.lay.L;
    return GlobalPos{e.mirror, e.mirror + GL, e.mirror + 2 * GL};
}

// builds the tree over all e.n bodies and leaves per-local-body acceleration (FAST) / force (EXACT) in w.acc
static void bh_forces(Engine& e, float theta) {
    BhWork& w = work(e);
    ensure_status(w);
    check_status(e, w, false);   // previous step's status (its copy has long finished by now)
    const int n = e.n;
    ensure_work(e, w, n);
    const int nl = local_count(e), ib = local_begin(e);
    if (nl > w.cap_acc) {
        if (w.acc) NB_CUDA(cudaFree(w.acc));
        NB_CUDA(cudaMalloc(&w.acc, sizeof(float2) * static_cast<size_t>(e.lay.L)));
        w.cap_acc = static_cast<int>(e.lay.L);
    }
    // sharded: slot index == body index only while every shard but the last is full (L-aligned shards)
    const GlobalPos gp = global_positions(e);
    cudaStream_t s = e.stream;
    const int T = 256, G = (n + T - 1) / T;
    {
        PhaseScope ps(e, 2);
        bh_reset_kernel<<<1, 32, 0, s>>>(w.status);
        bh_aabb_kernel<<<std::min(G, e.num_sms * 4), T, 0, s>>>(gp.x, gp.y, n, w.status);
        e.ctr.kernel_launches += 2;
    }
    if (e.mode == NBX_MODE_EXACT) {
        {
            PhaseScope ps(e, 5);
            bh_build_serial_kernel<<<1, 32, 0, s>>>(gp.x, gp.y, gp.m, n, w.cap_nodes, w.ndata, w.nbounds, w.nchild, w.status);
            bh_finalize_exact_kernel<<<(w.cap_nodes + T - 1) / T, T, 0, s>>>(w.cap_nodes, w.ndata, w.nbounds, w.nchild, w.status);
            e.ctr.kernel_launches += 2;
        }
        PhaseScope ps(e, 0);
        if (nl > 0) {
            bh_traverse_exact_kernel<<<(nl + 127) / 128, 128, 0, s>>>(w.ndata, w.nchild, gp.x, gp.y, gp.m, ib, nl, theta, w.acc);
            e.ctr.kernel_launches++;
        }
    } else {
        {
            PhaseScope ps(e, 3);
            bh_keys_kernel<<<G, T, 0, s>>>(gp.x, gp.y, n, w.status, w.keys, w.idx);
            e.ctr.kernel_launches++;
        }
        {
            PhaseScope ps(e, 4);
            size_t tb = w.cub_bytes;
            cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keys, w.keys_sorted, w.idx, w.idx_sorted, n, 0, kKeyBits, s);
        }
        {
            PhaseScope ps(e, 6);
            bh_gather_sorted_kernel<<<(n + 1 + T - 1) / T, T, 0, s>>>(gp.x, gp.y, gp.m, w.idx_sorted, n, w.sx, w.sy, w.sm, w.w3);
            e.ctr.kernel_launches++;
            const int len = n + 1, ntiles = (len + kScanTile - 1) / kScanTile;
            const size_t stride = static_cast<size_t>(n) + 1;
            scan_tile_sums_kernel<<<dim3(ntiles, 3), kScanThreads, 0, s>>>(w.w3, len, stride, w.tile_sums, ntiles);
            scan_tile_offsets_kernel<<<3, kScanThreads, 0, s>>>(w.tile_sums, ntiles);
            scan_apply_kernel<<<dim3(ntiles, 3), kScanThreads, 0, s>>>(w.w3, w.p3, len, stride, w.tile_sums, ntiles);
            e.ctr.kernel_launches += 3;
        }
        {
            PhaseScope ps(e, 5);
            bh_delta_kernel<<<G, T, 0, s>>>(w.keys_sorted, w.sx, w.sy, n, w.delta, w.close);
            bh_cap_kernel<<<G, T, 0, s>>>(w.delta, w.close, n, w.dcap, w.count);
            size_t tb = w.cub_bytes;
            cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.count, w.base, n, s);   // integer: deterministic
            BuildArgs ba{w.keys_sorted, w.sx, w.sy, w.sm, w.p3, w.dcap, w.base, w.nblk, w.ncblk, n, (w.cap_nodes - 4) / 4};
            bh_owner_kernel<<<G, T, 0, s>>>(ba, w.owner, w.status);
            bh_emit_kernel<<<std::min(G, e.num_sms * 8), T, 0, s>>>(ba, w.owner, w.status);
            e.ctr.kernel_launches += 4;
        }
        const int* mine = nullptr;
        int n_list = n;
        if (e.dist && e.world > 1) {
            size_t tb = w.cub_bytes;
            cub::DeviceSelect::If(w.cub_tmp, tb, Iota(0), w.mine, &w.status->n_mine, n, InRange{w.idx_sorted, ib, ib + nl}, s);
            mine = w.mine;
            n_list = nl;
        }
        PhaseScope ps(e, 0);
        if (n_list > 0) {
            const int blocks = (n_list + kTravWarps * 32 - 1) / (kTravWarps * 32);
            if (e.bh_count)
                bh_traverse_fast_kernel<true><<<blocks, kTravWarps * 32, 0, s>>>(w.nblk, w.ncblk, w.sx, w.sy, w.idx_sorted, mine,
                                                                                n_list, ib, theta * theta, w.acc, w.status);
            else
                bh_traverse_fast_kernel<false><<<blocks, kTravWarps * 32, 0, s>>>(w.nblk, w.ncblk, w.sx, w.sy, w.idx_sorted, mine,
                                                                                 n_list, ib, theta * theta, w.acc, w.status);
            e.ctr.kernel_launches++;
        }
    }
    NB_CUDA(cudaGetLastError());
    NB_CUDA(cudaMemcpyAsync(w.status_host, w.status, sizeof(BhStatus), cudaMemcpyDeviceToHost, s));
    NB_CUDA(cudaEventRecord(w.status_ev, s));
    w.status_pending = true;
}

// Deferred: the status block of a step is copied to pinned memory behind the step and looked at when the
// NEXT Barnes-Hut call (or any synchronising call) arrives, so stepping never stalls the host on the GPU.
static void check_status(Engine& e, BhWork& w, bool sync_now) {
    if (!w.status_pending) return;
    if (sync_now) NB_CUDA(cudaEventSynchronize(w.status_ev));
    else if (cudaEventQuery(w.status_ev) != cudaSuccess) { (void)cudaGetLastError(); NB_CUDA(cudaEventSynchronize(w.status_ev)); }
    w.status_pending = false;
    const BhStatus& h = *w.status_host;
    if (h.depth_error) fatal("Node::insert() - recursion depth > 50 (the reference panics here, rs-src/nbody.rs:230-232)", __FILE__, __LINE__);
    if (h.overflow && !w.warned) {
        fprintf(stderr, "nbody_b200: warning: quadtree node pool exhausted (%d nodes); deepest cells were merged\n", w.cap_nodes);
        w.warned = true;
    }
    e.ctr.bh_nodes_built += static_cast<uint64_t>(h.node_count);
    e.ctr.bh_interactions += h.interactions;
    e.ctr.bh_nodes_visited += h.visited;
}
void bh_poll(Engine& e) {
    if (e.bh) check_status(e, work(e), true);
}

void bh_step(Engine& e, float theta, float dt) {
    if (e.n == 0) return;
    BhWork& w = work(e);
    bh_forces(e, theta);
    {
        PhaseScope ps(e, 1);
        if (e.mode == NBX_MODE_EXACT) launch_integrate_exact(e, w.acc, dt, true);
        else launch_integrate_fast(e, w.acc, 1, dt, true);
    }
    e.step_count++;
    dist_signal_step_done(e);
    e.ctr.steps++;
    if (e.phase_timing) e.ev_slot++;
}

void bh_accelerations(Engine& e, float theta, float2* out) {
    if (e.n == 0) return;
    BhWork& w = work(e);
    bh_forces(e, theta);
    if (e.mode == NBX_MODE_EXACT) launch_accel_from_force(e, w.acc, out);
    else launch_accel_from_partial(e, w.acc, 1, out);
    e.step_count++;
    dist_signal_step_done(e);
    check_status(e, w, true);
}

void bh_shutdown(Engine& e) {
    if (!e.bh) return;
    BhWork& w = work(e);
    auto fr = [](void* p) { if (p) cudaFree(p); };
    fr(w.keys); fr(w.keys_sorted); fr(w.idx); fr(w.idx_sorted); fr(w.mine); fr(w.sx); fr(w.sy); fr(w.sm);
    fr(w.w3); fr(w.p3); fr(w.tile_sums); fr(w.ndata); fr(w.nbounds); fr(w.nchild); fr(w.nblk); fr(w.ncblk); fr(w.delta); fr(w.dcap); fr(w.close); fr(w.count); fr(w.base); fr(w.owner); fr(w.cub_tmp); fr(w.status); fr(w.acc);
    if (w.status_host) cudaFreeHost(w.status_host);
    if (w.status_ev) cudaEventDestroy(w.status_ev);
    delete &w;
    e.bh = nullptr;
}

}  // namespace nb
