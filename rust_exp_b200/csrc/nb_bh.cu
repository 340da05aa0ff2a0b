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
#include <vector>
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
    int n_part;               // partitioned step: bodies this part received (device-side size of its arrays)
    unsigned delta_hist[kLevels + 1];   // delta_hist[d] = sorted neighbours sharing exactly d key levels (sizes the next sort)
    int sort_long_run;        // longest run the sort fix-up had to order (> 4096 only)
    int n_interior;
    int n_deep;               // partitioned build: interior nodes at or below the cut level
    int n_top;                // partitioned build: interior nodes of the shared top tree
    unsigned long long interactions;
    unsigned long long visited;
    // walk scheduling + lane-efficiency instrumentation (COUNT builds of the walk only)
    int tickets[kMaxRanks + 1];           // next group of 32 bodies per walk launch (dynamic scheduling)
    unsigned long long pops;              // stack entries popped by all warps
    unsigned long long pop_lanes;         // sum of popcount(lane mask) over those pops
    unsigned long long pop_hist[33];      // histogram of popcount(lane mask)
};

constexpr int kCutLevel = kBhCutLevel;
constexpr int kNumCells = kBhNumCells;                   // 1024
constexpr int kTopNodes = (kNumCells * 4 - 1) / 3;       // levels 0..kCutLevel: 1365
constexpr int kPartShift = 22;                           // 4M blocks per part
constexpr int kCellShift = 2 * (kLevels - kCutLevel);

struct __align__(16) CellEntry {
    double M, MX, MY;
    int count;
    int child;        // global block index of the cell's child block (interior cell), else -1
    float x, y, m;    // leaf record (single body, or merged group)
    int pad;
};
static_assert(sizeof(CellEntry) <= 64, "the BH arena reserves 64 bytes per cell entry");

// The partition of the cut-level cells over the parts: part g owns cells [cut[g], cut[g+1]).  Device-resident and
// identical on every rank (computed by every rank from the same published cell table, one step ahead).
struct PartPlan {
    int cut[kMaxRanks + 1];
};

// One part of the domain-partitioned step = one rank (real ranks: this process holds exactly one; virtual ranks on a
// single GPU, nbx_bh_partition(G): this process holds all G and "peer" pointers are simply the other parts' buffers).
struct PartBufs {
    bool ready = false;
    size_t cap = 0;                   // body capacity of the destination-side arrays
    int cap_blocks = 0;
    size_t R = 0;                     // inbox region capacity per source
    // source side: this part's index shard, keyed and sorted locally
    int src_cap = 0;
    unsigned long long *keys = nullptr, *keys_sorted = nullptr;
    unsigned *k32 = nullptr, *k32_sorted = nullptr;
    int *idx = nullptr, *idx_sorted = nullptr, *idx_fixed = nullptr;
    // "arena" (peer-visible): real ranks -> pointers into the IPC arena; virtual ranks -> one cudaMalloc with the same layout
    char* arena = nullptr;
    bool arena_owned = false;
    // destination side: the merged (key-sorted) bodies of this part's cells, and the build workspace over them
    unsigned long long* mkeys = nullptr;
    float *sx = nullptr, *sy = nullptr, *sm = nullptr;
    int* gidx = nullptr;
    double *w3 = nullptr, *p3 = nullptr, *tile_sums = nullptr;
    signed char *delta = nullptr, *dcap = nullptr;
    unsigned char* close = nullptr;
    int *count = nullptr, *base = nullptr, *owner = nullptr, *itile = nullptr;
    BhStatus* status = nullptr;       // per part: n_part, n_interior, aabb, overflow ...
};

struct TopBufs {
    int* tcount = nullptr;
    double* tm3 = nullptr;        // [3][kTopNodes]
    float4* tleaf = nullptr;      // x, y, m, (unused)
    int* tchild = nullptr;        // level-kCutLevel nodes: child block or -1
    float4* blk = nullptr;        // top blocks: 1 + (kTopNodes - kNumCells) blocks
    int4* cblk = nullptr;
    PartPlan* plan = nullptr;     // [2]: plan[e & 1] is the partition step e uses; step e writes plan[(e + 1) & 1]
    uint32_t* epoch_dev = nullptr; // the step counter e of the partitioned path (device-resident: graphs replay unchanged)
    bool plan_valid = false;
    int plan_n = -1, plan_parts = -1;
};

// Block index space: (part << shift) | block.  One tree: shift = 31, part 0.  Partitioned trees: part g < world
// is rank g's subtree forest (possibly in a PEER GPU's memory, walked in place over NVLink), part kMaxRanks is
// the small shared top tree.  Accelerations are scattered to the rank that owns the body's index shard.
struct TreeTable {
    const float4* blk[kMaxRanks + 1];
    const int4* cblk[kMaxRanks + 1];
    float2* acc[kMaxRanks];
    int shift;
    unsigned root;
    int shard_len;     // L: owner of body i is i / L, its slot i % L
};

struct NodeInfo { int first, end, blk, slot_level; };   // slot_level = slot | (level << 8); own record = nblk[blk].slot
struct PeerU32 { uint32_t* p[kMaxRanks]; };

struct BhWork {
    int cap_n = 0;
    unsigned long long *keys = nullptr, *keys_sorted = nullptr;
    int *idx = nullptr, *idx_sorted = nullptr, *mine = nullptr;
    const int* order = nullptr;   // final sorted order of the last single-tree build (idx_sorted, or idx_l after a partial sort)
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
    // EXACT parallel build
    unsigned *xk = nullptr, *xk_sorted = nullptr;
    int *xorder = nullptr, *idx_l = nullptr, *xflags = nullptr, *xflags_host = nullptr;
    unsigned long long* lk_tmp = nullptr;
    NodeInfo* info = nullptr;
    BhStatus* status = nullptr;
    // Two pinned status slots (slot = position-buffer parity at the start of the step): a step's status block is copied
    // behind it and read when the slot comes round again, so the host runs up to two steps ahead of the GPU.
    struct StatusSlot {
        BhStatus* host = nullptr;
        cudaEvent_t ev = nullptr;
        bool pending = false;
        bool partitioned = false;
        int n = 0;
    } slot[2];
    int cur_slot = 0;                // slot of the step being enqueued / most recently enqueued
    BhStatus* status_host = nullptr; // == slot[cur_slot].host
    int last_nparts_hint = 1;        // parts the last partitioned step split into (per-part body count for the sort sizing)
    int sort_levels = kLevels;       // key levels the next sort orders on (from the delta histogram of earlier steps)
    uint64_t seen_set_gen = ~0ull;   // a replaced particle set resets sort_levels: no history to trust
    float2* acc = nullptr;     // per local body: acceleration (FAST) or force (EXACT)
    int cap_acc = 0;
    bool warned = false;
    // CUDA graph of the single-GPU FAST step (one per position-buffer parity): ~20 short launches replayed as one
    struct GraphSlot {
        cudaGraphExec_t exec = nullptr;
        int n = -1, cur = -1;
        float theta = 0.f, dt = 0.f;
        cudaStream_t stream = nullptr;
        uint64_t launches = 0;
        const char* arena = nullptr;
        uint64_t alloc_gen = 0;
        size_t L = 0;
        int sort_levels = 0;
        float kick = 0.f;
        int square = 0;
        int nparts = 0;                  // 1 = single tree, > 1 = domain-partitioned step
        bool partitioned = false;
        float2* acc_src = nullptr;       // host-side results of the capture, restored on every replay
        TreeTable tt{};
    } graph[2];
    uint64_t alloc_gen = 0;          // bumped whenever a buffer a captured graph points at is (re)allocated
    bool capturing = false;
    // domain-partitioned mode
    std::vector<PartBufs> parts;
    TopBufs top;
    uint32_t bh_epoch = 0;
    uint32_t sync_test_epoch = 0;
    float2* acc_src = nullptr;       // where this step's per-local-body accelerations ended up
    TreeTable last_tt{};             // the tree of the most recent FAST step (for nbx_bh_flatten)
    int last_nparts = 0;             // 0 = no FAST tree yet, 1 = single tree, >1 = partitioned
    bool last_partitioned = false;
    uint64_t pop_hist[33] = {};      // lane-mask popcount histogram of the walk (counting mode), since the last read
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
        st->n_deep = 0;
        st->n_top = 0;
        for (int d = 0; d <= kLevels; d++) st->delta_hist[d] = 0u;
        st->sort_long_run = 0;
        for (int g = 0; g <= kMaxRanks; g++) st->tickets[g] = 0;
        st->pops = 0;
        st->pop_lanes = 0;
        for (int k = 0; k < 33; k++) st->pop_hist[k] = 0;
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

// opt-in (nbx_set_square_aabb): the reference author's commented-out variant, rs-src/nbody.rs:400-407
//     if x2 - x1 > y2 - y1 { y2 = y1 + (x2 - x1) } else { x2 = x1 + (y2 - y1) }
__device__ __forceinline__ void square_up(int* enc) {
    const float x1 = ord2f(enc[0]), y1 = ord2f(enc[1]), x2 = ord2f(enc[2]), y2 = ord2f(enc[3]);
    if (__fsub_rn(x2, x1) > __fsub_rn(y2, y1)) enc[3] = f2ord(__fadd_rn(y1, __fsub_rn(x2, x1)));
    else enc[2] = f2ord(__fadd_rn(x1, __fsub_rn(y2, y1)));
}
__global__ void bh_square_aabb_kernel(BhStatus* st) {
    if (threadIdx.x == 0 && blockIdx.x == 0) square_up(st->aabb_enc);
}

// ---- keys: the reference's midpoint recursion, kLevels deep ------------------------------------------
__global__ void bh_keys_kernel(const float* __restrict__ x, const float* __restrict__ y, int n, const BhStatus* st,
                               unsigned long long* __restrict__ keys, int* __restrict__ idx, unsigned* __restrict__ k32 = nullptr) {
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
    if (k32) k32[i] = static_cast<unsigned>(key >> (kKeyBits - 32));   // the top 16 levels
}

// ---- sort: only the key levels the tree needs ------------------------------------------------------------------------
// The tree of a set whose deepest split is at level D only needs the bodies ordered by the top D+1 key levels; bodies
// move less than a cell per step, so last step's depth (+ margin) tells how many levels to sort on: 3-4 radix passes
// over 32-bit keys instead of 6 over 64-bit ones.  Correctness never depends on the guess: bh_sort_fixup_kernel
// completes the order inside every run of bodies that share the sorted prefix (normally none or pairs).
struct SortBufs {
    unsigned long long *keys, *keys_sorted;   // full 48-bit keys (unsorted) / CUB output of the 64-bit path
    unsigned *k32, *k32_sorted;               // top 32 bits
    int *idx, *idx_sorted, *idx_fixed;        // identity, CUB output, final order
};

// prefix (the sorted-on levels) of sorted position i
struct SortedPrefix {
    const unsigned long long* k64;
    const unsigned* k32;
    int shift;
    __device__ __forceinline__ unsigned long long operator()(int i) const {
        return k32 ? static_cast<unsigned long long>(k32[i] >> shift) : (k64[i] >> shift);
    }
};

// every body finds its final position inside its run of equal prefixes: rank by (full key, sorted position) -- stable
__global__ void bh_sort_fixup_kernel(SortedPrefix pre, const unsigned long long* __restrict__ keys, const int* __restrict__ idx_sorted, int n,
                                     int* __restrict__ idx_fixed, BhStatus* st) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const unsigned long long p = pre(i);
        const bool same_l = i > 0 && pre(i - 1) == p, same_r = i + 1 < n && pre(i + 1) == p;
        const int me = idx_sorted[i];
        if (!same_l && !same_r) { idx_fixed[i] = me; continue; }
        int lo = i, hi = i + 1;   // run [lo, hi): galloping + bisection on the sorted prefixes
        if (same_l) {
            int step = 1, l = i - 1;
            while (l - step >= 0 && pre(l - step) == p) { l -= step; step <<= 1; }
            int a = max(l - step, -1), b = l;          // pre(a) != p (or a == -1), pre(b) == p
            while (b - a > 1) { const int mid = (a + b) >> 1; if (pre(mid) == p) b = mid; else a = mid; }
            lo = b;
        }
        if (same_r) {
            int step = 1, r = i + 1;
            while (r + step < n && pre(r + step) == p) { r += step; step <<= 1; }
            int a = r, b = min(r + step, n);            // pre(a) == p, pre(b) != p (or b == n)
            while (b - a > 1) { const int mid = (a + b) >> 1; if (pre(mid) == p) a = mid; else b = mid; }
            hi = a + 1;
        }
        const unsigned long long k = keys[me];
        int rank = 0;
        for (int j = lo; j < hi; j++) {
            const unsigned long long kj = keys[idx_sorted[j]];
            rank += (kj < k || (kj == k && j < i)) ? 1 : 0;
        }
        idx_fixed[lo + rank] = me;
        if (i == lo && hi - lo > 4096) atomicMax(&st->sort_long_run, hi - lo);   // the depth guess was far off: tell the host
    }
}

// sorts (key, index) on the top `levels` key levels and returns the array holding the final order
static const int* sort_bodies(Engine& e, BhWork& w, const SortBufs& b, int n, int levels, BhStatus* st) {
    cudaStream_t s = e.stream;
    size_t tb = w.cub_bytes;
    if (levels >= kLevels) {
        cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, b.keys, b.keys_sorted, b.idx, b.idx_sorted, n, 0, kKeyBits, s);
        return b.idx_sorted;
    }
    SortedPrefix pre{};
    if (levels <= 16) {
        cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, b.k32, b.k32_sorted, b.idx, b.idx_sorted, n, 32 - 2 * levels, 32, s);
        pre.k32 = b.k32_sorted; pre.shift = 32 - 2 * levels;
    } else {
        cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, b.keys, b.keys_sorted, b.idx, b.idx_sorted, n, kKeyBits - 2 * levels, kKeyBits, s);
        pre.k64 = b.keys_sorted; pre.shift = kKeyBits - 2 * levels;
    }
    bh_sort_fixup_kernel<<<std::max(1, std::min((n + 255) / 256, e.num_sms * 16)), 256, 0, s>>>(pre, b.keys, b.idx_sorted, n, b.idx_fixed, st);
    e.ctr.kernel_launches++;
    return b.idx_fixed;
}

__global__ void bh_gather_sorted_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                        const float* __restrict__ m, const int* __restrict__ idx_sorted, int n,
                                        float* __restrict__ sx, float* __restrict__ sy, float* __restrict__ sm,
                                        double* __restrict__ w3, const unsigned long long* __restrict__ keys = nullptr,
                                        unsigned long long* __restrict__ keys_sorted = nullptr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    double wm = 0.0, wx = 0.0, wy = 0.0;
    if (i < n) {
        const int j = idx_sorted[i];
        const float xx = x[j], yy = y[j], mm = m[j];
        sx[i] = xx; sy[i] = yy; sm[i] = mm;
        if (keys) keys_sorted[i] = keys[j];   // partial sort: the full keys in final order
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

// len = len_arg, or *len_dev + len_add when the length only exists on the device (partitioned step): blocks stride
// over the tiles the actual length needs, so a grid sized for the capacity costs nothing.
template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const T* __restrict__ in, int len_arg, const int* __restrict__ len_dev,
                                                                      int len_add, size_t stride, T* __restrict__ tile_sums, int ntiles) {
    using BR = cub::BlockReduce<T, kScanThreads>;
    __shared__ typename BR::TempStorage tmp;
    const int len = len_dev ? *len_dev + len_add : len_arg;
    const int nt = min(ntiles, (len + kScanTile - 1) / kScanTile);
    const T* a = in + blockIdx.y * stride;
    for (int tile = blockIdx.x; tile < nt; tile += gridDim.x) {
        const int base = tile * kScanTile + threadIdx.x * kScanItems;
        T v = T(0);
#pragma unroll
        for (int k = 0; k < kScanItems; k++) v += (base + k < len) ? a[base + k] : T(0);
        const T t = BR(tmp).Sum(v);
        if (threadIdx.x == 0) tile_sums[blockIdx.y * ntiles + tile] = t;
        __syncthreads();
    }
}
template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_tile_offsets_kernel(T* tile_sums, int ntiles, int len_arg, const int* __restrict__ len_dev,
                                                                         int len_add) {
    // one block per array: chunks of kScanThreads tiles, fixed block-scan tree + a sequential carry
    using BS = cub::BlockScan<T, kScanThreads>;
    __shared__ typename BS::TempStorage tmp;
    __shared__ T carry;
    const int len = len_dev ? *len_dev + len_add : len_arg;
    const int nt = min(ntiles, (len + kScanTile - 1) / kScanTile);
    T* t = tile_sums + blockIdx.x * ntiles;
    if (threadIdx.x == 0) carry = T(0);
    __syncthreads();
    for (int base = 0; base < nt; base += kScanThreads) {
        const int i = base + threadIdx.x;
        const T v = i < nt ? t[i] : T(0);
        T ex, total;
        BS(tmp).ExclusiveSum(v, ex, total);
        const T c = carry;
        if (i < nt) t[i] = c + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
}
template <typename T>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const T* __restrict__ in, T* __restrict__ out, int len_arg,
                                                                  const int* __restrict__ len_dev, int len_add, size_t stride,
                                                                  const T* __restrict__ tile_offs, int ntiles) {
    using BS = cub::BlockScan<T, kScanThreads>;
    __shared__ typename BS::TempStorage tmp;
    const int len = len_dev ? *len_dev + len_add : len_arg;
    const int nt = min(ntiles, (len + kScanTile - 1) / kScanTile);
    const T* a = in + blockIdx.y * stride;
    T* o = out + blockIdx.y * stride;
    for (int tile = blockIdx.x; tile < nt; tile += gridDim.x) {
        const int base = tile * kScanTile + threadIdx.x * kScanItems;
        T v[kScanItems];
#pragma unroll
        for (int k = 0; k < kScanItems; k++) v[k] = (base + k < len) ? a[base + k] : T(0);
        BS(tmp).ExclusiveSum(v, v);
        const T off = tile_offs[blockIdx.y * ntiles + tile];
#pragma unroll
        for (int k = 0; k < kScanItems; k++) if (base + k < len) o[base + k] = off + v[k];
        __syncthreads();
    }
}

// exclusive scan of `arrays` arrays of `len` items each (plane stride `stride`), fixed association => deterministic
template <typename T>
static void launch_scan(cudaStream_t s, const T* in, T* out, T* tile_sums, int arrays, int len_cap, const int* len_dev, int len_add,
                        size_t stride, int max_blocks) {
    const int ntiles = (len_cap + kScanTile - 1) / kScanTile;
    const int gx = std::max(1, std::min(ntiles, max_blocks));
    scan_tile_sums_kernel<T><<<dim3(gx, arrays), kScanThreads, 0, s>>>(in, len_cap, len_dev, len_add, stride, tile_sums, ntiles);
    scan_tile_offsets_kernel<T><<<arrays, kScanThreads, 0, s>>>(tile_sums, ntiles, len_cap, len_dev, len_add);
    scan_apply_kernel<T><<<dim3(gx, arrays), kScanThreads, 0, s>>>(in, out, len_cap, len_dev, len_add, stride, tile_sums, ntiles);
}

// ---- build: single pass over the sorted keys, no level synchronisation ---------------------------------
//
// delta(i) = number of quadtree levels on which sorted bodies i and i+1 share their cell (common key
// prefix / 2 bits).  A cell with >= 2 bodies is an interior node unless all its bodies merge (reference rule
// rs-src/nbody.rs:249-260: closer than EPS in both axes).  An interior node at level l whose first body is i
// exists exactly for l in (dcap(i-1), dcap(i)], where dcap is delta capped (a) at kLevels-1 and (b), for
// pairs inside a chain of mutually close bodies, below the first level at which the chain is alone in
// its cell -- that cell becomes a merged leaf.  A prefix sum over count(i) = max(0, dcap(i)-dcap(i-1)) numbers
// the interior nodes; node k owns child block 1+k (one 64-byte SoA record x[4] y[4] m[4] s[4] plus four child
// block indices), and every node's record lives in its parent's block (the root in slot 0 of block 0).  One
// thread per interior node then emits its block: range end and child boundaries by binary search on the keys,
// mass/COM from the f64 prefix sums, cell bounds by replaying the key bits through the reference's midpoint
// recursion.
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

// delta[i] for the pair (i, i+1), i in [0, n-1); delta[n-1] = sentinel -1.  n = *n_dev when given (partitioned step).
// Also histograms delta into st->delta_hist (block-aggregated): the host sizes the next step's sort from it.
__global__ void bh_delta_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ sx,
                                const float* __restrict__ sy, int n_arg, const int* __restrict__ n_dev, signed char* __restrict__ delta,
                                unsigned char* __restrict__ close, BhStatus* st) {
    __shared__ unsigned hist[kLevels + 1];
    if (threadIdx.x <= kLevels) hist[threadIdx.x] = 0u;
    __syncthreads();
    const int n = n_dev ? *n_dev : n_arg;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i == n - 1) { delta[i] = -1; close[i] = 0; continue; }
        const unsigned long long x = keys[i] ^ keys[i + 1];
        const int d = x ? (__clzll(static_cast<long long>(x)) - (64 - kKeyBits)) >> 1 : kLevels;
        delta[i] = static_cast<signed char>(d);
        close[i] = (fabsf(sx[i] - sx[i + 1]) < kEps && fabsf(sy[i] - sy[i + 1]) < kEps) ? 1 : 0;
        // warp-aggregated: neighbours in Morton order mostly share the same delta
        const unsigned peers = __match_any_sync(__activemask(), d);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[d], static_cast<unsigned>(__popc(peers)));
    }
    __syncthreads();
    if (threadIdx.x <= kLevels && hist[threadIdx.x]) atomicAdd(&st->delta_hist[threadIdx.x], hist[threadIdx.x]);
}

__global__ void bh_cap_kernel(const signed char* __restrict__ delta, const unsigned char* __restrict__ close, int n_arg,
                              const int* __restrict__ n_dev, signed char* __restrict__ dcap, int* __restrict__ count) {
    const int n = n_dev ? *n_dev : n_arg;
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
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = cap_of(i), cl = cap_of(i - 1);
        dcap[i] = static_cast<signed char>(c);
        count[i] = max(0, c - cl);
    }
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
    unsigned part_off;   // added to every child block index (global block index space of partitioned trees)
    int cut_level;       // partitioned build: interior nodes at levels >= cut_level are counted in n_deep
    struct NodeInfo* info;   // EXACT parallel build: per interior node its body range, level and own record slot
    float q_scale;           // FAST: 1/theta^2 -- the 4th record field is qrec(s, q_scale), the folded opening test; EXACT (0): s
    const int* n_dev;        // partitioned step: the number of bodies only exists on the device (overrides n)
    size_t stride;           // plane stride of p3 (n + 1 for the single tree, capacity + 1 for a part)
};
// The FAST walk's opening test, folded into the record by the build: s/d < theta (rs-src/nbody.rs:345)  <=>
// s^2/theta^2 + EPS < d^2 + EPS.  theta so small that 1/theta^2 overflows gives +inf (or NaN for s = 0): never accepted.
__device__ __forceinline__ float qrec(float s, float inv_theta2) { return __fmaf_rn(__fmul_rn(s, s), inv_theta2, kEps); }
__device__ __forceinline__ int build_n(const BuildArgs& a) { return a.n_dev ? *a.n_dev : a.n; }

__device__ __forceinline__ int interior_id(const BuildArgs& a, int first, int level) {
    // id of the interior node that starts at body `first` on `level`
    const int dl = first > 0 ? static_cast<int>(a.dcap[first - 1]) : -1;
    return a.base[first] + (level - dl - 1);
}

// owner[id] = first body of interior node id (one thread per body; bodies start 0.76 nodes on average)
__global__ void bh_owner_kernel(const BuildArgs a, int* __restrict__ owner, BhStatus* st) {
    const int n = build_n(a);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int dhi = a.dcap[i];
        const int dlo = i > 0 ? static_cast<int>(a.dcap[i - 1]) : -1;
        const int b = a.base[i];
        for (int k = 0; k < dhi - dlo && b + k < a.cap_interior; k++) owner[b + k] = i;
        if (a.cut_level > 0) {   // (the grid is a multiple of the warp size and n is uniform: whole warps get here together or not at all)
            const int deep = max(0, dhi - max(dlo, a.cut_level - 1));
            const int wsum = __reduce_add_sync(__activemask(), deep);
            if (wsum > 0 && (threadIdx.x & 31) == (__ffs(__activemask()) - 1)) atomicAdd(&st->n_deep, wsum);
        }
        if (i == 0) {
            const int total = a.base[n - 1];   // count[n-1] == 0
            st->node_count = 1 + 4 * min(total, a.cap_interior);
            st->n_interior = min(total, a.cap_interior);
            if (total > a.cap_interior) st->overflow = 1;
            if (dhi < 0) {
                // no interior node at all: the root is a leaf (one body, or everything merges)
                float cx = 0.f, cy = 0.f, mm = 0.f;
                for (int k = 0; k < n; k++) add_mass_ref(cx, cy, mm, a.sx[k], a.sy[k], a.sm[k]);
                a.nblk[0] = make_float4(cx, 0.f, 0.f, 0.f);
                a.nblk[1] = make_float4(cy, 0.f, 0.f, 0.f);
                a.nblk[2] = make_float4(mm, 0.f, 0.f, 0.f);
                a.nblk[3] = make_float4(-1.f, -1.f, -1.f, -1.f);
                a.ncblk[0] = make_int4(-1, -1, -1, -1);
            }
        }
    }
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0) { st->node_count = 1; st->n_interior = 0; }
}

// one thread per interior node
__global__ void bh_emit_kernel(const BuildArgs a, const int* __restrict__ owner, const BhStatus* st) {
    const int T = st->n_interior;
    const int n = build_n(a);
    const size_t stride = a.stride;
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
            end = n;
        } else {
            const unsigned long long pre = key >> (shift + 2);
            int lo = i + 1, step = 2;
            int hi = min(n, lo + step);
            while (hi < n && (a.keys[hi - 1] >> (shift + 2)) == pre) { lo = hi; step <<= 1; hi = min(n, lo + step); }
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
        if (a.info) {   // range + level of this node; its own record slot is filled in by its parent below
            a.info[id].first = i; a.info[id].end = end;
            if (l == 0) { a.info[id].blk = 0; a.info[id].slot_level = 0; }
        }
        if (l == 0) {   // the root's own record
            const double M = a.p3[n] - a.p3[0];
            a.nblk[0] = make_float4(static_cast<float>((a.p3[stride + n] - a.p3[stride]) / M), 0.f, 0.f, 0.f);
            a.nblk[1] = make_float4(static_cast<float>((a.p3[2 * stride + n] - a.p3[2 * stride]) / M), 0.f, 0.f, 0.f);
            a.nblk[2] = make_float4(static_cast<float>(M), 0.f, 0.f, 0.f);
            const float s0 = __fsub_rn(x2, x1);
            a.nblk[3] = make_float4(a.q_scale != 0.f ? qrec(s0, a.q_scale) : s0, -1.f, -1.f, -1.f);
            a.ncblk[0] = make_int4(static_cast<int>(a.part_off) + blk, -1, -1, -1);
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
                    rec[q] = make_float4(static_cast<float>(MX / M), static_cast<float>(MY / M), static_cast<float>(M),
                                         a.q_scale != 0.f ? qrec(s, a.q_scale) : s);
                    chp[q] = static_cast<int>(a.part_off) + 1 + cid;
                    if (a.info) { a.info[cid].blk = blk; a.info[cid].slot_level = q | ((l + 1) << 8); }
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
// One warp per 32 Morton-consecutive bodies; warps fetch groups from a ticket counter (persistent grid, no tail).
// Stack entries are (node block, mask of lanes that must look at it).  A block is the four children of an opened
// node in SoA form (x[4] y[4] m[4] q[4], 64 bytes): the warp evaluates all four branch-free with packed FP32.
// q is the opening test folded into one number by the build (qrec()): for an interior child q = s^2/theta^2 + EPS, so
// that the reference's per-body test s/d < theta (rs-src/nbody.rs:345) reads q < d^2 + EPS -- the right-hand side is
// the denominator of the force law, which the lane computes anyway; a leaf or an empty slot has q = -1 and always
// passes.  A lane of the mask that fails the test asks for the child to be opened; only children that some lane must
// open are pushed, with that lane mask.  Every body therefore evaluates exactly the reference's interaction list
// (:333-377).  Empty leaves (m = 0, :367) and the body's own leaf (d = 0, :365) need no test: their contribution is
// an exact zero because EPS > 0.
// The lane's membership of the entry's mask is the third input of the comparison (one FSETP yields both "accepts"
// and "must open"), so lanes outside the mask cost no extra instruction and never contribute.
__device__ __forceinline__ uint2 stack_load(unsigned addr) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
    return v;
}
// push (block, mask) if the mask is not empty.  The mask is warp-uniform: every lane stores the same entry to the same
// slot, and a slot only ever receives one value between two warp syncs.
__device__ __forceinline__ void stack_push_if(unsigned& addr, unsigned block, unsigned mask) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.v2.u32 [%0], {%1, %2};\n\t@p add.u32 %0, %0, 8;\n\t}"
        : "+r"(addr) : "r"(block), "r"(mask) : "memory");
}
// One child of the popped block for one lane: c = the lane's (m * 1/(d^2+EPS)) for that child, kept only if the lane is in the
// entry's mask and passes the opening test q < d^2+EPS; returns the ballot of the mask's lanes that fail it (must open).
// One dual-output FSETP gives both predicates, with the mask membership as its third input.
__device__ __forceinline__ unsigned test_child(float q, float e, unsigned in_mask, float& c) {
    unsigned open;
    asm volatile(
        "{\n\t.reg .pred pa, pr, pm;\n\t"
        "setp.ne.u32 pm, %4, 0;\n\t"
        "setp.lt.and.f32 pa|pr, %2, %3, pm;\n\t"
        "selp.f32 %0, %0, 0f00000000, pa;\n\t"
        "vote.sync.ballot.b32 %1, pr, 0xffffffff;\n\t}"
        : "+f"(c), "=r"(open) : "f"(q), "f"(e), "r"(in_mask));
    return open;
}
// EARLYC: when the child-block indices of the popped block are fetched.  0 = only if something is pushed (a second,
// dependent load); 1 = together with the block, into registers (default: 1-2.5 % faster, gpurun r03d; an L1 prefetch
// instead was 3 % slower).  NB_BH_EARLYC=0 selects the former for A/B runs.
template <bool COUNT, bool PARTS, int MINB, int EARLYC>
__global__ void __launch_bounds__(kTravWarps * 32, MINB) bh_traverse_fast_kernel(
    const TreeTable tt, const float* __restrict__ sx,
    const float* __restrict__ sy, const int* __restrict__ idx_sorted, const int* __restrict__ mine, int n_list,
    BhStatus* st, int ticket_slot, const unsigned long long* __restrict__ keys_sorted,
    unsigned* __restrict__ cell_work, const int* __restrict__ n_dev, const uint32_t* __restrict__ epoch_dev) {
    __shared__ uint2 stk[kTravWarps][kStackPerWarp];
    if (cell_work != nullptr && epoch_dev != nullptr) cell_work += (*epoch_dev & 1u) * kNumCells;   // this step's parity
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lanebit = 1u << lane;
    if (n_dev) n_list = *n_dev;     // partitioned step: the size of the part's body list only exists on the device
    const int ngroups = (n_list + 31) >> 5;
    const float2 eps2 = make_float2(kEps, kEps);
    const unsigned sbase = static_cast<unsigned>(__cvta_generic_to_shared(stk[warp]));
    unsigned long long n_int = 0, n_vis = 0, n_pop = 0, n_lane = 0;
    for (;;) {
        int w = 0;
        if (lane == 0) w = atomicAdd(&st->tickets[ticket_slot], 1);
        w = __shfl_sync(0xffffffffu, w, 0);
        if (w >= ngroups) break;
        const int li = w * 32 + lane;
        const bool live = li < n_list;
        const int pos = live ? (mine ? mine[li] : li) : 0;  // position in sorted order
        const float px = sx[pos], py = sy[pos];
        const float2 npx = make_float2(-px, -px), npy = make_float2(-py, -py);
        float2 ax = make_float2(0.f, 0.f), ay = make_float2(0.f, 0.f);
        unsigned sa = sbase;          // shared-memory address of the first free stack slot
        unsigned my_pops = 0;
        {
            const unsigned m0 = __ballot_sync(0xffffffffu, live);
            stack_push_if(sa, tt.root, m0);             // root block = {root, empty, empty, empty}
            __syncwarp();
            if (COUNT) n_vis -= live ? 3 : 0;           // the three padding slots of block 0 are not nodes
        }
        while (sa != sbase) {
            sa -= 8u;
            const uint2 e = stack_load(sa);
            __syncwarp();
            const unsigned in_mask = e.y & lanebit;
            const float4* __restrict__ nb4;
            unsigned part = 0, bi = e.x;
            if (PARTS) {
                part = e.x >> kPartShift; bi = e.x & ((1u << kPartShift) - 1u);
                nb4 = tt.blk[part] + 4 * static_cast<size_t>(bi);
            } else {
                nb4 = tt.blk[0] + 4 * static_cast<size_t>(bi);
            }
            const float4 X = __ldg(nb4 + 0), Y = __ldg(nb4 + 1), M = __ldg(nb4 + 2), Q = __ldg(nb4 + 3);
            const int4* __restrict__ cptr = (PARTS ? tt.cblk[part] : tt.cblk[0]) + bi;
            int4 C = make_int4(0, 0, 0, 0);
            if (EARLYC == 1) C = __ldg(cptr);
            const float2 dx01 = __fadd2_rn(make_float2(X.x, X.y), npx), dx23 = __fadd2_rn(make_float2(X.z, X.w), npx);
            const float2 dy01 = __fadd2_rn(make_float2(Y.x, Y.y), npy), dy23 = __fadd2_rn(make_float2(Y.z, Y.w), npy);
            // d^2 + EPS: the denominator of the force law (rs-src/nbody.rs:180) and the right-hand side of the opening test
            const float2 e01 = __ffma2_rn(dy01, dy01, __ffma2_rn(dx01, dx01, eps2));
            const float2 e23 = __ffma2_rn(dy23, dy23, __ffma2_rn(dx23, dx23, eps2));
            float2 c01 = __fmul2_rn(make_float2(M.x, M.y), make_float2(rcp_approx(e01.x), rcp_approx(e01.y)));
            float2 c23 = __fmul2_rn(make_float2(M.z, M.w), make_float2(rcp_approx(e23.x), rcp_approx(e23.y)));
            const unsigned o0 = test_child(Q.x, e01.x, in_mask, c01.x), o1 = test_child(Q.y, e01.y, in_mask, c01.y);
            const unsigned o2 = test_child(Q.z, e23.x, in_mask, c23.x), o3 = test_child(Q.w, e23.y, in_mask, c23.y);
            ax = __ffma2_rn(c01, dx01, ax); ay = __ffma2_rn(c01, dy01, ay);
            ax = __ffma2_rn(c23, dx23, ax); ay = __ffma2_rn(c23, dy23, ay);
            if (COUNT) {
                const bool act = in_mask != 0u;
                const bool a0 = act && Q.x < e01.x, a1 = act && Q.y < e01.y, a2 = act && Q.z < e23.x, a3 = act && Q.w < e23.y;
                n_vis += act ? 4 : 0;
                n_int += (a0 && M.x != 0.f && !(X.x == px && Y.x == py)) + (a1 && M.y != 0.f && !(X.y == px && Y.y == py)) +
                         (a2 && M.z != 0.f && !(X.z == px && Y.z == py)) + (a3 && M.w != 0.f && !(X.w == px && Y.w == py));
                if (lane == 0) { n_pop++; n_lane += __popc(e.y); atomicAdd(&st->pop_hist[__popc(e.y)], 1ull); }
            }
            if (PARTS) my_pops++;
            if (o0 | o1 | o2 | o3) {
                if (EARLYC != 1) C = __ldg(cptr);
                // stack order bottom -> top: child 3, 2, 1, 0 (non-empty masks only), so that child 0 is opened first
                stack_push_if(sa, static_cast<unsigned>(C.w), o3);
                stack_push_if(sa, static_cast<unsigned>(C.z), o2);
                stack_push_if(sa, static_cast<unsigned>(C.y), o1);
                stack_push_if(sa, static_cast<unsigned>(C.x), o0);
            }
            __syncwarp();
        }
        if (live) {
            const int gi = idx_sorted[pos];
            const int own = gi / tt.shard_len;
            tt.acc[own][gi - own * tt.shard_len] = make_float2(ax.x + ax.y, ay.x + ay.y);
        }
        if (PARTS && cell_work != nullptr && lane == 0)   // partition weights for the next step: walk cost per cut-level cell
            atomicAdd(&cell_work[static_cast<unsigned>(keys_sorted[pos] >> kCellShift)], my_pops);
    }
    if (COUNT) {
        for (int o = 16; o > 0; o >>= 1) {
            n_int += __shfl_xor_sync(0xffffffffu, n_int, o);
            n_vis += __shfl_xor_sync(0xffffffffu, n_vis, o);
        }
        if (lane == 0) {
            atomicAdd(&st->interactions, n_int); atomicAdd(&st->visited, n_vis);
            atomicAdd(&st->pops, n_pop); atomicAdd(&st->pop_lanes, n_lane);
        }
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

// ---- EXACT build, parallel variant (merge-free sets) ---------------------------------------------------------
// If no two bodies are closer than EPS in both axes, Node::insert never merges (rs-src/nbody.rs:249-260), the
// reference tree is the pure refinement tree -- exactly what the sort-based build produces -- and an interior
// node's (px,py,m) is add_mass folded over the bodies of its subtree in ascending particle index
// (rs-src/nbody.rs:234-236,303-320).  So: sort-based topology, then per level one STABLE sort of the body
// indices by the level's key prefix (index order survives inside every cell) and one thread per interior node
// folding its bodies sequentially with the reference's rounding.  Any too-close pair (or a pair that shares all
// 24 key levels) sends the step to the serial build instead, which reproduces merges bit for bit.
__global__ void bh_xkeys_kernel(const float* __restrict__ x, int n, unsigned* __restrict__ xk, int* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    xk[i] = static_cast<unsigned>(f2ord(x[i])) ^ 0x80000000u;   // order-preserving unsigned key
    idx[i] = i;
}

// bodies sorted by x: any later body within EPS in x and in y?  (window capped: a pathological pile-up counts as "close")
__global__ void bh_close_scan_kernel(const int* __restrict__ order, const float* __restrict__ x, const float* __restrict__ y,
                                     int n, int* __restrict__ flag) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int i = order[p];
    const float xi = x[i], yi = y[i];
    for (int q = p + 1; q < n; q++) {
        const int j = order[q];
        if (!(fabsf(__fsub_rn(x[j], xi)) < kEps)) break;
        if (fabsf(__fsub_rn(y[j], yi)) < kEps || q - p > 4096) { *flag = 1; return; }
    }
}

// flag[1] = some pair shares all key levels (the reference would split deeper); flag[2] = deepest interior level
__global__ void bh_maxdelta_kernel(const signed char* __restrict__ delta, const signed char* __restrict__ dcap, int n,
                                   int* __restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    if (delta[i] >= kLevels) flag[1] = 1;
    atomicMax(&flag[2], static_cast<int>(dcap[i]));
}

__global__ void bh_exact_fold_kernel(int level, const NodeInfo* __restrict__ info, const BhStatus* st,
                                     const int* __restrict__ idx_l, const float* __restrict__ x, const float* __restrict__ y,
                                     const float* __restrict__ m, float4* nblk) {
    const int T = st->n_interior;
    float* nf = reinterpret_cast<float*>(nblk);
    for (int id = blockIdx.x * blockDim.x + threadIdx.x; id < T; id += gridDim.x * blockDim.x) {
        const NodeInfo ni = info[id];
        if ((ni.slot_level >> 8) != level) continue;
        float px = 0.f, py = 0.f, mm = 0.f;
        for (int p = ni.first; p < ni.end; p++) {
            const int j = idx_l ? idx_l[p] : p;
            add_mass_ref(px, py, mm, x[j], y[j], m[j]);
        }
        const int slot = ni.slot_level & 3;
        nf[(4 * static_cast<size_t>(ni.blk) + 0) * 4 + slot] = px;
        nf[(4 * static_cast<size_t>(ni.blk) + 1) * 4 + slot] = py;
        nf[(4 * static_cast<size_t>(ni.blk) + 2) * 4 + slot] = mm;
    }
}

// rs-src/nbody.rs:333-377 on the block-SoA tree, nested summation order preserved
struct FrameB { unsigned blk; int slot; int next; unsigned cb; float fx, fy; };

__global__ void bh_traverse_exact_blk_kernel(const float4* __restrict__ nblk, const int4* __restrict__ ncblk,
                                             const float* __restrict__ x, const float* __restrict__ y,
                                             const float* __restrict__ m, int i_begin, int n_local, float theta,
                                             float2* __restrict__ force_out) {
    const int il = blockIdx.x * blockDim.x + threadIdx.x;
    if (il >= n_local) return;
    const float px = x[i_begin + il], py = y[i_begin + il], pm = m[i_begin + il];
    const float* nf = reinterpret_cast<const float*>(nblk);
    const int* cf = reinterpret_cast<const int*>(ncblk);
    FrameB st[64];
    int sp = 0;
    st[0] = FrameB{0u, 0, -1, 0u, 0.f, 0.f};
    float rx = 0.f, ry = 0.f;
    while (sp >= 0) {
        FrameB& f = st[sp];
        if (f.next < 0) {
            const size_t b4 = 4 * static_cast<size_t>(f.blk);
            const float nx = nf[(b4 + 0) * 4 + f.slot], ny = nf[(b4 + 1) * 4 + f.slot];
            const float nm = nf[(b4 + 2) * 4 + f.slot], ns = nf[(b4 + 3) * 4 + f.slot];
            if (ns >= 0.f) {  // interior
                const float dx = __fsub_rn(nx, px), dy = __fsub_rn(ny, py);
                const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));   // :344
                if (__fdiv_rn(ns, d) < theta) {                                             // :345
                    force_ref(px, py, pm, nx, ny, nm, rx, ry);
                    sp--;
                } else {
                    f.fx = 0.f; f.fy = 0.f;
                    f.cb = static_cast<unsigned>(cf[4 * static_cast<size_t>(f.blk) + f.slot]);
                    f.next = 1;
                    st[sp + 1] = FrameB{f.cb, 0, -1, 0u, 0.f, 0.f};
                    sp++;
                    continue;
                }
            } else {  // leaf :363-374
                if ((nx == px && ny == py) || nm == 0.0f) { rx = 0.f; ry = 0.f; }
                else force_ref(px, py, pm, nx, ny, nm, rx, ry);
                sp--;
            }
        } else {
            f.fx = __fadd_rn(f.fx, rx);   // :358-359
            f.fy = __fadd_rn(f.fy, ry);
            if (f.next < 4) {
                st[sp + 1] = FrameB{f.cb, f.next, -1, 0u, 0.f, 0.f};
                f.next++;
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
static void status_begin(Engine& e, struct BhWork& w);
static void status_commit(Engine& e, struct BhWork& w, bool partitioned);
static void bh_forces_partitioned(Engine& e, float theta, int nparts);

// How many domain parts the FAST Barnes-Hut step uses: the world size when sharded over real GPUs (and the
// set is large enough for a level-5 cut to make sense), NB_BH_PARTS virtual parts on one GPU (testing), else 1.
static int bh_partition_count(const Engine& e) {
    if (e.mode != NBX_MODE_FAST) return 1;
    int min_n = 65536;
    if (const char* s = getenv("NB_BH_PARTS_MIN_N")) min_n = atoi(s);
    if (e.n < min_n) return 1;
    if (e.dist && e.world > 1) {
        if (e.bh_partition == 1 || e.bh_arena == nullptr) return 1;
        return e.world;
    }
    int parts = e.bh_partition;
    if (parts <= 0) { if (const char* s = getenv("NB_BH_PARTS")) parts = atoi(s); }
    if (parts > kMaxRanks) parts = kMaxRanks;
    return parts > 1 ? parts : 1;
}

// Resident CTAs per SM the walk is compiled for (its register budget follows: 6 -> 40 registers, 7 -> 36): NB_BH_WALK_MINB
// selects the 5 / 6 / 7 variant for experiments; NB_BH_WALK_BLOCKS only shrinks the persistent grid.
static int walk_minb() {
    static const int v = [] { const char* s = getenv("NB_BH_WALK_MINB"); const int c = s ? atoi(s) : 6; return (c == 5 || c == 7) ? c : 6; }();
    return v;
}
static int traverse_resident_blocks(Engine& e) {
    static int per_sm = 0;
    if (!per_sm) {
        const int mb = walk_minb();
        if (mb == 5) NB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bh_traverse_fast_kernel<false, false, 5, 0>, kTravWarps * 32, 0));
        else if (mb == 7) NB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bh_traverse_fast_kernel<false, false, 7, 0>, kTravWarps * 32, 0));
        else NB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bh_traverse_fast_kernel<false, false, 6, 1>, kTravWarps * 32, 0));
        if (per_sm < 1) per_sm = 1;
        if (const char* v = getenv("NB_BH_WALK_BLOCKS")) { const int c = atoi(v); if (c >= 1 && c < per_sm) per_sm = c; }   // experiments
    }
    return per_sm * e.num_sms;
}

// ticket_slot: which st->tickets[] counter this launch draws its groups from (one per launch between two resets)
static void launch_traverse(Engine& e, const TreeTable& tt, const float* sx, const float* sy, const int* idx_sorted,
                            const int* mine, int n_list, BhStatus* st, int ticket_slot = 0,
                            const unsigned long long* keys_sorted = nullptr, unsigned* cell_work = nullptr, const int* n_dev = nullptr,
                            const uint32_t* epoch_dev = nullptr) {
    // n_dev given: n_list is only an estimate for sizing the (persistent) grid
    const int want = std::max(1, (n_list + kTravWarps * 32 - 1) / (kTravWarps * 32));
    const int blocks = std::min(want, traverse_resident_blocks(e));
    const bool parts = tt.shift != 31;
    const int mb = walk_minb();
    static const int earlyc = [] { const char* v = getenv("NB_BH_EARLYC"); return (v && atoi(v) == 0) ? 0 : 1; }();
#define NB_TRAV(C, P, B, E) bh_traverse_fast_kernel<C, P, B, E><<<blocks, kTravWarps * 32, 0, e.stream>>>(tt, sx, sy, idx_sorted, mine, n_list, st, ticket_slot, keys_sorted, cell_work, n_dev, epoch_dev)
#define NB_TRAV2(C, P) do { if (mb == 5) NB_TRAV(C, P, 5, 0); else if (mb == 7) NB_TRAV(C, P, 7, 0); else if (earlyc == 1) NB_TRAV(C, P, 6, 1); else NB_TRAV(C, P, 6, 0); } while (0)
    if (e.bh_count) { if (parts) NB_TRAV2(true, true); else NB_TRAV2(true, false); }
    else { if (parts) NB_TRAV2(false, true); else NB_TRAV2(false, false); }
#undef NB_TRAV2
#undef NB_TRAV
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
}

// ---- host orchestration ----------------------------------------------------------------------------------------
static void ensure_work(Engine& e, BhWork& w, int n) {
    if (n > w.cap_n) {
        NB_CUDA(cudaStreamSynchronize(e.stream));
        auto fr = [](void* p) { if (p) cudaFree(p); };
        fr(w.keys); fr(w.keys_sorted); fr(w.idx); fr(w.idx_sorted); fr(w.mine); fr(w.sx); fr(w.sy); fr(w.sm);
        fr(w.w3); fr(w.p3); fr(w.tile_sums); fr(w.ndata); fr(w.nbounds); fr(w.nchild); fr(w.nblk); fr(w.ncblk); fr(w.delta); fr(w.dcap); fr(w.close); fr(w.count); fr(w.base); fr(w.owner); fr(w.xk); fr(w.xk_sorted); fr(w.xorder); fr(w.idx_l); fr(w.lk_tmp); fr(w.info); fr(w.cub_tmp);
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
        NB_CUDA(cudaMalloc(&w.xk, N * 4)); NB_CUDA(cudaMalloc(&w.xk_sorted, N * 4));
        NB_CUDA(cudaMalloc(&w.xorder, N * 4)); NB_CUDA(cudaMalloc(&w.idx_l, N * 4)); NB_CUDA(cudaMalloc(&w.lk_tmp, N * 8));
        NB_CUDA(cudaMalloc(&w.info, sizeof(NodeInfo) * (w.cap_nodes / 4)));
        NB_CUDA(cudaMalloc(&w.tile_sums, 3 * ((N + 1 + kScanTile - 1) / kScanTile) * 8));
        size_t b1 = 0, b2 = 0, b3 = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, b1, w.keys, w.keys_sorted, w.idx, w.idx_sorted, n, 0, kKeyBits, e.stream);
        cub::DeviceScan::ExclusiveSum(nullptr, b2, w.count, w.base, n, e.stream);
        cub::DeviceSelect::If(nullptr, b3, Iota(0), w.mine, &w.status->n_mine, n, InRange{w.idx_sorted, 0, n}, e.stream);
        w.cub_bytes = std::max(b1, std::max(b2, b3)) + 256;
        NB_CUDA(cudaMalloc(&w.cub_tmp, w.cub_bytes));
        w.cap_n = n;
        w.alloc_gen++;
    }
}

static void ensure_status(BhWork& w) {
    if (!w.status) {
        NB_CUDA(cudaMalloc(&w.status, sizeof(BhStatus)));
        for (auto& sl : w.slot) {
            NB_CUDA(cudaMallocHost(&sl.host, sizeof(BhStatus)));
            memset(sl.host, 0, sizeof(BhStatus));
            NB_CUDA(cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming));
        }
        w.status_host = w.slot[0].host;
        NB_CUDA(cudaMalloc(&w.xflags, 4 * sizeof(int)));
        NB_CUDA(cudaMallocHost(&w.xflags_host, 4 * sizeof(int)));
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

// keys -> sort -> gather/scan -> single-pass build of ONE tree over all n bodies (nblk / ncblk)
static float inv_theta2(float theta) { return 1.0f / (theta * theta); }
static void build_single_tree(Engine& e, BhWork& w, const GlobalPos& gp, int n, NodeInfo* info, float theta) {
    cudaStream_t s = e.stream;
    const int T = 256, G = (n + T - 1) / T;
    // EXACT (info != nullptr) keeps the full sort: its per-level index sorts below rely on w.idx staying the identity
    const int levels = info ? kLevels : w.sort_levels;
    {
        PhaseScope ps(e, 3);
        bh_keys_kernel<<<G, T, 0, s>>>(gp.x, gp.y, n, w.status, w.keys, w.idx, levels <= 16 ? w.xk : nullptr);
        e.ctr.kernel_launches++;
    }
    {
        PhaseScope ps(e, 4);
        SortBufs sb{w.keys, w.keys_sorted, w.xk, w.xk_sorted, w.idx, w.idx_sorted, w.idx_l};
        w.order = sort_bodies(e, w, sb, n, levels, w.status);
    }
    const size_t stride = static_cast<size_t>(n) + 1;
    {
        PhaseScope ps(e, 6);
        bh_gather_sorted_kernel<<<(n + 1 + T - 1) / T, T, 0, s>>>(gp.x, gp.y, gp.m, w.order, n, w.sx, w.sy, w.sm, w.w3,
                                                                  levels < kLevels ? w.keys : nullptr, w.keys_sorted);
        e.ctr.kernel_launches++;
        launch_scan<double>(s, w.w3, w.p3, w.tile_sums, 3, n + 1, nullptr, 0, stride, 1 << 30);
        e.ctr.kernel_launches += 3;
    }
    {
        PhaseScope ps(e, 5);
        bh_delta_kernel<<<G, T, 0, s>>>(w.keys_sorted, w.sx, w.sy, n, nullptr, w.delta, w.close, w.status);
        bh_cap_kernel<<<G, T, 0, s>>>(w.delta, w.close, n, nullptr, w.dcap, w.count);
        size_t tb = w.cub_bytes;
        cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.count, w.base, n, s);   // integer: deterministic
        BuildArgs ba{w.keys_sorted, w.sx, w.sy, w.sm, w.p3, w.dcap, w.base, w.nblk, w.ncblk, n, (w.cap_nodes - 4) / 4, 0u, 0, info,
                     info ? 0.f : inv_theta2(theta), nullptr, stride};
        bh_owner_kernel<<<G, T, 0, s>>>(ba, w.owner, w.status);
        bh_emit_kernel<<<std::min(G, e.num_sms * 8), T, 0, s>>>(ba, w.owner, w.status);
        e.ctr.kernel_launches += 4;
    }
}

// builds the tree over all e.n bodies and leaves per-local-body acceleration (FAST) / force (EXACT) in w.acc
static void bh_forces(Engine& e, float theta) {
    BhWork& w = work(e);
    ensure_status(w);
    if (!w.capturing) status_begin(e, w);   // claims this step's status slot (its previous user finished two steps ago)
    const int n = e.n;
    ensure_work(e, w, n);
    const int nl = local_count(e), ib = local_begin(e);
    if (nl > w.cap_acc) {
        NB_CUDA(cudaStreamSynchronize(e.stream));
        if (w.acc) NB_CUDA(cudaFree(w.acc));
        NB_CUDA(cudaMalloc(&w.acc, sizeof(float2) * static_cast<size_t>(e.lay.L)));
        w.cap_acc = static_cast<int>(e.lay.L);
        w.alloc_gen++;
    }
    w.acc_src = w.acc;
    w.last_partitioned = false;
    const int nparts = bh_partition_count(e);
    if (nparts > 1) {
        bh_forces_partitioned(e, theta, nparts);
        status_commit(e, w, true);
        return;
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
        if (e.square_aabb) { bh_square_aabb_kernel<<<1, 32, 0, s>>>(w.status); e.ctr.kernel_launches++; }
    }
    if (e.mode == NBX_MODE_EXACT) {
        bool done = false;
        bool try_parallel = n >= 2;
        if (const char* ev = getenv("NB_EXACT_PARALLEL")) try_parallel = try_parallel && atoi(ev) != 0;
        if (try_parallel) {
            PhaseScope ps(e, 5);
            // (a) is the set merge-free?  bodies sorted by x, each looks ahead through its EPS window
            NB_CUDA(cudaMemsetAsync(w.xflags, 0, 4 * sizeof(int), s));
            bh_xkeys_kernel<<<G, T, 0, s>>>(gp.x, n, w.xk, w.idx);
            size_t tb = w.cub_bytes;
            cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.xk, w.xk_sorted, w.idx, w.xorder, n, 0, 32, s);
            bh_close_scan_kernel<<<G, T, 0, s>>>(w.xorder, gp.x, gp.y, n, w.xflags);
            // (b) the refinement tree, with per-node body ranges
            build_single_tree(e, w, gp, n, w.info, theta);
            bh_maxdelta_kernel<<<G, T, 0, s>>>(w.delta, w.dcap, n, w.xflags);
            e.ctr.kernel_launches += 3;
            NB_CUDA(cudaMemcpyAsync(w.xflags_host, w.xflags, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
            NB_CUDA(cudaMemcpyAsync(w.status_host, w.status, sizeof(BhStatus), cudaMemcpyDeviceToHost, s));
            NB_CUDA(cudaStreamSynchronize(s));
            if (!w.xflags_host[0] && !w.xflags_host[1] && !w.status_host->overflow) {
                // (c) exact (px,py,m) of every interior node: fold its bodies in ascending particle index
                const int maxlev = std::min(w.xflags_host[2], kLevels - 1);
                for (int l = 0; l <= maxlev; l++) {
                    const int* idx_l = nullptr;
                    if (l > 0) {
                        tb = w.cub_bytes;
                        cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keys, w.lk_tmp, w.idx, w.idx_l, n, kKeyBits - 2 * l, kKeyBits, s);
                        idx_l = w.idx_l;
                    }
                    bh_exact_fold_kernel<<<e.num_sms * 4, 128, 0, s>>>(l, w.info, w.status, idx_l, gp.x, gp.y, gp.m, w.nblk);
                    e.ctr.kernel_launches++;
                }
                PhaseScope ps2(e, 0);
                if (nl > 0) {
                    bh_traverse_exact_blk_kernel<<<(nl + 127) / 128, 128, 0, s>>>(w.nblk, w.ncblk, gp.x, gp.y, gp.m, ib, nl, theta, w.acc);
                    e.ctr.kernel_launches++;
                }
                TreeTable tt{};
                tt.blk[0] = w.nblk; tt.cblk[0] = w.ncblk; tt.shift = 31; tt.root = 0u;
                w.last_tt = tt; w.last_nparts = 1;
                done = true;
            }
        }
        if (!done) {
            {
                PhaseScope ps(e, 5);
                bh_reset_kernel<<<1, 32, 0, s>>>(w.status);
                bh_aabb_kernel<<<std::min(G, e.num_sms * 4), T, 0, s>>>(gp.x, gp.y, n, w.status);
                if (e.square_aabb) { bh_square_aabb_kernel<<<1, 32, 0, s>>>(w.status); e.ctr.kernel_launches++; }
                bh_build_serial_kernel<<<1, 32, 0, s>>>(gp.x, gp.y, gp.m, n, w.cap_nodes, w.ndata, w.nbounds, w.nchild, w.status);
                bh_finalize_exact_kernel<<<(w.cap_nodes + T - 1) / T, T, 0, s>>>(w.cap_nodes, w.ndata, w.nbounds, w.nchild, w.status);
                e.ctr.kernel_launches += 4;
            }
            PhaseScope ps(e, 0);
            if (nl > 0) {
                bh_traverse_exact_kernel<<<(nl + 127) / 128, 128, 0, s>>>(w.ndata, w.nchild, gp.x, gp.y, gp.m, ib, nl, theta, w.acc);
                e.ctr.kernel_launches++;
            }
            w.last_nparts = 0;
        }
    } else {
        build_single_tree(e, w, gp, n, nullptr, theta);
        const int* mine = nullptr;
        int n_list = n;
        if (e.dist && e.world > 1) {
            size_t tb = w.cub_bytes;
            cub::DeviceSelect::If(w.cub_tmp, tb, Iota(0), w.mine, &w.status->n_mine, n, InRange{w.order, ib, ib + nl}, s);
            mine = w.mine;
            n_list = nl;
        }
        PhaseScope ps(e, 0);
        if (n_list > 0) {
            // replicated tree: this rank walks the bodies of its own index shard, so every write is local
            TreeTable tt{};
            tt.blk[0] = w.nblk; tt.cblk[0] = w.ncblk;
            for (int g = 0; g < kMaxRanks; g++) tt.acc[g] = w.acc;
            tt.acc[e.rank] = w.acc;
            tt.shift = 31; tt.root = 0u; tt.shard_len = static_cast<int>(e.lay.L);
            w.last_tt = tt; w.last_nparts = 1;
            launch_traverse(e, tt, w.sx, w.sy, w.order, mine, n_list, w.status);
        }
    }
    NB_CUDA(cudaGetLastError());
    status_commit(e, w, false);
}

// =================================================================================================
// Domain-partitioned Barnes-Hut (multi-GPU; SURVEY.md section 8e: the ORB split + LET exchange of the brief,
// re-designed for NVLink-connected B200s).  Per-rank work is O(N/G); no host round trip inside a step.
//
// The (global, reference-identical) quadtree is cut at level kCutLevel: its 1,024 cells, in Morton order, are dealt
// to the G ranks as contiguous ranges of equal WEIGHT -- a space-filling-curve bisection, the quadtree-aligned
// equivalent of ORB; weight = last step's measured walk cost per cell + a per-body build cost, so the split
// follows the work, not the body count.  Every subtree below the cut then belongs to exactly one rank ("part").
// One step, on every rank, all on the device:
//   boxes   bounding box of the rank's own index shard -> stored into every peer's arena; min/max over the G boxes
//           (exact: min/max do not round) is the reference's tight global AABB (rs-src/nbody.rs:388-398)
//   keys    Morton keys of the OWN shard only (N/G bodies), CUB radix sort of (key, index) -- N/G items
//   send    the sorted shard falls apart into G contiguous runs, one per destination part (parts are contiguous key
//           ranges): each run is stored straight into the destination's inbox region over NVLink (key + {x,y,m,index},
//           24 B per body -- the all-to-all-v that replaces both body migration and the LET exchange's body half)
//   merge   a destination holds G sorted runs: every body finds its slot by ranking its key in the other G-1 runs
//           (binary searches; ties by source rank => deterministic) -- a G-way merge without any host-known size.
//           From here on the number of bodies of the part exists on the device only.
//   build   the same single-pass build as the single-GPU path, over the part's bodies, in the part's block space;
//           one 48-byte entry per owned cell is published (count, f64 moments, leaf record or child block)
//   top     the <= 341 nodes above the cut are rebuilt by every rank from the G published cell tables -- and, from
//           the same table + the published walk costs, the partition of the NEXT step (identical on all ranks)
//   walk    a rank walks the bodies of its own cells.  No locally-essential tree is materialised or exchanged: block
//           indices carry a part id and a remote subtree is followed IN PLACE in the owning peer's HBM over NVLink --
//           the opening test decides which remote nodes are touched, which is exactly the LET, fetched on demand.
//           Each acceleration is stored to the rank that owns the body by index (8-byte peer stores).
//   integrate  by the index owner, as on one GPU.
// Four device-side ordering points per step (epoch flags, bounded waits): boxes, bodies, trees, walks.
// With nbx_bh_partition(G) on ONE GPU the same code runs with G "virtual ranks" whose "peer" pointers are local --
// that is how the single-GPU test box covers this path bit for bit.
// =================================================================================================
enum { kFlagBoxes = 0, kFlagBodies = 1, kFlagTrees = 2, kFlagWalks = 3 };
constexpr unsigned long long kBodyWeight = 8ull;   // build cost of one body in units of one walk pop (measured ratio)

struct PeerArenas { char* a[kMaxRanks]; };

// The step counter of the partitioned path lives on the device (EpochRef.p), so that a captured CUDA graph of the whole
// step replays unchanged: kernels read it, and select the step-parity buffers (partition plan, walk costs) themselves.
struct EpochRef {
    const uint32_t* p;   // device counter, or nullptr ...
    uint32_t v;          // ... then this value
    __device__ __forceinline__ uint32_t get() const { return p ? *p : v; }
};
__global__ void bhp_epoch_inc_kernel(uint32_t* epoch) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *epoch += 1u;
}

__global__ void bhp_signal_kernel(PeerArenas peers, size_t off_flags, int row, int world, int me, EpochRef er) {
    const uint32_t epoch = er.get();
    const int g = threadIdx.x;
    if (g < world) {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(peers.a[g] + off_flags + (static_cast<size_t>(row) * 64 + me) * sizeof(uint32_t)) = epoch;
        __threadfence_system();
    }
}
__global__ void bhp_wait_kernel(const char* arena, size_t off_flags, int row, int world, EpochRef er, unsigned long long timeout_ns) {
    const uint32_t epoch = er.get();
    const int g = threadIdx.x;
    if (g < world)
        wait_epoch(reinterpret_cast<const uint32_t*>(arena + off_flags) + row * 64 + g, epoch, timeout_ns, g, "Barnes-Hut epoch");
}

// signal + wait in one launch: thread g tells peer g "I reached `row` of this epoch", then waits for peer g's
__global__ void bhp_sync_kernel(PeerArenas peers, const char* arena, size_t off_flags, int row, int world, int me, EpochRef er,
                                unsigned long long timeout_ns) {
    const uint32_t epoch = er.get();
    const int g = threadIdx.x;
    if (g < world) {
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(peers.a[g] + off_flags + (static_cast<size_t>(row) * 64 + me) * sizeof(uint32_t)) = epoch;
        wait_epoch(reinterpret_cast<const uint32_t*>(arena + off_flags) + row * 64 + g, epoch, timeout_ns, g, "Barnes-Hut epoch");
    }
}

// my local box -> slot `me` of every peer's box table, then the "boxes" flag
__global__ void bhp_boxes_publish_kernel(const BhStatus* st, PeerArenas peers, size_t off_aabb, size_t off_flags, int world, int me,
                                         EpochRef er, unsigned* cellwork2) {
    const uint32_t epoch = er.get();
    // this step's walk-cost table (the parity nobody reads any more) starts from zero
    for (int c = threadIdx.x; c < kNumCells; c += blockDim.x) cellwork2[(epoch & 1u) * kNumCells + c] = 0u;
    const int g = threadIdx.x;
    if (g < world) {
        volatile int* dst = reinterpret_cast<volatile int*>(peers.a[g] + off_aabb) + 4 * me;
        for (int k = 0; k < 4; k++) dst[k] = st->aabb_enc[k];
        __threadfence_system();
        *reinterpret_cast<volatile uint32_t*>(peers.a[g] + off_flags + (static_cast<size_t>(kFlagBoxes) * 64 + me) * sizeof(uint32_t)) = epoch;
        __threadfence_system();
    }
}
// wait for all G boxes, reduce (exact), store the global box where the key / build / top kernels read it
__global__ void bhp_boxes_reduce_kernel(const char* arena, size_t off_aabb, size_t off_flags, int world, EpochRef er,
                                        unsigned long long timeout_ns, BhStatus* st_part, BhStatus* st_global, int square) {
    const uint32_t epoch = er.get();
    const int g = threadIdx.x;
    if (g < world)
        wait_epoch(reinterpret_cast<const uint32_t*>(arena + off_flags) + kFlagBoxes * 64 + g, epoch, timeout_ns, g, "Barnes-Hut boxes");
    __syncwarp();
    if (g == 0) {
        const volatile int* b = reinterpret_cast<const volatile int*>(arena + off_aabb);
        int mnx = b[0], mny = b[1], mxx = b[2], mxy = b[3];
        for (int r = 1; r < world; r++) {
            mnx = min(mnx, b[4 * r + 0]); mny = min(mny, b[4 * r + 1]);
            mxx = max(mxx, b[4 * r + 2]); mxy = max(mxy, b[4 * r + 3]);
        }
        int enc[4] = {mnx, mny, mxx, mxy};
        if (square) square_up(enc);
        for (int k = 0; k < 4; k++) { st_part->aabb_enc[k] = enc[k]; st_global->aabb_enc[k] = enc[k]; }
    }
}

// Warp-cooperative 32-ary search: first index in [lo, hi) at which the monotone predicate (false..false true..true) holds,
// hi if none.  Every lane probes one position and a ballot narrows the range 32-fold per round: 4 dependent rounds for a
// 524,288-element run where a per-thread bisection needs 19.  The whole warp must call it; the result is warp-uniform.
template <typename Pred>
__device__ __forceinline__ int warp_first_true(int lo, int hi, const int lane, Pred pred) {
    while (hi - lo > 32) {
        const int step = (hi - lo + 31) >> 5;                 // >= 2
        const int p = min(hi - 1, lo + (lane + 1) * step - 1);
        const unsigned m = __ballot_sync(0xffffffffu, pred(p));
        if (m == 0u) return hi;                               // lane 31 probed hi - 1
        const int f = __ffs(m) - 1;
        hi = min(hi - 1, lo + (f + 1) * step - 1) + 1;        // the probe of lane f is true: the answer is at or before it
        lo = lo + f * step;                                   // the probe of lane f - 1 (lo + f*step - 1) was false
    }
    const int p = lo + lane;
    const unsigned m = __ballot_sync(0xffffffffu, p < hi && pred(p));
    return m ? lo + __ffs(m) - 1 : hi;
}

// ---- send: the key-sorted shard -> the inbox regions of the parts that own its cells --------------------------------
struct SendArgs {
    const unsigned long long* keys;   // full keys of the shard (unsorted); the key of sorted position i is keys[order[i]]
    const int* order;
    const float *x, *y, *m;
    int n_local, gbegin;
    const PartPlan* plan2;      // [2]; this step's is plan2[epoch & 1]
    EpochRef er;
    int nparts, me;
    PeerArenas peers;
    size_t off_in_key, off_in_rec, off_count_in, R;
};
__global__ void __launch_bounds__(256) bhp_send_kernel(const SendArgs a) {
    __shared__ int lo[kMaxRanks + 1];
    {   // first sorted body whose cut-level cell is >= cut[p]: the start of part p's run.  One warp per boundary (every CTA
        // repeats this prologue, so it is kept to a few dependent rounds: it was half of this kernel's time as a bisection)
        const PartPlan* plan = a.plan2 + (a.er.get() & 1u);
        const int lane = threadIdx.x & 31;
        for (int p = threadIdx.x >> 5; p <= a.nparts; p += blockDim.x >> 5) {
            const unsigned long long kmin = static_cast<unsigned long long>(plan->cut[p]) << kCellShift;
            const int l = warp_first_true(0, a.n_local, lane, [&](int i) { return a.keys[a.order[i]] >= kmin; });
            if (lane == 0) lo[p] = l;
        }
    }
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < a.nparts)
        reinterpret_cast<int*>(a.peers.a[threadIdx.x] + a.off_count_in)[a.me] = lo[threadIdx.x + 1] - lo[threadIdx.x];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < a.n_local; i += gridDim.x * blockDim.x) {
        int p = 0;
        while (p + 1 < a.nparts && i >= lo[p + 1]) p++;
        const int j = a.order[i];
        const size_t slot = static_cast<size_t>(a.me) * a.R + static_cast<size_t>(i - lo[p]);
        reinterpret_cast<unsigned long long*>(a.peers.a[p] + a.off_in_key)[slot] = a.keys[j];
        reinterpret_cast<float4*>(a.peers.a[p] + a.off_in_rec)[slot] = make_float4(a.x[j], a.y[j], a.m[j], __int_as_float(a.gbegin + j));
    }
}

// ---- merge: G sorted runs -> one sorted sequence, by ranking ---------------------------------------------------------
struct MergeArgs {
    const char* arena;
    size_t off_in_key, off_in_rec, off_count_in, R;
    int nparts;
    unsigned long long* mkeys;
    float *sx, *sy, *sm;
    int* gidx;
    double* w3;
    size_t stride;
    BhStatus* st;
};
// Tile of 256 consecutive inbox elements (run-major order).  The keys of a tile span [kmin, kmax]; in every run r the
// ranks of all the tile's keys lie inside the window [#keys < kmin, #keys <= kmax], found once per tile by 2 binary
// searches per run -- so the per-element searches run over a few hundred keys that sit in L1, not over the whole run.
__global__ void __launch_bounds__(256) bhp_merge_kernel(const MergeArgs a) {
    __shared__ int cnt[kMaxRanks], off[kMaxRanks + 1], wlo[kMaxRanks], whi[kMaxRanks];
    __shared__ unsigned long long rmin[kMaxRanks], rmax[kMaxRanks];
    const unsigned long long* keys = reinterpret_cast<const unsigned long long*>(a.arena + a.off_in_key);
    const float4* recs = reinterpret_cast<const float4*>(a.arena + a.off_in_rec);
    if (threadIdx.x == 0) {
        const int* ci = reinterpret_cast<const int*>(a.arena + a.off_count_in);
        int o = 0;
        for (int s = 0; s < a.nparts; s++) { cnt[s] = ci[s]; off[s] = o; o += cnt[s]; }
        off[a.nparts] = o;
        if (blockIdx.x == 0) {
            a.st->n_part = o;
            a.w3[o] = 0.0; a.w3[a.stride + o] = 0.0; a.w3[2 * a.stride + o] = 0.0;   // sentinel of the scans over n + 1 entries
        }
    }
    __syncthreads();
    const int n = off[a.nparts];
    const int ntiles = (n + 255) >> 8;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int t0 = tile << 8, t1 = min(t0 + 255, n - 1);
        if (threadIdx.x < a.nparts) {   // key range of the tile: runs are sorted, so the ends of every run segment inside it
            const int r = threadIdx.x;
            const int lo = max(t0, off[r]), hi = min(t1, off[r + 1] - 1);
            rmin[r] = ~0ull; rmax[r] = 0ull;
            if (lo <= hi) {
                rmin[r] = keys[static_cast<size_t>(r) * a.R + (lo - off[r])];
                rmax[r] = keys[static_cast<size_t>(r) * a.R + (hi - off[r])];
            }
        }
        __syncthreads();
        {   // one warp per run (kMaxRanks == warps per CTA), 32-ary searches: this phase was 40 % of the kernel as 8 bisecting threads
            static_assert(kMaxRanks <= 8, "one warp of the 256-thread CTA per run");
            const int r = threadIdx.x >> 5, lane = threadIdx.x & 31;
            if (r < a.nparts) {
                unsigned long long kmin = ~0ull, kmax = 0ull;
                for (int q = 0; q < a.nparts; q++) { kmin = min(kmin, rmin[q]); kmax = max(kmax, rmax[q]); }
                const unsigned long long* kr = keys + static_cast<size_t>(r) * a.R;
                const int l = warp_first_true(0, cnt[r], lane, [&](int i) { return kr[i] >= kmin; });
                const int h = warp_first_true(l, cnt[r], lane, [&](int i) { return kr[i] > kmax; });   // at or after the lower end
                if (lane == 0) { wlo[r] = l; whi[r] = h; }
            }
        }
        __syncthreads();
        const int t = t0 + threadIdx.x;
        if (t < n) {
            int s = 0;
            while (s + 1 < a.nparts && t >= off[s + 1]) s++;
            const int i = t - off[s];
            const unsigned long long k = keys[static_cast<size_t>(s) * a.R + i];
            int pos = i;
            for (int r = 0; r < a.nparts; r++) {
                if (r == s || cnt[r] == 0) continue;
                // rank of k in run r: elements < k, or <= k for runs of lower source rank (ties: lower source first)
                const unsigned long long* kr = keys + static_cast<size_t>(r) * a.R;
                int l = wlo[r], h = whi[r];
                if (r < s) { while (l < h) { const int mid = (l + h) >> 1; if (kr[mid] <= k) l = mid + 1; else h = mid; } }
                else       { while (l < h) { const int mid = (l + h) >> 1; if (kr[mid] < k) l = mid + 1; else h = mid; } }
                pos += l;
            }
            const float4 rc = recs[static_cast<size_t>(s) * a.R + i];
            a.mkeys[pos] = k;
            a.sx[pos] = rc.x; a.sy[pos] = rc.y; a.sm[pos] = rc.z; a.gidx[pos] = __float_as_int(rc.w);
            a.w3[pos] = rc.z;
            a.w3[a.stride + pos] = static_cast<double>(rc.z) * rc.x;
            a.w3[2 * a.stride + pos] = static_cast<double>(rc.z) * rc.y;
        }
        __syncthreads();
    }
}

// one thread per cut-level cell: the entries of the cells this part owns are stored into EVERY rank's copy of the cell
// table (posted NVLink writes), and so is last step's walk cost of the cells this part walked then -- so that the top
// build of every rank reads local memory only.
struct PubArgs {
    PeerArenas peers;
    size_t off_celltab, off_workpub;
    int nparts;
    const PartPlan* plan2;       // [2]: this step's partition is plan2[epoch & 1]; the other one still holds last step's
    EpochRef er;
    int part;
    const unsigned* cellwork2;   // [2][cells]: this rank's measured walk cost per cell; last step's is parity (epoch + 1) & 1
};
// one WARP per cut-level cell (its body range by two 32-ary searches; lane 0 then writes the entry)
__global__ void __launch_bounds__(256) bh_celltab_kernel(const BuildArgs a, const PubArgs pub) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (c >= kNumCells) return;
    const uint32_t epoch = pub.er.get();
    const PartPlan* plan = pub.plan2 + (epoch & 1u);
    const PartPlan* plan_prev = pub.plan2 + ((epoch + 1u) & 1u);
    if (lane == 0 && c >= plan_prev->cut[pub.part] && c < plan_prev->cut[pub.part + 1]) {
        const unsigned wk = pub.cellwork2[((epoch + 1u) & 1u) * kNumCells + c];
        for (int g = 0; g < pub.nparts; g++) reinterpret_cast<unsigned*>(pub.peers.a[g] + pub.off_workpub)[c] = wk;
    }
    if (c < plan->cut[pub.part] || c >= plan->cut[pub.part + 1]) return;
    const int n = build_n(a);
    const size_t stride = a.stride;
    // first sorted body whose cell index is >= c, and >= c + 1
    const int first = warp_first_true(0, n, lane, [&](int i) { return (a.keys[i] >> kCellShift) >= static_cast<unsigned long long>(c); });
    const int end = warp_first_true(first, n, lane, [&](int i) { return (a.keys[i] >> kCellShift) > static_cast<unsigned long long>(c); });
    if (lane != 0) return;
    const int cnt = end - first;
    CellEntry en;
    en.M = en.MX = en.MY = 0.0; en.count = cnt; en.child = -1; en.x = en.y = en.m = 0.f; en.pad = 0;
    if (cnt >= 1) {
        en.M = a.p3[end] - a.p3[first];
        en.MX = a.p3[stride + end] - a.p3[stride + first];
        en.MY = a.p3[2 * stride + end] - a.p3[2 * stride + first];
    }
    if (cnt == 1) {
        en.x = a.sx[first]; en.y = a.sy[first]; en.m = a.sm[first];
    } else if (cnt >= 2) {
        int cid = -1;
        if (static_cast<int>(a.dcap[first]) >= kCutLevel) {
            cid = interior_id(a, first, kCutLevel);
            if (cid >= a.cap_interior) cid = -1;
        }
        if (cid >= 0) {
            en.child = static_cast<int>(a.part_off) + 1 + cid;
        } else {   // the whole cell merges (rs-src/nbody.rs:249-260)
            float bx = 0.f, by = 0.f, bm = 0.f;
            for (int k = 0; k < cnt; k++) add_mass_ref(bx, by, bm, a.sx[first + k], a.sy[first + k], a.sm[first + k]);
            en.x = bx; en.y = by; en.m = bm;
        }
    }
    for (int g = 0; g < pub.nparts; g++) reinterpret_cast<CellEntry*>(pub.peers.a[g] + pub.off_celltab)[c] = en;
}

struct TopArgs {
    const CellEntry* tab;              // the complete cell table (local copy, filled by all ranks)
    const unsigned* work;              // work[c] = walk cost measured for cell c on the previous step (local copy)
    PartPlan* plan2;                   // [2]: plan2[(epoch + 1) & 1] receives the NEXT step's partition, computed here
    EpochRef er;
    int nparts;
    int* tcount; double* tm3; float4* tleaf; int* tchild;
    float4* blk; int4* cblk;
    unsigned top_off;                  // part id of the top tree << kPartShift
    float q_scale;                     // 1/theta^2 (see qrec())
};

__device__ __forceinline__ int top_off_level(int l) { return ((1 << (2 * l)) - 1) / 3; }

// cell width s = x2 - x1 of cell `path` at `level` (rs-src/nbody.rs:286-301 replayed from the root box)
__device__ __forceinline__ float cell_width(const BhStatus* st, int level, int path) {
    float x1 = ord2f(st->aabb_enc[0]), x2 = ord2f(st->aabb_enc[2]);
    for (int t = 0; t < level; t++) {
        const int q = (path >> (2 * (level - 1 - t))) & 3;
        const float cx = __fmul_rn(__fadd_rn(x1, x2), 0.5f);
        if (q & 1) x1 = cx; else x2 = cx;
    }
    return __fsub_rn(x2, x1);
}

// one block: rebuild the nodes above the cut from the published cell tables (identical on every rank), and derive
// the next step's partition: contiguous cell ranges of equal weight (walk cost + kBodyWeight per body)
__global__ void __launch_bounds__(256) bh_top_build_kernel(const TopArgs a, BhStatus* st) {
    const int tid = threadIdx.x;
    const int offk = top_off_level(kCutLevel);
    __shared__ unsigned long long wts[kNumCells];
    __shared__ int s_top;
    for (int c = tid; c < kNumCells; c += blockDim.x) {
        const CellEntry en = a.tab[c];
        a.tcount[offk + c] = en.count;
        a.tm3[offk + c] = en.M; a.tm3[kTopNodes + offk + c] = en.MX; a.tm3[2 * kTopNodes + offk + c] = en.MY;
        a.tleaf[offk + c] = make_float4(en.x, en.y, en.m, 0.f);
        a.tchild[offk + c] = en.child;
        wts[c] = kBodyWeight * static_cast<unsigned long long>(en.count) + a.work[c];
    }
    __shared__ unsigned long long warp_tot[8];
    __shared__ int s_cut[kMaxRanks + 1];
    if (tid == 0) s_top = 0;
    if (tid <= kMaxRanks) s_cut[tid] = kNumCells;
    __syncthreads();
    {   // the next partition: block-wide exclusive prefix of the cell weights (thread t owns cells 4t..4t+3), then every
        // cell whose prefix reaches g/G of the total bids for cut g; the lowest bidder is the cut
        static_assert(kNumCells == 4 * 256, "one thread per four cells");
        const unsigned long long w0 = wts[4 * tid], w1 = wts[4 * tid + 1], w2 = wts[4 * tid + 2], w3 = wts[4 * tid + 3];
        unsigned long long inc = w0 + w1 + w2 + w3;
        const int ln = tid & 31, wp = tid >> 5;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long v = __shfl_up_sync(0xffffffffu, inc, o);
            if (ln >= o) inc += v;
        }
        if (ln == 31) warp_tot[wp] = inc;
        __syncthreads();
        unsigned long long base = 0, total = 0;
        for (int k = 0; k < 8; k++) { if (k < wp) base += warp_tot[k]; total += warp_tot[k]; }
        const unsigned long long e0 = base + inc - (w0 + w1 + w2 + w3), e1 = e0 + w0, e2 = e1 + w1, e3 = e2 + w2;
        const unsigned long long G = static_cast<unsigned long long>(a.nparts);
        for (int g = 1; g < a.nparts; g++) {
            const unsigned long long target = static_cast<unsigned long long>(g) * total;
            int c = kNumCells;
            if (e3 * G >= target) c = 4 * tid + 3;
            if (e2 * G >= target) c = 4 * tid + 2;
            if (e1 * G >= target) c = 4 * tid + 1;
            if (e0 * G >= target) c = 4 * tid;
            if (c < kNumCells) atomicMin(&s_cut[g], c);
        }
        __syncthreads();
        PartPlan* plan_next = a.plan2 + ((a.er.get() + 1u) & 1u);
        if (tid <= a.nparts) plan_next->cut[tid] = tid == 0 ? 0 : (tid == a.nparts ? kNumCells : s_cut[tid]);
    }
    for (int l = kCutLevel - 1; l >= 0; l--) {
        const int off = top_off_level(l), offc = top_off_level(l + 1);
        for (int p = tid; p < (1 << (2 * l)); p += blockDim.x) {
            int cnt = 0; double M = 0.0, MX = 0.0, MY = 0.0; float4 leaf = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int q = 0; q < 4; q++) {
                const int ch = offc + 4 * p + q;
                const int cc = a.tcount[ch];
                cnt += cc; M += a.tm3[ch]; MX += a.tm3[kTopNodes + ch]; MY += a.tm3[2 * kTopNodes + ch];
                if (cc == 1) leaf = a.tleaf[ch];
            }
            a.tcount[off + p] = cnt;
            a.tm3[off + p] = M; a.tm3[kTopNodes + off + p] = MX; a.tm3[2 * kTopNodes + off + p] = MY;
            a.tleaf[off + p] = leaf;
            a.tchild[off + p] = -1;
            if (cnt >= 2) atomicAdd(&s_top, 1);
        }
        __syncthreads();
    }
    auto record = [&](int level, int path, float4& rec, int& child) {
        const int node = top_off_level(level) + path;
        const int cnt = a.tcount[node];
        rec = make_float4(0.f, 0.f, 0.f, -1.0f);
        child = -1;
        if (cnt == 0) return;
        const float4 lf = a.tleaf[node];
        if (cnt == 1) { rec = make_float4(lf.x, lf.y, lf.z, -1.0f); return; }
        int ch = -1;
        if (level == kCutLevel) ch = a.tchild[node];
        else ch = static_cast<int>(a.top_off) + 1 + node;   // top block of an interior top node
        if (ch < 0) { rec = make_float4(lf.x, lf.y, lf.z, -1.0f); return; }   // merged cell at the cut level
        const double M = a.tm3[node];
        const float cw = cell_width(st, level, path);
        rec = make_float4(static_cast<float>(a.tm3[kTopNodes + node] / M), static_cast<float>(a.tm3[2 * kTopNodes + node] / M),
                          static_cast<float>(M), qrec(cw, a.q_scale));
        child = ch;
    };
    // block 0: {root, empty, empty, empty}; block 1 + node: the four children of interior top node `node`
    if (tid == 0) {
        float4 r; int ch;
        record(0, 0, r, ch);
        a.blk[0] = make_float4(r.x, 0.f, 0.f, 0.f); a.blk[1] = make_float4(r.y, 0.f, 0.f, 0.f);
        a.blk[2] = make_float4(r.z, 0.f, 0.f, 0.f); a.blk[3] = make_float4(r.w, -1.f, -1.f, -1.f);
        a.cblk[0] = make_int4(ch, -1, -1, -1);
        st->n_top = s_top;
    }
    for (int l = 0; l < kCutLevel; l++) {
        for (int p = tid; p < (1 << (2 * l)); p += blockDim.x) {
            const int node = top_off_level(l) + p;
            float4 r[4]; int ch[4];
            for (int q = 0; q < 4; q++) record(l + 1, 4 * p + q, r[q], ch[q]);
            const int b = 1 + node;
            a.blk[4 * b + 0] = make_float4(r[0].x, r[1].x, r[2].x, r[3].x);
            a.blk[4 * b + 1] = make_float4(r[0].y, r[1].y, r[2].y, r[3].y);
            a.blk[4 * b + 2] = make_float4(r[0].z, r[1].z, r[2].z, r[3].z);
            a.blk[4 * b + 3] = make_float4(r[0].w, r[1].w, r[2].w, r[3].w);
            a.cblk[b] = make_int4(ch[0], ch[1], ch[2], ch[3]);
        }
    }
}

struct PartStatusPtrs { BhStatus* p[kMaxRanks]; };
__global__ void bhp_fold_status_kernel(BhStatus* global, PartStatusPtrs parts, int nlocal) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        int deep = 0, ovf = 0, run = 0, np = 0;
        for (int r = 0; r < nlocal; r++) {
            np += parts.p[r]->n_part;
            deep += parts.p[r]->n_deep; ovf |= parts.p[r]->overflow; run = max(run, parts.p[r]->sort_long_run);
            for (int d = 0; d <= kLevels; d++) global->delta_hist[d] += parts.p[r]->delta_hist[d];
        }
        global->n_deep += deep;
        global->n_part = np;
        global->overflow |= ovf;
        global->sort_long_run = max(global->sort_long_run, run);
    }
}
__global__ void bhp_plan_init_kernel(PartPlan* plan2, int nparts) {
    if (threadIdx.x == 0 && blockIdx.x == 0)
        for (int b = 0; b < 2; b++)
            for (int g = 0; g <= nparts; g++) plan2[b].cut[g] = static_cast<int>(static_cast<long long>(g) * kNumCells / nparts);
}

// ---- partitioned step: host side -------------------------------------------------------------------------
static void part_free(PartBufs& P) {
    auto fr = [](void* p) { if (p) cudaFree(p); };
    fr(P.keys); fr(P.keys_sorted); fr(P.idx); fr(P.idx_sorted); fr(P.k32); fr(P.k32_sorted); fr(P.idx_fixed);
    fr(P.mkeys); fr(P.sx); fr(P.sy); fr(P.sm); fr(P.gidx); fr(P.w3); fr(P.p3); fr(P.tile_sums);
    fr(P.delta); fr(P.dcap); fr(P.close); fr(P.count); fr(P.base); fr(P.owner); fr(P.itile); fr(P.status);
    if (P.arena_owned) fr(P.arena);
    P = PartBufs();
}

// (re)allocate one part: src_cap bodies on the source side, cap bodies on the destination side
static void part_ensure(Engine& e, BhWork& w, PartBufs& P, int src_cap, size_t cap, const BhArenaLayout& lay, char* ipc_arena) {
    if (P.ready && src_cap <= P.src_cap && cap <= P.cap && lay.R == P.R && lay.cap_blocks == P.cap_blocks &&
        (ipc_arena == nullptr || ipc_arena == P.arena)) return;
    NB_CUDA(cudaStreamSynchronize(e.stream));
    part_free(P);
    const size_t S = static_cast<size_t>(src_cap), N = cap;
    NB_CUDA(cudaMalloc(&P.keys, S * 8)); NB_CUDA(cudaMalloc(&P.keys_sorted, S * 8));
    NB_CUDA(cudaMalloc(&P.idx, S * 4)); NB_CUDA(cudaMalloc(&P.idx_sorted, S * 4)); NB_CUDA(cudaMalloc(&P.idx_fixed, S * 4));
    NB_CUDA(cudaMalloc(&P.k32, S * 4)); NB_CUDA(cudaMalloc(&P.k32_sorted, S * 4));
    NB_CUDA(cudaMalloc(&P.mkeys, N * 8));
    NB_CUDA(cudaMalloc(&P.sx, N * 4)); NB_CUDA(cudaMalloc(&P.sy, N * 4)); NB_CUDA(cudaMalloc(&P.sm, N * 4)); NB_CUDA(cudaMalloc(&P.gidx, N * 4));
    NB_CUDA(cudaMalloc(&P.w3, 3 * (N + 1) * 8)); NB_CUDA(cudaMalloc(&P.p3, 3 * (N + 1) * 8));
    const size_t ntiles = (N + 1 + kScanTile - 1) / kScanTile;
    NB_CUDA(cudaMalloc(&P.tile_sums, 3 * ntiles * 8)); NB_CUDA(cudaMalloc(&P.itile, ntiles * 4));
    NB_CUDA(cudaMalloc(&P.delta, N)); NB_CUDA(cudaMalloc(&P.dcap, N)); NB_CUDA(cudaMalloc(&P.close, N));
    NB_CUDA(cudaMalloc(&P.count, N * 4)); NB_CUDA(cudaMalloc(&P.base, N * 4));
    NB_CUDA(cudaMalloc(&P.owner, sizeof(int) * static_cast<size_t>(lay.cap_blocks)));
    NB_CUDA(cudaMalloc(&P.status, sizeof(BhStatus)));
    NB_CUDA(cudaMemset(P.status, 0, sizeof(BhStatus)));
    if (ipc_arena) {
        P.arena = ipc_arena;
        P.arena_owned = false;
    } else {
        NB_CUDA(cudaMalloc(&P.arena, lay.bytes));
        NB_CUDA(cudaMemset(P.arena, 0, lay.off_in_key));   // flags, boxes, counts, walk costs, cell table
        P.arena_owned = true;
    }
    P.src_cap = src_cap; P.cap = cap; P.cap_blocks = lay.cap_blocks; P.R = lay.R;
    P.ready = true;
    w.alloc_gen++;
    // the source-side sort needs CUB scratch for src_cap items
    size_t b1 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b1, P.keys, P.keys_sorted, P.idx, P.idx_sorted, src_cap, 0, kKeyBits, e.stream);
    if (b1 + 256 > w.cub_bytes) {
        if (w.cub_tmp) NB_CUDA(cudaFree(w.cub_tmp));
        w.cub_bytes = b1 + 256;
        NB_CUDA(cudaMalloc(&w.cub_tmp, w.cub_bytes));
    }
}

static void top_ensure(TopBufs& t) {
    if (t.blk) return;
    NB_CUDA(cudaMalloc(&t.tcount, sizeof(int) * kTopNodes));
    NB_CUDA(cudaMalloc(&t.tm3, sizeof(double) * 3 * kTopNodes));
    NB_CUDA(cudaMalloc(&t.tleaf, sizeof(float4) * kTopNodes));
    NB_CUDA(cudaMalloc(&t.tchild, sizeof(int) * kTopNodes));
    NB_CUDA(cudaMalloc(&t.blk, sizeof(float4) * 4 * (kTopNodes + 1)));
    NB_CUDA(cudaMalloc(&t.cblk, sizeof(int4) * (kTopNodes + 1)));
    NB_CUDA(cudaMalloc(&t.plan, sizeof(PartPlan) * 2));
    NB_CUDA(cudaMalloc(&t.epoch_dev, sizeof(uint32_t)));
    NB_CUDA(cudaMemset(t.epoch_dev, 0, sizeof(uint32_t)));
}

// geometry of the parts this process runs (real ranks: one; virtual ranks: all nparts)
struct PartGeom {
    bool real;
    int nlocal;
    size_t shard;          // index-shard length: owner of body i is i / shard
    BhArenaLayout lay;
};
static PartGeom part_geometry(const Engine& e, int nparts) {
    PartGeom g;
    g.real = e.dist && e.world > 1;
    g.nlocal = g.real ? 1 : nparts;
    g.shard = e.lay.L;
    if (g.real) {
        g.lay = e.bh_lay;
    } else {
        g.shard = ((static_cast<size_t>(e.n) + nparts - 1) / nparts + kShardAlign - 1) / kShardAlign * kShardAlign;
        g.lay.set(g.shard, nparts, static_cast<size_t>(e.n));   // virtual ranks: a private layout sized for this set
    }
    return g;
}

// everything of the partitioned step that allocates, synchronises or bootstraps: must run OUTSIDE a graph capture
static void bh_prepare_partitioned(Engine& e, int nparts) {
    BhWork& w = work(e);
    cudaStream_t s = e.stream;
    const PartGeom pg = part_geometry(e, nparts);
    top_ensure(w.top);
    if (static_cast<int>(w.parts.size()) != pg.nlocal) {
        NB_CUDA(cudaStreamSynchronize(s));
        for (PartBufs& P : w.parts) part_free(P);
        w.parts.assign(pg.nlocal, PartBufs());
    }
    for (int r = 0; r < pg.nlocal; r++)
        part_ensure(e, w, w.parts[r], static_cast<int>(pg.real ? e.L_cap : pg.shard),
                    pg.real ? static_cast<size_t>(e.max_particles) : static_cast<size_t>(e.n), pg.lay, pg.real ? e.bh_arena : nullptr);
    if (pg.real) dist_require_peers(e);
    if (!w.top.plan_valid || w.top.plan_n != e.n || w.top.plan_parts != nparts) {
        bhp_plan_init_kernel<<<1, 32, 0, s>>>(w.top.plan, nparts);   // bootstrap: equal cell ranges; balanced from the next step on
        w.top.plan_valid = true; w.top.plan_n = e.n; w.top.plan_parts = nparts;
        for (int r = 0; r < pg.nlocal; r++)
            NB_CUDA(cudaMemsetAsync(w.parts[r].arena + pg.lay.off_cellwork, 0, 2 * kNumCells * sizeof(unsigned), s));
        e.ctr.kernel_launches++;
    }
}

static void bh_forces_partitioned(Engine& e, float theta, int nparts) {
    BhWork& w = work(e);
    if (!w.capturing) bh_prepare_partitioned(e, nparts);
    const PartGeom pg = part_geometry(e, nparts);
    const bool real = pg.real;
    const int n = e.n, nlocal = pg.nlocal;
    const size_t shard = pg.shard;
    const BhArenaLayout& lay = pg.lay;
    cudaStream_t s = e.stream;
    const int T = 256;
    PeerArenas peers{};
    for (int g = 0; g < nparts; g++) peers.a[g] = real ? e.bh_peer[g] : w.parts[g].arena;
    // the step counter advances on the device (a captured graph replays this launch); the host keeps a mirror
    bhp_epoch_inc_kernel<<<1, 32, 0, s>>>(w.top.epoch_dev);
    e.ctr.kernel_launches++;
    ++w.bh_epoch;
    const EpochRef er{w.top.epoch_dev, 0u};
    PartPlan* plan2 = w.top.plan;
    auto part_id = [&](int r) { return real ? e.rank : r; };
    auto src_begin = [&](int r) { return real ? local_begin(e) : static_cast<int>(std::min<size_t>(static_cast<size_t>(r) * shard, static_cast<size_t>(n))); };
    auto src_count = [&](int r) {
        if (real) return local_count(e);
        const long long b = static_cast<long long>(r) * static_cast<long long>(shard);
        return static_cast<int>(std::max<long long>(0, std::min<long long>(static_cast<long long>(shard), n - b)));
    };
    auto src_x = [&](int r) { return e.arena.x(e.lay, e.cur) + (real ? 0 : src_begin(r)); };
    auto src_y = [&](int r) { return e.arena.y(e.lay, e.cur) + (real ? 0 : src_begin(r)); };
    auto src_m = [&](int r) { return e.arena.m(e.lay) + (real ? 0 : src_begin(r)); };
    // a cross-rank ordering point: real ranks signal + wait in one launch; virtual ranks (one stream) signal all, then wait
    auto sync_point = [&](int row) {
        if (real) {
            bhp_sync_kernel<<<1, 32, 0, s>>>(peers, w.parts[0].arena, lay.off_flags, row, nparts, e.rank, er, e.peer_timeout_ns);
            e.ctr.kernel_launches++;
        } else {
            for (int r = 0; r < nlocal; r++) bhp_signal_kernel<<<1, 32, 0, s>>>(peers, lay.off_flags, row, nparts, r, er);
            for (int r = 0; r < nlocal; r++) bhp_wait_kernel<<<1, 32, 0, s>>>(w.parts[r].arena, lay.off_flags, row, nparts, er, e.peer_timeout_ns);
            e.ctr.kernel_launches += 2 * nlocal;
        }
    };
    // ---- boxes ------------------------------------------------------------------------------------------------------
    {
        PhaseScope ps(e, 2);
        bh_reset_kernel<<<1, 32, 0, s>>>(w.status);
        e.ctr.kernel_launches++;
        for (int r = 0; r < nlocal; r++) {
            PartBufs& P = w.parts[r];
            const int nl = src_count(r);
            bh_reset_kernel<<<1, 32, 0, s>>>(P.status);
            if (nl > 0) bh_aabb_kernel<<<std::min((nl + T - 1) / T, e.num_sms * 4), T, 0, s>>>(src_x(r), src_y(r), nl, P.status);
            bhp_boxes_publish_kernel<<<1, 32, 0, s>>>(P.status, peers, lay.off_aabb, lay.off_flags, nparts, part_id(r), er,
                                                      reinterpret_cast<unsigned*>(P.arena + lay.off_cellwork));
            e.ctr.kernel_launches += 3;
        }
    }
    {
        PhaseScope ps(e, 7);
        for (int r = 0; r < nlocal; r++) {
            bhp_boxes_reduce_kernel<<<1, 32, 0, s>>>(w.parts[r].arena, lay.off_aabb, lay.off_flags, nparts, er, e.peer_timeout_ns,
                                                     w.parts[r].status, w.status, e.square_aabb);
            e.ctr.kernel_launches++;
        }
    }
    // ---- keys + sort of the own shard ----------------------------------------------------------------------------------
    const int levels = w.sort_levels;
    const int* order[kMaxRanks] = {};
    {
        PhaseScope ps(e, 3);
        for (int r = 0; r < nlocal; r++) {
            const int nl = src_count(r);
            if (nl > 0) {
                bh_keys_kernel<<<(nl + T - 1) / T, T, 0, s>>>(src_x(r), src_y(r), nl, w.parts[r].status, w.parts[r].keys, w.parts[r].idx,
                                                            levels <= 16 ? w.parts[r].k32 : nullptr);
                e.ctr.kernel_launches++;
            }
        }
    }
    {
        PhaseScope ps(e, 4);
        for (int r = 0; r < nlocal; r++) {
            PartBufs& P = w.parts[r];
            const int nl = src_count(r);
            order[r] = P.idx_sorted;
            if (nl > 0) {
                SortBufs sb{P.keys, P.keys_sorted, P.k32, P.k32_sorted, P.idx, P.idx_sorted, P.idx_fixed};
                order[r] = sort_bodies(e, w, sb, nl, levels, P.status);
            }
        }
    }
    // ---- send + merge ---------------------------------------------------------------------------------------------------
    {
        PhaseScope ps(e, 7);
        for (int r = 0; r < nlocal; r++) {
            PartBufs& P = w.parts[r];
            const int nl = src_count(r);
            SendArgs sa{P.keys, order[r], src_x(r), src_y(r), src_m(r), nl, src_begin(r), plan2, er, nparts, part_id(r), peers,
                        lay.off_in_key, lay.off_in_rec, lay.off_count_in, lay.R};
            bhp_send_kernel<<<std::max(1, std::min((nl + T - 1) / T, e.num_sms * 8)), T, 0, s>>>(sa);
            e.ctr.kernel_launches++;
        }
        sync_point(kFlagBodies);
    }
    const int est = static_cast<int>(std::min<size_t>(real ? static_cast<size_t>(e.max_particles) : static_cast<size_t>(n),
                                                      2 * ((static_cast<size_t>(n) + nparts - 1) / nparts) + 4096));   // grid sizing only
    const int GE = std::max(1, std::min((est + T - 1) / T, e.num_sms * 16));
    {
        PhaseScope ps(e, 6);
        for (int r = 0; r < nlocal; r++) {
            PartBufs& P = w.parts[r];
            const size_t stride = P.cap + 1;
            MergeArgs ma{P.arena, lay.off_in_key, lay.off_in_rec, lay.off_count_in, lay.R, nparts, P.mkeys, P.sx, P.sy, P.sm, P.gidx, P.w3, stride, P.status};
            bhp_merge_kernel<<<GE, T, 0, s>>>(ma);
            launch_scan<double>(s, P.w3, P.p3, P.tile_sums, 3, static_cast<int>(P.cap) + 1, &P.status->n_part, 1, stride, e.num_sms * 4);
            e.ctr.kernel_launches += 4;
        }
    }
    // ---- build of the part's subtree forest + its cell table ---------------------------------------------------------------
    {
        PhaseScope ps(e, 5);
        for (int r = 0; r < nlocal; r++) {
            PartBufs& P = w.parts[r];
            const int* nd = &P.status->n_part;
            bh_delta_kernel<<<GE, T, 0, s>>>(P.mkeys, P.sx, P.sy, 0, nd, P.delta, P.close, P.status);
            bh_cap_kernel<<<GE, T, 0, s>>>(P.delta, P.close, 0, nd, P.dcap, P.count);
            launch_scan<int>(s, P.count, P.base, P.itile, 1, static_cast<int>(P.cap), nd, 0, 0, e.num_sms * 4);
            BuildArgs ba{P.mkeys, P.sx, P.sy, P.sm, P.p3, P.dcap, P.base, reinterpret_cast<float4*>(P.arena + lay.off_nblk),
                         reinterpret_cast<int4*>(P.arena + lay.off_ncblk), 0, P.cap_blocks - 2,
                         static_cast<unsigned>(part_id(r)) << kPartShift, kCutLevel, nullptr, inv_theta2(theta), nd, P.cap + 1};
            bh_owner_kernel<<<GE, T, 0, s>>>(ba, P.owner, P.status);
            bh_emit_kernel<<<std::min(GE, e.num_sms * 8), T, 0, s>>>(ba, P.owner, P.status);
            PubArgs pa{peers, lay.off_celltab, lay.off_workpub, nparts, plan2, er, part_id(r),
                       reinterpret_cast<const unsigned*>(P.arena + lay.off_cellwork)};
            bh_celltab_kernel<<<kNumCells / 8, 256, 0, s>>>(ba, pa);   // one warp per cell
            e.ctr.kernel_launches += 8;
        }
    }
    {
        PhaseScope ps(e, 7);
        sync_point(kFlagTrees);
    }
    // ---- top tree (+ the next partition), then the walk ------------------------------------------------------------------
    TreeTable tt{};
    tt.shift = kPartShift;
    tt.root = static_cast<unsigned>(kMaxRanks) << kPartShift;
    tt.shard_len = static_cast<int>(shard);
    {
        PhaseScope ps(e, 6);
        TopArgs ta{};
        for (int g = 0; g < nparts; g++) {
            tt.blk[g] = reinterpret_cast<const float4*>(peers.a[g] + lay.off_nblk);
            tt.cblk[g] = reinterpret_cast<const int4*>(peers.a[g] + lay.off_ncblk);
            tt.acc[g] = real ? reinterpret_cast<float2*>(peers.a[g] + lay.off_acc) : (w.acc + static_cast<size_t>(g) * shard);
        }
        ta.tab = reinterpret_cast<const CellEntry*>(w.parts[0].arena + lay.off_celltab);      // every local part holds a complete copy
        ta.work = reinterpret_cast<const unsigned*>(w.parts[0].arena + lay.off_workpub);
        ta.plan2 = plan2; ta.er = er; ta.nparts = nparts;
        ta.tcount = w.top.tcount; ta.tm3 = w.top.tm3; ta.tleaf = w.top.tleaf; ta.tchild = w.top.tchild;
        ta.blk = w.top.blk; ta.cblk = w.top.cblk;
        ta.top_off = static_cast<unsigned>(kMaxRanks) << kPartShift;
        ta.q_scale = inv_theta2(theta);
        bh_top_build_kernel<<<1, 256, 0, s>>>(ta, w.status);
        e.ctr.kernel_launches++;
        tt.blk[kMaxRanks] = w.top.blk;
        tt.cblk[kMaxRanks] = w.top.cblk;
    }
    {
        PhaseScope ps(e, 0);
        for (int r = 0; r < nlocal; r++) {
            PartBufs& P = w.parts[r];
            launch_traverse(e, tt, P.sx, P.sy, P.gidx, nullptr, est, w.status, real ? 0 : r, P.mkeys,
                            reinterpret_cast<unsigned*>(P.arena + lay.off_cellwork), &P.status->n_part, w.top.epoch_dev);
        }
    }
    {
        PhaseScope ps(e, 7);
        // accelerations were stored to their index owners (over NVLink): publish "my walk is done" and wait until every
        // rank's is (=> my acc buffer is complete, nobody reads my subtrees any more, the inboxes may be overwritten)
        sync_point(kFlagWalks);
    }
    PartStatusPtrs sp{};
    for (int r = 0; r < nlocal; r++) sp.p[r] = w.parts[r].status;
    bhp_fold_status_kernel<<<1, 32, 0, s>>>(w.status, sp, nlocal);
    e.ctr.kernel_launches++;
    w.acc_src = real ? reinterpret_cast<float2*>(e.bh_arena + lay.off_acc) : w.acc;
    w.last_partitioned = true;
    w.last_tt = tt; w.last_nparts = nparts; w.last_nparts_hint = nparts;
    NB_CUDA(cudaGetLastError());
}

// Deferred: the status block of a step is copied to pinned memory behind the step and looked at when its slot comes
// round again (two steps later) or at any synchronising call, so stepping never stalls the host on the GPU.
static int choose_sort_levels(const BhStatus& h, int n) {
    if (n < 8192) return kLevels;
    // Sorting on L levels leaves the neighbours that share >= L levels to the fix-up (cheap for pairs and short runs).
    // tail[L] = number of such neighbour pairs.  Prefer the 32-bit path (L <= 16): take the smallest L that leaves at
    // most n/32 pairs; if that needs more than 16 levels, 16 is still taken as long as no more than n/4 pairs remain.
    unsigned long long tail[kLevels + 2] = {};
    for (int d = kLevels; d >= 0; d--) tail[d] = tail[d + 1] + h.delta_hist[d];
    int L = kLevels;
    while (L > 6 && tail[L - 1] * 32ull <= static_cast<unsigned long long>(n)) L--;
    if (L > 16 && tail[16] * 4ull <= static_cast<unsigned long long>(n)) L = 16;
    // 12 levels = 24 key bits = one radix pass fewer.  Measured on B200 (profiles/r02_sort_depth_sweep.jsonl and the round-2 probes):
    // a pass costs about 10 us + 8 us per million keys, the fix-up about 64 us per million neighbour pairs it has to order.
    if (L > 12 && static_cast<double>(tail[12]) * 64e-6 < 10.0 + 8e-6 * static_cast<double>(n)) L = 12;
    if (getenv("NB_DEBUG_SORT"))
        fprintf(stderr, "nbody_b200: sort depth: n=%d tail[12]=%llu tail[14]=%llu tail[16]=%llu tail[18]=%llu tail[20]=%llu -> %d levels\n", n,
                tail[12], tail[14], tail[16], tail[18], tail[20], L);
    return L;
}
static void process_slot(Engine& e, BhWork& w, int k, bool wait) {
    BhWork::StatusSlot& sl = w.slot[k];
    if (!sl.pending) return;
    if (wait) NB_CUDA(cudaEventSynchronize(sl.ev));
    else if (cudaEventQuery(sl.ev) != cudaSuccess) { (void)cudaGetLastError(); return; }
    sl.pending = false;
    const BhStatus& h = *sl.host;
    if (h.depth_error) fatal("Node::insert() - recursion depth > 50 (the reference panics here, rs-src/nbody.rs:230-232)", __FILE__, __LINE__);
    if (h.overflow && !w.warned) {
        fprintf(stderr, "nbody_b200: warning: quadtree node pool exhausted (%d nodes); deepest cells were merged\n", w.cap_nodes);
        w.warned = true;
    }
    if (h.sort_long_run > 0)
        fprintf(stderr, "nbody_b200: note: the partial sort left a run of %d bodies to its fix-up (the set changed abruptly)\n", h.sort_long_run);
    if (sl.partitioned)   // this rank's share; rank 0 (or the single process) also counts the shared top tree
        e.ctr.bh_nodes_built += 4ull * static_cast<uint64_t>(h.n_deep) + (e.rank == 0 ? 1ull + 4ull * static_cast<uint64_t>(h.n_top) : 0ull);
    else
        e.ctr.bh_nodes_built += static_cast<uint64_t>(h.node_count);
    e.ctr.bh_interactions += h.interactions;
    e.ctr.bh_nodes_visited += h.visited;
    e.ctr.bh_pops += h.pops;
    e.ctr.bh_pop_lanes += h.pop_lanes;
    e.ctr.bh_part_bodies = sl.partitioned ? static_cast<uint64_t>(h.n_part) : static_cast<uint64_t>(sl.n);
    e.ctr.bh_sort_levels = static_cast<uint64_t>(w.sort_levels);
    for (int q = 0; q < 33; q++) w.pop_hist[q] += h.pop_hist[q];
    if (e.mode == NBX_MODE_FAST && sl.n > 0) w.sort_levels = choose_sort_levels(h, sl.n);
}
static void check_status(Engine& e, BhWork& w, bool sync_now) {
    // older slot first, so that sort_levels ends up from the most recent step
    const int older = w.cur_slot ^ 1;
    process_slot(e, w, older, sync_now);
    process_slot(e, w, w.cur_slot, sync_now);
}
// start of a Barnes-Hut call: claim the status slot of this position-buffer parity (wait for its previous user)
static void status_begin(Engine& e, BhWork& w) {
    const int k = e.cur & 1;
    process_slot(e, w, k ^ 1, false);
    process_slot(e, w, k, true);
    w.cur_slot = k;
    w.status_host = w.slot[k].host;
    if (w.seen_set_gen != e.set_gen) { w.seen_set_gen = e.set_gen; w.sort_levels = kLevels; }
    if (const char* f = getenv("NB_SORT_LEVELS")) { const int v = atoi(f); if (v >= 1 && v <= kLevels) w.sort_levels = v; }   // experiments
}
// end of the device work of a Barnes-Hut call: queue the status copy behind it
static void status_commit(Engine& e, BhWork& w, bool partitioned) {
    BhWork::StatusSlot& sl = w.slot[w.cur_slot];
    NB_CUDA(cudaMemcpyAsync(sl.host, w.status, sizeof(BhStatus), cudaMemcpyDeviceToHost, e.stream));
    sl.partitioned = partitioned;
    sl.n = (partitioned && e.dist && e.world > 1) ? e.n / std::max(1, w.last_nparts_hint) : e.n;   // bodies the histogram covers
    if (!w.capturing) {
        NB_CUDA(cudaEventRecord(sl.ev, e.stream));
        sl.pending = true;
    }
}
void bh_pop_histogram(Engine& e, uint64_t* out33, bool reset) {
    for (int k = 0; k < 33; k++) out33[k] = 0;
    if (!e.bh) return;
    BhWork& w = work(e);
    check_status(e, w, true);
    for (int k = 0; k < 33; k++) { out33[k] = w.pop_hist[k]; if (reset) w.pop_hist[k] = 0; }
}
// Diagnostic: `iters` back-to-back ordering points (signal + wait, one launch each) on a private flag row, timed with
// events -- the floor the four per-step ordering points of the partitioned step cannot go below.
float bh_sync_test(Engine& e, int iters) {
    if (!(e.dist && e.world > 1) || !e.bh_arena || iters <= 0) return 0.f;
    dist_require_peers(e);
    BhWork& w = work(e);
    PeerArenas peers{};
    for (int g = 0; g < e.world; g++) peers.a[g] = e.bh_peer[g];
    cudaEvent_t a, b;
    NB_CUDA(cudaEventCreate(&a)); NB_CUDA(cudaEventCreate(&b));
    for (int k = 0; k < 4; k++)
        bhp_sync_kernel<<<1, 32, 0, e.stream>>>(peers, e.bh_arena, e.bh_lay.off_flags, 4, e.world, e.rank, EpochRef{nullptr, ++w.sync_test_epoch}, e.peer_timeout_ns);
    NB_CUDA(cudaEventRecord(a, e.stream));
    for (int k = 0; k < iters; k++)
        bhp_sync_kernel<<<1, 32, 0, e.stream>>>(peers, e.bh_arena, e.bh_lay.off_flags, 4, e.world, e.rank, EpochRef{nullptr, ++w.sync_test_epoch}, e.peer_timeout_ns);
    NB_CUDA(cudaEventRecord(b, e.stream));
    NB_CUDA(cudaStreamSynchronize(e.stream));
    float ms = 0.f;
    NB_CUDA(cudaEventElapsedTime(&ms, a, b));
    cudaEventDestroy(a); cudaEventDestroy(b);
    return ms / iters;
}

void bh_poll(Engine& e) {
    if (e.bh) check_status(e, work(e), true);
}

static void bh_step_body(Engine& e, BhWork& w, float theta, float dt) {
    bh_forces(e, theta);
    PhaseScope ps(e, 1);
    if (e.mode == NBX_MODE_EXACT) launch_integrate_exact(e, w.acc, dt, true);
    else launch_integrate_fast(e, w.acc_src, 1, dt, true);
}

// The single-GPU FAST step is a fixed sequence of ~20 short kernels whose launch configuration depends only on
// n: capture it once per position-buffer parity and replay it (NB_BH_GRAPH=0 disables).
static bool bh_graph_eligible(const Engine& e) {
    if (e.mode != NBX_MODE_FAST || e.phase_timing || e.bh_count) return false;
    // sharded with ONE replicated tree: the position gather carries epoch values as launch parameters -> no replay.
    // The domain-partitioned step (real or virtual ranks) keeps its step counter on the device and replays as a graph.
    if (e.dist && e.world > 1 && bh_partition_count(e) == 1) return false;
    if (const char* s = getenv("NB_BH_GRAPH")) return atoi(s) != 0;
    return true;
}

void bh_step(Engine& e, float theta, float dt) {
    if (e.n == 0) return;
    BhWork& w = work(e);
    e.kdk_theta = theta;
    if (bh_graph_eligible(e)) {
        ensure_status(w);
        status_begin(e, w);
        ensure_work(e, w, e.n);
        if (local_count(e) > w.cap_acc) {   // same sizing rule as bh_forces: allocate outside the capture
            NB_CUDA(cudaStreamSynchronize(e.stream));
            if (w.acc) NB_CUDA(cudaFree(w.acc));
            NB_CUDA(cudaMalloc(&w.acc, sizeof(float2) * static_cast<size_t>(e.lay.L)));
            w.cap_acc = static_cast<int>(e.lay.L);
            w.alloc_gen++;
        }
        const int nparts_now = bh_partition_count(e);
        if (nparts_now > 1) bh_prepare_partitioned(e, nparts_now);   // allocations / bootstrap never happen inside a capture
        BhWork::GraphSlot& g = w.graph[e.cur];
        // every pointer and scalar a captured launch carries is a function of this key (alloc_gen covers the workspace)
        const bool hit = g.exec && g.nparts == nparts_now && g.n == e.n && g.cur == e.cur && g.theta == theta && g.dt == dt && g.stream == e.stream &&
                         g.arena == e.arena.base && g.alloc_gen == w.alloc_gen && g.L == e.lay.L && g.sort_levels == w.sort_levels &&
                         g.kick == kick_dt(e, dt) && g.square == e.square_aabb;
        if (!hit) {
            // Re-capture (host-only work, ~20 nodes) and patch the instantiated graph in place: a theta / dt change
            // from the reference UI (hs-src/RustNBodyExperiment.hs:88-93) or a reallocation only changes kernel
            // parameters, not the topology, so cudaGraphExecUpdate succeeds; a different n changes grid sizes, which
            // it also accepts.  Only if the update is refused is the graph re-instantiated.
            const int cur0 = e.cur;
            const uint64_t l0 = e.ctr.kernel_launches;
            const float kick0 = kick_dt(e, dt);
            cudaGraph_t graph = nullptr;
            w.capturing = true;
            NB_CUDA(cudaStreamBeginCapture(e.stream, cudaStreamCaptureModeThreadLocal));
            bh_step_body(e, w, theta, dt);
            NB_CUDA(cudaStreamEndCapture(e.stream, &graph));
            w.capturing = false;
            bool updated = false;
            // a different n / sort depth / part count / box option changes the number or shape of the launches: no point in
            // asking for an in-place update then (it would be refused); theta, dt and reallocations only change parameters
            const bool same_topology = g.exec && g.n == e.n && g.sort_levels == w.sort_levels && g.nparts == nparts_now && g.square == e.square_aabb;
            if (g.exec && !same_topology) { NB_CUDA(cudaGraphExecDestroy(g.exec)); g.exec = nullptr; }
            if (g.exec) {
                cudaGraphExecUpdateResultInfo info;
                if (cudaGraphExecUpdate(g.exec, graph, &info) == cudaSuccess) updated = true;
                else { (void)cudaGetLastError(); NB_CUDA(cudaGraphExecDestroy(g.exec)); g.exec = nullptr; }
            }
            if (!updated) NB_CUDA(cudaGraphInstantiate(&g.exec, graph, 0));
            NB_CUDA(cudaGraphDestroy(graph));
            g.n = e.n; g.cur = cur0; g.theta = theta; g.dt = dt; g.stream = e.stream; g.arena = e.arena.base;
            g.alloc_gen = w.alloc_gen; g.L = e.lay.L; g.sort_levels = w.sort_levels; g.kick = kick0; g.square = e.square_aabb;
            g.nparts = nparts_now; g.partitioned = w.last_partitioned; g.acc_src = w.acc_src; g.tt = w.last_tt;
            g.launches = e.ctr.kernel_launches - l0;
            // the capture already advanced the host-side state exactly like a replay does below
        } else {
            e.ctr.kernel_launches += g.launches;
            e.cur ^= 1;
            if (e.integrator == NBX_INTEGRATOR_LEAPFROG_KDK) e.kdk_pending = 0.5f * dt;
            w.acc_src = g.acc_src;
            w.last_partitioned = g.partitioned;
            w.last_tt = g.tt; w.last_nparts = g.nparts;
            if (g.partitioned) { ++w.bh_epoch; w.last_nparts_hint = g.nparts; }   // mirrors of what the replayed launches do on the device
        }
        NB_CUDA(cudaGraphLaunch(g.exec, e.stream));
        BhWork::StatusSlot& sl = w.slot[w.cur_slot];   // the graph's last node copied the status block into this slot
        NB_CUDA(cudaEventRecord(sl.ev, e.stream));
        sl.pending = true; sl.partitioned = g.partitioned;
        sl.n = (g.partitioned && e.dist && e.world > 1) ? e.n / std::max(1, g.nparts) : e.n;
    } else {
        bh_step_body(e, w, theta, dt);
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
    else launch_accel_from_partial(e, w.acc_src, 1, out);
    e.step_count++;
    dist_signal_step_done(e);
    check_status(e, w, true);
}

// Debug / parity aid: the FAST tree of the most recent Barnes-Hut call in the oracle's flatten format (DFS
// pre-order, children in index order; 9 floats per node: x1,y1,x2,y2,px,py,m,has_children,depth), so that tests
// can compare it with the reference tree node by node.  Walks downloaded copies of the block arrays on the host.
int bh_flatten(Engine& e, float* out9, int cap) {
    if (!e.bh) return 0;
    BhWork& w = work(e);
    if (w.last_nparts == 0) return 0;
    check_status(e, w, true);
    NB_CUDA(cudaStreamSynchronize(e.stream));
    const BhStatus h = *w.status_host;
    const TreeTable& tt = w.last_tt;
    const int nslots = kMaxRanks + 1;
    std::vector<std::vector<float4>> blk(nslots);
    std::vector<std::vector<int4>> cblk(nslots);
    auto fetch = [&](int part, size_t nblocks) {
        if (!tt.blk[part] || nblocks == 0) return;
        blk[part].resize(4 * nblocks);
        cblk[part].resize(nblocks);
        NB_CUDA(cudaMemcpy(blk[part].data(), tt.blk[part], sizeof(float4) * 4 * nblocks, cudaMemcpyDeviceToHost));
        NB_CUDA(cudaMemcpy(cblk[part].data(), tt.cblk[part], sizeof(int4) * nblocks, cudaMemcpyDeviceToHost));
    };
    if (w.last_nparts == 1) {
        fetch(0, static_cast<size_t>(h.n_interior) + 1);
    } else {
        if (e.dist && e.world > 1) return -1;   // peers' arrays are not fetched here: single-process (virtual ranks) only
        for (int r = 0; r < w.last_nparts; r++) fetch(r, static_cast<size_t>(w.parts[r].cap_blocks));   // virtual ranks: local arrays
        fetch(kMaxRanks, kTopNodes + 1);
    }
    auto ord2f_h = [](int i) { int j = i >= 0 ? i : i ^ 0x7fffffff; float f; memcpy(&f, &j, 4); return f; };
    struct Item { unsigned blockidx; int slot; float x1, y1, x2, y2; int depth; };
    std::vector<Item> stack;
    stack.push_back(Item{tt.root, 0, ord2f_h(h.aabb_enc[0]), ord2f_h(h.aabb_enc[1]), ord2f_h(h.aabb_enc[2]), ord2f_h(h.aabb_enc[3]), 0});
    int count = 0;
    const unsigned mask = tt.shift >= 31 ? 0x7fffffffu : ((1u << tt.shift) - 1u);
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        const unsigned part = tt.shift >= 31 ? 0u : (it.blockidx >> tt.shift), bi = it.blockidx & mask;
        if (part >= static_cast<unsigned>(nslots) || 4 * static_cast<size_t>(bi) + 3 >= blk[part].size()) return -2;
        const float* X = reinterpret_cast<const float*>(&blk[part][4 * bi + 0]);
        const float* Y = reinterpret_cast<const float*>(&blk[part][4 * bi + 1]);
        const float* M = reinterpret_cast<const float*>(&blk[part][4 * bi + 2]);
        const float* S = reinterpret_cast<const float*>(&blk[part][4 * bi + 3]);
        const int* C = reinterpret_cast<const int*>(&cblk[part][bi]);
        const bool interior = S[it.slot] >= 0.0f;
        if (count < cap) {
            float* r = out9 + 9 * static_cast<size_t>(count);
            r[0] = it.x1; r[1] = it.y1; r[2] = it.x2; r[3] = it.y2;
            r[4] = X[it.slot]; r[5] = Y[it.slot]; r[6] = M[it.slot];
            r[7] = interior ? 1.0f : 0.0f; r[8] = static_cast<float>(it.depth);
        }
        count++;
        if (interior) {
            const float cx = (it.x1 + it.x2) * 0.5f, cy = (it.y1 + it.y2) * 0.5f;   // rs-src/nbody.rs:289-300
            const Item ch[4] = {Item{static_cast<unsigned>(C[it.slot]), 0, it.x1, cy, cx, it.y2, it.depth + 1},
                                Item{static_cast<unsigned>(C[it.slot]), 1, cx, cy, it.x2, it.y2, it.depth + 1},
                                Item{static_cast<unsigned>(C[it.slot]), 2, it.x1, it.y1, cx, cy, it.depth + 1},
                                Item{static_cast<unsigned>(C[it.slot]), 3, cx, it.y1, it.x2, cy, it.depth + 1}};
            for (int q = 3; q >= 0; q--) stack.push_back(ch[q]);
        }
    }
    return count;
}

void bh_shutdown(Engine& e) {
    if (!e.bh) return;
    BhWork& w = work(e);
    auto fr = [](void* p) { if (p) cudaFree(p); };
    fr(w.keys); fr(w.keys_sorted); fr(w.idx); fr(w.idx_sorted); fr(w.mine); fr(w.sx); fr(w.sy); fr(w.sm);
    fr(w.w3); fr(w.p3); fr(w.tile_sums); fr(w.ndata); fr(w.nbounds); fr(w.nchild); fr(w.nblk); fr(w.ncblk); fr(w.delta); fr(w.dcap); fr(w.close); fr(w.count); fr(w.base); fr(w.owner); fr(w.xk); fr(w.xk_sorted); fr(w.xorder); fr(w.idx_l); fr(w.lk_tmp); fr(w.info); fr(w.cub_tmp); fr(w.status); fr(w.acc);
    for (auto& sl : w.slot) {
        if (sl.host) cudaFreeHost(sl.host);
        if (sl.ev) cudaEventDestroy(sl.ev);
    }
    for (auto& g : w.graph) if (g.exec) cudaGraphExecDestroy(g.exec);
    for (PartBufs& P : w.parts) part_free(P);
    fr(w.top.tcount); fr(w.top.tm3); fr(w.top.tleaf); fr(w.top.tchild); fr(w.top.blk); fr(w.top.cblk); fr(w.top.plan); fr(w.top.epoch_dev);
    delete &w;
    e.bh = nullptr;
}

}  // namespace nb
