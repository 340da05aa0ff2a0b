// nb_allpairs.cu -- the O(N^2) force-accumulation kernels (rs-src/nbody.rs:129-144, :164-184).
//
// Two kernels, one semantics:
//
//  * allpairs_fast_kernel<I>  (NBX_MODE_FAST) -- the throughput path.  Persistent, warp-specialised:
//    one producer warp streams j tiles (SoA x|y|m, 3 x 2 KB) from HBM/L2 -- or straight from a PEER
//    GPU's HBM over NVLink -- into a 3-stage shared-memory ring with 1-D bulk TMA (cp.async.bulk ->
//    SASS UBLKCP) + mbarrier transaction counts; eight consumer warps read the tile with broadcast
//    LDS.128 (4 j per load per array) and evaluate pairs two at a time with Blackwell packed-FP32
//    instructions (FADD2/FFMA2/FMUL2) and one MUFU.RCP per pair.  Per pair this is 3.5 FMA-pipe issue
//    slots + 1 MUFU, against 7 + 1 for scalar code: the kernel is bound by the MUFU pipe
//    (16 lanes/SM/clk => 4 pairs/clk/SMSP), see DESIGN.md section 4.
//    It computes a_i = sum_j m_j * d_ij / (|d_ij|^2 + EPS)   (= F_i/m_i of the reference, algebraically),
//    as partial sums per j slice; the integrate kernel adds the slices in a fixed order, so results are
//    run-to-run deterministic and independent of the CTA schedule.
//    No i==j test is needed: d_ii = 0 exactly and EPS > 0, so the self term is an exact +0.
//
//  * allpairs_exact_kernel (NBX_MODE_EXACT) -- the semantics pin.  One thread per body i, j ascending,
//    every operation a separately rounded IEEE binary32 op in the reference's order ((m1*m2)/(d2+EPS),
//    true division, i==j skipped).  Bit-identical to the CPU restatement of rs-src/nbody.rs.
#include "nb_engine.h"

namespace nb {

constexpr int kTJ = 512;             // j bodies per shared-memory tile
constexpr int kStages = 3;           // TMA ring depth
#ifndef NB_APQ_UNROLL
#define NB_APQ_UNROLL 16
#endif
constexpr int kQuadUnroll = NB_APQ_UNROLL;   // j quads per unrolled inner-loop body
constexpr int kDefaultShare = 0;     // reciprocal sharing (see pair4_shared_rcp); measured choice
constexpr int kComputeWarps = 8;
constexpr int kComputeThreads = kComputeWarps * 32;
constexpr int kThreads = kComputeThreads + 32;  // + 1 producer warp
constexpr uint32_t kTileBytes = 3u * kTJ * sizeof(float);
static_assert(kShardAlign % kTJ == 0, "j tiles must divide the shard granularity");

struct __align__(128) FastSmem {
    float x[kStages][kTJ];
    float y[kStages][kTJ];
    float m[kStages][kTJ];
    uint64_t full[kStages];
    uint64_t empty[kStages];
};

// Two pairs (one i body against j bodies a,b) in packed FP32.
//   dx = xj - xi ; d2 = dx*dx + (dy*dy + EPS) ; s = mj * rcp(d2) ; acc += s * d
// 2 FADD2 + 2 FFMA2 + 2 MUFU.RCP + 1 FMUL2 + 2 FFMA2.
__device__ __forceinline__ void pair2(float xa, float xb, float ya, float yb, float ma, float mb,
                                      const float2 nxi, const float2 nyi, float2& ax, float2& ay) {
    const float2 dx = __fadd2_rn(make_float2(xa, xb), nxi);
    const float2 dy = __fadd2_rn(make_float2(ya, yb), nyi);
    float2 d2 = __ffma2_rn(dx, dx, make_float2(kEps, kEps));
    d2 = __ffma2_rn(dy, dy, d2);
    const float2 inv = make_float2(rcp_approx(d2.x), rcp_approx(d2.y));
    const float2 s = __fmul2_rn(make_float2(ma, mb), inv);
    ax = __ffma2_rn(s, dx, ax);
    ay = __ffma2_rn(s, dy, ay);
}

// Four pairs (one i body against j bodies a,b,c,d) sharing TWO reciprocals: with A = (d2a,d2b), B = (d2c,d2d),
// R = rcp(A*B) gives 1/A = R*B and 1/B = R*A -- 3 extra FMUL2 (6 FMA-pipe cycles) buy back 2 MUFU.RCP
// (16 MUFU-pipe cycles).  d2+EPS lies in [1e-4, ~1e5], so the product neither under- nor overflows.
// Used for a fraction of the quads to balance the two pipes.
__device__ __forceinline__ void pair4_shared_rcp(const float4 X, const float4 Y, const float4 M, const float2 nxi,
                                                 const float2 nyi, float2& ax, float2& ay) {
    const float2 dxa = __fadd2_rn(make_float2(X.x, X.y), nxi), dxb = __fadd2_rn(make_float2(X.z, X.w), nxi);
    const float2 dya = __fadd2_rn(make_float2(Y.x, Y.y), nyi), dyb = __fadd2_rn(make_float2(Y.z, Y.w), nyi);
    const float2 A = __ffma2_rn(dya, dya, __ffma2_rn(dxa, dxa, make_float2(kEps, kEps)));
    const float2 B = __ffma2_rn(dyb, dyb, __ffma2_rn(dxb, dxb, make_float2(kEps, kEps)));
    const float2 P = __fmul2_rn(A, B);
    const float2 R = make_float2(rcp_approx(P.x), rcp_approx(P.y));
    const float2 sa = __fmul2_rn(make_float2(M.x, M.y), __fmul2_rn(R, B));
    const float2 sb = __fmul2_rn(make_float2(M.z, M.w), __fmul2_rn(R, A));
    ax = __ffma2_rn(sa, dxa, ax); ay = __ffma2_rn(sa, dya, ay);
    ax = __ffma2_rn(sb, dxb, ax); ay = __ffma2_rn(sb, dyb, ay);
}

template <int I, int MINB, int SHARE>
__global__ void __launch_bounds__(kThreads, MINB) allpairs_fast_kernel(const AllPairsArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FastSmem& sm = *reinterpret_cast<FastSmem*>(smem_raw);

    constexpr int TI = kComputeThreads * I;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; s++) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kComputeWarps);
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int n_itiles = (a.n_local + TI - 1) / TI;
    const int total_slices = a.nseg * a.slices_per_seg;
    const int n_items = n_itiles * total_slices;
    const int tiles_full_slice = a.slice_len / kTJ;

    if (a.stage && blockIdx.x < 2 * (a.nseg - 1)) {
        // ===== stager prologue: this CTA brings one half of one remote segment into local HBM, once per step =====
        const int g = (a.my_rank + 1 + (blockIdx.x >> 1)) % a.nseg;
        if (threadIdx.x == 0) wait_epoch(a.flags + g, a.wait_step, a.timeout_ns, g, "step");   // the peer finished writing it
        __syncthreads();
        const int L4 = a.seg_len >> 2, lo = (blockIdx.x & 1) ? (L4 >> 1) : 0, hi = (blockIdx.x & 1) ? L4 : (L4 >> 1);
        const float4* __restrict__ sx = reinterpret_cast<const float4*>(a.src[g].x);
        const float4* __restrict__ sy = reinterpret_cast<const float4*>(a.src[g].y);
        const float4* __restrict__ sm4 = reinterpret_cast<const float4*>(a.src[g].m);
        float4* dx = reinterpret_cast<float4*>(const_cast<float*>(a.dst[g].x));
        float4* dy = reinterpret_cast<float4*>(const_cast<float*>(a.dst[g].y));
        float4* dm = reinterpret_cast<float4*>(const_cast<float*>(a.dst[g].m));
        for (int i = lo + threadIdx.x; i < hi; i += 4 * kThreads) {
            float4 vx[4], vy[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const int k = i + u * kThreads; if (k < hi) { vx[u] = sx[k]; vy[u] = sy[k]; } }
#pragma unroll
            for (int u = 0; u < 4; u++) { const int k = i + u * kThreads; if (k < hi) { dx[k] = vx[u]; dy[k] = vy[u]; } }
            if (a.stage_mass) {
#pragma unroll
                for (int u = 0; u < 4; u++) { const int k = i + u * kThreads; if (k < hi) dm[k] = sm4[k]; }
            }
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(a.staged + g, 1u);
    }

    if (warp == kComputeWarps) {
        // ===== producer warp: one elected lane feeds the ring for every item of this CTA =====
        if (lane == 0) {
            uint32_t it = 0;  // running tile counter -> stage / phase
            uint32_t seen_mask = 1u << a.my_rank;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int gslice = item / n_itiles;  // items are slice-major: local segment first
                const int q = gslice / a.slices_per_seg;
                const int sl = gslice - q * a.slices_per_seg;
                const int g = (a.my_rank + q) % a.nseg;
                if (a.flags != nullptr && !(seen_mask & (1u << g))) {
                    if (a.stage) wait_epoch(a.staged + g, a.stage_want, a.timeout_ns, g, "staged halves");   // local flag
                    else wait_epoch(a.flags + g, a.wait_step, a.timeout_ns, g, "step");
                    seen_mask |= 1u << g;
                }
                const int j0 = sl * a.slice_len;
                const int j1 = min(j0 + a.slice_len, a.seg_len);
                const JSeg src = a.seg[g];
                for (int j = j0; j < j1; j += kTJ, it++) {
                    const uint32_t s = it % kStages;
                    const uint32_t ph = (it / kStages) & 1u;
                    mbar_wait(&sm.empty[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&sm.full[s], kTileBytes);
                    tma_load_1d(sm.x[s], src.x + j, kTJ * sizeof(float), &sm.full[s]);
                    tma_load_1d(sm.y[s], src.y + j, kTJ * sizeof(float), &sm.full[s]);
                    tma_load_1d(sm.m[s], src.m + j, kTJ * sizeof(float), &sm.full[s]);
                }
            }
        }
        return;
    }

    // ===== consumer warps =====
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int gslice = item / n_itiles;
        const int itile = item - gslice * n_itiles;
        const int sl = gslice % a.slices_per_seg;
        const int j0 = sl * a.slice_len;
        const int j1 = min(j0 + a.slice_len, a.seg_len);
        const int ntiles = (j1 - j0) / kTJ;
        (void)tiles_full_slice;

        float2 nxi[I], nyi[I], ax[I], ay[I];
#pragma unroll
        for (int k = 0; k < I; k++) {
            const int i = itile * TI + k * kComputeThreads + threadIdx.x;  // < seg_len (TI | kShardAlign)
            const float xi = a.xi[i], yi = a.yi[i];
            nxi[k] = make_float2(-xi, -xi);
            nyi[k] = make_float2(-yi, -yi);
            ax[k] = make_float2(0.f, 0.f);
            ay[k] = make_float2(0.f, 0.f);
        }

        for (int t = 0; t < ntiles; t++, it++) {
            const uint32_t s = it % kStages;
            const uint32_t ph = (it / kStages) & 1u;
            mbar_wait(&sm.full[s], ph);
            const float4* __restrict__ px = reinterpret_cast<const float4*>(sm.x[s]);
            const float4* __restrict__ py = reinterpret_cast<const float4*>(sm.y[s]);
            const float4* __restrict__ pm = reinterpret_cast<const float4*>(sm.m[s]);
#pragma unroll kQuadUnroll
            for (int q4 = 0; q4 < kTJ / 4; q4++) {
                const float4 X = px[q4], Y = py[q4], M = pm[q4];  // broadcast LDS.128
#pragma unroll
                for (int k = 0; k < I; k++) {
                    // SHARE: 0 = one MUFU.RCP per pair; 2 = every quad shares reciprocals; 1 = every other quad
                    if (SHARE == 2 || (SHARE == 1 && (q4 & 1))) {
                        pair4_shared_rcp(X, Y, M, nxi[k], nyi[k], ax[k], ay[k]);
                    } else {
                        pair2(X.x, X.y, Y.x, Y.y, M.x, M.y, nxi[k], nyi[k], ax[k], ay[k]);
                        pair2(X.z, X.w, Y.z, Y.w, M.z, M.w, nxi[k], nyi[k], ax[k], ay[k]);
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[s]);
        }

        float2* out = a.partial + static_cast<size_t>(gslice) * a.seg_len;
#pragma unroll
        for (int k = 0; k < I; k++) {
            const int i = itile * TI + k * kComputeThreads + threadIdx.x;
            out[i] = make_float2(ax[k].x + ax[k].y, ay[k].x + ay[k].y);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Exact kernel: rs-src/nbody.rs:132-144 with force() of :164-184, one rounding per operation.
// ---------------------------------------------------------------------------------------------
constexpr int kExactThreads = 128;

__global__ void __launch_bounds__(kExactThreads) allpairs_exact_kernel(const AllPairsArgs a, float2* __restrict__ force_out) {
    __shared__ float sx[kExactThreads], sy[kExactThreads], smass[kExactThreads];
    const int il = blockIdx.x * kExactThreads + threadIdx.x;
    const bool live = il < a.n_local;
    const int ig = a.i_global_begin + il;
    const float xi = live ? a.xi[il] : 0.f, yi = live ? a.yi[il] : 0.f, mi = live ? a.mi[il] : 0.f;
    float fx = 0.0f, fy = 0.0f;
    for (int g = 0; g < a.nseg; g++) {
        const int gbase = g * a.seg_len;
        const int cnt = min(a.seg_len, a.n_total - gbase);  // real bodies in this segment
        const JSeg src = a.seg[g];
        for (int j0 = 0; j0 < cnt; j0 += kExactThreads) {
            const int jl = j0 + threadIdx.x;
            __syncthreads();
            if (jl < cnt) {
                sx[threadIdx.x] = src.x[jl];
                sy[threadIdx.x] = src.y[jl];
                smass[threadIdx.x] = src.m[jl];
            }
            __syncthreads();
            const int lim = min(kExactThreads, cnt - j0);
            for (int jj = 0; jj < lim; jj++) {
                if (gbase + j0 + jj == ig) continue;  // rs-src/nbody.rs:136
                const float dx = __fsub_rn(sx[jj], xi);  // rs-src/nbody.rs:174
                const float dy = __fsub_rn(sy[jj], yi);
                const float dist_sq = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
                const float f = __fdiv_rn(__fmul_rn(mi, smass[jj]), __fadd_rn(dist_sq, kEps));  // :180
                fx = __fadd_rn(fx, __fmul_rn(f, dx));  // :141, :183
                fy = __fadd_rn(fy, __fmul_rn(f, dy));
            }
        }
    }
    if (live) force_out[il] = make_float2(fx, fy);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
static int pick_bodies_per_thread(const Engine& e) {
    if (e.tune.bodies_per_thread == 1 || e.tune.bodies_per_thread == 2 || e.tune.bodies_per_thread == 4)
        return e.tune.bodies_per_thread;
    return 4;   // measured best with 2 CTAs/SM and the 8-quad unroll (profiles/r01_sweep_unroll.txt)
}
static int pick_ctas_per_sm(const Engine& e, int I) {
    int c = e.tune.ctas_per_sm > 0 ? e.tune.ctas_per_sm : (I == 1 ? 5 : (I == 2 ? 4 : 2));  // measured best, profiles/
    // clamp to the instantiated variants
    if (I == 1) c = c < 4 ? 4 : (c > 5 ? 5 : c);
    else if (I == 4) c = c < 2 ? 2 : (c > 3 ? 3 : c);
    else c = c < 3 ? 3 : (c > 4 ? 4 : c);
    return c;
}

// j-slice decomposition of the work (shared with the 3-D extension so that <3,REF> with z = 0 groups its partial
// sums exactly like this kernel): enough (i tile, j slice) items for `W` waves over the resident CTAs.
void allpairs_slices(const Engine& e, int n_local, int L, int nseg, int* slice_len_out, int* per_seg_out) {
    const int I = pick_bodies_per_thread(e);
    const int TI = kComputeThreads * I;
    const int n_itiles = (n_local + TI - 1) / TI;
    const int R = e.num_sms * pick_ctas_per_sm(e, I);
    // measured (profiles/r01_sweep_*): many fine slices pay off once there are >= 256 i tiles; below that the
    // partial-sum traffic and per-item prologue outweigh the better tail
    const int W = e.tune.target_waves > 0 ? e.tune.target_waves : (n_itiles >= 256 ? 64 : 16);
    int want_total = (W * R + n_itiles - 1) / (n_itiles > 0 ? n_itiles : 1);
    int per_seg = (want_total + nseg - 1) / nseg;
    const int max_per_seg = L / kTJ;
    if (per_seg < 1) per_seg = 1;
    if (per_seg > max_per_seg) per_seg = max_per_seg;
    int slice_len = (L + per_seg - 1) / per_seg;
    slice_len = ((slice_len + kTJ - 1) / kTJ) * kTJ;
    per_seg = (L + slice_len - 1) / slice_len;
    *slice_len_out = slice_len;
    *per_seg_out = per_seg;
}

void allpairs_plan(Engine& e, AllPairsArgs& a) {
    const int L = static_cast<int>(e.lay.L);
    a.nseg = e.world;
    a.seg_len = L;
    a.n_total = e.n;
    a.i_global_begin = local_begin(e);
    a.n_local = local_count(e);
    a.my_rank = e.rank;
    a.timeout_ns = e.peer_timeout_ns;
    allpairs_slices(e, a.n_local, L, a.nseg, &a.slice_len, &a.slices_per_seg);
    const size_t need = static_cast<size_t>(a.nseg) * a.slices_per_seg * L;
    if (need > e.partial_cap) {
        if (e.partial) NB_CUDA(cudaFree(e.partial));
        NB_CUDA(cudaMalloc(&e.partial, need * sizeof(float2)));
        e.partial_cap = need;
    }
    a.partial = e.partial;
}

template <int I, int MINB, int SHARE>
static void launch_fast_t(Engine& e, const AllPairsArgs& a) {
    static bool attr_set = false;
    if (!attr_set) {
        NB_CUDA(cudaFuncSetAttribute(allpairs_fast_kernel<I, MINB, SHARE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(sizeof(FastSmem))));
        attr_set = true;
    }
    constexpr int TI = kComputeThreads * I;
    const int n_itiles = (a.n_local + TI - 1) / TI;
    const int n_items = n_itiles * a.nseg * a.slices_per_seg;
    if (n_items == 0) return;
    int grid = e.num_sms * MINB;
    if (grid > n_items) grid = n_items;
    allpairs_fast_kernel<I, MINB, SHARE><<<grid, kThreads, sizeof(FastSmem), e.stream>>>(a);
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
}

template <int SHARE>
static void launch_fast_s(Engine& e, const AllPairsArgs& a, int I, int C) {
    // (bodies per thread, resident CTAs per SM) variants; the register budget follows from MINB
    if (I == 1) {
        if (C <= 4) launch_fast_t<1, 4, SHARE>(e, a);
        else launch_fast_t<1, 5, SHARE>(e, a);
    } else if (I == 4) {
        if (C <= 2) launch_fast_t<4, 2, SHARE>(e, a);
        else launch_fast_t<4, 3, SHARE>(e, a);
    } else {
        if (C <= 3) launch_fast_t<2, 3, SHARE>(e, a);
        else launch_fast_t<2, 4, SHARE>(e, a);
    }
}

void launch_allpairs_fast(Engine& e, const AllPairsArgs& a) {
    const int I = pick_bodies_per_thread(e);
    const int C = pick_ctas_per_sm(e, I);
    const int share = e.tune.share_rcp >= 0 ? e.tune.share_rcp : kDefaultShare;
    if (share == 0) launch_fast_s<0>(e, a, I, C);
    else if (share == 1) launch_fast_s<1>(e, a, I, C);
    else launch_fast_s<2>(e, a, I, C);
    e.ctr.allpairs_pairs += static_cast<uint64_t>(a.n_local) * static_cast<uint64_t>(a.n_total > 0 ? a.n_total - 1 : 0);
}

void launch_allpairs_exact(Engine& e, const AllPairsArgs& a, float2* force_out) {
    if (a.n_local == 0) return;
    const int grid = (a.n_local + kExactThreads - 1) / kExactThreads;
    allpairs_exact_kernel<<<grid, kExactThreads, 0, e.stream>>>(a, force_out);
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
    e.ctr.allpairs_pairs += static_cast<uint64_t>(a.n_local) * static_cast<uint64_t>(a.n_total > 0 ? a.n_total - 1 : 0);
}

}  // namespace nb
