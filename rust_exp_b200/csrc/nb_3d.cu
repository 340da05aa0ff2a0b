// nb_3d.cu -- the 3-D extension surface nbx3_* (SURVEY.md D1/D2): all-pairs accelerations in three
// dimensions with a selectable pair law.
//
// The reference (rs-src/nbody.rs) is 2-D and uses the un-normalised law F = m1 m2 d / (|d|^2 + EPS); the
// project brief (BASELINE.json north_star) describes a 3-D, rsqrt-based Newtonian kernel.  This file is that
// kernel, kept OFF the drop-in surface: its state is separate from the nb_* particle set.
//   LAW_NEWTON  a_i = sum_j m_j d_ij / (|d_ij|^2 + eps2)^(3/2)      one MUFU.RSQ per pair, 20 flop/pair
//   LAW_REF     a_i = sum_j m_j d_ij / (|d_ij|^2 + eps2)            the reference's law lifted to 3-D
// Parity: LAW_REF with z == 0 reproduces the 2-D FAST kernel bit for bit (every z term is an exact zero);
// LAW_NEWTON has no reference counterpart and is validated against an f64 evaluation of the same law.
// Same structure as allpairs_fast_kernel: persistent CTAs, a producer warp bulk-TMA-streaming x|y|z|m tiles
// through an mbarrier ring, consumers on broadcast LDS.128 + packed FP32, deterministic per-slice partials.
// Integrator: the reference's semi-implicit Euler (v += dt*a ; p += dt*v), rs-src/nbody.rs:153-160.
#include <algorithm>

#include "nb_engine.h"

namespace nb {

constexpr int k3TJ = 512;
constexpr int k3Stages = 3;
constexpr int k3Warps = 8;
constexpr int k3Compute = k3Warps * 32;
constexpr int k3Threads = k3Compute + 32;
constexpr int k3I = 2;                     // bodies per thread
constexpr int k3TI = k3Compute * k3I;

struct __align__(128) Smem3 {
    float x[k3Stages][k3TJ], y[k3Stages][k3TJ], z[k3Stages][k3TJ], m[k3Stages][k3TJ];
    uint64_t full[k3Stages], empty[k3Stages];
};

struct Args3 {
    const float *x, *y, *z, *m;   // padded to n_pad (multiple of 1024); padding has m = 0
    int n, n_pad;
    int slice_len, nslices;
    float eps2;
    float4* partial;              // [nslices][n_pad] (ax, ay, az, -)
    int itile_begin, n_itiles;    // the i tiles this launch evaluates (all of them, or this rank's rows when sharded)
};

template <int LAW>
__device__ __forceinline__ void pair2_3d(float xa, float xb, float ya, float yb, float za, float zb, float ma, float mb,
                                         float nxi, float nyi, float nzi, float2 eps, float2& ax, float2& ay, float2& az) {
    const float2 dx = __fadd2_rn(make_float2(xa, xb), make_float2(nxi, nxi));
    const float2 dy = __fadd2_rn(make_float2(ya, yb), make_float2(nyi, nyi));
    const float2 dz = __fadd2_rn(make_float2(za, zb), make_float2(nzi, nzi));
    float2 d2 = __ffma2_rn(dx, dx, eps);
    d2 = __ffma2_rn(dy, dy, d2);
    d2 = __ffma2_rn(dz, dz, d2);
    float2 s;
    if (LAW == NBX3_LAW_NEWTON) {
        const float2 r = make_float2(rsqrt_approx(d2.x), rsqrt_approx(d2.y));
        const float2 r3 = __fmul2_rn(__fmul2_rn(r, r), r);
        s = __fmul2_rn(make_float2(ma, mb), r3);
    } else {
        s = __fmul2_rn(make_float2(ma, mb), make_float2(rcp_approx(d2.x), rcp_approx(d2.y)));
    }
    ax = __ffma2_rn(s, dx, ax);
    ay = __ffma2_rn(s, dy, ay);
    az = __ffma2_rn(s, dz, az);
}

template <int LAW>
__global__ void __launch_bounds__(k3Threads, 3) allpairs3_kernel(const Args3 a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem3& sm = *reinterpret_cast<Smem3*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < k3Stages; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], k3Warps); }
        mbar_fence_init();
    }
    __syncthreads();
    const int n_itiles = a.n_itiles;
    const int n_items = n_itiles * a.nslices;
    if (warp == k3Warps) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
                const int sl = item / n_itiles;
                const int j0 = sl * a.slice_len, j1 = min(j0 + a.slice_len, a.n_pad);
                for (int j = j0; j < j1; j += k3TJ, it++) {
                    const uint32_t s = it % k3Stages, ph = (it / k3Stages) & 1u;
                    mbar_wait(&sm.empty[s], ph ^ 1u);
                    mbar_arrive_expect_tx(&sm.full[s], 4u * k3TJ * sizeof(float));
                    tma_load_1d(sm.x[s], a.x + j, k3TJ * sizeof(float), &sm.full[s]);
                    tma_load_1d(sm.y[s], a.y + j, k3TJ * sizeof(float), &sm.full[s]);
                    tma_load_1d(sm.z[s], a.z + j, k3TJ * sizeof(float), &sm.full[s]);
                    tma_load_1d(sm.m[s], a.m + j, k3TJ * sizeof(float), &sm.full[s]);
                }
            }
        }
        return;
    }
    const float2 eps = make_float2(a.eps2, a.eps2);
    uint32_t it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int sl = item / n_itiles, itile = a.itile_begin + (item - sl * n_itiles);
        const int j0 = sl * a.slice_len, j1 = min(j0 + a.slice_len, a.n_pad);
        const int ntiles = (j1 - j0) / k3TJ;
        float nxi[k3I], nyi[k3I], nzi[k3I];
        float2 ax[k3I], ay[k3I], az[k3I];
#pragma unroll
        for (int k = 0; k < k3I; k++) {
            const int i = itile * k3TI + k * k3Compute + threadIdx.x;   // < n_pad
            nxi[k] = -a.x[i]; nyi[k] = -a.y[i]; nzi[k] = -a.z[i];
            ax[k] = ay[k] = az[k] = make_float2(0.f, 0.f);
        }
        for (int t = 0; t < ntiles; t++, it++) {
            const uint32_t s = it % k3Stages, ph = (it / k3Stages) & 1u;
            mbar_wait(&sm.full[s], ph);
            const float4* __restrict__ px = reinterpret_cast<const float4*>(sm.x[s]);
            const float4* __restrict__ py = reinterpret_cast<const float4*>(sm.y[s]);
            const float4* __restrict__ pz = reinterpret_cast<const float4*>(sm.z[s]);
            const float4* __restrict__ pm = reinterpret_cast<const float4*>(sm.m[s]);
#pragma unroll 8
            for (int q4 = 0; q4 < k3TJ / 4; q4++) {
                const float4 X = px[q4], Y = py[q4], Z = pz[q4], M = pm[q4];
#pragma unroll
                for (int k = 0; k < k3I; k++) {
                    pair2_3d<LAW>(X.x, X.y, Y.x, Y.y, Z.x, Z.y, M.x, M.y, nxi[k], nyi[k], nzi[k], eps, ax[k], ay[k], az[k]);
                    pair2_3d<LAW>(X.z, X.w, Y.z, Y.w, Z.z, Z.w, M.z, M.w, nxi[k], nyi[k], nzi[k], eps, ax[k], ay[k], az[k]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[s]);
        }
        float4* out = a.partial + static_cast<size_t>(sl) * a.n_pad;
#pragma unroll
        for (int k = 0; k < k3I; k++) {
            const int i = itile * k3TI + k * k3Compute + threadIdx.x;
            out[i] = make_float4(ax[k].x + ax[k].y, ay[k].x + ay[k].y, az[k].x + az[k].y, 0.f);
        }
    }
}

// rows [i_begin, i_end) only (this rank's rows when sharded)
__global__ void integrate3_kernel(const float4* __restrict__ partial, int nslices, int n_pad, int i_begin, int i_end, float dt, int update,
                                  float* x, float* y, float* z, float* vx, float* vy, float* vz, float* acc_out) {
    const int i = i_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= i_end) return;
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int s = 0; s < nslices; s++) {
        const float4 p = partial[static_cast<size_t>(s) * n_pad + i];
        ax += p.x; ay += p.y; az += p.z;
    }
    if (acc_out) { acc_out[3 * i] = ax; acc_out[3 * i + 1] = ay; acc_out[3 * i + 2] = az; }
    if (update) {
        const float qx = fmaf(dt, ax, vx[i]), qy = fmaf(dt, ay, vy[i]), qz = fmaf(dt, az, vz[i]);
        vx[i] = qx; vy[i] = qy; vz[i] = qz;
        x[i] = fmaf(dt, qx, x[i]); y[i] = fmaf(dt, qy, y[i]); z[i] = fmaf(dt, qz, z[i]);
    }
}

__global__ void aos7_to_soa_kernel(const float* __restrict__ aos, int n, int n_pad, float* x, float* y, float* z, float* vx,
                                   float* vy, float* vz, float* m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    float v[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (i < n) for (int k = 0; k < 7; k++) v[k] = aos[7 * static_cast<size_t>(i) + k];
    x[i] = v[0]; y[i] = v[1]; z[i] = v[2]; vx[i] = v[3]; vy[i] = v[4]; vz[i] = v[5]; m[i] = v[6];
}
__global__ void soa_to_aos7_kernel(float* __restrict__ aos, int n, const float* x, const float* y, const float* z,
                                   const float* vx, const float* vy, const float* vz, const float* m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* r = aos + 7 * static_cast<size_t>(i);
    r[0] = x[i]; r[1] = y[i]; r[2] = z[i]; r[3] = vx[i]; r[4] = vy[i]; r[5] = vz[i]; r[6] = m[i];
}

// Sharding (round 2): with the process group wired (nbx_dist_init + nbx_dist_nccl_init) every rank keeps the whole
// position set, evaluates and integrates only its own rows [rank * shard, (rank + 1) * shard) and the new positions
// are exchanged with one in-place ncclAllGather per coordinate -- the "index shards + per-step NCCL all-gather of
// positions" the project brief describes.  The j-slice plan stays the single-GPU one, so the sharded result equals
// the single-GPU result bit for bit.  Velocities stay with their rows and are gathered by nbx3_get_particles.
struct State3 {
    int n = 0, n_pad = 0, cap = 0;
    int shard = 0;                 // rows per rank when sharded (multiple of 1024), else 0
    bool want_shard = true;
    float* buf = nullptr;          // 7 arrays of cap floats
    float4* partial = nullptr;
    size_t partial_cap = 0;
    float* stage = nullptr;        // device staging for AoS / accelerations
    size_t stage_cap = 0;
    float eps2 = kEps;
    int law = NBX3_LAW_NEWTON;
    float* arr(int k) const { return buf + static_cast<size_t>(k) * cap; }
};

static State3& st3(Engine& e) {
    if (!e.ext3) e.ext3 = new State3();
    return *static_cast<State3*>(e.ext3);
}

static void stage3(Engine& e, State3& s, size_t floats) {
    if (floats > s.stage_cap) {
        NB_CUDA(cudaStreamSynchronize(e.stream));
        if (s.stage) NB_CUDA(cudaFree(s.stage));
        NB_CUDA(cudaMalloc(&s.stage, floats * sizeof(float)));
        s.stage_cap = floats;
    }
}

static bool sharded3(const Engine& e, const State3& s) { return s.want_shard && e.dist && e.world > 1 && dist_nccl_ready(); }

void x3_set_sharded(Engine& e, int on) {
    State3& s = st3(e);
    if (e.dist && e.world > 1 && s.shard > 0 && !on) {
        // leaving sharded mode: every rank needs every velocity from here on
        float* v[3] = {s.arr(3), s.arr(4), s.arr(5)};
        dist_nccl_allgather_inplace(e, v, 3, static_cast<size_t>(s.shard));
        s.shard = 0;
    }
    s.want_shard = on != 0;
    if (s.want_shard && s.n_pad > 0 && sharded3(e, s)) {
        const int shard = ((s.n_pad / 1024 + e.world - 1) / e.world) * 1024;
        if (static_cast<long long>(shard) * e.world <= s.cap) s.shard = shard;   // else: takes effect at the next nbx3_set_particles
    }
}

void x3_set(Engine& e, const float* aos7, int n) {
    State3& s = st3(e);
    if (n < 0) n = 0;
    const int n_pad = ((n + 1023) / 1024) * 1024 + (n == 0 ? 1024 : 0);
    // sharded: the arrays are the receive buffers of the all-gather, world * shard long (>= n_pad)
    const int shard = sharded3(e, s) ? ((n_pad / 1024 + e.world - 1) / e.world) * 1024 : 0;
    const int n_alloc = shard ? shard * e.world : n_pad;
    if (n_alloc > s.cap) {
        NB_CUDA(cudaStreamSynchronize(e.stream));
        if (s.buf) NB_CUDA(cudaFree(s.buf));
        NB_CUDA(cudaMalloc(&s.buf, 7 * sizeof(float) * static_cast<size_t>(n_alloc)));
        s.cap = n_alloc;
    }
    s.n = n; s.n_pad = n_pad; s.shard = shard;
    stage3(e, s, 7 * static_cast<size_t>(n > 0 ? n : 1));
    if (n > 0) NB_CUDA(cudaMemcpyAsync(s.stage, aos7, 7 * sizeof(float) * static_cast<size_t>(n), cudaMemcpyHostToDevice, e.stream));
    aos7_to_soa_kernel<<<(n_alloc + 255) / 256, 256, 0, e.stream>>>(s.stage, n, n_alloc, s.arr(0), s.arr(1), s.arr(2), s.arr(3), s.arr(4),
                                                                 s.arr(5), s.arr(6));
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
}

void x3_get(Engine& e, float* aos7, int n) {
    State3& s = st3(e);
    if (n > s.n) n = s.n;
    if (n <= 0) return;
    if (s.shard) {   // velocities live with their rows: collect them (collective call)
        float* v[3] = {s.arr(3), s.arr(4), s.arr(5)};
        dist_nccl_allgather_inplace(e, v, 3, static_cast<size_t>(s.shard));
    }
    stage3(e, s, 7 * static_cast<size_t>(n));
    soa_to_aos7_kernel<<<(n + 255) / 256, 256, 0, e.stream>>>(s.stage, n, s.arr(0), s.arr(1), s.arr(2), s.arr(3), s.arr(4), s.arr(5), s.arr(6));
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
    NB_CUDA(cudaMemcpyAsync(aos7, s.stage, 7 * sizeof(float) * static_cast<size_t>(n), cudaMemcpyDeviceToHost, e.stream));
    NB_CUDA(cudaStreamSynchronize(e.stream));
}

int x3_num(Engine& e) { return st3(e).n; }
void x3_config(Engine& e, int law, float eps2) { State3& s = st3(e); s.law = law; s.eps2 = eps2; }

// one all-pairs evaluation; update != 0 integrates, acc_host != nullptr returns accelerations (3 floats per body)
void x3_step(Engine& e, float dt, int update, float* acc_host) {
    State3& s = st3(e);
    if (s.n == 0) return;
    static bool attr_set = false;
    if (!attr_set) {
        NB_CUDA(cudaFuncSetAttribute(allpairs3_kernel<NBX3_LAW_NEWTON>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(Smem3))));
        NB_CUDA(cudaFuncSetAttribute(allpairs3_kernel<NBX3_LAW_REF>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(Smem3))));
        attr_set = true;
    }
    // the 2-D kernel's own decomposition rule, so that <3,REF> with z = 0 groups its sums exactly like it
    // (always the plan of the whole set: a sharded step then adds up exactly what the single-GPU step adds up)
    int i_begin = 0, i_end = s.n;
    if (s.shard) {
        i_begin = std::min(s.n, e.rank * s.shard);
        i_end = std::min(s.n, i_begin + s.shard);
    }
    const int itile_begin = i_begin / k3TI;               // shard is a multiple of 1024
    const int n_itiles = i_end > i_begin ? (i_end + k3TI - 1) / k3TI - itile_begin : 0;
    int slice_len = 0, per = 1;
    allpairs_slices(e, s.n, s.n_pad, 1, &slice_len, &per);
    const size_t need = static_cast<size_t>(per) * s.n_pad;
    if (need > s.partial_cap) {
        NB_CUDA(cudaStreamSynchronize(e.stream));
        if (s.partial) NB_CUDA(cudaFree(s.partial));
        NB_CUDA(cudaMalloc(&s.partial, need * sizeof(float4)));
        s.partial_cap = need;
    }
    Args3 a{s.arr(0), s.arr(1), s.arr(2), s.arr(6), s.n, s.n_pad, slice_len, per, s.eps2, s.partial, itile_begin, n_itiles};
    int grid = e.num_sms * 3;
    if (grid > n_itiles * per) grid = n_itiles * per;
    if (grid > 0) {
        if (s.law == NBX3_LAW_NEWTON) allpairs3_kernel<NBX3_LAW_NEWTON><<<grid, k3Threads, sizeof(Smem3), e.stream>>>(a);
        else allpairs3_kernel<NBX3_LAW_REF><<<grid, k3Threads, sizeof(Smem3), e.stream>>>(a);
        NB_CUDA(cudaGetLastError());
        e.ctr.kernel_launches++;
    }
    float* acc_dev = nullptr;
    if (acc_host) { stage3(e, s, 3 * static_cast<size_t>(s.shard ? s.shard * e.world : s.n) + 4); acc_dev = s.stage; }
    if (i_end > i_begin) {
        integrate3_kernel<<<(i_end - i_begin + 255) / 256, 256, 0, e.stream>>>(s.partial, per, s.n_pad, i_begin, i_end, dt, update, s.arr(0),
                                                                               s.arr(1), s.arr(2), s.arr(3), s.arr(4), s.arr(5), acc_dev);
        NB_CUDA(cudaGetLastError());
        e.ctr.kernel_launches++;
    }
    e.ctr.allpairs_pairs += static_cast<uint64_t>(i_end - i_begin) * static_cast<uint64_t>(s.n - 1);
    if (s.shard) {
        if (update) {   // the new positions of every rank's rows, in place
            float* pos[3] = {s.arr(0), s.arr(1), s.arr(2)};
            dist_nccl_allgather_inplace(e, pos, 3, static_cast<size_t>(s.shard));
        }
        if (acc_dev) dist_nccl_allgather_inplace(e, &acc_dev, 1, 3 * static_cast<size_t>(s.shard));   // rows are 3 floats, shards contiguous
    }
    if (acc_host) {
        NB_CUDA(cudaMemcpyAsync(acc_host, acc_dev, 3 * sizeof(float) * static_cast<size_t>(s.n), cudaMemcpyDeviceToHost, e.stream));
        NB_CUDA(cudaStreamSynchronize(e.stream));
    }
}

void x3_shutdown(Engine& e) {
    if (!e.ext3) return;
    State3& s = st3(e);
    if (s.buf) cudaFree(s.buf);
    if (s.partial) cudaFree(s.partial);
    if (s.stage) cudaFree(s.stage);
    delete &s;
    e.ext3 = nullptr;
}

}  // namespace nb
