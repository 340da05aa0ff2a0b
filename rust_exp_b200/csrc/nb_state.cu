// nb_state.cu -- particle-state kernels: AoS<->SoA conversion at the ABI boundary, the integrators
// (rs-src/nbody.rs:153-160, :453-471) and the device-side initial-condition generators
// (rs-src/nbody.rs:39-104).
#include "nb_engine.h"

namespace nb {

constexpr int kBlk = 256;

// ---------------------------------------------------------------------------------------------
// AoS {px,py,vx,vy,m} (rs-src/nbody.rs:19-26) <-> SoA shard.  `aos` holds the GLOBAL set; a rank
// keeps rows [begin, begin+count) and zero-fills its padding slots (m = 0 => no force, never moved).
// ---------------------------------------------------------------------------------------------
__global__ void aos_to_soa_kernel(const float* __restrict__ aos, int begin, int count, int L, float* __restrict__ x0,
                                  float* __restrict__ y0, float* __restrict__ x1, float* __restrict__ y1,
                                  float* __restrict__ m, float* __restrict__ vx, float* __restrict__ vy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    float px = 0.f, py = 0.f, qx = 0.f, qy = 0.f, mm = 0.f;
    if (i < count) {
        const float* r = aos + 5 * static_cast<size_t>(begin + i);
        px = r[0]; py = r[1]; qx = r[2]; qy = r[3]; mm = r[4];
    }
    x0[i] = px; y0[i] = py; x1[i] = px; y1[i] = py;
    m[i] = mm; vx[i] = qx; vy[i] = qy;
}

__global__ void soa_to_aos_kernel(float* __restrict__ aos, int count, const float* __restrict__ x,
                                  const float* __restrict__ y, const float* __restrict__ vx,
                                  const float* __restrict__ vy, const float* __restrict__ m) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float* r = aos + 5 * static_cast<size_t>(i);
    r[0] = x[i]; r[1] = y[i]; r[2] = vx[i]; r[3] = vy[i]; r[4] = m[i];
}

static void ensure_stage(Engine& e, size_t floats) {
    if (floats > e.stage_dev_cap) {
        if (e.stage_dev) NB_CUDA(cudaFree(e.stage_dev));
        NB_CUDA(cudaMalloc(&e.stage_dev, floats * sizeof(float)));
        e.stage_dev_cap = floats;
    }
}

void state_upload_aos(Engine& e, const float* aos5, int n) {
    // caller has already sized the arena for n.  Every rank receives the full array and keeps its shard.
    const int begin = local_begin(e), count = local_count(e);
    const int L = static_cast<int>(e.lay.L);
    if (count > 0) {
        ensure_stage(e, 5 * static_cast<size_t>(count));
        // copy only this rank's rows: host -> device staging (pageable source is fine; the pinned bounce
        // would cost a second host pass)
        NB_CUDA(cudaMemcpyAsync(e.stage_dev, aos5 + 5 * static_cast<size_t>(begin), 5 * sizeof(float) * count,
                                cudaMemcpyHostToDevice, e.stream));
    } else {
        ensure_stage(e, 5);
    }
    if (L > 0) {
        aos_to_soa_kernel<<<(L + kBlk - 1) / kBlk, kBlk, 0, e.stream>>>(
            e.stage_dev, 0, count, L, e.arena.x(e.lay, 0), e.arena.y(e.lay, 0), e.arena.x(e.lay, 1),
            e.arena.y(e.lay, 1), e.arena.m(e.lay), e.arena.vx(e.lay), e.arena.vy(e.lay));
        NB_CUDA(cudaGetLastError());
        e.ctr.kernel_launches++;
    }
    e.cur = 0;
    (void)n;
}

void state_download_aos(Engine& e, float* aos5, int n) {
    // Gather every rank's shard (peers are read over NVLink through their mapped arenas).
    if (n > e.n) n = e.n;
    if (n <= 0) return;
    dist_require_peers(e);
    ensure_stage(e, 5 * static_cast<size_t>(n));
    const int L = static_cast<int>(e.lay.L);
    for (int g = 0; g < e.world; g++) {
        const int b = g * L;
        int c = n - b;
        if (c > L) c = L;
        if (c <= 0) break;
        const ArenaView& av = (g == e.rank) ? e.arena : e.peer[g];
        soa_to_aos_kernel<<<(c + kBlk - 1) / kBlk, kBlk, 0, e.stream>>>(e.stage_dev + 5 * static_cast<size_t>(b), c,
                                                                       av.x(e.lay, e.cur), av.y(e.lay, e.cur),
                                                                       av.vx(e.lay), av.vy(e.lay), av.m(e.lay));
        NB_CUDA(cudaGetLastError());
        e.ctr.kernel_launches++;
    }
    NB_CUDA(cudaMemcpyAsync(aos5, e.stage_dev, 5 * sizeof(float) * static_cast<size_t>(n), cudaMemcpyDeviceToHost,
                            e.stream));
    NB_CUDA(cudaStreamSynchronize(e.stream));
}

// Owner-only read-back: this rank's rows [begin, begin+count) of the global AoS array, nothing else.  No peer
// memory is touched, so it is not a collective epoch; it orders after this rank's own enqueued steps.
void state_download_local(Engine& e, float* aos5_full, int n) {
    const int b = local_begin(e);
    int c = local_count(e);
    if (b + c > n) c = n - b;
    if (c <= 0) { NB_CUDA(cudaStreamSynchronize(e.stream)); return; }
    ensure_stage(e, 5 * static_cast<size_t>(c));
    soa_to_aos_kernel<<<(c + kBlk - 1) / kBlk, kBlk, 0, e.stream>>>(e.stage_dev, c, e.arena.x(e.lay, e.cur), e.arena.y(e.lay, e.cur),
                                                                   e.arena.vx(e.lay), e.arena.vy(e.lay), e.arena.m(e.lay));
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
    NB_CUDA(cudaMemcpyAsync(aos5_full + 5 * static_cast<size_t>(b), e.stage_dev, 5 * sizeof(float) * static_cast<size_t>(c),
                            cudaMemcpyDeviceToHost, e.stream));
    NB_CUDA(cudaStreamSynchronize(e.stream));
}

// ---------------------------------------------------------------------------------------------
// Integrators.  Both write the OTHER position buffer (ping-pong) so that peers may still be reading
// the current one.
// ---------------------------------------------------------------------------------------------

// FAST: a = sum of per-slice partial accelerations in ascending slice order (deterministic), then
// v += dt*a ; p += dt*v   (rs-src/nbody.rs:155-159 with F/m folded into a), optional velocity kill
// (rs-src/nbody.rs:466-471).
// dt_kick != dt only for the opt-in leapfrog (kick-drift-kick): the kick then closes the previous step's half kick and
// opens this step's, dt_kick = dt_prev/2 + dt/2 (nb_engine.cu::kick_dt).
__global__ void integrate_fast_kernel(const float2* __restrict__ partial, int nslices, int L, int count, float dt_kick, float dt,
                                      int kill, const float* __restrict__ xc, const float* __restrict__ yc,
                                      float* __restrict__ xn, float* __restrict__ yn, float* __restrict__ vx,
                                      float* __restrict__ vy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float ax = 0.f, ay = 0.f;
    for (int s = 0; s < nslices; s++) {
        const float2 p = partial[static_cast<size_t>(s) * L + i];
        ax += p.x;
        ay += p.y;
    }
    float qx = fmaf(dt_kick, ax, vx[i]);
    float qy = fmaf(dt_kick, ay, vy[i]);
    const float px = fmaf(dt, qx, xc[i]);
    const float py = fmaf(dt, qy, yc[i]);
    if (kill && (fabsf(px) > kKillLimit || fabsf(py) > kKillLimit)) { qx = 0.f; qy = 0.f; }
    vx[i] = qx; vy[i] = qy;
    xn[i] = px; yn[i] = py;
}

// EXACT: rs-src/nbody.rs:155-159 -- v += (dt*F)/m ; p += dt*v, each op rounded separately.
__global__ void integrate_exact_kernel(const float2* __restrict__ force, int count, float dt, int kill,
                                       const float* __restrict__ xc, const float* __restrict__ yc,
                                       const float* __restrict__ m, float* __restrict__ xn, float* __restrict__ yn,
                                       float* __restrict__ vx, float* __restrict__ vy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float2 f = force[i];
    const float mi = m[i];
    float qx = __fadd_rn(vx[i], __fdiv_rn(__fmul_rn(dt, f.x), mi));
    float qy = __fadd_rn(vy[i], __fdiv_rn(__fmul_rn(dt, f.y), mi));
    const float px = __fadd_rn(xc[i], __fmul_rn(dt, qx));
    const float py = __fadd_rn(yc[i], __fmul_rn(dt, qy));
    // rs-src/nbody.rs:466-467: (VP_ORG - p).abs() > VP_WDH*0.55 ; VP_ORG = 0 and 0 - p is exact
    if (kill && (fabsf(__fsub_rn(0.0f, px)) > __fmul_rn(kVpWdh, 0.55f) ||
                 fabsf(__fsub_rn(0.0f, py)) > __fmul_rn(kVpWdh, 0.55f))) {
        qx = 0.f; qy = 0.f;
    }
    vx[i] = qx; vy[i] = qy;
    xn[i] = px; yn[i] = py;
}

void launch_integrate_fast(Engine& e, const float2* partial, int nslices, float dt, bool kill) {
    const int count = local_count(e);
    const int nxt = e.cur ^ 1;
    if (count > 0) {
        integrate_fast_kernel<<<(count + kBlk - 1) / kBlk, kBlk, 0, e.stream>>>(
            partial, nslices, static_cast<int>(e.lay.L), count, kick_dt(e, dt), dt, kill ? 1 : 0, e.arena.x(e.lay, e.cur),
            e.arena.y(e.lay, e.cur), e.arena.x(e.lay, nxt), e.arena.y(e.lay, nxt), e.arena.vx(e.lay),
            e.arena.vy(e.lay));
        NB_CUDA(cudaGetLastError());
        e.ctr.kernel_launches++;
    }
    if (e.integrator == NBX_INTEGRATOR_LEAPFROG_KDK) e.kdk_pending = 0.5f * dt;   // the closing half kick is still owed
    e.cur = nxt;
}

// v += h * a for the local shard: the closing half kick of the leapfrog, applied before state is read back
__global__ void kick_kernel(const float2* __restrict__ acc, int count, float h, float* __restrict__ vx, float* __restrict__ vy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float2 a = acc[i];
    vx[i] = fmaf(h, a.x, vx[i]);
    vy[i] = fmaf(h, a.y, vy[i]);
}
void launch_kick(Engine& e, const float2* acc, float h) {
    const int count = local_count(e);
    if (count <= 0) return;
    kick_kernel<<<(count + kBlk - 1) / kBlk, kBlk, 0, e.stream>>>(acc, count, h, e.arena.vx(e.lay), e.arena.vy(e.lay));
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
}

void launch_integrate_exact(Engine& e, const float2* force, float dt, bool kill) {
    const int count = local_count(e);
    const int nxt = e.cur ^ 1;
    if (count > 0) {
        integrate_exact_kernel<<<(count + kBlk - 1) / kBlk, kBlk, 0, e.stream>>>(
            force, count, dt, kill ? 1 : 0, e.arena.x(e.lay, e.cur), e.arena.y(e.lay, e.cur), e.arena.m(e.lay),
            e.arena.x(e.lay, nxt), e.arena.y(e.lay, nxt), e.arena.vx(e.lay), e.arena.vy(e.lay));
        NB_CUDA(cudaGetLastError());
        e.ctr.kernel_launches++;
    }
    e.cur = nxt;
}

// accelerations for nbx_accelerations
__global__ void accel_from_partial_kernel(const float2* __restrict__ partial, int nslices, int L, int count,
                                          float2* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float ax = 0.f, ay = 0.f;
    for (int s = 0; s < nslices; s++) {
        const float2 p = partial[static_cast<size_t>(s) * L + i];
        ax += p.x;
        ay += p.y;
    }
    out[i] = make_float2(ax, ay);
}
__global__ void accel_from_force_kernel(const float2* __restrict__ force, const float* __restrict__ m, int count,
                                        float2* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float2 f = force[i];
    out[i] = make_float2(__fdiv_rn(f.x, m[i]), __fdiv_rn(f.y, m[i]));
}
void launch_accel_from_partial(Engine& e, const float2* partial, int nslices, float2* out) {
    const int count = local_count(e);
    if (count <= 0) return;
    accel_from_partial_kernel<<<(count + kBlk - 1) / kBlk, kBlk, 0, e.stream>>>(partial, nslices,
                                                                               static_cast<int>(e.lay.L), count, out);
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
}
void launch_accel_from_force(Engine& e, const float2* force, float2* out) {
    const int count = local_count(e);
    if (count <= 0) return;
    accel_from_force_kernel<<<(count + kBlk - 1) / kBlk, kBlk, 0, e.stream>>>(force, e.arena.m(e.lay), count, out);
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
}

__global__ void fill_zero_kernel(float* p, size_t n) {
    size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i < n) p[i] = 0.f;
}
void launch_fill_zero_f32(Engine& e, float* p, size_t n) {
    if (n == 0) return;
    fill_zero_kernel<<<static_cast<unsigned>((n + kBlk - 1) / kBlk), kBlk, 0, e.stream>>>(p, n);
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
}

// ---------------------------------------------------------------------------------------------
// Initial conditions on the device (rs-src/nbody.rs:39-104).  The reference draws from an unseeded
// thread_rng, so only the distributions are reproducible: a counter-based generator (splitmix64 of
// seed + global body index + stream) makes the result independent of the sharding.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float u01(uint64_t seed, uint64_t idx, uint32_t stream) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx * 8ull + stream + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
}

__global__ void gen_kernel(int kind, uint64_t seed, int begin, int count, int L, float n_total, float rmin, float rmax,
                           float* __restrict__ x0, float* __restrict__ y0, float* __restrict__ x1,
                           float* __restrict__ y1, float* __restrict__ m, float* __restrict__ vx,
                           float* __restrict__ vy) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= L) return;
    float px = 0.f, py = 0.f, qx = 0.f, qy = 0.f, mm = 0.f;
    if (i < count) {
        const uint64_t gi = static_cast<uint64_t>(begin + i);
        if (kind == 0) {
            // nb_random_disk: rs-src/nbody.rs:51-63 with uniform_sample_disk (:66-71)
            const float u = u01(seed, gi, 0), v = u01(seed, gi, 1);
            const float r = sqrtf(u), th = 2.0f * 3.14159265358979323846f * v;
            px = r * cosf(th) * 23.0f;
            py = r * sinf(th) * 23.0f;
            qx = -3.5f + 7.0f * u01(seed, gi, 2);
            qy = -3.5f + 7.0f * u01(seed, gi, 3);
            mm = 0.1f + 1.4f * u01(seed, gi, 4);
        } else if (kind == 2) {
            // nbx_plummer: Plummer model of scale a = rmin, equal masses rmax, truncated at 10 a, projected to z = 0
            // (configurations C2-C5; the same construction as rust_exp_b200/ic.py::plummer_2d).
            // Radius from the cumulative mass profile M(r)/M = r^3 / (r^2 + a^2)^(3/2); isotropic direction; velocities
            // Gaussian (Box-Muller) with the local dispersion sqrt(N m / 6) (1 + r^2/a^2)^(-1/4).
            const float a = rmin;
            const float xmax = 1000.0f / (101.0f * sqrtf(101.0f));
            const float u = fmaxf(u01(seed, gi, 0), 5.9604645e-8f) * xmax;
            const float r = a * rsqrtf(powf(u, -2.0f / 3.0f) - 1.0f);
            const float cz = 2.0f * u01(seed, gi, 1) - 1.0f;
            const float ph = 2.0f * 3.14159265358979323846f * u01(seed, gi, 2);
            const float sn = sqrtf(fmaxf(0.0f, 1.0f - cz * cz));
            px = r * sn * cosf(ph);
            py = r * sn * sinf(ph);
            mm = rmax;
            const float q = r / a;
            const float sigma = sqrtf(n_total * mm / 6.0f) * rsqrtf(sqrtf(1.0f + q * q));
            const float g = sqrtf(-2.0f * logf(fmaxf(u01(seed, gi, 3), 5.9604645e-8f)));
            const float ga = 2.0f * 3.14159265358979323846f * u01(seed, gi, 4);
            qx = sigma * g * cosf(ga);
            qy = sigma * g * sinf(ga);
        } else {
            // nb_stable_orbits: rs-src/nbody.rs:85-103
            if (gi == 0) {
                mm = 1000.0f;
            } else {
                const float speed = sqrtf(1.0f * 1000.0f);
                const float r = (rmax - rmin) * u01(seed, gi, 0) + rmin;
                const float th = 2.0f * 3.14159265358979323846f * u01(seed, gi, 1);
                px = r * cosf(th); py = r * sinf(th);
                qx = -speed * sinf(th); qy = speed * cosf(th);
                mm = 1.0f;
            }
        }
    }
    x0[i] = px; y0[i] = py; x1[i] = px; y1[i] = py;
    m[i] = mm; vx[i] = qx; vy[i] = qy;
}

static void launch_gen(Engine& e, int kind, float rmin, float rmax) {
    const int L = static_cast<int>(e.lay.L);
    if (L == 0) return;
    gen_kernel<<<(L + kBlk - 1) / kBlk, kBlk, 0, e.stream>>>(kind, e.seed, local_begin(e), local_count(e), L, static_cast<float>(e.n), rmin,
                                                            rmax, e.arena.x(e.lay, 0), e.arena.y(e.lay, 0),
                                                            e.arena.x(e.lay, 1), e.arena.y(e.lay, 1), e.arena.m(e.lay),
                                                            e.arena.vx(e.lay), e.arena.vy(e.lay));
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
    e.cur = 0;
}
void generate_random_disk(Engine& e, int n) { (void)n; launch_gen(e, 0, 0.f, 0.f); }
void generate_stable_orbits(Engine& e, int n, float rmin, float rmax) { (void)n; launch_gen(e, 1, rmin, rmax); }
void generate_plummer(Engine& e, int n, float a_scale, float mass_per_body) { (void)n; launch_gen(e, 2, a_scale, mass_per_body); }

}  // namespace nb
