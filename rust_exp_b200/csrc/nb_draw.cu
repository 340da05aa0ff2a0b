// nb_draw.cu -- nb_draw on the device (rs-src/nbody.rs:482-617).
//
// The reference read-modify-writes the caller's (write-combined, PBO-mapped) framebuffer once per
// particle.  Here the image is composed in HBM: per-channel saturating add is commutative and
// associative (min(255, a+b) folds in any order to min(255, sum)), so a CAS loop per touched pixel gives
// the identical image regardless of the order in which particles land, and only w*h*4 bytes cross
// PCIe -- not 20 bytes per particle.
#include "nb_engine.h"

namespace nb {

// rs-src/nbody.rs:595-617
__device__ __forceinline__ uint32_t add_abgr32(uint32_t c1, uint32_t c2) {
    const uint32_t a = min(255u, (c1 >> 24) + (c2 >> 24));
    const uint32_t b = min(255u, ((c1 >> 16) & 255u) + ((c2 >> 16) & 255u));
    const uint32_t g = min(255u, ((c1 >> 8) & 255u) + ((c2 >> 8) & 255u));
    const uint32_t r = min(255u, (c1 & 255u) + (c2 & 255u));
    return (a << 24) | (b << 16) | (g << 8) | r;
}

__device__ __forceinline__ void blend(uint32_t* px, uint32_t col) {
    uint32_t old = *px, assumed;
    do {
        assumed = old;
        const uint32_t nw = add_abgr32(assumed, col);
        if (nw == assumed) return;  // saturated: nothing to write
        old = atomicCAS(px, assumed, nw);
    } while (old != assumed);
}

// Rust `as i32` on f32: saturating, NaN -> 0 -- exactly what cvt.rzi.s32.f32 does.
__device__ __forceinline__ int f32_as_i32(float v) { return __float2int_rz(v); }

struct DrawParams {
    int w, h;
    float x1, y1, scalex, scaley;
    uint32_t col_body, col_tail;
};

__global__ void draw_scatter_kernel(DrawParams d, int count, const float* __restrict__ x, const float* __restrict__ y,
                                    const float* __restrict__ vx, const float* __restrict__ vy, uint32_t* fb) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    // rs-src/nbody.rs:525-526
    const float fx = __fmul_rn(__fsub_rn(x[i], d.x1), d.scalex);
    const float fy = __fmul_rn(__fsub_rn(y[i], d.y1), d.scaley);
    const int xb = f32_as_i32(fx), yb = f32_as_i32(fy);
    if (xb >= 0 && xb < d.w && yb >= 0 && yb < d.h) blend(fb + xb + yb * d.w, d.col_body);
    // tail: rs-src/nbody.rs:541-555.  atan2 is evaluated in f64 and rounded once, which reproduces a
    // correctly rounded atan2f; the rest is f32 in the reference's order.
    const float angle = static_cast<float>(atan2(static_cast<double>(vy[i]), static_cast<double>(vx[i])));
    const float t = __fadd_rn(__fdiv_rn(__fmul_rn(8.0f, angle), __fmul_rn(2.0f, 3.14159265358979323846f)), 8.0f);
    int oct = f32_as_i32(t) % 8;
    if (oct < 0) oct += 8;
    const int dxs[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    const int dys[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    const int xt = xb - dxs[oct], yt = yb - dys[oct];
    if (xt >= 0 && xt < d.w && yt >= 0 && yt < d.h) blend(fb + xt + yt * d.w, d.col_tail);
}

__global__ void draw_cross_kernel(int w, int h, uint32_t* fb) {
    // rs-src/nbody.rs:571-577
    // (the reference writes these five pixels unchecked; out-of-range ones are dropped here instead of
    // corrupting memory for degenerate w,h < 3)
    const int ox[5] = {0, 1, 0, -1, 0}, oy[5] = {0, 0, 1, 0, -1};
    const int k = threadIdx.x;
    if (blockIdx.x == 0 && k < 5) {
        const int xx = w / 2 + ox[k], yy = h / 2 + oy[k];
        if (xx >= 0 && xx < w && yy >= 0 && yy < h) fb[xx + yy * w] = 0x00FF00FFu;
    }
}

// rs-src/nbody.rs:585-593 (host; Rust `as u32` saturates)
static uint32_t rgb_to_abgr32(unsigned r8, unsigned g8, unsigned b8, float factor) {
    auto cv = [&](unsigned c) -> uint32_t {
        float v = static_cast<float>(c) * factor;
        if (!(v > 0.f)) return 0u;
        uint32_t u = v >= 4294967296.0f ? 0xFFFFFFFFu : static_cast<uint32_t>(v);
        return u > 255u ? 255u : u;
    };
    return (cv(r8) << 0) | (cv(b8) << 16) | (cv(g8) << 8);
}

static uint32_t* g_fb_dev = nullptr;
static size_t g_fb_cap = 0;

void draw_shutdown() {
    if (g_fb_dev) cudaFree(g_fb_dev);
    g_fb_dev = nullptr;
    g_fb_cap = 0;
}

void draw_to_host(Engine& e, int w, int h, uint32_t* fb) {
    if (w <= 0 || h <= 0) return;
    const size_t px = static_cast<size_t>(w) * static_cast<size_t>(h);
    if (px > g_fb_cap) {
        if (g_fb_dev) NB_CUDA(cudaFree(g_fb_dev));
        NB_CUDA(cudaMalloc(&g_fb_dev, px * sizeof(uint32_t)));
        g_fb_cap = px;
    }
    NB_CUDA(cudaMemsetAsync(g_fb_dev, 0, px * sizeof(uint32_t), e.stream));  // rs-src/nbody.rs:490
    // rs-src/nbody.rs:494-506, evaluated in f32 exactly as written there
    DrawParams d{};
    d.w = w; d.h = h;
    const float aspect = static_cast<float>(h) / static_cast<float>(w);
    const float x1 = 0.0f - kVpWdh / 2.0f;
    const float y1 = (0.0f - kVpWdh / 2.0f) * aspect;
    const float x2 = 0.0f + kVpWdh / 2.0f;
    const float y2 = (0.0f + kVpWdh / 2.0f) * aspect;
    const float vpw = x2 - x1, vph = y2 - y1;
    d.x1 = x1; d.y1 = y1;
    d.scalex = (1.0f / vpw) * static_cast<float>(w);
    d.scaley = (1.0f / vph) * static_cast<float>(h);
    d.col_body = rgb_to_abgr32(255, 215, 130, 0.3f);
    d.col_tail = rgb_to_abgr32(255, 215, 130, 0.25f);
    if (e.n > 0) {
        if (e.dist && e.world > 1) dist_wait_all(e, e.step_count);
        const int L = static_cast<int>(e.lay.L);
        for (int g = 0; g < e.world; g++) {
            int c = e.n - g * L;
            if (c > L) c = L;
            if (c <= 0) break;
            const ArenaView& av = (g == e.rank) ? e.arena : e.peer[g];
            draw_scatter_kernel<<<(c + 255) / 256, 256, 0, e.stream>>>(d, c, av.x(e.lay, e.cur), av.y(e.lay, e.cur),
                                                                       av.vx(e.lay), av.vy(e.lay), g_fb_dev);
            NB_CUDA(cudaGetLastError());
            e.ctr.kernel_launches++;
        }
    }
    draw_cross_kernel<<<1, 32, 0, e.stream>>>(w, h, g_fb_dev);
    NB_CUDA(cudaGetLastError());
    e.ctr.kernel_launches++;
    NB_CUDA(cudaMemcpyAsync(fb, g_fb_dev, px * sizeof(uint32_t), cudaMemcpyDeviceToHost, e.stream));
    NB_CUDA(cudaStreamSynchronize(e.stream));
    if (e.dist && e.world > 1) {
        e.step_count++;
        dist_signal_step_done(e);
    }
}

}  // namespace nb
