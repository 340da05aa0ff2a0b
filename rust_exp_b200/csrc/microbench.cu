// microbench.cu -- pipe-rate microbenchmarks that fix the roofline of the all-pairs kernel on B200.
//
// Prints one line per test: lane-ops per clock per SM (from clock64 inside the kernel) and per second
// chip-wide (from CUDA events).  Used to (a) validate the FP32 peak denominator (2*128*SMs*clk, SURVEY.md
// H9), (b) measure the MUFU.RCP rate that bounds the <2,REF> pair law, (c) show what the
// packed FP32x2 instructions buy.  Build: make -C rust_exp_b200/csrc ../nb_microbench
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int kIters = 4096;
constexpr int kUnroll = 8;

__device__ __forceinline__ float rcp_approx(float x) { float r; asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

enum Test { T_FFMA = 0, T_FFMA2, T_FADD, T_FADD2, T_FMUL2, T_MUFU, T_PAIR_SCALAR, T_PAIR_X2, T_MUFU_FFMA2_MIX, T_FFMA2_PLUS_FFMA, T_FFMA2_PLUS_FADD, T_FFMA2_PLUS_IMAD, T_COUNT };
static const char* kNames[T_COUNT] = {"ffma", "ffma2", "fadd", "fadd2", "fmul2", "mufu_rcp", "pair_scalar", "pair_x2", "mufu+3.5ffma2", "ffma2+ffma(1:1)", "ffma2+fadd(1:1)", "ffma2+imad(1:1)"};
// lane-ops (per thread per inner iteration) and flop per lane-op for reporting
static const double kOpsPerIter[T_COUNT] = {kUnroll, kUnroll, kUnroll, kUnroll, kUnroll, kUnroll, kUnroll, kUnroll, kUnroll, kUnroll, kUnroll, kUnroll};

template <int T>
__global__ void __launch_bounds__(256) bench_kernel(float* out, long long* cycles, float seed) {
    float a[kUnroll], b[kUnroll];
    float2 a2[kUnroll], b2[kUnroll];
    int ia[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; k++) {
        a[k] = seed + k + threadIdx.x * 1e-3f;
        b[k] = seed * 0.5f + k;
        a2[k] = make_float2(a[k], a[k] + 1.f);
        b2[k] = make_float2(b[k], b[k] + 1.f);
        ia[k] = threadIdx.x + k;
    }
    const float c = seed * 1.0001f, d = seed * 0.999f;
    const float2 c2 = make_float2(c, d), d2 = make_float2(d, c);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < kIters; it++) {
#pragma unroll
        for (int k = 0; k < kUnroll; k++) {
            if (T == T_FFMA) a[k] = fmaf(a[k], c, d);
            if (T == T_FFMA2) a2[k] = __ffma2_rn(a2[k], c2, d2);
            if (T == T_FADD) a[k] = a[k] + c;
            if (T == T_FADD2) a2[k] = __fadd2_rn(a2[k], c2);
            if (T == T_FMUL2) a2[k] = __fmul2_rn(a2[k], c2);
            if (T == T_MUFU) a[k] = rcp_approx(a[k]);
            if (T == T_PAIR_SCALAR) {
                // one pair, scalar: 2 FADD + 2 FFMA + MUFU + FMUL + 2 FFMA
                const float dx = b[k] - c, dy = b[(k + 1) % kUnroll] - d;
                const float r2 = fmaf(dy, dy, fmaf(dx, dx, 1e-4f));
                const float s = seed * rcp_approx(r2);
                a[k] = fmaf(s, dx, a[k]);
                b[k] = fmaf(s, dy, b[k]);
            }
            if (T == T_PAIR_X2) {
                // two pairs, packed: 2 FADD2 + 2 FFMA2 + 2 MUFU + FMUL2 + 2 FFMA2
                const float2 dx = __fadd2_rn(b2[k], c2), dy = __fadd2_rn(b2[(k + 1) % kUnroll], d2);
                float2 r2 = __ffma2_rn(dx, dx, make_float2(1e-4f, 1e-4f));
                r2 = __ffma2_rn(dy, dy, r2);
                const float2 s = __fmul2_rn(c2, make_float2(rcp_approx(r2.x), rcp_approx(r2.y)));
                a2[k] = __ffma2_rn(s, dx, a2[k]);
                b2[k] = __ffma2_rn(s, dy, b2[k]);
            }
            if (T == T_FFMA2_PLUS_FFMA) { a2[k] = __ffma2_rn(a2[k], c2, d2); a[k] = fmaf(a[k], c, d); }
            if (T == T_FFMA2_PLUS_FADD) { a2[k] = __ffma2_rn(a2[k], c2, d2); a[k] = a[k] + c; }
            if (T == T_FFMA2_PLUS_IMAD) { a2[k] = __ffma2_rn(a2[k], c2, d2); ia[k] = ia[k] * 3 + it; }
            if (T == T_MUFU_FFMA2_MIX) {
                a[k] = rcp_approx(a[k]);
                a2[k] = __ffma2_rn(a2[k], c2, d2);
                b2[k] = __ffma2_rn(b2[k], c2, d2);
                if (k & 1) a2[k] = __ffma2_rn(a2[k], d2, c2);
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < kUnroll; k++) acc += a[k] + b[k] + a2[k].x + a2[k].y + b2[k].x + b2[k].y + (float)ia[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int T>
static void run(int sms, int blocks_per_sm, float* out, long long* cyc_d) {
    const int grid = sms * blocks_per_sm;
    bench_kernel<T><<<grid, 256>>>(out, cyc_d, 1.25f);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int r = 0; r < reps; r++) bench_kernel<T><<<grid, 256>>>(out, cyc_d, 1.25f);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    long long* cyc = (long long*)malloc(sizeof(long long) * grid);
    CK(cudaMemcpy(cyc, cyc_d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    double cmax = 0;
    for (int i = 0; i < grid; i++) if (cyc[i] > cmax) cmax = (double)cyc[i];
    free(cyc);
    const double thread_iters = (double)kIters * kOpsPerIter[T];          // "units" per thread
    const double units_per_sm = thread_iters * 256.0 * blocks_per_sm;    // per kernel
    const double per_clk_sm = units_per_sm / cmax;
    const double per_s = units_per_sm * sms * reps / (ms * 1e-3);
    printf("{\"test\": \"%s\", \"blocks_per_sm\": %d, \"units_per_clk_per_sm\": %.2f, \"units_per_s\": %.4e, \"ms_per_launch\": %.4f, \"eff_clock_ghz\": %.3f}\n",
           kNames[T], blocks_per_sm, per_clk_sm, per_s, ms / reps, cmax / (ms / reps * 1e-3) / 1e9);
}

int main() {
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, dev));
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, p.multiProcessorCount, p.clockRate);
    const int sms = p.multiProcessorCount;
    float* out;
    long long* cyc;
    CK(cudaMalloc(&out, sizeof(float) * sms * 8 * 256));
    CK(cudaMalloc(&cyc, sizeof(long long) * sms * 8));
    // "units": ffma/fadd = lane-instructions (x2 forms = one unit per packed instruction, i.e. 2 lane-results);
    // mufu = lane-results; pair_scalar = pairs; pair_x2 = 2 pairs per unit.
    for (int bps : {4}) {
        run<T_FFMA>(sms, bps, out, cyc);
        run<T_FFMA2>(sms, bps, out, cyc);
        run<T_FADD>(sms, bps, out, cyc);
        run<T_FADD2>(sms, bps, out, cyc);
        run<T_FMUL2>(sms, bps, out, cyc);
        run<T_MUFU>(sms, bps, out, cyc);
        run<T_PAIR_SCALAR>(sms, bps, out, cyc);
        run<T_PAIR_X2>(sms, bps, out, cyc);
        run<T_MUFU_FFMA2_MIX>(sms, bps, out, cyc);
        run<T_FFMA2_PLUS_FFMA>(sms, bps, out, cyc);
        run<T_FFMA2_PLUS_FADD>(sms, bps, out, cyc);
        run<T_FFMA2_PLUS_IMAD>(sms, bps, out, cyc);
    }
    return 0;
}
