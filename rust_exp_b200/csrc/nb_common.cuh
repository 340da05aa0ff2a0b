// nb_common.cuh -- shared device/host helpers for libnbody_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libnbody_b200 is written for sm_100a (B200) only"
#endif

namespace nb {

// rs-src/nbody.rs:13-17
constexpr float kEps = 0.0001f;
constexpr float kVpWdh = 100.0f;
constexpr float kKillLimit = 55.0f;  // VP_WDH * 0.55 (rs-src/nbody.rs:466-467); 100*0.55f rounds to 55.0f

// The nb_* ABI has no error channel (the reference panics): report and abort.
void fatal(const char* what, const char* file, int line);
void set_error(const char* fmt, ...);

#define NB_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e__ = (call);                                            \
        if (e__ != cudaSuccess) ::nb::fatal(cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk -> SASS UBLKCP) ------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> this CTA's shared memory; completion is signalled on `bar` as transaction bytes.
// size and both addresses must be multiples of 16 bytes.  src may be a peer-GPU (NVLink) address.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// One MUFU.RCP.  d2+eps is never denormal here (>= 1e-4), so ftz costs nothing and avoids the
// denormal fix-up sequence of the non-ftz form.
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// Spin until *flag >= want (wrap-safe).  Bounded by timeout_ns (0 = no bound): a dead peer must not hang the GPU
// for ever -- message + trap, which the host turns into abort().
__device__ __forceinline__ void wait_epoch(const uint32_t* flag, uint32_t want, unsigned long long timeout_ns, int who,
                                           const char* what) {
    const volatile uint32_t* f = flag;
    if (static_cast<int32_t>(*f - want) < 0) {
        const unsigned long long t0 = globaltimer_ns();
        while (static_cast<int32_t>(*f - want) < 0) {
            __nanosleep(100);
            if (timeout_ns != 0ull && globaltimer_ns() - t0 > timeout_ns) {
                printf("nbody_b200: timeout (%llu ms) waiting for rank %d to reach %s %u (at %u)\n", timeout_ns / 1000000ull, who,
                       what, want, *f);
                __trap();
            }
        }
    }
    __threadfence_system();
}

__device__ __forceinline__ float ld_volatile_f32(const float* p) { return *reinterpret_cast<const volatile float*>(p); }

#endif  // __CUDACC__

}  // namespace nb
