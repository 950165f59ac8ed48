// Shared helpers for the hm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/hm_b200.h"

void hm_set_error(const char* fmt, ...);

#define HM_CUDA_TRY(expr)                                                                   \
    do {                                                                                    \
        cudaError_t _e = (expr);                                                            \
        if (_e != cudaSuccess) {                                                            \
            hm_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                         __LINE__);                                                         \
            return (int)_e;                                                                 \
        }                                                                                   \
    } while (0)

#define HM_REQUIRE(cond, msg)                                           \
    do {                                                                \
        if (!(cond)) {                                                  \
            hm_set_error("invalid argument: %s (%s)", msg, #cond);      \
            return HM_E_INVALID;                                        \
        }                                                               \
    } while (0)

int hm_num_sms();

// ---- checked build (make checked -> libhm_b200_checked.so, -DHM_CHECKED) ---------------------------------------------
// compute-sanitizer is not available on the GPU pool this library is developed on (profiles/r02_sanitizer_unavailable.txt),
// so the hand-rolled synchronisation code carries its own invariants: every index into a shared-memory ring / queue /
// candidate buffer and every staged-node address is checked in the checked build; a violation is COUNTED (the kernel
// keeps running, nothing traps, so a failing run cannot wedge the GPU) and hm_checked_failures() reports the total and
// the source line of the last one.  The normal build compiles the checks away.
#ifdef HM_CHECKED
static __device__ unsigned hm_chk_fails = 0;
static __device__ int hm_chk_line = 0;
#define HM_DEV_ASSERT(cond)                     \
    do {                                        \
        if (!(cond)) {                          \
            atomicAdd(&hm_chk_fails, 1u);       \
            hm_chk_line = __LINE__;             \
        }                                       \
    } while (0)
#define HM_CHECKED_EXPORT(tu)                                                        \
    extern "C" unsigned hm_chk_##tu(int* line) {                                     \
        unsigned v = 0;                                                              \
        cudaDeviceSynchronize();                                                     \
        cudaMemcpyFromSymbol(&v, hm_chk_fails, sizeof(v));                           \
        if (line) cudaMemcpyFromSymbol(line, hm_chk_line, sizeof(int));              \
        return v;                                                                    \
    }
#else
#define HM_DEV_ASSERT(cond) ((void)0)
#define HM_CHECKED_EXPORT(tu)
#endif

// ---- order-preserving key for the stable (value, index) selection ---------------------------
// np.argmin semantics: NaN is "smallest" (first NaN wins), -0.0 == +0.0, ties -> lowest index.
__host__ __device__ __forceinline__ unsigned long long hm_order_key(double v) {
    if (v != v) return 0ull;                   // NaN first
    v = v + 0.0;                               // -0.0 -> +0.0
#ifdef __CUDA_ARCH__
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
#else
    unsigned long long b;
    memcpy(&b, &v, 8);
#endif
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

struct HmPair {
    unsigned long long key;   // hm_order_key(value)
    long long idx;            // global index
};
__host__ __device__ __forceinline__ bool hm_pair_less(const HmPair& a, const HmPair& b) {
    return (a.key < b.key) || (a.key == b.key && a.idx < b.idx);
}

// ---- mbarrier / bulk-copy (TMA 1-D) wrappers --------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t hm_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void hm_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hm_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void hm_fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void hm_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hm_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void hm_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(hm_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk async copy (UBLKCP), completion signalled on an mbarrier; 16 B aligned, size % 16 == 0
__device__ __forceinline__ void hm_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            hm_smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(hm_smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void hm_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
#endif
