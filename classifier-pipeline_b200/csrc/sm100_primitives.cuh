// sm100_primitives.cuh -- thin wrappers over the sm_100a synchronisation / bulk-copy PTX the kernels use:
// shared-memory mbarriers (init / arrive / expect_tx / try_wait with a suspend-time hint) and 1-D TMA bulk copies
// (cp.async.bulk global -> shared, completion counted on an mbarrier) with L2 eviction policies.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cpt {
namespace prim {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_inval(unsigned long long *bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// A waiting warp sleeps in hardware for up to hint_ns instead of re-issuing try_wait (a spinning warp takes issue slots
// from the warps it is waiting for).
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity, uint32_t hint_ns = 4000u) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PRIM_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra.uni PRIM_WAIT_DONE;\n"
        "bra.uni PRIM_WAIT_LOOP;\n"
        "PRIM_WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(hint_ns) : "memory");
}

// the same on a shared-space address (loops that step through a ring of barriers keep the address, not the pointer)
__device__ __forceinline__ void mbar_arrive_addr(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity, uint32_t hint_ns = 4000u) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "PRIM_WAITA_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra.uni PRIM_WAITA_DONE;\n"
        "bra.uni PRIM_WAITA_LOOP;\n"
        "PRIM_WAITA_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity), "r"(hint_ns) : "memory");
}

__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// global -> shared bulk copy (TMA, 1-D): bytes a multiple of 16, both addresses 16-byte aligned; completion is counted
// on the mbarrier.
__device__ __forceinline__ void bulk_g2s_hint(void *dst, const void *src, uint32_t bytes, unsigned long long *bar,
                                              unsigned long long policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

}  // namespace prim
}  // namespace cpt
