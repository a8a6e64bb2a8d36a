// skrample_b200 - small device/host helpers shared by the kernels (sm_100a).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>

#include "../../include/skrample_b200.h"

namespace skr {

__host__ __device__ __forceinline__ uint32_t dtype_size(int dtype) {
    return dtype == SKR_F64 ? 8u : (dtype == SKR_F32 ? 4u : 2u);
}
static inline uint32_t dtype_size_host(int dtype) { return dtype_size(dtype); }

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk), the sm_90+/sm_100 async-proxy path -------------

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void fence_barrier_init() {
    // make the initialised barriers visible to the async proxy (TMA unit)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t tx_bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(tx_bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(addr),
        "r"(parity)
        : "memory");
}

// global -> shared bulk copy of `bytes` (multiple of 16, both sides 16-byte aligned); completion is
// signalled on `bar` as transaction bytes.
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}


// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may
// become resident while its predecessor in the stream is still running.  launch_dependents lets the successor of
// THIS grid start its prologue; wait blocks until every predecessor grid has completed and its memory is visible.
// Both are no-ops for a kernel that was launched without the attribute.
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

}  // namespace skr
