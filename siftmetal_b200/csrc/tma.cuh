// tma.cuh — TMA (cp.async.bulk.tensor) and mbarrier helpers shared by the blur and match kernels,
// plus the host-side tensor-map encoder (driver entry point fetched through the runtime: the
// library does not link libcuda).
#pragma once
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace sift {

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
// makes the initialised barriers visible to the async proxy (TMA / tcgen05.commit arrivals)
__device__ __forceinline__ void mbarInitFence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smemAddr(bar);
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmaLoad2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smemAddr(dst)), "l"((uint64_t)map), "r"(smemAddr(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmaLoad3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smemAddr(dst)), "l"((uint64_t)map), "r"(smemAddr(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
#endif

inline PFN_cuTensorMapEncodeTiled_v12000 tensorMapEncoder() {
    static std::atomic<void*> cached{nullptr};
    void* fn = cached.load(std::memory_order_acquire);
    if (!fn) {
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        cached.store(fn, std::memory_order_release);
    }
    return (PFN_cuTensorMapEncodeTiled_v12000)fn;
}

}  // namespace sift
