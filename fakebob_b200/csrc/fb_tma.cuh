// mbarrier + bulk-copy (TMA engine, non-tensor form) wrappers for the streaming kernels of the i-vector path: one elected
// thread of a producer warp issues cp.async.bulk copies global -> shared that complete on a "full" mbarrier, consumer warps
// release a slot by arriving on its "empty" mbarrier.  (fb_gmm.cu carries its own copies next to its tcgen05 wrappers.)
#pragma once
#include <cstdint>

__device__ __forceinline__ uint32_t tma_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_bar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void tma_bar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_bar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_bar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
// Orders this thread's view of earlier generic-proxy accesses to shared memory (the consumers' loads, made visible to it by
// the "empty" mbarrier) before async-proxy writes it issues afterwards (the bulk copy that refills the slot).
__device__ __forceinline__ void tma_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// bytes: multiple of 16; dst and src 16-byte aligned
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
