// Microbenchmark: per-SM throughput of cp.async.bulk (1-D TMA bulk copy) global->shared as a function of
// copy size, number of copies in flight and whether all CTAs read the same addresses.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t b) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(b) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// Each CTA streams `total` bytes in stages of `stage` bytes (split into `split` copies), `depth` stages in flight.
__global__ void probe(const uint8_t *src, size_t region, int shared_addr, uint32_t stage, int split, int depth, int n_stages, long long *clk) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[8];
  if (threadIdx.x == 0) {
    for (int i = 0; i < depth; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const uint8_t *base = src + (shared_addr ? 0 : (size_t)blockIdx.x * region);
  long long t0 = clock64();
  int issued = 0;
  for (int s = 0; s < n_stages + depth; ++s) {
    if (s >= depth) mbar_wait(smem_u32(&bars[(s - depth) % depth]), ((s - depth) / depth) & 1);
    if (issued < n_stages) {
      const int slot = issued % depth;
      mbar_expect_tx(smem_u32(&bars[slot]), stage);
      const size_t off = ((size_t)issued * stage) % region;
      for (int q = 0; q < split; ++q)
        bulk_g2s(smem_u32(smem) + slot * stage + q * (stage / split), base + off + q * (stage / split), stage / split, smem_u32(&bars[slot]));
      ++issued;
    }
  }
  clk[blockIdx.x] = clock64() - t0;
}
int main() {
  const size_t region = 8u << 20;      // 8 MB per CTA region (L2 resident overall: use shared or first touch)
  uint8_t *src; long long *clk;
  cudaMalloc(&src, region * 148); cudaMemset(src, 1, region * 148); cudaMalloc(&clk, 148 * sizeof(long long));
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  long long h[148];
  printf("%-8s %-7s %-6s %-6s %-10s %s\n", "shared", "stage", "split", "depth", "B/clk/SM", "TB/s@1.9GHz chip");
  for (int shared_addr = 0; shared_addr <= 1; ++shared_addr)
    for (uint32_t stage : {4096u, 16384u, 38912u})
      for (int split : {1, 4})
        for (int depth : {1, 2, 4}) {
          if ((size_t)stage * depth > 190 * 1024) continue;
          const int n_stages = (int)((4u << 20) / stage);
          for (int rep = 0; rep < 2; ++rep) {
            probe<<<148, 32, stage * depth>>>(src, shared_addr ? (1u << 20) : region, shared_addr, stage, split, depth, n_stages, clk);
            cudaDeviceSynchronize();
          }
          cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
          double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
          double bpc = (double)n_stages * stage / avg;
          printf("%-8d %-7u %-6d %-6d %-10.1f %.2f\n", shared_addr, stage, split, depth, bpc, bpc * 148 * 1.9e9 / 1e12);
        }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
