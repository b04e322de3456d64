// Microbenchmark: two warps of one CTA issue tcgen05.mma M128 N64 K16 (A in TMEM) jobs concurrently into different
// accumulators.  Does the interleaving cost tensor-pipe throughput?  Prints clk per MMA (aggregate), nominal 32.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{.reg .pred P; elect.sync _|P, 0xffffffff; selp.u32 %0, 1, 0, P;}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
// n_issuers warps; each does `iters` jobs of J MMAs into accumulator (warp * 2 + (job & 1)) * 64; nwaits dummy waits per job
template <int J>
__global__ void __launch_bounds__(128, 1) probe(int n_issuers, int iters, int nwaits, long long *clk) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar[4], bar3;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar3)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (warp < n_issuers) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t b_desc0 = make_desc(smem_u32(smem) + warp * 40960, 1024, 128);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      for (int w = 0; w < nwaits; ++w) mbar_wait(smem_u32(&bar3), 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < J; ++k)
          asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}"
                       ::"r"(tmem + 64 * (warp * 2 + (i & 1))), "r"(tmem + 256 + 128 * warp + 8 * (k % 10)),
                         "l"(b_desc0 + (uint64_t)(((k % 5) * 2048) >> 4)), "r"(idesc), "r"(k ? 1u : 0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[2 + warp])) : "memory");
      }
      __syncwarp();
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[warp])) : "memory");
    __syncwarp();
    mbar_wait(smem_u32(&bar[warp]), 0);
    if ((threadIdx.x & 31) == 0) clk[blockIdx.x * 2 + warp] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}
int main() {
  long long *clk; cudaMalloc(&clk, 148 * 2 * 8);
  cudaFuncSetAttribute(probe<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 2000;
  for (int ni = 1; ni <= 2; ++ni)
    for (int nw : {0, 1, 2, 3}) {
      for (int rep = 0; rep < 2; ++rep) { probe<15><<<148, 128, 96 * 1024>>>(ni, iters, nw, clk); cudaDeviceSynchronize(); }
      long long h[296]; cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
      double mx = 0; for (int w = 0; w < ni; ++w) mx = h[w] > mx ? h[w] : mx;
      printf("issuers %d, waits/job %d: clk per MMA (aggregate) %.1f, clk per job (aggregate) %.0f  err=%s\n", ni, nw, mx / (iters * 15.0 * ni), mx / (iters * ni),
             cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
