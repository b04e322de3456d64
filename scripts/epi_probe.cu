// Microbenchmark: tcgen05.ld 32x32b.x32 throughput with 4/8/16 reader warps, alone and against a concurrent stream of
// tcgen05.mma M128 N64 K16 (A in TMEM).  Prints clk per LDTM instruction per warp and the MMA rate.
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
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t *r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
// warps 0: MMA issuer (if mma_jobs > 0); warps 1..: readers
__global__ void __launch_bounds__(544, 1) probe(int n_readers, int reps, int mma_iters, long long *clk_ld, long long *clk_mma, uint32_t *sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
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
  if (warp == 0) {
    if (mma_iters > 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
      const uint64_t b_desc = make_desc(smem_u32(smem) + 8192, 64 * 16, 128);
      long long t0 = clock64();
      if (elect_one()) {
        for (int i = 0; i < mma_iters; ++i) {
#pragma unroll
          for (int k = 0; k < 15; ++k)
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}"
                         ::"r"(tmem + 64 * (i % 3)), "r"(tmem + 256 + 8 * (k % 10)), "l"(b_desc), "r"(idesc), "r"(k ? 1u : 0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
      }
      __syncwarp();
      uint32_t done;
      do {
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      } while (!done);
      if (lane == 0) clk_mma[blockIdx.x] = clock64() - t0;
    }
  } else if (warp <= n_readers) {
    const int quad = warp & 3;
    const uint32_t taddr = tmem + ((uint32_t)(quad * 32) << 16);
    uint32_t acc = 0;
    uint32_t v[32], w[32];
    long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      tc_ld32(taddr + 64 * (i % 3), v);
      tc_ld32(taddr + 64 * (i % 3) + 32, w);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int k = 0; k < 32; ++k) acc ^= v[k] + w[k];
    }
    long long t1 = clock64();
    if (lane == 0) clk_ld[blockIdx.x * 32 + warp] = t1 - t0;
    if (acc == 0x12345u) sink[0] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}
int main() {
  long long *clk_ld, *clk_mma; uint32_t *sink;
  cudaMalloc(&clk_ld, 148 * 32 * 8); cudaMalloc(&clk_mma, 148 * 8); cudaMalloc(&sink, 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 2000;
  for (int mma = 0; mma < 2; ++mma)
    for (int nr : {4, 8, 16}) {
      // mma iterations sized to cover the readers: each reader does reps pairs of LDTM
      const int mma_iters = mma ? 4000 : 0;
      for (int rep = 0; rep < 2; ++rep) { probe<<<148, 32 * (nr + 1), 48 * 1024>>>(nr, reps, mma_iters, clk_ld, clk_mma, sink); cudaDeviceSynchronize(); }
      long long h[148 * 32], hm[148];
      cudaMemcpy(h, clk_ld, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hm, clk_mma, sizeof(hm), cudaMemcpyDeviceToHost);
      double avg = 0; for (int w = 1; w <= nr; ++w) avg += h[w]; avg /= nr;
      // per SM bytes per clk: nr warps x reps x 2 x 4 KB over avg clk
      printf("readers %2d  mma %d: clk per LDTM.x32 per warp %.1f  -> TMEM read %.0f B/clk/SM", nr, mma, avg / (2.0 * reps), nr * reps * 2.0 * 4096 / avg);
      if (mma) printf("   MMA clk/MMA %.1f (nominal 32)", hm[0] / (mma_iters * 15.0));
      printf("  err=%s\n", cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
