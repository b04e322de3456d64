// Microbenchmark: does the tcgen05.mma M128 N64 K16 (A in TMEM, B in shared memory) rate depend on where B comes from?
//   bmode 0: the same B descriptor for every MMA            bmode 1: B walks through a 40 KB slot like the GMM kernel
//   tma   0: shared memory otherwise idle                   tma   1: a second warp streams 40 KB bulk copies into other slots
//   lbo: byte distance between K-adjacent core matrices (1024 = GMM kernel's W layout; 128 = K-contiguous alternative)
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
template <int N>
__global__ void __launch_bounds__(96, 1) probe(int bmode, int tma, int iters, uint32_t lbo, uint32_t sbo, const uint8_t *gsrc, long long *clk, int *copies) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar, bar_t;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    stop = 0;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar_t)));
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
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t b_desc0 = make_desc(smem_u32(smem), lbo, sbo);
    long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < iters; ++i) {
        // slot = i % 2 (slots 0,1 are MMA-read; slots 2..4 receive the bulk copies)
        const uint64_t b_slot = b_desc0 + (uint64_t)(((i & 1) * 40960u) >> 4);
#pragma unroll
        for (int k = 0; k < 15; ++k) {
          // 5 k-blocks x 3 parts; part 2 reads the second half of the slot
          const uint32_t off = bmode ? ((k % 5) * 2 * (N * 16) + (k >= 10 ? 20480u : 0u)) : 0u;
          asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}"
                       ::"r"(tmem + 64 * (i % 3)), "r"(tmem + 256 + 8 * (k % 10)), "l"((bmode ? b_slot : b_desc0) + (uint64_t)(off >> 4)), "r"(idesc), "r"(k ? 1u : 0u) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0);
    if (threadIdx.x == 0) { clk[blockIdx.x] = clock64() - t0; stop = 1; }
  } else if (warp == 1 && tma) {
    int n = 0;
    if (elect_one()) {
      while (!stop) {
        const uint32_t dst = smem_u32(smem) + (2 + n % 3) * 40960u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar_t)), "r"(40960u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(dst), "l"(gsrc + (size_t)((n * 7 + blockIdx.x) % 100) * 40960), "r"(40960u), "r"(smem_u32(&bar_t)) : "memory");
        mbar_wait(smem_u32(&bar_t), n & 1);
        ++n;
      }
      copies[blockIdx.x] = n;
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}
template <int N>
void run(int bmode, int tma, uint32_t lbo, uint32_t sbo, const uint8_t *g, long long *clk, int *copies) {
  const int iters = 3000;
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int rep = 0; rep < 2; ++rep) { probe<N><<<148, 96, 201 * 1024>>>(bmode, tma, iters, lbo, sbo, g, clk, copies); cudaDeviceSynchronize(); }
  long long h[148]; int c[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(c, copies, sizeof(c), cudaMemcpyDeviceToHost);
  double avg = 0, ac = 0; for (int i = 0; i < 148; ++i) { avg += h[i]; ac += c[i]; } avg /= 148; ac /= 148;
  printf("N=%3d bmode %d tma %d lbo %4u sbo %4u: clk/MMA %.1f (nominal %d)", N, bmode, tma, lbo, sbo, avg / (iters * 15.0), N / 2);
  if (tma) printf("   bulk copies/SM %.0f -> %.1f B/clk/SM, %.0f clk per 40 KB copy", ac, ac * 40960 / avg, avg / ac);
  printf("  err=%s\n", cudaGetErrorString(cudaGetLastError()));
}
int main() {
  long long *clk; int *copies; uint8_t *g;
  cudaMalloc(&clk, 148 * 8); cudaMalloc(&copies, 148 * 4); cudaMalloc(&g, 100 * 40960); cudaMemset(g, 0, 100 * 40960); cudaMemset(copies, 0, 148 * 4);
  for (int tma = 0; tma < 2; ++tma) {
    run<64>(0, tma, 1024, 128, g, clk, copies);
    run<64>(1, tma, 1024, 128, g, clk, copies);
    run<64>(1, tma, 128, 256, g, clk, copies);     // K-contiguous core matrices: [n-group][k-slab][8 rows][16 B]
    run<128>(1, tma, 2048, 128, g, clk, copies);
    run<128>(1, tma, 128, 256, g, clk, copies);
    run<256>(1, tma, 4096, 128, g, clk, copies);
  }
  return 0;
}
