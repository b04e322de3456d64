// Microbenchmark: issue rate of tcgen05.mma kind::f16 (M=128) for N = 64/128/256, A from TMEM or from shared memory.
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
template <int N, bool TA>
__global__ void __launch_bounds__(128, 1) probe(int iters, long long *clk) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;  // fp16 1.0
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
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t a_desc = make_desc(smem_u32(smem), 2048, 128);
    const uint64_t b_desc = make_desc(smem_u32(smem) + 8192, N * 16, 128);
    long long t0 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (TA)
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}"
                         ::"r"(tmem), "r"(tmem + 256 + 8 * k), "l"(b_desc), "r"(idesc), "r"(1u) : "memory");
          else
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}"
                         ::"r"(tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(1u) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    __syncwarp();
    uint32_t done;
    do {
      asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    } while (!done);
    if (threadIdx.x == 0 || t0) { if (t0) clk[blockIdx.x] = clock64() - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}
// jobs of J MMAs; after each job: commit to `bar2` (never waited) and try_wait once on an already-completed barrier
template <int J>
__global__ void __launch_bounds__(128, 1) probe_jobs(int jobs, int nwaits, long long *clk) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint32_t tmem_slot;
  __shared__ __align__(8) uint64_t bar, bar2, bar3;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2)));
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
  if (warp == 0) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
    const uint64_t b_desc = make_desc(smem_u32(smem) + 8192, 64 * 16, 128);
    long long t0 = clock64();
    for (int j = 0; j < jobs; ++j) {
      for (int w = 0; w < nwaits; ++w) {
        uint32_t done;   // bar3 is fresh: waiting for parity 1 returns immediately (already-complete fast path)
        do {
          asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&bar3)), "r"(1u) : "memory");
        } while (!done);
      }
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < J; ++k)
          asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}"
                       ::"r"(tmem + 64 * (j % 3)), "r"(tmem + 256 + 8 * (k % 10)), "l"(b_desc), "r"(idesc), "r"(k ? 1u : 0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
      }
      __syncwarp();
    }
    if (elect_one())
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    __syncwarp();
    uint32_t done;
    do {
      asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    } while (!done);
    if (threadIdx.x == 0) clk[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}
template <int J>
void run_jobs(long long *clk, int nwaits) {
  const int jobs = 2000;
  cudaFuncSetAttribute(probe_jobs<J>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  long long h[148];
  for (int rep = 0; rep < 2; ++rep) { probe_jobs<J><<<148, 128, 48 * 1024>>>(jobs, nwaits, clk); cudaDeviceSynchronize(); }
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
  printf("jobs of %2d MMAs (N=64, A tmem), %d waits/job: clk/MMA %.1f  clk/job %.0f  err=%s\n", J, nwaits, avg / (jobs * (double)J), avg / jobs, cudaGetErrorString(cudaGetLastError()));
}
template <int N, bool TA>
void run(long long *clk, const char *name) {
  const int iters = 2000;
  cudaFuncSetAttribute(probe<N, TA>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  long long h[148];
  for (int rep = 0; rep < 2; ++rep) {
    probe<N, TA><<<148, 128, 48 * 1024>>>(iters, clk);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  printf("%-22s N=%-4d clk/MMA %.1f   (nominal %d)  err=%s\n", name, N, avg / (iters * 8.0), N / 2, cudaGetErrorString(cudaGetLastError()));
}
int main() {
  long long *clk;
  cudaMalloc(&clk, 148 * sizeof(long long));
  run<64, false>(clk, "A smem (SS)");
  run<128, false>(clk, "A smem (SS)");
  run<256, false>(clk, "A smem (SS)");
  run<64, true>(clk, "A tmem (TS)");
  run<128, true>(clk, "A tmem (TS)");
  run<256, true>(clk, "A tmem (TS)");
  run_jobs<15>(clk, 0); run_jobs<15>(clk, 1); run_jobs<15>(clk, 2); run_jobs<15>(clk, 4);
  run_jobs<30>(clk, 0); run_jobs<30>(clk, 2);
  return 0;
}
