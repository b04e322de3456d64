// Microbenchmark: sustained DFMA rate of the whole GPU (what bounds ivec_quad / ivec_solve), and DMMA m8n8k4 for comparison.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/fp64_probe scripts/fp64_probe.cu && scripts/fp64_probe
#include <cuda_runtime.h>
#include <stdio.h>
template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_probe(int iters, double *out, double a, double b) {
  double acc[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) dmma_probe(int iters, double *out, double a, double b) {
  double c0[4], c1[4];
  for (int i = 0; i < 4; ++i) { c0[i] = threadIdx.x; c1[i] = i; }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
  }
  double s = 0;
  for (int i = 0; i < 4; ++i) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// dependent-chain latencies (one warp): DFMA, rsqrt(double), shuffle of a double
__global__ void latency_probe(int iters, double *out, long long *clk, double a, double b) {
  double x = threadIdx.x + 1.5;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) x = fma(x, a, b);
  long long t1 = clock64();
  double y = x;
  for (int i = 0; i < iters; ++i) y = rsqrt(y) + 1.0;
  long long t2 = clock64();
  double z = y;
  for (int i = 0; i < iters; ++i) z = __shfl_sync(0xffffffffu, z, (i + 1) & 31) + 1.0;
  long long t3 = clock64();
  float f = (float)z;
  for (int i = 0; i < iters; ++i) f = fmaf(f, 1.0001f, 0.5f);
  long long t4 = clock64();
  out[threadIdx.x] = z + f;
  if (threadIdx.x == 0) { clk[0] = t1 - t0; clk[1] = t2 - t1; clk[2] = t3 - t2; clk[3] = t4 - t3; }
}
int main() {
  {
    double *o; long long *c, h[4];
    cudaMalloc(&o, 32 * 8); cudaMalloc(&c, 4 * 8);
    latency_probe<<<1, 32>>>(4096, o, c, 1.0000001, 1e-9);
    cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
    printf("latency (clk / dependent op, one warp): DFMA %.1f, rsqrt(double)+DADD %.1f, shfl(double)+DADD %.1f, FFMA %.1f\n",
           h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0, h[3] / 4096.0);
  }
  double *out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(double));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int ctas_per_sm = 1; ctas_per_sm <= 8; ctas_per_sm *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      dfma_probe<8><<<148 * ctas_per_sm, 256>>>(iters, out, 1.0000001, 1e-9);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep) printf("DFMA  %d CTAs/SM x 256 thr x 8 chains: %.2f TDFMA/s (%.1f TFLOP/s)\n", ctas_per_sm,
                      148.0 * ctas_per_sm * 256 * 8 * iters / ms / 1e9, 2 * 148.0 * ctas_per_sm * 256 * 8 * iters / ms / 1e9);
    }
  }
  for (int ctas_per_sm = 1; ctas_per_sm <= 8; ctas_per_sm *= 2) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      dmma_probe<<<148 * ctas_per_sm, 256>>>(iters, out, 1.0000001, 1e-9);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      // one m8n8k4 = 256 FMA per warp
      if (rep) printf("DMMA  %d CTAs/SM x 8 warps x 4 chains: %.2f TDFMA/s (%.1f TFLOP/s)\n", ctas_per_sm,
                      148.0 * ctas_per_sm * 8 * 4 * 256.0 * iters / ms / 1e9, 2 * 148.0 * ctas_per_sm * 8 * 4 * 256.0 * iters / ms / 1e9);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
