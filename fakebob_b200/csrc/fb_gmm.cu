// Diagonal-GMM frame log-likelihoods on the 5th-generation tensor cores.
//
// Replaces `gmm-global-get-frame-likes --average=true <model> <feats>` run once per model per
// score() call by the reference (gmm_ubm_kaldiHelper.py:202-221); upstream arithmetic
// DiagGmm::LogLikelihood + LogSumExp, SURVEY.md Appendix A.7:
//     ll[t,c] = gconst_c + (mu/var)_c . x_t - 0.5 (1/var)_c . x_t^2 ,   LL_t = logsumexp_c ll[t,c]
//
// As a contraction:  [x | x^2] (rows x 144)  .  [mu/var | -0.5/var]^T (144 x C).
// Precision: operands are split into fp16 hi + lo parts (after an exact power-of-two
// per-dimension scaling) and three MMAs  hi.hi + lo.hi + hi.lo  accumulate in fp32 in TMEM,
// which keeps ~22 significant bits per operand (fp32 Kaldi keeps 24).  W is pre-multiplied by
// log2(e) so the epilogue works in the log2 domain with ex2.approx.
//
// Kernel shape (cta_group::1, persistent, 1 CTA / SM, 192 threads):
//   warp 0   bulk-copy (TMA engine, cp.async.bulk) producer: A super-tile (2 x 128 rows, hi+lo,
//            144 KB, resident for all models/columns it is used with) and W stages (64 columns,
//            hi+lo, 36 KB, 2-deep ring).  Operand images are stored in global memory already in
//            the UMMA canonical no-swizzle K-major core-matrix order, so every copy is one
//            contiguous bulk transfer.
//   warp 1   single-thread tcgen05.mma issuer: per stage 2 tiles x 3 parts x 9 k-blocks of
//            M128 N64 K16, accumulators double-buffered in TMEM (2 x [2 tiles x 128 cols]).
//   warps 2-5 epilogue: tcgen05.ld 32x32b.x32, + gconst, online max / sum of ex2 with Kaldi's
//            log(FLT_EPSILON) pruning, one (max,sum) partial per row per 128-column unit.
// Work unit = (super-tile, model, 128-column chunk); units are split evenly over the CTAs.
#include "fb_common.cuh"
#include <math.h>

#define GMM_THREADS 192
static constexpr uint32_t kAHalfBytes = FB_KSLABS * FB_TILE_M * 16;           // 36864: one of hi/lo of one tile
static constexpr uint32_t kATileBytes = 2 * kAHalfBytes;                      // 73728
static constexpr uint32_t kASuperBytes = 2 * kATileBytes;                     // 147456
static constexpr uint32_t kWHalfBytes = FB_KSLABS * FB_STAGE_N * 16;          // 18432
static constexpr uint32_t kWStageBytes = 2 * kWHalfBytes;                     // 36864
static constexpr uint32_t kNumWStages = 2;
static constexpr uint32_t kSmemA = 0;
static constexpr uint32_t kSmemW = kASuperBytes;
static constexpr uint32_t kSmemBar = kSmemW + kNumWStages * kWStageBytes;     // 221184
static constexpr uint32_t kSmemTotal = kSmemBar + 128;
static constexpr uint32_t kSmemLaunch = kSmemTotal + 1024;                    // alignment slack

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
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
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// K-major, no-swizzle canonical layout: core matrix = 8 rows x 16 bytes (contiguous 128 B);
// LBO = byte distance between core matrices adjacent in K, SBO = between 8-row groups in M/N.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= 1ull << 46;                             // descriptor version (Blackwell)
  return d;                                    // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

// kind::f16 instruction descriptor: D=f32, A=B=f16, both K-major, N>>3 at [17,23), M>>4 at [24,29)
static constexpr uint32_t kIdesc = (1u << 4) | ((FB_STAGE_N >> 3) << 17) | ((FB_TILE_M >> 4) << 24);

struct GmmArgs {
  const __half *a_img;      // [super-tile][tile 2][hl 2][18][128][8]
  const __half *w_img;      // [model][C/64][hl 2][18][64][8]
  const float *gconst2;     // [model][C]
  float2 *part;             // [model][C/128][rows_cap]
  const int *misc;          // misc[2] = total voiced rows
  const int *done_flag;
  int n_models, C, rows_cap;
};

__global__ void __launch_bounds__(GMM_THREADS, 1) gmm_umma_kernel(GmmArgs g) {
  if (g.done_flag && *g.done_flag) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t *smem = smem_raw + (base - raw_addr);
  const uint32_t bar0 = base + kSmemBar;
  const uint32_t bar_a_full = bar0, bar_a_empty = bar0 + 8;
  const uint32_t bar_w_full = bar0 + 16, bar_w_empty = bar0 + 32;       // [2] each
  const uint32_t bar_acc_full = bar0 + 48, bar_acc_empty = bar0 + 64;   // [2] each
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + kSmemBar + 96);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = g.misc[2];
  const int nch = g.C / FB_CHUNK_N;
  const int n_super = (M + 2 * FB_TILE_M - 1) / (2 * FB_TILE_M);
  const long long n_units = (long long)n_super * g.n_models * nch;
  const int u0 = (int)(n_units * blockIdx.x / gridDim.x);
  const int u1 = (int)(n_units * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0 && lane == 0) {
    mbar_init(bar_a_full, 1);
    mbar_init(bar_a_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_w_full + 8 * i, 1);
      mbar_init(bar_w_empty + 8 * i, 1);
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 4);     // one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void *)tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= producer =================
    if (lane == 0) {
      int cur_super = -1;
      uint32_t a_cnt = 0, w_cnt = 0;
      for (int u = u0; u < u1; ++u) {
        const int ch = u % nch;
        const int item = u / nch;
        const int model = item % g.n_models;
        const int sp = item / g.n_models;
        if (sp != cur_super) {
          if (a_cnt > 0) mbar_wait(bar_a_empty, (a_cnt - 1) & 1);
          mbar_expect_tx(bar_a_full, kASuperBytes);
          const uint8_t *src = reinterpret_cast<const uint8_t *>(g.a_img) + (size_t)sp * kASuperBytes;
#pragma unroll
          for (int q = 0; q < 4; ++q)
            bulk_g2s(base + kSmemA + q * kAHalfBytes, src + (size_t)q * kAHalfBytes, kAHalfBytes, bar_a_full);
          ++a_cnt;
          cur_super = sp;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t s = w_cnt & 1;
          mbar_wait(bar_w_empty + 8 * s, ((w_cnt >> 1) & 1) ^ 1);
          mbar_expect_tx(bar_w_full + 8 * s, kWStageBytes);
          const size_t stage_idx = (size_t)model * (g.C / FB_STAGE_N) + (size_t)ch * 2 + h;
          const uint8_t *src = reinterpret_cast<const uint8_t *>(g.w_img) + stage_idx * kWStageBytes;
          bulk_g2s(base + kSmemW + s * kWStageBytes, src, kWHalfBytes, bar_w_full + 8 * s);
          bulk_g2s(base + kSmemW + s * kWStageBytes + kWHalfBytes, src + kWHalfBytes, kWHalfBytes, bar_w_full + 8 * s);
          ++w_cnt;
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      int cur_super = -1;
      uint32_t a_cnt = 0, w_cnt = 0, acc_cnt = 0;
      for (int u = u0; u < u1; ++u) {
        const int item = u / nch;
        const int sp = item / g.n_models;
        if (sp != cur_super) {
          mbar_wait(bar_a_full, a_cnt & 1);
          ++a_cnt;
          cur_super = sp;
        }
        const uint32_t buf = acc_cnt & 1;
        mbar_wait(bar_acc_empty + 8 * buf, ((acc_cnt >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          const uint32_t s = w_cnt & 1;
          mbar_wait(bar_w_full + 8 * s, (w_cnt >> 1) & 1);
          tc_fence_after();
          const uint32_t w_hi = base + kSmemW + s * kWStageBytes;
          const uint32_t w_lo = w_hi + kWHalfBytes;
#pragma unroll
          for (int tile = 0; tile < 2; ++tile) {
            const uint32_t a_hi = base + kSmemA + tile * kATileBytes;
            const uint32_t a_lo = a_hi + kAHalfBytes;
            const uint32_t d_tmem = tmem_base + buf * 256 + tile * FB_CHUNK_N + h * FB_STAGE_N;
#pragma unroll
            for (int part = 0; part < 3; ++part) {
              const uint32_t a_base = (part == 1) ? a_lo : a_hi;
              const uint32_t b_base = (part == 2) ? w_lo : w_hi;
#pragma unroll
              for (int kb = 0; kb < FB_KSLABS / 2; ++kb) {
                const uint64_t ad = make_desc(a_base + kb * 2 * (FB_TILE_M * 16), FB_TILE_M * 16, 128);
                const uint64_t bd = make_desc(b_base + kb * 2 * (FB_STAGE_N * 16), FB_STAGE_N * 16, 128);
                tc_mma_f16(d_tmem, ad, bd, kIdesc, (part | kb) ? 1u : 0u);
              }
            }
          }
          tc_commit(bar_w_empty + 8 * s);
          ++w_cnt;
        }
        tc_commit(bar_acc_full + 8 * buf);
        ++acc_cnt;
        const bool last_of_super = (u + 1 == u1) || ((u + 1) / nch / g.n_models != sp);
        if (last_of_super) tc_commit(bar_a_empty);
      }
    }
  } else {
    // ================= epilogue (warps 2..5; TMEM lane quadrant = warp % 4) =================
    const int quad = warp & 3;
    uint32_t acc_cnt = 0;
    for (int u = u0; u < u1; ++u) {
      const int ch = u % nch;
      const int item = u / nch;
      const int model = item % g.n_models;
      const int sp = item / g.n_models;
      const uint32_t buf = acc_cnt & 1;
      mbar_wait(bar_acc_full + 8 * buf, (acc_cnt >> 1) & 1);
      tc_fence_after();
      const float *gc = g.gconst2 + (size_t)model * g.C + ch * FB_CHUNK_N;
#pragma unroll 1
      for (int tile = 0; tile < 2; ++tile) {
        float m = -INFINITY, s = 0.f;
#pragma unroll 1
        for (int c8 = 0; c8 < 4; ++c8) {
          float v[32];
          tc_ld32(tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 256 + tile * FB_CHUNK_N + c8 * 32, v);
          tc_wait_ld();
          float cmax = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const float4 gq = __ldg(reinterpret_cast<const float4 *>(gc + c8 * 32 + i));
            v[i] += gq.x; v[i + 1] += gq.y; v[i + 2] += gq.z; v[i + 3] += gq.w;
            cmax = fmaxf(cmax, fmaxf(fmaxf(v[i], v[i + 1]), fmaxf(v[i + 2], v[i + 3])));
          }
          if (cmax > m) {
            s *= ex2_approx(m - cmax);
            m = cmax;
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float t = v[i] - m;
            const float e = ex2_approx(t);
            s += (t >= -23.0f) ? e : 0.f;          // Kaldi LogSumExp cutoff log(FLT_EPSILON) = -23 in log2
          }
        }
        const int row = sp * (2 * FB_TILE_M) + tile * FB_TILE_M + quad * 32 + lane;
        g.part[((size_t)model * nch + ch) * g.rows_cap + row] = make_float2(m, s);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * buf);
      ++acc_cnt;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 CUDA-core cross-check kernel (debug / bring-up): same partial format, operands rebuilt
// from the hi+lo image so it sees exactly the features the tensor-core kernel sees.
// grid (rows/32, n_models * nch), block 128: each thread = one column of the 128-column chunk.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gmm_simt_kernel(const __half *__restrict__ a_img, const float *__restrict__ w_f32, const float *__restrict__ gconst_nat,
                const float *__restrict__ feat_scale, float2 *__restrict__ part, const int *__restrict__ misc,
                int n_models, int C, int rows_cap) {
  __shared__ float s_x[32][2 * FB_DIM + 1];
  __shared__ float s_ll[32][FB_CHUNK_N + 1];
  const int M = misc[2];
  const int row0 = blockIdx.x * 32;
  if (row0 >= M) return;
  const int nch = C / FB_CHUNK_N;
  const int model = blockIdx.y / nch, ch = blockIdx.y % nch;
  for (int idx = threadIdx.x; idx < 32 * 2 * FB_DIM; idx += blockDim.x) {
    const int r = idx / (2 * FB_DIM), k = idx % (2 * FB_DIM);
    const int row = row0 + r;
    const int tile = row >> 7, rr = row & 127, slab = k >> 3, e = k & 7;
    const size_t b = ((size_t)tile * 2 * FB_KSLABS + slab) * (FB_TILE_M * 8) + rr * 8 + e;
    float v = __half2float(a_img[b]) + __half2float(a_img[b + (size_t)FB_KSLABS * FB_TILE_M * 8]);
    const int d = (k < FB_DIM) ? k : k - FB_DIM;
    const float sc = feat_scale[d];
    v = (k < FB_DIM) ? v / sc : v / (sc * sc);
    s_x[r][k] = v;
  }
  __syncthreads();
  const int c = ch * FB_CHUNK_N + threadIdx.x;
  const float *w = w_f32 + ((size_t)model * C + c) * (2 * FB_DIM);
  const float gcv = gconst_nat[(size_t)model * C + c];
  for (int r = 0; r < 32; ++r) {
    float acc = gcv;
    for (int k = 0; k < 2 * FB_DIM; ++k) acc += w[k] * s_x[r][k];
    s_ll[r][threadIdx.x] = acc * 1.4426950408889634f;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int r = threadIdx.x;
    float m = -INFINITY;
    for (int k = 0; k < FB_CHUNK_N; ++k) m = fmaxf(m, s_ll[r][k]);
    float s = 0.f;
    for (int k = 0; k < FB_CHUNK_N; ++k) {
      const float t = s_ll[r][k] - m;
      if (t >= -23.0f) s += exp2f(t);
    }
    part[((size_t)model * nch + ch) * rows_cap + row0 + r] = make_float2(m, s);
  }
}

// ------------------------------------------------------------------------------------------------
// Merge the per-chunk partials into per-frame log-likelihoods and the per-utterance average
// (gmm-global-get-frame-likes --average=true: float frame values, double sum, float quotient).
// grid (B, n_models), block 128, fixed reduction order (deterministic).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gmm_reduce_kernel(const float2 *__restrict__ part, const int *__restrict__ row_off, float *__restrict__ frame_ll,
                  double *__restrict__ avg_ll, int n_models, int nch, int rows_cap, const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  __shared__ double s_red[4];
  const int b = blockIdx.x, model = blockIdx.y;
  const int r0 = row_off[b], r1 = row_off[b + 1];
  double acc = 0.0;
  for (int row = r0 + threadIdx.x; row < r1; row += blockDim.x) {
    float mx = -INFINITY;
    for (int k = 0; k < nch; ++k) mx = fmaxf(mx, part[((size_t)model * nch + k) * rows_cap + row].x);
    double s = 0.0;
    for (int k = 0; k < nch; ++k) {
      const float2 p = part[((size_t)model * nch + k) * rows_cap + row];
      if (p.x >= mx - 23.0f) s += (double)p.y * exp2((double)(p.x - mx));
    }
    const float ll = (float)(((double)mx + log2(s)) * 0.6931471805599453);
    frame_ll[(size_t)model * rows_cap + row] = ll;
    acc += (double)ll;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double tot = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
    const int n = r1 - r0;
    avg_ll[(size_t)b * n_models + model] = (n > 0) ? (double)__fdiv_rn((float)tot, (float)n) : 0.0;
  }
}

// ------------------------------------------------------------------------------------------------
// Host: model upload
// ------------------------------------------------------------------------------------------------
extern "C" int fb_load_diag_gmm(fb_ctx *ctx, int slot, const float *weights, const float *means_invvars,
                                const float *inv_vars, const float *gconsts, int C, int D) {
  FB_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  FB_CHECK_ARG(slot >= 0 && slot < FB_MAX_MODELS, "slot out of range");
  FB_CHECK_ARG(D == FB_DIM, "feature dimension must be 72");
  FB_CHECK_ARG(C > 0 && C % FB_CHUNK_N == 0, "number of components must be a positive multiple of 128");
  FB_CHECK_ARG(weights && means_invvars && inv_vars && gconsts, "NULL parameter array");
  FbHostGmm &h = ctx->host_gmm[slot];
  h.weights.assign(weights, weights + C);
  h.means_invvars.assign(means_invvars, means_invvars + (size_t)C * D);
  h.inv_vars.assign(inv_vars, inv_vars + (size_t)C * D);
  h.gconsts.assign(gconsts, gconsts + C);
  h.loaded = true;
  return FB_OK;
}

extern "C" int fb_finalize_gmms(fb_ctx *ctx, int n_models) {
  FB_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  FB_CHECK_ARG(n_models > 0 && n_models <= FB_MAX_MODELS, "n_models out of range");
  const int C = (int)ctx->host_gmm[0].weights.size();
  for (int m = 0; m < n_models; ++m) {
    FB_CHECK_ARG(ctx->host_gmm[m].loaded, "a model slot below n_models was never loaded");
    FB_CHECK_ARG((int)ctx->host_gmm[m].weights.size() == C, "all models must have the same number of components");
  }
  FB_CUDA(cudaSetDevice(ctx->device));
  // per-dimension power-of-two scale from slot 0's second moments: x' = x * s_d is O(1)
  const FbHostGmm &g0 = ctx->host_gmm[0];
  for (int d = 0; d < FB_DIM; ++d) {
    double acc = 0.0, wsum = 0.0;
    for (int c = 0; c < C; ++c) {
      const double iv = g0.inv_vars[(size_t)c * FB_DIM + d];
      const double mu = g0.means_invvars[(size_t)c * FB_DIM + d] / iv;
      acc += (double)g0.weights[c] * (mu * mu + 1.0 / iv);
      wsum += g0.weights[c];
    }
    const double rms = sqrt(acc / (wsum > 0 ? wsum : 1.0));
    int e = (int)lrint(log2(rms > 1e-30 ? rms : 1.0));
    if (e > 60) e = 60;
    if (e < -60) e = -60;
    ctx->tables_host.feat_scale[d] = (float)ldexp(1.0, -e);
  }
  ctx->tables_dirty = true;
  const int n_stage = C / FB_STAGE_N;
  const size_t img_halfs = (size_t)n_models * n_stage * 2 * FB_KSLABS * FB_STAGE_N * 8;
  std::vector<__half> img(img_halfs);
  std::vector<float> gc2((size_t)n_models * C), gcn((size_t)n_models * C), wf((size_t)n_models * C * 2 * FB_DIM);
  const double log2e = 1.4426950408889634;
  for (int m = 0; m < n_models; ++m) {
    const FbHostGmm &h = ctx->host_gmm[m];
    for (int c = 0; c < C; ++c) {
      gc2[(size_t)m * C + c] = (float)((double)h.gconsts[c] * log2e);
      gcn[(size_t)m * C + c] = h.gconsts[c];
      const int st = c / FB_STAGE_N, cc = c % FB_STAGE_N;
      for (int k = 0; k < 2 * FB_DIM; ++k) {
        const int d = (k < FB_DIM) ? k : k - FB_DIM;
        const double s = ctx->tables_host.feat_scale[d];
        const double raw = (k < FB_DIM) ? (double)h.means_invvars[(size_t)c * FB_DIM + d]
                                        : -0.5 * (double)h.inv_vars[(size_t)c * FB_DIM + d];
        wf[((size_t)m * C + c) * 2 * FB_DIM + k] = (float)raw;
        const double w = ((k < FB_DIM) ? raw / s : raw / (s * s)) * log2e;
        if (!(fabs(w) < 60000.0)) {
          fb_set_error("model %d component %d dim %d: scaled weight %g exceeds the fp16 range", m, c, k, w);
          return FB_ERR_UNSUPPORTED;
        }
        const __half hi = __float2half_rn((float)w);
        const __half lo = __float2half_rn((float)(w - (double)__half2float(hi)));
        const int slab = k >> 3, e = k & 7;
        const size_t b = ((((size_t)m * n_stage + st) * 2 + 0) * FB_KSLABS + slab) * (FB_STAGE_N * 8) + cc * 8 + e;
        img[b] = hi;
        img[b + (size_t)FB_KSLABS * FB_STAGE_N * 8] = lo;
      }
    }
  }
  int rc;
  if ((rc = ctx->w_img.ensure(img_halfs))) return rc;
  if ((rc = ctx->gconst2.ensure(gc2.size()))) return rc;
  if ((rc = ctx->gconst_nat.ensure(gcn.size()))) return rc;
  if ((rc = ctx->w_f32.ensure(wf.size()))) return rc;
  FB_CUDA(cudaMemcpy(ctx->w_img.p, img.data(), img_halfs * sizeof(__half), cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(ctx->gconst2.p, gc2.data(), gc2.size() * sizeof(float), cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(ctx->gconst_nat.p, gcn.data(), gcn.size() * sizeof(float), cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(ctx->w_f32.p, wf.data(), wf.size() * sizeof(float), cudaMemcpyHostToDevice));
  ctx->n_models = n_models;
  ctx->C = C;
  // buffers that depend on n_models are (re)sized at the next fb_reserve_batch
  ctx->part.release();
  ctx->frame_ll.release();
  ctx->avg_ll.release();
  static bool attr_set = false;
  if (!attr_set) {
    FB_CUDA(cudaFuncSetAttribute(gmm_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
    attr_set = true;
  }
  return FB_OK;
}

int fb_run_gmm_flag(fb_ctx *ctx, const int *done_flag) {
  FB_CHECK_ARG(ctx->n_models > 0, "no GMMs loaded (fb_finalize_gmms)");
  const int nch = ctx->C / FB_CHUNK_N;
  if (ctx->gmm_impl == 0) {
    GmmArgs a;
    a.a_img = ctx->a_img.p;
    a.w_img = ctx->w_img.p;
    a.gconst2 = ctx->gconst2.p;
    a.part = ctx->part.p;
    a.misc = ctx->misc.p;
    a.done_flag = done_flag;
    a.n_models = ctx->n_models;
    a.C = ctx->C;
    a.rows_cap = ctx->rows_cap;
    // upper bound on useful CTAs: one unit each
    const long long max_units = (long long)fb_div_up(ctx->total_frames, 2 * FB_TILE_M) * ctx->n_models * nch;
    int grid = ctx->num_sms;
    if (max_units < grid) grid = (int)max_units;
    gmm_umma_kernel<<<grid, GMM_THREADS, kSmemLaunch, ctx->stream>>>(a);
  } else {
    dim3 grid(fb_div_up(ctx->total_frames, 32), ctx->n_models * nch);
    gmm_simt_kernel<<<grid, 128, 0, ctx->stream>>>(ctx->a_img.p, ctx->w_f32.p, ctx->gconst_nat.p,
                                                   ctx->tables_dev->feat_scale, ctx->part.p, ctx->misc.p,
                                                   ctx->n_models, ctx->C, ctx->rows_cap);
  }
  fb_prof_mark(ctx, 4);
  dim3 g2(ctx->B, ctx->n_models);
  gmm_reduce_kernel<<<g2, 128, 0, ctx->stream>>>(ctx->part.p, ctx->row_off.p, ctx->frame_ll.p, ctx->avg_ll.p,
                                                 ctx->n_models, nch, ctx->rows_cap, done_flag);
  fb_prof_mark(ctx, 5);
  ctx->launches += 2;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fb_run_gmm(fb_ctx *ctx) { return fb_run_gmm_flag(ctx, nullptr); }
