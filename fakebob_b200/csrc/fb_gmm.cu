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
// Kernel shape (cta_group::1, persistent, 1 CTA / SM, 320 threads).  Measured on B200 (profiles/r01_*): with both
// operands in shared memory the N=64 MMAs need 192 B/clk of smem reads and run at ~68 clk instead of 32, and they starve
// the bulk-copy writes, so the A operand lives in TENSOR MEMORY instead:
//   warp 0   bulk-copy (TMA engine, cp.async.bulk) producer into a 5-slot x 40 KB ring: per 256-row super-tile four
//            A entries (tile0 hi, tile0 lo, tile1 hi, tile1 lo), then per 128-column unit two W stages (64 columns,
//            hi+lo).  Operand images are stored in global memory already in the UMMA canonical no-swizzle K-major
//            core-matrix order, so every copy is one contiguous transfer.
//   warp 1   single thread: tcgen05.cp moves A entries smem -> TMEM (304 columns: 2 tiles x (hi 80 + lo 72)), then per W
//            stage and tile 28 tcgen05.mma M128 N64 K16 with A from TMEM and B from the ring slot; the two tiles'
//            accumulators (64 TMEM columns each) ping-pong against the epilogue.
//   warps 2-9 epilogue (4 per tile, TMEM lane quadrant = warp % 4): tcgen05.ld 32x32b.x32, online max / sum of ex2 with
//            Kaldi's log(FLT_EPSILON) pruning; gconst is already inside the accumulator (extra k-block of the hi.hi
//            part: "ones" columns of A times [g_hi g_mid g_lo] rows of W); one (max,sum) partial per row per unit.
// Work unit = (super-tile, model, 128-column chunk); units are split evenly over the CTAs.
#include "fb_common.cuh"
#include <math.h>

#ifndef GMM_PARTS
#define GMM_PARTS 3
#endif
#define GMM_THREADS 320                       // producer warp + MMA warp + 8 epilogue warps
// A image per 128-row tile: hi = 20 slabs (9 x, 9 x^2, 1 "ones" slab carrying the gconst columns, 1 zero slab), lo = 18 slabs.
// W image per 64-column stage: hi = 20 slabs (18 + gconst slab [g_hi g_mid g_lo 0..] + zero slab), lo = 18 slabs.
static constexpr uint32_t kSlabA = FB_TILE_M * 16;                             // 2048 B
static constexpr uint32_t kSlabW = FB_STAGE_N * 16;                            // 1024 B
static constexpr uint32_t kAHiBytes = FB_A_HI_SLABS * kSlabA;                  // 40960
static constexpr uint32_t kALoBytes = FB_KSLABS * kSlabA;                      // 36864
static constexpr uint32_t kATileBytes = kAHiBytes + kALoBytes;                 // 77824
static constexpr uint32_t kWHiBytes = FB_W_HI_SLABS * kSlabW;                  // 20480
static constexpr uint32_t kWLoBytes = FB_KSLABS * kSlabW;                      // 18432
static constexpr uint32_t kWStageBytes = kWHiBytes + kWLoBytes;                // 38912
static constexpr uint32_t kSlotBytes = 40960;
static constexpr uint32_t kNumSlots = 5;
static constexpr uint32_t kSmemBar = kNumSlots * kSlotBytes;                   // 204800
static constexpr uint32_t kSmemTotal = kSmemBar + 256;
static constexpr uint32_t kSmemLaunch = kSmemTotal + 128;                      // 128 B alignment slack
static_assert(kAHiBytes <= kSlotBytes && kWStageBytes <= kSlotBytes, "ring slot too small");
static_assert(kSmemLaunch <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
// TMEM columns: accumulators [0,128) (tile0 | tile1, 64 each); A operand: tile t hi at 128 + 152 t, lo 80 columns later
static constexpr uint32_t kTmemAcc = 0;
static constexpr uint32_t kTmemA = 128;
static constexpr uint32_t kTmemAHiCols = FB_A_HI_SLABS * 4;                    // 80
static constexpr uint32_t kTmemATileCols = kTmemAHiCols + FB_KSLABS * 4;       // 152

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
__device__ __forceinline__ void tc_mma_f16_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
// smem (matrix descriptor: 128 rows x 32 bytes = one K=16 fp16 block) -> TMEM lanes 0..127, 8 columns
__device__ __forceinline__ void tc_cp_128x256b(uint32_t taddr, uint64_t s_desc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(s_desc) : "memory");
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

// One lane of a converged warp; lets the compiler treat the guarded region as single-threaded (uniform registers).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float y;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
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
  const __half *a_img;      // [tile][hi 19 slabs | lo 18 slabs][128][8]
  const __half *w_img;      // [model][C/64][hi 20 slabs | lo 18 slabs][64][8]
  float2 *part;             // [model][C/128][rows_cap]
  const int *misc;          // misc[2] = total voiced rows
  const int *done_flag;
  float *ll_out;            // STORE mode: [rows_cap][C] natural-log component log-likelihoods (Gaussian selection)
  int n_models, C, rows_cap;
};

// One 32-column group of one accumulator row: online max / sum of 2^(v - max) with Kaldi's cutoff.
__device__ __forceinline__ void lse_group(const float *v, float &m, float &s) {
  float c0 = max3(v[0], v[1], v[2]), c1 = max3(v[3], v[4], v[5]);
#pragma unroll
  for (int i = 6; i < 30; i += 6) {
    c0 = max3(c0, v[i], v[i + 1]);
    c0 = fmaxf(c0, v[i + 2]);
    c1 = max3(c1, v[i + 3], v[i + 4]);
    c1 = fmaxf(c1, v[i + 5]);
  }
  const float cmax = max3(c0, c1, fmaxf(v[30], v[31]));
  if (cmax > m) {
    s *= ex2_approx(m - cmax);
    m = cmax;
  }
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const float t0 = v[i] - m, t1 = v[i + 1] - m;
    const float e0 = ex2_approx(t0), e1 = ex2_approx(t1);
    if (t0 >= -23.0f) s0 += e0;                 // Kaldi LogSumExp cutoff log(FLT_EPSILON) = -23 in log2
    if (t1 >= -23.0f) s1 += e1;
  }
  s += s0 + s1;
}

template <bool kStore>
__global__ void __launch_bounds__(GMM_THREADS, 1) gmm_umma_kernel(GmmArgs g) {
  if (g.done_flag && *g.done_flag) return;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 127u) & ~127u;
  uint8_t *smem = smem_raw + (base - raw_addr);
  const uint32_t bar0 = base + kSmemBar;
  const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * kNumSlots;           // ring slots
  const uint32_t bar_acc_full = bar0 + 16 * kNumSlots, bar_acc_empty = bar_acc_full + 16;   // [2 tiles] each
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + kSmemBar + 16 * kNumSlots + 32);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int M = g.misc[2];
  const int nch = g.C / FB_CHUNK_N;
  const int n_super = (M + 2 * FB_TILE_M - 1) / (2 * FB_TILE_M);
  const long long n_units = (long long)n_super * g.n_models * nch;
  const int u0 = (int)(n_units * blockIdx.x / gridDim.x);
  const int u1 = (int)(n_units * (blockIdx.x + 1) / gridDim.x);

  if (warp == 0 && lane == 0) {
    for (uint32_t i = 0; i < kNumSlots; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 4);     // one arrive per epilogue warp of that tile
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
    // ================= producer: ring of kNumSlots slots (whole warp loops, one elected lane issues) =================
    int cur_super = -1;
    uint32_t cnt = 0;
    for (int u = u0; u < u1; ++u) {
      const int ch = u % nch;
      const int item = u / nch;
      const int model = item % g.n_models;
      const int sp = item / g.n_models;
      if (sp != cur_super) {
        const uint8_t *src = reinterpret_cast<const uint8_t *>(g.a_img) + (size_t)sp * 2 * kATileBytes;
#pragma unroll 1
        for (int e = 0; e < 4; ++e) {          // tile0 hi, tile0 lo, tile1 hi, tile1 lo
          const uint32_t slot = cnt % kNumSlots;
          mbar_wait(bar_empty + 8 * slot, ((cnt / kNumSlots) & 1) ^ 1);
          if (elect_one()) {
            const uint32_t bytes = (e & 1) ? kALoBytes : kAHiBytes;
            mbar_expect_tx(bar_full + 8 * slot, bytes);
            bulk_g2s(base + slot * kSlotBytes, src + (size_t)(e >> 1) * kATileBytes + ((e & 1) ? kAHiBytes : 0), bytes,
                     bar_full + 8 * slot);
          }
          __syncwarp();
          ++cnt;
        }
        cur_super = sp;
      }
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const uint32_t slot = cnt % kNumSlots;
        mbar_wait(bar_empty + 8 * slot, ((cnt / kNumSlots) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(bar_full + 8 * slot, kWStageBytes);
          const size_t stage_idx = (size_t)model * (g.C / FB_STAGE_N) + (size_t)ch * 2 + h;
          bulk_g2s(base + slot * kSlotBytes, reinterpret_cast<const uint8_t *>(g.w_img) + stage_idx * kWStageBytes,
                   kWStageBytes, bar_full + 8 * slot);
        }
        __syncwarp();
        ++cnt;
      }
    }
  } else if (warp == 1) {
    // ===== tcgen05 issuer (whole warp loops and waits; one elected lane issues copies / MMAs / commits) =====
    // smem descriptor with LBO / SBO / version bits and a zero start address; only the 14-bit address field varies
    const uint64_t desc_w = make_desc(0, kSlabW, 128);
    const uint64_t desc_a = make_desc(0, kSlabA, 128);
    int cur_super = -1;
    uint32_t cnt = 0, sub0 = 0, sub1 = 0;
    for (int u = u0; u < u1; ++u) {
      const int item = u / nch;
      const int sp = item / g.n_models;
      if (sp != cur_super) {
#pragma unroll 1
        for (int e = 0; e < 4; ++e) {
          const uint32_t slot = cnt % kNumSlots;
          mbar_wait(bar_full + 8 * slot, (cnt / kNumSlots) & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t src = desc_a + (uint64_t)(((base + slot * kSlotBytes) & 0x3FFFFu) >> 4);
            const uint32_t dst = tmem_base + kTmemA + (e >> 1) * kTmemATileCols + ((e & 1) ? kTmemAHiCols : 0);
            const int nkb = (e & 1) ? FB_KSLABS / 2 : FB_A_HI_SLABS / 2;
            for (int kb = 0; kb < nkb; ++kb) tc_cp_128x256b(dst + kb * 8, src + (uint64_t)(kb * ((2 * kSlabA) >> 4)));
            tc_commit(bar_empty + 8 * slot);
          }
          __syncwarp();
          ++cnt;
        }
        cur_super = sp;
      }
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const uint32_t slot = cnt % kNumSlots;
        mbar_wait(bar_full + 8 * slot, (cnt / kNumSlots) & 1);
        const uint64_t w_hi = desc_w + (uint64_t)(((base + slot * kSlotBytes) & 0x3FFFFu) >> 4);
        const uint64_t w_lo = w_hi + (uint64_t)(kWHiBytes >> 4);
#pragma unroll
        for (int tile = 0; tile < 2; ++tile) {
          mbar_wait(bar_acc_empty + 8 * tile, ((tile ? sub1 : sub0) & 1) ^ 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_hi = tmem_base + kTmemA + tile * kTmemATileCols;
            const uint32_t a_lo = a_hi + kTmemAHiCols;
            const uint32_t d_tmem = tmem_base + kTmemAcc + tile * FB_STAGE_N;
#pragma unroll
            for (int part = 0; part < GMM_PARTS; ++part) {
              const uint32_t a_base = (part == 1) ? a_lo : a_hi;
              const uint64_t b_base = (part == 2) ? w_lo : w_hi;
              const int nkb = (part == 0) ? FB_A_HI_SLABS / 2 : FB_KSLABS / 2;      // part 0 carries the gconst k-block
#pragma unroll
              for (int kb = 0; kb < nkb; ++kb)
                tc_mma_f16_ta(d_tmem, a_base + kb * 8, b_base + (uint64_t)(kb * ((2 * kSlabW) >> 4)), kIdesc, (part | kb) ? 1u : 0u);
            }
            tc_commit(bar_acc_full + 8 * tile);
            if (tile == 1) tc_commit(bar_empty + 8 * slot);
          }
          __syncwarp();
          if (tile) ++sub1; else ++sub0;
        }
        ++cnt;
      }
    }
  } else {
    // ===== epilogue: warps 2..9; TMEM lane quadrant = warp % 4, tile = (warp - 2) / 4 =====
    const int quad = warp & 3;
    const int tile = (warp - 2) >> 2;
    uint32_t sub = 0;
    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + kTmemAcc + tile * FB_STAGE_N;
    for (int u = u0; u < u1; ++u) {
      const int ch = u % nch;
      const int item = u / nch;
      const int model = item % g.n_models;
      const int sp = item / g.n_models;
      float m = -INFINITY, s = 0.f;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        mbar_wait(bar_acc_full + 8 * tile, sub & 1);
        tc_fence_after();
        float va[32], vb[32];
        tc_ld32(taddr, va);
        tc_ld32(taddr + 32, vb);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * tile);     // accumulator is in registers: MMA may overwrite
        ++sub;
        if (kStore) {
          const int row = sp * (2 * FB_TILE_M) + tile * FB_TILE_M + quad * 32 + lane;
          float4 *dst = reinterpret_cast<float4 *>(g.ll_out + (size_t)row * g.C + ch * FB_CHUNK_N + h * FB_STAGE_N);
          const float ln2 = 0.6931471805599453f;
#pragma unroll
          for (int i = 0; i < 32; i += 4) dst[i >> 2] = make_float4(va[i] * ln2, va[i + 1] * ln2, va[i + 2] * ln2, va[i + 3] * ln2);
#pragma unroll
          for (int i = 0; i < 32; i += 4) dst[8 + (i >> 2)] = make_float4(vb[i] * ln2, vb[i + 1] * ln2, vb[i + 2] * ln2, vb[i + 3] * ln2);
        } else {
          lse_group(va, m, s);
          lse_group(vb, m, s);
        }
      }
      if (!kStore) {
        const int row = sp * (2 * FB_TILE_M) + tile * FB_TILE_M + quad * 32 + lane;
        g.part[((size_t)model * nch + ch) * g.rows_cap + row] = make_float2(m, s);
      }
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
    const size_t b = ((size_t)tile * FB_A_TILE_SLABS + slab) * (FB_TILE_M * 8) + rr * 8 + e;
    float v = __half2float(a_img[b]) + __half2float(a_img[b + (size_t)FB_A_HI_SLABS * FB_TILE_M * 8]);
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
  const size_t stage_halfs = (size_t)(FB_W_HI_SLABS + FB_KSLABS) * FB_STAGE_N * 8;
  const size_t img_halfs = (size_t)n_models * n_stage * stage_halfs;
  std::vector<__half> img(img_halfs, __float2half_rn(0.f));
  std::vector<float> gc2((size_t)n_models * C), gcn((size_t)n_models * C), wf((size_t)n_models * C * 2 * FB_DIM);
  const double log2e = 1.4426950408889634;
  for (int m = 0; m < n_models; ++m) {
    const FbHostGmm &h = ctx->host_gmm[m];
    for (int c = 0; c < C; ++c) {
      gc2[(size_t)m * C + c] = (float)((double)h.gconsts[c] * log2e);
      gcn[(size_t)m * C + c] = h.gconsts[c];
      const int st = c / FB_STAGE_N, cc = c % FB_STAGE_N;
      const size_t stage_base = ((size_t)m * n_stage + st) * stage_halfs;
      {
        // gconst * log2(e) as three fp16 terms in the extra hi slab (multiplied by the ones slab of A)
        double gv = (double)h.gconsts[c] * log2e;
        if (!(gv > -60000.0)) gv = -60000.0;         // zero-weight components: exp2 underflows to 0 anyway
        if (gv > 60000.0) gv = 60000.0;
        const __half g0 = __float2half_rn((float)gv);
        const double r1 = gv - (double)__half2float(g0);
        const __half g1 = __float2half_rn((float)r1);
        const __half g2 = __float2half_rn((float)(r1 - (double)__half2float(g1)));
        const size_t gb = stage_base + (size_t)FB_KSLABS * (FB_STAGE_N * 8) + cc * 8;
        img[gb] = g0; img[gb + 1] = g1; img[gb + 2] = g2;
      }
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
        const size_t b = stage_base + (size_t)slab * (FB_STAGE_N * 8) + cc * 8 + e;
        img[b] = hi;
        img[b + (size_t)FB_W_HI_SLABS * FB_STAGE_N * 8] = lo;
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
    FB_CUDA(cudaFuncSetAttribute(gmm_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
    FB_CUDA(cudaFuncSetAttribute(gmm_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
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
    a.ll_out = nullptr;
    gmm_umma_kernel<false><<<grid, GMM_THREADS, kSmemLaunch, ctx->stream>>>(a);
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

// Gaussian-selection pass of the i-vector path: slot 0 only, raw component log-likelihoods to ll_out [rows_cap][C].
int fb_run_gmm_store(fb_ctx *ctx, float *ll_out, const int *done_flag) {
  FB_CHECK_ARG(ctx->n_models >= 1, "no GMMs loaded");
  const int nch = ctx->C / FB_CHUNK_N;
  GmmArgs a;
  a.a_img = ctx->a_img.p;
  a.w_img = ctx->w_img.p;
  a.part = nullptr;
  a.misc = ctx->misc.p;
  a.done_flag = done_flag;
  a.ll_out = ll_out;
  a.n_models = 1;
  a.C = ctx->C;
  a.rows_cap = ctx->rows_cap;
  const long long max_units = (long long)fb_div_up(ctx->total_frames, 2 * FB_TILE_M) * nch;
  int grid = ctx->num_sms;
  if (max_units < grid) grid = (int)max_units;
  gmm_umma_kernel<true><<<grid, GMM_THREADS, kSmemLaunch, ctx->stream>>>(a);
  fb_prof_mark(ctx, 4);
  ctx->launches += 1;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}
