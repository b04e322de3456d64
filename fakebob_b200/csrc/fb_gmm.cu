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
// Kernel shape (cta_group::1, persistent, 1 CTA / SM, 608 threads).  Measured on B200 (profiles/r01_*, scripts/*_probe.cu):
//  * with both operands in shared memory the N=64 MMAs need 192 B/clk of smem reads and run at 48-68 clk instead of 32,
//    so the A operand lives in TENSOR MEMORY;
//  * the tensor pipe's issue queue is shallow: every mbarrier wait one thread does between two 15-MMA jobs idles the
//    pipe ~130 clk (480 -> 700+ clk per job), while two issuing warps feeding different accumulators reach exactly
//    32 clk/MMA whatever they wait for -> TWO issuer warps, one per 128-row tile;
//  * tcgen05.ld moves 190-310 B/clk/SM and does not slow the MMAs down; a bulk copy of <= 40 KB takes ~970 clk.
//   warp 0   bulk-copy (TMA engine, cp.async.bulk) producer into a 5-slot x 40 KB ring: per 256-row super-tile four
//            A entries (tile0 hi, tile0 lo, tile1 hi, tile1 lo), then the W stages (general: one 40 KB stage of 64
//            columns, hi+lo; shared-variance mode: 20 KB sub-stages, two per ring entry).  Operand images are stored in
//            global memory already in the UMMA canonical no-swizzle K-major core-matrix order, so every copy is one
//            contiguous transfer.
//   warp 1, warp 2    issuers (one elected lane each) for tile 0 / tile 1: tcgen05.cp moves their tile's A entries
//            smem -> TMEM (304 columns: 2 tiles x (hi 80 + lo 72)), then per job 15 (shared mode) or 30 tcgen05.mma
//            M128 N64 K16 with A from TMEM and B from the ring slot.  Jobs are numbered globally (2 * step + tile) and use
//            accumulator job % 3 (three 64-column accumulators); a ring slot is released by one arrival per issuer.
//            Barriers between issuers and epilogue are per (accumulator, tile) so that every barrier has one waiter that
//            sees all of its phases (a parity wait cannot tell phase n from n+2).
//   warps 3-18 epilogue (8 per tile: TMEM lane quadrant = warp % 4, two warps per quadrant each taking 32 of the 64
//            accumulator columns): tcgen05.ld 32x32b.x32, online max / sum of ex2 in groups of 8 values (packed FADD2 /
//            EX2 on the groups that can still contribute) with Kaldi's log(FLT_EPSILON) pruning; gconst is already inside
//            the accumulator (extra k-block of the hi.hi part: "ones" columns of A times [g_hi g_mid g_lo] rows of W).
//            The per-slot running (max, sum) live in registers for the reference's slot counts (template kNM = 2 / 5 / 6).
// Work = the sequence of 64-column stages ordered (super-tile, [model,] stage), cut into 148 equal contiguous ranges; a CTA
// writes one (max, sum) partial per row and segment (run of its stages inside one super-tile), gmm_frame_kernel recomputes
// the cut points to merge them.
// Shared-variance mode (all slots share their inverse variances, as MAP mean-only adaptation produces): per (rows, 64-column
// stage) the ring carries  q = 0: x^2 . (-0.5/var),  q = 1: x . w_0 + g_0  (slot 0, normally the UBM),
// q = 1 + m: x . (w_m - w_0) + (g_m - g_0)  for every other slot.  The issuers accumulate q = 0 and q = 1 into ONE accumulator
// (ll_0, 30 MMAs); each difference sub-stage D_m gets its own (5 MMAs per product term) and the epilogue forms ll_m = ll_0 + D_m.
// ll_0 always uses the three-term split; the difference sub-stages use `delta_terms` of the three products
// (1: hi.hi, 2: hi.hi + hi.lo, 3: all) -- the differences are small, so their fp16 rounding error is small in absolute
// terms (scripts/gmm_precision_study.py; fb_set_gmm_delta_terms in the header).
// Round-2 measurements (profiles/r02_gmm_variants_ab.txt): with the difference sub-stages the kernel is EPILOGUE bound
// (123 us vs 72 us with the log-sum-exp compiled out), hence sixteen epilogue warps, slot 0 merged into one accumulator
// and the packed log-sum-exp: 129 -> ~101-107 us at C2 (tensor pipe 61 % busy).  Tried and not paying (same-box A/B): five
// accumulators, 4-column LSE groups, suspended (parked) barrier waits for producer / issuers, programmatic dependent
// launch at S = 50, kNM register-resident slot state (+-0).  In the general (non-shared) mode of round 1: four
// accumulators with the x^2-lo operand in shared memory and a 4-slot ring (+7 %).
#include "fb_common.cuh"
#include <math.h>
#include <atomic>

#ifndef GMM_PARTS
#define GMM_PARTS 3
#endif
// Warps: 0 = bulk-copy producer, 1 / 2 = MMA issuers (tile 0 / tile 1), 3.. = epilogue.  FB_GMM_EPI_HALVES = 2: sixteen
// epilogue warps, each owning 32 lanes x 32 of the 64 accumulator columns (the kernel is epilogue-bound once the speaker
// slots cost 5 MMAs per job: measured 123 us with eight 64-column warps vs 72 us with the log-sum-exp compiled out; the
// per-job work is one dependent chain -- tcgen05.ld, add, max tree, votes, exponentials -- so more warps per scheduler
// hide its latency).  1: eight warps x 64 columns.
#define GMM_EPI_WARP0 3
#define GMM_EPI_COLS (FB_STAGE_N / FB_GMM_EPI_HALVES)
#define GMM_THREADS (32 * (GMM_EPI_WARP0 + 8 * FB_GMM_EPI_HALVES))
#define GMM_ISSUER1 2
// Operand K layout (slabs of 8 fp16) for the hi and lo halves of A and W alike:
//   [x 0..8][ones | gconst 9][x^2 10..18][zero 19]   (20 slabs = 10 k-blocks of K=16; each half is 5 k-blocks)
// gconst*log2(e) sits in W's slab 9 as three fp16 terms (g_hi, g_mid, g_lo) against three 1.0 columns of A's "ones" slab,
// so it is added inside the MMA and the epilogue never touches it.
static constexpr uint32_t kSlabA = FB_TILE_M * 16;                             // 2048 B
static constexpr uint32_t kSlabW = FB_STAGE_N * 16;                            // 1024 B
static constexpr uint32_t kAHalfBytes = FB_A_HI_SLABS * kSlabA;                // 40960 (hi or lo of one tile)
static constexpr uint32_t kATileBytes = 2 * kAHalfBytes;                       // 81920
static constexpr uint32_t kWHalfBytes = FB_W_HI_SLABS * kSlabW;                // 20480 (hi or lo of one general stage)
static constexpr uint32_t kWStageBytes = 2 * kWHalfBytes;                      // 40960
static constexpr uint32_t kWShHalfBytes = 10 * kSlabW;                         // 10240 (hi or lo of one shared-mode sub-stage)
static constexpr uint32_t kWShBytes = 2 * kWShHalfBytes;                       // 20480
static constexpr uint32_t kSlotBytes = 40960;
static constexpr uint32_t kNumSlots = 5;
static constexpr uint32_t kSmemBar = kNumSlots * kSlotBytes;                   // 204800
static constexpr uint32_t kSmemTotal = kSmemBar + 256;
static constexpr uint32_t kSmemLaunch = kSmemTotal + 128;                      // 128 B alignment slack
static_assert(kAHalfBytes <= kSlotBytes && kWStageBytes <= kSlotBytes, "ring slot too small");
static_assert(kSmemLaunch <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
// TMEM columns (all 512 used): three 64-column accumulators [0,192) used as a ring over (tile, sub-step) jobs, so the
// MMA <-> epilogue hand-shake latency is hidden behind two jobs; A operand: tile t at 192 + 160 t: hi 80 columns, lo 80 columns
static constexpr uint32_t kTmemAcc = 0;
static constexpr uint32_t kNumAcc = 3;
static constexpr uint32_t kTmemA = 192;
static constexpr uint32_t kTmemAHalfCols = FB_A_HI_SLABS * 4;                  // 80
static constexpr uint32_t kTmemATileCols = 2 * kTmemAHalfCols;                 // 160
static constexpr uint32_t kTmemX2Cols = FB_SLAB_X2 * 4;                        // 40: column offset of the x^2 half

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
// Waiting warps must not steal issue slots from the epilogue warps of their scheduler (ncu: 45.8 M warp instructions per
// launch, issue slots 47 % busy, while the epilogue -- the bottleneck -- needs ~26 M): the suspend-time hint lets the
// hardware park the warp until the phase completes or the hint expires instead of re-issuing try_wait in a tight loop.
__device__ __forceinline__ void mbar_wait_parked(uint32_t bar, uint32_t parity, uint32_t hint_ns) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity), "r"(hint_ns) : "memory");
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
  float2 *part;             // [model][C/64][column range][rows_cap]: (max, sum) of the segment of 64-column stages that STARTS at that stage
  const int *misc;          // misc[2] = total voiced rows
  const int *done_flag;
  float *ll_out;            // STORE mode: [rows_cap][C] natural-log component log-likelihoods (Gaussian selection)
  int n_models, C, rows_cap;
  int delta_terms;          // shared mode: fp16 product terms of the difference sub-stages (1, 2 or 3)
};

// Shared-mode ring entries of one 64-column stage: entry 0 = {q = 0, q = 1} (2 x 20 KB: hi + lo), then the difference
// sub-stages in groups that fill a 40 KB slot: 4 per entry when only their hi half is stored (delta_terms == 1, 10 KB each),
// else 2 per entry.
struct SharedLayout {
  uint32_t dbytes;          // bytes of one difference sub-stage in the W image
  uint32_t group;           // difference sub-stages per ring entry
  uint32_t stage_bytes;     // bytes of one stage of the W image
  uint32_t mask;            // product terms of a difference sub-stage: bit 0 hi.hi, bit 1 lo.hi, bit 2 hi.lo
};
__host__ __device__ __forceinline__ SharedLayout shared_layout(int n_models, int delta_terms) {
  SharedLayout L;
  L.dbytes = (delta_terms == 1) ? 10240u : 20480u;
  L.group = 40960u / L.dbytes;
  L.stage_bytes = 40960u + (uint32_t)(n_models - 1) * L.dbytes;
  L.mask = (delta_terms == 1) ? 1u : (delta_terms == 2 ? 5u : 7u);
  return L;
}

// 64 accumulator columns of one row: online max / sum of 2^(v - max) with Kaldi's cutoff (LogSumExp drops terms below
// max + log(FLT_EPSILON) = max - 23 in log2).  Only ~0.5 % of the terms survive that cutoff, so the exponentials are
// evaluated per 8-column group and only when a warp vote finds a lane that still needs them (measured on the C2 model:
// ~75 % of the warp x 8-column groups are dead).  All eight votes are taken before the first exponential so their
// latencies overlap.  `m` is the running row maximum, `s` the sum relative to it.
#ifndef GMM_LSE_GROUP
#define GMM_LSE_GROUP 8
#endif
template <int NC>
__device__ __forceinline__ void lse_stage(const float *v, float &m, float &s) {
  constexpr int GS = GMM_LSE_GROUP;
  constexpr int NG = NC / GS;
  float gmax[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    const float *q = v + GS * k;
    if constexpr (GS == 8) gmax[k] = max3(max3(q[0], q[1], q[2]), max3(q[3], q[4], q[5]), fmaxf(q[6], q[7]));
    else gmax[k] = fmaxf(max3(q[0], q[1], q[2]), q[3]);
  }
  float cmax = gmax[0];
#pragma unroll
  for (int k = 1; k + 1 < NG; k += 2) cmax = max3(cmax, gmax[k], gmax[k + 1]);
  if constexpr ((NG & 1) == 0) cmax = fmaxf(cmax, gmax[NG - 1]);
  if (cmax > m) {
    s *= ex2_approx(m - cmax);
    m = cmax;
  }
  const float thr = m - 23.0f;
  bool alive[NG];
#pragma unroll
  for (int k = 0; k < NG; ++k) alive[k] = __any_sync(0xffffffffu, gmax[k] >= thr);
#pragma unroll
  for (int k = 0; k < NG; ++k) {
    if (alive[k]) {
      // Kaldi's cutoff is applied per group: a group all of whose terms lie below max - 23 is skipped; inside a surviving
      // group every term is added (those below the cutoff contribute < 2^-23 of the largest term each), which keeps the
      // arithmetic packed: 4 FADD2, 8 EX2, 3 FADD2 + 1 FADD per group of 8
      const float *q = v + GS * k;
      const float2 nm = make_float2(-m, -m);
      float2 acc2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < GS; i += 2) {
        const float2 t = __fadd2_rn(make_float2(q[i], q[i + 1]), nm);
        acc2 = __fadd2_rn(acc2, make_float2(ex2_approx(t.x), ex2_approx(t.y)));
      }
      s += acc2.x + acc2.y;
    }
  }
}

// kShared: all models share the inverse variances (MAP mean-only adaptation, build_spk_models.py:170), so per
// (super-tile, 64-column stage) the x^2 contraction  Q = x^2 . (-0.5/var)  is computed ONCE, read by the epilogue into
// registers, and every model only needs  T_m = x . (mu_m/var) + g_m  (15 MMAs instead of 30); the epilogue forms
// ll_m = Q + T_m.  Every accumulator starts from zero, so the (truncating) tensor-core accumulation error is the same for
// all models and cancels in log-likelihood-ratio scores.
#ifdef GMM_STATS
// [cta][16]: 0 entry globaltimer, 1 exit-entry ns, 2 setup clk, 3 total clk, 4 producer total, 5 producer wait empty,
// 6/7 issuer0 total / waits, 8/9 issuer1 total / waits, 10/11 epilogue warp 2 total / wait full, 12/13 epilogue warp 6
__device__ long long g_gmm_stats[148 * 16];
extern "C" int fb_debug_gmm_stats(long long *out_host) {
  return cudaMemcpyFromSymbol(out_host, g_gmm_stats, sizeof(g_gmm_stats)) == cudaSuccess ? 0 : -1;
}
#define STAT_DECL(x) long long x = 0
#define STAT_WAIT(acc, bar, par) do { const long long t_ = clock64(); mbar_wait(bar, par); acc += clock64() - t_; } while (0)
#define STAT_WAIT_PARKED(acc, bar, par, ns) do { const long long t_ = clock64(); mbar_wait_parked(bar, par, ns); acc += clock64() - t_; } while (0)
#else
#define STAT_DECL(x)
#define STAT_WAIT(acc, bar, par) mbar_wait(bar, par)
#define STAT_WAIT_PARKED(acc, bar, par, ns) mbar_wait_parked(bar, par, ns)
#endif
#ifndef GMM_PARK_PRODUCER_NS
#define GMM_PARK_PRODUCER_NS 2000
#endif
#ifndef GMM_PARK_ISSUER_NS
#define GMM_PARK_ISSUER_NS 400
#endif
// kNM > 0: the number of slots is known at compile time (shared mode), so the epilogue's per-slot running (max, sum) live in
// registers and the slot loop is unrolled; kNM = 0: run-time slot count, the state sits in local memory.
template <bool kStore, bool kShared, int kNM = 0>
__global__ void __launch_bounds__(GMM_THREADS, 1) gmm_umma_kernel(GmmArgs g) {
#ifdef GMM_STATS
  unsigned long long gt_entry;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_entry));
  const long long ck_entry = clock64();
#endif
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 127u) & ~127u;
  uint8_t *smem = smem_raw + (base - raw_addr);
  const uint32_t bar0 = base + kSmemBar;
  const uint32_t bar_full = bar0, bar_empty = bar0 + 8 * kNumSlots;           // ring slots
  // accumulator barriers are per (accumulator, tile): full[a][t] issuer t -> epilogue t, empty[a][t] epilogue t -> issuer 1-t
  // (the next user of accumulator a).  One barrier per accumulator would be waited on by the two issuers / epilogue groups
  // alternately, each skipping every other phase, which a parity wait cannot tell apart.
  const uint32_t bar_acc_full = bar0 + 16 * kNumSlots, bar_acc_empty = bar_acc_full + 16 * kNumAcc;   // [kNumAcc][2] each
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(smem + kSmemBar + 16 * kNumSlots + 32 * kNumAcc + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nst = g.C / FB_STAGE_N;
  // barrier init and the TMEM allocation touch no global memory: they run while the previous kernel drains
  if (warp == 0 && lane == 0) {
    for (uint32_t i = 0; i < kNumSlots; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 2);          // one arrive per issuer warp
    }
    for (uint32_t i = 0; i < 2 * kNumAcc; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 4 * FB_GMM_EPI_HALVES);     // one arrive per epilogue warp of the tile that read it
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32((const void *)tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  FB_GRID_DEP_SYNC();
  const bool skip = g.done_flag && *g.done_flag;     // NES early stop: fall through with no work (TMEM must still be freed)
  const int M = skip ? 0 : g.misc[2];
  const int n_super = (M + 2 * FB_TILE_M - 1) / (2 * FB_TILE_M);
  // Work = the sequence of 64-column stages ordered (super-tile, [model,] stage); general mode: one model per stage,
  // shared mode: all models inside a stage.  CTA b takes the contiguous range [total b / grid, total (b+1) / grid)
  // (gmm_frame_kernel recomputes these boundaries to find the partials of a row).
  const int models_per_unit = kShared ? 1 : g.n_models;
  const long long n_units = (long long)n_super * models_per_unit * nst;
  const int u0 = (int)(n_units * blockIdx.x / gridDim.x);
  const int u1 = (int)(n_units * (blockIdx.x + 1) / gridDim.x);
  const int n_sub = kShared ? g.n_models : 1;                                  // jobs per tile per 64-column stage
  const SharedLayout SL = shared_layout(g.n_models, g.delta_terms);

  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
#ifdef GMM_STATS
  const long long ck_setup = clock64();
#endif

  if (warp == 0) {
    // ================= producer: ring of kNumSlots slots (whole warp loops, one elected lane issues) =================
    int cur_super = -1;
    uint32_t cnt = 0;
    STAT_DECL(st_prod_empty);
#ifdef GMM_STATS
    const long long st_t0 = clock64();
#endif
    for (int u = u0; u < u1; ++u) {
      const size_t stage = u % nst;
      const int item = u / nst;
      const int model = item % models_per_unit;
      const int sp = item / models_per_unit;
      if (sp != cur_super) {
        const uint8_t *src = reinterpret_cast<const uint8_t *>(g.a_img) + (size_t)sp * 2 * kATileBytes;
#pragma unroll 1
        for (int e = 0; e < 4; ++e) {          // tile0 hi, tile0 lo, tile1 hi, tile1 lo
          const uint32_t slot = cnt % kNumSlots;
          STAT_WAIT_PARKED(st_prod_empty, bar_empty + 8 * slot, ((cnt / kNumSlots) & 1) ^ 1, GMM_PARK_PRODUCER_NS);
          if (elect_one()) {
            mbar_expect_tx(bar_full + 8 * slot, kAHalfBytes);
            bulk_g2s(base + slot * kSlotBytes, src + (size_t)e * kAHalfBytes, kAHalfBytes, bar_full + 8 * slot);
          }
          __syncwarp();
          ++cnt;
        }
        cur_super = sp;
      }
      if constexpr (kShared) {
        // a bulk copy costs ~930 clk whatever its size (<= 40 KB) and copies of one SM do not overlap, so the sub-stages
        // are fetched in groups that fill a slot: {x^2, slot 0}, then `group` difference sub-stages per entry
        const uint8_t *wst = reinterpret_cast<const uint8_t *>(g.w_img) + stage * (size_t)SL.stage_bytes;
        const int n_delta = g.n_models - 1;
#pragma unroll 1
        for (int e = -1; e * (int)SL.group < n_delta; ++e) {
          const uint32_t slot = cnt % kNumSlots;
          STAT_WAIT_PARKED(st_prod_empty, bar_empty + 8 * slot, ((cnt / kNumSlots) & 1) ^ 1, GMM_PARK_PRODUCER_NS);
          if (elect_one()) {
            uint32_t bytes, off;
            if (e < 0) { bytes = 40960u; off = 0; }
            else {
              const int first = e * (int)SL.group;
              const int n = min((int)SL.group, n_delta - first);
              bytes = (uint32_t)n * SL.dbytes;
              off = 40960u + (uint32_t)first * SL.dbytes;
            }
            mbar_expect_tx(bar_full + 8 * slot, bytes);
            bulk_g2s(base + slot * kSlotBytes, wst + off, bytes, bar_full + 8 * slot);
          }
          __syncwarp();
          ++cnt;
        }
      } else {
        const uint32_t slot = cnt % kNumSlots;
        STAT_WAIT_PARKED(st_prod_empty, bar_empty + 8 * slot, ((cnt / kNumSlots) & 1) ^ 1, GMM_PARK_PRODUCER_NS);
        if (elect_one()) {
          const size_t off = ((size_t)model * (g.C / FB_STAGE_N) + stage) * (size_t)kWStageBytes;
          mbar_expect_tx(bar_full + 8 * slot, kWStageBytes);
          bulk_g2s(base + slot * kSlotBytes, reinterpret_cast<const uint8_t *>(g.w_img) + off, kWStageBytes, bar_full + 8 * slot);
        }
        __syncwarp();
        ++cnt;
      }
    }
#ifdef GMM_STATS
    if (lane == 0) { g_gmm_stats[blockIdx.x * 16 + 4] = clock64() - st_t0; g_gmm_stats[blockIdx.x * 16 + 5] = st_prod_empty; }
#endif
  } else if (warp == 1 || warp == GMM_ISSUER1) {
    // ===== tcgen05 issuers: warp 1 owns tile 0, warp GMM_ISSUER1 owns tile 1 (whole warp loops and waits; one elected lane
    // issues copies / MMAs / commits).  Two issuers because the tensor pipe's issue queue is shallow (measured,
    // scripts/mma_probe.cu: every mbarrier wait between two 15-MMA jobs of ONE thread idles the pipe ~130 clk, 480 -> 700+
    // clk per job); while one issuer is in its wait / commit overhead the other one's MMAs keep the pipe busy.
    // Both walk the same sequence of ring entries; a slot is released by two arrivals (one per issuer).
    const int tile = (warp == 1) ? 0 : 1;
    // smem descriptor with LBO / SBO / version bits and a zero start address; only the 14-bit address field varies
    const uint64_t desc_w = make_desc(0, kSlabW, 128);
    const uint64_t desc_a = make_desc(0, kSlabA, 128);
    int cur_super = -1;
    uint32_t cnt = 0, job = tile;                   // job = global (sub-step, tile) counter = 2 * step + tile; accumulator = job % 3
    STAT_DECL(st_mma_full); STAT_DECL(st_mma_acc); STAT_DECL(st_mma_afull);
#ifdef GMM_STATS
    const long long st_t0 = clock64();
#endif
    for (int u = u0; u < u1; ++u) {
      const int item = u / nst;
      const int sp = item / models_per_unit;
      if (sp != cur_super) {
#pragma unroll 1
        for (int e = 0; e < 4; ++e) {
          const uint32_t slot = cnt % kNumSlots;
          // the other tile's entries are waited for as well: "full" proves the slot's previous use was released, so the
          // plain arrive below cannot land in the previous phase of the empty barrier
          STAT_WAIT(st_mma_afull, bar_full + 8 * slot, (cnt / kNumSlots) & 1);
          tc_fence_after();
          if (elect_one()) {
            if ((e >> 1) == tile) {
              const uint64_t src = desc_a + (uint64_t)(((base + slot * kSlotBytes) & 0x3FFFFu) >> 4);
              const uint32_t dst = tmem_base + kTmemA + tile * kTmemATileCols + (e & 1) * kTmemAHalfCols;
#pragma unroll
              for (int kb = 0; kb < FB_A_HI_SLABS / 2; ++kb) tc_cp_128x256b(dst + kb * 8, src + (uint64_t)(kb * ((2 * kSlabA) >> 4)));
              tc_commit(bar_empty + 8 * slot);
            } else {
              mbar_arrive(bar_empty + 8 * slot);
            }
          }
          __syncwarp();
          ++cnt;
        }
        cur_super = sp;
      }
      {
#pragma unroll 1
        for (int q = 0; q < n_sub; ++q) {
          const uint32_t slot = cnt % kNumSlots;
          // shared mode: job 0 = slot 0 complete (x^2 and x sub-blocks of ring entry {0, 1}, one accumulator), job q >= 1 = the
          // difference sub-stage of slot q, in groups of SL.group per ring entry
          bool first_in_slot = true, last_in_slot = true;
          uint32_t sub_off = 0, mask = 7u;
          if constexpr (kShared) {
            if (q >= 1) {
              const uint32_t j = (uint32_t)(q - 1), r = j % SL.group;
              first_in_slot = r == 0; last_in_slot = r == SL.group - 1 || q + 1 == n_sub;
              sub_off = r * SL.dbytes; mask = SL.mask;
            }
          }
          if (first_in_slot) STAT_WAIT_PARKED(st_mma_full, bar_full + 8 * slot, (cnt / kNumSlots) & 1, GMM_PARK_ISSUER_NS);
          const uint64_t w_hi = desc_w + (uint64_t)(((base + slot * kSlotBytes + sub_off) & 0x3FFFFu) >> 4);
          const uint32_t abuf = job % kNumAcc;
          // accumulator abuf was last used by job - 3, a job of the other tile: wait until that tile's epilogue has read it
          if (job >= kNumAcc) STAT_WAIT_PARKED(st_mma_acc, bar_acc_empty + 8 * (2 * abuf + (tile ^ 1)), ((job - kNumAcc) / (2 * kNumAcc)) & 1, GMM_PARK_ISSUER_NS);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_hi = tmem_base + kTmemA + tile * kTmemATileCols;
            const uint32_t a_lo = a_hi + kTmemAHalfCols;
            const uint32_t d_tmem = tmem_base + kTmemAcc + abuf * FB_STAGE_N;
            if constexpr (kShared) {
              if (q == 0) {
                // ll_0 = x^2 . (-0.5/var) + x . w_0 + g_0 in one accumulator: 3 terms x (5 + 5) k-blocks.  Ring entry:
                // [x^2 block: hi 10 slabs | lo 10 slabs][x block: hi | lo]; A: x at column 0, x^2 at kTmemX2Cols.
                const uint64_t q_hi = w_hi, q_lo = w_hi + (uint64_t)(kWShHalfBytes >> 4);
                const uint64_t t_hi = w_hi + (uint64_t)(kWShBytes >> 4), t_lo = t_hi + (uint64_t)(kWShHalfBytes >> 4);
#pragma unroll
                for (int part = 0; part < GMM_PARTS; ++part) {
                  const uint32_t a_base = (part == 1) ? a_lo : a_hi;
#pragma unroll
                  for (int kb = 0; kb < 5; ++kb)
                    tc_mma_f16_ta(d_tmem, a_base + kb * 8, ((part == 2) ? t_lo : t_hi) + (uint64_t)(kb * ((2 * kSlabW) >> 4)), kIdesc, (part | kb) ? 1u : 0u);
#pragma unroll
                  for (int kb = 0; kb < 5; ++kb)
                    tc_mma_f16_ta(d_tmem, a_base + kTmemX2Cols + kb * 8, ((part == 2) ? q_lo : q_hi) + (uint64_t)(kb * ((2 * kSlabW) >> 4)), kIdesc, 1u);
                }
              } else {
                const uint64_t w_lo = w_hi + (uint64_t)(kWShHalfBytes >> 4);
#pragma unroll
                for (int part = 0; part < GMM_PARTS; ++part) {
                  if (!((mask >> part) & 1u)) continue;
                  const uint32_t a_base = (part == 1) ? a_lo : a_hi;
                  const uint64_t b_base = (part == 2) ? w_lo : w_hi;
#pragma unroll
                  for (int kb = 0; kb < 5; ++kb)
                    tc_mma_f16_ta(d_tmem, a_base + kb * 8, b_base + (uint64_t)(kb * ((2 * kSlabW) >> 4)), kIdesc, (part | kb) ? 1u : 0u);
                }
              }
            } else {
              const uint64_t w_lo = w_hi + (uint64_t)(kWHalfBytes >> 4);
#pragma unroll
              for (int part = 0; part < GMM_PARTS; ++part) {
                const uint32_t a_base = (part == 1) ? a_lo : a_hi;
                const uint64_t b_base = (part == 2) ? w_lo : w_hi;
#pragma unroll
                for (int kb = 0; kb < 10; ++kb)
                  tc_mma_f16_ta(d_tmem, a_base + kb * 8, b_base + (uint64_t)(kb * ((2 * kSlabW) >> 4)), kIdesc, (part | kb) ? 1u : 0u);
              }
            }
            tc_commit(bar_acc_full + 8 * (2 * abuf + tile));
            if (last_in_slot) tc_commit(bar_empty + 8 * slot);
          }
          __syncwarp();
          job += 2;
          if (last_in_slot) ++cnt;
        }
      }
    }
#ifdef GMM_STATS
    if (lane == 0) { g_gmm_stats[blockIdx.x * 16 + 6 + 2 * tile] = clock64() - st_t0; g_gmm_stats[blockIdx.x * 16 + 7 + 2 * tile] = st_mma_full + st_mma_afull + st_mma_acc; }
#endif
  } else {
    // ===== epilogue: warps 3..; TMEM lane quadrant = warp % 4; (warp - 3) / 4 = 2 * half + tile: an accumulator is read by
    // 4 lane quadrants x FB_GMM_EPI_HALVES column ranges =====
    constexpr int NC = GMM_EPI_COLS;
    const int quad = warp & 3;
    const int tile = ((warp - GMM_EPI_WARP0) >> 2) & 1;
    const int half = (warp - GMM_EPI_WARP0) >> 3;
    uint32_t job = tile;                               // this tile's jobs are tile, tile + 2, tile + 4, ...
    const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + kTmemAcc + half * NC;
    constexpr int kState = kShared ? (kNM > 0 ? kNM : FB_MAX_MODELS) : 1;
    float mm[kState], ss[kState];
    int run_item = -1;                                  // the running maxima belong to this (super[, model])
    STAT_DECL(st_epi_full);
#ifdef GMM_STATS
    const long long st_t0 = clock64();
#endif
    // wait for this tile's next accumulator, pull this warp's 32 lanes x NC columns into registers, hand it back
    auto fetch = [&](float (&v)[NC]) {
      const uint32_t abuf = job % kNumAcc;
      STAT_WAIT(st_epi_full, bar_acc_full + 8 * (2 * abuf + tile), (job / (2 * kNumAcc)) & 1);
      tc_fence_after();
      const uint32_t taddr = taddr0 + abuf * FB_STAGE_N;
#ifndef GMM_NO_LDTM
      tc_ld32(taddr, v);
      if constexpr (NC == 64) tc_ld32(taddr + 32, v + 32);
      tc_wait_ld();
#else
#pragma unroll
      for (int i = 0; i < NC; ++i) v[i] = (float)(i + job);
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * (2 * abuf + tile));   // accumulator is in registers: the next job may overwrite
      job += 2;
    };
    // The running (max, sum) of a row is kept over the whole run of consecutive stages this CTA computes for the same
    // (rows[, model]) -- a "segment" -- so the cutoff is relative to the running maximum; one partial per segment and column
    // range, stored at the index of the segment's first stage.
    const int n_run = kShared ? (kNM > 0 ? kNM : g.n_models) : 1;
    int seg_start = 0, seg_row = 0, seg_model = 0;
    auto flush = [&]() {
      if constexpr (!kStore) {
        if constexpr (kNM > 0) {
#pragma unroll
          for (int r = 0; r < kNM; ++r)
            g.part[(((size_t)r * nst + seg_start) * FB_GMM_EPI_HALVES + half) * g.rows_cap + seg_row] = make_float2(mm[r], ss[r]);
        } else {
          for (int r = 0; r < n_run; ++r) {
            const int mdl = kShared ? r : seg_model;
            g.part[(((size_t)mdl * nst + seg_start) * FB_GMM_EPI_HALVES + half) * g.rows_cap + seg_row] = make_float2(mm[r], ss[r]);
          }
        }
      }
    };
    for (int u = u0; u < u1; ++u) {
      const int stage = u % nst;
      const int item = u / nst;
      const int model = item % models_per_unit;
      const int sp = item / models_per_unit;
      const int row = sp * (2 * FB_TILE_M) + tile * FB_TILE_M + quad * 32 + lane;
      if (item != run_item) {
        if (run_item >= 0) flush();
        if constexpr (kNM > 0) {
#pragma unroll
          for (int i = 0; i < kNM; ++i) { mm[i] = -INFINITY; ss[i] = 0.f; }
        } else {
          for (int i = 0; i < n_run; ++i) { mm[i] = -INFINITY; ss[i] = 0.f; }
        }
        run_item = item;
        seg_start = stage; seg_row = row; seg_model = model;
      }
      if constexpr (kShared) {
        float q[NC];                                    // ll_0 of this (rows, column range): the base every other slot adds to
        fetch(q);
        {
          float m = mm[0], sacc = ss[0];
#ifndef GMM_NO_LSE
          lse_stage<NC>(q, m, sacc);
#else
          m = fmaxf(m, q[0] + q[NC - 1]);
#endif
          mm[0] = m;
          ss[0] = sacc;
        }
        auto slot_job = [&](float &m_r, float &s_r) {
          float v[NC];
          fetch(v);
#pragma unroll
          for (int i = 0; i < NC; i += 2) {
            const float2 t = __fadd2_rn(make_float2(v[i], v[i + 1]), make_float2(q[i], q[i + 1]));
            v[i] = t.x; v[i + 1] = t.y;
          }
          float m = m_r, sacc = s_r;
#ifndef GMM_NO_LSE
          lse_stage<NC>(v, m, sacc);
#else
          m = fmaxf(m, v[0] + v[NC - 1]);
#endif
          m_r = m;
          s_r = sacc;
        };
        if constexpr (kNM > 0) {
#pragma unroll
          for (int r = 1; r < kNM; ++r) slot_job(mm[r], ss[r]);
        } else {
#pragma unroll 1
          for (int r = 1; r < g.n_models; ++r) slot_job(mm[r], ss[r]);
        }
      } else {
        float v[NC];
        fetch(v);
        if constexpr (kStore) {
          float4 *dst = reinterpret_cast<float4 *>(g.ll_out + (size_t)row * g.C + stage * FB_STAGE_N + half * NC);
          const float ln2 = 0.6931471805599453f;
#pragma unroll
          for (int i = 0; i < NC; i += 4) dst[i >> 2] = make_float4(v[i] * ln2, v[i + 1] * ln2, v[i + 2] * ln2, v[i + 3] * ln2);
        } else {
          float m = mm[0], sacc = ss[0];
#ifndef GMM_NO_LSE
          lse_stage<NC>(v, m, sacc);
#else
          m = fmaxf(m, v[0] + v[NC - 1]);
#endif
          mm[0] = m;
          ss[0] = sacc;
        }
      }
    }
    if (run_item >= 0) flush();
#ifdef GMM_STATS
    if (lane == 0 && (warp == 3 || warp == 7)) { g_gmm_stats[blockIdx.x * 16 + 10 + (warp == 7 ? 2 : 0)] = clock64() - st_t0; g_gmm_stats[blockIdx.x * 16 + 11 + (warp == 7 ? 2 : 0)] = st_epi_full; }
#endif
  }
  tc_fence_before();
  __syncthreads();
#ifdef GMM_STATS
  if (threadIdx.x == 0) {
    unsigned long long gt_exit;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_exit));
    g_gmm_stats[blockIdx.x * 16 + 0] = (long long)gt_entry; g_gmm_stats[blockIdx.x * 16 + 1] = (long long)(gt_exit - gt_entry);
    g_gmm_stats[blockIdx.x * 16 + 2] = ck_setup - ck_entry; g_gmm_stats[blockIdx.x * 16 + 3] = clock64() - ck_entry;
  }
#endif
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Merge the per-chunk partials into per-frame log-likelihoods (one thread per (row, model)), then the per-utterance
// average (gmm-global-get-frame-likes --average=true: float frame values, double sum, float quotient).
// Fixed reduction order (deterministic).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gmm_frame_kernel(const float2 *__restrict__ part, const int *__restrict__ misc, float *__restrict__ frame_ll, int nst,
                 int rows_cap, const int *__restrict__ done_flag, int umma_grid, int models_per_unit) {
  FB_GRID_DEP_SYNC();
  if (done_flag && *done_flag) return;
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  const int model = blockIdx.y;
  const int M = misc[2];
  if (row >= M) return;
  // The partials of this row are the segments gmm_umma_kernel cut its stage sequence into: recompute the CTA boundaries
  // floor(total * b / grid) that fall inside this (super-tile[, model]) item.
  float2 p[64 * FB_GMM_EPI_HALVES];              // C <= 4096: at most 64 stages, hence at most 64 segments x column ranges
  int n = 0;
  {
    const int n_super = (M + 2 * FB_TILE_M - 1) / (2 * FB_TILE_M);
    const long long total = (long long)n_super * models_per_unit * nst;
    const int sp = row / (2 * FB_TILE_M);
    const long long g0 = ((long long)sp * models_per_unit + (models_per_unit > 1 ? model : 0)) * nst;
    long long st = g0;
    while (st < g0 + nst) {
      const long long b = ((st + 1) * umma_grid - 1) / total;                 // the CTA whose range contains stage st
      long long end = total * (b + 1) / umma_grid;
      if (end > g0 + nst) end = g0 + nst;
#pragma unroll
      for (int h = 0; h < FB_GMM_EPI_HALVES; ++h)
        p[n++] = part[(((size_t)model * nst + (int)(st - g0)) * FB_GMM_EPI_HALVES + h) * rows_cap + row];
      st = end;
    }
  }
  float mx = -INFINITY;
  for (int k = 0; k < n; ++k) mx = fmaxf(mx, p[k].x);
  double s = 0.0;
  for (int k = 0; k < n; ++k)
    if (p[k].x >= mx - 23.0f) s += (double)(p[k].y * exp2f(p[k].x - mx));
  frame_ll[(size_t)model * rows_cap + row] = (float)(((double)mx + log2(s)) * 0.6931471805599453);
}

__global__ void __launch_bounds__(128)
gmm_reduce_kernel(const float *__restrict__ frame_ll, const int *__restrict__ row_off, double *__restrict__ avg_ll,
                  int n_models, int rows_cap, const int *__restrict__ done_flag, int text7) {
  FB_GRID_DEP_SYNC();
  if (done_flag && *done_flag) return;
  __shared__ double s_red[4];
  const int b = blockIdx.x, model = blockIdx.y;
  const int r0 = row_off[b], r1 = row_off[b + 1];
  double acc = 0.0;
  for (int row = r0 + threadIdx.x; row < r1; row += blockDim.x) acc += (double)frame_ll[(size_t)model * rows_cap + row];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double tot = (s_red[0] + s_red[1]) + (s_red[2] + s_red[3]);
    const int n = r1 - r0;
    double a = (n > 0) ? (double)__fdiv_rn((float)tot, (float)n) : 0.0;
    if (text7) a = fb_round_sig7(a);               // the reference parses this number from 7-digit text
    avg_ll[(size_t)b * n_models + model] = a;
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
  FB_CHECK_ARG(C > 0 && C % FB_CHUNK_N == 0 && C <= 32 * FB_CHUNK_N, "number of components must be a multiple of 128, at most 4096");
  FB_CHECK_ARG(weights && means_invvars && inv_vars && gconsts, "NULL parameter array");
  FbHostGmm &h = ctx->host_gmm[slot];
  h.weights.assign(weights, weights + C);
  h.means_invvars.assign(means_invvars, means_invvars + (size_t)C * D);
  h.inv_vars.assign(inv_vars, inv_vars + (size_t)C * D);
  h.gconsts.assign(gconsts, gconsts + C);
  h.loaded = true;
  return FB_OK;
}

// fp16 hi/lo split helpers for the W image (values already scaled and multiplied by log2(e))
static inline void put_split(std::vector<__half> &img, size_t hi_idx, size_t lo_off, double w) {
  const __half hi = __float2half_rn((float)w);
  const __half lo = __float2half_rn((float)(w - (double)__half2float(hi)));
  img[hi_idx] = hi;
  img[hi_idx + lo_off] = lo;
}
static inline void put_gconst3(std::vector<__half> &img, size_t idx, double gv) {
  if (!(gv > -60000.0)) gv = -60000.0;         // zero-weight components: exp2 underflows to 0 anyway
  if (gv > 60000.0) gv = 60000.0;
  const __half g0 = __float2half_rn((float)gv);
  const double r1 = gv - (double)__half2float(g0);
  const __half g1 = __float2half_rn((float)r1);
  const __half g2 = __float2half_rn((float)(r1 - (double)__half2float(g1)));
  img[idx] = g0; img[idx + 1] = g1; img[idx + 2] = g2;
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
  // shared-variance mode: every model has exactly slot 0's inverse variances (MAP with --update-flags=m)
  bool shared = n_models >= 2 && getenv("FB_GMM_NO_SHARED") == nullptr;
  for (int m = 1; m < n_models && shared; ++m)
    shared = ctx->host_gmm[m].inv_vars == g0.inv_vars;
  ctx->gmm_shared = shared;
  const int n_stage = C / FB_STAGE_N;
  const double log2e = 1.4426950408889634;
  // scaled weights in double: w1[m][c][d] = miv / s * log2e ; w2[c][d] = -0.5 iv / s^2 * log2e ; g[m][c] = gconst * log2e
  auto w1 = [&](int m, int c, int d) {
    return (double)ctx->host_gmm[m].means_invvars[(size_t)c * FB_DIM + d] / (double)ctx->tables_host.feat_scale[d] * log2e;
  };
  auto w2 = [&](int m, int c, int d) {
    const double sc = ctx->tables_host.feat_scale[d];
    return -0.5 * (double)ctx->host_gmm[m].inv_vars[(size_t)c * FB_DIM + d] / (sc * sc) * log2e;
  };
  auto gl = [&](int m, int c) { return (double)ctx->host_gmm[m].gconsts[c] * log2e; };
  std::vector<__half> img;
  const size_t slabW = (size_t)FB_STAGE_N * 8;                 // halfs per slab
  bool range_err = false;
  if (!shared) {
    // [model][stage][hi 20 slabs | lo 20 slabs][64 cols][8]
    const size_t stage_halfs = (size_t)2 * FB_W_HI_SLABS * slabW;
    img.assign((size_t)n_models * n_stage * stage_halfs, __float2half_rn(0.f));
    for (int m = 0; m < n_models; ++m)
      for (int c = 0; c < C; ++c) {
        const int st = c / FB_STAGE_N, cc = c % FB_STAGE_N;
        const size_t sb = ((size_t)m * n_stage + st) * stage_halfs + (size_t)cc * 8;
        const size_t lo_off = (size_t)FB_W_HI_SLABS * slabW;
        for (int d = 0; d < FB_DIM; ++d) {
          const double a1 = w1(m, c, d), a2 = w2(m, c, d);
          if (!(fabs(a1) < 60000.0) || !(fabs(a2) < 60000.0)) range_err = true;
          put_split(img, sb + (size_t)(d >> 3) * slabW + (d & 7), lo_off, a1);
          put_split(img, sb + (size_t)(FB_SLAB_X2 + (d >> 3)) * slabW + (d & 7), lo_off, a2);
        }
        put_gconst3(img, sb + (size_t)FB_SLAB_ONES * slabW, gl(m, c));
      }
  } else {
    // How many of the three fp16 products the difference sub-stages need.  The rounding error of a one-term product
    // x_hi . dw_hi is ~2^-12 |x| |dw| per dimension; with |x| ~ 1 after the power-of-two scaling, the per-frame error of
    // component c is ~2^-12 ||dw_c||.  E = 2^-12 sqrt(sum_c weight_c ||dw_c||^2), maximised over the slots, predicts it
    // (measured on the B200, tests/test_gpu_gmm.py / test_gpu_fullsize.py: E = 6.0e-4 (2048-mixture bench models) -> utterance
    // score deviation from the three-term result 1.1e-4 with one term, 3.8e-5 with two; E = 1.2e-3 (256-mixture test models)
    // -> 7.2e-4 and 1.2e-4).  The automatic choice keeps the deviation near the 1e-4 resolution of the reference's own
    // 7-significant-digit text scores: two terms for MAP-adapted speaker models, one only for very small offsets, three
    // when the slots are unrelated models.
    double E = 0.0;
    for (int m = 1; m < n_models; ++m) {
      double acc = 0.0, wsum = 0.0;
      for (int c = 0; c < C; ++c) {
        double n2 = 0.0;
        for (int d = 0; d < FB_DIM; ++d) { const double dv = w1(m, c, d) - w1(0, c, d); n2 += dv * dv; }
        acc += (double)g0.weights[c] * n2;
        wsum += g0.weights[c];
      }
      const double e = ldexp(sqrt(acc / (wsum > 0 ? wsum : 1.0)), -12);
      if (e > E) E = e;
    }
    ctx->delta_err_est = E;
    int terms = ctx->delta_terms_req;
    if (terms == 0) terms = (E <= 2.5e-4) ? 1 : ((E <= 1.5e-3) ? 2 : 3);
    ctx->delta_terms = terms;
    // [stage][q = 0: x^2 part | q = 1: x part + gconst of slot 0 | q = 1+m: (slot m - slot 0) x part + gconst difference]
    // q < 2: [hi 10 slabs | lo 10 slabs][64 cols][8]; q >= 2: the same, or the hi half alone when terms == 1
    const SharedLayout SL = shared_layout(n_models, terms);
    const size_t stage_halfs = SL.stage_bytes / 2;
    img.assign((size_t)n_stage * stage_halfs, __float2half_rn(0.f));
    const size_t lo_off = (size_t)10 * slabW;
    std::vector<__half> scratch(2 * lo_off + 16);
    for (int c = 0; c < C; ++c) {
      const int st = c / FB_STAGE_N, cc = c % FB_STAGE_N;
      for (int q = 0; q <= n_models; ++q) {
        const size_t q_off = (q < 2) ? (size_t)q * 10240 : (size_t)20480 + (size_t)(q - 2) * (SL.dbytes / 2);   // in halfs
        const size_t sb = (size_t)st * stage_halfs + q_off + (size_t)cc * 8;
        const bool has_lo = q < 2 || terms >= 2;
        for (int d = 0; d < FB_DIM; ++d) {
          double v;
          if (q == 0) v = w2(0, c, d);
          else if (q == 1) v = w1(0, c, d);
          else v = w1(q - 1, c, d) - w1(0, c, d);
          if (!(fabs(v) < 60000.0)) range_err = true;
          const size_t idx = sb + (size_t)(d >> 3) * slabW + (d & 7);
          if (has_lo) put_split(img, idx, lo_off, v);
          else img[idx] = __float2half_rn((float)v);
        }
        if (q == 1) put_gconst3(img, sb + (size_t)9 * slabW, gl(0, c));
        else if (q >= 2) {
          double g_m = gl(q - 1, c), g_0 = gl(0, c);
          if (!(g_m > -60000.0)) g_m = -60000.0;
          if (!(g_0 > -60000.0)) g_0 = -60000.0;
          put_gconst3(img, sb + (size_t)9 * slabW, g_m - g_0);
        }
      }
    }
  }
  if (range_err) {
    fb_set_error("a scaled GMM weight exceeds the fp16 range (model dynamic range not supported by the split-fp16 kernel)");
    return FB_ERR_UNSUPPORTED;
  }
  int rc;
  if ((rc = ctx->w_img.ensure(img.size()))) return rc;
  FB_CUDA(cudaMemcpy(ctx->w_img.p, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
  ctx->n_models = n_models;
  ctx->C = C;
  // buffers that depend on n_models are (re)sized at the next fb_reserve_batch
  ctx->part.release();
  ctx->frame_ll.release();
  ctx->avg_ll.release();
  ctx->batch_tag = -1;
  static std::atomic<unsigned long long> attr_set_mask{0};
  if (fb_once_per_device(attr_set_mask, ctx->device)) {
    FB_CUDA(cudaFuncSetAttribute(gmm_umma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
    FB_CUDA(cudaFuncSetAttribute(gmm_umma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
    FB_CUDA(cudaFuncSetAttribute(gmm_umma_kernel<false, true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
    FB_CUDA(cudaFuncSetAttribute(gmm_umma_kernel<false, true, 5>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
    FB_CUDA(cudaFuncSetAttribute(gmm_umma_kernel<false, true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
    FB_CUDA(cudaFuncSetAttribute(gmm_umma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLaunch));
  }
  return FB_OK;
}

int fb_run_gmm_flag(fb_ctx *ctx, const int *done_flag) {
  FB_CHECK_ARG(ctx->n_models > 0, "no GMMs loaded (fb_finalize_gmms)");
  const int nst = ctx->C / FB_STAGE_N;
  int umma_grid = 0;
  FbNvtxSeq nv;
  nv.next("fb:gmm_umma");
  {
    GmmArgs a;
    a.a_img = ctx->a_img.p;
    a.w_img = ctx->w_img.p;
    a.part = ctx->part.p;
    a.misc = ctx->misc.p;
    a.done_flag = done_flag;
    a.n_models = ctx->n_models;
    a.C = ctx->C;
    a.rows_cap = ctx->rows_cap;
    a.delta_terms = ctx->delta_terms;
    // upper bound on useful CTAs: one unit each
    const long long max_units = (long long)fb_div_up(ctx->total_frames, 2 * FB_TILE_M) * (ctx->gmm_shared ? 1 : ctx->n_models) * nst;
    int grid = ctx->num_sms;
    if (max_units < grid) grid = (int)max_units;
    umma_grid = grid;
    a.ll_out = nullptr;
    if (ctx->gmm_shared) {
      // the slot counts of the reference's configurations get register-resident epilogue state: SV (UBM + 1), CSI with
      // the default 5 speakers, OSI (UBM + 5); any other count takes the general instantiation
      static const bool generic_only = getenv("FB_GMM_GENERIC_SLOTS") != nullptr;
      const int nm = generic_only ? 0 : ctx->n_models;
      if (nm == 6) FB_CUDA(fb_launch(gmm_umma_kernel<false, true, 6>, dim3(grid), dim3(GMM_THREADS), kSmemLaunch, ctx->stream, a));
      else if (nm == 5) FB_CUDA(fb_launch(gmm_umma_kernel<false, true, 5>, dim3(grid), dim3(GMM_THREADS), kSmemLaunch, ctx->stream, a));
      else if (nm == 2) FB_CUDA(fb_launch(gmm_umma_kernel<false, true, 2>, dim3(grid), dim3(GMM_THREADS), kSmemLaunch, ctx->stream, a));
      else FB_CUDA(fb_launch(gmm_umma_kernel<false, true>, dim3(grid), dim3(GMM_THREADS), kSmemLaunch, ctx->stream, a));
    } else {
      FB_CUDA(fb_launch(gmm_umma_kernel<false, false>, dim3(grid), dim3(GMM_THREADS), kSmemLaunch, ctx->stream, a));
    }
  }
  fb_prof_mark(ctx, 4);
  nv.next("fb:gmm_frame_reduce");
  FB_CUDA(fb_launch(gmm_frame_kernel, dim3(fb_div_up(ctx->total_frames, 128), ctx->n_models), dim3(128), 0, ctx->stream,
                    ctx->part.p, ctx->misc.p, ctx->frame_ll.p, nst, ctx->rows_cap, done_flag, umma_grid, ctx->gmm_shared ? 1 : ctx->n_models));
  FB_CUDA(fb_launch(gmm_reduce_kernel, dim3(ctx->B, ctx->n_models), dim3(128), 0, ctx->stream, ctx->frame_ll.p, ctx->row_off.p,
                    ctx->avg_ll.p, ctx->n_models, ctx->rows_cap, done_flag, ctx->kx_text ? 1 : 0));
  fb_prof_mark(ctx, 5);
  ctx->launches += 3;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fb_run_gmm(fb_ctx *ctx) { return fb_run_gmm_flag(ctx, nullptr); }

// Gaussian-selection pass of the i-vector path: slot 0 only, raw component log-likelihoods to ll_out [rows_cap][C].
int fb_run_gmm_store(fb_ctx *ctx, float *ll_out, const int *done_flag) {
  FB_CHECK_ARG(ctx->n_models >= 1, "no GMMs loaded");
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
  a.delta_terms = 3;
  const long long max_units = (long long)fb_div_up(ctx->total_frames, 2 * FB_TILE_M) * (ctx->C / FB_STAGE_N);
  int grid = ctx->num_sms;
  if (max_units < grid) grid = (int)max_units;
  FB_CHECK_ARG(!ctx->gmm_shared, "Gaussian selection needs the general W image");
  FbNvtxSeq nv;
  nv.next("fb:gmm_umma_store");
  FB_CUDA(fb_launch(gmm_umma_kernel<true, false>, dim3(grid), dim3(GMM_THREADS), kSmemLaunch, ctx->stream, a));
  fb_prof_mark(ctx, 4);
  ctx->launches += 1;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}
