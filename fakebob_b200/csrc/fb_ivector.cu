// i-vector extraction + PLDA scoring on the device.
//
// Replaces sid/extract_ivectors.sh and ivector-plda-scoring as the reference runs them per score() call
// (ivector_PLDA_kaldiHelper.py:202-211, 262-271); upstream arithmetic SURVEY.md Appendix A.8:
//   gselect_kernel     gmm-gselect --n=20 on the diagonalised full UBM (log-likelihoods from the tcgen05 GMM kernel in
//                      STORE mode), best-first indices
//   fgmm_post_kernel   fgmm-global-gselect-to-post --min-post=0.025: 20 full-covariance log-likelihoods per frame (float),
//                      softmax, pruning, renormalisation
//   ivec_stats_kernel  Baum-Welch statistics gamma_c, X_c per utterance (float64, frame order like Kaldi's AccStats)
//   ivec_lin_kernel    lin  = sum_c (Sigma_c^-1 M_c)^T X_c              (fp32 parameters streamed once, float64 accumulation)
//   ivec_quad_kernel   quad = sum_c gamma_c vech(M_c^T Sigma_c^-1 M_c)  (idem)
//   ivec_solve_kernel  (quad + I) w = lin + prior e_0   by Cholesky (float64), w_0 -= prior offset, stored as float
//   plda_kernel        ivector-subtract-global-mean | transform-vec | ivector-normalize-length (float), Plda::TransformIvector
//                      (normalize_length, float64), LogLikelihoodRatio against every enrolled speaker
// The derived extractor matrices (Sigma^-1 M, U) are computed ONCE at load on the device (Kaldi recomputes them on every
// ivector-extract invocation).
#include "fb_common.cuh"
#include "fb_tma.cuh"
#include <cooperative_groups.h>
#include "fb_ivector.cuh"
#include <math.h>
#include <string.h>

int fb_run_frontend_flag(fb_ctx *ctx, const int *done_flag);
int fb_run_gmm_store(fb_ctx *ctx, float *ll_out, const int *done_flag);

#define IV_NSEL 20

// ------------------------------------------------------------------------------------------------
// Top-20 Gaussian selection: one warp per frame; lanes hold C/32 values each (C <= 2048).
// Ties resolve to the lower component index (stable descending sort).
// Fast path: find a threshold that between 20 and 64 values pass (normally the 20th largest lane maximum, one counting
// pass; else a bisection below the row maximum), compact those candidates into shared memory and rank them exactly
// (value descending, index ascending).  Rows where no such threshold exists (massive ties) take the plain 20-round
// arg-max selection.
// ------------------------------------------------------------------------------------------------
#define IV_CAND 64

__global__ void __launch_bounds__(256, 3)
gselect_kernel(const float *__restrict__ ll, const int *__restrict__ misc, int C, int *__restrict__ gsel,
               const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  __shared__ unsigned long long s_key[8][IV_CAND];
  const int w = threadIdx.x >> 5;
  const int row = blockIdx.x * 8 + w;
  const int lane = threadIdx.x & 31;
  if (row >= misc[2]) return;
  const int per = C >> 5;
  float v[64];
  const float *src = ll + (size_t)row * C;
#pragma unroll
  for (int i = 0; i < 64; ++i) v[i] = (i < per) ? src[lane + 32 * i] : -INFINITY;
  float best = -INFINITY;
#pragma unroll
  for (int i = 0; i < 64; ++i) best = fmaxf(best, v[i]);
  float M = best;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  // ---- threshold.  First guess: the 20th largest of the 32 lane maxima (bitonic sort across the lanes) -- twenty distinct
  // values are >= it, so at least 20 pass, and in a 2048-component log-likelihood row rarely more than a few dozen do.
  // Only if more than IV_CAND pass, bisect between it and the row maximum (delta_lo passes fewer than 20, delta_hi more
  // than IV_CAND values).
  float srt = best;
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1)
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float other = __shfl_xor_sync(0xffffffffu, srt, j);
      const bool up = ((lane & k) == 0) == ((lane & j) == 0);      // this lane keeps the smaller value
      srt = up ? fminf(srt, other) : fmaxf(srt, other);
    }
  float tau = __shfl_sync(0xffffffffu, srt, 32 - IV_NSEL);          // ascending: position 12 holds the 20th largest
  int cnt = 0, mine = 0;
  unsigned m_lo = 0u, m_hi = 0u;                                    // which of this lane's 64 values pass
  auto count_pass = [&]() {
    m_lo = 0u; m_hi = 0u;
#pragma unroll
    for (int i = 0; i < 32; ++i) m_lo |= (v[i] >= tau) ? (1u << i) : 0u;
#pragma unroll
    for (int i = 0; i < 32; ++i) m_hi |= (v[32 + i] >= tau) ? (1u << i) : 0u;
    mine = __popc(m_lo) + __popc(m_hi);
    cnt = mine;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  };
  count_pass();
  bool ok = cnt <= IV_CAND;                                         // cnt >= IV_NSEL by construction
  if (!ok) {
    float d_lo = 0.f, d_hi = M - tau, delta = 0.5f * (M - tau);
    for (int it = 0; it < 24 && !ok && d_hi > 0.f; ++it) {
      tau = M - delta;
      count_pass();
      if (cnt < IV_NSEL) { d_lo = delta; delta = 0.5f * (d_lo + d_hi); }
      else if (cnt > IV_CAND) { d_hi = delta; delta = 0.5f * (d_lo + d_hi); }
      else ok = true;
    }
  }
  if (ok) {
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    int pos = incl - mine;
    // a lane passes one or two values: walk its mask and re-read them (L1 / L2 hits) instead of testing all 64 registers.
    // A candidate is stored as ONE sortable 64-bit key: order-preserving bits of the value above, inverted component index
    // below, so "value descending, index ascending" is a plain unsigned comparison and the ranking loop below costs one
    // 64-bit load and two compare / add pairs per step for both candidates of a lane.
    auto key_of = [&](int i) {
      const unsigned u = __float_as_uint(src[lane + 32 * i]);
      const unsigned ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);       // monotone map float -> unsigned
      return ((unsigned long long)ord << 32) | (unsigned)(0xFFFFFFFFu - (unsigned)(lane + 32 * i));
    };
    for (unsigned m = m_lo; m; m &= m - 1) s_key[w][pos++] = key_of(__ffs(m) - 1);
    for (unsigned m = m_hi; m; m &= m - 1) s_key[w][pos++] = key_of(32 + __ffs(m) - 1);
    __syncwarp();
    const unsigned long long k0 = (lane < cnt) ? s_key[w][lane] : 0ull, k1 = (lane + 32 < cnt) ? s_key[w][lane + 32] : 0ull;
    int rank0 = 0, rank1 = 0;
    for (int j = 0; j < cnt; ++j) {
      const unsigned long long o = s_key[w][j];
      rank0 += (o > k0) ? 1 : 0;
      rank1 += (o > k1) ? 1 : 0;
    }
    if (lane < cnt && rank0 < IV_NSEL) gsel[(size_t)row * IV_NSEL + rank0] = (int)(0xFFFFFFFFu - (unsigned)k0);
    if (lane + 32 < cnt && rank1 < IV_NSEL) gsel[(size_t)row * IV_NSEL + rank1] = (int)(0xFFFFFFFFu - (unsigned)k1);
    return;
  }
  int bi = 0;                                        // slow path: this lane's arg-max (first maximum)
  best = -INFINITY;
#pragma unroll
  for (int i = 0; i < 64; ++i)
    if (v[i] > best) { best = v[i]; bi = i; }
  for (int k = 0; k < IV_NSEL; ++k) {
    float wv = best;
    int wi = lane + 32 * bi;                      // component index
    int wl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, wv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, wi, o);
      const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
      if (ov > wv || (ov == wv && oi < wi)) { wv = ov; wi = oi; wl = ol; }
    }
    if (lane == 0) gsel[(size_t)row * IV_NSEL + k] = wi;
    if (lane == wl) {
      const int slot = wi >> 5;
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (i == slot) v[i] = -INFINITY;
      best = -INFINITY;
      bi = 0;
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (v[i] > best) { best = v[i]; bi = i; }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Full-covariance log-likelihoods of the 20 selected components + posteriors.  One warp per frame.
// inv_covars are packed lower-triangular row-major (Kaldi SpMatrix), D = 72 -> 2628 entries.
// ------------------------------------------------------------------------------------------------
#define IV_PACKED (FB_DIM * (FB_DIM + 1) / 2)
#define IV_STATS_THREADS 1024  // the per-component accumulation is a chain of L2 round trips: many warps per CTA
#define IV_STATS_SPLIT 2       // CTAs per utterance in ivec_stats_kernel (component ranges); ~144 KB smem each: one wave at B = 51

__global__ void __launch_bounds__(256)
fgmm_post_kernel(const float *__restrict__ feats, const int *__restrict__ gsel, const float *__restrict__ gconsts,
                 const float *__restrict__ means_invcovars, const float *__restrict__ inv_covars_packed,
                 const unsigned short *__restrict__ rc_table, const int *__restrict__ misc, float min_post,
                 float *__restrict__ post, const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  __shared__ float s_x[8][FB_DIM];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + w;
  if (row >= misc[2]) return;
  for (int d = lane; d < FB_DIM; d += 32) s_x[w][d] = feats[(size_t)row * FB_DIM + d];
  __syncwarp();
  float my_ll = -INFINITY;
  for (int j = 0; j < IV_NSEL; ++j) {
    const int c = gsel[(size_t)row * IV_NSEL + j];
    const float *S = inv_covars_packed + (size_t)c * IV_PACKED;
    float q = 0.f;
    for (int e = lane; e < IV_PACKED; e += 32) {
      const unsigned short rc = rc_table[e];
      const int r = rc >> 8, cc = rc & 255;
      const float xx = s_x[w][r] * s_x[w][cc];
      q += S[e] * ((r == cc) ? 0.5f * xx : xx);
    }
    float lin = 0.f;
    for (int d = lane; d < FB_DIM; d += 32) lin += means_invcovars[(size_t)c * FB_DIM + d] * s_x[w][d];
    float tot = lin - q;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == j) my_ll = gconsts[c] + tot;
  }
  // softmax over lanes 0..19 (VectorBase::ApplySoftMax: float exp, sequential float sum, Scale(1/sum))
  float m = my_ll;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float e = (lane < IV_NSEL) ? expf(my_ll - m) : 0.f;
  float sum = 0.f;
  for (int j = 0; j < IV_NSEL; ++j) sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, e, j));
  float p = e * (float)(1.0 / (double)sum);
  if (min_post != 0.f) {
    // argmax (first maximum), prune, renormalise (fgmm-global-gselect-to-post.cc)
    float bv = (lane < IV_NSEL) ? p : -1.f;
    int bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (ov > bv || (ov == bv && ol < bl)) { bv = ov; bl = ol; }
    }
    if (p < min_post) p = 0.f;
    double s2 = 0.0;
    for (int j = 0; j < IV_NSEL; ++j) s2 += (double)__shfl_sync(0xffffffffu, p, j);
    const float s2f = (float)s2;
    if (s2f == 0.f) p = (lane == bl) ? 1.f : 0.f;
    else p = p * (float)(1.0 / (double)s2f);
  }
  if (lane < IV_NSEL) post[(size_t)row * IV_NSEL + lane] = p;
}

// ------------------------------------------------------------------------------------------------
// Same quantity, rows grouped: one CTA handles the rows of EIGHT audios at the same frame index t.
// The plain kernel above executes ~9 instructions per (row, component, packed entry) -- table lookup, two feature loads,
// products, select -- and gathers 20 x 10.5 KB of packed inverse covariances per row from L2 (4.8 GB per NES iteration at
// C3).  Here
//  * the audios of an NES batch are perturbations of one utterance, so rows with the same frame index select (almost) the
//    same components: the CTA stages the UNION of its rows' selections, each component once (~25 instead of 8 x 20).  Any
//    batch is handled correctly: rows that share no component simply do not share loads;
//  * a component is staged as one block [packed inverse covariance | mean x inverse covariance] (2700 floats) by two bulk
//    copies of the TMA engine that complete on the slot's "full" mbarrier; a ring stage holds IV_POST_NCS components
//    (ring of IV_POST_STAGES stages, released through "empty" mbarriers; thread 0 issues stage it + 3 right before it
//    consumes stage it -- a separate producer warp was measured slower: 9 warps of 110 registers do not fit twice into
//    the SM's four register files);
//  * the multipliers of a row -- x_r x_c (halved on the diagonal) against the covariance entries, -x_d against the linear
//    entries, so one pass gives q - lin -- live in REGISTERS as pairs, one 64-bit shared-memory load + one packed FFMA2 per
//    two entries;
//  * warp w = (row group w / 4, entry slice w % 4): it multiplies ONE QUARTER of the staged block with FOUR rows, so a
//    staged value is read from shared memory twice per CTA instead of once per selecting row (the row-per-warp version
//    moved 5 GB through shared memory per launch, half of its run time).  The four slice partials of a (row, component)
//    meet in shared memory (fixed order: deterministic) IV_COMB_LAG stages later, combined by warp (stage % 8).
//    Slot reuse: warp (it % 8) combines stage it inside its iteration it + IV_COMB_LAG, BEFORE it releases that stage's ring
//    slot; partial slot it % IV_PSLOTS is written again in iteration it + IV_PSLOTS, whose data thread 0 only requests
//    after every warp has released stage it + IV_PSLOTS - IV_POST_STAGES >= it + IV_COMB_LAG -- so the combine is done.
// Summation order differs from fgmm_post_kernel (float rounding only); selection, soft-max and pruning are the same.
// ------------------------------------------------------------------------------------------------
#define IV_GROUP 8
#define IV_ENT (IV_PACKED + FB_DIM)                 // a staged component: packed inverse covariance, then mean x inverse covariance
#define IV_PAIRS (IV_ENT / 2)                       // 1350
#define IV_ES 4                                     // entry slices
#define IV_RPG (IV_GROUP / 2)                       // rows per row group (two row groups)
#define IV_SLICE_PAIRS ((IV_PAIRS + IV_ES - 1) / IV_ES)        // 338
#define IV_LP ((IV_SLICE_PAIRS + 31) / 32)          // 11 register pairs per lane and row
#define IV_POST_NCS 2                               // components per ring stage: both are multiplied in one pass over the row
                                                    // multipliers (8 independent FFMA2 chains) and share the barrier traffic
#define IV_POST_STAGES 4                            // stages in flight (2 x 10.8 KB each; a bulk copy takes ~2 k cycles to land)
#define IV_COMB_LAG 2                               // a stage's partials are combined this many stages later
#define IV_PSLOTS 8                                 // partial-sum slots, >= IV_POST_STAGES + IV_COMB_LAG (see the slot-reuse argument)
static_assert(IV_PSLOTS >= IV_POST_STAGES + IV_COMB_LAG, "a partial slot must be combined before the ring lets anybody refill it");
static_assert(IV_PACKED % 4 == 0 && FB_DIM % 4 == 0, "bulk copies need 16-byte aligned component blocks of 16 n bytes");
static_assert(IV_PACKED % 2 == 0 && IV_ENT % 2 == 0, "pairs of entries");

struct __align__(16) PostSmem {
  float S[IV_POST_STAGES][IV_POST_NCS][IV_ENT];
  float gc[IV_GROUP * IV_NSEL];
  float x[IV_GROUP][FB_DIM];
  float ll[IV_GROUP][IV_NSEL];
  float part[IV_PSLOTS][IV_POST_NCS][2][IV_ES][IV_RPG];
  unsigned bitmap[128];                             // C <= 4096
  unsigned short prefix[128];                       // union members before bitmap word i
  int list[IV_GROUP * IV_NSEL];
  unsigned pos[IV_GROUP * IV_NSEL][2];              // per union member: byte r = position of it in row r's selection, 0xFF = not selected
  uint64_t bar[2 * IV_POST_STAGES + IV_PSLOTS];
  int n;
};

__global__ void __launch_bounds__(256, 2)
fgmm_post_group_kernel(const float *__restrict__ feats, const int *__restrict__ gsel, const float *__restrict__ gconsts,
                       const float *__restrict__ means_invcovars, const float *__restrict__ inv_covars_packed,
                       const unsigned short *__restrict__ rc_table, const int *__restrict__ frame_off,
                       const int *__restrict__ vrank, const int *__restrict__ row_off, int B, int C, int n_chunks,
                       float min_post, float *__restrict__ post, const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  extern __shared__ __align__(16) unsigned char post_smem_raw[];
  PostSmem &sm = *reinterpret_cast<PostSmem *>(post_smem_raw);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = w / IV_ES, es = w % IV_ES;
  const int t = blockIdx.x / n_chunks, chunk = blockIdx.x - t * n_chunks;
  const uint32_t full0 = tma_smem_u32(sm.bar), empty0 = tma_smem_u32(sm.bar + IV_POST_STAGES),
                 pfull0 = tma_smem_u32(sm.bar + 2 * IV_POST_STAGES);
  if (threadIdx.x == 0) {
    for (int k = 0; k < IV_POST_STAGES; ++k) {
      tma_bar_init(full0 + 8 * k, 1);
      tma_bar_init(empty0 + 8 * k, IV_GROUP);        // one arrival per warp
    }
    for (int k = 0; k < IV_PSLOTS; ++k) tma_bar_init(pfull0 + 8 * k, IV_GROUP);
    tma_bar_init_fence();
  }
  // row owned by this warp for loading, selection and the soft-max: audio chunk * 8 + w at frame index t,
  // -1 = not voiced / beyond the utterance
  int row = -1;
  {
    const int b = chunk * IV_GROUP + w;
    if (b < B) {
      const int f0 = frame_off[b];
      if (t < frame_off[b + 1] - f0) {
        const int r = vrank[f0 + t];
        if (r >= 0) row = row_off[b] + r;
      }
    }
  }
  if (threadIdx.x < 128) sm.bitmap[threadIdx.x] = 0u;
  for (int i = threadIdx.x; i < IV_GROUP * IV_NSEL * 2; i += blockDim.x) (&sm.pos[0][0])[i] = 0xFFFFFFFFu;
  const int sel = (row >= 0 && lane < IV_NSEL) ? gsel[(size_t)row * IV_NSEL + lane] : -1;
  for (int d = lane; d < FB_DIM; d += 32) sm.x[w][d] = (row >= 0) ? feats[(size_t)row * FB_DIM + d] : 0.f;
  __syncthreads();
  if (sel >= 0) atomicOr(&sm.bitmap[sel >> 5], 1u << (sel & 31));
  // the multipliers of this warp's four rows for its entry slice: pair p = es * IV_SLICE_PAIRS + lane + 32 m
  float2 xx[IV_RPG][IV_LP];
#pragma unroll
  for (int m = 0; m < IV_LP; ++m) {
    const int lp = lane + 32 * m, pr = es * IV_SLICE_PAIRS + lp;
    const bool valid = lp < IV_SLICE_PAIRS && pr < IV_PAIRS;
    const int e = 2 * pr;
    int r0 = 0, c0 = 0, r1 = 0, c1 = 0;
    const bool quad = valid && e < IV_PACKED;        // IV_PACKED is even: a pair is on one side as a whole
    if (quad) {
      const unsigned short rc0 = rc_table[e], rc1 = rc_table[e + 1];
      r0 = rc0 >> 8; c0 = rc0 & 255; r1 = rc1 >> 8; c1 = rc1 & 255;
    }
#pragma unroll
    for (int k = 0; k < IV_RPG; ++k) {
      const float *xr = sm.x[rg * IV_RPG + k];
      float2 v = make_float2(0.f, 0.f);
      if (quad) {
        const float p0 = xr[r0] * xr[c0], p1 = xr[r1] * xr[c1];
        v = make_float2((r0 == c0) ? 0.5f * p0 : p0, (r1 == c1) ? 0.5f * p1 : p1);
      } else if (valid) {
        v = make_float2(-xr[e - IV_PACKED], -xr[e + 1 - IV_PACKED]);
      }
      xx[k][m] = v;
    }
  }
  __syncthreads();
  if (w == 0) {                                       // ordered list of the components any row of the CTA selected
    const int nw = (C + 31) >> 5;                     // bitmap words, 4 per lane
    int cnt = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) cnt += (lane * 4 + k < nw) ? __popc(sm.bitmap[lane * 4 + k]) : 0;
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    int pos = incl - cnt;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      sm.prefix[lane * 4 + k] = (unsigned short)pos;
      if (lane * 4 + k < nw) {
        unsigned bits = sm.bitmap[lane * 4 + k];
        while (bits) {
          const int bpos = __ffs(bits) - 1;
          bits &= bits - 1;
          sm.list[pos++] = (lane * 4 + k) * 32 + bpos;
        }
      }
    }
    if (lane == 31) sm.n = incl;
  }
  __syncthreads();
  const int n_union = sm.n;
  if (n_union == 0) return;
  for (int i = threadIdx.x; i < n_union; i += blockDim.x) sm.gc[i] = gconsts[sm.list[i]];      // not a global load per loop trip
  if (sel >= 0) {                                     // where in the union is my selection, and which of my 20 is it
    const int ui = sm.prefix[sel >> 5] + __popc(sm.bitmap[sel >> 5] & ((1u << (sel & 31)) - 1u));
    reinterpret_cast<unsigned char *>(&sm.pos[ui][0])[w] = (unsigned char)lane;
  }
  if (lane < IV_NSEL) sm.ll[w][lane] = -INFINITY;
  __syncthreads();
  const int n_it = (n_union + IV_POST_NCS - 1) / IV_POST_NCS;       // ring stages: IV_POST_NCS union members each
  auto issue = [&](int it) {
    if (it >= n_it) return;
    const int slot = it % IV_POST_STAGES, round = it / IV_POST_STAGES;
    if (round > 0) { tma_bar_wait(empty0 + 8 * slot, (round - 1) & 1); tma_fence_proxy_async(); }
    const int nc = min(IV_POST_NCS, n_union - it * IV_POST_NCS);
    tma_bar_expect_tx(full0 + 8 * slot, nc * IV_ENT * 4);
    for (int k = 0; k < nc; ++k) {
      const int c = sm.list[it * IV_POST_NCS + k];
      tma_bulk_g2s(tma_smem_u32(sm.S[slot][k]), inv_covars_packed + (size_t)c * IV_PACKED, IV_PACKED * 4, full0 + 8 * slot);
      tma_bulk_g2s(tma_smem_u32(sm.S[slot][k] + IV_PACKED), means_invcovars + (size_t)c * FB_DIM, FB_DIM * 4, full0 + 8 * slot);
    }
  };
  // the four slice partials of (row r, union member uc) -> log-likelihood, in slice order
  auto combine = [&](int itc) {
    const int ps = itc % IV_PSLOTS;
    tma_bar_wait(pfull0 + 8 * ps, (itc / IV_PSLOTS) & 1);
    if (lane < IV_GROUP) {
#pragma unroll
      for (int k = 0; k < IV_POST_NCS; ++k) {
        const int uc = itc * IV_POST_NCS + k;
        if (uc < n_union) {
          const unsigned p = reinterpret_cast<const unsigned char *>(&sm.pos[uc][0])[lane];
          if (p != 0xFFu) {
            const int g2 = lane / IV_RPG, r = lane % IV_RPG;
            const float q = ((sm.part[ps][k][g2][0][r] + sm.part[ps][k][g2][1][r]) + sm.part[ps][k][g2][2][r]) + sm.part[ps][k][g2][3][r];
            sm.ll[lane][p] = sm.gc[uc] - q;           // gconst + lin - q
          }
        }
      }
    }
  };
  if (threadIdx.x == 0)
    for (int u0 = 0; u0 < IV_POST_STAGES - 1; ++u0) issue(u0);
  const bool hi = (lane & 16) != 0, b8 = (lane & 8) != 0;
  // 4 rows x 32 lanes -> one total per row: exchange halves (rows 0,1 | 2,3), then (row a | row b), then a butterfly
  auto reduce4 = [&](const float2 (&acc)[IV_RPG]) {
    const float t0 = acc[0].x + acc[0].y, t1 = acc[1].x + acc[1].y, t2 = acc[2].x + acc[2].y, t3 = acc[3].x + acc[3].y;
    float k0 = hi ? t2 : t0, k1 = hi ? t3 : t1;
    k0 += __shfl_xor_sync(0xffffffffu, hi ? t0 : t2, 16);
    k1 += __shfl_xor_sync(0xffffffffu, hi ? t1 : t3, 16);
    float kk = b8 ? k1 : k0;
    kk += __shfl_xor_sync(0xffffffffu, b8 ? k0 : k1, 8);
    kk += __shfl_xor_sync(0xffffffffu, kk, 4);
    kk += __shfl_xor_sync(0xffffffffu, kk, 2);
    kk += __shfl_xor_sync(0xffffffffu, kk, 1);
    return kk;
  };
  for (int it = 0; it < n_it; ++it) {
    const int slot = it % IV_POST_STAGES;
    if (threadIdx.x == 0) issue(it + IV_POST_STAGES - 1);            // into the slot stage it - 1 used
    tma_bar_wait(full0 + 8 * slot, (it / IV_POST_STAGES) & 1);
    {
      // both components of the stage in one pass (an absent second one of the last stage reads stale, finite data and is
      // never combined); a row group none of whose rows selected a component computes it anyway (~1 % of the cases)
      const float2 *Sa = reinterpret_cast<const float2 *>(sm.S[slot][0]) + es * IV_SLICE_PAIRS + lane;
      const float2 *Sb = reinterpret_cast<const float2 *>(sm.S[slot][1]) + es * IV_SLICE_PAIRS + lane;
      float2 acca[IV_RPG], accb[IV_RPG];
#pragma unroll
      for (int k = 0; k < IV_RPG; ++k) acca[k] = accb[k] = make_float2(0.f, 0.f);
#pragma unroll
      for (int m = 0; m < IV_LP; ++m) {
        // only the last pair index is partly outside the slice (xx is zero there, but the buffer holds no defined value)
        const bool inside = 32 * m + 31 < IV_SLICE_PAIRS && (IV_ES - 1) * IV_SLICE_PAIRS + 32 * m + 31 < IV_PAIRS;
        if (inside || (lane + 32 * m < IV_SLICE_PAIRS && es * IV_SLICE_PAIRS + lane + 32 * m < IV_PAIRS)) {
          const float2 sa = Sa[32 * m], sb = Sb[32 * m];
#pragma unroll
          for (int k = 0; k < IV_RPG; ++k) {
            acca[k] = __ffma2_rn(sa, xx[k][m], acca[k]);
            accb[k] = __ffma2_rn(sb, xx[k][m], accb[k]);
          }
        }
      }
      const float ka = reduce4(acca), kb = reduce4(accb);
      if ((lane & 7) == 0) {
        const int r = (hi ? 2 : 0) + (b8 ? 1 : 0);
        sm.part[it % IV_PSLOTS][0][rg][es][r] = ka;
        sm.part[it % IV_PSLOTS][1][rg][es][r] = kb;
      }
    }
    if (it >= IV_COMB_LAG && ((it - IV_COMB_LAG) % IV_GROUP) == w) combine(it - IV_COMB_LAG);
    __syncwarp();
    if (lane == 0) {
      tma_bar_arrive(pfull0 + 8 * (it % IV_PSLOTS));
      tma_bar_arrive(empty0 + 8 * slot);
    }
  }
  for (int itc = max(0, n_it - IV_COMB_LAG); itc < n_it; ++itc)
    if ((itc % IV_GROUP) == w) combine(itc);
  __syncthreads();
  if (row < 0) return;
  const float my_ll = (lane < IV_NSEL) ? sm.ll[w][lane] : -INFINITY;
  // softmax / pruning, exactly as in fgmm_post_kernel
  float m = my_ll;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  const float e = (lane < IV_NSEL) ? expf(my_ll - m) : 0.f;
  float sum = 0.f;
  for (int j = 0; j < IV_NSEL; ++j) sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, e, j));
  float p = e * (float)(1.0 / (double)sum);
  if (min_post != 0.f) {
    // argmax (first maximum), prune, renormalise (fgmm-global-gselect-to-post.cc)
    float bv = (lane < IV_NSEL) ? p : -1.f;
    int bl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
      if (ov > bv || (ov == bv && ol < bl)) { bv = ov; bl = ol; }
    }
    if (p < min_post) p = 0.f;
    double s2 = 0.0;
    for (int j = 0; j < IV_NSEL; ++j) s2 += (double)__shfl_sync(0xffffffffu, p, j);
    const float s2f = (float)s2;
    if (s2f == 0.f) p = (lane == bl) ? 1.f : 0.f;
    else p = p * (float)(1.0 / (double)s2f);
  }
  if (lane < IV_NSEL) post[(size_t)row * IV_NSEL + lane] = p;
}

// ------------------------------------------------------------------------------------------------
// Baum-Welch statistics per utterance: gamma[b][c], X[b][c][72] (float64, dense).  One CTA per utterance.
// Pairs are bucketed by component in shared memory, each bucket is put in frame order (rank sort), then one warp
// per component accumulates sequentially -- the same order as Kaldi's per-frame AccStats, deterministic.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(IV_STATS_THREADS)
ivec_stats_kernel(const float *__restrict__ feats, const int *__restrict__ gsel, const float *__restrict__ post,
                  const int *__restrict__ row_off, int C, int max_pairs, double *__restrict__ gamma,
                  double *__restrict__ Xs, int *__restrict__ err, const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  extern __shared__ unsigned char s_raw[];
  int *cnt = reinterpret_cast<int *>(s_raw);                  // [C]
  int *off = cnt + C;                                         // [C + 1]
  int *cur = off + C + 1;                                     // [C]
  unsigned short *lt = reinterpret_cast<unsigned short *>(cur + C);   // [max_pairs] frame index, unsorted
  unsigned short *ls = lt + max_pairs;                        // [max_pairs] sorted
  float *lp = reinterpret_cast<float *>(ls + max_pairs + (max_pairs & 1));   // [max_pairs] unsorted
  float *lq = lp + max_pairs;                                 // [max_pairs] sorted
  __shared__ int s_scan[256];
  const int b = blockIdx.x;
  // blockIdx.y splits the components: a CTA buckets and accumulates only the pairs of its component range
  const int cq0 = (int)((long long)C * blockIdx.y / gridDim.y), cq1 = (int)((long long)C * (blockIdx.y + 1) / gridDim.y);
  const int r00 = row_off[b], Tv_all = row_off[b + 1] - r00;
  const int tid = threadIdx.x;
  // Utterances longer than the shared-memory bucket capacity are processed in frame chunks, in frame order: the
  // accumulators of a component continue from the values the previous chunk stored, so the sequence of additions is the
  // same as for one pass (and as Kaldi's per-frame AccStats).
  const int chunk_frames = max_pairs / IV_NSEL;
  (void)err;
  for (int f0 = 0; f0 < Tv_all || f0 == 0; f0 += chunk_frames) {
  const int r0 = r00 + f0;
  const int Tv = min(chunk_frames, Tv_all - f0);
  __syncthreads();
  for (int c = tid; c < C; c += blockDim.x) { cnt[c] = 0; cur[c] = 0; }
  __syncthreads();
  for (int i = tid; i < Tv * IV_NSEL; i += blockDim.x) {
    const int c = gsel[(size_t)r0 * IV_NSEL + i];
    if (c >= cq0 && c < cq1 && post[(size_t)r0 * IV_NSEL + i] != 0.f) atomicAdd(&cnt[c], 1);
  }
  __syncthreads();
  // exclusive scan of cnt by the first 256 threads (C <= 2048: 8 per thread); the block may be larger
  const int per = (C + 255) / 256;
  if (tid < 256) {
    int local = 0;
    for (int k = 0; k < per; ++k) { const int c = tid * per + k; if (c < C) local += cnt[c]; }
    s_scan[tid] = local;
  }
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int i = 0; i < 256; ++i) { const int t = s_scan[i]; s_scan[i] = run; run += t; }
  }
  __syncthreads();
  if (tid < 256) {
    int run = s_scan[tid];
    for (int k = 0; k < per; ++k) { const int c = tid * per + k; if (c < C) { off[c] = run; run += cnt[c]; } }
    if (tid == 255) off[C] = run;
  }
  __syncthreads();
  for (int i = tid; i < Tv * IV_NSEL; i += blockDim.x) {
    const float p = post[(size_t)r0 * IV_NSEL + i];
    const int c = gsel[(size_t)r0 * IV_NSEL + i];
    if (p != 0.f && c >= cq0 && c < cq1) {
      const int pos = off[c] + atomicAdd(&cur[c], 1);
      lt[pos] = (unsigned short)(i / IV_NSEL);
      lp[pos] = p;
    }
  }
  __syncthreads();
  const int w = tid >> 5, lane = tid & 31, n_warps = blockDim.x >> 5;
  for (int c = cq0 + w; c < cq1; c += n_warps) {
    const int n = cnt[c], o = off[c];
    if (f0 > 0 && n == 0) continue;                           // nothing to add in this chunk
    // rank sort by frame index (frame indices within a bucket are distinct)
    for (int i = lane; i < n; i += 32) {
      const unsigned short t = lt[o + i];
      int rank = 0;
      for (int k = 0; k < n; ++k) rank += (lt[o + k] < t) ? 1 : 0;
      ls[o + rank] = t;
      lq[o + rank] = lp[o + i];
    }
    __syncwarp();
    double *xo = Xs + ((size_t)b * C + c) * FB_DIM;
    double g = 0.0, x0 = 0.0, x1 = 0.0, x2 = 0.0;
    if (f0 > 0) {
      g = gamma[(size_t)b * C + c];
      x0 = xo[lane]; x1 = xo[lane + 32];
      if (lane < 8) x2 = xo[lane + 64];
    }
    for (int i = 0; i < n; ++i) {
      const double p = (double)lq[o + i];
      const float *xr = feats + (size_t)(r0 + ls[o + i]) * FB_DIM;
      g += p;
      x0 += p * (double)xr[lane];
      x1 += p * (double)xr[lane + 32];
      if (lane < 8) x2 += p * (double)xr[lane + 64];
    }
    xo[lane] = x0;
    xo[lane + 32] = x1;
    if (lane < 8) xo[lane + 64] = x2;
    if (lane == 0) gamma[(size_t)b * C + c] = g;
  }
  }
}

// ------------------------------------------------------------------------------------------------
// Derived extractor parameters (once, at load): SIM[c] = Sigma_inv[c] * M[c] (D x R),  U[c] = vech(M[c]^T SIM[c]).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ivec_derive_sim_kernel(const double *__restrict__ M, const double *__restrict__ sigma_inv, int R, double *__restrict__ sim64,
                       float *__restrict__ sim32) {
  const int c = blockIdx.x;
  for (int idx = threadIdx.x; idx < FB_DIM * R; idx += blockDim.x) {
    const int d = idx / R, r = idx - d * R;
    double acc = 0.0;
    for (int e = 0; e < FB_DIM; ++e) acc += sigma_inv[((size_t)c * FB_DIM + d) * FB_DIM + e] * M[((size_t)c * FB_DIM + e) * R + r];
    sim64[(size_t)c * FB_DIM * R + idx] = acc;
    sim32[(size_t)c * FB_DIM * R + idx] = (float)acc;
  }
}

__global__ void __launch_bounds__(256)
ivec_derive_u_kernel(const double *__restrict__ M, const double *__restrict__ sim64, int R, int n_packed, float *__restrict__ U) {
  const int c = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_packed) return;
  // packed lower-triangular index e -> (r, s), s <= r
  int r = (int)((sqrt(8.0 * (double)e + 1.0) - 1.0) * 0.5);
  while ((r + 1) * (r + 2) / 2 <= e) ++r;
  while (r * (r + 1) / 2 > e) --r;
  const int s = e - r * (r + 1) / 2;
  double acc = 0.0;
  for (int d = 0; d < FB_DIM; ++d) acc += M[((size_t)c * FB_DIM + d) * R + r] * sim64[((size_t)c * FB_DIM + d) * R + s];
  U[(size_t)c * n_packed + e] = (float)acc;
}

// ------------------------------------------------------------------------------------------------
// lin and quad: float64 accumulation of fp32 parameter streams, register-tiled: each thread owns 4 consecutive output
// columns x 8 utterances (32 accumulators); parameters come in as one float4 per thread per k, statistics as 4 LDS.128.
// A CTA covers IV_BCHUNK = 32 utterances; components whose gamma is zero for all 32 utterances are skipped entirely
// (NES batches are perturbations of one audio, so their active component sets nearly coincide).
// ------------------------------------------------------------------------------------------------
#define IV_BCHUNK 32

// lin partials: grid (n_splits, ceil(B/32)); block = 4 utterance-groups x ceil(R/4) column-groups (R <= 512).
static_assert(FB_DIM % 4 == 0, "ivec_lin_kernel walks the feature dimension four rows at a time");
__global__ void __launch_bounds__(512)
ivec_lin_kernel(const float *__restrict__ sim32, const double *__restrict__ Xs, const int *__restrict__ act_list, int B, int C,
                int R, int n_splits, double *__restrict__ part, const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  __shared__ __align__(16) double s_x[FB_DIM][IV_BCHUNK];
  const int split = blockIdx.x;
  const int b0 = blockIdx.y * IV_BCHUNK;
  const int nb = min(IV_BCHUNK, B - b0);
  // the ACTIVE components of this utterance chunk (ivec_active_kernel), split evenly over the CTAs of the row
  const int *list = act_list + (size_t)blockIdx.y * (C + 1);
  const int n_act = list[0];
  const int a_lo = (int)((long long)n_act * split / n_splits), a_hi = (int)((long long)n_act * (split + 1) / n_splits);
  const int ncg = (R + 3) / 4;
  const int ug = threadIdx.x / ncg, cg = threadIdx.x - ug * ncg;      // utterance group 0..3, column group
  const bool active = ug < 4 && ug * 8 < nb;            // utterance groups beyond the batch (B = 51: the last 8 slots) skip the work
  const int r0 = cg * 4;
  double acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  for (int ai = a_lo; ai < a_hi; ++ai) {
    const int c = list[1 + ai];
    __syncthreads();
    for (int idx = threadIdx.x; idx < FB_DIM * IV_BCHUNK; idx += blockDim.x) {
      const int d = idx / IV_BCHUNK, i = idx - d * IV_BCHUNK;
      s_x[d][i] = (i < nb) ? Xs[((size_t)(b0 + i) * C + c) * FB_DIM + d] : 0.0;
    }
    __syncthreads();
    if (!active) continue;
    const float *col = sim32 + (size_t)c * FB_DIM * R + r0;
    const bool vec_ok = r0 + 3 < R && (R & 3) == 0;
    // four parameter rows in flight per thread: the 236 MB stream is latency bound with one
    for (int d0 = 0; d0 < FB_DIM; d0 += 4) {
      float p[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (vec_ok) {
          const float4 q = __ldcs(reinterpret_cast<const float4 *>(col + (size_t)(d0 + u) * R));
          p[u][0] = q.x; p[u][1] = q.y; p[u][2] = q.z; p[u][3] = q.w;
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) p[u][i] = (r0 + i < R) ? col[(size_t)(d0 + u) * R + i] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const double2 *xr = reinterpret_cast<const double2 *>(&s_x[d0 + u][ug * 8]);
        const double2 x01 = xr[0], x23 = xr[1], x45 = xr[2], x67 = xr[3];
        const double x[8] = {x01.x, x01.y, x23.x, x23.y, x45.x, x45.y, x67.x, x67.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] += (double)p[u][i] * x[j];
      }
    }
  }
  if (active)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int r = r0 + i, bi = ug * 8 + j;
        if (r < R && bi < nb) part[((size_t)split * B + b0 + bi) * R + r] = acc[i][j];
      }
}

// Active components of a 32-utterance chunk: c is active when gamma[b][c] != 0 for some utterance of the chunk (top-20
// selection + posterior pruning leave ~18 % of the components at C3).  One CTA per chunk; ordered (ascending) compaction so
// that consumers accumulate in the same order as a plain loop over c.  list[0] = count, list[1..] = components.
__global__ void __launch_bounds__(1024)
ivec_active_kernel(const double *__restrict__ gamma, int B, int C, int *__restrict__ act_list, const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int b0 = blockIdx.x * IV_BCHUNK;
  const int nb = min(IV_BCHUNK, B - b0);
  int *list = act_list + (size_t)blockIdx.x * (C + 1);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int c0 = 0; c0 < C; c0 += blockDim.x) {
    const int c = c0 + threadIdx.x;
    bool act = false;
    if (c < C)
      for (int i = 0; i < nb; ++i) act |= gamma[(size_t)(b0 + i) * C + c] != 0.0;
    const unsigned bal = __ballot_sync(0xffffffffu, act);
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    int before = s_base;
    for (int k = 0; k < w; ++k) before += s_warp[k];
    if (act) list[1 + before + __popc(bal & ((1u << lane) - 1u))] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += s_warp[k];
      s_base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) list[0] = s_base;
}

// quad: grid (ceil(n_packed / 256), ceil(B/32)); block 256 = 4 utterance-groups x 64 column-groups of 4 packed entries.
// The kernel streams the ACTIVE rows of U (C x n_packed fp32, 657 MB at C = 2048, R = 400; ~120 MB active at C3) with one
// 16-byte load per thread and component, so it is latency bound unless several loads are in flight and the per-chunk
// bookkeeping is rare: it walks the compacted active list 64 components at a time (their gammas gathered into shared
// memory once per chunk) and issues the loads of four components before the first FMA.
__global__ void __launch_bounds__(256, 2)
ivec_quad_kernel(const float *__restrict__ U, const double *__restrict__ gamma, const int *__restrict__ act_list, int B, int C,
                 int n_packed, double *__restrict__ quad, const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  __shared__ __align__(16) double s_g[64][IV_BCHUNK];
  __shared__ int s_c[64];
  const int b0 = blockIdx.y * IV_BCHUNK;
  const int nb = min(IV_BCHUNK, B - b0);
  const int *list = act_list + (size_t)blockIdx.y * (C + 1);
  const int n_act = list[0];
  const int ug = threadIdx.x >> 6, cg = threadIdx.x & 63;
  const int e0 = blockIdx.x * 256 + cg * 4;
  const bool vec_ok = (n_packed & 3) == 0 && e0 + 3 < n_packed;
  double acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  auto load_row = [&](int c, float (&p)[4]) {
    const float *row = U + (size_t)c * n_packed + e0;
    if (vec_ok) {
      const float4 q = __ldcs(reinterpret_cast<const float4 *>(row));      // streamed once: do not keep in L2
      p[0] = q.x; p[1] = q.y; p[2] = q.z; p[3] = q.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = (e0 + i < n_packed) ? row[i] : 0.f;
    }
  };
  auto fma_row = [&](int k, const float (&p)[4]) {
    const double2 *gr = reinterpret_cast<const double2 *>(&s_g[k][ug * 8]);
    const double2 g01 = gr[0], g23 = gr[1], g45 = gr[2], g67 = gr[3];
    const double gv[8] = {g01.x, g01.y, g23.x, g23.y, g45.x, g45.y, g67.x, g67.y};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] += (double)p[i] * gv[j];
  };
  for (int a0 = 0; a0 < n_act; a0 += 64) {
    const int n = min(64, n_act - a0);
    __syncthreads();
    if (threadIdx.x < 64) s_c[threadIdx.x] = (threadIdx.x < n) ? list[1 + a0 + threadIdx.x] : 0;
    __syncthreads();
    for (int idx = threadIdx.x; idx < 64 * IV_BCHUNK; idx += blockDim.x) {
      const int k = idx / IV_BCHUNK, i = idx - k * IV_BCHUNK;
      s_g[k][i] = (i < nb && k < n) ? gamma[(size_t)(b0 + i) * C + s_c[k]] : 0.0;
    }
    __syncthreads();
    int i = 0;
    for (; i + 4 <= n; i += 4) {
      float p0[4], p1[4], p2[4], p3[4];
      load_row(s_c[i], p0); load_row(s_c[i + 1], p1); load_row(s_c[i + 2], p2); load_row(s_c[i + 3], p3);
      fma_row(i, p0); fma_row(i + 1, p1); fma_row(i + 2, p2); fma_row(i + 3, p3);
    }
    for (; i < n; ++i) {
      float p[4];
      load_row(s_c[i], p);
      fma_row(i, p);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = e0 + i, bi = ug * 8 + j;
      if (e < n_packed && bi < nb) quad[(size_t)(b0 + bi) * n_packed + e] = acc[i][j];
    }
}

// The same sum with the rows of U staged by the TMA engine (used whenever n_packed is a multiple of 4, i.e. 16-byte
// aligned row segments).  ncu of the kernel above at C3: 282 us, issue slots 32 % busy, 5.4 warps stalled on global loads per
// issued instruction -- a thread loads four rows, waits, then runs its 128 DFMAs, and with 126 registers only 16 warps per SM
// hide that.  Here one thread streams the active rows (one bulk copy of this CTA's column segment per component, sixteen
// components per stage -- with four, the per-stage barrier traffic cost 17 %: 196 -> 162 us -- IV_QUAD_STAGES stages in
// flight, completion on mbarriers) while 8 consumer warps only read shared
// memory and feed the FP64 pipe; the gammas of the next 64 components are prefetched into registers during the current
// 64.  Grid (b-chunks, column blocks): the CTAs of one column block run side by side, so its rows come from HBM once and
// from L2 for the other utterance chunks.  Arithmetic and summation order per entry are those of the kernel above.
#define IV_QUAD_STAGES 4
#define IV_QUAD_STAGE_COMPS 16
#define IV_QUAD_COLS 256
#define IV_QUAD_RING_BYTES (IV_QUAD_STAGES * IV_QUAD_STAGE_COMPS * IV_QUAD_COLS * 4)
#define IV_QUAD_GSTRIDE (IV_BCHUNK + 2)              // row stride (doubles) of the multiplier chunk: even (16-byte rows), and the
                                                    // transposed stores of the gather hit 8 banks instead of 1
#define IV_QUAD_G_BYTES (64 * IV_QUAD_GSTRIDE * 8)
static inline size_t ivec_quad_tma_smem(int C) { return IV_QUAD_RING_BYTES + IV_QUAD_G_BYTES + 2 * IV_QUAD_STAGES * 8 + ((size_t)C + 4) * 4; }

// One body for both sums: the streamed matrix W has RPC rows of `row_len` floats per component (quad: U, RPC = 1,
// row_len = R (R + 1) / 2; lin: Sigma^-1 M, RPC = 72, row_len = R), the multiplier of row (c, d) for utterance b is
// mult[(b C + c) RPC + d] (gamma, resp. the first-order statistics X), the components a CTA covers are the split-th part of
// the chunk's active list, and out[(split B + b) row_len + e] receives its partial sum (quad: one split).
template <int RPC>
__device__ __forceinline__ void ivec_stream_body(const float *__restrict__ W, const double *__restrict__ mult,
                                                 const int *__restrict__ act_list, int B, int C, int row_len, int n_splits,
                                                 double *__restrict__ out) {
  extern __shared__ __align__(128) unsigned char q_smem[];
  float *ring = reinterpret_cast<float *>(q_smem);
  double (*s_g)[IV_QUAD_GSTRIDE] = reinterpret_cast<double (*)[IV_QUAD_GSTRIDE]>(q_smem + IV_QUAD_RING_BYTES);
  uint64_t *bars = reinterpret_cast<uint64_t *>(q_smem + IV_QUAD_RING_BYTES + IV_QUAD_G_BYTES);
  int *s_list = reinterpret_cast<int *>(bars + 2 * IV_QUAD_STAGES);
  const uint32_t full0 = tma_smem_u32(bars), empty0 = tma_smem_u32(bars + IV_QUAD_STAGES);
  const int b0 = blockIdx.x * IV_BCHUNK;
  const int nb = min(IV_BCHUNK, B - b0);
  const int col0 = blockIdx.y * IV_QUAD_COLS;
  const int split = blockIdx.z;
  const int *list = act_list + (size_t)blockIdx.x * (C + 1);
  const int n_act_all = list[0];
  const int a_lo = (int)((long long)n_act_all * split / n_splits), a_hi = (int)((long long)n_act_all * (split + 1) / n_splits);
  const int n_act = (a_hi - a_lo) * RPC;             // items = rows of W this CTA streams, in (component, d) order
  for (int i = threadIdx.x; i < a_hi - a_lo; i += blockDim.x) s_list[i] = list[1 + a_lo + i];
  if (threadIdx.x == 0) {
    for (int s = 0; s < IV_QUAD_STAGES; ++s) {
      tma_bar_init(full0 + 8 * s, 1);
      tma_bar_init(empty0 + 8 * s, 8);                // one arrival per consumer warp
    }
    tma_bar_init_fence();
  }
  __syncthreads();
  const int n_stages = (n_act + IV_QUAD_STAGE_COMPS - 1) / IV_QUAD_STAGE_COMPS;
  const int lane = threadIdx.x & 31;
  // thread 0 is also the producer: stage st + IV_QUAD_STAGES - 1 is issued right before stage st is consumed, into the slot
  // that stage st - 1 used (its "empty" barrier completes when all 8 warps have released it).  A separate producer warp
  // (288 threads) costs the second CTA per SM: 9 warps do not split evenly over the 4 sub-partitions' register files.
  const uint32_t seg_bytes = (uint32_t)min(IV_QUAD_COLS, row_len - col0) * 4u;
  const uint32_t ring0 = tma_smem_u32(ring);
  auto row_of = [&](int item) -> size_t {            // row index of W for an item
    if (RPC == 1) return (size_t)s_list[item];
    return (size_t)s_list[item / RPC] * RPC + (item % RPC);
  };
  auto issue = [&](int st) {
    if (st >= n_stages) return;
    const int slot = st % IV_QUAD_STAGES, round = st / IV_QUAD_STAGES;
    if (round > 0) { tma_bar_wait(empty0 + 8 * slot, (round - 1) & 1); tma_fence_proxy_async(); }
    const int nc = min(IV_QUAD_STAGE_COMPS, n_act - st * IV_QUAD_STAGE_COMPS);
    tma_bar_expect_tx(full0 + 8 * slot, nc * seg_bytes);
    for (int k = 0; k < nc; ++k)
      tma_bulk_g2s(ring0 + (uint32_t)((slot * IV_QUAD_STAGE_COMPS + k) * IV_QUAD_COLS * 4),
                   W + row_of(st * IV_QUAD_STAGE_COMPS + k) * row_len + col0, seg_bytes, full0 + 8 * slot);
  };
  if (threadIdx.x == 0)
    for (int st0 = 0; st0 < IV_QUAD_STAGES - 1; ++st0) issue(st0);
  const int ug = threadIdx.x >> 6, cg = threadIdx.x & 63;
  const int e0 = col0 + cg * 4;
  const bool work = ug * 8 < nb;                      // warp-uniform: utterance groups beyond the batch skip the arithmetic
  double acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  // multiplier gather of a 64-item chunk: element idx = threadIdx.x + 256 q -> item idx % 64, utterance idx / 64, so that
  // neighbouring threads read neighbouring rows of one utterance (lin: consecutive d of a component are contiguous in X;
  // quad: neighbouring active components) instead of striding over the utterances
  double gnext[8];
  auto gather = [&](int a0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = threadIdx.x + 256 * q, k = idx & 63, i = idx >> 6;
      gnext[q] = (i < nb && a0 + k < n_act) ? mult[(size_t)(b0 + i) * C * RPC + row_of(a0 + k)] : 0.0;
    }
  };
  gather(0);
  int st = 0;
  for (int a0 = 0; a0 < n_act; a0 += 64) {
    const int n = min(64, n_act - a0);
    __syncthreads();                                  // everybody finished reading the previous chunk's multipliers
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int idx = threadIdx.x + 256 * q;
      s_g[idx & 63][idx >> 6] = gnext[q];
    }
    __syncthreads();
    if (a0 + 64 < n_act) gather(a0 + 64);
    for (int k0 = 0; k0 < n; k0 += IV_QUAD_STAGE_COMPS, ++st) {
      const int slot = st % IV_QUAD_STAGES;
      if (threadIdx.x == 0) issue(st + IV_QUAD_STAGES - 1);
      tma_bar_wait(full0 + 8 * slot, (st / IV_QUAD_STAGES) & 1);
      if (work) {
        const float4 *r = reinterpret_cast<const float4 *>(ring + (size_t)slot * IV_QUAD_STAGE_COMPS * IV_QUAD_COLS) + cg;
#pragma unroll
        for (int kk = 0; kk < IV_QUAD_STAGE_COMPS; ++kk) {
          if (k0 + kk < n) {
            const float4 pv = r[kk * (IV_QUAD_COLS / 4)];
            const float p[4] = {pv.x, pv.y, pv.z, pv.w};
            const double2 *gr = reinterpret_cast<const double2 *>(&s_g[k0 + kk][ug * 8]);
            const double2 g01 = gr[0], g23 = gr[1], g45 = gr[2], g67 = gr[3];
            const double gv[8] = {g01.x, g01.y, g23.x, g23.y, g45.x, g45.y, g67.x, g67.y};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[i][j] += (double)p[i] * gv[j];
          }
        }
      }
      __syncwarp();
      if (lane == 0) tma_bar_arrive(empty0 + 8 * slot);
    }
  }
  if (work && e0 < row_len) {                          // row_len % 4 == 0: a thread's four entries are inside or outside together
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int bi = ug * 8 + j;
      if (bi < nb) {
        double2 *dst = reinterpret_cast<double2 *>(out + ((size_t)split * B + b0 + bi) * row_len + e0);
        dst[0] = make_double2(acc[0][j], acc[1][j]);
        dst[1] = make_double2(acc[2][j], acc[3][j]);
      }
    }
  }
}

__global__ void __launch_bounds__(256, 2)
ivec_quad_tma_kernel(const float *__restrict__ U, const double *__restrict__ gamma, const int *__restrict__ act_list, int B, int C,
                     int n_packed, double *__restrict__ quad, const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  ivec_stream_body<1>(U, gamma, act_list, B, C, n_packed, 1, quad);
}

// lin through the same pipeline: grid (b-chunks, ceil(R / 256), splits); replaces ivec_lin_kernel whenever R % 4 == 0 (that
// kernel has the structure ivec_quad_kernel had: four loads, wait, 128 DFMAs, 13 warps per SM).
__global__ void __launch_bounds__(256, 2)
ivec_lin_tma_kernel(const float *__restrict__ sim32, const double *__restrict__ Xs, const int *__restrict__ act_list, int B, int C,
                    int R, int n_splits, double *__restrict__ part, const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  ivec_stream_body<FB_DIM>(sim32, Xs, act_list, B, C, R, n_splits, part);
}

// ------------------------------------------------------------------------------------------------
// Per utterance: A = unpack(quad) + I, rhs = sum of lin partials (+ prior offset on element 0), right-looking blocked
// Cholesky (block 32) with the right-hand side carried as an extra matrix row, so the forward substitution falls out of the
// factorisation; blocked backward substitution.  One CTA (512 threads) per utterance; A lives in global scratch
// [b][R][R] (L2 resident), lower triangle only.  Per block column (all float64, reciprocal diagonal like LAPACK dpotf2):
//   diagonal block   warp 0, one row per lane in registers, shuffles (no shared-memory round trips)
//   L21 = A21 L11^-T one thread per row, the row in registers, L11 broadcast from shared memory
//   trailing update  warp tiles of 8 rows x 128 columns (lane owns columns l, l+32, l+64, l+96: conflict-free panel reads,
//                    coalesced global read-modify-write), 32 accumulators per lane
// A thread-block cluster of IV_SOLVE_CLUSTER CTAs works on one utterance: every CTA factors the (small) panel redundantly
// -- identical arithmetic, identical results -- and takes every IV_SOLVE_CLUSTER-th tile of the trailing update; a
// cluster barrier (release / acquire) per block column makes the updates visible to the partner.
// ------------------------------------------------------------------------------------------------
#define IV_NB 32
#define IV_PSTRIDE 33      // padded panel row stride (doubles): consecutive rows map to different banks
#define IV_SDIAG (IV_NB * (IV_NB - 1) / 2)          // strict lower triangle of a diagonal block
#define IV_SOLVE_CLUSTER 2 // CTAs (SMs) per utterance: at B = 51 one CTA per utterance would leave 97 of 148 SMs idle

__global__ void __launch_bounds__(512)
ivec_solve_kernel(const double *__restrict__ quad, const double *__restrict__ lin_part, int n_splits, int B, int R, int n_packed,
                  double prior_offset, double *__restrict__ Awork, float *__restrict__ ivec, int *__restrict__ err,
                  const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
#ifdef IV_SOLVE_STATS
  long long st[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long st_t = clock64();
#define IV_LAP(i) do { const long long n_ = clock64(); st[i] += n_ - st_t; st_t = n_; } while (0)
  long long st2[4] = {0, 0, 0, 0};
#define IV_LAP2(i) do { const long long n_ = clock64(); st2[i] += n_ - st_t; st_t = n_; } while (0)
#else
#define IV_LAP(i)
#define IV_LAP2(i)
#endif
  extern __shared__ double s_dyn[];
  double *rhs = s_dyn;                               // [R]   right-hand side -> y -> w
  double *invd = s_dyn + R;                          // [R]   1 / L[k][k]
  double *panel = s_dyn + 2 * R;                     // [R + 2][IV_PSTRIDE]: rows k0..R-1 of the block column, the rhs row, a zero row
  double *sdiag = panel + (size_t)(R + 2) * IV_PSTRIDE;   // [blocks][496]: strict lower triangles of the factored diagonal blocks,
                                                     // entry (k, l) times 1 / L[l][l], kept for the back substitution
  __shared__ double s_l11[IV_NB][IV_PSTRIDE];        // factored diagonal block, identity-padded to 32 x 32
  __shared__ double s_a[IV_NB][IV_PSTRIDE];          // working copy during its factorisation
  __shared__ __align__(16) double s_l11t[IV_NB][IV_NB + 2];   // [q][j] = L[j][q] / L[j][j]: one FMA per substitution step, read in pairs
  __shared__ double s_invd[IV_NB];
  __shared__ double s_blk[IV_NB];
  __shared__ int s_fail;
  cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
  const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
  const int b = blockIdx.x / csize, tid = threadIdx.x, nt = blockDim.x;
  const int warp = tid >> 5, lane = tid & 31, nw = nt >> 5;
  double *A = Awork + (size_t)b * R * R;
  if (tid == 0) s_fail = 0;
  // A = unpack(quad) + I is never materialised: block column 0 is loaded straight from the packed triangle and the first
  // trailing update writes quad + I - L21 L21^T instead of updating in place.
  const double *qb = quad + (size_t)b * n_packed;
  for (int r = tid; r < R; r += nt) {
    double acc = 0.0;
    for (int s0 = 0; s0 < n_splits; s0 += 32) {      // 32 partials in flight (the plain loop ran at one L2 round trip per 4);
      double v[32];                                  // the additions keep their order (+ 0.0 beyond the last split)
#pragma unroll
      for (int q = 0; q < 32; ++q) v[q] = (s0 + q < n_splits) ? __ldg(lin_part + ((size_t)(s0 + q) * B + b) * R + r) : 0.0;
#pragma unroll
      for (int q = 0; q < 32; ++q) acc += v[q];
    }
    rhs[r] = acc + ((r == 0) ? prior_offset : 0.0);
  }
  __syncthreads();
  IV_LAP(0);
  for (int k0 = 0; k0 < R; k0 += IV_NB) {
    const int nbk = min(IV_NB, R - k0);
    const int rows = R - k0;                         // matrix rows in the panel (global rows k0..R-1); local row `rows` = rhs
    const int zrow = rows + 1;                       // all-zero row: target of out-of-range tile rows
    // ---- load the block column (columns beyond nbk are zero), the rhs block and the zero row.  Branch-free body so that
    // the unrolled loads are in flight together (the loop is L2-latency bound); entries above the diagonal are read
    // (never-written scratch) and discarded.
    if (k0 == 0) {
#pragma unroll 4
      for (int idx = tid; idx < rows * IV_NB; idx += nt) {
        const int i = idx >> 5, j = idx & 31;
        const bool use = j < nbk && j <= i;
        const double v = qb[use ? (size_t)i * (i + 1) / 2 + j : 0];
        panel[i * IV_PSTRIDE + j] = use ? v + ((i == j) ? 1.0 : 0.0) : 0.0;
      }
    } else {
      // all of a thread's loads in one batch (rows / 16 <= 24 of them for R <= 400): one L2 round trip per block column
      for (int base = tid; base < rows * IV_NB; base += 24 * nt) {
        double v[24];
#pragma unroll
        for (int u = 0; u < 24; ++u) {
          const int idx = base + u * nt, i = idx >> 5, j = idx & 31;
          v[u] = (idx < rows * IV_NB) ? __ldcg(A + (size_t)(k0 + i) * R + k0 + (j < nbk ? j : 0)) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 24; ++u) {
          const int idx = base + u * nt, i = idx >> 5, j = idx & 31;
          if (idx < rows * IV_NB) panel[i * IV_PSTRIDE + j] = (j < nbk && j <= i) ? v[u] : 0.0;
        }
      }
    }
    if (tid < 2 * IV_NB) {
      const int i = rows + (tid >> 5), j = tid & 31;
      panel[i * IV_PSTRIDE + j] = (i == rows && j < nbk) ? rhs[k0 + j] : 0.0;
    }
    __syncthreads();
    IV_LAP(1);
    // ---- diagonal block (identity padding for nbk < 32), in four 8-column sub-blocks.  The serial part of a Cholesky
    // factorisation is one rsqrt + update chain per column; here that chain runs without any communication: every lane
    // of warp 0 factors the 8 x 8 diagonal sub-block redundantly in its own registers, then lane l solves row l below it
    // against that register copy.  All warps then apply the rank-8 update to the rest of the block (two positions per
    // thread).  Two barriers per sub-block instead of one per column.  (Cycles per 32 x 32 block: 29.6 k for one warp with
    // the rows in registers and a shuffle per element, 17 k for all warps with a barrier per column, ~6 k this way.)
    {
      const int ia = tid >> 5, ja = tid & 31, ib = ia + 16, jb = ja;      // positions tid and tid + 512
      auto initial = [&](int i, int j) {
        return (i < nbk && j < nbk) ? ((j <= i) ? panel[i * IV_PSTRIDE + j] : 0.0) : ((i == j) ? 1.0 : 0.0);
      };
      s_a[ia][ja] = initial(ia, ja);
      s_a[ib][jb] = initial(ib, jb);
      s_l11[ia][ja] = 0.0;
      s_l11[ib][jb] = 0.0;
      __syncthreads();
      bool ok = true;
      IV_LAP(11);
#pragma unroll 1
      for (int kb = 0; kb < IV_NB; kb += 8) {
        if (warp == 0) {
          double d[8][8];                                // lower triangle of the sub-block, the same in every lane
#pragma unroll
          for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) d[r][c] = s_a[kb + r][kb + c];
          double x[8];                                   // this lane's row below the sub-block
          const int xi = kb + 8 + lane;
          const bool has_row = xi < IV_NB;
#pragma unroll
          for (int c = 0; c < 8; ++c) x[c] = has_row ? s_a[xi][kb + c] : 0.0;
          double dinv[8];
          IV_LAP2(0);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            if (!(d[c][c] > 0.0)) ok = false;
            const double inv = rsqrt(d[c][c]);
            dinv[c] = inv;
            d[c][c] *= inv;
#pragma unroll
            for (int r = c + 1; r < 8; ++r) d[r][c] *= inv;
#pragma unroll
            for (int c2 = c + 1; c2 < 8; ++c2)
#pragma unroll
              for (int r = c2; r < 8; ++r) d[r][c2] = fma(-d[r][c], d[c2][c], d[r][c2]);
            x[c] *= inv;
#pragma unroll
            for (int c2 = c + 1; c2 < 8; ++c2) x[c2] = fma(-x[c], d[c2][c], x[c2]);
          }
          IV_LAP2(1);
          // publish the factored sub-block: every lane holds the same values, lane 0 stores them in one straight run
          // ("lane r writes row r" took 2 k cycles per sub-block in 8 serialised branches, all lanes storing the same
          // value to the same address 0.8 k); then every lane its solved row
          if (lane == 0)
#pragma unroll
          for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int c = 0; c <= r; ++c) s_l11[kb + r][kb + c] = d[r][c];
            s_invd[kb + r] = dinv[r];
          }
          if (has_row) {
#pragma unroll
            for (int c = 0; c < 8; ++c) s_l11[xi][kb + c] = x[c];
          }
          IV_LAP2(2);
        }
        __syncthreads();
        IV_LAP(2);
        if (kb + 8 < IV_NB) {                            // rank-8 update of the remaining lower triangle
          auto update = [&](int i, int j) {
            if (j >= kb + 8 && i >= j) {
              double v = s_a[i][j];
#pragma unroll
              for (int c = 0; c < 8; ++c) v = fma(-s_l11[i][kb + c], s_l11[j][kb + c], v);
              s_a[i][j] = v;
            }
          };
          if (ia >= kb + 8) update(ia, ja);              // warp-uniform: rows above the remaining block skip the loads
          update(ib, jb);
          __syncthreads();
        }
        IV_LAP(11);
      }
      if (!ok && tid == 0) s_fail = 1;
      // all threads: the factor into the panel (it is written back to A from there), 1 / diagonal for the back substitution,
      // and the pre-scaled transposed copy for the panel solve -- kept out of warp 0's serial section (a lone warp gets one
      // shared-memory instruction through per ~9 cycles)
      auto finish = [&](int i, int j) {
        const double l = s_l11[i][j];
        s_l11t[j][i] = l * s_invd[i];
        if (i < nbk && j <= i) panel[i * IV_PSTRIDE + j] = l;
        if (j < i) sdiag[(size_t)(k0 / IV_NB) * IV_SDIAG + i * (i - 1) / 2 + j] = l * s_invd[j];
      };
      finish(ia, ja);
      finish(ib, jb);
      if (tid < nbk) invd[k0 + tid] = s_invd[tid];
    }
    __syncthreads();
    IV_LAP(2);
    if (s_fail) {                                    // both CTAs of the cluster take this exit in the same block column
      if (tid == 0) atomicExch(err, 3);
      return;
    }
    // ---- rows below the diagonal block and the rhs row: x = a L11^-T, the row in registers
    for (int i = nbk + tid; i <= rows; i += nt) {
      double *row = panel + i * IV_PSTRIDE;
      double x[IV_NB];
#pragma unroll
      for (int j = 0; j < IV_NB; ++j) x[j] = row[j] * s_invd[j];
      // column-oriented: once x[q] is final, every later entry takes its update (31 - q independent FMAs); the empty asm
      // keeps the compiler from hoisting all 496 broadcast loads of L11 to the top (it spilled 4 KB per thread)
#pragma unroll
      for (int q = 0; q < IV_NB; ++q) {
        asm volatile("" ::: "memory");
#pragma unroll
        for (int jj = 0; jj < IV_NB / 2; ++jj)
          if (2 * jj + 1 > q) {                        // the phase is bound by these broadcast loads: two entries per load
            const double2 l2 = *reinterpret_cast<const double2 *>(&s_l11t[q][2 * jj]);
            if (2 * jj > q) x[2 * jj] = fma(-x[q], l2.x, x[2 * jj]);
            x[2 * jj + 1] = fma(-x[q], l2.y, x[2 * jj + 1]);
          }
      }
#pragma unroll
      for (int j = 0; j < IV_NB; ++j) row[j] = x[j];
    }
    __syncthreads();
    IV_LAP(3);
    // ---- rhs: y block out, trailing part updated (the extra row of the trailing update)
    const double *yrow = panel + rows * IV_PSTRIDE;
    for (int j = tid; j < rows; j += nt) {
      if (j < nbk) {
        rhs[k0 + j] = yrow[j];
      } else {
        double v = rhs[k0 + j];
        const double *lr = panel + j * IV_PSTRIDE;
#pragma unroll 8
        for (int q = 0; q < IV_NB; ++q) v -= yrow[q] * lr[q];
        rhs[k0 + j] = v;
      }
    }
    // ---- write the factored block column back (rank 0; the partner holds the same values)
    if (crank == 0)
      for (int idx = tid; idx < rows * IV_NB; idx += nt) {
        const int i = idx >> 5, j = idx & 31;
        if (j < nbk && j <= i) A[(size_t)(k0 + i) * R + k0 + j] = panel[i * IV_PSTRIDE + j];
      }
    IV_LAP(4);
    // ---- trailing update A22 -= L21 L21^T on the lower triangle, on the FP64 tensor-core path (mma.sync m8n8k4: one
    // instruction = 256 FMAs of a warp; 18.6 T FMA/s measured, scripts/fp64_probe.cu): warp tiles of 16 rows x 32 columns =
    // 2 x 4 fragments, A fragment = 8 rows x 4 panel columns of -L21, B fragment = the same of the column block's rows.
    // Against the earlier DFMA tiles (8 x 128, 32 accumulators per lane) this takes 8x fewer FP64 instructions and 4x
    // fewer shared-memory loads per FMA, wastes less on the diagonal (tiles are smaller: 92 % useful instead of 73 %) and
    // needs only 16 accumulators, so the old values of the NEXT tile are prefetched from L2 during the current one.
    const int rem = rows - nbk;
    if (rem > 0) {
      const int nrb = (rem + 15) >> 4, ncb = (rem + 31) >> 5;
      int n_items = 0;
      for (int cb = 0; cb < ncb; ++cb) n_items += max(nrb - 2 * cb, 0);
      const int g = lane >> 2, tq = lane & 3;
      auto coords = [&](int it, int &i0, int &j0) {
        int cb = 0, rb = it;
        while (rb >= nrb - 2 * cb) { rb -= nrb - 2 * cb; ++cb; }
        rb += 2 * cb;
        i0 = nbk + 16 * rb;                            // panel-local rows
        j0 = nbk + 32 * cb;
      };
      // C fragment of tile (mt, nt): row i0 + 8 mt + g, columns j0 + 8 nt + 2 tq + {0, 1}
      auto load_old = [&](int i0, int j0, double (&o)[2][4][2]) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt2 = 0; nt2 < 4; ++nt2)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int li = i0 + 8 * mt + g, lj = j0 + 8 * nt2 + 2 * tq + h;
              double v = 0.0;
              if (li < rows && lj <= li)
                v = (k0 == 0) ? __ldg(qb + (size_t)li * (li + 1) / 2 + lj) + ((li == lj) ? 1.0 : 0.0)
                              : __ldcg(A + (size_t)(k0 + li) * R + k0 + lj);
              o[mt][nt2][h] = v;
            }
      };
      const int stride_it = nw * csize;
      IV_LAP(7);
      for (int it = warp * csize + crank; it < n_items; it += stride_it) {
        int i0, j0;
        coords(it, i0, j0);
        double cur[2][4][2], old[2][4][2];           // products accumulate from zero: the old values (an L2 round trip) are
        load_old(i0, j0, old);                       // only needed after the 64 MMAs
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt2 = 0; nt2 < 4; ++nt2) cur[mt][nt2][0] = cur[mt][nt2][1] = 0.0;
        int ra[2], rb4[4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) ra[mt] = ((i0 + 8 * mt + g < rows) ? i0 + 8 * mt + g : zrow) * IV_PSTRIDE + tq;
#pragma unroll
        for (int nt2 = 0; nt2 < 4; ++nt2) rb4[nt2] = ((j0 + 8 * nt2 + g < rows) ? j0 + 8 * nt2 + g : zrow) * IV_PSTRIDE + tq;
#pragma unroll
        for (int ks = 0; ks < IV_NB / 4; ++ks) {
          double af[2], bf[4];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) af[mt] = -panel[ra[mt] + 4 * ks];
#pragma unroll
          for (int nt2 = 0; nt2 < 4; ++nt2) bf[nt2] = panel[rb4[nt2] + 4 * ks];
#pragma unroll
          for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt2 = 0; nt2 < 4; ++nt2)
              asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                           : "+d"(cur[mt][nt2][0]), "+d"(cur[mt][nt2][1]) : "d"(af[mt]), "d"(bf[nt2]));
        }
        IV_LAP(8);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt2 = 0; nt2 < 4; ++nt2)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int li = i0 + 8 * mt + g, lj = j0 + 8 * nt2 + 2 * tq + h;
              if (li < rows && lj <= li) A[(size_t)(k0 + li) * R + k0 + lj] = old[mt][nt2][h] + cur[mt][nt2][h];
            }
        IV_LAP(9);
      }
    }
    IV_LAP(5);
    if (csize > 1) cluster.sync();                   // the partner's trailing-update tiles are visible (release / acquire)
    else __syncthreads();
    IV_LAP(10);
  }
  // rhs now holds y = L^-1 b.  Backward substitution L^T w = y, blocked from the last block.
  const int n_blocks = (R + IV_NB - 1) / IV_NB;
  for (int bi = n_blocks - 1; bi >= 0; --bi) {
    const int k0 = bi * IV_NB;
    const int nbk = min(IV_NB, R - k0);
    // the update of the rows above needs 32 values of L per row from L2: in flight while warp 0 runs the block's chain
    const int iu = tid - 32;
    const bool upd = tid >= 32 && iu < k0;
    double a[IV_NB];
    if (warp != 0) {
#pragma unroll
      for (int j = 0; j < IV_NB; ++j) a[j] = (upd && j < nbk) ? __ldcg(A + (size_t)(k0 + j) * R + iu) : 0.0;
    } else {
      // w_lane = y_lane / L_ll - sum_k (L[k][lane] / L_ll) w_k: the column comes pre-scaled from shared memory (sdiag), so a
      // step is one shuffle + one FMA and the block starts without an L2 round trip
      const double *sd = sdiag + (size_t)bi * IV_SDIAG;
      double lcol[IV_NB];                              // column `lane` of the diagonal block
#pragma unroll
      for (int k = 0; k < IV_NB; ++k) lcol[k] = (lane < k) ? sd[k * (k - 1) / 2 + lane] : 0.0;
      double wv = (lane < nbk) ? rhs[k0 + lane] * invd[k0 + lane] : 0.0;
#pragma unroll
      for (int k = IV_NB - 1; k >= 0; --k) {
        if (k < nbk) {
          const double wk = __shfl_sync(0xffffffffu, wv, k);
          if (lane < k) wv = fma(-lcol[k], wk, wv);
        }
      }
      if (lane < nbk) rhs[k0 + lane] = wv;
      s_blk[lane] = (lane < nbk) ? wv : 0.0;
#pragma unroll
      for (int j = 0; j < IV_NB; ++j) a[j] = 0.0;
    }
    __syncthreads();
    if (upd) {
      double v = rhs[iu];
#pragma unroll
      for (int j = 0; j < IV_NB; ++j) v -= a[j] * s_blk[j];
      rhs[iu] = v;
    }
    __syncthreads();
  }
  IV_LAP(6);
#ifdef IV_SOLVE_STATS
  if (b == 0 && (tid == 0 || tid == 511))
    printf("solve tid %d clk: init %lld load %lld diag %lld l21 %lld rhs+wb %lld trailing: setup %lld fma %lld rmw %lld rest %lld sync %lld; backsub %lld; diag init+update %lld\n",
           tid, st[0], st[1], st[2], st[3], st[4], st[7], st[8], st[9], st[5], st[10], st[6], st[11]);
  if (b == 0 && tid == 0) printf("solve warp 0 diag: loads %lld factor %lld publish %lld\n", st2[0], st2[1], st2[2]);
#endif
  if (crank == 0)
    for (int r = tid; r < R; r += nt) ivec[(size_t)b * R + r] = (float)(rhs[r] - ((r == 0) ? prior_offset : 0.0));
}

// ------------------------------------------------------------------------------------------------
// PLDA back-end, one CTA per utterance.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
plda_kernel(const float *__restrict__ ivec, const float *__restrict__ mean_vec, const float *__restrict__ lda, int lda_cols,
            const double *__restrict__ plda_T, const double *__restrict__ plda_off, const double *__restrict__ psi,
            const double *__restrict__ u_train, int R, int L, int K, double *__restrict__ scores,
            const int *__restrict__ done_flag, int text7) {
  if (done_flag && *done_flag) return;
  extern __shared__ double s_d[];
  double *u = s_d;                                   // [L]
  float *v = reinterpret_cast<float *>(s_d + L);    // [R] centred i-vector
  float *t = v + R;                                  // [L] LDA output
  __shared__ double s_red[8];
  __shared__ double s_bcast;
  const int b = blockIdx.x, tid = threadIdx.x;
  // Kaldi-exact text mode: the i-vector reaches the back-end through `ark,t:` (7 significant digits, read back as float)
  for (int r = tid; r < R; r += blockDim.x) {
    float w = ivec[(size_t)b * R + r];
    if (text7) w = (float)fb_round_sig7((double)w);
    v[r] = __fadd_rn(w, -mean_vec[r]);
  }
  __syncthreads();
  for (int l = tid; l < L; l += blockDim.x) {
    float acc = (lda_cols == R + 1) ? lda[(size_t)l * lda_cols + R] : 0.f;
    for (int r = 0; r < R; ++r) acc += lda[(size_t)l * lda_cols + r] * v[r];
    t[l] = acc;
  }
  __syncthreads();
  // ivector-normalize-length: scale to norm sqrt(L)
  double part = 0.0;
  for (int l = tid; l < L; l += blockDim.x) part += (double)t[l] * (double)t[l];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((tid & 31) == 0) s_red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += s_red[i];
    const float norm = (float)sqrt(tot);
    const float ratio = norm / sqrtf((float)L);
    s_bcast = (double)(1.0f / ratio);
  }
  __syncthreads();
  const float inv_ratio = (float)s_bcast;
  for (int l = tid; l < L; l += blockDim.x) t[l] = t[l] * inv_ratio;
  __syncthreads();
  // Plda::TransformIvector (double)
  for (int l = tid; l < L; l += blockDim.x) {
    double acc = plda_off[l];
    for (int k = 0; k < L; ++k) acc += plda_T[(size_t)l * L + k] * (double)t[k];
    u[l] = acc;
  }
  __syncthreads();
  part = 0.0;
  for (int l = tid; l < L; l += blockDim.x) part += u[l] * u[l] / (psi[l] + 1.0);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  __syncthreads();
  if ((tid & 31) == 0) s_red[tid >> 5] = part;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += s_red[i];
    s_bcast = sqrt((double)L / tot);
  }
  __syncthreads();
  const double factor = s_bcast;
  for (int l = tid; l < L; l += blockDim.x) u[l] *= factor;
  __syncthreads();
  // LogLikelihoodRatio against each enrolled speaker (n = 1); one warp per speaker
  const int w = tid >> 5, lane = tid & 31;
  for (int k = w; k < K; k += 8) {
    double given = 0.0, without = 0.0;
    for (int l = lane; l < L; l += 32) {
      const double ps = psi[l];
      const double mean = ps / (ps + 1.0) * u_train[(size_t)k * L + l];
      const double var = 1.0 + ps / (ps + 1.0);
      const double d = u[l] - mean;
      given += log(var) + d * d / var;
      const double var0 = 1.0 + ps;
      without += log(var0) + u[l] * u[l] / var0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      given += __shfl_xor_sync(0xffffffffu, given, o);
      without += __shfl_xor_sync(0xffffffffu, without, o);
    }
    if (lane == 0) {
      const double llr = -0.5 * given + 0.5 * without;
      scores[(size_t)b * K + k] = text7 ? fb_round_sig7(llr) : llr;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
static void iv_release(FbIvector *v) {
  if (!v) return;
  v->gconsts.release(); v->means_invcovars.release(); v->inv_covars.release(); v->rc_table.release();
  v->sim32.release(); v->U.release(); v->mean_vec.release(); v->lda.release(); v->plda_T.release();
  v->plda_off.release(); v->psi.release(); v->u_train.release();
  v->ll.release(); v->gsel.release(); v->post.release(); v->act_list.release(); v->gamma.release(); v->Xs.release(); v->lin_part.release();
  v->quad.release(); v->Awork.release(); v->ivec.release(); v->scores.release();
  delete v;
}

void fb_ivector_destroy(fb_ctx *ctx) {
  iv_release(ctx->iv);
  ctx->iv = nullptr;
}

static FbIvector *iv_get(fb_ctx *ctx) {
  if (!ctx->iv) ctx->iv = new FbIvector();
  return ctx->iv;
}

extern "C" int fb_load_full_gmm(fb_ctx *ctx, const float *weights, const float *means_invcovars, const float *inv_covars,
                                const float *gconsts, int C, int D) {
  FB_CHECK_ARG(ctx && weights && means_invcovars && inv_covars && gconsts, "NULL argument");
  FB_CHECK_ARG(D == FB_DIM, "feature dimension must be 72");
  FB_CHECK_ARG(C > 0 && C % FB_CHUNK_N == 0 && C <= 2048, "number of components must be a multiple of 128, at most 2048");
  FB_CUDA(cudaSetDevice(ctx->device));
  FbIvector *v = iv_get(ctx);
  v->C = C;
  // packed lower-triangular copies + (row, col) table
  std::vector<float> packed((size_t)C * IV_PACKED);
  for (int c = 0; c < C; ++c) {
    size_t e = 0;
    for (int i = 0; i < D; ++i)
      for (int j = 0; j <= i; ++j) packed[(size_t)c * IV_PACKED + e++] = inv_covars[((size_t)c * D + i) * D + j];
  }
  std::vector<unsigned short> rc(IV_PACKED);
  {
    size_t e = 0;
    for (int i = 0; i < D; ++i)
      for (int j = 0; j <= i; ++j) rc[e++] = (unsigned short)((i << 8) | j);
  }
  int rcode;
  if ((rcode = v->gconsts.ensure(C))) return rcode;
  if ((rcode = v->means_invcovars.ensure((size_t)C * D))) return rcode;
  if ((rcode = v->inv_covars.ensure(packed.size()))) return rcode;
  if ((rcode = v->rc_table.ensure(IV_PACKED))) return rcode;
  FB_CUDA(cudaMemcpy(v->gconsts.p, gconsts, C * sizeof(float), cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(v->means_invcovars.p, means_invcovars, (size_t)C * D * sizeof(float), cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(v->inv_covars.p, packed.data(), packed.size() * sizeof(float), cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(v->rc_table.p, rc.data(), IV_PACKED * sizeof(unsigned short), cudaMemcpyHostToDevice));
  // fgmm-global-to-gmm (DiagGmm::CopyFromFullGmm): invert each covariance in double (Cholesky-free Gauss-Jordan on SPD)
  std::vector<float> dw(weights, weights + C), dmiv((size_t)C * D), div_((size_t)C * D), dgc(C);
  std::vector<double> a((size_t)D * D), inv((size_t)D * D);
  for (int c = 0; c < C; ++c) {
    for (int i = 0; i < D * D; ++i) a[i] = inv_covars[(size_t)c * D * D + i];
    for (int i = 0; i < D; ++i)
      for (int j = 0; j < D; ++j) inv[(size_t)i * D + j] = (i == j) ? 1.0 : 0.0;
    for (int k = 0; k < D; ++k) {
      int piv = k;
      for (int i = k + 1; i < D; ++i)
        if (fabs(a[(size_t)i * D + k]) > fabs(a[(size_t)piv * D + k])) piv = i;
      if (fabs(a[(size_t)piv * D + k]) < 1e-300) { fb_set_error("component %d: singular inverse covariance", c); return FB_ERR_ARG; }
      if (piv != k)
        for (int j = 0; j < D; ++j) { std::swap(a[(size_t)k * D + j], a[(size_t)piv * D + j]); std::swap(inv[(size_t)k * D + j], inv[(size_t)piv * D + j]); }
      const double d = 1.0 / a[(size_t)k * D + k];
      for (int j = 0; j < D; ++j) { a[(size_t)k * D + j] *= d; inv[(size_t)k * D + j] *= d; }
      for (int i = 0; i < D; ++i)
        if (i != k) {
          const double f = a[(size_t)i * D + k];
          if (f != 0.0)
            for (int j = 0; j < D; ++j) { a[(size_t)i * D + j] -= f * a[(size_t)k * D + j]; inv[(size_t)i * D + j] -= f * inv[(size_t)k * D + j]; }
        }
    }
    double gc = log((double)weights[c]) - 0.5 * D * 1.8378770664093454835606594728112;
    for (int i = 0; i < D; ++i) {
      double mean = 0.0;
      for (int j = 0; j < D; ++j) mean += inv[(size_t)i * D + j] * (double)means_invcovars[(size_t)c * D + j];
      const float var_f = (float)inv[(size_t)i * D + i];
      const float iv_f = 1.0f / var_f;
      const float mean_f = (float)mean;
      const float miv_f = mean_f * iv_f;
      div_[(size_t)c * D + i] = iv_f;
      dmiv[(size_t)c * D + i] = miv_f;
      gc += 0.5 * log((double)iv_f) - 0.5 * (double)miv_f * (double)miv_f / (double)iv_f;
    }
    dgc[c] = (float)gc;
  }
  if ((rcode = fb_load_diag_gmm(ctx, 0, dw.data(), dmiv.data(), div_.data(), dgc.data(), C, D))) return rcode;
  if ((rcode = fb_finalize_gmms(ctx, 1))) return rcode;
  v->have_ubm = true;
  return FB_OK;
}

extern "C" int fb_load_ivector_extractor(fb_ctx *ctx, const double *M, const double *sigma_inv, double prior_offset, int C, int D,
                                         int R) {
  FB_CHECK_ARG(ctx && M && sigma_inv, "NULL argument");
  FB_CHECK_ARG(D == FB_DIM, "feature dimension must be 72");
  FB_CHECK_ARG(R >= 2 && R <= 512, "i-vector dimension must be in [2, 512]");
  FB_CUDA(cudaSetDevice(ctx->device));
  FbIvector *v = iv_get(ctx);
  FB_CHECK_ARG(!v->have_ubm || v->C == C, "extractor and full UBM disagree on the number of components");
  v->C = C;
  v->R = R;
  v->n_packed = R * (R + 1) / 2;
  v->prior_offset = prior_offset;
  double *dM = nullptr, *dS = nullptr, *dSim = nullptr;
  const size_t nM = (size_t)C * D * R, nS = (size_t)C * D * D;
  FB_CUDA(cudaMalloc(&dM, nM * sizeof(double)));
  FB_CUDA(cudaMalloc(&dS, nS * sizeof(double)));
  FB_CUDA(cudaMalloc(&dSim, nM * sizeof(double)));
  FB_CUDA(cudaMemcpy(dM, M, nM * sizeof(double), cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(dS, sigma_inv, nS * sizeof(double), cudaMemcpyHostToDevice));
  int rc;
  if ((rc = v->sim32.ensure(nM))) return rc;
  if ((rc = v->U.ensure((size_t)C * v->n_packed))) return rc;
  ivec_derive_sim_kernel<<<C, 256, 0, ctx->stream>>>(dM, dS, R, dSim, v->sim32.p);
  ivec_derive_u_kernel<<<dim3(fb_div_up(v->n_packed, 256), C), 256, 0, ctx->stream>>>(dM, dSim, R, v->n_packed, v->U.p);
  FB_CUDA(cudaGetLastError());
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  cudaFree(dM); cudaFree(dS); cudaFree(dSim);
  v->have_ie = true;
  return FB_OK;
}

// Plda::TransformIvector (normalize_length, num_examples = 1) on the host, float64
static void plda_transform_host(const FbIvector *v, const float *raw, double *out) {
  const int R = v->R, L = v->L;
  std::vector<float> c(R), t(L);
  for (int r = 0; r < R; ++r) c[r] = raw[r] - v->h_mean_vec[r];
  for (int l = 0; l < L; ++l) {
    float acc = (v->lda_cols == R + 1) ? v->h_lda[(size_t)l * v->lda_cols + R] : 0.f;
    for (int r = 0; r < R; ++r) acc += v->h_lda[(size_t)l * v->lda_cols + r] * c[r];
    t[l] = acc;
  }
  double tot = 0.0;
  for (int l = 0; l < L; ++l) tot += (double)t[l] * (double)t[l];
  const float ratio = (float)sqrt(tot) / sqrtf((float)L);
  const float inv_ratio = 1.0f / ratio;
  for (int l = 0; l < L; ++l) t[l] *= inv_ratio;
  double dot = 0.0;
  for (int l = 0; l < L; ++l) {
    double acc = v->h_plda_off[l];
    for (int k = 0; k < L; ++k) acc += v->h_plda_T[(size_t)l * L + k] * (double)t[k];
    out[l] = acc;
    dot += acc * acc / (v->h_psi[l] + 1.0);
  }
  const double f = sqrt((double)L / dot);
  for (int l = 0; l < L; ++l) out[l] *= f;
}

extern "C" int fb_load_plda_backend(fb_ctx *ctx, const float *mean_vec, const float *transform_mat, int transform_cols,
                                    const double *plda_mean, const double *plda_transform, const double *plda_psi, int R, int L) {
  FB_CHECK_ARG(ctx && mean_vec && transform_mat && plda_mean && plda_transform && plda_psi, "NULL argument");
  FB_CHECK_ARG(transform_cols == R || transform_cols == R + 1, "transform.mat must have R or R+1 columns");
  FB_CHECK_ARG(L >= 1 && L <= 512, "PLDA dimension must be in [1, 512]");
  FB_CUDA(cudaSetDevice(ctx->device));
  FbIvector *v = iv_get(ctx);
  FB_CHECK_ARG(!v->have_ie || v->R == R, "back-end and extractor disagree on the i-vector dimension");
  v->R = R;
  v->L = L;
  v->lda_cols = transform_cols;
  v->h_mean_vec.assign(mean_vec, mean_vec + R);
  v->h_lda.assign(transform_mat, transform_mat + (size_t)L * transform_cols);
  v->h_plda_T.assign(plda_transform, plda_transform + (size_t)L * L);
  v->h_psi.assign(plda_psi, plda_psi + L);
  v->h_plda_off.assign(L, 0.0);
  for (int l = 0; l < L; ++l) {
    double acc = 0.0;
    for (int k = 0; k < L; ++k) acc += plda_transform[(size_t)l * L + k] * plda_mean[k];
    v->h_plda_off[l] = -acc;
  }
  int rc;
  if ((rc = v->mean_vec.ensure(R))) return rc;
  if ((rc = v->lda.ensure(v->h_lda.size()))) return rc;
  if ((rc = v->plda_T.ensure(v->h_plda_T.size()))) return rc;
  if ((rc = v->plda_off.ensure(L))) return rc;
  if ((rc = v->psi.ensure(L))) return rc;
  FB_CUDA(cudaMemcpy(v->mean_vec.p, mean_vec, R * sizeof(float), cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(v->lda.p, v->h_lda.data(), v->h_lda.size() * sizeof(float), cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(v->plda_T.p, v->h_plda_T.data(), v->h_plda_T.size() * sizeof(double), cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(v->plda_off.p, v->h_plda_off.data(), L * sizeof(double), cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(v->psi.p, v->h_psi.data(), L * sizeof(double), cudaMemcpyHostToDevice));
  v->have_backend = true;
  return FB_OK;
}

extern "C" int fb_set_enrolled_ivectors(fb_ctx *ctx, const float *enrolled, int K) {
  FB_CHECK_ARG(ctx && enrolled && K >= 1 && K < FB_MAX_MODELS, "bad argument");
  FbIvector *v = ctx->iv;
  FB_CHECK_ARG(v && v->have_backend, "fb_load_plda_backend must be called first");
  FB_CUDA(cudaSetDevice(ctx->device));
  std::vector<double> u((size_t)K * v->L);
  for (int k = 0; k < K; ++k) plda_transform_host(v, enrolled + (size_t)k * v->R, u.data() + (size_t)k * v->L);
  int rc;
  if ((rc = v->u_train.ensure(u.size()))) return rc;
  FB_CUDA(cudaMemcpy(v->u_train.p, u.data(), u.size() * sizeof(double), cudaMemcpyHostToDevice));
  v->K = K;
  ctx->arch = 1;
  return FB_OK;
}

int fb_ivector_reserve(fb_ctx *ctx) {
  FbIvector *v = ctx->iv;
  const int B = ctx->B;
  int rc;
  if ((rc = v->ll.ensure((size_t)ctx->rows_cap * v->C))) return rc;
  if ((rc = v->gsel.ensure((size_t)ctx->rows_cap * IV_NSEL))) return rc;
  if ((rc = v->post.ensure((size_t)ctx->rows_cap * IV_NSEL))) return rc;
  if ((rc = v->gamma.ensure((size_t)B * v->C))) return rc;
  if ((rc = v->Xs.ensure((size_t)B * v->C * FB_DIM))) return rc;
  v->n_splits = ctx->num_sms < v->C ? ctx->num_sms : v->C;
  if ((rc = v->lin_part.ensure((size_t)v->n_splits * B * v->R))) return rc;
  if ((rc = v->quad.ensure((size_t)B * v->n_packed))) return rc;
  if ((rc = v->act_list.ensure((size_t)fb_div_up(B, IV_BCHUNK) * (v->C + 1)))) return rc;
  if ((rc = v->Awork.ensure((size_t)B * v->R * v->R))) return rc;
  if ((rc = v->ivec.ensure((size_t)B * v->R))) return rc;
  if ((rc = v->scores.ensure((size_t)B * (v->K > 0 ? v->K : 1)))) return rc;
  return FB_OK;
}

// i-vectors (and, when speakers are enrolled, PLDA scores) for the batch whose features are in the context.
int fb_run_ivector_flag(fb_ctx *ctx, const int *done_flag, bool with_plda) {
  FbIvector *v = ctx->iv;
  FB_CHECK_ARG(v && v->have_ubm && v->have_ie, "full UBM and i-vector extractor must be loaded");
  FB_CHECK_ARG(!with_plda || (v->have_backend && v->K > 0), "PLDA back-end / enrolled speakers missing");
  int rc;
  if ((rc = fb_ivector_reserve(ctx))) return rc;
  const int B = ctx->B;
  const int rows = ctx->total_frames;           // upper bound of the voiced rows (device knows the exact count)
  if ((rc = fb_run_gmm_store(ctx, v->ll.p, done_flag))) return rc;
  FbNvtxSeq nv;
  nv.next("fb:gselect");
  gselect_kernel<<<fb_div_up(rows, 8), 256, 0, ctx->stream>>>(v->ll.p, ctx->misc.p, v->C, v->gsel.p, done_flag);
  fb_prof_mark(ctx, 8);
  nv.next("fb:fgmm_post");
  static const bool plain_post = getenv("FB_IV_PLAIN_POST") != nullptr;       // diagnostic: the one-row-per-warp kernel
  if (plain_post) {
    fgmm_post_kernel<<<fb_div_up(rows, 8), 256, 0, ctx->stream>>>(ctx->feats_f32.p, v->gsel.p, v->gconsts.p, v->means_invcovars.p,
                                                                    v->inv_covars.p, v->rc_table.p, ctx->misc.p, v->min_post, v->post.p,
                                                                    done_flag);
  } else {
    const int n_chunks = fb_div_up(B, IV_GROUP);
    static std::atomic<unsigned long long> attr_post_mask{0};
    if (fb_once_per_device(attr_post_mask, ctx->device)) {
      FB_CUDA(cudaFuncSetAttribute(fgmm_post_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PostSmem)));
    }
    fgmm_post_group_kernel<<<ctx->max_frames * n_chunks, 256, sizeof(PostSmem), ctx->stream>>>(
        ctx->feats_f32.p, v->gsel.p, v->gconsts.p, v->means_invcovars.p, v->inv_covars.p, v->rc_table.p, ctx->frame_off.p,
        ctx->vrank.p, ctx->row_off.p, B, v->C, n_chunks, v->min_post, v->post.p, done_flag);
  }
  fb_prof_mark(ctx, 9);
  nv.next("fb:ivec_stats");
  // bucket capacity: the longest utterance, or as many frames as fit in shared memory (longer utterances go in chunks)
  const size_t smem_fixed = (size_t)(3 * v->C + 1) * sizeof(int) + 2 * sizeof(unsigned short) + 16;
  const int cap_frames = (int)((220 * 1024 - smem_fixed) / (IV_NSEL * 12));
  FB_CHECK_ARG(cap_frames >= 16, "too many UBM components for the i-vector statistics kernel");
  const int max_pairs = (ctx->max_frames < cap_frames ? ctx->max_frames : cap_frames) * IV_NSEL;
  const size_t smem_stats = (size_t)(3 * v->C + 1) * sizeof(int) + (size_t)(2 * max_pairs + 2) * sizeof(unsigned short) +
                            (size_t)2 * max_pairs * sizeof(float) + 16;
  static std::atomic<unsigned long long> attr_mask{0};
  if (fb_once_per_device(attr_mask, ctx->device)) {
    FB_CUDA(cudaFuncSetAttribute(ivec_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  }
  ivec_stats_kernel<<<dim3(B, IV_STATS_SPLIT), IV_STATS_THREADS, smem_stats, ctx->stream>>>(ctx->feats_f32.p, v->gsel.p, v->post.p, ctx->row_off.p, v->C, max_pairs,
                                                         v->gamma.p, v->Xs.p, ctx->misc.p + 1, done_flag);
  fb_prof_mark(ctx, 10);
  nv.next("fb:ivec_lin");
  const int bch = fb_div_up(B, IV_BCHUNK);
  const int lin_threads = ((4 * ((v->R + 3) / 4) + 31) / 32) * 32;
  ivec_active_kernel<<<bch, 1024, 0, ctx->stream>>>(v->gamma.p, B, v->C, v->act_list.p, done_flag);
  static const bool plain_lin = getenv("FB_IV_PLAIN_LIN") != nullptr;         // diagnostic: the register-staged kernel
  static std::atomic<unsigned long long> attr_lin_mask{0};
  if (fb_once_per_device(attr_lin_mask, ctx->device)) {
    FB_CUDA(cudaFuncSetAttribute(ivec_lin_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  }
  int n_splits = v->n_splits;                          // partial sums over the component range, added up by the solve
  if ((v->R & 3) == 0 && !plain_lin && ivec_quad_tma_smem(v->C) <= 100 * 1024) {
    const int colb = fb_div_up(v->R, IV_QUAD_COLS);
    n_splits = (2 * ctx->num_sms) / (bch * colb);      // two CTAs per SM, one wave
    if (n_splits < 1) n_splits = 1;
    if (n_splits > v->n_splits) n_splits = v->n_splits;
    ivec_lin_tma_kernel<<<dim3(bch, colb, n_splits), 256, ivec_quad_tma_smem(v->C), ctx->stream>>>(
        v->sim32.p, v->Xs.p, v->act_list.p, B, v->C, v->R, n_splits, v->lin_part.p, done_flag);
  } else {
    ivec_lin_kernel<<<dim3(v->n_splits, bch), lin_threads, 0, ctx->stream>>>(v->sim32.p, v->Xs.p, v->act_list.p, B, v->C, v->R,
                                                                            v->n_splits, v->lin_part.p, done_flag);
  }
  v->n_splits_used = n_splits;
  fb_prof_mark(ctx, 11);
  nv.next("fb:ivec_quad");
  static const bool plain_quad = getenv("FB_IV_PLAIN_QUAD") != nullptr;       // diagnostic: the register-staged kernel
  if ((v->n_packed & 3) == 0 && !plain_quad) {
    static std::atomic<unsigned long long> attr_quad_mask{0};
    if (fb_once_per_device(attr_quad_mask, ctx->device)) {
      FB_CUDA(cudaFuncSetAttribute(ivec_quad_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    }
    if (ivec_quad_tma_smem(v->C) > 100 * 1024) return FB_ERR_ARG;
    ivec_quad_tma_kernel<<<dim3(bch, fb_div_up(v->n_packed, IV_QUAD_COLS)), 256, ivec_quad_tma_smem(v->C), ctx->stream>>>(
        v->U.p, v->gamma.p, v->act_list.p, B, v->C, v->n_packed, v->quad.p, done_flag);
  } else {
    ivec_quad_kernel<<<dim3(fb_div_up(v->n_packed, 256), bch), 256, 0, ctx->stream>>>(v->U.p, v->gamma.p, v->act_list.p, B, v->C,
                                                                                     v->n_packed, v->quad.p, done_flag);
  }
  fb_prof_mark(ctx, 12);
  nv.next("fb:ivec_solve");
  const size_t smem_solve = (2 * (size_t)v->R + ((size_t)v->R + 2) * IV_PSTRIDE + (size_t)fb_div_up(v->R, IV_NB) * IV_SDIAG) * sizeof(double);
  static std::atomic<unsigned long long> attr_solve_mask{0};
  if (fb_once_per_device(attr_solve_mask, ctx->device)) {
    FB_CUDA(cudaFuncSetAttribute(ivec_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * IV_SOLVE_CLUSTER);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem_solve;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = IV_SOLVE_CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FB_CUDA(cudaLaunchKernelEx(&cfg, ivec_solve_kernel, (const double *)v->quad.p, (const double *)v->lin_part.p, v->n_splits_used, B, v->R,
                               v->n_packed, v->prior_offset, v->Awork.p, v->ivec.p, ctx->misc.p + 1, done_flag));
  }
  fb_prof_mark(ctx, 13);
  ctx->launches += 7;
  if (with_plda) {
    nv.next("fb:plda");
    const size_t smem_plda = (size_t)v->L * sizeof(double) + (size_t)(v->R + v->L) * sizeof(float) + 16;
    plda_kernel<<<B, 256, smem_plda, ctx->stream>>>(v->ivec.p, v->mean_vec.p, v->lda.p, v->lda_cols, v->plda_T.p, v->plda_off.p,
                                                    v->psi.p, v->u_train.p, v->R, v->L, v->K, v->scores.p, done_flag, ctx->kx_text ? 1 : 0);
    ctx->launches += 1;
  }
  fb_prof_mark(ctx, 14);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

static int iv_check(fb_ctx *ctx) { return fb_check_device_error(ctx); }

extern "C" int fb_score_ivector_host(fb_ctx *ctx, const int16_t *wave, const int64_t *offsets, int B, double *out_scores,
                                     float *out_ivectors) {
  FB_CHECK_ARG(ctx && wave && offsets, "NULL argument");
  FB_CHECK_ARG(offsets[0] == 0, "offsets[0] must be 0");
  FbIvector *v = ctx->iv;
  FB_CHECK_ARG(v && v->have_ubm && v->have_ie, "full UBM and i-vector extractor must be loaded");
  FB_CHECK_ARG(!out_scores || (v->have_backend && v->K > 0), "PLDA back-end / enrolled speakers missing");
  FB_CUDA(cudaSetDevice(ctx->device));
  int rc;
  ctx->need_feats_f32 = true;
  if ((rc = fb_prepare_tables(ctx))) return rc;
  if ((rc = fb_reserve_batch(ctx, B, offsets))) return rc;
  ctx->batch_tag = 0;
  if ((rc = ctx->wave.ensure((size_t)offsets[B] + 8))) return rc;
  FB_CUDA(cudaMemcpyAsync(ctx->wave.p, wave, (size_t)offsets[B] * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
  fb_prof_mark(ctx, -1);
  if ((rc = fb_run_frontend_flag(ctx, nullptr))) return rc;
  if ((rc = fb_run_ivector_flag(ctx, nullptr, out_scores != nullptr))) return rc;
  if (out_scores)
    FB_CUDA(cudaMemcpyAsync(out_scores, v->scores.p, (size_t)B * v->K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (out_ivectors)
    FB_CUDA(cudaMemcpyAsync(out_ivectors, v->ivec.p, (size_t)B * v->R * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  int rc2 = iv_check(ctx);
  if (rc2 == FB_OK && out_ivectors && ctx->kx_text)          // what a reader of ivector-extract's text archive gets
    for (size_t i = 0; i < (size_t)B * v->R; ++i) out_ivectors[i] = (float)fb_round_sig7((double)out_ivectors[i]);
  return rc2;
}

extern "C" int fb_get_ivector_stats(fb_ctx *ctx, int b, double *gamma_host, double *x_host, double *lin_host, double *quad_host) {
  FB_CHECK_ARG(ctx && ctx->iv && b >= 0 && b < ctx->B, "bad argument");
  FbIvector *v = ctx->iv;
  FB_CHECK_ARG(v->gamma.p && v->quad.p, "no i-vector batch has been scored");
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (gamma_host) FB_CUDA(cudaMemcpy(gamma_host, v->gamma.p + (size_t)b * v->C, v->C * sizeof(double), cudaMemcpyDeviceToHost));
  if (x_host)
    FB_CUDA(cudaMemcpy(x_host, v->Xs.p + (size_t)b * v->C * FB_DIM, (size_t)v->C * FB_DIM * sizeof(double), cudaMemcpyDeviceToHost));
  if (quad_host)
    FB_CUDA(cudaMemcpy(quad_host, v->quad.p + (size_t)b * v->n_packed, (size_t)v->n_packed * sizeof(double), cudaMemcpyDeviceToHost));
  if (lin_host) {
    std::vector<double> part((size_t)v->n_splits * v->R);
    for (int s = 0; s < v->n_splits_used; ++s)
      FB_CUDA(cudaMemcpy(part.data() + (size_t)s * v->R, v->lin_part.p + ((size_t)s * ctx->B + b) * v->R, v->R * sizeof(double),
                         cudaMemcpyDeviceToHost));
    for (int r = 0; r < v->R; ++r) {
      double acc = 0.0;
      for (int s = 0; s < v->n_splits_used; ++s) acc += part[(size_t)s * v->R + r];
      lin_host[r] = acc;
    }
  }
  return FB_OK;
}

extern "C" int fb_get_posteriors(fb_ctx *ctx, int32_t *gsel_host, float *post_host, int64_t capacity_rows) {
  FB_CHECK_ARG(ctx && ctx->iv && gsel_host && post_host, "bad argument");
  int misc[3];
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  FB_CUDA(cudaMemcpy(misc, ctx->misc.p, sizeof(misc), cudaMemcpyDeviceToHost));
  const int rows = misc[2];
  FB_CHECK_ARG(capacity_rows >= rows, "output buffers too small");
  FB_CUDA(cudaMemcpy(gsel_host, ctx->iv->gsel.p, (size_t)rows * IV_NSEL * sizeof(int), cudaMemcpyDeviceToHost));
  FB_CUDA(cudaMemcpy(post_host, ctx->iv->post.p, (size_t)rows * IV_NSEL * sizeof(float), cudaMemcpyDeviceToHost));
  return rows;
}
