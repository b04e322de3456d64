// Front-end kernels: MFCC, energy VAD + voiced-row scan, delta / sliding-CMN / voiced compaction
// with the fp16 hi/lo operand packing for the tensor-core GMM kernel.
//
// Replaces the Kaldi binaries the reference runs per score() call:
//   compute-mfcc-feats  (gmm_ubm_kaldiHelper.py:138)      -> mfcc_kernel
//   compute-vad         (gmm_ubm_kaldiHelper.py:158)      -> vad_scan_kernel
//   add-deltas | apply-cmvn-sliding | select-voiced-frames (gmm_ubm_kaldiHelper.py:195-198) -> feats_kernel
// Upstream arithmetic: SURVEY.md Appendix A.2-A.6.
#include "fb_common.cuh"
#include <math.h>
#include <string.h>

// ------------------------------------------------------------------------------------------------
// Host: tables
// ------------------------------------------------------------------------------------------------
static float mel_scale_f(float f) { return 1127.0f * logf(1.0f + f / 700.0f); }

int fb_prepare_tables(fb_ctx *ctx) {
  if (!ctx->tables_dirty) return FB_OK;
  const fb_feat_config &c = ctx->cfg;
  FB_CHECK_ARG(c.num_ceps == FB_NCEPS, "num_ceps must be 24");
  FB_CHECK_ARG(c.num_mel_bins >= c.num_ceps && c.num_mel_bins <= 32, "num_mel_bins must be in [24,32]");
  FbTables &t = ctx->tables_host;
  // keep feat_scale (set by fb_finalize_gmms) across re-preparation
  float keep_scale[FB_DIM];
  memcpy(keep_scale, t.feat_scale, sizeof(keep_scale));
  memset(&t, 0, sizeof(t));
  memcpy(t.feat_scale, keep_scale, sizeof(keep_scale));
  for (int d = 0; d < FB_DIM; ++d)
    if (t.feat_scale[d] == 0.f) t.feat_scale[d] = 1.f;
  const double a = 2.0 * M_PI / (FB_FRAME_LEN - 1);
  for (int i = 0; i < FB_FRAME_LEN; ++i) t.window[i] = (float)pow(0.5 - 0.5 * cos(a * i), 0.85);
  for (int q = 0; q < FB_FFT_N; ++q) {
    double ang = -2.0 * M_PI * q / FB_FFT_N;
    t.tw512[q] = make_float2((float)cos(ang), (float)sin(ang));
  }
  // mel banks (mel-computations.cc, no VTLN), float arithmetic like Kaldi's BaseFloat
  const int nb = c.num_mel_bins;
  const int n_fft_bins = FB_FFT_N / 2;
  const float fft_bin_width = c.sample_frequency / FB_FFT_N;
  const float mel_low = mel_scale_f(c.low_freq), mel_high = mel_scale_f(c.high_freq);
  const float delta = (mel_high - mel_low) / (float)(nb + 1);
  for (int b = 0; b < nb; ++b) {
    const float left = mel_low + b * delta, center = mel_low + (b + 1) * delta, right = mel_low + (b + 2) * delta;
    int first = -1, last = -1;
    for (int i = 0; i < n_fft_bins; ++i) {
      float mel = mel_scale_f(fft_bin_width * i);
      if (mel > left && mel < right) {
        if (first < 0) first = i;
        last = i;
      }
    }
    if (first < 0) { first = 0; last = -1; }
    int len = last - first + 1;
    if (len > FB_MEL_MAXLEN) { fb_set_error("mel filter %d spans %d bins (> %d)", b, len, FB_MEL_MAXLEN); return FB_ERR_UNSUPPORTED; }
    t.mel_start[b] = first;
    t.mel_len[b] = len;
    for (int i = 0; i < len; ++i) {
      float mel = mel_scale_f(fft_bin_width * (first + i));
      t.mel_w_t[i][b] = (mel <= center) ? (mel - left) / (center - left) : (right - mel) / (right - center);
    }
  }
  t.num_mel = nb;
  for (int k = 0; k < FB_NCEPS; ++k)
    for (int n = 0; n < nb; ++n)
      t.dct_t[n][k] = (k == 0) ? (float)sqrt(1.0 / nb) : (float)(sqrt(2.0 / nb) * cos(M_PI / nb * (n + 0.5) * k));
  for (int k = 0; k < FB_NCEPS; ++k) t.lifter[k] = (float)(1.0 + 0.5 * c.cepstral_lifter * sin(M_PI * k / c.cepstral_lifter));
  // delta scales (feature-functions.cc DeltaFeatures ctor), window 3, order 2, float arithmetic
  {
    float s0[1] = {1.f};
    float s1[7] = {0}, s2[13] = {0};
    float normalizer = 0.f;
    for (int j = -3; j <= 3; ++j) { normalizer += (float)(j * j); s1[j + 3] += (float)j * s0[0]; }
    for (int i = 0; i < 7; ++i) s1[i] = s1[i] * (1.0f / normalizer);
    normalizer = 0.f;
    for (int j = -3; j <= 3; ++j) {
      normalizer += (float)(j * j);
      for (int k = -3; k <= 3; ++k) s2[j + k + 6] += (float)j * s1[k + 3];
    }
    for (int i = 0; i < 13; ++i) s2[i] = s2[i] * (1.0f / normalizer);
    memcpy(t.dscale1, s1, sizeof(s1));
    memcpy(t.dscale2, s2, sizeof(s2));
  }
  t.preemph = c.preemph;
  t.vad_thr = c.vad_energy_threshold;
  t.vad_mean_scale = c.vad_energy_mean_scale;
  t.vad_prop = c.vad_proportion_threshold;
  t.vad_ctx = c.vad_frames_context;
  t.cmn_window = c.cmn_window;
  if (!ctx->tables_dev) FB_CUDA(cudaMalloc(&ctx->tables_dev, sizeof(FbTables)));
  FB_CUDA(cudaMemcpyAsync(ctx->tables_dev, &t, sizeof(FbTables), cudaMemcpyHostToDevice, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->tables_dirty = false;
  return FB_OK;
}

// ------------------------------------------------------------------------------------------------
// MFCC: one warp per frame, 8 frames per CTA.  512-point real FFT as a 256-point complex
// radix-4 Stockham transform in shared memory (two 2 KB ping-pong buffers per warp).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#define MFCC_WARPS 8
// float2 index paddings of the two FFT buffers, chosen per buffer so that EVERY access pattern of the three stages is
// conflict-free (a warp's 32 float2 = 2 wavefronts; one slot per 8 elements for both buffers cost 3 on five of the seven
// access groups: ncu counted 42 % of the kernel's shared-memory wavefronts as bank conflicts):
//   buffer A: stage-A stores (index 8 lane + c), unit-stride loads / stores (lane + 32 r)      -> one slot per 16
//   buffer B: stage-B stores (index 32 (lane >> 2) + (lane & 3) + 4 r), unit-stride loads      -> four slots per 32
#define FFT_PAD_A(i) ((i) + ((i) >> 4))
#define FFT_PAD_B(i) ((i) + 4 * ((i) >> 5))
#define FFT_BUF 320                             // >= 256 + 32 padding

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }            // -i * a

// forward 4-point DFT, natural order
__device__ __forceinline__ void dft4(float2 &v0, float2 &v1, float2 &v2, float2 &v3) {
  const float2 s0 = cadd(v0, v2), s1 = csub(v0, v2), s2 = cadd(v1, v3), s3 = mul_mi(csub(v1, v3));
  v0 = cadd(s0, s2); v1 = cadd(s1, s3); v2 = csub(s0, s2); v3 = csub(s1, s3);
}
// forward 8-point DFT, natural order: y[r] = a[r] + W8^r b[r], y[r+4] = a[r] - W8^r b[r]
__device__ __forceinline__ void dft8(float2 *u) {
  float2 a0 = u[0], a1 = u[2], a2 = u[4], a3 = u[6], b0 = u[1], b1 = u[3], b2 = u[5], b3 = u[7];
  dft4(a0, a1, a2, a3);
  dft4(b0, b1, b2, b3);
  const float h = 0.70710678118654752f;
  const float2 t1 = make_float2(h * (b1.x + b1.y), h * (b1.y - b1.x));        // b1 * (1 - i)/sqrt2
  const float2 t2 = mul_mi(b2);                                                // b2 * -i
  const float2 t3 = make_float2(h * (b3.y - b3.x), -h * (b3.x + b3.y));       // b3 * (-1 - i)/sqrt2
  u[0] = cadd(a0, b0); u[4] = csub(a0, b0);
  u[1] = cadd(a1, t1); u[5] = csub(a1, t1);
  u[2] = cadd(a2, t2); u[6] = csub(a2, t2);
  u[3] = cadd(a3, t3); u[7] = csub(a3, t3);
}

// One warp per frame.  Lane l holds samples 128 r + 4 l + k (r = 0..3 rounds, k = 0..3), i.e. exactly the inputs of the
// two first-stage radix-4 butterflies j = 2l, 2l+1 of the 256-point complex FFT (256 = 4 x 8 x 8), so framing, DC removal,
// pre-emphasis, windowing and the first FFT stage never touch shared memory.
__global__ void __launch_bounds__(MFCC_WARPS * 32)
mfcc_kernel(const int16_t *__restrict__ wave, const int64_t *__restrict__ wave_off,
            const int *__restrict__ frame_off, const FbTables *__restrict__ tb,
            float *__restrict__ mfcc, const int *__restrict__ done_flag) {
  FB_GRID_DEP_SYNC();
  if (done_flag && *done_flag) return;
  __shared__ float2 s_buf[MFCC_WARPS][2][FFT_BUF];
  __shared__ float2 s_tw[FB_FFT_N];
  __shared__ float s_lm[MFCC_WARPS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < FB_FFT_N; i += blockDim.x) s_tw[i] = tb->tw512[i];
  __syncthreads();
  const int f0 = frame_off[b];
  const int T = frame_off[b + 1] - f0;
  const int t = blockIdx.x * MFCC_WARPS + warp;
  if (t >= T) return;
  const int64_t w0 = wave_off[b];
  const int n_samp = (int)(wave_off[b + 1] - w0);
  const int16_t *w = wave + w0;
  float2 *bufA = s_buf[warp][0];
  float2 *bufB = s_buf[warp][1];

  // ---- extract window (snip_edges=false, reflected edges)
  const int start = FB_FRAME_SHIFT * t + FB_FRAME_SHIFT / 2 - FB_FRAME_LEN / 2;
  float x[4][4];
  const bool interior = start >= 0 && start + FB_FRAME_LEN <= n_samp && ((w0 + start) & 3) == 0;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i0 = 128 * r + 4 * lane;
    if (i0 < FB_FRAME_LEN) {
      if (interior) {
        const short4 q = *reinterpret_cast<const short4 *>(w + start + i0);
        x[r][0] = (float)q.x; x[r][1] = (float)q.y; x[r][2] = (float)q.z; x[r][3] = (float)q.w;
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          int sidx = start + i0 + k;
          while (sidx < 0 || sidx >= n_samp) sidx = (sidx < 0) ? (-sidx - 1) : (2 * n_samp - 1 - sidx);
          x[r][k] = (float)w[sidx];
        }
      }
    } else {
      x[r][0] = x[r][1] = x[r][2] = x[r][3] = 0.f;
    }
  }
  // ---- DC removal, raw log-energy
  float sum = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) sum += (x[r][0] + x[r][1]) + (x[r][2] + x[r][3]);
  sum = warp_sum(sum);                       // exact: integers, |sum| < 2^24
  const float mean = sum / (float)FB_FRAME_LEN;
  float e = 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r)
    if (128 * r + 4 * lane < FB_FRAME_LEN) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { x[r][k] -= mean; e += x[r][k] * x[r][k]; }
    }
  e = warp_sum(e);
  const float log_energy = logf(fmaxf(e, 1.1920928955078125e-07f));
  // ---- pre-emphasis (needs the previous sample: neighbouring lane / previous round) + Povey window
  const float pe = tb->preemph;
  float y[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const float up = __shfl_up_sync(0xffffffffu, x[r][3], 1);
    const float wrap = __shfl_sync(0xffffffffu, x[r > 0 ? r - 1 : 0][3], 31);
    const float prev0 = (lane > 0) ? up : ((r > 0) ? wrap : x[0][0]);
    const int i0 = 128 * r + 4 * lane;
    if (i0 < FB_FRAME_LEN) {
      const float4 wn = __ldg(reinterpret_cast<const float4 *>(&tb->window[i0]));
      y[r][0] = (x[r][0] - pe * prev0) * wn.x;
      y[r][1] = (x[r][1] - pe * x[r][0]) * wn.y;
      y[r][2] = (x[r][2] - pe * x[r][1]) * wn.z;
      y[r][3] = (x[r][3] - pe * x[r][2]) * wn.w;
    } else {
      y[r][0] = y[r][1] = y[r][2] = y[r][3] = 0.f;
    }
  }
  // ---- FFT stage A (radix 4, Ns = 1), in registers: butterflies j = 2 lane and 2 lane + 1; out[4 j + r]
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float2 v0 = make_float2(y[0][2 * h], y[0][2 * h + 1]), v1 = make_float2(y[1][2 * h], y[1][2 * h + 1]);
    float2 v2 = make_float2(y[2][2 * h], y[2][2 * h + 1]), v3 = make_float2(y[3][2 * h], y[3][2 * h + 1]);
    dft4(v0, v1, v2, v3);
    const int o = 4 * (2 * lane + h);
    bufA[FFT_PAD_A(o)] = v0; bufA[FFT_PAD_A(o + 1)] = v1; bufA[FFT_PAD_A(o + 2)] = v2; bufA[FFT_PAD_A(o + 3)] = v3;
  }
  __syncwarp();
  // ---- stage B (radix 8, Ns = 4): butterfly j = lane, k = j & 3, twiddle exp(-2 pi i r k / 32)
  {
    const int k = lane & 3;
    float2 u[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) u[r] = bufA[FFT_PAD_A(lane + 32 * r)];
#pragma unroll
    for (int r = 1; r < 8; ++r) u[r] = cmul(u[r], s_tw[r * k * 16]);
    dft8(u);
    const int j0 = ((lane >> 2) << 5) + k;
#pragma unroll
    for (int r = 0; r < 8; ++r) bufB[FFT_PAD_B(j0 + 4 * r)] = u[r];
  }
  __syncwarp();
  // ---- stage C (radix 8, Ns = 32): k = lane, twiddle exp(-2 pi i r k / 256); out[k + 32 r]
  {
    float2 u[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) u[r] = bufB[FFT_PAD_B(lane + 32 * r)];
#pragma unroll
    for (int r = 1; r < 8; ++r) u[r] = cmul(u[r], s_tw[r * lane * 2]);
    dft8(u);
#pragma unroll
    for (int r = 0; r < 8; ++r) bufA[FFT_PAD_A(lane + 32 * r)] = u[r];
  }
  __syncwarp();
  // ---- real-FFT post-processing -> power spectrum bins 0..255 in bufB viewed as float[256]
  float *pw = reinterpret_cast<float *>(bufB);
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int k = lane + 32 * m;
    const float2 zk = bufA[FFT_PAD_A(k)];
    float2 zc = bufA[FFT_PAD_A((256 - k) & 255)];
    zc.y = -zc.y;
    const float2 ev = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
    const float2 df = csub(zk, zc);
    const float2 od = make_float2(0.5f * df.y, -0.5f * df.x);       // -i/2 * (zk - zc)
    const float2 xo = cmul(od, s_tw[k]);
    const float re = ev.x + xo.x, im = ev.y + xo.y;
    pw[k] = re * re + im * im;
  }
  __syncwarp();
  // ---- mel filterbank (lane = filter), log, DCT (lane = cepstrum), lifter, C0 <- log-energy
  float lm = 0.f;
  if (lane < tb->num_mel) {
    const int st = tb->mel_start[lane], ln = tb->mel_len[lane];
    float acc = 0.f;
#pragma unroll 4
    for (int i = 0; i < ln; ++i) acc += __ldg(&tb->mel_w_t[i][lane]) * pw[st + i];
    lm = logf(fmaxf(acc, 1.1920928955078125e-07f));
  }
  s_lm[warp][lane] = lm;
  __syncwarp();
  if (lane < FB_NCEPS) {
    float c = 0.f;
    const int nm = tb->num_mel;
#pragma unroll 6
    for (int m = 0; m < nm; ++m) c += __ldg(&tb->dct_t[m][lane]) * s_lm[warp][m];
    c *= __ldg(&tb->lifter[lane]);
    if (lane == 0) c = log_energy;
    mfcc[(int64_t)(f0 + t) * FB_NCEPS + lane] = c;
  }
}

// ------------------------------------------------------------------------------------------------
// Kaldi-exact mode: CompressedMatrix (speech-feature format) round trip of one utterance's MFCC matrix, in place.
// matrix/compressed-matrix.cc: global min / range; per column the values at sorted positions 0, T/4, 3(T/4), T-1 as
// uint16 (forced strictly increasing); every value as one byte, piecewise linear between those four points.
// One CTA per utterance; the order statistics by rank counting (T ~ 500, not a hot path).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int cm_float_to_u16(float vmin, float vrange, float v) {
  float f = __fdiv_rn(__fadd_rn(v, -vmin), vrange);
  f = fminf(fmaxf(f, 0.f), 1.f);
  return (int)__fadd_rn(__fmul_rn(f, 65535.0f), 0.499f);
}
__device__ __forceinline__ float cm_u16_to_float(float vmin, float vrange, int u) {
  return __fadd_rn(vmin, __fmul_rn(__fmul_rn(vrange, 1.52590218966964e-05f), (float)u));
}

__global__ void __launch_bounds__(256)
mfcc_compress_kernel(float *__restrict__ mfcc, const int *__restrict__ frame_off, int *__restrict__ err, const int *__restrict__ done_flag) {
  FB_GRID_DEP_SYNC();
  if (done_flag && *done_flag) return;
  __shared__ float s_mn[8], s_mx[8];
  __shared__ float s_q[4];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int f0 = frame_off[b];
  const int T = frame_off[b + 1] - f0;
  float *M = mfcc + (size_t)f0 * FB_NCEPS;
  if (T <= 8) {                                   // Kaldi would pick the two-byte format: not implemented
    if (tid == 0) atomicExch(err, 2);
    return;
  }
  float mn = INFINITY, mx = -INFINITY;
  for (int i = tid; i < T * FB_NCEPS; i += 256) { const float v = M[i]; mn = fminf(mn, v); mx = fmaxf(mx, v); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o)); mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); }
  if (lane == 0) { s_mn[warp] = mn; s_mx[warp] = mx; }
  __syncthreads();
  mn = s_mn[0]; mx = s_mx[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) { mn = fminf(mn, s_mn[i]); mx = fmaxf(mx, s_mx[i]); }
  if (mx == mn) mx = __fadd_rn(mn, __fadd_rn(1.0f, fabsf(mn)));
  const float vmin = mn, vrange = __fadd_rn(mx, -mn);
  const int q = T / 4;
  for (int d = 0; d < FB_NCEPS; ++d) {
    __syncthreads();
    // the values at sorted positions 0, q, 3q, T-1: an element's position = #smaller + #equal with a lower index
    for (int i = tid; i < T; i += 256) {
      const float v = M[(size_t)i * FB_NCEPS + d];
      int rank = 0;
      for (int j = 0; j < T; ++j) {
        const float w = M[(size_t)j * FB_NCEPS + d];
        rank += (w < v || (w == v && j < i)) ? 1 : 0;
      }
      if (rank == 0) s_q[0] = v;
      if (rank == q) s_q[1] = v;
      if (rank == 3 * q) s_q[2] = v;
      if (rank == T - 1) s_q[3] = v;
    }
    __syncthreads();
    const int p0 = min(cm_float_to_u16(vmin, vrange, s_q[0]), 65532);
    const int p25 = min(max(cm_float_to_u16(vmin, vrange, s_q[1]), p0 + 1), 65533);
    const int p75 = min(max(cm_float_to_u16(vmin, vrange, s_q[2]), p25 + 1), 65534);
    const int p100 = max(cm_float_to_u16(vmin, vrange, s_q[3]), p75 + 1);
    const float g0 = cm_u16_to_float(vmin, vrange, p0), g25 = cm_u16_to_float(vmin, vrange, p25);
    const float g75 = cm_u16_to_float(vmin, vrange, p75), g100 = cm_u16_to_float(vmin, vrange, p100);
    __syncthreads();                              // every rank pass of this column has finished reading it
    for (int i = tid; i < T; i += 256) {
      const float v = M[(size_t)i * FB_NCEPS + d];
      int c;
      if (v < g25) {
        const float f = __fdiv_rn(__fadd_rn(v, -g0), __fadd_rn(g25, -g0));
        c = min(max((int)__fadd_rn(__fmul_rn(f, 64.f), 0.5f), 0), 64);
      } else if (v < g75) {
        const float f = __fdiv_rn(__fadd_rn(v, -g25), __fadd_rn(g75, -g25));
        c = 64 + min(max((int)__fadd_rn(__fmul_rn(f, 128.f), 0.5f), 0), 128);
      } else {
        const float f = __fdiv_rn(__fadd_rn(v, -g75), __fadd_rn(g100, -g75));
        c = 192 + min(max((int)__fadd_rn(__fmul_rn(f, 63.f), 0.5f), 0), 63);
      }
      float o;
      if (c <= 64) o = __fadd_rn(g0, __fmul_rn(__fmul_rn(__fadd_rn(g25, -g0), (float)c), 1.0f / 64.0f));
      else if (c <= 192) o = __fadd_rn(g25, __fmul_rn(__fmul_rn(__fadd_rn(g75, -g25), (float)(c - 64)), 1.0f / 128.0f));
      else o = __fadd_rn(g75, __fmul_rn(__fmul_rn(__fadd_rn(g100, -g75), (float)(c - 192)), 1.0f / 63.0f));
      M[(size_t)i * FB_NCEPS + d] = o;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// VAD + per-utterance voiced ranks + cross-utterance row offsets (last CTA done performs the scan).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vad_scan_kernel(const float *__restrict__ mfcc, const int *__restrict__ frame_off, const FbTables *__restrict__ tb,
                int *__restrict__ vrank, int *__restrict__ nvoiced, int *__restrict__ row_off,
                int *__restrict__ misc, int B, int c0_cap, const int *__restrict__ done_flag) {
  FB_GRID_DEP_SYNC();
  if (done_flag && *done_flag) return;
  __shared__ double s_red[8];
  __shared__ int s_wtot[8];
  __shared__ int s_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const int f0 = frame_off[b];
  const int T = frame_off[b + 1] - f0;
  extern __shared__ float s_c0[];                       // [min(T, cap)] log-energies of this utterance
  const int cap = c0_cap;
  double s = 0.0;
  for (int t = tid; t < T; t += 256) {
    const float v = mfcc[(int64_t)(f0 + t) * FB_NCEPS];
    if (t < cap) s_c0[t] = v;
    s += (double)v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) s_red[warp] = s;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += s_red[i];
  float thr = tb->vad_thr;
  if (tb->vad_mean_scale != 0.f) thr = thr + __fdiv_rn(__fmul_rn(tb->vad_mean_scale, (float)tot), (float)T);
  const int ctx = tb->vad_ctx;
  const float prop = tb->vad_prop;
  int running = 0;
  for (int base = 0; base < T; base += 256) {
    const int t = base + tid;
    int v = 0;
    if (t < T) {
      int num = 0, den = 0;
      for (int t2 = t - ctx; t2 <= t + ctx; ++t2)
        if (t2 >= 0 && t2 < T) {
          den++;
          const float c0v = (t2 < cap) ? s_c0[t2] : mfcc[(int64_t)(f0 + t2) * FB_NCEPS];
          num += (c0v > thr) ? 1 : 0;
        }
      v = ((float)num >= __fmul_rn((float)den, prop)) ? 1 : 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, v);
    if (lane == 0) s_wtot[warp] = __popc(bal);
    __syncthreads();
    int woff = 0, ctot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int c = s_wtot[i];
      if (i < warp) woff += c;
      ctot += c;
    }
    if (t < T) vrank[f0 + t] = v ? (running + woff + __popc(bal & ((1u << lane) - 1u))) : -1;
    running += ctot;
    __syncthreads();
  }
  if (tid == 0) {
    nvoiced[b] = running;
    if (running == 0) atomicExch(&misc[1], 16 + b);
    __threadfence();
    int ticket = atomicAdd(&misc[0], 1);
    s_flag = (ticket == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  // exclusive scan of nvoiced over utterances
  int run = 0;
  for (int base = 0; base < B; base += 256) {
    const int i = base + tid;
    int c = (i < B) ? ((volatile int *)nvoiced)[i] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    int woff = 0, ctot = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      int cc = s_wtot[k];
      if (k < warp) woff += cc;
      ctot += cc;
    }
    if (i < B) row_off[i] = run + woff + incl - c;
    run += ctot;
    __syncthreads();
  }
  if (tid == 0) {
    row_off[B] = run;
    misc[2] = run;
    misc[0] = 0;
  }
}

// ------------------------------------------------------------------------------------------------
// Deltas (order 2, window 3), sliding CMN (centred window, float64 prefix sums), voiced-row compaction,
// per-dimension power-of-two scaling and the fp16 hi/lo split of [x | x^2], written straight into the
// tensor-core operand image  a_img[tile][hi: 20 slabs | lo: 20 slabs][row 128][8 halfs]  (slab order: x, ones, x^2, zero).
// One CTA per (utterance, output slab): slab sl = 3*order + cepstra-group holds dims d = 24*order + 8*cg + 0..7,
// so each CTA owns x-slab sl and x^2-slab 9+sl and every store is a full 16-byte row chunk.
// ------------------------------------------------------------------------------------------------
#define FEATS_THREADS 512
#define FEATS_NSEG (FEATS_THREADS / 8)

__device__ __forceinline__ void split_pack8(const float *v, uint4 &hi, uint4 &lo) {
  __half h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float c = fminf(fmaxf(v[i], -60000.f), 60000.f);
    h[i] = __float2half_rn(c);
    l[i] = __float2half_rn(c - __half2float(h[i]));
  }
  hi = *reinterpret_cast<uint4 *>(h);
  lo = *reinterpret_cast<uint4 *>(l);
}

__global__ void __launch_bounds__(FEATS_THREADS)
feats_kernel(const float *__restrict__ mfcc, const int *__restrict__ frame_off, const int *__restrict__ vrank,
             const int *__restrict__ row_off, const FbTables *__restrict__ tb, __half *__restrict__ a_img,
             float *__restrict__ feats_f32, float *__restrict__ raw_global, double *__restrict__ pre_global,
             int use_smem, const int *__restrict__ done_flag) {
  FB_GRID_DEP_SYNC();
  if (done_flag && *done_flag) return;
  extern __shared__ double s_dyn[];
  __shared__ double s_seg[FEATS_NSEG][8];
  const int sl = blockIdx.x;                  // output slab 0..8
  const int ord = sl / 3, cg = sl - 3 * ord;
  const int b = blockIdx.y;
  const int f0 = frame_off[b];
  const int T = frame_off[b + 1] - f0;
  // issue every small global read now so their latencies overlap (they are consumed in the last phase)
  const int W = tb->cmn_window;
  const int r0 = row_off[b];
  float scale[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) scale[i] = tb->feat_scale[ord * FB_NCEPS + cg * 8 + i];
  const int vr_first = (threadIdx.x < T) ? vrank[f0 + threadIdx.x] : -1;
  // P[(T+1)][8] float64 prefix sums, raw[T][8] float features of this slab, stat[T][8] the cepstra group's statics
  double *P = use_smem ? s_dyn : pre_global + ((size_t)(f0 + b) * 9 + (size_t)sl * (T + 1)) * 8;
  float *raw = use_smem ? reinterpret_cast<float *>(s_dyn + (size_t)(T + 1) * 8)
                        : raw_global + ((size_t)f0 * 9 + (size_t)sl * T) * 8;
  const float *mfg = mfcc + (size_t)f0 * FB_NCEPS + cg * 8;
  const float *mf = mfg;
  int mstride = FB_NCEPS;
  if (use_smem) {
    float *stat = reinterpret_cast<float *>(P);       // aliases the prefix-sum area, which is only written after phase A
    for (int idx = threadIdx.x; idx < T * 2; idx += blockDim.x) {          // 2 float4 per frame
      const int t = idx >> 1, h = idx & 1;
      reinterpret_cast<float4 *>(stat)[idx] = *reinterpret_cast<const float4 *>(mfg + (size_t)t * FB_NCEPS + 4 * h);
    }
    __syncthreads();
    mf = stat;
    mstride = 8;
  }
  // ---- phase A: this slab's static / delta / delta-delta values, float accumulation in Kaldi's tap order
  float sc1[7], sc2[13];
#pragma unroll
  for (int j = 0; j < 7; ++j) sc1[j] = tb->dscale1[j];
#pragma unroll
  for (int j = 0; j < 13; ++j) sc2[j] = tb->dscale2[j];
  for (int idx = threadIdx.x; idx < T * 8; idx += blockDim.x) {
    const int t = idx >> 3, ci = idx & 7;
    float v;
    if (ord == 0) {
      v = mf[t * mstride + ci];
    } else if (ord == 1) {
      v = 0.f;
#pragma unroll
      for (int j = -3; j <= 3; ++j)
        if (sc1[j + 3] != 0.f) v = __fadd_rn(v, __fmul_rn(sc1[j + 3], mf[min(max(t + j, 0), T - 1) * mstride + ci]));
    } else {
      v = 0.f;
#pragma unroll
      for (int j = -6; j <= 6; ++j)
        if (sc2[j + 6] != 0.f) v = __fadd_rn(v, __fmul_rn(sc2[j + 6], mf[min(max(t + j, 0), T - 1) * mstride + ci]));
    }
    raw[idx] = v;
  }
  __syncthreads();
  // ---- phase B: float64 exclusive prefix sums over frames (segment sums -> scan of segments -> local prefixes)
  const int ci = threadIdx.x & 7, g = threadIdx.x >> 3;
  const int L = (T + FEATS_NSEG - 1) / FEATS_NSEG;
  const int t0 = min(g * L, T), t1 = min(t0 + L, T);
  double acc = 0.0;
  for (int t = t0; t < t1; ++t) acc = __dadd_rn(acc, (double)raw[t * 8 + ci]);
  s_seg[g][ci] = acc;
  __syncthreads();
  double run = 0.0;
  for (int q = 0; q < g; ++q) run = __dadd_rn(run, s_seg[q][ci]);
  for (int t = t0; t < t1; ++t) {
    P[t * 8 + ci] = run;
    run = __dadd_rn(run, (double)raw[t * 8 + ci]);
  }
  if (t1 == T && t0 < T) P[T * 8 + ci] = run;
  if (T == 0) return;
  __syncthreads();
  // ---- phase C: x - window mean, scale, split, pack; one thread per voiced frame (8 dims = one 16-byte chunk)
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int r = (t == (int)threadIdx.x) ? vr_first : vrank[f0 + t];
    if (r < 0) continue;
    int ws = t - W / 2, we = ws + W;
    if (ws < 0) { we -= ws; ws = 0; }
    if (we > T) { ws -= (we - T); we = T; if (ws < 0) ws = 0; }
    const float alpha = __fdiv_rn(-1.0f, (float)(we - ws));
    float xs[8], x2[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const double cur = __dadd_rn(P[we * 8 + i], -P[ws * 8 + i]);
      o[i] = (float)__dadd_rn((double)raw[t * 8 + i], __dmul_rn((double)alpha, cur));
      xs[i] = o[i] * scale[i];                 // exact: power-of-two scale
      x2[i] = xs[i] * xs[i];
    }
    const int row = r0 + r;
    if (feats_f32) {
      float4 *dst = reinterpret_cast<float4 *>(feats_f32 + (size_t)row * FB_DIM + ord * FB_NCEPS + cg * 8);
      dst[0] = make_float4(o[0], o[1], o[2], o[3]);
      dst[1] = make_float4(o[4], o[5], o[6], o[7]);
    }
    const int tile = row >> 7, rr = row & 127;
    uint4 hi, lo;
    __half *tbase = a_img + (size_t)tile * FB_A_TILE_SLABS * (FB_TILE_M * 8) + rr * 8;
    split_pack8(xs, hi, lo);
    *reinterpret_cast<uint4 *>(tbase + (size_t)sl * (FB_TILE_M * 8)) = hi;
    *reinterpret_cast<uint4 *>(tbase + (size_t)(FB_A_HI_SLABS + sl) * (FB_TILE_M * 8)) = lo;
    split_pack8(x2, hi, lo);
    *reinterpret_cast<uint4 *>(tbase + (size_t)(FB_SLAB_X2 + sl) * (FB_TILE_M * 8)) = hi;
    *reinterpret_cast<uint4 *>(tbase + (size_t)(FB_A_HI_SLABS + FB_SLAB_X2 + sl) * (FB_TILE_M * 8)) = lo;
  }
}

// The "ones" slab (hi slab 9) of every tile: columns 0..2 = 1.0 so the three fp16 terms of gconst in W add up in the MMA.
__global__ void a_img_init_kernel(__half *a_img, int n_tiles) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_tiles * FB_TILE_M) return;
  const int tile = idx / FB_TILE_M, rr = idx - tile * FB_TILE_M;
  __half *p = a_img + ((size_t)tile * FB_A_TILE_SLABS + FB_SLAB_ONES) * (FB_TILE_M * 8) + rr * 8;
  const __half one = __float2half_rn(1.f), zero = __float2half_rn(0.f);
  p[0] = one; p[1] = one; p[2] = one;
#pragma unroll
  for (int i = 3; i < 8; ++i) p[i] = zero;
}

// ------------------------------------------------------------------------------------------------
// Host: batch reservation and launch sequence
// ------------------------------------------------------------------------------------------------
#define FB_FEATS_SMEM_MAX (200 * 1024)
static size_t fb_feats_smem_bytes(int max_frames) {
  return (size_t)(max_frames + 1) * 8 * sizeof(double) + (size_t)max_frames * 8 * sizeof(float);
}

int fb_reserve_batch(fb_ctx *ctx, int B, const int64_t *offsets_host) {
  FB_CHECK_ARG(B > 0, "B must be positive");
  FB_CHECK_ARG(offsets_host != nullptr, "offsets is NULL");
  int rc;
  ctx->off_host.assign(offsets_host, offsets_host + B + 1);
  ctx->frame_off_host.resize(B + 1);
  int total_frames = 0, max_frames = 0;
  for (int b = 0; b < B; ++b) {
    int64_t n = offsets_host[b + 1] - offsets_host[b];
    FB_CHECK_ARG(n > 0 && n < (1ll << 30), "utterance length out of range");
    int T = (int)((n + FB_FRAME_SHIFT / 2) / FB_FRAME_SHIFT);
    FB_CHECK_ARG(T > 0, "utterance shorter than half a frame shift");
    ctx->frame_off_host[b] = total_frames;
    total_frames += T;
    if (T > max_frames) max_frames = T;
  }
  ctx->frame_off_host[B] = total_frames;
  ctx->B = B;
  ctx->total_samples = offsets_host[B];
  ctx->total_frames = total_frames;
  ctx->max_frames = max_frames;
  if ((rc = ctx->wave_off.ensure(B + 1))) return rc;
  if ((rc = ctx->frame_off.ensure(B + 1))) return rc;
  if ((rc = ctx->mfcc.ensure((size_t)total_frames * FB_NCEPS))) return rc;
  if ((rc = ctx->vrank.ensure(total_frames))) return rc;
  if ((rc = ctx->nvoiced.ensure(B))) return rc;
  if ((rc = ctx->row_off.ensure(B + 1))) return rc;
  if ((rc = ctx->misc.ensure(8, true))) return rc;
  const int rows_pad = ((total_frames + 255) / 256) * 256;
  if (rows_pad > ctx->rows_cap) {
    const int cap = rows_pad + rows_pad / 8;
    const int cap_pad = ((cap + 255) / 256) * 256;
    ctx->a_img.release();
    const int n_tiles = cap_pad / FB_TILE_M;
    if ((rc = ctx->a_img.ensure((size_t)n_tiles * FB_A_TILE_SLABS * FB_TILE_M * 8, true))) return rc;
    a_img_init_kernel<<<fb_div_up((int64_t)n_tiles * FB_TILE_M, 256), 256, 0, ctx->stream>>>(ctx->a_img.p, n_tiles);
    FB_CUDA(cudaGetLastError());
    ctx->rows_cap = cap_pad;
    ctx->part.release();
    ctx->frame_ll.release();
  }
  if (ctx->n_models > 0) {
    const size_t nst = ctx->C / FB_STAGE_N;          // one partial slot per 64-column stage (gmm_umma_kernel segments)
    if ((rc = ctx->part.ensure((size_t)ctx->n_models * nst * FB_GMM_EPI_HALVES * ctx->rows_cap))) return rc;
    if ((rc = ctx->frame_ll.ensure((size_t)ctx->n_models * ctx->rows_cap))) return rc;
    if ((rc = ctx->avg_ll.ensure((size_t)B * ctx->n_models))) return rc;
  }
  if (ctx->debug_feats || ctx->need_feats_f32)
    if ((rc = ctx->feats_f32.ensure((size_t)total_frames * FB_DIM))) return rc;
  if (fb_feats_smem_bytes(max_frames) > FB_FEATS_SMEM_MAX) {
    if ((rc = ctx->raw72.ensure((size_t)total_frames * FB_DIM))) return rc;
    if ((rc = ctx->cmn_prefix.ensure((size_t)(total_frames + B) * FB_DIM))) return rc;
  }
  FB_CUDA(cudaMemcpyAsync(ctx->wave_off.p, ctx->off_host.data(), (B + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  FB_CUDA(cudaMemcpyAsync(ctx->frame_off.p, ctx->frame_off_host.data(), (B + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  return FB_OK;
}

int fb_run_frontend_flag(fb_ctx *ctx, const int *done_flag) {
  int rc;
  if ((rc = fb_prepare_tables(ctx))) return rc;
  const int B = ctx->B;
  FbNvtxSeq nv;
  nv.next("fb:mfcc");
  dim3 g1(fb_div_up(ctx->max_frames, MFCC_WARPS), B);
  FB_CUDA(fb_launch(mfcc_kernel, g1, dim3(MFCC_WARPS * 32), 0, ctx->stream, ctx->wave.p, ctx->wave_off.p, ctx->frame_off.p,
                    ctx->tables_dev, ctx->mfcc.p, done_flag));
  if (ctx->kx_compress) {      // Kaldi-exact mode: copy-feats --compress=true round trip before anything reads the MFCCs
    nv.next("fb:mfcc_compress");
    FB_CUDA(fb_launch(mfcc_compress_kernel, dim3(B), dim3(256), 0, ctx->stream, ctx->mfcc.p, ctx->frame_off.p, ctx->misc.p + 1, done_flag));
    ctx->launches += 1;
  }
  fb_prof_mark(ctx, 1);
  nv.next("fb:vad_scan");
  const int c0_cap = ctx->max_frames < 8192 ? ctx->max_frames : 8192;
  FB_CUDA(fb_launch(vad_scan_kernel, dim3(B), dim3(256), (size_t)c0_cap * sizeof(float), ctx->stream, ctx->mfcc.p, ctx->frame_off.p,
                    ctx->tables_dev, ctx->vrank.p, ctx->nvoiced.p, ctx->row_off.p, ctx->misc.p, B, c0_cap, done_flag));
  fb_prof_mark(ctx, 2);
  nv.next("fb:feats");
  const size_t smem = fb_feats_smem_bytes(ctx->max_frames);
  const int use_smem = smem <= FB_FEATS_SMEM_MAX;
  static std::atomic<unsigned long long> configured_mask{0};
  if (fb_once_per_device(configured_mask, ctx->device)) {
    FB_CUDA(cudaFuncSetAttribute(feats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_FEATS_SMEM_MAX));
  }
  FB_CUDA(fb_launch(feats_kernel, dim3(9, B), dim3(FEATS_THREADS), use_smem ? smem : 0, ctx->stream, ctx->mfcc.p, ctx->frame_off.p,
                    ctx->vrank.p, ctx->row_off.p, ctx->tables_dev, ctx->a_img.p,
                    (ctx->debug_feats || ctx->need_feats_f32) ? ctx->feats_f32.p : nullptr, use_smem ? nullptr : ctx->raw72.p,
                    use_smem ? nullptr : ctx->cmn_prefix.p, use_smem, done_flag));
  fb_prof_mark(ctx, 3);
  ctx->launches += 3;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fb_run_frontend(fb_ctx *ctx) { return fb_run_frontend_flag(ctx, nullptr); }
