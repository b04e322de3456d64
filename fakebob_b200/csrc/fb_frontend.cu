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
      t.mel_w[b][i] = (mel <= center) ? (mel - left) / (center - left) : (right - mel) / (right - center);
    }
  }
  t.num_mel = nb;
  for (int k = 0; k < FB_NCEPS; ++k)
    for (int n = 0; n < nb; ++n)
      t.dct[k][n] = (k == 0) ? (float)sqrt(1.0 / nb) : (float)(sqrt(2.0 / nb) * cos(M_PI / nb * (n + 0.5) * k));
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

__global__ void __launch_bounds__(MFCC_WARPS * 32)
mfcc_kernel(const int16_t *__restrict__ wave, const int64_t *__restrict__ wave_off,
            const int *__restrict__ frame_off, const FbTables *__restrict__ tb,
            float *__restrict__ mfcc, const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  __shared__ float2 s_buf[MFCC_WARPS][2][256];
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
  float *bufA = reinterpret_cast<float *>(s_buf[warp][0]);
  float *bufB = reinterpret_cast<float *>(s_buf[warp][1]);

  // ---- extract window (snip_edges=false, reflected edges), DC removal, raw log-energy
  const int start = FB_FRAME_SHIFT * t + FB_FRAME_SHIFT / 2 - FB_FRAME_LEN / 2;
  float x[13];
  float sum = 0.f;
#pragma unroll
  for (int q = 0; q < 13; ++q) {
    int i = lane + 32 * q;
    float v = 0.f;
    if (i < FB_FRAME_LEN) {
      int s = start + i;
      while (s < 0 || s >= n_samp) s = (s < 0) ? (-s - 1) : (2 * n_samp - 1 - s);
      v = (float)w[s];
    }
    x[q] = v;
    sum += v;
  }
  sum = warp_sum(sum);                       // exact: integers, |sum| < 2^24
  const float mean = sum / (float)FB_FRAME_LEN;
  float e = 0.f;
#pragma unroll
  for (int q = 0; q < 13; ++q) {
    int i = lane + 32 * q;
    if (i < FB_FRAME_LEN) {
      x[q] -= mean;
      e += x[q] * x[q];
      bufA[i] = x[q];
    }
  }
  e = warp_sum(e);
  const float log_energy = logf(fmaxf(e, 1.1920928955078125e-07f));
  __syncwarp();
  // ---- pre-emphasis + Povey window, zero-padded to 512 -> bufB (interleaved complex z[n] = x[2n] + i x[2n+1])
  const float pe = tb->preemph;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    int i = lane + 32 * q;
    float v = 0.f;
    if (i < FB_FRAME_LEN) {
      float prev = bufA[i > 0 ? i - 1 : 0];
      v = (x[q < 13 ? q : 0] - pe * prev) * __ldg(&tb->window[i]);
    }
    bufB[i] = v;
  }
  __syncwarp();
  // ---- 256-point complex FFT, radix-4 Stockham, stages Ns = 1, 4, 16, 64
  float2 *in = reinterpret_cast<float2 *>(bufB);
  float2 *out = reinterpret_cast<float2 *>(bufA);
#pragma unroll
  for (int ls = 0; ls < 8; ls += 2) {
    const int Ns = 1 << ls;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = lane + 32 * h;
      const int k = j & (Ns - 1);
      const int tws = (k << (7 - ls));       // k * (512 / (4*Ns)) in the 512-entry table (= 2 * k*64/Ns)
      float2 v0 = in[j], v1 = in[j + 64], v2 = in[j + 128], v3 = in[j + 192];
      if (ls > 0) {
        v1 = cmul(v1, s_tw[tws]);
        v2 = cmul(v2, s_tw[2 * tws]);
        v3 = cmul(v3, s_tw[3 * tws]);
      }
      float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y);
      float2 a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
      float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
      float2 d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
      float2 a3 = make_float2(d13.y, -d13.x);                 // -i * (v1 - v3)
      const int j0 = ((j - k) << 2) + k;
      out[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
      out[j0 + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
      out[j0 + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
      out[j0 + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
    }
    __syncwarp();
    float2 *tmp = in; in = out; out = tmp;
  }
  // result Z in `in` (= bufB after 4 swaps); power spectrum bins 0..255 -> `out` viewed as float[256]
  float *pw = reinterpret_cast<float *>(out);
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int k = lane + 32 * m;
    float2 zk = in[k];
    float2 zc = in[(256 - k) & 255];
    zc.y = -zc.y;
    float2 ev = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y + zc.y));
    float2 df = make_float2(zk.x - zc.x, zk.y - zc.y);
    float2 od = make_float2(0.5f * df.y, -0.5f * df.x);       // -i/2 * (zk - zc)
    float2 xo = cmul(od, s_tw[k]);
    float re = ev.x + xo.x, im = ev.y + xo.y;
    pw[k] = re * re + im * im;
  }
  __syncwarp();
  // ---- mel filterbank (lane = filter), log, DCT (lane = cepstrum), lifter, C0 <- log-energy
  float lm = 0.f;
  if (lane < tb->num_mel) {
    const int st = tb->mel_start[lane], ln = tb->mel_len[lane];
    float acc = 0.f;
    for (int i = 0; i < ln; ++i) acc += __ldg(&tb->mel_w[lane][i]) * pw[st + i];
    lm = logf(fmaxf(acc, 1.1920928955078125e-07f));
  }
  s_lm[warp][lane] = lm;
  __syncwarp();
  if (lane < FB_NCEPS) {
    float c = 0.f;
    const int nm = tb->num_mel;
    for (int m = 0; m < nm; ++m) c += __ldg(&tb->dct[lane][m]) * s_lm[warp][m];
    c *= __ldg(&tb->lifter[lane]);
    if (lane == 0) c = log_energy;
    mfcc[(int64_t)(f0 + t) * FB_NCEPS + lane] = c;
  }
}

// ------------------------------------------------------------------------------------------------
// VAD + per-utterance voiced ranks + cross-utterance row offsets (last CTA done performs the scan).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vad_scan_kernel(const float *__restrict__ mfcc, const int *__restrict__ frame_off, const FbTables *__restrict__ tb,
                int *__restrict__ vrank, int *__restrict__ nvoiced, int *__restrict__ row_off,
                int *__restrict__ misc, int B, const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  __shared__ double s_red[8];
  __shared__ int s_wtot[8];
  __shared__ int s_flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const int f0 = frame_off[b];
  const int T = frame_off[b + 1] - f0;
  double s = 0.0;
  for (int t = tid; t < T; t += 256) s += (double)mfcc[(int64_t)(f0 + t) * FB_NCEPS];
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
          num += (mfcc[(int64_t)(f0 + t2) * FB_NCEPS] > thr) ? 1 : 0;
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
    if (running == 0) atomicExch(&misc[1], 1 + b);
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
// Deltas (order 2, window 3), sliding CMN (centred window, double running sums), voiced-row
// compaction, per-dimension power-of-two scaling and the fp16 hi/lo split of [x | x^2] written
// straight into the tensor-core operand image  a_img[tile][hi|lo][slab 18][row 128][8].
// One CTA per utterance; the (T x 72) delta block lives in shared memory when it fits.
// ------------------------------------------------------------------------------------------------
#define FEATS_THREADS 576   // 8 segments x 72 dims

__device__ __forceinline__ void store_split(__half *__restrict__ a_img, int row, int d, float v) {
  v = fminf(fmaxf(v, -60000.f), 60000.f);
  const __half hi = __float2half_rn(v);
  const __half lo = __float2half_rn(v - __half2float(hi));
  const int tile = row >> 7, rr = row & 127, slab = d >> 3, e = d & 7;
  const size_t base = ((size_t)tile * 2 * FB_KSLABS + slab) * (FB_TILE_M * 8) + rr * 8 + e;
  a_img[base] = hi;
  a_img[base + (size_t)FB_KSLABS * FB_TILE_M * 8] = lo;
}

__global__ void __launch_bounds__(FEATS_THREADS)
feats_kernel(const float *__restrict__ mfcc, const int *__restrict__ frame_off, const int *__restrict__ vrank,
             const int *__restrict__ row_off, const FbTables *__restrict__ tb, __half *__restrict__ a_img,
             float *__restrict__ feats_f32, float *__restrict__ raw_global, int use_smem,
             const int *__restrict__ done_flag) {
  if (done_flag && *done_flag) return;
  extern __shared__ float s_raw[];
  const int b = blockIdx.x;
  const int f0 = frame_off[b];
  const int T = frame_off[b + 1] - f0;
  float *raw = use_smem ? s_raw : (raw_global + (size_t)f0 * FB_DIM);
  const float *mf = mfcc + (size_t)f0 * FB_NCEPS;
  // ---- phase A: [static | delta | delta-delta], float accumulation in Kaldi's tap order
  for (int idx = threadIdx.x; idx < T * FB_DIM; idx += blockDim.x) {
    const int t = idx / FB_DIM, d = idx - t * FB_DIM;
    const int k = d / FB_NCEPS, c = d - k * FB_NCEPS;
    float v;
    if (k == 0) {
      v = mf[t * FB_NCEPS + c];
    } else if (k == 1) {
      v = 0.f;
#pragma unroll
      for (int j = -3; j <= 3; ++j) {
        const float s = tb->dscale1[j + 3];
        if (s != 0.f) {
          int tt = min(max(t + j, 0), T - 1);
          v = __fadd_rn(v, __fmul_rn(s, mf[tt * FB_NCEPS + c]));
        }
      }
    } else {
      v = 0.f;
#pragma unroll
      for (int j = -6; j <= 6; ++j) {
        const float s = tb->dscale2[j + 6];
        if (s != 0.f) {
          int tt = min(max(t + j, 0), T - 1);
          v = __fadd_rn(v, __fmul_rn(s, mf[tt * FB_NCEPS + c]));
        }
      }
    }
    raw[idx] = v;
  }
  __syncthreads();
  // ---- phase B: sliding-window mean subtraction, running sums in double (SlidingWindowCmn)
  const int nseg = blockDim.x / FB_DIM;
  const int g = threadIdx.x / FB_DIM, d = threadIdx.x - g * FB_DIM;
  if (g >= nseg) return;
  const int L = (T + nseg - 1) / nseg;
  const int t0 = g * L, t1 = min(T, t0 + L);
  const int W = tb->cmn_window;
  const float scale = tb->feat_scale[d];
  const int r0 = row_off[b];
  int pws = -1, pwe = -1;
  double cur = 0.0;
  for (int t = t0; t < t1; ++t) {
    int ws = t - W / 2, we = ws + W;
    if (ws < 0) { we -= ws; ws = 0; }
    if (we > T) { ws -= (we - T); we = T; if (ws < 0) ws = 0; }
    if (pws < 0) {
      cur = 0.0;
      for (int u = ws; u < we; ++u) cur = __dadd_rn(cur, (double)raw[u * FB_DIM + d]);
    } else {
      if (ws > pws) cur = __dadd_rn(cur, -(double)raw[pws * FB_DIM + d]);
      if (we > pwe) cur = __dadd_rn(cur, (double)raw[pwe * FB_DIM + d]);
    }
    pws = ws; pwe = we;
    const int r = vrank[f0 + t];
    if (r >= 0) {
      const float alpha = __fdiv_rn(-1.0f, (float)(we - ws));
      const float o = (float)__dadd_rn((double)raw[t * FB_DIM + d], __dmul_rn((double)alpha, cur));
      const int row = r0 + r;
      if (feats_f32) feats_f32[(size_t)row * FB_DIM + d] = o;
      const float xs = o * scale;                 // exact: scale is a power of two
      store_split(a_img, row, d, xs);
      store_split(a_img, row, FB_DIM + d, xs * xs);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Host: batch reservation and launch sequence
// ------------------------------------------------------------------------------------------------
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
    if ((rc = ctx->a_img.ensure((size_t)cap_pad * 2 * FB_KSLABS * 8, true))) return rc;
    ctx->rows_cap = cap_pad;
    ctx->part.release();
    ctx->frame_ll.release();
  }
  if (ctx->n_models > 0) {
    const size_t nch = ctx->C / FB_CHUNK_N;
    if ((rc = ctx->part.ensure((size_t)ctx->n_models * nch * ctx->rows_cap))) return rc;
    if ((rc = ctx->frame_ll.ensure((size_t)ctx->n_models * ctx->rows_cap))) return rc;
    if ((rc = ctx->avg_ll.ensure((size_t)B * ctx->n_models))) return rc;
  }
  if (ctx->debug_feats)
    if ((rc = ctx->feats_f32.ensure((size_t)total_frames * FB_DIM))) return rc;
  if ((size_t)max_frames * FB_DIM * sizeof(float) > 200 * 1024)
    if ((rc = ctx->raw72.ensure((size_t)total_frames * FB_DIM))) return rc;
  FB_CUDA(cudaMemcpyAsync(ctx->wave_off.p, ctx->off_host.data(), (B + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
  FB_CUDA(cudaMemcpyAsync(ctx->frame_off.p, ctx->frame_off_host.data(), (B + 1) * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  return FB_OK;
}

int fb_run_frontend_flag(fb_ctx *ctx, const int *done_flag) {
  int rc;
  if ((rc = fb_prepare_tables(ctx))) return rc;
  const int B = ctx->B;
  dim3 g1(fb_div_up(ctx->max_frames, MFCC_WARPS), B);
  mfcc_kernel<<<g1, MFCC_WARPS * 32, 0, ctx->stream>>>(ctx->wave.p, ctx->wave_off.p, ctx->frame_off.p, ctx->tables_dev,
                                                       ctx->mfcc.p, done_flag);
  fb_prof_mark(ctx, 1);
  vad_scan_kernel<<<B, 256, 0, ctx->stream>>>(ctx->mfcc.p, ctx->frame_off.p, ctx->tables_dev, ctx->vrank.p,
                                              ctx->nvoiced.p, ctx->row_off.p, ctx->misc.p, B, done_flag);
  fb_prof_mark(ctx, 2);
  const size_t smem = (size_t)ctx->max_frames * FB_DIM * sizeof(float);
  const int use_smem = smem <= 200 * 1024;
  static size_t configured = 0;
  if (use_smem && smem > configured) {
    FB_CUDA(cudaFuncSetAttribute(feats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = 200 * 1024;
  }
  feats_kernel<<<B, FEATS_THREADS, use_smem ? smem : 0, ctx->stream>>>(
      ctx->mfcc.p, ctx->frame_off.p, ctx->vrank.p, ctx->row_off.p, ctx->tables_dev, ctx->a_img.p,
      ctx->debug_feats ? ctx->feats_f32.p : nullptr, use_smem ? nullptr : ctx->raw72.p, use_smem, done_flag);
  fb_prof_mark(ctx, 3);
  ctx->launches += 3;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fb_run_frontend(fb_ctx *ctx) { return fb_run_frontend_flag(ctx, nullptr); }
