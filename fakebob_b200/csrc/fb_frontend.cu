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
#define FFT_PAD(i) ((i) + ((i) >> 2))          // float2 index padding: radix-4 Stockham stores become conflict-free
#define FFT_BUF 320                             // 256 + 64 padding

__global__ void __launch_bounds__(MFCC_WARPS * 32)
mfcc_kernel(const int16_t *__restrict__ wave, const int64_t *__restrict__ wave_off,
            const int *__restrict__ frame_off, const FbTables *__restrict__ tb,
            float *__restrict__ mfcc, const int *__restrict__ done_flag) {
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
  // ---- pre-emphasis + Povey window, zero-padded to 512 -> bufB as complex z[n] = x[2n] + i x[2n+1] (padded index)
  const float pe = tb->preemph;
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int i = lane + 32 * q;
    float v = 0.f;
    if (q < 13 && i < FB_FRAME_LEN) {
      const float prev = bufA[i > 0 ? i - 1 : 0];
      v = (x[q < 13 ? q : 0] - pe * prev) * __ldg(&tb->window[i]);
    }
    bufB[2 * FFT_PAD(i >> 1) + (i & 1)] = v;
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
      const int tws = (k << (7 - ls));       // k * 512 / (4 Ns) in the 512-entry table
      float2 v0 = in[FFT_PAD(j)], v1 = in[FFT_PAD(j + 64)], v2 = in[FFT_PAD(j + 128)], v3 = in[FFT_PAD(j + 192)];
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
      out[FFT_PAD(j0)] = make_float2(a0.x + a2.x, a0.y + a2.y);
      out[FFT_PAD(j0 + Ns)] = make_float2(a1.x + a3.x, a1.y + a3.y);
      out[FFT_PAD(j0 + 2 * Ns)] = make_float2(a0.x - a2.x, a0.y - a2.y);
      out[FFT_PAD(j0 + 3 * Ns)] = make_float2(a1.x - a3.x, a1.y - a3.y);
    }
    __syncwarp();
    float2 *tmp = in; in = out; out = tmp;
  }
  // result Z in `in` (= bufB after 4 swaps); power spectrum bins 0..255 -> `out` viewed as float[256]
  float *pw = reinterpret_cast<float *>(out);
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int k = lane + 32 * m;
    float2 zk = in[FFT_PAD(k)];
    float2 zc = in[FFT_PAD((256 - k) & 255)];
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
    for (int i = 0; i < ln; ++i) acc += __ldg(&tb->mel_w_t[i][lane]) * pw[st + i];
    lm = logf(fmaxf(acc, 1.1920928955078125e-07f));
  }
  s_lm[warp][lane] = lm;
  __syncwarp();
  if (lane < FB_NCEPS) {
    float c = 0.f;
    const int nm = tb->num_mel;
    for (int m = 0; m < nm; ++m) c += __ldg(&tb->dct_t[m][lane]) * s_lm[warp][m];
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
#define FEATS_THREADS 256
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
  if (done_flag && *done_flag) return;
  extern __shared__ double s_dyn[];
  __shared__ double s_seg[FEATS_NSEG][8];
  const int sl = blockIdx.x;                  // output slab 0..8
  const int ord = sl / 3, cg = sl - 3 * ord;
  const int b = blockIdx.y;
  const int f0 = frame_off[b];
  const int T = frame_off[b + 1] - f0;
  // P[(T+1)][8] float64 prefix sums, raw[T][8] float features of this slab
  double *P = use_smem ? s_dyn : pre_global + ((size_t)(f0 + b) * 9 + (size_t)sl * (T + 1)) * 8;
  float *raw = use_smem ? reinterpret_cast<float *>(s_dyn + (size_t)(T + 1) * 8)
                        : raw_global + ((size_t)f0 * 9 + (size_t)sl * T) * 8;
  const float *mf = mfcc + (size_t)f0 * FB_NCEPS + cg * 8;
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
      v = mf[t * FB_NCEPS + ci];
    } else if (ord == 1) {
      v = 0.f;
#pragma unroll
      for (int j = -3; j <= 3; ++j)
        if (sc1[j + 3] != 0.f) v = __fadd_rn(v, __fmul_rn(sc1[j + 3], mf[min(max(t + j, 0), T - 1) * FB_NCEPS + ci]));
    } else {
      v = 0.f;
#pragma unroll
      for (int j = -6; j <= 6; ++j)
        if (sc2[j + 6] != 0.f) v = __fadd_rn(v, __fmul_rn(sc2[j + 6], mf[min(max(t + j, 0), T - 1) * FB_NCEPS + ci]));
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
  const int W = tb->cmn_window;
  const int r0 = row_off[b];
  float scale[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) scale[i] = tb->feat_scale[ord * FB_NCEPS + cg * 8 + i];
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const int r = vrank[f0 + t];
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
    const size_t nch = ctx->C / FB_CHUNK_N;
    if ((rc = ctx->part.ensure((size_t)ctx->n_models * nch * ctx->rows_cap))) return rc;
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
  dim3 g1(fb_div_up(ctx->max_frames, MFCC_WARPS), B);
  mfcc_kernel<<<g1, MFCC_WARPS * 32, 0, ctx->stream>>>(ctx->wave.p, ctx->wave_off.p, ctx->frame_off.p, ctx->tables_dev,
                                                       ctx->mfcc.p, done_flag);
  fb_prof_mark(ctx, 1);
  vad_scan_kernel<<<B, 256, 0, ctx->stream>>>(ctx->mfcc.p, ctx->frame_off.p, ctx->tables_dev, ctx->vrank.p,
                                              ctx->nvoiced.p, ctx->row_off.p, ctx->misc.p, B, done_flag);
  fb_prof_mark(ctx, 2);
  const size_t smem = fb_feats_smem_bytes(ctx->max_frames);
  const int use_smem = smem <= FB_FEATS_SMEM_MAX;
  static bool configured = false;
  if (!configured) {
    FB_CUDA(cudaFuncSetAttribute(feats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FB_FEATS_SMEM_MAX));
    configured = true;
  }
  feats_kernel<<<dim3(9, B), FEATS_THREADS, use_smem ? smem : 0, ctx->stream>>>(
      ctx->mfcc.p, ctx->frame_off.p, ctx->vrank.p, ctx->row_off.p, ctx->tables_dev, ctx->a_img.p,
      (ctx->debug_feats || ctx->need_feats_f32) ? ctx->feats_f32.p : nullptr, use_smem ? nullptr : ctx->raw72.p,
      use_smem ? nullptr : ctx->cmn_prefix.p, use_smem, done_flag);
  fb_prof_mark(ctx, 3);
  ctx->launches += 3;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

int fb_run_frontend(fb_ctx *ctx) { return fb_run_frontend_flag(ctx, nullptr); }
