// Shared declarations for libfakebob_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include <string>
#include <vector>

#include "../../include/fakebob_b200.h"

#define FB_FRAME_LEN   400
#define FB_FRAME_SHIFT 160
#define FB_FFT_N       512
#define FB_NCEPS       24
#define FB_DIM         72      // 24 x (static, delta, delta-delta)
#define FB_KSLABS      18      // K = 144 = [x | x^2] in slabs of 8 fp16
// Operand K layout (slabs of 8 fp16), identical for the hi and the lo half of A and W:
//   [x: 0..8] [ones / gconst: 9] [x^2: 10..18] [zero: 19]      -> the x half and the x^2 half are 5 k-blocks (K=80) each
#define FB_A_HI_SLABS  20
#define FB_A_TILE_SLABS 40     // hi 20 + lo 20 slabs per 128-row tile
#define FB_W_HI_SLABS  20
#define FB_SLAB_ONES   9
#define FB_SLAB_X2     10
#define FB_TILE_M      128
#define FB_STAGE_N     64      // W columns per smem stage
#define FB_CHUNK_N     128     // accumulator columns per (tile, unit)
#define FB_MAX_MODELS  32
#ifndef FB_GMM_EPI_HALVES
#define FB_GMM_EPI_HALVES 2    // column ranges per accumulator in the GMM epilogue (fb_gmm.cu): partials are [model][stage][range][row]
#endif
#define FB_MEL_MAXLEN  48

void fb_set_error(const char *fmt, ...);
uint64_t fb_alloc_epoch();
void fb_bump_alloc_epoch();

#define FB_CUDA(call)                                                                        \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      fb_set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      return FB_ERR_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

#define FB_CHECK_ARG(cond, msg)                 \
  do {                                          \
    if (!(cond)) {                              \
      fb_set_error("bad argument: %s", msg);    \
      return FB_ERR_ARG;                        \
    }                                           \
  } while (0)

// Device-resident front-end tables (built once per feature config).
struct FbTables {
  float  window[FB_FRAME_LEN];
  float2 tw512[FB_FFT_N];                 // exp(-2 pi i q / 512)
  int    mel_start[32];
  int    mel_len[32];
  float  mel_w_t[FB_MEL_MAXLEN][32];      // [bin offset][filter]: lanes (filters) read consecutive words
  float  dct_t[32][32];                   // [mel bin][cepstrum]; lifter applied separately like Kaldi
  float  lifter[FB_NCEPS];
  float  dscale1[7];                      // delta scales (window 3)
  float  dscale2[13];
  float  feat_scale[FB_DIM];              // power-of-two per-dimension scale applied before the fp16 split
  int    num_mel;
  float  preemph;
  float  vad_thr, vad_mean_scale, vad_prop;
  int    vad_ctx;
  int    cmn_window;
};

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  int ensure(size_t want, bool zero = false) {
    if (want <= n) return FB_OK;
    if (p) cudaFree(p);
    p = nullptr; n = 0;
    fb_bump_alloc_epoch();
    size_t cap = want + want / 8 + 64;
    FB_CUDA(cudaMalloc(&p, cap * sizeof(T)));
    if (zero) {
      FB_CUDA(cudaMemset(p, 0, cap * sizeof(T)));
      FB_CUDA(cudaDeviceSynchronize());          // the context's stream is non-blocking w.r.t. the legacy stream
    }
    n = cap;
    return FB_OK;
  }
  void release() { if (p) { cudaFree(p); fb_bump_alloc_epoch(); } p = nullptr; n = 0; }
};

struct FbHostGmm {
  std::vector<float> weights, means_invvars, inv_vars, gconsts;
  bool loaded = false;
};

struct FbNes;   // fb_nes.cu
struct FbIvector;  // fb_ivector.cu
struct FbComm;  // fb_comm.cu

struct fb_ctx {
  int device = 0;
  int num_sms = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  fb_feat_config cfg;
  bool tables_dirty = true;
  FbTables *tables_dev = nullptr;
  FbTables tables_host;

  // models
  FbHostGmm host_gmm[FB_MAX_MODELS];
  int n_models = 0, C = 0;
  int delta_terms_req = 0;     // fb_set_gmm_delta_terms: 0 = automatic
  int delta_terms = 3;         // fp16 product terms of the (slot m - slot 0) sub-stages actually in use (shared-variance mode)
  DevBuf<__half> w_img;        // general: [model][C/64][hi 20 | lo 20 slabs][64][8]; shared-variance: [C/64][x^2 | slot 0 | slot m - slot 0 ...]
  bool gmm_shared = false;     // all models share inv_vars (MAP mean-only adaptation): x^2 contraction done once
  double delta_err_est = 0.0;  // predicted per-frame error of a one-term difference product (fb_finalize_gmms)

  // batch workspace (shared by score() calls [tag 0] and the NES state [tag 1])
  int batch_tag = -1;
  int B = 0;
  int64_t total_samples = 0;
  int total_frames = 0, max_frames = 0;
  bool debug_feats = false;
  bool kx_compress = false;    // fb_set_kaldi_exact: CompressedMatrix round trip of the MFCCs
  bool kx_text = false;        // fb_set_kaldi_exact: 7-significant-digit text round trip of scores / i-vectors
  DevBuf<int16_t> wave;
  DevBuf<int64_t> wave_off;    // B+1
  DevBuf<int>     frame_off;   // B+1
  DevBuf<float>   mfcc;        // [total_frames][24]
  DevBuf<int>     vrank;       // [total_frames]
  DevBuf<int>     nvoiced;     // [B]
  DevBuf<int>     row_off;     // [B+1]
  DevBuf<int>     misc;        // [0]=ticket, [1]=error flag, [2]=M (total voiced rows)
  DevBuf<__half>  a_img;       // [n_tiles][hi 19 | lo 18 slabs][128][8]
  DevBuf<float>   raw72;       // global fallback scratch for long utterances
  DevBuf<double>  cmn_prefix;  // idem (float64 prefix sums)
  DevBuf<float>   feats_f32;   // [rows][72] when debug
  DevBuf<float2>  part;        // [model][C/128][rows_pad] (max, sum) in log2 domain
  DevBuf<float>   frame_ll;    // [model][rows_pad]
  DevBuf<double>  avg_ll;      // [B][n_models]
  std::vector<int64_t> off_host;
  std::vector<int> frame_off_host;
  int rows_cap = 0;            // padded row capacity of a_img / part (multiple of 256)

  // per-stage profiler
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  std::vector<int> prof_tag;      // stage id closed by event i (event 0 of a sequence has tag -1)
  double prof_ms[FB_PROF_STAGES] = {0};
  int64_t prof_cnt[FB_PROF_STAGES] = {0};

  int arch = 0;                // 0 = GMM-UBM scoring, 1 = i-vector / PLDA scoring
  bool need_feats_f32 = false; // the i-vector path consumes the float32 features
  FbIvector *iv = nullptr;
  FbNes *nes = nullptr;
  FbComm *comm = nullptr;
  int64_t launches = 0;
};

// fb_frontend.cu
int fb_prepare_tables(fb_ctx *ctx);
int fb_reserve_batch(fb_ctx *ctx, int B, const int64_t *offsets_host);
int fb_run_frontend(fb_ctx *ctx);             // mfcc -> vad_scan -> feats (wave already on device)
// fb_gmm.cu
int fb_run_gmm(fb_ctx *ctx);                  // gmm -> reduce into avg_ll
void fb_prof_mark(fb_ctx *ctx, int tag);
int fb_check_device_error(fb_ctx *ctx);       // reads + clears misc[1] on the stream; distinct error codes
int fb_map_device_error(fb_ctx *ctx, int code);
// helpers
// cudaFuncSetAttribute is per device: one flag per device ordinal (a process may hold contexts on several GPUs)
static inline bool fb_once_per_device(std::atomic<unsigned long long> &mask, int device) {
  const unsigned long long bit = 1ull << (device & 63);
  return (mask.fetch_or(bit) & bit) == 0;
}
static inline int fb_div_up(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Kaldi's text output stream has precision(7): what survives writing a value and reading it back
// (round to 7 significant decimal digits; double arithmetic, exact powers of ten up to 1e22)
#ifdef __CUDACC__
__host__ __device__ static inline double fb_round_sig7(double v) {
  if (v == 0.0 || !(fabs(v) < 1e300)) return v;
  int e = (int)floor(log10(fabs(v)));
  int p = 6 - e;
  if (p > 22 || p < -22) return v;
  const double s = pow(10.0, (double)(p < 0 ? -p : p));
  double r = (p >= 0) ? rint(v * s) / s : rint(v / s) * s;
  // log10 may land one decade off right at a power of ten: the result then has 8 or 6 digits, fix by one retry
  if (fabs(r) >= pow(10.0, (double)(e + 1))) { ++e; p = 6 - e; const double s2 = pow(10.0, (double)(p < 0 ? -p : p)); r = (p >= 0) ? rint(v * s2) / s2 : rint(v / s2) * s2; }
  return r;
}
#endif

// ---- NVTX ranges around the enqueue of every stage (header-only NVTX3: no cost unless a tool is attached) ---------------
#include <nvtx3/nvToolsExt.h>
struct FbNvtxSeq {
  bool open = false;
  void next(const char *name) {
    if (open) nvtxRangePop();
    nvtxRangePushA(name);
    open = true;
  }
  ~FbNvtxSeq() { if (open) nvtxRangePop(); }
};

// ---- programmatic dependent launch ---------------------------------------------------------------------------------
// With FB_PDL=1 every kernel of the scoring / NES sequence is launched with the programmatic-stream-serialization attribute;
// all of them start with FB_GRID_DEP_SYNC(): the next kernel's CTAs are scheduled (launch latency, prologue) while the
// previous kernel drains, and block at griddepcontrol.wait until it has completed and flushed its memory.  The wait is
// executed unconditionally, before any early return, so completion stays transitive along the chain.  Measured inside the
// captured iteration graph it does not pay (equal at S = 50, 14 % slower at S = 6), so it is opt-in; without the attribute
// griddepcontrol.wait is a no-op.
bool fb_pdl_enabled();
#ifdef __CUDACC__
#define FB_GRID_DEP_SYNC()                                          \
  do {                                                              \
    asm volatile("griddepcontrol.wait;" ::: "memory");              \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)
template <typename... KArgs, typename... Args>
static inline cudaError_t fb_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = fb_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif
