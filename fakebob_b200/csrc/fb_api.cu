// C-ABI glue: context lifetime, configuration, batch scoring entry points and stage read-backs.
// See include/fakebob_b200.h for the reference call each entry point replaces.
#include "fb_common.cuh"
#include "fb_nes.cuh"
#include "fb_ivector.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>

static thread_local char g_err[1024] = "";
static std::atomic<uint64_t> g_epoch{1};

void fb_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
bool fb_pdl_enabled() {
  // Measured on B200 inside the captured iteration graph (gpurun session s10): 229.1 vs 230.0 us / iteration at S = 50
  // (no difference), 91.3 vs 80.1 us at S = 6 (programmatic edges are SLOWER when every kernel is launch-bound).
  // Off unless FB_PDL=1.
  static const bool on = getenv("FB_PDL") != nullptr && getenv("FB_NO_PDL") == nullptr;
  return on;
}
uint64_t fb_alloc_epoch() { return g_epoch.load(); }
void fb_bump_alloc_epoch() { g_epoch.fetch_add(1); }

int fb_run_frontend_flag(fb_ctx *ctx, const int *done_flag);
int fb_run_gmm_flag(fb_ctx *ctx, const int *done_flag);
int fb_comm_destroy_impl(fb_ctx *ctx);

extern "C" const char *fb_last_error(void) { return g_err; }
extern "C" int fb_version(void) { return 100; }

static void default_config(fb_feat_config *c) {
  c->sample_frequency = 16000.f;
  c->low_freq = 20.f;
  c->high_freq = 7600.f;
  c->num_mel_bins = 30;
  c->num_ceps = 24;
  c->preemph = 0.97f;
  c->cepstral_lifter = 22.f;
  c->vad_energy_threshold = 5.5f;
  c->vad_energy_mean_scale = 0.5f;
  c->vad_proportion_threshold = 0.12f;
  c->vad_frames_context = 2;
  c->cmn_window = 300;
}

extern "C" int fb_ctx_create(int device, fb_ctx **out) {
  FB_CHECK_ARG(out != nullptr, "out is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    fb_set_error("no CUDA device available (%s); the B200 path has no CPU fallback", cudaGetErrorString(e));
    return FB_ERR_CUDA;
  }
  FB_CHECK_ARG(device >= 0 && device < n, "device index out of range");
  FB_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  FB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    fb_set_error("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return FB_ERR_UNSUPPORTED;
  }
  fb_ctx *ctx = new fb_ctx();
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  default_config(&ctx->cfg);
  memset(&ctx->tables_host, 0, sizeof(ctx->tables_host));
  FB_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  *out = ctx;
  return FB_OK;
}

extern "C" int fb_ctx_destroy(fb_ctx *ctx) {
  if (!ctx) return FB_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  fb_nes_destroy(ctx);
  fb_ivector_destroy(ctx);
  fb_comm_destroy_impl(ctx);
  ctx->w_img.release();
  ctx->wave.release(); ctx->wave_off.release(); ctx->frame_off.release(); ctx->mfcc.release();
  ctx->vrank.release(); ctx->nvoiced.release(); ctx->row_off.release(); ctx->misc.release();
  ctx->a_img.release(); ctx->raw72.release(); ctx->cmn_prefix.release(); ctx->feats_f32.release(); ctx->part.release();
  ctx->frame_ll.release(); ctx->avg_ll.release();
  if (ctx->tables_dev) cudaFree(ctx->tables_dev);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return FB_OK;
}

extern "C" int fb_set_stream(fb_ctx *ctx, void *cuda_stream) {
  FB_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
  fb_bump_alloc_epoch();     // captured graphs are bound to the old stream's work; force re-capture
  return FB_OK;
}

extern "C" int fb_synchronize(fb_ctx *ctx) {
  FB_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  return FB_OK;
}

extern "C" int fb_set_feature_config(fb_ctx *ctx, const fb_feat_config *cfg) {
  FB_CHECK_ARG(ctx && cfg, "NULL argument");
  FB_CHECK_ARG(cfg->num_ceps == FB_NCEPS, "num_ceps must be 24");
  FB_CHECK_ARG(cfg->num_mel_bins >= FB_NCEPS && cfg->num_mel_bins <= 32, "num_mel_bins must be in [24,32]");
  FB_CHECK_ARG(cfg->sample_frequency == 16000.f, "sample_frequency must be 16000 (25 ms / 10 ms framing is compiled in)");
  FB_CHECK_ARG(cfg->cmn_window > 0 && cfg->vad_frames_context >= 0 && cfg->vad_frames_context <= 16, "bad CMN / VAD option");
  ctx->cfg = *cfg;
  ctx->tables_dirty = true;
  return FB_OK;
}

extern "C" int fb_set_kaldi_exact(fb_ctx *ctx, int compress_features, int text_precision) {
  FB_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  ctx->kx_compress = compress_features != 0;
  ctx->kx_text = text_precision != 0;
  fb_bump_alloc_epoch();          // captured NES graphs hold the old kernel sequence / arguments
  return FB_OK;
}

extern "C" int fb_set_gmm_delta_terms(fb_ctx *ctx, int terms) {
  FB_CHECK_ARG(ctx && terms >= 0 && terms <= 3, "terms must be 0 (automatic), 1, 2 or 3");
  ctx->delta_terms_req = terms;
  return FB_OK;
}

extern "C" int fb_get_gmm_info(fb_ctx *ctx, int *shared_variances, int *delta_terms, double *err_estimate) {
  FB_CHECK_ARG(ctx && ctx->n_models > 0, "no GMMs loaded (fb_finalize_gmms)");
  if (shared_variances) *shared_variances = ctx->gmm_shared ? 1 : 0;
  if (delta_terms) *delta_terms = ctx->gmm_shared ? ctx->delta_terms : 3;
  if (err_estimate) *err_estimate = ctx->gmm_shared ? ctx->delta_err_est : 0.0;
  return FB_OK;
}

extern "C" int fb_set_debug(fb_ctx *ctx, int keep_f32_features) {
  FB_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  ctx->debug_feats = keep_f32_features != 0;
  ctx->batch_tag = -1;
  fb_bump_alloc_epoch();
  return FB_OK;
}

// Device-side failure flag misc[1] (set by the kernels, 0 = none): 2 = utterance too long for a kernel's per-utterance
// capacity, 3 = matrix not positive definite, 16 + b = utterance b has no voiced frames.  Reads it on the context's stream
// (so it also waits for the enqueued work), clears it, and maps it to a distinct error code.
int fb_check_device_error(fb_ctx *ctx) {
  int code = 0;
  FB_CUDA(cudaMemcpyAsync(&code, ctx->misc.p + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  return fb_map_device_error(ctx, code);
}

int fb_map_device_error(fb_ctx *ctx, int code) {
  if (code == 0) return FB_OK;
  FB_CUDA(cudaMemsetAsync(ctx->misc.p + 1, 0, sizeof(int), ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (code == 2) {
    fb_set_error("an utterance is too long for the i-vector statistics kernel");
    return FB_ERR_TOO_LONG;
  }
  if (code == 4) {
    fb_set_error("multi-GPU exchange timed out waiting for a peer rank's partial gradient");
    return FB_ERR_NCCL;
  }
  if (code == 3) {
    fb_set_error("i-vector posterior precision matrix (quad + I) is not positive definite");
    return FB_ERR_NOT_SPD;
  }
  fb_set_error("utterance %d has no voiced frames (Kaldi's select-voiced-frames would drop it)", code - 16);
  return FB_ERR_NO_VOICED;
}

static int check_voiced(fb_ctx *ctx) { return fb_check_device_error(ctx); }

static int score_common(fb_ctx *ctx, const int64_t *offsets, int B) {
  int rc;
  FB_CHECK_ARG(ctx->n_models > 0, "no GMMs loaded (fb_load_diag_gmm + fb_finalize_gmms)");
  if ((rc = fb_prepare_tables(ctx))) return rc;
  if ((rc = fb_reserve_batch(ctx, B, offsets))) return rc;
  ctx->batch_tag = 0;
  return FB_OK;
}

extern "C" int fb_score_gmm_host(fb_ctx *ctx, const int16_t *wave, const int64_t *offsets, int B, double *out_avg_ll) {
  FB_CHECK_ARG(ctx && wave && offsets && out_avg_ll, "NULL argument");
  FB_CHECK_ARG(offsets[0] == 0, "offsets[0] must be 0");
  FB_CUDA(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = score_common(ctx, offsets, B))) return rc;
  if ((rc = ctx->wave.ensure((size_t)offsets[B] + 8))) return rc;
  FB_CUDA(cudaMemcpyAsync(ctx->wave.p, wave, (size_t)offsets[B] * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream));
  fb_prof_mark(ctx, -1);
  if ((rc = fb_run_frontend_flag(ctx, nullptr))) return rc;
  if ((rc = fb_run_gmm_flag(ctx, nullptr))) return rc;
  FB_CUDA(cudaMemcpyAsync(out_avg_ll, ctx->avg_ll.p, (size_t)B * ctx->n_models * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  return check_voiced(ctx);
}

extern "C" int fb_score_gmm_dev(fb_ctx *ctx, const int16_t *wave_dev, const int64_t *offsets_host, int B, double *out_avg_ll_dev) {
  FB_CHECK_ARG(ctx && wave_dev && offsets_host && out_avg_ll_dev, "NULL argument");
  FB_CHECK_ARG(offsets_host[0] == 0, "offsets[0] must be 0");
  FB_CUDA(cudaSetDevice(ctx->device));
  int rc;
  if ((rc = score_common(ctx, offsets_host, B))) return rc;
  // run the kernels directly on the caller's buffer
  int16_t *saved = ctx->wave.p;
  ctx->wave.p = const_cast<int16_t *>(wave_dev);
  rc = fb_run_frontend_flag(ctx, nullptr);
  ctx->wave.p = saved;
  if (rc) return rc;
  if ((rc = fb_run_gmm_flag(ctx, nullptr))) return rc;
  FB_CUDA(cudaMemcpyAsync(out_avg_ll_dev, ctx->avg_ll.p, (size_t)B * ctx->n_models * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
  return FB_OK;
}

// ---- read-backs ---------------------------------------------------------------------------------
extern "C" int fb_get_num_frames(fb_ctx *ctx, int B, int32_t *frames_host, int32_t *voiced_host) {
  FB_CHECK_ARG(ctx && B == ctx->B, "B does not match the last scored batch");
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (frames_host)
    for (int b = 0; b < B; ++b) frames_host[b] = ctx->frame_off_host[b + 1] - ctx->frame_off_host[b];
  if (voiced_host) FB_CUDA(cudaMemcpy(voiced_host, ctx->nvoiced.p, B * sizeof(int), cudaMemcpyDeviceToHost));
  return FB_OK;
}

extern "C" int fb_get_mfcc(fb_ctx *ctx, float *out_host, int64_t capacity_floats) {
  FB_CHECK_ARG(ctx && out_host, "NULL argument");
  const int64_t n = (int64_t)ctx->total_frames * FB_NCEPS;
  FB_CHECK_ARG(capacity_floats >= n, "output buffer too small");
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  FB_CUDA(cudaMemcpy(out_host, ctx->mfcc.p, n * sizeof(float), cudaMemcpyDeviceToHost));
  return FB_OK;
}

extern "C" int fb_get_vad(fb_ctx *ctx, int32_t *out_host, int64_t capacity) {
  FB_CHECK_ARG(ctx && out_host && capacity >= ctx->total_frames, "bad argument");
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  FB_CUDA(cudaMemcpy(out_host, ctx->vrank.p, (size_t)ctx->total_frames * sizeof(int), cudaMemcpyDeviceToHost));
  return FB_OK;
}

static int total_rows(fb_ctx *ctx, int *rows) {
  int misc[3];
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  FB_CUDA(cudaMemcpy(misc, ctx->misc.p, sizeof(misc), cudaMemcpyDeviceToHost));
  *rows = misc[2];
  return FB_OK;
}

extern "C" int fb_get_features(fb_ctx *ctx, float *out_host, int64_t capacity_floats) {
  FB_CHECK_ARG(ctx && out_host, "NULL argument");
  FB_CHECK_ARG(ctx->debug_feats || ctx->need_feats_f32, "fb_set_debug(ctx, 1) must be called before scoring");
  int rows = 0, rc;
  if ((rc = total_rows(ctx, &rows))) return rc;
  FB_CHECK_ARG(capacity_floats >= (int64_t)rows * FB_DIM, "output buffer too small");
  FB_CUDA(cudaMemcpy(out_host, ctx->feats_f32.p, (size_t)rows * FB_DIM * sizeof(float), cudaMemcpyDeviceToHost));
  return rows;
}

extern "C" int fb_get_frame_loglikes(fb_ctx *ctx, float *out_host, int64_t capacity_floats) {
  FB_CHECK_ARG(ctx && out_host, "NULL argument");
  int rows = 0, rc;
  if ((rc = total_rows(ctx, &rows))) return rc;
  FB_CHECK_ARG(capacity_floats >= (int64_t)rows * ctx->n_models, "output buffer too small");
  for (int m = 0; m < ctx->n_models; ++m)
    FB_CUDA(cudaMemcpy(out_host + (size_t)m * rows, ctx->frame_ll.p + (size_t)m * ctx->rows_cap, (size_t)rows * sizeof(float),
                       cudaMemcpyDeviceToHost));
  return rows;
}

// ---- profiler -----------------------------------------------------------------------------------
void fb_prof_mark(fb_ctx *ctx, int tag) {
  if (!ctx->prof_on) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, ctx->stream);
  ctx->prof_ev.push_back(e);
  ctx->prof_tag.push_back(tag);
}

static void prof_drain(fb_ctx *ctx) {
  cudaStreamSynchronize(ctx->stream);
  for (size_t i = 1; i < ctx->prof_ev.size(); ++i) {
    const int tag = ctx->prof_tag[i];
    if (tag < 0 || tag >= FB_PROF_STAGES) continue;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->prof_ev[i - 1], ctx->prof_ev[i]) == cudaSuccess) {
      ctx->prof_ms[tag] += ms;
      ctx->prof_cnt[tag] += 1;
    }
  }
  for (cudaEvent_t e : ctx->prof_ev) cudaEventDestroy(e);
  ctx->prof_ev.clear();
  ctx->prof_tag.clear();
}

extern "C" int fb_profile_enable(fb_ctx *ctx, int on) {
  FB_CHECK_ARG(ctx != nullptr, "ctx is NULL");
  prof_drain(ctx);
  ctx->prof_on = on != 0;
  if (on) {
    memset(ctx->prof_ms, 0, sizeof(ctx->prof_ms));
    memset(ctx->prof_cnt, 0, sizeof(ctx->prof_cnt));
  }
  return FB_OK;
}

extern "C" int fb_profile_read(fb_ctx *ctx, double *ms_host, int64_t *count_host) {
  FB_CHECK_ARG(ctx && ms_host && count_host, "NULL argument");
  prof_drain(ctx);
  memcpy(ms_host, ctx->prof_ms, sizeof(ctx->prof_ms));
  memcpy(count_host, ctx->prof_cnt, sizeof(ctx->prof_cnt));
  return FB_OK;
}

extern "C" int fb_get_voiced_rows(fb_ctx *ctx, int *rows) {
  FB_CHECK_ARG(ctx && rows, "NULL argument");
  return total_rows(ctx, rows);
}
