// Speaker enrolment on the device: MAP mean-only adaptation of the diagonal UBM.
//
// Replaces, per enrolled speaker, the two Kaldi processes the reference starts in build_spk_models.py:202-219
//     gmm-global-acc-stats --update-flags=m final.dubm <feats> acc   |   gmm-global-est-map --update-flags=m final.dubm acc out
// (the second one is the reference's own native file, gmm-global-est-map.cc: MapDiagGmmUpdate with MapDiagGmmOptions,
// mean_tau = 10).  Upstream arithmetic (gmm/mle-diag-gmm.cc):
//     post[t,c]  = softmax_c(loglike[t,c])                        float  (DiagGmm::ComponentPosteriors -> ApplySoftMax)
//     occ[c]     = sum_t post[t,c]                                 double (AccumDiagGmm::AccumulateFromPosteriors)
//     macc[c,:]  = sum_t post[t,c] x[t,:]                          double, frame order
//     mean'[c,:] = (macc[c,:] + tau mean[c,:]) / (occ[c] + tau)    double (MapDiagGmmUpdate), variances and weights kept
//     gconst'    = DiagGmm::ComputeGconsts
// The component log-likelihoods come from the tcgen05 GMM kernel in STORE mode (fb_gmm.cu); the front-end is the scoring one.
#include "fb_common.cuh"
#include <math.h>

int fb_run_frontend_flag(fb_ctx *ctx, const int *done_flag);
int fb_run_gmm_store(fb_ctx *ctx, float *ll_out, const int *done_flag);

// One warp per voiced row: in-place softmax over the C component log-likelihoods (float, like VectorBase<float>::ApplySoftMax).
__global__ void __launch_bounds__(256)
map_post_kernel(float *__restrict__ ll, const int *__restrict__ misc, int C) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= misc[2]) return;
  float *p = ll + (size_t)row * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, p[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float e = expf(p[c] - mx);
    p[c] = e;
    s += e;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.0f / s;
  for (int c = lane; c < C; c += 32) p[c] *= inv;
}

// One CTA per component, thread d < 72 owns macc[c][d], thread 72 owns occ[c]; rows in frame order, unfused double
// multiply-add (the order and rounding of Kaldi's per-frame AddVecVec).
__global__ void __launch_bounds__(96)
map_acc_kernel(const float *__restrict__ post, const float *__restrict__ feats, const int *__restrict__ misc, int C,
               double *__restrict__ occ, double *__restrict__ macc) {
  const int c = blockIdx.x, d = threadIdx.x;
  if (d > FB_DIM) return;
  const int M = misc[2];
  double acc = 0.0;
  for (int t = 0; t < M; ++t) {
    const double pt = (double)post[(size_t)t * C + c];
    const double x = (d < FB_DIM) ? (double)feats[(size_t)t * FB_DIM + d] : 1.0;
    acc = __dadd_rn(acc, __dmul_rn(pt, x));
  }
  if (d < FB_DIM) macc[(size_t)c * FB_DIM + d] = acc;
  else occ[c] = acc;
}

extern "C" int fb_map_adapt_host(fb_ctx *ctx, const int16_t *wave, const int64_t *offsets, int B, double mean_tau,
                                 float *out_means_invvars, float *out_gconsts, double *out_occupancy) {
  FB_CHECK_ARG(ctx && wave && offsets && out_means_invvars && out_gconsts, "NULL argument");
  FB_CHECK_ARG(offsets[0] == 0, "offsets[0] must be 0");
  FB_CHECK_ARG(ctx->n_models == 1 && !ctx->gmm_shared, "load the UBM alone into slot 0 (fb_load_diag_gmm + fb_finalize_gmms(ctx, 1))");
  FB_CHECK_ARG(mean_tau >= 0.0, "mean_tau must be non-negative");
  FB_CUDA(cudaSetDevice(ctx->device));
  int rc;
  const bool saved_need = ctx->need_feats_f32;
  ctx->need_feats_f32 = true;
  const int C = ctx->C;
  if ((rc = fb_prepare_tables(ctx)) || (rc = fb_reserve_batch(ctx, B, offsets))) { ctx->need_feats_f32 = saved_need; return rc; }
  ctx->batch_tag = 0;
  DevBuf<float> ll;
  DevBuf<double> acc;
  auto done = [&](int code) {
    ctx->need_feats_f32 = saved_need;
    cudaStreamSynchronize(ctx->stream);
    ll.release();
    acc.release();
    return code;
  };
  if ((rc = ctx->wave.ensure((size_t)offsets[B] + 8))) return done(rc);
  if ((rc = ll.ensure((size_t)ctx->rows_cap * C))) return done(rc);
  if ((rc = acc.ensure((size_t)C * (FB_DIM + 1)))) return done(rc);
  if (cudaMemcpyAsync(ctx->wave.p, wave, (size_t)offsets[B] * sizeof(int16_t), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
    return done(FB_ERR_CUDA);
  if ((rc = fb_run_frontend_flag(ctx, nullptr))) return done(rc);
  if ((rc = fb_run_gmm_store(ctx, ll.p, nullptr))) return done(rc);
  map_post_kernel<<<fb_div_up(ctx->total_frames, 8), 256, 0, ctx->stream>>>(ll.p, ctx->misc.p, C);
  map_acc_kernel<<<C, 96, 0, ctx->stream>>>(ll.p, ctx->feats_f32.p, ctx->misc.p, C, acc.p, acc.p + C);
  ctx->launches += 2;
  if (cudaGetLastError() != cudaSuccess) { fb_set_error("enrolment kernel launch failed"); return done(FB_ERR_CUDA); }
  std::vector<double> h((size_t)C * (FB_DIM + 1));
  int misc[3];
  if (cudaMemcpyAsync(h.data(), acc.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaMemcpyAsync(misc, ctx->misc.p, sizeof(misc), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
      cudaStreamSynchronize(ctx->stream) != cudaSuccess) {
    fb_set_error("enrolment read-back failed: %s", cudaGetErrorString(cudaGetLastError()));
    return done(FB_ERR_CUDA);
  }
  if (misc[1] != 0) return done(fb_map_device_error(ctx, misc[1]));
  // MapDiagGmmUpdate (means only) + CopyToDiagGmm + ComputeGconsts, double arithmetic, float storage
  const FbHostGmm &g = ctx->host_gmm[0];
  const double *occ = h.data(), *macc = h.data() + C;
  const double log2pi = 1.8378770664093454835606594728112;
  for (int c = 0; c < C; ++c) {
    double gc = log((double)g.weights[c]) - 0.5 * FB_DIM * log2pi;
    for (int d = 0; d < FB_DIM; ++d) {
      const double iv = (double)g.inv_vars[(size_t)c * FB_DIM + d];
      const double old_mean = (double)g.means_invvars[(size_t)c * FB_DIM + d] / iv;
      const double mean = (macc[(size_t)c * FB_DIM + d] + mean_tau * old_mean) * (1.0 / (occ[c] + mean_tau));
      const float miv = (float)(mean * iv);
      out_means_invvars[(size_t)c * FB_DIM + d] = miv;
      gc += 0.5 * log(iv) - 0.5 * (double)miv * (double)miv / iv;
    }
    out_gconsts[c] = (float)gc;
    if (out_occupancy) out_occupancy[c] = occ[c];
  }
  return done(FB_OK);
}
