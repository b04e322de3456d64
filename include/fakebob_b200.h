/*
 * fakebob_b200 -- C ABI of the B200-native FAKEBOB hot path (libfakebob_b200.so).
 *
 * The reference has no FFI: its extension point is Python duck typing
 * (README.md:136 "function score and make_decisions") and every scorer shells out to
 * Kaldi binaries through gmm_ubm_kaldiHelper / ivector_PLDA_kaldiHelper.  This header
 * is the boundary a binding for that path needs: each entry point names the reference
 * call it replaces.  Plain C, caller-owned buffers, explicit sizes, int return codes
 * (0 = ok, negative = error, text via fb_last_error()).  No torch / C++ types cross it.
 * A context is bound to one CUDA device and one stream; a context is not thread-safe,
 * different contexts are independent.
 *
 * Conventions
 *   - "host" pointers are CPU memory, "dev" pointers are device memory on the context's GPU.
 *   - audio is 16-bit PCM exactly as the reference writes it to <i>.wav
 *     (gmm_ubm_kaldiHelper.py:56-68) after (x * 2^15).astype(int16) (gmm_ubm_OSI.py:83-85).
 *   - utterance b of a batch occupies samples [offsets[b], offsets[b+1]) of `wave`.
 */
#ifndef FAKEBOB_B200_H
#define FAKEBOB_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fb_ctx fb_ctx;

/* ---- errors / lifetime ------------------------------------------------------------ */
#define FB_OK               0
#define FB_ERR_CUDA        -1
#define FB_ERR_ARG         -2
#define FB_ERR_STATE       -3
#define FB_ERR_NO_VOICED   -4   /* an utterance has zero voiced frames (Kaldi would drop it) */
#define FB_ERR_NCCL        -5
#define FB_ERR_UNSUPPORTED -6
#define FB_ERR_TOO_LONG    -7   /* an utterance exceeds a kernel's per-utterance capacity */
#define FB_ERR_NOT_SPD     -8   /* i-vector posterior precision (quad + I) is not positive definite */

const char *fb_last_error(void);
int  fb_version(void);
int  fb_ctx_create(int device, fb_ctx **out);
int  fb_ctx_destroy(fb_ctx *ctx);
/* cudaStream_t passed as void*; NULL = the context's own stream. */
int  fb_set_stream(fb_ctx *ctx, void *cuda_stream);
int  fb_synchronize(fb_ctx *ctx);

/* ---- feature configuration --------------------------------------------------------
 * Replaces: --config=conf/mfcc.conf (gmm_ubm_kaldiHelper.py:138), --vad-config conf/vad.conf (:158),
 * delta_opts (:191-193), apply-cmvn-sliding options (:196). */
typedef struct fb_feat_config {
  float sample_frequency;      /* 16000 */
  float low_freq, high_freq;   /* 20, 7600 */
  int   num_mel_bins;          /* 30 (<= 32) */
  int   num_ceps;              /* 24 (fixed by the kernels) */
  float preemph;               /* 0.97 */
  float cepstral_lifter;       /* 22 */
  float vad_energy_threshold;  /* 5.5 */
  float vad_energy_mean_scale; /* 0.5 */
  float vad_proportion_threshold; /* 0.12 */
  int   vad_frames_context;    /* 2 */
  int   cmn_window;            /* 300 */
} fb_feat_config;
int fb_set_feature_config(fb_ctx *ctx, const fb_feat_config *cfg);
/* Non-ideal effects of the reference's real Kaldi path, off by default:
 *   compress_features != 0: the MFCC matrix of every utterance goes through Kaldi's CompressedMatrix speech-feature codec
 *     (steps/make_mfcc.sh writes with copy-feats --compress=true, gmm_ubm_kaldiHelper.py:138-147) before VAD / deltas / CMN;
 *   text_precision != 0: the values the reference parses from Kaldi's text output are rounded to 7 significant digits:
 *     average log-likelihoods (ark,t:, gmm_ubm_kaldiHelper.py:204-208), i-vectors and PLDA scores
 *     (ivector_PLDA_kaldiHelper.py:202-211,262-271).
 * Kaldi's default dither (libc rand() noise of +-1 LSB per sample) is not reproduced on the device. */
int fb_set_kaldi_exact(fb_ctx *ctx, int compress_features, int text_precision);

/* ---- diagonal GMMs ---------------------------------------------------------------
 * Replaces: the model rxfilename argument of gmm-global-get-frame-likes
 * (gmm_ubm_kaldiHelper.py:206-208); arrays are Kaldi's stored DiagGmm members
 * (host, float32, row-major C x D).  Slots are scored in slot order.  All slots must share C and D=72. */
int fb_load_diag_gmm(fb_ctx *ctx, int slot, const float *weights, const float *means_invvars,
                     const float *inv_vars, const float *gconsts, int C, int D);
int fb_finalize_gmms(fb_ctx *ctx, int n_models);
/* Number of fp16 product terms used for the speaker-minus-slot-0 part of the contraction when all slots share their
 * inverse variances (MAP mean-only adaptation, build_spk_models.py:170):  ll_m = ll_0 + x.(w_m - w_0) + (g_m - g_0).
 * Slot 0 and the x^2 term always use the full three-term split (hi.hi + lo.hi + hi.lo, ~22 bits per operand).
 * 3 = the same for the difference (bit-for-bit the fp32-class result), 2 = x_hi.(dw_hi + dw_lo), 1 = x_hi.dw_hi.
 * 0 (default) = choose from the size of the differences so that the expected score deviation stays below
 * the 1e-4 the reference's own 7-significant-digit text scores resolve (gmm_ubm_kaldiHelper.py:204-208).
 * Call before fb_finalize_gmms. */
int fb_set_gmm_delta_terms(fb_ctx *ctx, int terms);
/* After fb_finalize_gmms: *shared_variances = 1 when the difference formulation is in use, *delta_terms = the number of
 * product terms in effect, *err_estimate = the predicted per-frame error of a one-term difference product (any may be NULL). */
int fb_get_gmm_info(fb_ctx *ctx, int *shared_variances, int *delta_terms, double *err_estimate);

/* ---- batch scoring: serves gmm_{CSI,OSI,SV}.score() --------------------------------
 * Replaces: gmm_ubm_kaldiHelper.score() (gmm_ubm_kaldiHelper.py:270-291): write_audio, data_prepare,
 * make_mfcc, compute_vad, get_frames_likes, resolce_scores.
 * out_avg_ll[b * n_models + k] = average per-voiced-frame log-likelihood of utterance b under slot k
 * (what `gmm-global-get-frame-likes --average=true` prints), float64.
 * _host: pageable or pinned host buffers, copies included, synchronous.
 * _dev : device buffers, asynchronous on the context's stream. */
int fb_score_gmm_host(fb_ctx *ctx, const int16_t *wave, const int64_t *offsets, int B, double *out_avg_ll);
int fb_score_gmm_dev(fb_ctx *ctx, const int16_t *wave_dev, const int64_t *offsets_host, int B, double *out_avg_ll_dev);

/* ---- enrolment: serves build_spk_models.py step 2 ------------------------------------------------
 * Replaces: `gmm-global-acc-stats --update-flags=m final.dubm <feats> acc` + `gmm-global-est-map --update-flags=m final.dubm
 * acc <spk>-identity.gmm` (build_spk_models.py:202-219; gmm-global-est-map.cc is the reference's one native source file).
 * MAP mean-only adaptation (MapDiagGmmUpdate, mean_tau = 10 by default in Kaldi) of slot 0 -- load the UBM alone,
 * fb_finalize_gmms(ctx, 1) -- to the voiced frames of the batch, all utterances pooled like one feature archive.
 * Weights and variances are kept.  out_means_invvars: C x 72 float32, out_gconsts: C float32 (DiagGmm members, ready for
 * fb_load_diag_gmm or a Kaldi model file), out_occupancy: C float64 or NULL. */
int fb_map_adapt_host(fb_ctx *ctx, const int16_t *wave, const int64_t *offsets, int B, double mean_tau,
                      float *out_means_invvars, float *out_gconsts, double *out_occupancy);

/* ---- i-vector / PLDA scoring: serves iv_{CSI,OSI,SV}.score() ---------------------------------
 * Replaces: ivector_PLDA_kaldiHelper.score() (ivector_PLDA_kaldiHelper.py:340-363): write_audio, data_prepare, make_mfcc,
 * compute_vad, extract_ivector (sid/extract_ivectors.sh, :202-211), write_trials, plda_scoring (:251-280), resolve_score.
 * Parameters are Kaldi's stored members: final.ubm (FullGmm: inv_covars as full symmetric C x D x D float32),
 * final.ie (M: C x D x R, SigmaInv: C x D x D, float64; no weight projection), mean.vec, transform.mat (L x R or
 * L x (R+1)), plda (mean, transform, psi; float64).  Derived matrices are computed once, on the device. */
int fb_load_full_gmm(fb_ctx *ctx, const float *weights, const float *means_invcovars, const float *inv_covars,
                     const float *gconsts, int C, int D);
int fb_load_ivector_extractor(fb_ctx *ctx, const double *M, const double *sigma_inv, double prior_offset, int C, int D, int R);
int fb_load_plda_backend(fb_ctx *ctx, const float *mean_vec, const float *transform_mat, int transform_cols,
                         const double *plda_mean, const double *plda_transform, const double *plda_psi, int R, int L);
/* Raw enrolled i-vectors (K x R float32, as read from the pickle's identity_location); back-end applied on load. */
int fb_set_enrolled_ivectors(fb_ctx *ctx, const float *enrolled, int K);
/* out_scores[b * K + k] = PLDA log-likelihood ratio of test utterance b against enrolled speaker k (float64; what
 * ivector-plda-scoring prints), or NULL; out_ivectors[b * R + r] = raw i-vectors (what ivector-extract writes), or NULL. */
int fb_score_ivector_host(fb_ctx *ctx, const int16_t *wave, const int64_t *offsets, int B, double *out_scores,
                          float *out_ivectors);
/* Intermediate results of utterance b of the last i-vector batch (parity tests): gamma [C], X [C x 72] (Baum-Welch
 * statistics), lin [R] (before the prior offset), quad [R(R+1)/2] (packed lower triangle, row-major, before + I);
 * float64, any may be NULL. */
int fb_get_ivector_stats(fb_ctx *ctx, int b, double *gamma_host, double *x_host, double *lin_host, double *quad_host);
/* Gaussian selection and pruned posteriors of the last i-vector batch: [rows][20] each; returns rows. */
int fb_get_posteriors(fb_ctx *ctx, int32_t *gsel_host, float *post_host, int64_t capacity_rows);

/* ---- stage read-backs (used by the parity tests; valid after a score call) ------------ */
int fb_set_debug(fb_ctx *ctx, int keep_f32_features);
int fb_get_num_frames(fb_ctx *ctx, int B, int32_t *frames_host, int32_t *voiced_host);
int fb_get_mfcc(fb_ctx *ctx, float *out_host, int64_t capacity_floats);          /* [sum T_b][24] */
int fb_get_vad(fb_ctx *ctx, int32_t *out_host, int64_t capacity);                /* [sum T_b] rank or -1 */
int fb_get_features(fb_ctx *ctx, float *out_host, int64_t capacity_floats);      /* [sum Tv_b][72], needs debug */
int fb_get_frame_loglikes(fb_ctx *ctx, float *out_host, int64_t capacity_floats);/* [n_models][sum Tv_b] */

/* ---- NES attack state: serves FakeBob.attack()/get_grad()/estimate_threshold() --------
 * Replaces: the body of the loop at FAKEBOB.py:168-214 (get_grad :223-246, loss_fn :248-299,
 * momentum :193, plateau LR :195-200, sign step + clip :202-203, early stop :181, log rows :209-214). */
enum { FB_TASK_CSI = 0, FB_TASK_OSI = 1, FB_TASK_SV = 2 };
enum { FB_RNG_HOST = 0, FB_RNG_PHILOX = 1 };

typedef struct fb_nes_params {
  int    task;               /* FB_TASK_* */
  int    targeted;           /* 1 = targeted */
  int    label;              /* target (targeted) or true (CSI untargeted) speaker index; -1 if unused */
  int    n_speakers;         /* K; model slots are [ubm, spk_0..spk_{K-1}] for OSI/SV, [spk_0..] for CSI */
  int    samples_per_draw;   /* S; S/2 antithetic pairs (odd S uses S-1, FAKEBOB.py:234-235) */
  int    max_iter;
  int    rng;                /* FB_RNG_* */
  int    plateau_length;
  double threshold;          /* theta */
  double adver_thresh;       /* kappa */
  double epsilon, sigma, max_lr, min_lr, momentum, plateau_drop;
  uint64_t seed;             /* Philox key */
  uint64_t draw_base;        /* Philox iteration counter of the first draw */
  const double *z_norm_means;/* CSI only, host, K */
  const double *z_norm_stds; /* CSI only, host, K */
  int    external_scorer;    /* 1 = black-box model (README.md:136: any object with score()/make_decisions()): the context needs
                                no resident models; iterate with fb_nes_ext_perturb / fb_nes_ext_update */
} fb_nes_params;

/* Creates device state for one utterance of n_samples (float64 audio in [-1,1], host). */
int fb_nes_init(fb_ctx *ctx, const fb_nes_params *p, const double *audio_host, int64_t n_samples);
/* Change theta / label / attack type without re-uploading the audio (estimate_threshold outer loop). */
int fb_nes_set_threshold(fb_ctx *ctx, double threshold);
/* Enqueue up to n_iters iterations (stops early on device when adver_loss < 0).  rng = HOST: noise_host
 * is n_iters x (S/2) x n_samples float64 (pair-major), else NULL.  Asynchronous. */
int fb_nes_run(fb_ctx *ctx, int n_iters, const double *noise_host);
/* Blocks until enqueued work finished; returns iterations executed so far and the stop code (0 = running, 1 = early stop /
 * candidate threshold reached, 2 = accepted by the system in estimate mode). */
int fb_nes_status(fb_ctx *ctx, int *iters_done, int *stopped);
/* Log rows [iter] = {distance, adver_loss, final_loss, lr, score_0..score_{K-1}} (float64, 4+K per row). */
int fb_nes_read_log(fb_ctx *ctx, double *rows_host, int max_rows);
int fb_nes_read_adver(fb_ctx *ctx, double *adver_host, int64_t n_samples);
int fb_nes_read_grad(fb_ctx *ctx, double *grad_host, int64_t n_samples);
/* The last gradient estimate (before momentum) and the losses [S+1] / clean scores [K] it was computed from. */
int fb_nes_read_gest(fb_ctx *ctx, double *gest_host, int64_t n_samples, double *losses_host, double *score0_host);
/* FakeBob.estimate_threshold on the device (FAKEBOB.py:76-137).  The reference scores the current adversarial audio with
 * make_decisions() and then, separately, the S+1 batch of get_grad(); column 0 of that batch IS the current audio, so one
 * batch per inner iteration serves both.  After fb_nes_init (untargeted OSI / SV, max_iter = an upper bound on the total
 * number of inner iterations): fb_nes_estimate_begin(accept_threshold = the system's own threshold, model.threshold),
 * then per outer iteration fb_nes_continue(theta) -- sets theta, lr = max_lr, empties the plateau window, clears the stop
 * flag -- and fb_nes_run / fb_nes_status until `stopped` is 1 (score >= theta: next outer iteration) or 2 (accepted: done).
 * The stopping iteration applies no update and does not consume its noise draw, like the reference's break / return. */
int fb_nes_estimate_begin(fb_ctx *ctx, double accept_threshold);
int fb_nes_continue(fb_ctx *ctx, double threshold);
/* Black-box scorers (fb_nes_params.external_scorer = 1): the device keeps the attack state, draws the noise, quantises and
 * applies the update; the S+1 audios of an iteration go to the host, the caller scores them with its own system and hands
 * the scores back.  fb_nes_ext_perturb writes the batch as int16 PCM [S+1][n_samples] (row 0 = the current adversarial audio,
 * rows 1..S/2 = +noise, the rest = -noise: the column order of FAKEBOB.py:234-237).  fb_nes_ext_update takes the scores
 * [S+1][K] (K = n_speakers; SV: K = 1) exactly as score() returned them, computes the losses, the gradient estimate and the
 * update (FAKEBOB.py:239-246,181-203); gradient_only = 1 stops after the estimate (FakeBob.get_grad; read it with
 * fb_nes_read_gest, apply it with fb_nes_apply_update).  Stop state / log as with fb_nes_status / fb_nes_read_log. */
int fb_nes_ext_perturb(fb_ctx *ctx, int16_t *wave_host);
int fb_nes_ext_update(fb_ctx *ctx, const double *scores_host, int gradient_only);
/* One gradient estimate without the update: FakeBob.get_grad (FAKEBOB.py:223-246). */
int fb_nes_get_grad(fb_ctx *ctx, const double *noise_host, double *final_loss, double *adver_loss,
                    double *score0_host, double *grad_host);
/* adver <- clip(adver - lr*sign(momentum*grad + (1-momentum)*g), lower, upper) with g from the last get_grad. */
int fb_nes_apply_update(fb_ctx *ctx, double lr);
int fb_nes_kernel_launches(fb_ctx *ctx, int64_t *count);

/* ---- per-stage device timing (CUDA events on the context's stream; disables graph replay while on) ----
 * Stage ids: 0 perturb, 1 mfcc, 2 vad_scan, 3 feats, 4 gmm, 5 gmm_reduce, 6 loss, 7 update(+collective),
 * 8 gselect, 9 fgmm_post, 10 ivec_stats, 11 ivec_lin, 12 ivec_quad, 13 ivec_solve, 14 plda.
 * fb_profile_read returns accumulated milliseconds and launch counts per stage since fb_profile_enable(ctx, 1). */
#define FB_PROF_STAGES 16
int fb_profile_enable(fb_ctx *ctx, int on);
int fb_profile_read(fb_ctx *ctx, double *ms_host, int64_t *count_host);
/* Total voiced rows (frames) of the last scored batch: the GMM kernel's M dimension. */
int fb_get_voiced_rows(fb_ctx *ctx, int *rows);

/* ---- multi-GPU: one process per GPU, antithetic pairs sharded across ranks ---------------
 * New (nothing in the reference communicates): a single ncclAllReduce(sum, float64) of
 * [grad partial (N) | losses (S+1) | score_0 (K)] per iteration on the context's stream.
 * The unique id is produced by rank 0 and distributed by the host (torch.distributed). */
int fb_comm_unique_id(void *out_128_bytes);
int fb_comm_init(fb_ctx *ctx, const void *unique_id_128_bytes, int rank, int world);
int fb_comm_destroy(fb_ctx *ctx);
/* Optional, after fb_comm_init (<= 8 ranks of one node): exchange buffers in peer memory.  Every rank calls
 * fb_comm_p2p_export (allocates its buffer for utterances of up to max_samples, returns a 64-byte CUDA IPC handle), the host
 * gathers the handles of all ranks in rank order, every rank calls fb_comm_p2p_import.  From then on an iteration's
 * gradient partials are published in each rank's own buffer and pulled over NVLink inside the update kernel (summed in rank
 * order, identical on every rank) instead of going through ncclAllReduce.  FB_ERR_UNSUPPORTED (mapping refused): the
 * ncclAllReduce path stays in use. */
int fb_comm_p2p_export(fb_ctx *ctx, int64_t max_samples, void *handle_out_64_bytes);
int fb_comm_p2p_import(fb_ctx *ctx, const void *handles_world_x_64_bytes);

#ifdef __cplusplus
}
#endif
#endif /* FAKEBOB_B200_H */
