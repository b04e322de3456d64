// i-vector / PLDA device state (see fb_ivector.cu).
#pragma once
#include "fb_common.cuh"

struct FbIvector {
  int C = 0, R = 0, L = 0, K = 0, n_packed = 0, lda_cols = 0, n_splits = 0, n_splits_used = 0;
  double prior_offset = 0.0;
  float min_post = 0.025f;
  bool have_ubm = false, have_ie = false, have_backend = false;
  // resident parameters
  DevBuf<float> gconsts, means_invcovars, inv_covars;      // full UBM (inv_covars packed lower-triangular)
  DevBuf<unsigned short> rc_table;
  DevBuf<float> sim32;                                      // [C][72][R]   Sigma^-1 M
  DevBuf<float> U;                                          // [C][R(R+1)/2] vech(M' Sigma^-1 M)
  DevBuf<float> mean_vec, lda;
  DevBuf<double> plda_T, plda_off, psi, u_train;
  std::vector<float> h_mean_vec, h_lda;
  std::vector<double> h_plda_T, h_plda_off, h_psi;
  // per-batch workspace
  DevBuf<float> ll;                                         // [rows_cap][C]
  DevBuf<int> gsel;                                         // [rows_cap][20]
  DevBuf<float> post;                                       // [rows_cap][20]
  DevBuf<double> gamma, Xs, lin_part, quad, Awork, scores;
  DevBuf<float> ivec;
  DevBuf<int> act_list;                                     // [ceil(B/32)][C + 1]: [0] = count, then the active components, ascending
};

void fb_ivector_destroy(fb_ctx *ctx);
int fb_run_ivector_flag(fb_ctx *ctx, const int *done_flag, bool with_plda);
int fb_ivector_reserve(fb_ctx *ctx);      // allocate the per-batch workspace (never inside a stream capture)
