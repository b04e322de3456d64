// NES device state (see fb_nes.cu).
#pragma once
#include "fb_common.cuh"

struct FbNesDev {
  // float64 vectors of length N (FAKEBOB.py:156-164: adver, grad, lower, upper)
  double *audio, *adver, *lower, *upper, *grad, *gest;
  double *red;         // [N grad sum | S+1 losses by global column | K clean scores]   (all-reduce payload)
  double *zmean, *zstd;
  double *state_f64;   // [0] lr, [8..8+L) plateau window
  double *log;         // [max_iter+1][4+K]
  double *threshold;   // [0] theta (device-resident so a captured graph sees updates)
  double *noise;       // [pairs_local][N]
  int *flags;          // [0] stopped, [1] iterations done, [2] stop iteration, [3] plateau window fill
  unsigned long long *state_u64;   // [0] Philox draw counter
  unsigned long long *dist_bits;   // [iter] L-inf distance before iteration `iter` (as double bits)
  int64_t N;
  int S, K, pairs_total, pairs_local, pair0, B_local, has_clean;
  int task, targeted, label, plateau_length, auto_stop, znorm;
  double kappa, sigma, epsilon, momentum, one_minus_momentum, min_lr, plateau_drop;
  unsigned long long seed;
};

struct FbNes {
  fb_nes_params p;
  FbNesDev dev;
  int64_t N = 0;
  int pairs_local = 0, pair0 = 0, B_local = 0, rank = 0, world = 1;
  bool has_clean = true;
  double *f64_pool = nullptr;
  double *noise = nullptr;
  int *flags = nullptr;
  unsigned long long *dist_bits = nullptr;
  size_t red_count = 0;
  int enqueued = 0;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  uint64_t graph_epoch = 0;
  int64_t launches_per_iter = 0;
  std::vector<int64_t> offsets;
};

void fb_nes_destroy(fb_ctx *ctx);
void fb_comm_info(fb_ctx *ctx, int *rank, int *world);
