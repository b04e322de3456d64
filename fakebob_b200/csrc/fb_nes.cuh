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
  double *noise;       // [pairs_local][N] float64: host-supplied numpy noise (rng = HOST)
  float *noise32;      // [pairs_local][N] float32: device Philox noise (the Box-Muller deviates are float32 values), else NULL
  int *flags;          // [0] stopped, [1] iterations done, [2] stop iteration, [3] plateau window fill
  unsigned long long *state_u64;   // [0] Philox draw counter
  unsigned long long *dist_bits;   // [iter] L-inf distance before iteration `iter` (as double bits)
  int64_t N;
  int S, K, pairs_total, pairs_local, pair0, B_local, has_clean;
  int task, targeted, label, plateau_length, auto_stop, znorm;
  double kappa, sigma, epsilon, momentum, one_minus_momentum, min_lr, plateau_drop;
  unsigned long long seed;
  // estimate_threshold mode (FAKEBOB.py:82-113): the clean column of every batch is the make_decisions() score of the current
  // adversarial audio, so the stop tests run on the device: accepted by the system (score >= accept_threshold, flags[0] = 2)
  // or candidate threshold reached (score >= theta, flags[0] = 1); neither consumes the iteration's noise draw.
  int est_mode;
  double accept_threshold;
  // multi-GPU one-shot exchange over peer memory (fb_comm.cu): every rank publishes [grad partial (N) | losses | clean
  // scores] in its OWN buffer (two parities) and raises a flag; every rank then pulls the W partials over NVLink and adds
  // them in rank order inside the update kernel -- no ring, no separate collective kernel.  NULL: ncclAllReduce instead.
  double *xown;                    // [2][xstride]
  unsigned long long *xflag;       // [2] own flags, value = (session << 32) | (iteration + 1)
  const unsigned long long *xsess; // device word holding the session number (so the captured graph survives new sessions)
  const double *xpeer[8];          // every rank's buffer (own included), mapped
  const unsigned long long *xpeer_flag[8];
  unsigned long long xstride;
  int world, rank;
};

// One per context, created by the first fb_nes_init and reused by every later session: the device buffers only grow
// (no cudaMalloc / cudaFree per attack), and the captured iteration graph is kept while the kernel arguments it baked in
// (buffer addresses, sizes, hyper-parameters) are unchanged.
struct FbNes {
  fb_nes_params p;
  FbNesDev dev;
  int64_t N = 0;
  int pairs_local = 0, pair0 = 0, B_local = 0, rank = 0, world = 1;
  bool has_clean = true;
  DevBuf<double> f64_pool;
  DevBuf<double> noise64;
  DevBuf<float> noise32;
  DevBuf<unsigned long long> flags;       // 8 x u64 holding 16 ints, then 16 u64 counters
  DevBuf<unsigned long long> dist_bits;
  DevBuf<double> ext_scores;              // [B][K] scores handed in by a black-box scorer (fb_nes_ext_update)
  size_t red_count = 0;
  int enqueued = 0;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  uint64_t graph_epoch = 0;
  FbNesDev graph_dev;                     // the arguments the captured graph was built with
  int graph_arch = -1;
  int64_t launches_per_iter = 0;
  std::vector<int64_t> offsets;
};

void fb_nes_destroy(fb_ctx *ctx);
void fb_comm_info(fb_ctx *ctx, int *rank, int *world);
struct FbNesDev;
// fills the peer-exchange fields of `d` when the context has mapped peer buffers large enough for `count` doubles; bumps the session
int fb_comm_p2p_attach(fb_ctx *ctx, FbNesDev *d, size_t count);
