// NES attack loop on the device.
//
// Replaces the body of FakeBob.attack()'s loop (FAKEBOB.py:168-214) and FakeBob.get_grad()
// (FAKEBOB.py:223-246) / loss_fn() (:248-299):
//   perturb_kernel    noise draw (Philox4x32-10 + Box-Muller, float64) or host-supplied numpy noise,
//                     noise_audios = sigma * noise + adver (:237), int16 truncation (gmm_ubm_OSI.py:83-85)
//   [front-end + GMM kernels score the S+1 audios]
//   nes_loss_kernel   scores from average log-likelihoods (gmm_ubm_OSI.py:89 / gmm_ubm_CSI.py:93 /
//                     gmm_ubm_SV.py:77), margin loss (:248-299), early-stop test (:181), plateau LR
//                     schedule (:195-200), log row (:209-214)
//   nes_update_kernel grad = mean_i(loss_i * noise_i) / sigma in numpy's pairwise summation order (:244),
//                     momentum (:193), sign step + clip (:202-203), next iteration's L-inf distance (:178)
// All state is float64 like the reference's numpy arrays, and every operation is rounded separately
// (no FMA contraction), so given identical losses and noise the update is bit-identical to numpy.
#include "fb_common.cuh"
#include "fb_nes.cuh"
#include "fb_ivector.cuh"
#include <math.h>
#include <string.h>

int fb_run_frontend_flag(fb_ctx *ctx, const int *done_flag);
int fb_run_gmm_flag(fb_ctx *ctx, const int *done_flag);
int fb_comm_allreduce_f64(fb_ctx *ctx, double *buf, size_t count);

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (counter = (n/4, pair, draw_lo, draw_hi), key = seed) -> 4 float64 normals
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// Box-Muller in float32 on 23-bit uniforms (every integer -> float step is exact; logf / sincosf are the accurate
// <= 2 ulp library versions).  The normal deviates are float32 values widened to float64 for the state arithmetic.
__device__ __forceinline__ void box_muller(uint32_t xa, uint32_t xb, double &z0, double &z1) {
  const float ua = __fmul_rn(__fadd_rn((float)(xa >> 9), 0.5f), 1.1920928955078125e-07f);    // (k + 0.5) * 2^-23
  const float ub = __fmul_rn(__fadd_rn((float)(xb >> 9), 0.5f), 1.1920928955078125e-07f);
  const float rad = __fsqrt_rn(__fmul_rn(-2.0f, logf(ua)));
  const float ang = __fmul_rn(6.2831855f, ub);
  float sn, cs;
  sincosf(ang, &sn, &cs);
  z0 = (double)__fmul_rn(rad, cs);
  z1 = (double)__fmul_rn(rad, sn);
}

__device__ __forceinline__ int16_t quantise(double v) {
  // (v * 2^15).astype(np.int16): truncation toward zero, two's-complement wrap outside int16
  const int q = __double2int_rz(__dmul_rn(v, 32768.0));
  return (int16_t)(q & 0xFFFF);
}

// grid (ceil(N/4/256), pair chunks), block 256; each thread owns 4 consecutive samples.
__global__ void __launch_bounds__(256)
perturb_kernel(FbNesDev st, int16_t *__restrict__ wave, int64_t stride, int philox) {
  FB_GRID_DEP_SYNC();
  if (st.flags[0]) return;
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n0 = g * 4;
  if (n0 >= st.N) return;
  const int nv = (int)min((int64_t)4, st.N - n0);
  const bool vec = nv == 4 && (st.N & 3) == 0 && (stride & 3) == 0;     // rows stay 8-byte (int16) / 32-byte (f64) aligned
  double a[4];
  for (int i = 0; i < nv; ++i) a[i] = st.adver[n0 + i];
  const int bclean = st.has_clean ? 1 : 0;
  if (bclean && blockIdx.y == 0) {
    if (vec) *reinterpret_cast<short4 *>(wave + n0) = make_short4(quantise(a[0]), quantise(a[1]), quantise(a[2]), quantise(a[3]));
    else for (int i = 0; i < nv; ++i) wave[n0 + i] = quantise(a[i]);
  }
  const unsigned long long draw = st.state_u64[0];
  for (int j = blockIdx.y; j < st.pairs_local; j += gridDim.y) {
    double z[4];
    if (philox) {
      uint32_t r[4];
      philox4x32_10((uint32_t)g, (uint32_t)(st.pair0 + j), (uint32_t)draw, (uint32_t)(draw >> 32),
                    (uint32_t)st.seed, (uint32_t)(st.seed >> 32), r);
      box_muller(r[0], r[1], z[0], z[1]);
      box_muller(r[2], r[3], z[2], z[3]);
      float *np_ = st.noise32 + (int64_t)j * st.N + n0;      // the deviates are float32 values: stored exactly
      if (vec) {
        *reinterpret_cast<float4 *>(np_) = make_float4((float)z[0], (float)z[1], (float)z[2], (float)z[3]);
      } else {
        for (int i = 0; i < nv; ++i) np_[i] = (float)z[i];
      }
    } else {
      for (int i = 0; i < nv; ++i) z[i] = st.noise[(int64_t)j * st.N + n0 + i];
    }
    int16_t *wp = wave + (int64_t)(bclean + j) * stride + n0;
    int16_t *wm = wave + (int64_t)(bclean + st.pairs_local + j) * stride + n0;
    int16_t qp[4], qm[4];
    for (int i = 0; i < nv; ++i) {
      const double t = __dmul_rn(st.sigma, z[i]);
      qp[i] = quantise(__dadd_rn(t, a[i]));
      qm[i] = quantise(__dadd_rn(-t, a[i]));
    }
    if (vec) {                                        // 8-byte stores: a warp writes 256 contiguous bytes
      *reinterpret_cast<short4 *>(wp) = make_short4(qp[0], qp[1], qp[2], qp[3]);
      *reinterpret_cast<short4 *>(wm) = make_short4(qm[0], qm[1], qm[2], qm[3]);
    } else {
      for (int i = 0; i < nv; ++i) { wp[i] = qp[i]; wm[i] = qm[i]; }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// numpy pairwise summation (pairwise_sum_DOUBLE): blocks of <=128 with 8 accumulators, recursive halves above.
// ------------------------------------------------------------------------------------------------
template <typename F>
__device__ __forceinline__ double np_pairwise_block(const F &get, int lo, int n) {     // n <= 128
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = __dadd_rn(r, get(lo + i));
    return r;
  }
  double r0 = get(lo), r1 = get(lo + 1), r2 = get(lo + 2), r3 = get(lo + 3);
  double r4 = get(lo + 4), r5 = get(lo + 5), r6 = get(lo + 6), r7 = get(lo + 7);
  int i = 8;
  for (; i < n - (n % 8); i += 8) {
    r0 = __dadd_rn(r0, get(lo + i));     r1 = __dadd_rn(r1, get(lo + i + 1));
    r2 = __dadd_rn(r2, get(lo + i + 2)); r3 = __dadd_rn(r3, get(lo + i + 3));
    r4 = __dadd_rn(r4, get(lo + i + 4)); r5 = __dadd_rn(r5, get(lo + i + 5));
    r6 = __dadd_rn(r6, get(lo + i + 6)); r7 = __dadd_rn(r7, get(lo + i + 7));
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r0, r1), __dadd_rn(r2, r3)), __dadd_rn(__dadd_rn(r4, r5), __dadd_rn(r6, r7)));
  for (; i < n; ++i) res = __dadd_rn(res, get(lo + i));
  return res;
}

template <typename F>
__device__ __noinline__ double np_pairwise_rec(const F &get, int lo, int n) {          // n > 128: numpy's recursive halves
  if (n <= 128) return np_pairwise_block(get, lo, n);
  int n2 = n / 2;
  n2 -= n2 % 8;
  const double a = np_pairwise_rec(get, lo, n2);
  const double b = np_pairwise_rec(get, lo + n2, n - n2);
  return __dadd_rn(a, b);
}

template <typename F>
__device__ __forceinline__ double np_pairwise(const F &get, int lo, int n) {
  return (n <= 128) ? np_pairwise_block(get, lo, n) : np_pairwise_rec(get, lo, n);
}

// ------------------------------------------------------------------------------------------------
// Scores + margin loss for the local batch; optional bookkeeping (single-GPU: fused; multi-GPU: after allreduce).
// One CTA.  red layout: [0,N) gradient sum | [N, N+S+1) losses by global column | [N+S+1, N+S+1+K) clean scores.
// ------------------------------------------------------------------------------------------------
__device__ double sample_loss(const FbNesDev &st, const double *ll) {
  const int K = st.K;
  double sc_own = 0.0, sc_other = -INFINITY, sc_max = -INFINITY;
  for (int k = 0; k < K; ++k) {
    double s;
    if (st.znorm) s = __ddiv_rn(__dadd_rn(ll[k], -st.zmean[k]), st.zstd[k]);
    else s = __dadd_rn(ll[1 + k], -ll[0]);
    if (k == st.label) sc_own = s; else sc_other = fmax(sc_other, s);
    sc_max = fmax(sc_max, s);
  }
  if (st.task == FB_TASK_SV) return __dadd_rn(__dadd_rn(st.threshold[0], st.kappa), -sc_max);
  if (st.task == FB_TASK_OSI) {
    if (!st.targeted) return __dadd_rn(__dadd_rn(st.threshold[0], st.kappa), -sc_max);
    return __dadd_rn(__dadd_rn(fmax(sc_other, st.threshold[0]), st.kappa), -sc_own);
  }
  if (st.targeted) return __dadd_rn(__dadd_rn(sc_other, st.kappa), -sc_own);
  return __dadd_rn(__dadd_rn(sc_own, st.kappa), -sc_other);
}

struct PtrGetter {
  const double *p;
  __device__ double operator()(int i) const { return p[i]; }
};

__device__ void bookkeeping(const FbNesDev &st) {
  // executed by one thread
  const double *loss = st.red + st.N;
  const int S = st.S;
  const int it = st.flags[1];
  const double adver_loss = loss[0];
  PtrGetter gl{loss + 1};
  const double final_loss = __ddiv_rn(np_pairwise(gl, 0, S), (double)S);
  double *row = st.log + (size_t)it * (4 + st.K);
  row[0] = __longlong_as_double((long long)st.dist_bits[it]);
  row[1] = adver_loss;
  row[2] = final_loss;
  for (int k = 0; k < st.K; ++k) row[4 + k] = st.red[st.N + S + 1 + k];
  double lr = st.state_f64[0];
  if (st.est_mode) {
    double sc = -INFINITY;
    for (int k = 0; k < st.K; ++k) sc = fmax(sc, st.red[st.N + S + 1 + k]);     // np.max(score) for OSI, the score for SV
    const int code = (sc >= st.accept_threshold) ? 2 : ((sc >= st.threshold[0]) ? 1 : 0);
    if (code) {
      row[3] = lr;
      st.flags[0] = code;          // FAKEBOB.py:97-106: return / break before get_grad -- no update, the draw is not consumed
      st.flags[2] = it;
      st.flags[1] = it + 1;
      return;
    }
  } else if (st.auto_stop && adver_loss < 0.0) {
    row[3] = lr;
    st.flags[0] = 1;             // early stop: no update (FAKEBOB.py:181-191)
    st.flags[2] = it;
    st.flags[1] = it + 1;
    st.state_u64[0] += 1;
    return;
  }
  // plateau schedule (FAKEBOB.py:195-200)
  int n_ls = st.flags[3];
  double *ls = st.state_f64 + 8;
  const int L = st.plateau_length;
  if (n_ls == L) {
    for (int i = 0; i + 1 < L; ++i) ls[i] = ls[i + 1];
    ls[L - 1] = final_loss;
  } else {
    ls[n_ls++] = final_loss;
  }
  if (n_ls == L && ls[L - 1] > ls[0]) {
    if (lr > st.min_lr) lr = fmax(__ddiv_rn(lr, st.plateau_drop), st.min_lr);
    n_ls = 0;
  }
  st.flags[3] = n_ls;
  st.state_f64[0] = lr;
  row[3] = lr;
  st.flags[1] = it + 1;
  st.state_u64[0] += 1;          // next Philox draw
}

__global__ void __launch_bounds__(256)
nes_loss_kernel(FbNesDev st, const double *__restrict__ avg_ll, int n_models, int do_book) {
  FB_GRID_DEP_SYNC();
  if (st.flags[0]) return;
  const int bl = st.B_local;
  const int bclean = st.has_clean ? 1 : 0;
  double *loss = st.red + st.N;
  for (int b = threadIdx.x; b < bl; b += blockDim.x) {
    const double *ll = avg_ll + (size_t)b * n_models;
    int col;
    if (bclean && b == 0) col = 0;
    else {
      const int j = b - bclean;
      col = (j < st.pairs_local) ? (1 + st.pair0 + j) : (1 + st.pairs_total + st.pair0 + (j - st.pairs_local));
    }
    loss[col] = sample_loss(st, ll);
    if (col == 0) {
      for (int k = 0; k < st.K; ++k) {
        double s;
        if (st.znorm) s = __ddiv_rn(__dadd_rn(ll[k], -st.zmean[k]), st.zstd[k]);
        else s = __dadd_rn(ll[1 + k], -ll[0]);
        st.red[st.N + st.S + 1 + k] = s;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (do_book == 1) bookkeeping(st);
    else if (do_book == 2) st.state_u64[0] += 1;     // get_grad: consume the draw, nothing else
  }
}

__global__ void nes_book_kernel(FbNesDev st) {
  FB_GRID_DEP_SYNC();
  if (st.flags[0]) return;
  if (threadIdx.x == 0 && blockIdx.x == 0) bookkeeping(st);
}

// Zero the parts of the reduction buffer other ranks own (multi-GPU only).
__global__ void nes_zero_red_kernel(FbNesDev st) {
  FB_GRID_DEP_SYNC();
  if (st.flags[0]) return;
  const int n = st.S + 1 + st.K;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) st.red[st.N + i] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// Gradient estimate (+ optional momentum / sign / clip update).  One thread per sample.
// mode 0: single GPU, estimate in numpy order and update.      mode 1: partial sum over local pairs -> red[n]
// mode 2: red[n] holds the all-reduced sum -> update.          mode 3: estimate only -> gest (get_grad)
// ------------------------------------------------------------------------------------------------
struct ColGetter {
  const double *noise; const float *noise32; const double *loss; int64_t N; int64_t n; int S2;
  __device__ double z(int j) const { return noise32 ? (double)noise32[(int64_t)j * N + n] : noise[(int64_t)j * N + n]; }
  __device__ double operator()(int i) const {
    // column i of `loss.flatten() * noise[:, 1:]` for row n (FAKEBOB.py:244)
    return (i < S2) ? __dmul_rn(loss[1 + i], z(i)) : __dmul_rn(loss[1 + i], -z(i - S2));
  }
};

__global__ void __launch_bounds__(128)
nes_update_kernel(FbNesDev st, int mode, int use_state_lr, double lr_arg) {
  FB_GRID_DEP_SYNC();
  if (st.flags[0]) return;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double dist = 0.0;
  if (n < st.N) {
    const double *loss = st.red + st.N;
    double g;
    if (mode == 0 || mode == 3) {
      ColGetter get{st.noise, st.noise32, loss, st.N, n, st.pairs_total};
      const double sum = np_pairwise(get, 0, st.S);
      g = __ddiv_rn(__ddiv_rn(sum, (double)st.S), st.sigma);
    } else if (mode == 1) {
      double acc = 0.0;
      ColGetter get{st.noise, st.noise32, loss, st.N, n, st.pairs_total};
      for (int j = 0; j < st.pairs_local; ++j) {
        const double z = get.z(j);
        acc = __dadd_rn(acc, __dmul_rn(loss[1 + st.pair0 + j], z));
        acc = __dadd_rn(acc, __dmul_rn(loss[1 + st.pairs_total + st.pair0 + j], -z));
      }
      st.red[n] = acc;
      return;
    } else {
      g = __ddiv_rn(__ddiv_rn(st.red[n], (double)st.S), st.sigma);
    }
    if (mode == 3) {
      st.gest[n] = g;
      return;
    }
    const double lr = use_state_lr ? st.state_f64[0] : lr_arg;
    const double G = __dadd_rn(__dmul_rn(st.momentum, st.grad[n]), __dmul_rn(st.one_minus_momentum, g));
    st.grad[n] = G;
    const double sg = (G > 0.0) ? 1.0 : ((G < 0.0) ? -1.0 : G);      // np.sign (0 -> 0, nan -> nan)
    double a = __dadd_rn(st.adver[n], -__dmul_rn(lr, sg));
    a = fmin(fmax(a, st.lower[n]), st.upper[n]);
    st.adver[n] = a;
    dist = fabs(__dadd_rn(st.audio[n], -a));
  }
  if (mode == 3 || mode == 1) return;
  // next iteration's pre-update distance = max |audio - adver| now
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dist = fmax(dist, __shfl_xor_sync(0xffffffffu, dist, o));
  if ((threadIdx.x & 31) == 0 && dist > 0.0)
    atomicMax(&st.dist_bits[st.flags[1]], (unsigned long long)__double_as_longlong(dist));
}

// Single-GPU gradient + update, one thread per sample: numpy's pairwise block (8 <= S <= 128) keeps 8 interleaved
// accumulators r_c = a[c] + a[c+8] + ..., combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), tail elements added last -- the
// same additions in the same order, so the result is bit-identical to np.mean(loss * noise, axis=1).  Lane = sample: every
// noise load of a warp is one contiguous 128-byte (float32) or 256-byte (float64) line, eight independent loads and
// add chains in flight per thread; the losses sit in shared memory.
// mode 0: estimate + update, mode 3: estimate only (get_grad).
__global__ void __launch_bounds__(128)
nes_update8_kernel(FbNesDev st, int mode) {
  FB_GRID_DEP_SYNC();
  if (st.flags[0]) return;
  __shared__ double s_loss[129];
  const int S = st.S, S2 = st.pairs_total;
  for (int i = threadIdx.x; i <= S; i += blockDim.x) s_loss[i] = st.red[st.N + i];
  __syncthreads();
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = n < st.N;
  const int64_t nn = valid ? n : st.N - 1;
  ColGetter get{st.noise, st.noise32, s_loss, st.N, nn, S2};
  const int full = S - (S % 8);
  double r[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) r[c] = get(c);
  for (int i = 8; i < full; i += 8) {
    double v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = get(i + c);
#pragma unroll
    for (int c = 0; c < 8; ++c) r[c] = __dadd_rn(r[c], v[c]);
  }
  double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])), __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
  for (int i = full; i < S; ++i) res = __dadd_rn(res, get(i));
  double dist = 0.0;
  if (valid) {
    const double g = __ddiv_rn(__ddiv_rn(res, (double)S), st.sigma);
    if (mode == 3) {
      st.gest[n] = g;
    } else {
      const double lr = st.state_f64[0];
      const double G = __dadd_rn(__dmul_rn(st.momentum, st.grad[n]), __dmul_rn(st.one_minus_momentum, g));
      st.grad[n] = G;
      const double sg = (G > 0.0) ? 1.0 : ((G < 0.0) ? -1.0 : G);      // np.sign (0 -> 0, nan -> nan)
      double a = __dadd_rn(st.adver[n], -__dmul_rn(lr, sg));
      a = fmin(fmax(a, st.lower[n]), st.upper[n]);
      st.adver[n] = a;
      dist = fabs(__dadd_rn(st.audio[n], -a));
    }
  }
  if (mode == 3) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dist = fmax(dist, __shfl_xor_sync(0xffffffffu, dist, o));
  if ((threadIdx.x & 31) == 0 && dist > 0.0)
    atomicMax(&st.dist_bits[st.flags[1]], (unsigned long long)__double_as_longlong(dist));
}

// ------------------------------------------------------------------------------------------------
// Multi-GPU iteration over peer memory (FbNesDev::xown != NULL): partial -> publish -> gather + bookkeeping -> gather + update.
// Writes are local, reads are remote (volatile, straight to the owner's L2 over NVLink); a flag per parity orders them.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ld_peer(const double *p) { return *reinterpret_cast<const volatile double *>(p); }

// partial gradient sum over this rank's pairs, one thread per sample (coalesced noise rows); -> own exchange buffer (or red)
__global__ void __launch_bounds__(128)
nes_partial_kernel(FbNesDev st) {
  FB_GRID_DEP_SYNC();
  if (st.flags[0]) return;
  __shared__ double s_lp[128], s_lm[128];
  const int P = st.pairs_local;
  const double *loss = st.red + st.N;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int par = st.flags[1] & 1;
  double *out = st.xown ? st.xown + (size_t)par * st.xstride : st.red;
  double acc = 0.0;
  for (int j0 = 0; j0 < P; j0 += 128) {
    const int nj = min(128, P - j0);
    __syncthreads();
    if ((int)threadIdx.x < nj) {
      s_lp[threadIdx.x] = loss[1 + st.pair0 + j0 + threadIdx.x];
      s_lm[threadIdx.x] = loss[1 + st.pairs_total + st.pair0 + j0 + threadIdx.x];
    }
    __syncthreads();
    if (n < st.N) {
      for (int j = 0; j < nj; ++j) {
        const double z = st.noise32 ? (double)st.noise32[(int64_t)(j0 + j) * st.N + n] : st.noise[(int64_t)(j0 + j) * st.N + n];
        acc = __dadd_rn(acc, __dmul_rn(s_lp[j], z));
        acc = __dadd_rn(acc, __dmul_rn(s_lm[j], -z));
      }
    }
  }
  if (n < st.N) out[n] = acc;
}

// copies this rank's losses / clean scores next to its partial and raises the flag of the iteration's parity
__global__ void __launch_bounds__(256)
nes_publish_kernel(FbNesDev st) {
  FB_GRID_DEP_SYNC();
  if (st.flags[0]) return;
  const int it = st.flags[1];
  const int par = it & 1;
  double *own = st.xown + (size_t)par * st.xstride;
  const int tail = st.S + 1 + st.K;
  for (int i = threadIdx.x; i < tail; i += blockDim.x) own[st.N + i] = st.red[st.N + i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long v = (*st.xsess << 32) | (unsigned long long)(it + 1);
    *reinterpret_cast<volatile unsigned long long *>(&st.xflag[par]) = v;
    __threadfence_system();
  }
}

// waits for every rank's flag, adds the loss / score tails in rank order into the local reduction buffer, bookkeeping
__global__ void __launch_bounds__(256)
nes_gather_book_kernel(FbNesDev st, int *__restrict__ err) {
  FB_GRID_DEP_SYNC();
  if (st.flags[0]) return;
  __shared__ int s_bad;
  const int it = st.flags[1];
  const int par = it & 1;
  const unsigned long long want = (*st.xsess << 32) | (unsigned long long)(it + 1);
  if (threadIdx.x == 0) s_bad = 0;
  __syncthreads();
  if ((int)threadIdx.x < st.world) {
    const volatile unsigned long long *f = st.xpeer_flag[threadIdx.x] + par;
    const long long t0 = clock64();
    while (*f < want) {
      if (clock64() - t0 > (6ll << 30)) { s_bad = 1; break; }       // ~3 s: a peer died; report instead of hanging the box
      __nanosleep(200);
    }
    __threadfence_system();
  }
  __syncthreads();
  if (s_bad) {
    if (threadIdx.x == 0) { atomicExch(err, 4); st.flags[0] = 1; }
    return;
  }
  const int tail = st.S + 1 + st.K;
  for (int i = threadIdx.x; i < tail; i += blockDim.x) {
    double acc = 0.0;
    for (int r = 0; r < st.world; ++r) acc = __dadd_rn(acc, ld_peer(st.xpeer[r] + (size_t)par * st.xstride + st.N + i));
    st.red[st.N + i] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) bookkeeping(st);
}

// gradient = sum of the W partials in rank order (identical on every rank), then the update; one thread per sample
__global__ void __launch_bounds__(128)
nes_gather_update_kernel(FbNesDev st) {
  FB_GRID_DEP_SYNC();
  if (st.flags[0]) return;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int par = (st.flags[1] - 1) & 1;                        // bookkeeping already advanced the iteration counter
  double dist = 0.0;
  if (n < st.N) {
    double v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = (r < st.world) ? ld_peer(st.xpeer[r] + (size_t)par * st.xstride + n) : 0.0;
    double sum = v[0];
#pragma unroll
    for (int r = 1; r < 8; ++r)
      if (r < st.world) sum = __dadd_rn(sum, v[r]);
    const double g = __ddiv_rn(__ddiv_rn(sum, (double)st.S), st.sigma);
    const double lr = st.state_f64[0];
    const double G = __dadd_rn(__dmul_rn(st.momentum, st.grad[n]), __dmul_rn(st.one_minus_momentum, g));
    st.grad[n] = G;
    const double sg = (G > 0.0) ? 1.0 : ((G < 0.0) ? -1.0 : G);
    double a = __dadd_rn(st.adver[n], -__dmul_rn(lr, sg));
    a = fmin(fmax(a, st.lower[n]), st.upper[n]);
    st.adver[n] = a;
    dist = fabs(__dadd_rn(st.audio[n], -a));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) dist = fmax(dist, __shfl_xor_sync(0xffffffffu, dist, o));
  if ((threadIdx.x & 31) == 0 && dist > 0.0)
    atomicMax(&st.dist_bits[st.flags[1]], (unsigned long long)__double_as_longlong(dist));
}

// adver <- clip(adver - lr * sign(momentum * grad + (1 - momentum) * gest))   (after get_grad)
__global__ void __launch_bounds__(128) nes_apply_kernel(FbNesDev st, double lr) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= st.N) return;
  const double G = __dadd_rn(__dmul_rn(st.momentum, st.grad[n]), __dmul_rn(st.one_minus_momentum, st.gest[n]));
  st.grad[n] = G;
  const double sg = (G > 0.0) ? 1.0 : ((G < 0.0) ? -1.0 : G);
  double a = __dadd_rn(st.adver[n], -__dmul_rn(lr, sg));
  st.adver[n] = fmin(fmax(a, st.lower[n]), st.upper[n]);
}

__global__ void nes_init_kernel(FbNesDev st) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= st.N) return;
  const double a = st.audio[n];
  st.adver[n] = a;
  st.grad[n] = 0.0;
  st.lower[n] = fmin(fmax(__dadd_rn(a, -st.epsilon), -1.0), 1.0);
  st.upper[n] = fmin(fmax(__dadd_rn(a, st.epsilon), -1.0), 1.0);
}

// ------------------------------------------------------------------------------------------------
// Host side
// ------------------------------------------------------------------------------------------------
// The batch workspace is shared with the plain score() entry points; re-establish the NES layout when
// someone else used it, and drop the captured graph if any buffer was reallocated meanwhile.
static void nes_drop_graph(FbNes *s) {
  if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
  if (s->graph) cudaGraphDestroy(s->graph);
  s->graph_exec = nullptr;
  s->graph = nullptr;
}

static int nes_claim_batch(fb_ctx *ctx) {
  FbNes *s = ctx->nes;
  int rc;
  if (ctx->arch == 1) ctx->need_feats_f32 = true;
  if (ctx->batch_tag != 1) {
    if ((rc = ctx->wave.ensure((size_t)s->B_local * s->N + 8))) return rc;
    if ((rc = fb_reserve_batch(ctx, s->B_local, s->offsets.data()))) return rc;
    ctx->batch_tag = 1;
  }
  if (ctx->arch == 1 && ctx->iv)
    if ((rc = fb_ivector_reserve(ctx))) return rc;       // before any graph capture
  if (s->graph_exec && (s->graph_epoch != fb_alloc_epoch() || s->graph_arch != ctx->arch ||
                        memcmp(&s->graph_dev, &s->dev, sizeof(FbNesDev)) != 0))
    nes_drop_graph(s);
  return FB_OK;
}

void fb_nes_destroy(fb_ctx *ctx) {
  FbNes *s = ctx->nes;
  if (!s) return;
  nes_drop_graph(s);
  s->f64_pool.release(); s->noise64.release(); s->noise32.release(); s->flags.release(); s->dist_bits.release(); s->ext_scores.release();
  delete s;
  ctx->nes = nullptr;
}

static FbNesDev nes_dev(const FbNes *s) { return s->dev; }

extern "C" int fb_nes_init(fb_ctx *ctx, const fb_nes_params *p, const double *audio_host, int64_t n_samples) {
  FB_CHECK_ARG(ctx && p && audio_host, "NULL argument");
  FB_CHECK_ARG(n_samples >= 8 && n_samples < (1ll << 30), "n_samples out of range");
  FB_CHECK_ARG(p->samples_per_draw >= 2, "samples_per_draw must be >= 2");
  FB_CHECK_ARG(p->n_speakers >= 1 && p->n_speakers <= FB_MAX_MODELS - 1, "n_speakers out of range");
  FB_CHECK_ARG(p->plateau_length >= 1 && p->plateau_length <= 64, "plateau_length must be in [1,64]");
  FB_CHECK_ARG(p->max_iter >= 1, "max_iter must be >= 1");
  const int K = p->n_speakers;
  const bool ext = p->external_scorer != 0;
  const bool iv = !ext && ctx->arch == 1;
  if (ext) {
    int rk = 0, wd = 1;
    fb_comm_info(ctx, &rk, &wd);
    FB_CHECK_ARG(wd == 1, "black-box scorers run single-GPU");
  } else if (iv) {
    FB_CHECK_ARG(ctx->iv && ctx->iv->K == K, "enrolled i-vectors do not match n_speakers");
    FB_CHECK_ARG(p->z_norm_means && p->z_norm_stds, "i-vector scorers need z-norm statistics");
  } else {
    const int need_models = (p->task == FB_TASK_CSI) ? K : K + 1;
    FB_CHECK_ARG(ctx->n_models == need_models, "loaded model slots do not match task / n_speakers");
  }
  if (p->task == FB_TASK_SV) FB_CHECK_ARG(K == 1, "SV needs exactly one speaker");
  if (p->task == FB_TASK_CSI || (p->task == FB_TASK_OSI && p->targeted)) {
    FB_CHECK_ARG(K >= 2, "this loss needs at least two speakers");
    FB_CHECK_ARG(p->label >= 0 && p->label < K, "label (true/target) out of range");
  }
  if (p->task == FB_TASK_CSI && !ext) FB_CHECK_ARG(p->z_norm_means && p->z_norm_stds, "CSI needs z-norm statistics");
  FB_CUDA(cudaSetDevice(ctx->device));
  if (!ctx->nes) ctx->nes = new FbNes();
  FbNes *s = ctx->nes;
  s->p = *p;
  s->p.z_norm_means = s->p.z_norm_stds = nullptr;     // caller-owned host arrays: copied to the device below, not kept
  s->N = n_samples;
  s->enqueued = 0;
  const int pairs_total = p->samples_per_draw / 2;
  const int S = 2 * pairs_total;
  int rank = 0, world = 1;
  fb_comm_info(ctx, &rank, &world);
  // floor partition (sharding.pair_range): when the pairs do not divide evenly the LOW ranks get the smaller share, so
  // rank 0, which also scores the clean audio, never holds more audios than the largest rank
  // (25 pairs on 8 ranks: 3,3,3,3,3,3,3,4 pairs -> 7,6,6,6,6,6,6,8 audios).
  const int p0 = (int)((int64_t)pairs_total * rank / world), p1 = (int)((int64_t)pairs_total * (rank + 1) / world);
  s->pairs_local = p1 - p0;
  s->pair0 = p0;
  s->world = world;
  s->rank = rank;
  s->has_clean = (rank == 0);
  s->B_local = (s->has_clean ? 1 : 0) + 2 * s->pairs_local;
  FB_CHECK_ARG(s->B_local > 0, "more ranks than antithetic pairs: this rank has nothing to score");
  // float64 pool: audio adver lower upper grad gest | red (N + S+1 + K) | zmean zstd | state | log | threshold
  const size_t N = (size_t)n_samples;
  const size_t red_n = N + S + 1 + K;
  const size_t red_pad = (red_n + 7) & ~(size_t)7;
  const size_t log_n = (size_t)(p->max_iter + 1) * (4 + K);
  const size_t pool_n = 6 * N + red_pad + 2 * K + 128 + log_n + 8;
  const bool philox = (p->rng == FB_RNG_PHILOX);
  const size_t noise_n = (size_t)(s->pairs_local > 0 ? s->pairs_local : 1) * N;
  int rc;
  if ((rc = s->f64_pool.ensure(pool_n))) return rc;
  if (philox) { if ((rc = s->noise32.ensure(noise_n))) return rc; }
  else { if ((rc = s->noise64.ensure(noise_n))) return rc; }
  if ((rc = s->flags.ensure(8 + 16))) return rc;
  if ((rc = s->dist_bits.ensure((size_t)p->max_iter + 2))) return rc;
  // everything except the six state vectors (nes_init_kernel / the audio copy write those) starts from zero
  FB_CUDA(cudaMemsetAsync(s->f64_pool.p + 4 * N, 0, (pool_n - 4 * N) * sizeof(double), ctx->stream));
  FB_CUDA(cudaMemsetAsync(s->flags.p, 0, (8 + 16) * sizeof(unsigned long long), ctx->stream));
  FB_CUDA(cudaMemsetAsync(s->dist_bits.p, 0, ((size_t)p->max_iter + 2) * sizeof(unsigned long long), ctx->stream));
  double *q = s->f64_pool.p;
  FbNesDev &d = s->dev;
  memset(&d, 0, sizeof(d));                            // padding bytes too: the struct is compared to decide graph reuse
  d.audio = q; q += N; d.adver = q; q += N; d.lower = q; q += N; d.upper = q; q += N; d.grad = q; q += N; d.gest = q; q += N;
  d.red = q; q += red_pad;
  d.zmean = q; q += K; d.zstd = q; q += K;
  d.state_f64 = q; q += 128;
  d.log = q; q += log_n;
  d.threshold = q; q += 8;
  s->red_count = red_n;
  d.noise = philox ? nullptr : s->noise64.p;
  d.noise32 = philox ? s->noise32.p : nullptr;
  d.flags = reinterpret_cast<int *>(s->flags.p);
  d.state_u64 = s->flags.p + 8;
  d.dist_bits = s->dist_bits.p;
  d.N = n_samples; d.S = S; d.K = K; d.pairs_total = pairs_total; d.pairs_local = s->pairs_local; d.pair0 = p0;
  d.B_local = s->B_local; d.has_clean = s->has_clean ? 1 : 0;
  d.znorm = (ext || iv || p->task == FB_TASK_CSI) ? 1 : 0;      // black-box scores arrive final: (s - 0) / 1
  d.task = p->task; d.targeted = p->targeted; d.label = p->label; d.plateau_length = p->plateau_length;
  d.auto_stop = 1;
  d.kappa = p->adver_thresh; d.sigma = p->sigma; d.epsilon = p->epsilon; d.momentum = p->momentum;
  d.one_minus_momentum = 1.0 - p->momentum;
  d.min_lr = p->min_lr; d.plateau_drop = p->plateau_drop; d.seed = p->seed;
  d.world = world; d.rank = rank;
  if ((rc = fb_comm_p2p_attach(ctx, &d, red_n))) return rc;
  FB_CUDA(cudaMemcpyAsync(d.audio, audio_host, N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<double> ext_one(K, 1.0);
  if (ext) {
    FB_CUDA(cudaMemcpyAsync(d.zstd, ext_one.data(), K * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));   // zmean stays 0
  } else if (d.znorm) {
    FB_CUDA(cudaMemcpyAsync(d.zmean, p->z_norm_means, K * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    FB_CUDA(cudaMemcpyAsync(d.zstd, p->z_norm_stds, K * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  const double st0[1] = {p->max_lr};
  FB_CUDA(cudaMemcpyAsync(d.state_f64, st0, sizeof(st0), cudaMemcpyHostToDevice, ctx->stream));
  FB_CUDA(cudaMemcpyAsync(d.threshold, &p->threshold, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  FB_CUDA(cudaMemcpyAsync(d.state_u64, &p->draw_base, sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
  nes_init_kernel<<<fb_div_up(n_samples, 256), 256, 0, ctx->stream>>>(d);
  FB_CUDA(cudaGetLastError());
  // batch layout: B_local utterances of n_samples each, back to back.  Always re-established: the previous session (or a
  // score() call) may have left the shared batch workspace laid out for another length or batch size.
  s->offsets.resize(s->B_local + 1);
  for (int b = 0; b <= s->B_local; ++b) s->offsets[b] = (int64_t)b * n_samples;
  if (ext) {
    // no front-end / model workspace: the wave buffer for the batch and a score buffer [B][K]
    if ((rc = ctx->wave.ensure((size_t)s->B_local * s->N + 8))) return rc;
    if ((rc = s->ext_scores.ensure((size_t)s->B_local * K))) return rc;
    if ((rc = ctx->misc.ensure(8, true))) return rc;          // status word read by fb_nes_status
    ctx->batch_tag = -1;
  } else {
    if ((rc = fb_prepare_tables(ctx))) return rc;
    ctx->batch_tag = -1;
    if ((rc = nes_claim_batch(ctx))) return rc;
  }
  FB_CUDA(cudaStreamSynchronize(ctx->stream));        // the host arrays (audio, z-norm, locals above) may go away now
  return FB_OK;
}

extern "C" int fb_nes_ext_perturb(fb_ctx *ctx, int16_t *wave_host) {
  FB_CHECK_ARG(ctx && ctx->nes && wave_host, "fb_nes_init has not been called");
  FbNes *s = ctx->nes;
  FB_CHECK_ARG(s->p.external_scorer, "the session was not created for an external scorer");
  FB_CHECK_ARG(s->p.rng == FB_RNG_PHILOX, "external scorers use the device noise generator");
  FB_CUDA(cudaSetDevice(ctx->device));
  const FbNesDev d = s->dev;
  dim3 gp(fb_div_up(fb_div_up(s->N, 4), 256), s->pairs_local < 8 ? s->pairs_local : 8);
  FB_CUDA(fb_launch(perturb_kernel, gp, dim3(256), 0, ctx->stream, d, ctx->wave.p, s->N, 1));
  ctx->launches += 1;
  FB_CUDA(cudaMemcpyAsync(wave_host, ctx->wave.p, (size_t)s->B_local * s->N * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  return FB_OK;
}

extern "C" int fb_nes_read_gest(fb_ctx *ctx, double *gest_host, int64_t n_samples, double *losses_host, double *score0_host) {
  FB_CHECK_ARG(ctx && ctx->nes && n_samples == ctx->nes->N, "bad argument");
  FbNes *s = ctx->nes;
  if (gest_host) FB_CUDA(cudaMemcpyAsync(gest_host, s->dev.gest, n_samples * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (losses_host)
    FB_CUDA(cudaMemcpyAsync(losses_host, s->dev.red + s->N, (size_t)(s->dev.S + 1) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (score0_host)
    FB_CUDA(cudaMemcpyAsync(score0_host, s->dev.red + s->N + s->dev.S + 1, (size_t)s->dev.K * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  return FB_OK;
}

extern "C" int fb_nes_ext_update(fb_ctx *ctx, const double *scores_host, int gradient_only) {
  FB_CHECK_ARG(ctx && ctx->nes && scores_host, "fb_nes_init has not been called");
  FbNes *s = ctx->nes;
  FB_CHECK_ARG(s->p.external_scorer, "the session was not created for an external scorer");
  FB_CUDA(cudaSetDevice(ctx->device));
  const FbNesDev d = s->dev;
  FB_CUDA(cudaMemcpyAsync(s->ext_scores.p, scores_host, (size_t)s->B_local * d.K * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  const int g = gradient_only ? 1 : 0;
  FB_CUDA(fb_launch(nes_loss_kernel, dim3(1), dim3(256), 0, ctx->stream, d, (const double *)s->ext_scores.p, d.K, g ? 2 : 1));
  if (d.S >= 8 && d.S <= 128)
    FB_CUDA(fb_launch(nes_update8_kernel, dim3(fb_div_up(s->N, 128)), dim3(128), 0, ctx->stream, d, g ? 3 : 0));
  else
    FB_CUDA(fb_launch(nes_update_kernel, dim3(fb_div_up(s->N, 128)), dim3(128), 0, ctx->stream, d, g ? 3 : 0, 1, 0.0));
  ctx->launches += 2;
  if (!g) s->enqueued += 1;
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  return FB_OK;
}

static int nes_enqueue_iteration(fb_ctx *ctx, int mode_get_grad) {
  FbNes *s = ctx->nes;
  const FbNesDev d = nes_dev(s);
  const int philox = (s->p.rng == FB_RNG_PHILOX);
  int rc;
  dim3 gp(fb_div_up(fb_div_up(s->N, 4), 256), s->pairs_local > 0 ? (s->pairs_local < 8 ? s->pairs_local : 8) : 1);
  fb_prof_mark(ctx, -1);
  FbNvtxSeq nv;
  nv.next("fb:nes_perturb");
  FB_CUDA(fb_launch(perturb_kernel, gp, dim3(256), 0, ctx->stream, d, ctx->wave.p, s->N, philox));
  fb_prof_mark(ctx, 0);
  ctx->launches += 1;
  nv.next("fb:nes_score");
  if ((rc = fb_run_frontend_flag(ctx, d.flags))) return rc;
  const double *ll_dev;
  int ll_stride;
  if (ctx->arch == 1) {
    if ((rc = fb_run_ivector_flag(ctx, d.flags, true))) return rc;
    ll_dev = ctx->iv->scores.p;
    ll_stride = ctx->iv->K;
  } else {
    if ((rc = fb_run_gmm_flag(ctx, d.flags))) return rc;
    ll_dev = ctx->avg_ll.p;
    ll_stride = ctx->n_models;
  }
  const int multi = s->world > 1;
  nv.next("fb:nes_loss");
  if (multi) {
    FB_CUDA(fb_launch(nes_zero_red_kernel, dim3(1), dim3(256), 0, ctx->stream, d));
    ctx->launches += 1;
  }
  FB_CUDA(fb_launch(nes_loss_kernel, dim3(1), dim3(256), 0, ctx->stream, d, ll_dev, ll_stride, mode_get_grad ? 2 : (multi ? 0 : 1)));
  fb_prof_mark(ctx, 6);
  ctx->launches += 1;
  const int nb = fb_div_up(s->N, 128);
  nv.next(multi ? "fb:nes_exchange_update" : "fb:nes_update");
  if (!multi) {
    if (d.S >= 8 && d.S <= 128)
      FB_CUDA(fb_launch(nes_update8_kernel, dim3(fb_div_up(s->N, 128)), dim3(128), 0, ctx->stream, d, mode_get_grad ? 3 : 0));
    else
      FB_CUDA(fb_launch(nes_update_kernel, dim3(nb), dim3(128), 0, ctx->stream, d, mode_get_grad ? 3 : 0, 1, 0.0));
    ctx->launches += 1;
  } else {
    FB_CUDA(fb_launch(nes_partial_kernel, dim3(nb), dim3(128), 0, ctx->stream, d));
    if (d.xown) {
      // one-shot exchange over peer memory: no collective kernel, the W partials are added inside the update
      FB_CUDA(fb_launch(nes_publish_kernel, dim3(1), dim3(256), 0, ctx->stream, d));
      FB_CUDA(fb_launch(nes_gather_book_kernel, dim3(1), dim3(256), 0, ctx->stream, d, ctx->misc.p + 1));
      FB_CUDA(fb_launch(nes_gather_update_kernel, dim3(nb), dim3(128), 0, ctx->stream, d));
      ctx->launches += 4;
    } else {
      if ((rc = fb_comm_allreduce_f64(ctx, d.red, s->red_count))) return rc;
      FB_CUDA(fb_launch(nes_book_kernel, dim3(1), dim3(32), 0, ctx->stream, d));
      FB_CUDA(fb_launch(nes_update_kernel, dim3(nb), dim3(128), 0, ctx->stream, d, 2, 1, 0.0));
      ctx->launches += 3;
    }
  }
  fb_prof_mark(ctx, 7);
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_nes_run(fb_ctx *ctx, int n_iters, const double *noise_host) {
  FB_CHECK_ARG(ctx && ctx->nes, "fb_nes_init has not been called");
  FbNes *s = ctx->nes;
  FB_CHECK_ARG(n_iters >= 0, "n_iters must be >= 0");
  FB_CUDA(cudaSetDevice(ctx->device));
  const bool host_rng = (s->p.rng == FB_RNG_HOST);
  FB_CHECK_ARG(!host_rng || noise_host, "rng = HOST needs noise_host");
  const int remaining = s->p.max_iter - s->enqueued;
  if (n_iters > remaining) n_iters = remaining;
  int rc;
  if ((rc = nes_claim_batch(ctx))) return rc;
  const size_t per_iter = (size_t)s->dev.pairs_total * s->N;
  static const bool no_graph = getenv("FB_NO_GRAPH") != nullptr;
  for (int i = 0; i < n_iters; ++i) {
    if (host_rng) {
      // pair-major (S/2, N) float64 block for this iteration; this rank keeps its own pairs
      const double *src = noise_host + (size_t)i * per_iter + (size_t)s->pair0 * s->N;
      FB_CUDA(cudaMemcpyAsync(s->noise64.p, src, (size_t)s->pairs_local * s->N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    }
    if (no_graph || host_rng || ctx->prof_on) {
      if ((rc = nes_enqueue_iteration(ctx, 0))) return rc;
    } else {
      if (!s->graph_exec) {
        const int64_t before = ctx->launches;
        FB_CUDA(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        rc = nes_enqueue_iteration(ctx, 0);
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &s->graph);
        if (rc) return rc;
        if (e != cudaSuccess) { fb_set_error("graph capture failed: %s", cudaGetErrorString(e)); return FB_ERR_CUDA; }
        FB_CUDA(cudaGraphInstantiate(&s->graph_exec, s->graph, 0));
        s->graph_epoch = fb_alloc_epoch();
        s->graph_dev = s->dev;
        s->graph_arch = ctx->arch;
        s->launches_per_iter = ctx->launches - before;
        ctx->launches = before;
      }
      FB_CUDA(cudaGraphLaunch(s->graph_exec, ctx->stream));
      ctx->launches += s->launches_per_iter;
    }
    s->enqueued += 1;
  }
  return FB_OK;
}

extern "C" int fb_nes_status(fb_ctx *ctx, int *iters_done, int *stopped) {
  FB_CHECK_ARG(ctx && ctx->nes, "fb_nes_init has not been called");
  int h[4];
  int err[2] = {0, 0};
  // both read-backs ride the stream, one host wait for the pair
  FB_CUDA(cudaMemcpyAsync(h, ctx->nes->dev.flags, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(cudaMemcpyAsync(err, ctx->misc.p + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  if (err[0] != 0) return fb_map_device_error(ctx, err[0]);
  if (iters_done) *iters_done = h[1];
  if (stopped) *stopped = h[0];
  return FB_OK;
}

extern "C" int fb_nes_read_log(fb_ctx *ctx, double *rows_host, int max_rows) {
  FB_CHECK_ARG(ctx && ctx->nes && rows_host, "bad argument");
  int done = 0;
  int rc = fb_nes_status(ctx, &done, nullptr);
  if (rc) return rc;
  const int n = done < max_rows ? done : max_rows;
  FB_CUDA(cudaMemcpy(rows_host, ctx->nes->dev.log, (size_t)n * (4 + ctx->nes->dev.K) * sizeof(double), cudaMemcpyDeviceToHost));
  return n;
}

extern "C" int fb_nes_read_adver(fb_ctx *ctx, double *adver_host, int64_t n_samples) {
  FB_CHECK_ARG(ctx && ctx->nes && adver_host && n_samples == ctx->nes->N, "bad argument");
  FB_CUDA(cudaMemcpyAsync(adver_host, ctx->nes->dev.adver, n_samples * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  return FB_OK;
}

extern "C" int fb_nes_read_grad(fb_ctx *ctx, double *grad_host, int64_t n_samples) {
  FB_CHECK_ARG(ctx && ctx->nes && grad_host && n_samples == ctx->nes->N, "bad argument");
  FB_CUDA(cudaMemcpyAsync(grad_host, ctx->nes->dev.grad, n_samples * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  return FB_OK;
}

extern "C" int fb_nes_set_threshold(fb_ctx *ctx, double threshold) {
  FB_CHECK_ARG(ctx && ctx->nes, "fb_nes_init has not been called");
  ctx->nes->p.threshold = threshold;
  FB_CUDA(cudaMemcpyAsync(ctx->nes->dev.threshold, &ctx->nes->p.threshold, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  return FB_OK;
}

extern "C" int fb_nes_estimate_begin(fb_ctx *ctx, double accept_threshold) {
  FB_CHECK_ARG(ctx && ctx->nes, "fb_nes_init has not been called");
  FbNes *s = ctx->nes;
  FB_CHECK_ARG(s->p.task != FB_TASK_CSI && !s->p.targeted, "estimate mode is the untargeted OSI / SV search (FAKEBOB.py:41-43,73-74)");
  FB_CHECK_ARG(s->world == 1, "estimate mode is single-GPU only");
  s->dev.est_mode = 1;
  s->dev.accept_threshold = accept_threshold;        // the captured graph is keyed on the device arguments: re-captured on next run
  return FB_OK;
}

extern "C" int fb_nes_continue(fb_ctx *ctx, double threshold) {
  FB_CHECK_ARG(ctx && ctx->nes, "fb_nes_init has not been called");
  FbNes *s = ctx->nes;
  FB_CUDA(cudaSetDevice(ctx->device));
  s->p.threshold = threshold;
  const double lr0 = s->p.max_lr;
  FB_CUDA(cudaMemcpyAsync(s->dev.threshold, &s->p.threshold, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  FB_CUDA(cudaMemcpyAsync(s->dev.state_f64, &lr0, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));   // lr = max_lr (FAKEBOB.py:82)
  FB_CUDA(cudaMemsetAsync(s->dev.flags, 0, sizeof(int), ctx->stream));                                       // not stopped
  FB_CUDA(cudaMemsetAsync(s->dev.flags + 3, 0, sizeof(int), ctx->stream));                                   // last_ls = []
  int done = 0;                                                      // iterations enqueued after the stop were no-ops
  FB_CUDA(cudaMemcpyAsync(&done, s->dev.flags + 1, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  s->enqueued = done;
  return FB_OK;
}

extern "C" int fb_nes_get_grad(fb_ctx *ctx, const double *noise_host, double *final_loss, double *adver_loss,
                               double *score0_host, double *grad_host) {
  FB_CHECK_ARG(ctx && ctx->nes, "fb_nes_init has not been called");
  FbNes *s = ctx->nes;
  FB_CHECK_ARG(s->world == 1, "fb_nes_get_grad is single-GPU only");
  const bool host_rng = (s->p.rng == FB_RNG_HOST);
  FB_CHECK_ARG(!host_rng || noise_host, "rng = HOST needs noise_host");
  FB_CUDA(cudaSetDevice(ctx->device));
  if (host_rng)
    FB_CUDA(cudaMemcpyAsync(s->noise64.p, noise_host, (size_t)s->pairs_local * s->N * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  int rc;
  if ((rc = nes_claim_batch(ctx))) return rc;
  if ((rc = nes_enqueue_iteration(ctx, 1))) return rc;
  const int S = s->dev.S, K = s->dev.K;
  std::vector<double> tail(S + 1 + K);
  FB_CUDA(cudaMemcpyAsync(tail.data(), s->dev.red + s->N, tail.size() * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (grad_host)
    FB_CUDA(cudaMemcpyAsync(grad_host, s->dev.gest, s->N * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  FB_CUDA(cudaStreamSynchronize(ctx->stream));
  if ((rc = fb_check_device_error(ctx))) return rc;
  if (adver_loss) *adver_loss = tail[0];
  if (final_loss) {
    // np.mean(loss[1:]) in numpy's pairwise order, on the host
    struct H { static double pw(const double *a, int n) {
      if (n < 8) { double r = 0.0; for (int i = 0; i < n; ++i) r += a[i]; return r; }
      if (n <= 128) {
        double r[8]; for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i = 8;
        for (; i < n - (n % 8); i += 8) for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
      }
      int n2 = n / 2; n2 -= n2 % 8;
      return pw(a, n2) + pw(a + n2, n - n2);
    } };
    *final_loss = H::pw(tail.data() + 1, S) / (double)S;
  }
  if (score0_host) memcpy(score0_host, tail.data() + S + 1, K * sizeof(double));
  return FB_OK;
}

extern "C" int fb_nes_apply_update(fb_ctx *ctx, double lr) {
  FB_CHECK_ARG(ctx && ctx->nes, "fb_nes_init has not been called");
  FbNes *s = ctx->nes;
  nes_apply_kernel<<<fb_div_up(s->N, 128), 128, 0, ctx->stream>>>(s->dev, lr);
  ctx->launches += 1;
  FB_CUDA(cudaGetLastError());
  return FB_OK;
}

extern "C" int fb_nes_kernel_launches(fb_ctx *ctx, int64_t *count) {
  FB_CHECK_ARG(ctx && count, "NULL argument");
  *count = ctx->launches;
  return FB_OK;
}
