// Multi-GPU plumbing: one process per GPU; the antithetic pairs of a NES draw are sharded across
// ranks and one ncclAllReduce(sum, float64) per iteration combines
//   [gradient partial sums (N) | per-sample losses (S+1, each owned by one rank) | clean scores (K)]
// on the context's stream (capturable in the per-iteration CUDA graph).
// Nothing in the reference communicates; this is new for the B200 build.
// NCCL is resolved with dlopen at fb_comm_init time so the library loads on hosts without it.
#include "fb_common.cuh"
#include "fb_nes.cuh"
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef int ncclResult_t;

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId_t *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId_t, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

struct FbComm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  // peer-memory exchange buffers (one allocation per rank: [2][xcap] doubles, then flags [2] and the session word)
  void *xbase = nullptr;
  size_t xcap = 0;                 // doubles per parity
  void *peer_base[8] = {nullptr};
  bool p2p = false;
  unsigned long long session = 0;
};
static constexpr size_t kXTail = 256;   // bytes after the two data buffers: flags [2] at +0, session word at +64

static int load_nccl() {
  if (g_nccl.lib) return FB_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.lib) break;
  }
  if (!g_nccl.lib) {
    fb_set_error("cannot dlopen libnccl.so.2: %s", dlerror());
    return FB_ERR_NCCL;
  }
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.lib, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.lib, "ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.lib, "ncclCommDestroy");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.lib, "ncclAllReduce");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.lib, "ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce) {
    fb_set_error("libnccl is missing required symbols");
    return FB_ERR_NCCL;
  }
  return FB_OK;
}

#define FB_NCCL(call)                                                                              \
  do {                                                                                             \
    ncclResult_t r_ = (call);                                                                      \
    if (r_ != 0) {                                                                                 \
      fb_set_error("%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
      return FB_ERR_NCCL;                                                                          \
    }                                                                                              \
  } while (0)

extern "C" int fb_comm_unique_id(void *out_128_bytes) {
  FB_CHECK_ARG(out_128_bytes != nullptr, "out is NULL");
  int rc = load_nccl();
  if (rc) return rc;
  ncclUniqueId_t id;
  FB_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(out_128_bytes, &id, 128);
  return FB_OK;
}

extern "C" int fb_comm_init(fb_ctx *ctx, const void *unique_id_128_bytes, int rank, int world) {
  FB_CHECK_ARG(ctx && unique_id_128_bytes, "NULL argument");
  FB_CHECK_ARG(world >= 1 && rank >= 0 && rank < world, "bad rank / world");
  int rc = load_nccl();
  if (rc) return rc;
  FB_CUDA(cudaSetDevice(ctx->device));
  if (ctx->comm) fb_comm_destroy(ctx);
  FbComm *c = new FbComm();
  ncclUniqueId_t id;
  memcpy(&id, unique_id_128_bytes, 128);
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, id, rank);
  if (r != 0) {
    fb_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    delete c;
    return FB_ERR_NCCL;
  }
  c->rank = rank;
  c->world = world;
  ctx->comm = c;
  return FB_OK;
}

// ---- peer-memory exchange ---------------------------------------------------------------------------------------------
// New (nothing in the reference communicates).  Each rank allocates one buffer, exports it with cudaIpcGetMemHandle, the host
// (torch.distributed) gathers the handles, every rank maps all of them.  Data is only ever WRITTEN locally and READ remotely.
extern "C" int fb_comm_p2p_export(fb_ctx *ctx, int64_t max_samples, void *handle_out_64_bytes) {
  FB_CHECK_ARG(ctx && ctx->comm && handle_out_64_bytes, "fb_comm_init must be called first");
  FB_CHECK_ARG(max_samples > 0 && max_samples <= (1ll << 26), "max_samples out of range");
  FB_CHECK_ARG(ctx->comm->world <= 8, "the peer exchange supports up to 8 ranks");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  FbComm *c = ctx->comm;
  FB_CUDA(cudaSetDevice(ctx->device));
  if (c->xbase) { cudaFree(c->xbase); c->xbase = nullptr; }
  c->xcap = (size_t)max_samples + 2048;                       // + losses (S + 1) + clean scores (K)
  const size_t bytes = 2 * c->xcap * sizeof(double) + kXTail;
  FB_CUDA(cudaMalloc(&c->xbase, bytes));
  FB_CUDA(cudaMemset(c->xbase, 0, bytes));
  FB_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  FB_CUDA(cudaIpcGetMemHandle(&h, c->xbase));
  memcpy(handle_out_64_bytes, &h, 64);
  return FB_OK;
}

extern "C" int fb_comm_p2p_import(fb_ctx *ctx, const void *handles_world_x_64_bytes) {
  FB_CHECK_ARG(ctx && ctx->comm && handles_world_x_64_bytes, "NULL argument");
  FbComm *c = ctx->comm;
  FB_CHECK_ARG(c->xbase != nullptr, "fb_comm_p2p_export must be called first");
  FB_CUDA(cudaSetDevice(ctx->device));
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) { c->peer_base[r] = c->xbase; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char *)handles_world_x_64_bytes + 64 * r, 64);
    void *p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      fb_set_error("cudaIpcOpenMemHandle for rank %d failed: %s (falling back to ncclAllReduce)", r, cudaGetErrorString(e));
      c->p2p = false;
      return FB_ERR_UNSUPPORTED;
    }
    c->peer_base[r] = p;
  }
  c->p2p = true;
  fb_bump_alloc_epoch();
  return FB_OK;
}

int fb_comm_p2p_attach(fb_ctx *ctx, FbNesDev *d, size_t count) {
  d->xown = nullptr;
  FbComm *c = ctx->comm;
  if (!c || c->world == 1 || !c->p2p || count > c->xcap || getenv("FB_NO_P2P")) return FB_OK;
  char *own = (char *)c->xbase;
  d->xown = (double *)own;
  d->xflag = (unsigned long long *)(own + 2 * c->xcap * sizeof(double));
  d->xsess = (const unsigned long long *)(own + 2 * c->xcap * sizeof(double) + 64);
  d->xstride = c->xcap;
  d->world = c->world;
  d->rank = c->rank;
  for (int r = 0; r < 8; ++r) {
    char *pb = (char *)c->peer_base[r < c->world ? r : c->rank];
    d->xpeer[r] = (const double *)pb;
    d->xpeer_flag[r] = (const unsigned long long *)(pb + 2 * c->xcap * sizeof(double));
  }
  // a new session: every rank calls fb_nes_init the same number of times, so the numbers agree without communication
  c->session += 1;
  FB_CUDA(cudaMemcpyAsync((void *)d->xsess, &c->session, sizeof(unsigned long long), cudaMemcpyHostToDevice, ctx->stream));
  return FB_OK;
}

int fb_comm_destroy_impl(fb_ctx *ctx) {
  if (ctx && ctx->comm) {
    for (int r = 0; r < ctx->comm->world && r < 8; ++r)
      if (r != ctx->comm->rank && ctx->comm->peer_base[r]) cudaIpcCloseMemHandle(ctx->comm->peer_base[r]);
    if (ctx->comm->xbase) cudaFree(ctx->comm->xbase);
    if (ctx->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm->comm);
    delete ctx->comm;
    ctx->comm = nullptr;
  }
  return FB_OK;
}

extern "C" int fb_comm_destroy(fb_ctx *ctx) { return fb_comm_destroy_impl(ctx); }

void fb_comm_info(fb_ctx *ctx, int *rank, int *world) {
  *rank = ctx->comm ? ctx->comm->rank : 0;
  *world = ctx->comm ? ctx->comm->world : 1;
}

int fb_comm_allreduce_f64(fb_ctx *ctx, double *buf, size_t count) {
  if (!ctx->comm || ctx->comm->world == 1) return FB_OK;
  FB_NCCL(g_nccl.AllReduce(buf, buf, count, /*ncclFloat64*/ 8, /*ncclSum*/ 0, ctx->comm->comm, ctx->stream));
  return FB_OK;
}
