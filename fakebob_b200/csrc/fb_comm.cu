// Multi-GPU plumbing: one process per GPU; the antithetic pairs of a NES draw are sharded across
// ranks and one ncclAllReduce(sum, float64) per iteration combines
//   [gradient partial sums (N) | per-sample losses (S+1, each owned by one rank) | clean scores (K)]
// on the context's stream (capturable in the per-iteration CUDA graph).
// Nothing in the reference communicates; this is new for the B200 build.
// NCCL is resolved with dlopen at fb_comm_init time so the library loads on hosts without it.
#include "fb_common.cuh"
#include "fb_nes.cuh"
#include <dlfcn.h>
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
};

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

int fb_comm_destroy_impl(fb_ctx *ctx) {
  if (ctx && ctx->comm) {
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
