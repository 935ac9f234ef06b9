// vkp_comm.cu -- NCCL collectives on the context stream (additive: the reference is single-GPU,
// vulkpy/_vkarray.cc:485 just picks enumeratePhysicalDevices()[idx]).
//
// One process per GPU.  Used only where the sharded path has a real exchange step:
// full-array / axis-0 reduction partials (all-reduce with sum/prod/max/min), the row-sharded
// matmul's all-gather of B, and the data-parallel gradient all-reduce of vulkpy.nn.
// NCCL is resolved with dlopen at first use so that the library loads (and every other symbol
// works) on machines without NCCL; the unique id is exchanged by the caller (e.g. through the
// torch.distributed store that torchrun already provides).
#include "vkp_common.cuh"

#include <dlfcn.h>

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };
enum { ncclInt8 = 0, ncclFloat32 = 7 };

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

int load_nccl() {
  std::lock_guard<std::mutex> g(g_nccl_mu);
  if (g_nccl.lib) return VKP_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (h) break;
  }
  VKP_CHECK(h, "NCCL not found (dlopen libnccl.so.2): %s", dlerror());
#define SYM(field, name)                                         \
  *(void**)(&g_nccl.field) = dlsym(h, name);                     \
  VKP_CHECK(g_nccl.field, "NCCL symbol %s missing", name)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(AllReduce, "ncclAllReduce");
  SYM(AllGather, "ncclAllGather");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
  g_nccl.lib = h;
  return VKP_OK;
}

#define VKP_NCCL(call)                                                                       \
  do {                                                                                       \
    ncclResult_t _r = (call);                                                                \
    if (_r != 0) return vkp_set_error("%s failed: %s", #call, g_nccl.GetErrorString(_r));    \
  } while (0)

}  // namespace

struct vkp_comm_state {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
};

static_assert(VKP_COMM_ID_BYTES == sizeof(ncclUniqueId), "unique id size");

extern "C" int vkp_comm_unique_id(void* id_out) {
  VKP_CHECK(id_out, "vkp_comm_unique_id: null argument");
  VKP_TRY(load_nccl());
  ncclUniqueId id;
  VKP_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return VKP_OK;
}

extern "C" int vkp_comm_init(vkp_ctx* ctx, int nranks, int rank, const void* id) {
  VKP_CHECK(ctx && id, "vkp_comm_init: null argument");
  VKP_CHECK(nranks >= 1 && rank >= 0 && rank < nranks, "vkp_comm_init: bad rank %d of %d", rank, nranks);
  VKP_CHECK(!ctx->comm, "vkp_comm_init: communicator already initialised");
  VKP_TRY(load_nccl());
  VKP_TRY(vkp_make_current(ctx));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  vkp_comm_state* st = new vkp_comm_state();
  st->nranks = nranks;
  st->rank = rank;
  ncclResult_t r = g_nccl.CommInitRank(&st->comm, nranks, uid, rank);
  if (r != 0) {
    delete st;
    return vkp_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
  }
  ctx->comm = st;
  return VKP_OK;
}

extern "C" int vkp_comm_destroy(vkp_ctx* ctx) {
  VKP_CHECK(ctx, "vkp_comm_destroy: null context");
  if (!ctx->comm) return VKP_OK;
  VKP_TRY(vkp_make_current(ctx));
  cudaStreamSynchronize(ctx->stream);
  g_nccl.CommDestroy(ctx->comm->comm);
  delete ctx->comm;
  ctx->comm = nullptr;
  return VKP_OK;
}

extern "C" int vkp_comm_allreduce(vkp_ctx* ctx, const float* send, float* recv, size_t count, int op,
                                  vkp_job** job) {
  VKP_CHECK(ctx && ctx->comm, "vkp_comm_allreduce: communicator not initialised");
  VKP_CHECK(op >= 0 && op <= 3, "vkp_comm_allreduce: bad op %d", op);
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[2] = {(void*)send, (void*)recv};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 2));
  if (count) VKP_NCCL(g_nccl.AllReduce(send, recv, count, ncclFloat32, op, ctx->comm->comm, ctx->stream));
  return vkp_finish_op(ctx, job);
}

extern "C" int vkp_comm_allgather(vkp_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank,
                                  vkp_job** job) {
  VKP_CHECK(ctx && ctx->comm, "vkp_comm_allgather: communicator not initialised");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[2] = {(void*)send, (void*)recv};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 2));
  if (bytes_per_rank)
    VKP_NCCL(g_nccl.AllGather(send, recv, bytes_per_rank, ncclInt8, ctx->comm->comm, ctx->stream));
  return vkp_finish_op(ctx, job);
}
