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
#include <cstdlib>

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
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
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
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
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

// vkp_gemm_tc.cu
int vkp_tc_split_lo(vkp_ctx* ctx, cudaStream_t stream, const float* in, float* lo, size_t elems);
int vkp_tc_transpose_split(vkp_ctx* ctx, cudaStream_t stream, const float* in, uint32_t rows, uint32_t cols,
                           float* hi, float* lo, size_t ldo);
int vkp_gemm_tc_chunked(vkp_ctx* ctx, uint32_t M, uint32_t N, uint32_t K, const float* A, const float* Alo,
                        const float* Bt, const float* Btlo, float* C, vkp_tc_chunks ch, const vkp_tc_pull* pull);
int vkp_gemm_tc_supported(int transA, int transB, uint32_t M, uint32_t N, uint32_t K, const float* A,
                          const float* B, float* C, int forced);

#define VKP_MAX_BUCKET 16

struct vkp_comm_state {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  // ---- row-sharded matmul over peer memory (vkp_comm_matmul_allgather) ----
  float* barrier_word = nullptr;          // 1-element all-reduce = stream-ordered barrier
  uint32_t* flags = nullptr;              // [2 * VKP_MAX_RANKS]: range flags, then arrival counters;
                                          // K range c is valid once flags[c] == epoch
  uint32_t epoch = 0;
  // symmetric staging: two [N, K] K-major copies of B (alternating per call), exported with CUDA IPC
  void* symm = nullptr;
  size_t symm_bytes = 0;                  // bytes of ONE copy
  void* peer[VKP_MAX_RANKS] = {};         // peer[r] = rank r's symm mapped here (peer[rank] = symm)
  uint64_t calls = 0;
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
  vkp_comm_state* st = ctx->comm;
  for (int r = 0; r < st->nranks; r++)
    if (r != st->rank && st->peer[r]) cudaIpcCloseMemHandle(st->peer[r]);
  g_nccl.CommDestroy(st->comm);          // collective: every rank has stopped reading this rank's memory
  if (st->symm) cudaFree(st->symm);
  if (st->flags) cudaFree(st->flags);
  if (st->barrier_word) cudaFree(st->barrier_word);
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

// One bucket for a set of small tensors (the data-parallel gradient exchange): the
// all-reduces are grouped into a single NCCL launch and one kernel applies the 1/world scale to all
// of them, instead of a launch pair per parameter.
namespace {
struct ScaleMany {
  float* ptr[VKP_MAX_BUCKET];
  unsigned long long count[VKP_MAX_BUCKET];
  int n;
};
__global__ void __launch_bounds__(256) scale_many_kernel(ScaleMany p, float scale) {
  for (int t = 0; t < p.n; t++) {
    float* x = p.ptr[t];
    const size_t n = p.count[t];
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
      x[i] = x[i] * scale;
  }
}
}  // namespace

extern "C" int vkp_comm_allreduce_multi(vkp_ctx* ctx, float* const* bufs, const size_t* counts, int n, int op,
                                        float scale, vkp_job** job) {
  VKP_CHECK(ctx && ctx->comm, "vkp_comm_allreduce_multi: communicator not initialised");
  VKP_CHECK(bufs && counts && n >= 1 && n <= VKP_MAX_BUCKET, "vkp_comm_allreduce_multi: 1..%d tensors", VKP_MAX_BUCKET);
  VKP_CHECK(op >= 0 && op <= 3, "vkp_comm_allreduce_multi: bad op %d", op);
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  VKP_TRY(vkp_prepare_buffers(ctx, (void* const*)bufs, n));
  ScaleMany sm;
  sm.n = n;
  size_t most = 0;
  VKP_NCCL(g_nccl.GroupStart());
  for (int t = 0; t < n; t++) {
    sm.ptr[t] = bufs[t];
    sm.count[t] = counts[t];
    if (counts[t] > most) most = counts[t];
    if (counts[t]) {
      ncclResult_t r = g_nccl.AllReduce(bufs[t], bufs[t], counts[t], ncclFloat32, op, ctx->comm->comm, ctx->stream);
      if (r != 0) {
        g_nccl.GroupEnd();
        return vkp_set_error("ncclAllReduce (grouped) failed: %s", g_nccl.GetErrorString(r));
      }
    }
  }
  VKP_NCCL(g_nccl.GroupEnd());
  if (scale != 1.0f && most) {
    scale_many_kernel<<<vkp_grid_for(ctx, most, 256, 4), 256, 0, ctx->stream>>>(sm, scale);
    VKP_TRY(vkp_after_launch(ctx, "scale_many"));
  }
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

// ======================================================================================================
// Row-sharded matmul  C_r[M_r, N] = A_r[M_r, K] @ B[K, N],  B sharded by rows: rank s owns B_s[K/w, N]
// ("all-gather of B chunked by K-block and overlapped with the GEMM").
//
// No all-gather call, no gathered copy of B in its original layout, no communication kernel:
//   1. every rank transposes + TF32-splits ITS shard once into the K range it owns of a symmetric
//      [N, K] staging matrix (K-major, what the tensor-core kernel wants) that is mapped into every
//      peer with CUDA IPC; a one-word all-reduce is the barrier "all shards are staged";
//   2. ONE persistent tcgen05 GEMM (vkp_gemm_tc.cu) does the rest.  Its TMA producer starts on the
//      local K range at once and walks the others in ring order (rank+1, rank+2, ...), waiting on a
//      flag word before it enters a range; the four warps per CTA that the pre-split variant leaves
//      idle fetch exactly those ranges, in the same order, with 16-byte loads from the mapped peer
//      pointers over NVLink, store them with their lo parts and raise the flags (pull_ranges).  At
//      step j every rank reads from a different owner, so all NVSwitch ports are busy.
// The accumulators never leave TMEM between ranges, C is written once.  The staging matrix is
// double-buffered by call parity, which makes the single barrier sufficient: a rank can overwrite
// copy n%2 in call n only after it passed barrier n-1, i.e. after every peer finished GEMM n-2 and
// with it all reads of that copy.
// ======================================================================================================
namespace {

int symm_reserve(vkp_ctx* ctx, vkp_comm_state* st, size_t bytes_one) {
  if (bytes_one <= st->symm_bytes) return VKP_OK;
  // collective growth (all ranks see the same shapes): quiesce, drop the old mappings, re-export
  VKP_NCCL(g_nccl.AllReduce(st->barrier_word, st->barrier_word, 1, ncclFloat32, ncclSum, st->comm, ctx->stream));
  VKP_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int r = 0; r < st->nranks; r++) {
    if (r != st->rank && st->peer[r]) VKP_CUDA(cudaIpcCloseMemHandle(st->peer[r]));
    st->peer[r] = nullptr;
  }
  VKP_NCCL(g_nccl.AllReduce(st->barrier_word, st->barrier_word, 1, ncclFloat32, ncclSum, st->comm, ctx->stream));
  VKP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (st->symm) VKP_CUDA(cudaFree(st->symm));
  st->symm = nullptr;
  st->symm_bytes = 0;
  const size_t one = (bytes_one + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
  VKP_CUDA(cudaMalloc(&st->symm, 2 * one));
  cudaIpcMemHandle_t mine;
  VKP_CUDA(cudaIpcGetMemHandle(&mine, st->symm));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  unsigned char* dev_handles = nullptr;
  VKP_CUDA(cudaMalloc(&dev_handles, 64 * (size_t)st->nranks));
  VKP_CUDA(cudaMemcpyAsync(dev_handles + 64 * st->rank, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
  VKP_NCCL(g_nccl.AllGather(dev_handles + 64 * st->rank, dev_handles, 64, ncclInt8, st->comm, ctx->stream));
  std::vector<cudaIpcMemHandle_t> all(st->nranks);
  VKP_CUDA(cudaMemcpyAsync(all.data(), dev_handles, 64 * (size_t)st->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  VKP_CUDA(cudaStreamSynchronize(ctx->stream));
  VKP_CUDA(cudaFree(dev_handles));
  for (int r = 0; r < st->nranks; r++) {
    if (r == st->rank) { st->peer[r] = st->symm; continue; }
    cudaError_t e = cudaIpcOpenMemHandle(&st->peer[r], all[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      st->peer[r] = nullptr;
      return vkp_set_error("cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
    }
  }
  st->symm_bytes = one;
  return VKP_OK;
}

}  // namespace

extern "C" int vkp_comm_matmul_allgather(vkp_ctx* ctx, uint32_t M, uint32_t N, uint32_t K, const float* A,
                                         const float* B_shard, float* C, vkp_job** job) {
  VKP_CHECK(ctx && ctx->comm, "vkp_comm_matmul_allgather: communicator not initialised");
  VKP_CHECK(A && B_shard && C, "vkp_comm_matmul_allgather: null argument");
  vkp_comm_state* st = ctx->comm;
  const uint32_t w = (uint32_t)st->nranks, rank = (uint32_t)st->rank;
  VKP_CHECK(w <= VKP_MAX_RANKS, "vkp_comm_matmul_allgather: more than %d ranks", VKP_MAX_RANKS);
  VKP_CHECK(K % w == 0 && (K / w) % 32 == 0, "vkp_comm_matmul_allgather: K = %u must split into %u ranges of whole 32-wide k-blocks", K, w);
  VKP_CHECK(vkp_gemm_tc_supported(0, 1, M, N, K, A, B_shard, C, 1),
            "vkp_comm_matmul_allgather: shape (%u,%u,%u) is not supported by the tensor-core kernel", M, N, K);
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[3] = {(void*)A, (void*)B_shard, (void*)C};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 3));
  if (!st->flags) {
    VKP_CUDA(cudaMalloc(&st->barrier_word, 256));
    VKP_CUDA(cudaMemsetAsync(st->barrier_word, 0, 256, ctx->stream));
    VKP_CUDA(cudaMalloc(&st->flags, sizeof(uint32_t) * 2 * VKP_MAX_RANKS));
    VKP_CUDA(cudaMemsetAsync(st->flags, 0, sizeof(uint32_t) * 2 * VKP_MAX_RANKS, ctx->stream));
  }
  const size_t nk = (size_t)N * K, mk = (size_t)M * K;
  VKP_TRY(symm_reserve(ctx, st, nk * sizeof(float)));
  void* ws;
  VKP_TRY(vkp_workspace(ctx, 1, (nk + mk) * sizeof(float), &ws));
  float* bt_lo = static_cast<float*>(ws);
  float* a_lo = bt_lo + nk;
  const uint32_t kc = K / w;
  const size_t copy_off = (st->calls & 1) ? st->symm_bytes : 0;
  float* bt_hi = reinterpret_cast<float*>(static_cast<char*>(st->symm) + copy_off);
  st->calls++;
  st->epoch++;

  // 1. stage the local operands: A_lo, and B_shard^T (hi into the exported matrix, lo beside it)
  VKP_TRY(vkp_tc_split_lo(ctx, ctx->stream, A, a_lo, mk));
  VKP_TRY(vkp_tc_transpose_split(ctx, ctx->stream, B_shard, kc, N, bt_hi + (size_t)rank * kc, bt_lo + (size_t)rank * kc, K));
  // barrier: every rank's shard is staged (and every rank is done with the copy used two calls ago)
  VKP_NCCL(g_nccl.AllReduce(st->barrier_word, st->barrier_word, 1, ncclFloat32, ncclSum, st->comm, ctx->stream));

  // 2. one kernel: GEMM over all K ranges + the NVLink pulls of the ranges it does not have yet
  vkp_tc_chunks ch{st->flags, st->epoch, 0, rank, w};
  vkp_tc_pull pl;
  memset(&pl, 0, sizeof(pl));
  for (uint32_t r = 0; r < w; r++)
    pl.src[r] = reinterpret_cast<const float*>(static_cast<const char*>(st->peer[r]) + copy_off);
  pl.hi = bt_hi; pl.lo = bt_lo; pl.counters = st->flags + VKP_MAX_RANKS;
  pl.rows = N; pl.ld = K; pl.kc = kc;
  VKP_TRY(vkp_gemm_tc_chunked(ctx, M, N, K, A, a_lo, bt_hi, bt_lo, C, ch, &pl));
  return vkp_finish_op(ctx, job);
}
