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
int vkp_tc_split_lo_2d(vkp_ctx* ctx, cudaStream_t stream, const float* in, float* lo, uint32_t rows, uint32_t cols,
                       size_t ld);
int vkp_tc_transpose_split(vkp_ctx* ctx, cudaStream_t stream, const float* in, uint32_t rows, uint32_t cols,
                           float* hi, float* lo, size_t ldo);
int vkp_gemm_tc_chunked(vkp_ctx* ctx, uint32_t M, uint32_t N, uint32_t K, const float* A, const float* Alo,
                        const float* Bt, const float* Btlo, float* C, vkp_tc_chunks ch);
int vkp_gemm_tc_supported(int transA, int transB, uint32_t M, uint32_t N, uint32_t K, const float* A,
                          const float* B, float* C, int forced);

#define VKP_MAX_RANKS 64
#define VKP_MAX_BUCKET 16

struct vkp_comm_state {
  ncclComm_t comm = nullptr;
  int nranks = 1, rank = 0;
  // ---- row-sharded matmul over peer memory (vkp_comm_matmul_allgather) ----
  cudaStream_t pull_stream = nullptr;     // copy-engine pulls of the peers' shards + their lo split + flags
  // one copy engine does not fill an NVLink port with 4 KB rows: each range is pulled as `parts`
  // row bands on as many streams (part 0 on pull_stream, which then joins the others)
  static constexpr int MAX_PARTS = 8;
  int parts = 1;
  cudaStream_t part_stream[MAX_PARTS] = {};
  cudaEvent_t part_ev[MAX_PARTS] = {};
  cudaEvent_t pull_t0 = nullptr, pull_t1 = nullptr;   // timing of the last call's pull phase (diagnostics)
  cudaEvent_t ready_ev = nullptr;         // this rank's shard is staged and every rank passed the barrier
  cudaEvent_t pull_done_ev = nullptr;     // last pull of the previous call has finished
  float* barrier_word = nullptr;          // 1-element all-reduce = stream-ordered barrier
  uint32_t* flags = nullptr;              // [2 * VKP_MAX_RANKS]: flags, then the pull kernel's arrival counters;
                                          // K-range c is valid once flags[c] == epoch
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
  if (st->pull_stream) cudaStreamSynchronize(st->pull_stream);
  for (int r = 0; r < st->nranks; r++)
    if (r != st->rank && st->peer[r]) cudaIpcCloseMemHandle(st->peer[r]);
  g_nccl.CommDestroy(st->comm);          // collective: every rank has stopped reading this rank's memory
  if (st->symm) cudaFree(st->symm);
  if (st->flags) cudaFree(st->flags);
  if (st->barrier_word) cudaFree(st->barrier_word);
  if (st->ready_ev) cudaEventDestroy(st->ready_ev);
  if (st->pull_done_ev) cudaEventDestroy(st->pull_done_ev);
  for (int i = 1; i < vkp_comm_state::MAX_PARTS; i++) {
    if (st->part_stream[i]) { cudaStreamSynchronize(st->part_stream[i]); cudaStreamDestroy(st->part_stream[i]); }
    if (st->part_ev[i]) cudaEventDestroy(st->part_ev[i]);
  }
  if (st->pull_t0) cudaEventDestroy(st->pull_t0);
  if (st->pull_t1) cudaEventDestroy(st->pull_t1);
  if (st->pull_stream) cudaStreamDestroy(st->pull_stream);
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
// No all-gather call and no gathered copy of B in its original layout:
//   1. every rank transposes + TF32-splits ITS shard once into the K-range it owns of a symmetric
//      [N, K] staging matrix (K-major, what the tensor-core kernel wants) that is mapped into every
//      peer with CUDA IPC; a one-word all-reduce is the barrier "all shards are staged";
//   2. ONE persistent tcgen05 GEMM starts at once on the local K-range; its TMA producer walks the
//      other ranges in ring order (rank+1, rank+2, ...) and before entering a range waits on a flag
//      word in device memory;
//   3. meanwhile a second stream pulls the peers' ranges over NVLink with the copy engines (no SM
//      does communication), splits off their lo parts and raises the flags.  At step j every rank
//      reads from a different owner, so all NVSwitch ports are busy.
// The accumulators never leave TMEM between ranges, C is written once.  The staging matrix is
// double-buffered by call parity, which makes the single barrier sufficient: a rank can overwrite
// copy n%2 in call n only after it passed barrier n-1, i.e. after every peer finished GEMM n-2 and
// with it all reads of that copy.
// ======================================================================================================
namespace {

// All K ranges of the peers in one launch: CTAs stream each range out of the owner's exported
// staging matrix with 16-byte volatile loads (peer memory over NVLink, never cached), store it
// into the local matrix together with its TF32 low part, and the last CTA to finish a range
// raises its flag.  Ring order (rank+1, rank+2, ...): at any time every owner serves one reader.
// Co-resident with the persistent GEMM (no shared memory, 256 threads); no CTA waits on another.
struct PullArgs {
  const float* src[VKP_MAX_RANKS];   // src[s]: rank s's staging matrix (this call's copy) mapped here
  float* hi;
  float* lo;
  uint32_t* flags;
  uint32_t* counters;
  uint32_t N, K, kc, w, rank, epoch;
};

__device__ __forceinline__ uint4 ld_peer16(const float* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t tf32_lo_bits(uint32_t xb) {
  const float x = __uint_as_float(xb);
  const float h = __uint_as_float(xb & 0xffffe000u);
  return (__float_as_uint(x - h) + 0x1000u) & 0xffffe000u;
}

__global__ void __launch_bounds__(256) pull_split_kernel(const __grid_constant__ PullArgs a) {
  const uint32_t c4 = a.kc / 4;
  const size_t n4 = (size_t)a.N * c4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  constexpr int U = 8;
  for (uint32_t j = 1; j < a.w; j++) {
    uint32_t s = a.rank + j;
    if (s >= a.w) s -= a.w;
    const float* src = a.src[s] + (size_t)s * a.kc;
    float* hi = a.hi + (size_t)s * a.kc;
    float* lo = a.lo + (size_t)s * a.kc;
    for (size_t i0 = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i0 < n4; i0 += stride * U) {
      uint4 v[U];
      size_t off[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t i = i0 + u * stride;
        const size_t r = i / c4;
        off[u] = r * a.K + (i - r * c4) * 4;
        if (i < n4) v[u] = ld_peer16(src + off[u]);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (i0 + u * stride < n4) {
          *reinterpret_cast<uint4*>(hi + off[u]) = v[u];
          *reinterpret_cast<uint4*>(lo + off[u]) =
              make_uint4(tf32_lo_bits(v[u].x), tf32_lo_bits(v[u].y), tf32_lo_bits(v[u].z), tf32_lo_bits(v[u].w));
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(a.counters + s, 1u) == gridDim.x - 1) {
        a.counters[s] = 0;             // ready for the next call (launches on this stream are ordered)
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a.flags + s), "r"(a.epoch) : "memory");
      }
    }
  }
}

__global__ void set_flag_kernel(uint32_t* flag, uint32_t epoch) {
  __threadfence_system();
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
}

int symm_reserve(vkp_ctx* ctx, vkp_comm_state* st, size_t bytes_one) {
  if (bytes_one <= st->symm_bytes) return VKP_OK;
  // collective growth (all ranks see the same shapes): quiesce, drop the old mappings, re-export
  VKP_CUDA(cudaStreamSynchronize(ctx->stream));
  VKP_CUDA(cudaStreamSynchronize(st->pull_stream));
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
  if (!st->pull_stream) {
    int lo_p = 0, hi_p = 0;
    VKP_CUDA(cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p));
    VKP_CUDA(cudaStreamCreateWithPriority(&st->pull_stream, cudaStreamNonBlocking, hi_p));
    if (const char* e = getenv("VKP_PULL_PARTS")) st->parts = atoi(e);
    if (st->parts < 1) st->parts = 1;
    if (st->parts > vkp_comm_state::MAX_PARTS) st->parts = vkp_comm_state::MAX_PARTS;
    for (int i = 1; i < st->parts; i++) {
      VKP_CUDA(cudaStreamCreateWithPriority(&st->part_stream[i], cudaStreamNonBlocking, hi_p));
      VKP_CUDA(cudaEventCreateWithFlags(&st->part_ev[i], cudaEventDisableTiming));
    }
    VKP_CUDA(cudaEventCreate(&st->pull_t0));
    VKP_CUDA(cudaEventCreate(&st->pull_t1));
    VKP_CUDA(cudaEventCreateWithFlags(&st->ready_ev, cudaEventDisableTiming));
    VKP_CUDA(cudaEventCreateWithFlags(&st->pull_done_ev, cudaEventDisableTiming));
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

  // 1. stage the local operands (compute stream): A_lo, and B_shard^T (hi into the exported matrix)
  VKP_TRY(vkp_tc_split_lo(ctx, ctx->stream, A, a_lo, mk));
  VKP_TRY(vkp_tc_transpose_split(ctx, ctx->stream, B_shard, kc, N, bt_hi + (size_t)rank * kc, bt_lo + (size_t)rank * kc, K));
  // barrier: every rank's shard is staged (and every rank is done with the copy used two calls ago)
  VKP_NCCL(g_nccl.AllReduce(st->barrier_word, st->barrier_word, 1, ncclFloat32, ncclSum, st->comm, ctx->stream));
  VKP_CUDA(cudaEventRecord(st->ready_ev, ctx->stream));

  // 2. pulls over NVLink on the copy engines, ring order, then the lo split and the flag of that range
  VKP_CUDA(cudaStreamWaitEvent(st->pull_stream, st->ready_ev, 0));
  VKP_CUDA(cudaEventRecord(st->pull_t0, st->pull_stream));
  static const bool use_ce = getenv("VKP_PULL_CE") != nullptr;   // copy-engine pulls (first version; ~350 GB/s)
  if (!use_ce && w > 1) {
    PullArgs pa;
    for (uint32_t r = 0; r < w; r++)
      pa.src[r] = reinterpret_cast<const float*>(static_cast<const char*>(st->peer[r]) + copy_off);
    pa.hi = bt_hi; pa.lo = bt_lo; pa.flags = st->flags; pa.counters = st->flags + VKP_MAX_RANKS;
    pa.N = N; pa.K = K; pa.kc = kc; pa.w = w; pa.rank = rank; pa.epoch = st->epoch;
    static const int pull_ctas_per_sm = getenv("VKP_PULL_CTAS") ? atoi(getenv("VKP_PULL_CTAS")) : 1;
    pull_split_kernel<<<ctx->sms * pull_ctas_per_sm, 256, 0, st->pull_stream>>>(pa);
    VKP_TRY(vkp_after_launch(ctx, "pull_split"));
  }
  const int parts = (N >= 64u * st->parts) ? st->parts : 1;
  for (int i = 1; use_ce && i < parts; i++) VKP_CUDA(cudaStreamWaitEvent(st->part_stream[i], st->ready_ev, 0));
  for (uint32_t j = 1; use_ce && j < w; j++) {
    const uint32_t s = (rank + j) % w;
    const float* src = reinterpret_cast<const float*>(static_cast<const char*>(st->peer[s]) + copy_off) + (size_t)s * kc;
    float* dst = bt_hi + (size_t)s * kc;
    for (int i = 0; i < parts; i++) {
      const uint32_t r0 = (uint32_t)((uint64_t)N * i / parts), r1 = (uint32_t)((uint64_t)N * (i + 1) / parts);
      cudaStream_t cs = i ? st->part_stream[i] : st->pull_stream;
      VKP_CUDA(cudaMemcpy2DAsync(dst + (size_t)r0 * K, (size_t)K * 4, src + (size_t)r0 * K, (size_t)K * 4,
                                 (size_t)kc * 4, r1 - r0, cudaMemcpyDeviceToDevice, cs));
      if (i) {
        VKP_CUDA(cudaEventRecord(st->part_ev[i], cs));
        VKP_CUDA(cudaStreamWaitEvent(st->pull_stream, st->part_ev[i], 0));
      }
    }
    VKP_TRY(vkp_tc_split_lo_2d(ctx, st->pull_stream, dst, bt_lo + (size_t)s * kc, N, kc, K));
    set_flag_kernel<<<1, 1, 0, st->pull_stream>>>(st->flags + s, st->epoch);
    VKP_TRY(vkp_after_launch(ctx, "set_flag"));
  }
  VKP_CUDA(cudaEventRecord(st->pull_t1, st->pull_stream));
  VKP_CUDA(cudaEventRecord(st->pull_done_ev, st->pull_stream));

  // 3. one GEMM over all ranges, starting with the local one
  vkp_tc_chunks ch{st->flags, st->epoch, 0, rank, w};
  VKP_TRY(vkp_gemm_tc_chunked(ctx, M, N, K, A, a_lo, bt_hi, bt_lo, C, ch));
  // later work on the compute stream may reuse the workspace: order it after the pull stream too
  VKP_CUDA(cudaStreamWaitEvent(ctx->stream, st->pull_done_ev, 0));
  return vkp_finish_op(ctx, job);
}

/* diagnostics: device time of the last call's pull phase (first copy .. last flag); synchronises */
extern "C" int vkp_comm_last_pull_ms(vkp_ctx* ctx, float* ms) {
  VKP_CHECK(ctx && ctx->comm && ms, "vkp_comm_last_pull_ms: null argument");
  VKP_CHECK(ctx->comm->pull_t1, "vkp_comm_last_pull_ms: no row-sharded matmul has run");
  VKP_TRY(vkp_make_current(ctx));
  VKP_CUDA(cudaEventSynchronize(ctx->comm->pull_t1));
  VKP_CUDA(cudaEventElapsedTime(ms, ctx->comm->pull_t0, ctx->comm->pull_t1));
  return VKP_OK;
}
