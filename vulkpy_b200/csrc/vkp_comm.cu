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
  // ---- peer mailbox: small all-reduces / barriers as ONE kernel over NVLink peer memory ----
  // block layout: flags[set 2][phase 2][16 ranks] (64 words), two grid counters ... pad to 4 KiB,
  // then two data slots (call parity) of MBOX_SLOT_BYTES each
  void* mbox = nullptr;
  void* mbox_peer[VKP_MAX_RANKS] = {};    // mbox_peer[r] = rank r's mailbox mapped here
  uint32_t mbox_epoch = 0;
  int mbox_state = 0;                     // 0 untried, 1 mapped on all ranks, -1 unavailable (NCCL only)
  bool mbox_off = false;                  // vkp_comm_peer_mode(0): keep the mapping, route through NCCL (A/B runs)
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
  for (int r = 0; r < st->nranks; r++) {
    if (r != st->rank && st->peer[r]) cudaIpcCloseMemHandle(st->peer[r]);
    if (r != st->rank && st->mbox_peer[r]) cudaIpcCloseMemHandle(st->mbox_peer[r]);
  }
  g_nccl.CommDestroy(st->comm);          // collective: every rank has stopped reading this rank's memory
  if (st->symm) cudaFree(st->symm);
  if (st->mbox) cudaFree(st->mbox);
  if (st->flags) cudaFree(st->flags);
  if (st->barrier_word) cudaFree(st->barrier_word);
  delete ctx->comm;
  ctx->comm = nullptr;
  return VKP_OK;
}

extern "C" int vkp_comm_allgather(vkp_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank,
                                  vkp_job** job) {
  VKP_RANGE("vkp_comm_allgather");
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

int ensure_words(vkp_ctx* ctx, vkp_comm_state* st) {
  if (st->flags) return VKP_OK;
  VKP_CUDA(cudaMalloc(&st->barrier_word, 256));
  VKP_CUDA(cudaMemsetAsync(st->barrier_word, 0, 256, ctx->stream));
  VKP_CUDA(cudaMalloc(&st->flags, sizeof(uint32_t) * 2 * VKP_MAX_RANKS));
  VKP_CUDA(cudaMemsetAsync(st->flags, 0, sizeof(uint32_t) * 2 * VKP_MAX_RANKS, ctx->stream));
  return VKP_OK;
}

// Collective: export `local` (a whole cudaMalloc block) with CUDA IPC and map every peer's block into
// this process.  Opening a handle can fail on SOME ranks only (IPC not permitted for a device pair, a
// container without a shared PID namespace ...), and a rank that fell back alone would leave the
// others waiting for its flags: so every rank contributes 1 (all mapped) or 0 to a min all-reduce
// and all of them either keep the mappings (VKP_OK, peers[] filled, peers[rank] = local) or drop
// them and report the same "cudaIpc" error (nothing left mapped; `local` still belongs to the caller).
int ipc_map_all(vkp_ctx* ctx, vkp_comm_state* st, void* local, void** peers) {
  VKP_TRY(ensure_words(ctx, st));
  cudaIpcMemHandle_t mine;
  VKP_CUDA(cudaIpcGetMemHandle(&mine, local));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  unsigned char* dev_handles = nullptr;
  VKP_CUDA(cudaMalloc(&dev_handles, 64 * (size_t)st->nranks));
  VKP_CUDA(cudaMemcpyAsync(dev_handles + 64 * st->rank, &mine, 64, cudaMemcpyHostToDevice, ctx->stream));
  VKP_NCCL(g_nccl.AllGather(dev_handles + 64 * st->rank, dev_handles, 64, ncclInt8, st->comm, ctx->stream));
  std::vector<cudaIpcMemHandle_t> all(st->nranks);
  VKP_CUDA(cudaMemcpyAsync(all.data(), dev_handles, 64 * (size_t)st->nranks, cudaMemcpyDeviceToHost, ctx->stream));
  VKP_CUDA(cudaStreamSynchronize(ctx->stream));
  VKP_CUDA(cudaFree(dev_handles));
  float ok = 1.0f;
  std::string why;
  for (int r = 0; r < st->nranks; r++) {
    if (r == st->rank) { peers[r] = local; continue; }
    cudaError_t e = cudaIpcOpenMemHandle(&peers[r], all[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      peers[r] = nullptr;
      ok = 0.0f;
      if (why.empty()) why = std::string("rank ") + std::to_string(r) + ": " + cudaGetErrorString(e);
    }
  }
  float all_ok = 0.0f;
  VKP_CUDA(cudaMemcpyAsync(st->barrier_word + 1, &ok, sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  VKP_NCCL(g_nccl.AllReduce(st->barrier_word + 1, st->barrier_word + 1, 1, ncclFloat32, ncclMin, st->comm, ctx->stream));
  VKP_CUDA(cudaMemcpyAsync(&all_ok, st->barrier_word + 1, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  VKP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (all_ok == 1.0f) return VKP_OK;
  for (int r = 0; r < st->nranks; r++) {
    if (r != st->rank && peers[r]) cudaIpcCloseMemHandle(peers[r]);
    peers[r] = nullptr;
  }
  // every rank has closed (or never opened) its mapping of this rank's block before the caller frees it
  VKP_NCCL(g_nccl.AllReduce(st->barrier_word, st->barrier_word, 1, ncclFloat32, ncclSum, st->comm, ctx->stream));
  VKP_CUDA(cudaStreamSynchronize(ctx->stream));
  return vkp_set_error("cudaIpcOpenMemHandle failed on at least one rank (%s): peer memory disabled on all ranks",
                       why.empty() ? "a peer" : why.c_str());
}

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
  if (ipc_map_all(ctx, st, st->symm, st->peer) != VKP_OK) {
    cudaFree(st->symm);
    st->symm = nullptr;
    return VKP_ERR;
  }
  st->symm_bytes = one;
  return VKP_OK;
}

// ======================================================================================================
// Peer mailbox: all-reduce of small buffers (reduction partials: 1 .. 64 Ki floats; the data-parallel
// gradient bucket: a few MB) and barriers as ONE kernel over NVLink peer memory instead of an NCCL
// launch ("partial kernel's last block writes straight into the send buffer").
//
// Every rank owns a mailbox block mapped into all peers (CUDA IPC): two flag sets and two data slots,
// alternating by call parity.  Call n on rank q:
//   1. the kernel copies the local operands into q's slot n%2 (or a preceding reduction kernel wrote
//      its result there directly: `staged`); the last CTA to finish fences (system scope) and stores
//      n into flags[n%2][q] of EVERY rank (st.release.sys through the mapped pointers);
//   2. every CTA waits until its own flags[n%2][0..w) all equal n (ld.acquire.sys, local memory);
//   3. every thread reads its elements from all w slots (volatile 16-byte loads; w-1 of them cross
//      NVLink) and folds them in rank order 0..w-1 -- the same order on every rank, so all ranks get
//      bit-identical results -- applies the optional scale and stores to the destination.
// Slot re-use needs no extra barrier: q overwrites slot n%2 in call n+2 only after it left call n+1,
// i.e. after every peer pushed flag n+1, which each does (stream order) after finishing its reads of
// call n.  Grids are at most one CTA per SM, so all CTAs are co-resident and the flag waits cannot
// starve the copies that satisfy them.
// ======================================================================================================
constexpr size_t MBOX_HEADER_BYTES = 4096;
constexpr size_t MBOX_SLOT_BYTES = (size_t)16 << 20;   // data + (two-shot) result region

struct PeerBucket {
  const float* in[VKP_MAX_BUCKET];
  float* out[VKP_MAX_BUCKET];
  unsigned long long count[VKP_MAX_BUCKET];
  unsigned long long start[VKP_MAX_BUCKET];   // offset of tensor t in the slot, floats, multiple of 4
  int n;
};
struct PeerMbox {
  float* slot[VKP_MAX_RANKS];          // this call's data slot of every rank, as mapped here
  uint32_t* flag_out[VKP_MAX_RANKS];   // &flags[set][phase 0][my rank] inside rank r's mailbox (phase 1: + 16 words)
  uint32_t* flag_in;                   // my flags[set][phase 0][0..w)                            (phase 1: + 16 words)
  uint32_t* counter;                   // grid arrival counters (local, zero between calls), one per phase
  uint32_t epoch, w, rank;
};

// the last CTA to arrive at `counter` raises this rank's flag in every rank's mailbox
__device__ __forceinline__ void mbox_publish(uint32_t* counter, uint32_t* const* flag_out, uint32_t phase,
                                             uint32_t epoch, uint32_t w) {
  __syncthreads();
  __shared__ uint32_t s_last[2];
  if (threadIdx.x == 0) {
    __threadfence();
    s_last[phase] = (atomicAdd(counter + phase, 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (s_last[phase]) {
    if (threadIdx.x == 0) counter[phase] = 0;             // ready for the next call (stream-ordered)
    if (threadIdx.x < w) {
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag_out[threadIdx.x] + phase * VKP_MAX_RANKS), "r"(epoch)
                   : "memory");
    }
  }
}

// every CTA waits until its own flags of this phase all carry `epoch`
__device__ __forceinline__ void mbox_wait(const uint32_t* flag_in, uint32_t phase, uint32_t epoch, uint32_t w) {
  if (threadIdx.x < w) {
    // Ranks reach a collective at different times (first-use allocations, host work): wait up to a
    // minute of wall clock, then trap -- a lost flag must fail the call, not hang the GPU.
    uint32_t v;
    unsigned long long t0 = 0;
    for (uint32_t spin = 0;; spin++) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag_in + phase * VKP_MAX_RANKS + threadIdx.x) : "memory");
      if (v == epoch) break;
      __nanosleep(spin < 64 ? 20 : 200);
      if ((spin & 1023u) == 1023u) {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        if (t0 == 0) t0 = now;
        else if (now - t0 > 60000000000ull) __trap();
      }
    }
  }
  __syncthreads();
}

// local operands -> this rank's slot
__device__ __forceinline__ void mbox_stage(const PeerBucket& b, float* mine, size_t gtid, size_t gstride) {
  for (int t = 0; t < b.n; t++) {
    const float* src = b.in[t];
    float* dst = mine + b.start[t];
    const size_t n = b.count[t];
    if ((((uintptr_t)src) & 15) == 0) {
      const size_t n4 = n >> 2;
      for (size_t i = gtid; i < n4; i += gstride)
        reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
      for (size_t i = (n4 << 2) + gtid; i < n; i += gstride) dst[i] = src[i];
    } else {
      for (size_t i = gtid; i < n; i += gstride) dst[i] = src[i];
    }
  }
}

template <int OP> struct POp;
template <> struct POp<0> { static __device__ __forceinline__ float f(float a, float b) { return a + b; } };
template <> struct POp<1> { static __device__ __forceinline__ float f(float a, float b) { return a * b; } };
template <> struct POp<2> { static __device__ __forceinline__ float f(float a, float b) { return fmaxf(a, b); } };
template <> struct POp<3> { static __device__ __forceinline__ float f(float a, float b) { return fminf(a, b); } };

__device__ __forceinline__ float4 ld_vol4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_vol1(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

template <int OP>
__global__ void __launch_bounds__(256)
peer_allreduce_kernel(const __grid_constant__ PeerBucket b, const __grid_constant__ PeerMbox m, float scale, int staged) {
  const size_t gtid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t gstride = (size_t)gridDim.x * blockDim.x;
  // ---- 1. local operands -> my slot; the last CTA raises my flag in every rank's mailbox ----
  if (!staged) mbox_stage(b, m.slot[m.rank], gtid, gstride);
  mbox_publish(m.counter, m.flag_out, 0, m.epoch, m.w);
  // ---- 2. wait for every rank's flag ----
  mbox_wait(m.flag_in, 0, m.epoch, m.w);
  // ---- 3. fold the w slots in rank order ----
  for (int t = 0; t < b.n; t++) {
    float* dst = b.out[t];
    const size_t n = b.count[t];
    const size_t off = b.start[t];
    const bool vec = (((uintptr_t)dst) & 15) == 0;
    const size_t n4 = vec ? (n >> 2) : 0;
    for (size_t i = gtid; i < n4; i += gstride) {
      const size_t o = off + (i << 2);
      float4 acc = ld_vol4(m.slot[0] + o);
      for (uint32_t base = 1; base < m.w; base += 7) {
        float4 v[7];
#pragma unroll
        for (int j = 0; j < 7; j++)
          if (base + j < m.w) v[j] = ld_vol4(m.slot[base + j] + o);
#pragma unroll
        for (int j = 0; j < 7; j++)
          if (base + j < m.w) {
            acc.x = POp<OP>::f(acc.x, v[j].x); acc.y = POp<OP>::f(acc.y, v[j].y);
            acc.z = POp<OP>::f(acc.z, v[j].z); acc.w = POp<OP>::f(acc.w, v[j].w);
          }
      }
      if (scale != 1.0f) { acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale; }
      reinterpret_cast<float4*>(dst)[i] = acc;
    }
    for (size_t i = (n4 << 2) + gtid; i < n; i += gstride) {
      float acc = ld_vol1(m.slot[0] + off + i);
      for (uint32_t r = 1; r < m.w; r++) acc = POp<OP>::f(acc, ld_vol1(m.slot[r] + off + i));
      if (scale != 1.0f) acc *= scale;
      dst[i] = acc;
    }
  }
}

// Two-shot form for buckets where NVLink traffic matters (the gradient bucket: a few MB on 4-8 ranks):
// after the operands are staged (phase 0), rank q folds ONLY slice q of the concatenated bucket from
// all w slots (reduce-scatter: (w-1)/w of the bucket crosses NVLink per rank instead of w-1 buckets),
// writes it to its outputs and to a result region behind its data slot, raises the phase-1 flags, and
// pulls the other w-1 reduced slices from their owners (all-gather).  Each slice is reduced once, by
// its owner, in rank order: results are bit-identical on all ranks.
__device__ __forceinline__ void mbox_store4(const PeerBucket& b, size_t o, float4 v) {
  for (int t = 0; t < b.n; t++) {
    const size_t st = b.start[t];
    const size_t n = b.count[t];
    if (o >= st && o < st + n) {
      float* dst = b.out[t] + (o - st);
      if (o - st + 4 <= n && (((uintptr_t)dst) & 15) == 0) {
        *reinterpret_cast<float4*>(dst) = v;
      } else {
        const float e[4] = {v.x, v.y, v.z, v.w};
        for (int k = 0; k < 4 && o - st + k < n; k++) dst[k] = e[k];
      }
      return;
    }
  }
}

constexpr int PEER2_THREADS = 512;

__device__ __forceinline__ void st_vec4(float* p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Phase 1 of the two-shot form: rank q folds slice q from all w slots (every thread has its 7 remote loads
// in flight at once: one NVLink round trip per pass), stores it to its own outputs and PUSHES it into the
// result region (behind the data, `total` floats further) of EVERY rank's slot -- posted writes, no round
// trip -- then raises the phase-1 flags.  Phase 2 is local: copy the other ranks' slices from this rank's own
// result region to the outputs.  (The first version pulled the reduced slices with one 16-byte load in
// flight per thread: 6 NVLink round trips for the 4.3 MB gradient bucket, 42 us back to back on 8 GPUs.)
template <int OP>
__global__ void __launch_bounds__(PEER2_THREADS)
peer_allreduce2_kernel(const __grid_constant__ PeerBucket b, const __grid_constant__ PeerMbox m, float scale,
                       unsigned long long total, unsigned long long slice) {
  const size_t gtid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t gstride = (size_t)gridDim.x * blockDim.x;
  mbox_stage(b, m.slot[m.rank], gtid, gstride);
  mbox_publish(m.counter, m.flag_out, 0, m.epoch, m.w);
  mbox_wait(m.flag_in, 0, m.epoch, m.w);
  const size_t my_lo = (size_t)m.rank * slice;
  const size_t my_hi = my_lo + slice < total ? my_lo + slice : total;
  // ---- reduce-scatter my slice, push the result to everyone ----
  {
    const size_t n4 = my_hi > my_lo ? (my_hi - my_lo) >> 2 : 0;
    for (size_t i = gtid; i < n4; i += gstride) {
      const size_t o = my_lo + (i << 2);
      float4 acc = ld_vol4(m.slot[0] + o);
      for (uint32_t base = 1; base < m.w; base += 7) {
        float4 v[7];
#pragma unroll
        for (int j = 0; j < 7; j++)
          if (base + j < m.w) v[j] = ld_vol4(m.slot[base + j] + o);
#pragma unroll
        for (int j = 0; j < 7; j++)
          if (base + j < m.w) {
            acc.x = POp<OP>::f(acc.x, v[j].x); acc.y = POp<OP>::f(acc.y, v[j].y);
            acc.z = POp<OP>::f(acc.z, v[j].z); acc.w = POp<OP>::f(acc.w, v[j].w);
          }
      }
      if (scale != 1.0f) { acc.x *= scale; acc.y *= scale; acc.z *= scale; acc.w *= scale; }
      for (uint32_t r = 0; r < m.w; r++)
        if (r != m.rank) st_vec4(m.slot[r] + total + o, acc);
      mbox_store4(b, o, acc);
    }
  }
  __threadfence_system();                    // my pushes are performed before this CTA counts as arrived
  mbox_publish(m.counter, m.flag_out, 1, m.epoch, m.w);
  mbox_wait(m.flag_in, 1, m.epoch, m.w);
  // ---- the other ranks' reduced slices now sit in MY result region: local copy to the outputs ----
  {
    const float* res = m.slot[m.rank] + total;
    const size_t n4 = (size_t)total >> 2;
    for (size_t i = gtid; i < n4; i += gstride) {
      const size_t o = i << 2;
      if (o >= my_lo && o < my_hi) continue;
      mbox_store4(b, o, ld_vol4(res + o));
    }
  }
}

// collective, first use: allocate + map the mailboxes.  Returns with st->mbox_state = 1 or -1.
int mbox_reserve(vkp_ctx* ctx, vkp_comm_state* st) {
  if (st->mbox_state != 0) return VKP_OK;
  static const bool off = getenv("VKP_COMM_PEER") && getenv("VKP_COMM_PEER")[0] == '0';
  if (off || st->nranks < 2 || st->nranks > VKP_MAX_RANKS) { st->mbox_state = -1; return VKP_OK; }
  const size_t bytes = MBOX_HEADER_BYTES + 2 * MBOX_SLOT_BYTES;
  VKP_CUDA(cudaMalloc(&st->mbox, bytes));
  VKP_CUDA(cudaMemsetAsync(st->mbox, 0, MBOX_HEADER_BYTES, ctx->stream));
  VKP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ipc_map_all(ctx, st, st->mbox, st->mbox_peer) != VKP_OK) {   // the same verdict on every rank
    cudaFree(st->mbox);
    st->mbox = nullptr;
    st->mbox_state = -1;
    return VKP_OK;
  }
  st->mbox_state = 1;
  return VKP_OK;
}

// launches the mailbox kernel for the next epoch; `staged`: slot already holds the local operands
int peer_launch(vkp_ctx* ctx, vkp_comm_state* st, int op, PeerBucket& b, size_t total_floats, float scale, int staged) {
  st->mbox_epoch++;
  const uint32_t set = st->mbox_epoch & 1u;
  PeerMbox m;
  memset(&m, 0, sizeof(m));
  for (int r = 0; r < st->nranks; r++) {
    char* base = static_cast<char*>(st->mbox_peer[r]);
    m.slot[r] = reinterpret_cast<float*>(base + MBOX_HEADER_BYTES + set * MBOX_SLOT_BYTES);
    m.flag_out[r] = reinterpret_cast<uint32_t*>(base) + set * 2 * VKP_MAX_RANKS + st->rank;
  }
  m.flag_in = reinterpret_cast<uint32_t*>(st->mbox) + set * 2 * VKP_MAX_RANKS;
  m.counter = reinterpret_cast<uint32_t*>(st->mbox) + 4 * VKP_MAX_RANKS;
  m.epoch = st->mbox_epoch;
  m.w = (uint32_t)st->nranks;
  m.rank = (uint32_t)st->rank;
  // 16 bytes per thread per pass; at most one CTA per SM (co-residency, see above)
  size_t ctas = (total_floats / 4 + 255) / 256;
  if (ctas < 1) ctas = 1;
  if (ctas > (size_t)ctx->sms) ctas = ctx->sms;
  const unsigned grid = (unsigned)ctas;
  // two-shot (reduce-scatter + all-gather) once the bucket is big enough for NVLink bytes to matter
  static const size_t two_shot_min = getenv("VKP_COMM_TWO_SHOT_MIN") ? (size_t)atoll(getenv("VKP_COMM_TWO_SHOT_MIN")) : 32768;
  size_t slice = ((total_floats + st->nranks - 1) / st->nranks + 3) & ~(size_t)3;
  const int two_shot_ranks = getenv("VKP_COMM_TWO_SHOT_MIN") ? 2 : 3;   // 2 ranks: same NVLink bytes either way
  const bool two = !staged && st->nranks >= two_shot_ranks && total_floats >= two_shot_min &&
                   2 * total_floats * sizeof(float) <= MBOX_SLOT_BYTES;
  if (two) {
    // staging and the final copy walk the whole bucket, the reduce-scatter one slice: one CTA per SM
    size_t ctas2 = (total_floats / 4 + PEER2_THREADS - 1) / PEER2_THREADS;
    if (ctas2 < 1) ctas2 = 1;
    if (ctas2 > (size_t)ctx->sms) ctas2 = ctx->sms;
    const unsigned grid2 = (unsigned)ctas2;
    switch (op) {
      case 0: peer_allreduce2_kernel<0><<<grid2, PEER2_THREADS, 0, ctx->stream>>>(b, m, scale, total_floats, slice); break;
      case 1: peer_allreduce2_kernel<1><<<grid2, PEER2_THREADS, 0, ctx->stream>>>(b, m, scale, total_floats, slice); break;
      case 2: peer_allreduce2_kernel<2><<<grid2, PEER2_THREADS, 0, ctx->stream>>>(b, m, scale, total_floats, slice); break;
      default: peer_allreduce2_kernel<3><<<grid2, PEER2_THREADS, 0, ctx->stream>>>(b, m, scale, total_floats, slice); break;
    }
    return vkp_after_launch(ctx, "peer_allreduce(two-shot)");
  }
  switch (op) {
    case 0: peer_allreduce_kernel<0><<<grid, 256, 0, ctx->stream>>>(b, m, scale, staged); break;
    case 1: peer_allreduce_kernel<1><<<grid, 256, 0, ctx->stream>>>(b, m, scale, staged); break;
    case 2: peer_allreduce_kernel<2><<<grid, 256, 0, ctx->stream>>>(b, m, scale, staged); break;
    default: peer_allreduce_kernel<3><<<grid, 256, 0, ctx->stream>>>(b, m, scale, staged); break;
  }
  return vkp_after_launch(ctx, "peer_allreduce");
}

float* mbox_next_slot(vkp_comm_state* st) {   // the slot the NEXT peer_launch will use
  const uint32_t set = (st->mbox_epoch + 1) & 1u;
  return reinterpret_cast<float*>(static_cast<char*>(st->mbox) + MBOX_HEADER_BYTES + set * MBOX_SLOT_BYTES);
}

}  // namespace

// vkp_reduce.cu: local [prev, axis, post] -> [prev, post] reduction into any device pointer
int vkp_reduce_axis_into(vkp_ctx* ctx, int op, const float* in, float* out, uint32_t prev, uint32_t axis, uint32_t post);

static bool peer_ready(vkp_ctx* ctx, vkp_comm_state* st, size_t floats) {
  if (st->mbox_off) return false;
  if (st->mbox_state == 0 && mbox_reserve(ctx, st) != VKP_OK) return false;
  return st->mbox_state == 1 && (floats + 4 * VKP_MAX_BUCKET) * sizeof(float) <= MBOX_SLOT_BYTES;
}

extern "C" int vkp_comm_allreduce(vkp_ctx* ctx, const float* send, float* recv, size_t count, int op,
                                  vkp_job** job) {
  VKP_RANGE("vkp_comm_allreduce");
  VKP_CHECK(ctx && ctx->comm, "vkp_comm_allreduce: communicator not initialised");
  VKP_CHECK(op >= 0 && op <= 3, "vkp_comm_allreduce: bad op %d", op);
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[2] = {(void*)send, (void*)recv};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 2));
  vkp_comm_state* st = ctx->comm;
  if (count && peer_ready(ctx, st, count)) {       // one kernel over NVLink peer memory
    PeerBucket b;
    memset(&b, 0, sizeof(b));
    b.in[0] = send; b.out[0] = recv; b.count[0] = count; b.start[0] = 0; b.n = 1;
    VKP_TRY(peer_launch(ctx, st, op, b, (count + 3) & ~(size_t)3, 1.0f, 0));
  } else if (count) {
    VKP_NCCL(g_nccl.AllReduce(send, recv, count, ncclFloat32, op, st->comm, ctx->stream));
  }
  return vkp_finish_op(ctx, job);
}

// One bucket for a set of small tensors (the data-parallel gradient exchange): ONE
// mailbox kernel reduces all of them and applies the 1/world scale; buckets that do not fit a
// mailbox slot (or without peer memory) go through one grouped NCCL launch + one scale kernel.
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
  VKP_RANGE("vkp_comm_allreduce_multi");
  VKP_CHECK(ctx && ctx->comm, "vkp_comm_allreduce_multi: communicator not initialised");
  VKP_CHECK(bufs && counts && n >= 1 && n <= VKP_MAX_BUCKET, "vkp_comm_allreduce_multi: 1..%d tensors", VKP_MAX_BUCKET);
  VKP_CHECK(op >= 0 && op <= 3, "vkp_comm_allreduce_multi: bad op %d", op);
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  VKP_TRY(vkp_prepare_buffers(ctx, (void* const*)bufs, n));
  vkp_comm_state* st = ctx->comm;
  size_t total = 0, most = 0;
  for (int t = 0; t < n; t++) {
    total += (counts[t] + 3) & ~(size_t)3;
    if (counts[t] > most) most = counts[t];
  }
  if (total && peer_ready(ctx, st, total)) {
    PeerBucket b;
    memset(&b, 0, sizeof(b));
    size_t off = 0;
    for (int t = 0; t < n; t++) {
      b.in[t] = bufs[t]; b.out[t] = bufs[t]; b.count[t] = counts[t]; b.start[t] = off;
      off += (counts[t] + 3) & ~(size_t)3;
    }
    b.n = n;
    VKP_TRY(peer_launch(ctx, st, op, b, total, scale, 0));
    return vkp_finish_op(ctx, job);
  }
  ScaleMany sm;
  sm.n = n;
  VKP_NCCL(g_nccl.GroupStart());
  for (int t = 0; t < n; t++) {
    sm.ptr[t] = bufs[t];
    sm.count[t] = counts[t];
    if (counts[t]) {
      ncclResult_t r = g_nccl.AllReduce(bufs[t], bufs[t], counts[t], ncclFloat32, op, st->comm, ctx->stream);
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

// Sharded reduction with its exchange step fused (SURVEY 8(e) rows "full reduction" / "axis = 0"):
// the local [prev, axis, post] -> [prev, post] reduction writes its result straight into this rank's
// mailbox slot (no intermediate array, no copy), and the mailbox kernel folds the w slots into `out`.
// Full reductions pass prev = 1, axis = n, post = 1.  Results that do not fit a slot (or ranks
// without peer memory) reduce into `out` and all-reduce it in place with NCCL.
extern "C" int vkp_comm_reduce_allreduce(vkp_ctx* ctx, int op, const float* in, float* out, uint32_t prev,
                                         uint32_t axis, uint32_t post, vkp_job** job) {
  VKP_RANGE("vkp_comm_reduce_allreduce");
  VKP_CHECK(ctx && ctx->comm, "vkp_comm_reduce_allreduce: communicator not initialised");
  VKP_CHECK(in && out && op >= 0 && op <= 3, "vkp_comm_reduce_allreduce: bad argument");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[2] = {(void*)in, (void*)out};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 2));
  vkp_comm_state* st = ctx->comm;
  const size_t nout = (size_t)prev * post;
  if (nout == 0) return vkp_finish_op(ctx, job);
  VKP_CHECK(axis > 0, "vkp_comm_reduce_allreduce: empty axis");
  if (peer_ready(ctx, st, nout)) {
    VKP_TRY(vkp_reduce_axis_into(ctx, op, in, mbox_next_slot(st), prev, axis, post));
    PeerBucket b;
    memset(&b, 0, sizeof(b));
    b.in[0] = nullptr; b.out[0] = out; b.count[0] = nout; b.start[0] = 0; b.n = 1;
    VKP_TRY(peer_launch(ctx, st, op, b, nout, 1.0f, 1));
  } else {
    VKP_TRY(vkp_reduce_axis_into(ctx, op, in, out, prev, axis, post));
    VKP_NCCL(g_nccl.AllReduce(out, out, nout, ncclFloat32, op, st->comm, ctx->stream));
  }
  return vkp_finish_op(ctx, job);
}

// Stream-ordered barrier across the ranks: mailbox flags when peer memory is mapped (one tiny
// kernel), else a one-word NCCL all-reduce.
static int comm_barrier_locked(vkp_ctx* ctx, vkp_comm_state* st) {
  if (peer_ready(ctx, st, 0)) {
    PeerBucket b;
    memset(&b, 0, sizeof(b));
    return peer_launch(ctx, st, 0, b, 0, 1.0f, 1);
  }
  VKP_TRY(ensure_words(ctx, st));
  VKP_NCCL(g_nccl.AllReduce(st->barrier_word, st->barrier_word, 1, ncclFloat32, ncclSum, st->comm, ctx->stream));
  return VKP_OK;
}

// mode 0: every later collective goes through NCCL; 1: peer mailbox where it applies (default);
// -1: query only.  *active = 1 when the mailbox is mapped and selected.  Collective: all ranks must
// switch at the same point of their streams.
extern "C" int vkp_comm_peer_mode(vkp_ctx* ctx, int mode, int* active) {
  VKP_CHECK(ctx && ctx->comm, "vkp_comm_peer_mode: communicator not initialised");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  vkp_comm_state* st = ctx->comm;
  if (mode == 0) st->mbox_off = true;
  if (mode == 1) st->mbox_off = false;
  if (active) *active = peer_ready(ctx, st, 0) ? 1 : 0;
  return VKP_OK;
}

extern "C" int vkp_comm_barrier(vkp_ctx* ctx, vkp_job** job) {
  VKP_RANGE("vkp_comm_barrier");
  VKP_CHECK(ctx && ctx->comm, "vkp_comm_barrier: communicator not initialised");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  VKP_TRY(comm_barrier_locked(ctx, ctx->comm));
  return vkp_finish_op(ctx, job);
}


extern "C" int vkp_comm_matmul_allgather(vkp_ctx* ctx, uint32_t M, uint32_t N, uint32_t K, const float* A,
                                         const float* B_shard, float* C, vkp_job** job) {
  VKP_RANGE("vkp_comm_matmul_allgather");
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
  VKP_TRY(ensure_words(ctx, st));
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
  VKP_TRY(comm_barrier_locked(ctx, st));

  // 2. one kernel: GEMM over all K ranges + the NVLink pulls of the ranges it does not have yet
  vkp_tc_chunks ch{st->flags, st->epoch, 0, rank, w};
  vkp_tc_pull pl;
  memset(&pl, 0, sizeof(pl));
  for (uint32_t r = 0; r < w; r++)
    pl.src[r] = reinterpret_cast<const float*>(static_cast<const char*>(st->peer[r]) + copy_off);
  pl.hi = bt_hi; pl.lo = bt_lo; pl.counters = st->flags + VKP_MAX_RANKS;
  pl.rows = N; pl.ld = K; pl.kc = kc;
  // VKP_COMM_NO_PULL=1 (timing diagnostic only, WRONG results): the same GEMM over the staging matrix as it is,
  // no K ranges, no pulls -- what the kernel costs when nothing has to cross NVLink
  static const bool no_pull = getenv("VKP_COMM_NO_PULL") != nullptr;
  if (no_pull) {
    VKP_TRY(vkp_gemm_tc_chunked(ctx, M, N, K, A, a_lo, bt_hi, bt_lo, C, vkp_tc_chunks{nullptr, 0, 0, 0, 1}, nullptr));
    return vkp_finish_op(ctx, job);
  }
  VKP_TRY(vkp_gemm_tc_chunked(ctx, M, N, K, A, a_lo, bt_hi, bt_lo, C, ch, &pl));
  return vkp_finish_op(ctx, job);
}
