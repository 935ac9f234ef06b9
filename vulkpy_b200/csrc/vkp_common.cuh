// vkp_common.cuh -- shared internals of libvulkpy_b200 (context, error handling, launch helpers)
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdlib>
#include <cstdio>
#include <cstdarg>
#include <cstring>
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX 3: no-ops unless a tool (Nsight Systems / Compute) is attached

#include "../../include/vulkpy_b200.h"

#define VKP_OK 0
#define VKP_ERR 1
#define VKP_TIMEOUT 2

int vkp_set_error(const char* fmt, ...);

#define VKP_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess)                                                          \
      return vkp_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e),  \
                           __FILE__, __LINE__);                                     \
  } while (0)

#define VKP_CHECK(cond, ...)                        \
  do {                                              \
    if (!(cond)) return vkp_set_error(__VA_ARGS__); \
  } while (0)

#define VKP_TRY(expr)          \
  do {                         \
    int _r = (expr);           \
    if (_r != VKP_OK) return _r; \
  } while (0)

// NVTX range around one public entry point (named like the reference shader / C-ABI call it serves):
// the timeline of a profiler shows `add`, `sum_axis`, `vkp_gemm`, `vkp_comm_allreduce` ... as ranges with
// their kernels inside.  The reference's analogue is util.enable_debug's API dump (vulkpy/util.py:19-55).
struct vkp_nvtx_range {
  explicit vkp_nvtx_range(const char* name) { nvtxRangePushA(name); }
  ~vkp_nvtx_range() { nvtxRangePop(); }
  vkp_nvtx_range(const vkp_nvtx_range&) = delete;
  vkp_nvtx_range& operator=(const vkp_nvtx_range&) = delete;
};
#define VKP_RANGE(name) vkp_nvtx_range _vkp_range(name)

struct vkp_block {
  void* ptr = nullptr;
  size_t bytes = 0;         // rounded size class
  uint64_t guard_seq = 0;   // work submitted up to this sequence number may still touch the block
  // Plain device memory (cudaMalloc) until the host asks for a view of the buffer; then the
  // contents move into a managed block, whose pointer is valid on host and device (vkp_host_view).
  bool managed = false;
  bool host_dirty = false;  // managed pages may live in host memory -> prefetch before the next kernel
  bool in_use = false;
  uint64_t last_seq = 0;    // sequence number of the last compute-stream operation bound to the block
  // copy-engine transfers in flight on the side streams (vkp_upload_async / vkp_download_async):
  // the next compute operation bound to the block waits for them (vkp_prepare_buffers)
  cudaEvent_t h2d_ev = nullptr;
  cudaEvent_t d2h_ev = nullptr;
};

struct vkp_comm_state;  // vkp_comm.cu

struct vkp_ctx {
  int device = 0;
  int sms = 148;
  cudaStream_t stream = nullptr;
  // copy-engine streams: host->device and device->host transfers that overlap the compute stream
  // and each other (PCIe is full duplex); created on first use
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  // cudaMalloc bounce buffers for page-locked transfers, one per stream (0 h2d, 1 d2h, 2 compute).
  // Copies that touch managed memory directly do not overlap across directions and run ~8 % slower
  // (scripts/micro/copy_duplex.cu); DMA into plain device memory + a copy kernel does neither.
  void* stage[3] = {nullptr, nullptr, nullptr};
  std::mutex mu;
  uint64_t seq = 0;        // number of operations enqueued so far
  // all operations with sequence <= done_seq are known complete; written under `mu`, read without it
  // by vkp_job_wait / vkp_job_done (fast path of Array.wait())
  std::atomic<uint64_t> done_seq{0};
  uint64_t kernel_launches = 0;
  bool debug_sync = false;
  int refcount = 1;
  std::unordered_map<void*, vkp_block*> blocks;
  std::unordered_map<size_t, std::vector<vkp_block*>> free_lists[2];   // [0] device, [1] managed
  size_t pooled_bytes = 0, live_bytes = 0;
  std::vector<cudaEvent_t> event_pool;
  // device scratch: slot 0 = partials of two-pass reductions, slot 1 = reduced values awaiting
  // re-broadcast / GEMM staging (two slots so that nested users never alias)
  void* workspace[2] = {nullptr, nullptr};
  size_t workspace_bytes[2] = {0, 0};
  vkp_comm_state* comm = nullptr;
};

enum { VKP_JOB_COMPUTE = 0, VKP_JOB_UPLOAD = 1, VKP_JOB_DOWNLOAD = 2 };

struct vkp_job {
  vkp_ctx* ctx;
  cudaEvent_t ev;
  uint64_t seq;   // compute-stream operations <= seq are complete once `ev` has fired (0 for uploads)
  int kind = VKP_JOB_COMPUTE;
};

// K ranges of a GEMM that become valid while the kernel runs (vkp_gemm_tc.cu, vkp_comm.cu)
#define VKP_MAX_RANKS 16
struct vkp_tc_chunks {
  uint32_t* flags;         // [n_chunks] device words; range c may be read once flags[c] == epoch
  uint32_t epoch;
  uint32_t kb_per_chunk;   // filled in by the launcher
  uint32_t first;          // range that is valid from the start (no flag), walked first
  uint32_t n_chunks;       // <= 1: plain GEMM
};
// Where those ranges come from: the spare warps of the same GEMM kernel copy range s of the
// K-major [rows, ld] matrix of B^T out of rank s's memory (mapped peer pointer) into the local
// hi matrix, write its TF32 low part next to it, and the last CTA to finish a range raises its flag.
struct vkp_tc_pull {
  const float* src[VKP_MAX_RANKS];   // src[s]: rank s's hi matrix as mapped here (nullptr: nothing to pull)
  float* hi;
  float* lo;
  uint32_t* counters;                // [n_chunks] arrival counters, zero between calls
  uint32_t rows, ld, kc;             // N, K, K / n_chunks
};

struct vkp_timer {
  vkp_ctx* ctx;
  cudaEvent_t ev;
};

// ---- helpers implemented in vkp_runtime.cu ----------------------------------------
int vkp_make_current(vkp_ctx* ctx);
// prefetch host-dirty managed blocks that a kernel is about to touch
int vkp_prepare_buffers(vkp_ctx* ctx, void* const* bufs, int nbuf);
// marks one enqueued operation; records the job event if requested
int vkp_finish_op(vkp_ctx* ctx, vkp_job** job);
int vkp_workspace(vkp_ctx* ctx, int slot, size_t bytes, void** out);
// call right after each <<<>>> launch
int vkp_after_launch(vkp_ctx* ctx, const char* what);

// VKP_NORMAL_PRECISE=1: Box-Muller's log / sqrt / sin / cos as <= 2 ulp float32 routines instead of the
// special-function unit (vkpm::box_muller_fast); read per call so a test can flip it inside one process
static inline bool vkp_normal_precise() {
  const char* e = getenv("VKP_NORMAL_PRECISE");
  return e && e[0] == '1';
}

static inline unsigned vkp_grid_for(vkp_ctx* ctx, size_t work_items, unsigned per_block,
                                    unsigned blocks_per_sm) {
  size_t need = (work_items + per_block - 1) / per_block;
  size_t cap = (size_t)ctx->sms * blocks_per_sm;
  if (need < 1) need = 1;
  return (unsigned)(need < cap ? need : cap);
}

// ---- family entry points (one per .cu file) ------------------------------------------
int vkp_launch_elementwise(vkp_ctx* ctx, int fam, int sub, void* const* bufs, int nbuf,
                           const void* params, size_t pbytes);
int vkp_launch_broadcast(vkp_ctx* ctx, int fam, int sub, void* const* bufs, int nbuf,
                         const void* params, size_t pbytes);
int vkp_launch_reduce(vkp_ctx* ctx, int fam, int sub, void* const* bufs, int nbuf,
                      const void* params, size_t pbytes);
int vkp_launch_gather(vkp_ctx* ctx, int fam, int sub, void* const* bufs, int nbuf,
                      const void* params, size_t pbytes);
// Element-wise step applied to every output of a GEMM after bias / accumulate (fused activation):
//   relu : C = max(C, 0)                      ReLU.forward = x.max(0.0)       (nn/layers.py:186)
//   mask : C = max(sign(mask), 0) * C         ReLU.backward on the dx GEMM    (nn/layers.py:207-210)
struct vkp_gemm_post {
  int relu;
  const float* mask;   // [M, N] like C, or nullptr
};
int vkp_launch_gemm(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K,
                    const float* A, const float* B, float* C, const float* bias, int flags,
                    vkp_gemm_post post = vkp_gemm_post{0, nullptr});

// op families
enum {
  VKF_BIN = 0,       // c = a op b            (add.comp ...)
  VKF_IBIN,          // a = a op b            (iadd.comp ...)
  VKF_SCALAR,        // b = a op s / s op a   (add_scalar.comp, rsub_scalar.comp ...)
  VKF_ISCALAR,       // a = a op s            (iadd_scalar.comp ...)
  VKF_BCAST,         // c = a op b, NumPy broadcasting (add_broadcast.comp ...)
  VKF_IBCAST,        // a = a op b, b broadcast (iadd_broadcast.comp ...)
  VKF_BCAST_COPY,    // broadcast.comp
  VKF_UNARY,         // b = f(a)              (abs.comp ...)
  VKF_IUNARY,        // a = f(a)              (iabs.comp ...)
  VKF_CLAMP,         // clamp*.comp, iclamp*.comp (sub = variant)
  VKF_REDUCE,        // sum.comp / prod / maximum / minimum (MultiVector<2>)
  VKF_REDUCE_SG,     // sum_v1.3.comp ... (Vector)
  VKF_REDUCE_AXIS,   // sum_axis.comp ...
  VKF_REDUCE_AXIS_RB,// sum_axis_rebroadcast.comp ...
  VKF_GATHER,        // gather.comp
  VKF_GATHER_AXIS,   // gather_axis.comp
  VKF_MATMUL,        // matmul.comp
  VKF_BATCH_AFFINE,  // batch_affine.comp
  VKF_CE,            // nn_cross_entropy.comp
  VKF_CE_BWD,        // nn_cross_entropy_backward.comp
  VKF_BOX_MULLER,    // prng_box_muller.comp
  VKF_IBOX_MULLER,   // prng_ibox_muller.comp
  VKF_RANDRANGE,     // prng_randrange.comp
  VKF_PRNG_U32,      // prng_xoshiro128pp_uint32.comp (bufs[0] = state, one draw per lane)
  VKF_PRNG_F32,      // prng_xoshiro128pp_float.comp
};

// binary sub-ops
enum { VKB_ADD = 0, VKB_SUB, VKB_MUL, VKB_DIV, VKB_MAX, VKB_MIN, VKB_POW, VKB_RSUB, VKB_RDIV, VKB_RPOW };
// unary sub-ops
enum {
  VKU_ABS = 0, VKU_SIGN, VKU_SIN, VKU_COS, VKU_TAN, VKU_ASIN, VKU_ACOS, VKU_ATAN, VKU_SINH,
  VKU_COSH, VKU_TANH, VKU_ASINH, VKU_ACOSH, VKU_ATANH, VKU_EXP, VKU_LOG, VKU_EXP2, VKU_LOG2,
  VKU_SQRT, VKU_INVSQRT, VKU_COUNT
};
// clamp variants: bit0 = in-place, bits1-2: 0 vvv, 1 sv (scalar min), 2 vs (scalar max), 3 ss
enum { VKC_VV = 0, VKC_SV = 1, VKC_VS = 2, VKC_SS = 3 };
// reductions
enum { VKR_SUM = 0, VKR_PROD, VKR_MAX, VKR_MIN };
