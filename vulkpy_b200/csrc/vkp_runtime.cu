// vkp_runtime.cu -- context, stream-ordered managed-memory pool, jobs (pooled events), timers.
//
// Replaces class GPU / Buffer<T> / Job of the reference (vulkpy/_vkarray.cc:38-130,392-574).
// Design differences (B200-first):
//   * one in-order, non-blocking CUDA stream per device instead of a command pool + fence per op
//     and a host-side wait on every dependency (_vkarray.cc:412-438);
//   * buffers come from a size-bucketed pool of cudaMallocManaged blocks kept resident in HBM;
//     the host sees the same pointer (reference: HOST_VISIBLE|HOST_COHERENT mapped buffers,
//     _vkarray.cc:61-72) and pages migrate only when the host really touches them;
//   * a Job is a pooled cudaEvent recorded after the op.
#include "vkp_common.cuh"

#include <chrono>
#include <thread>

static thread_local char g_err[1024] = "";

int vkp_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return VKP_ERR;
}

static std::mutex g_ctx_mu;
static vkp_ctx* g_ctx_by_device[64] = {nullptr};

extern "C" int vkp_abi_version(void) { return VKP_ABI_VERSION; }
extern "C" const char* vkp_last_error(void) { return g_err; }

extern "C" int vkp_device_count(int* count) {
  VKP_CHECK(count != nullptr, "vkp_device_count: null argument");
  VKP_CUDA(cudaGetDeviceCount(count));
  return VKP_OK;
}

int vkp_make_current(vkp_ctx* ctx) {
  VKP_CUDA(cudaSetDevice(ctx->device));
  return VKP_OK;
}

extern "C" int vkp_ctx_create(int device, float priority, vkp_ctx** out) {
  VKP_CHECK(out != nullptr, "vkp_ctx_create: null argument");
  std::lock_guard<std::mutex> g(g_ctx_mu);
  int n = 0;
  VKP_CUDA(cudaGetDeviceCount(&n));
  VKP_CHECK(n > 0, "no CUDA device visible: the vulkpy B200 backend has no CPU fallback");
  VKP_CHECK(device >= 0 && device < n && device < 64, "GPU index %d out of range (%d devices)", device, n);
  if (g_ctx_by_device[device]) {  // GPU(idx) objects with the same index share one context
    g_ctx_by_device[device]->refcount++;
    *out = g_ctx_by_device[device];
    return VKP_OK;
  }
  VKP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  VKP_CUDA(cudaGetDeviceProperties(&prop, device));
  VKP_CHECK(prop.major >= 10,
            "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
            prop.major, prop.minor);
  int concurrent = 0;
  VKP_CUDA(cudaDeviceGetAttribute(&concurrent, cudaDevAttrConcurrentManagedAccess, device));
  VKP_CHECK(concurrent, "device %d lacks concurrent managed access (needed for the host view)", device);
  vkp_ctx* ctx = new vkp_ctx();
  ctx->device = device;
  ctx->sms = prop.multiProcessorCount;
  int lo = 0, hi = 0;
  VKP_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // lo = least priority (numerically largest)
  float p = priority < 0.f ? 0.f : (priority > 1.f ? 1.f : priority);
  int prio = lo + (int)((hi - lo) * p);
  VKP_CUDA(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio));
  g_ctx_by_device[device] = ctx;
  *out = ctx;
  return VKP_OK;
}

static int sync_locked(vkp_ctx* ctx) {
  uint64_t s = ctx->seq;
  VKP_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->h2d_stream) VKP_CUDA(cudaStreamSynchronize(ctx->h2d_stream));
  if (ctx->d2h_stream) VKP_CUDA(cudaStreamSynchronize(ctx->d2h_stream));
  if (s > ctx->done_seq) ctx->done_seq = s;
  return VKP_OK;
}

extern "C" int vkp_ctx_sync(vkp_ctx* ctx) {
  VKP_CHECK(ctx, "vkp_ctx_sync: null context");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  return sync_locked(ctx);
}

static int trim_locked(vkp_ctx* ctx) {
  VKP_TRY(sync_locked(ctx));
  for (int k = 0; k < 2; k++)
  for (auto& kv : ctx->free_lists[k]) {
    for (vkp_block* b : kv.second) {
      if (b->h2d_ev) ctx->event_pool.push_back(b->h2d_ev);   // all streams are idle after the sync
      if (b->d2h_ev) ctx->event_pool.push_back(b->d2h_ev);
      ctx->blocks.erase(b->ptr);
      cudaFree(b->ptr);
      ctx->pooled_bytes -= b->bytes;
      delete b;
    }
    kv.second.clear();
  }
  return VKP_OK;
}

extern "C" int vkp_ctx_trim(vkp_ctx* ctx) {
  VKP_CHECK(ctx, "vkp_ctx_trim: null context");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  return trim_locked(ctx);
}

extern "C" int vkp_ctx_destroy(vkp_ctx* ctx) {
  if (!ctx) return VKP_OK;
  {
    std::lock_guard<std::mutex> g(g_ctx_mu);
    if (--ctx->refcount > 0) return VKP_OK;
    // The context of a device stays cached for the life of the process: arrays that are
    // still alive (NumPy views keep their buffer) may outlive every GPU object.
    ctx->refcount = 0;
  }
  return VKP_OK;
}

extern "C" int vkp_ctx_device(vkp_ctx* ctx, int* device) {
  VKP_CHECK(ctx && device, "vkp_ctx_device: null argument");
  *device = ctx->device;
  return VKP_OK;
}

extern "C" int vkp_ctx_sm_count(vkp_ctx* ctx, int* sms) {
  VKP_CHECK(ctx && sms, "vkp_ctx_sm_count: null argument");
  *sms = ctx->sms;
  return VKP_OK;
}

extern "C" int vkp_ctx_set_debug_sync(vkp_ctx* ctx, int enable) {
  VKP_CHECK(ctx, "vkp_ctx_set_debug_sync: null context");
  ctx->debug_sync = enable != 0;
  return VKP_OK;
}

extern "C" int vkp_ctx_launch_count(vkp_ctx* ctx, uint64_t* kernels) {
  VKP_CHECK(ctx && kernels, "vkp_ctx_launch_count: null argument");
  *kernels = ctx->kernel_launches;
  return VKP_OK;
}

extern "C" int vkp_ctx_mem_info(vkp_ctx* ctx, size_t* pooled, size_t* live) {
  VKP_CHECK(ctx, "vkp_ctx_mem_info: null context");
  if (pooled) *pooled = ctx->pooled_bytes;
  if (live) *live = ctx->live_bytes;
  return VKP_OK;
}

// ---- allocator ------------------------------------------------------------------------
static size_t size_class(size_t bytes) {
  if (bytes < 512) return 512;
  if (bytes <= (1u << 20)) {  // next power of two up to 1 MiB
    size_t s = 512;
    while (s < bytes) s <<= 1;
    return s;
  }
  const size_t two_mib = 2u << 20;  // 2 MiB pages: keep large blocks page-granular
  return (bytes + two_mib - 1) / two_mib * two_mib;
}

// quiet != 0: prefer a cached block that no enqueued compute work can still touch, so that a
// copy-engine upload into it has nothing to wait for (the default is LIFO: hottest block first).
// managed != 0: host-visible block (cudaMallocManaged, prefetched into HBM); otherwise cudaMalloc.
static int alloc_locked(vkp_ctx* ctx, size_t bytes, void** ptr, int quiet, int managed) {
  const size_t cls = size_class(bytes);
  auto& lists = ctx->free_lists[managed ? 1 : 0];
  auto it = lists.find(cls);
  if (it != lists.end() && !it->second.empty()) {
    std::vector<vkp_block*>& fl = it->second;
    size_t pick = fl.size() - 1;
    if (quiet) {
      bool found = false;
      for (size_t i = 0; i < fl.size(); i++)
        if (fl[i]->guard_seq <= ctx->done_seq && !fl[i]->d2h_ev) { pick = i; found = true; break; }
      // Every cached block is still in the compute stream's future: an upload into one of them would wait for
      // ALL compute work enqueued so far (bench e2e: the next step's 2 GiB upload started only after this step's
      // kernels, 49 ms per step against 42 ms for the bare copies).  Memory is what a B200 has plenty of: take a
      // fresh block instead, as long as the pool stays below half of the device.
      if (!found) {
        static size_t total[64] = {};
        size_t& tot = total[ctx->device & 63];
        if (!tot) {
          size_t fr = 0;
          if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); tot = 1; }
        }
        if (ctx->pooled_bytes + cls <= tot / 2) goto fresh;
      }
    }
    vkp_block* b = fl[pick];
    fl.erase(fl.begin() + pick);
    b->in_use = true;
    ctx->live_bytes += b->bytes;
    *ptr = b->ptr;
    return VKP_OK;
  }
fresh:
  void* p = nullptr;
  auto raw_alloc = [&]() { return managed ? cudaMallocManaged(&p, cls, cudaMemAttachGlobal) : cudaMalloc(&p, cls); };
  cudaError_t e = raw_alloc();
  if (e != cudaSuccess) {
    cudaGetLastError();
    // release cached blocks and retry once
    VKP_TRY(trim_locked(ctx));
    e = raw_alloc();
    if (e != cudaSuccess) {
      cudaGetLastError();
      return vkp_set_error("%s(%zu bytes) failed: %s", managed ? "cudaMallocManaged" : "cudaMalloc", cls,
                           cudaGetErrorString(e));
    }
  }
  // populate managed pages in HBM now; kernels then run at full rate without GPU page faults
  if (managed) VKP_CUDA(cudaMemPrefetchAsync(p, cls, ctx->device, ctx->stream));
  vkp_block* b = new vkp_block();
  b->ptr = p;
  b->bytes = cls;
  b->managed = managed != 0;
  b->in_use = true;
  ctx->blocks[p] = b;
  ctx->pooled_bytes += cls;
  ctx->live_bytes += cls;
  *ptr = p;
  return VKP_OK;
}

static int alloc_impl(vkp_ctx* ctx, size_t bytes, void** ptr, int quiet) {
  VKP_CHECK(ctx && ptr, "vkp_alloc: null argument");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  return alloc_locked(ctx, bytes, ptr, quiet, 0);
}

extern "C" int vkp_alloc(vkp_ctx* ctx, size_t bytes, void** ptr) { return alloc_impl(ctx, bytes, ptr, 0); }
extern "C" int vkp_alloc_for_upload(vkp_ctx* ctx, size_t bytes, void** ptr) { return alloc_impl(ctx, bytes, ptr, 1); }

extern "C" int vkp_free(vkp_ctx* ctx, void* ptr) {
  if (!ptr) return VKP_OK;
  VKP_CHECK(ctx, "vkp_free: null context");
  std::lock_guard<std::mutex> g(ctx->mu);
  auto it = ctx->blocks.find(ptr);
  VKP_CHECK(it != ctx->blocks.end() && it->second->in_use, "vkp_free: %p is not a live buffer", ptr);
  vkp_block* b = it->second;
  b->in_use = false;
  // the last operation bound to the block may still read or write it (every entry point passes its
  // buffers through vkp_prepare_buffers, and buffers are always bound by their base pointer)
  b->guard_seq = b->last_seq;
  ctx->live_bytes -= b->bytes;
  ctx->free_lists[b->managed ? 1 : 0][b->bytes].push_back(b);
  return VKP_OK;
}

int vkp_prepare_buffers(vkp_ctx* ctx, void* const* bufs, int nbuf) {
  for (int i = 0; i < nbuf; i++) {
    if (!bufs[i]) continue;
    auto it = ctx->blocks.find(bufs[i]);
    if (it == ctx->blocks.end()) continue;
    vkp_block* b = it->second;
    if (b->host_dirty) {   // only managed blocks are ever host-dirty
      VKP_CUDA(cudaMemPrefetchAsync(b->ptr, b->bytes, ctx->device, ctx->stream));
      b->host_dirty = false;
    }
    // copy-engine transfers on the side streams: this and every later compute operation run after them
    if (b->h2d_ev) {
      VKP_CUDA(cudaStreamWaitEvent(ctx->stream, b->h2d_ev, 0));
      ctx->event_pool.push_back(b->h2d_ev);
      b->h2d_ev = nullptr;
    }
    if (b->d2h_ev) {
      VKP_CUDA(cudaStreamWaitEvent(ctx->stream, b->d2h_ev, 0));
      ctx->event_pool.push_back(b->d2h_ev);
      b->d2h_ev = nullptr;
    }
    b->last_seq = ctx->seq + 1;
  }
  return VKP_OK;
}

static int take_event(vkp_ctx* ctx, cudaEvent_t* ev) {
  if (!ctx->event_pool.empty()) {
    *ev = ctx->event_pool.back();
    ctx->event_pool.pop_back();
    return VKP_OK;
  }
  VKP_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
  return VKP_OK;
}

int vkp_finish_op(vkp_ctx* ctx, vkp_job** job) {
  ctx->seq++;
  if (job) {
    cudaEvent_t ev;
    VKP_TRY(take_event(ctx, &ev));
    VKP_CUDA(cudaEventRecord(ev, ctx->stream));
    vkp_job* j = new vkp_job{ctx, ev, ctx->seq, VKP_JOB_COMPUTE};
    *job = j;
  }
  return VKP_OK;
}

int vkp_after_launch(vkp_ctx* ctx, const char* what) {
  ctx->kernel_launches++;
  cudaError_t e = cudaPeekAtLastError();
  if (e == cudaSuccess && ctx->debug_sync) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return vkp_set_error("kernel %s failed: %s", what, cudaGetErrorString(e));
  }
  return VKP_OK;
}

int vkp_workspace(vkp_ctx* ctx, int slot, size_t bytes, void** out) {
  if (bytes > ctx->workspace_bytes[slot]) {
    if (ctx->workspace[slot]) {
      VKP_CUDA(cudaStreamSynchronize(ctx->stream));
      VKP_CUDA(cudaFree(ctx->workspace[slot]));
      ctx->workspace[slot] = nullptr;
      ctx->workspace_bytes[slot] = 0;
    }
    size_t want = bytes < (8u << 20) ? (8u << 20) : bytes;
    VKP_CUDA(cudaMalloc(&ctx->workspace[slot], want));
    ctx->workspace_bytes[slot] = want;
  }
  *out = ctx->workspace[slot];
  return VKP_OK;
}

// ---- host <-> buffer ------------------------------------------------------------------
static const size_t STAGE_BYTES = 128u << 20;   // one chunk = ~2.4 ms of PCIe 5 x16

static __global__ void __launch_bounds__(256)
stage_copy_kernel(uint4* __restrict__ dst, const uint4* __restrict__ src, size_t n16, size_t nwords) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
  const uint32_t* s = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d = reinterpret_cast<uint32_t*>(dst);
  for (size_t i = n16 * 4 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nwords; i += stride) d[i] = s[i];
}

static bool is_page_locked(const void* p) {
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) == cudaSuccess) return attr.type == cudaMemoryTypeHost;
  cudaGetLastError();
  return false;
}

static int stage_launch(vkp_ctx* ctx, cudaStream_t st, void* dst, const void* src, size_t bytes) {
  const size_t nwords = bytes / 4, n16 = bytes / 16;
  const unsigned grid = vkp_grid_for(ctx, n16 ? n16 : 1, 256 * 4, 4);
  stage_copy_kernel<<<grid, 256, 0, st>>>((uint4*)dst, (const uint4*)src, n16, nwords);
  return vkp_after_launch(ctx, "stage_copy");
}

// page-locked host memory <-> buffer through the bounce buffer of stream slot `slot`, chunk by chunk
// (everything is in order on `st`, so one bounce buffer per stream is enough)
static int staged_copy(vkp_ctx* ctx, cudaStream_t st, int slot, void* dst, const void* src, size_t bytes, bool h2d) {
  VKP_CHECK(bytes % 4 == 0, "staged copy: %zu bytes is not a whole number of elements", bytes);
  if (!ctx->stage[slot]) VKP_CUDA(cudaMalloc(&ctx->stage[slot], STAGE_BYTES));
  char* bounce = (char*)ctx->stage[slot];
  for (size_t off = 0; off < bytes; off += STAGE_BYTES) {
    const size_t n = bytes - off < STAGE_BYTES ? bytes - off : STAGE_BYTES;
    if (h2d) {
      VKP_CUDA(cudaMemcpyAsync(bounce, (const char*)src + off, n, cudaMemcpyHostToDevice, st));
      VKP_TRY(stage_launch(ctx, st, (char*)dst + off, bounce, n));
    } else {
      VKP_TRY(stage_launch(ctx, st, bounce, (const char*)src + off, n));
      VKP_CUDA(cudaMemcpyAsync((char*)dst + off, bounce, n, cudaMemcpyDeviceToHost, st));
    }
  }
  return VKP_OK;
}

static const size_t STAGE_MIN_BYTES = 4u << 20;   // below this a direct copy is as fast

// page-locked transfers bounce only when the device side is a managed block
static bool wants_bounce(vkp_ctx* ctx, const void* dev_ptr, size_t bytes) {
  if (bytes < STAGE_MIN_BYTES || bytes % 4 != 0) return false;
  auto it = ctx->blocks.find(const_cast<void*>(dev_ptr));
  return it != ctx->blocks.end() && it->second->managed;
}

extern "C" int vkp_upload(vkp_ctx* ctx, void* dst, const void* src_host, size_t bytes) {
  VKP_RANGE("vkp_upload");
  VKP_CHECK(ctx && (bytes == 0 || (dst && src_host)), "vkp_upload: null argument");
  if (bytes == 0) return VKP_OK;
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[1] = {dst};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 1));
  // stream-ordered: earlier kernels that still use a recycled block finish first
  const bool locked = is_page_locked(src_host);
  if (locked && wants_bounce(ctx, dst, bytes))
    VKP_TRY(staged_copy(ctx, ctx->stream, 2, dst, src_host, bytes, true));
  else
    VKP_CUDA(cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyDefault, ctx->stream));
  ctx->seq++;
  // Contract: the source may be reused as soon as this returns (the reference's Buffer::set is a
  // plain memcpy).  Pageable sources are already staged by the runtime; page-locked ones are
  // read by DMA asynchronously, so wait for them.
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, src_host) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered)
    return sync_locked(ctx);
  cudaGetLastError();
  return VKP_OK;
}

extern "C" int vkp_download(vkp_ctx* ctx, void* dst_host, const void* src, size_t bytes) {
  VKP_RANGE("vkp_download");
  VKP_CHECK(ctx && (bytes == 0 || (dst_host && src)), "vkp_download: null argument");
  if (bytes == 0) return VKP_OK;
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[1] = {const_cast<void*>(src)};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 1));
  if (wants_bounce(ctx, src, bytes) && is_page_locked(dst_host))
    VKP_TRY(staged_copy(ctx, ctx->stream, 2, dst_host, src, bytes, false));
  else
    VKP_CUDA(cudaMemcpyAsync(dst_host, src, bytes, cudaMemcpyDefault, ctx->stream));
  ctx->seq++;
  return sync_locked(ctx);
}

// ---- copy-engine transfers that overlap compute (and each other) ---------------------------
static int ensure_copy_streams(vkp_ctx* ctx) {
  if (!ctx->h2d_stream) VKP_CUDA(cudaStreamCreateWithFlags(&ctx->h2d_stream, cudaStreamNonBlocking));
  if (!ctx->d2h_stream) VKP_CUDA(cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
  return VKP_OK;
}

// `side` runs after everything enqueued on the compute stream so far
static int order_after_compute(vkp_ctx* ctx, cudaStream_t side) {
  cudaEvent_t ev;
  VKP_TRY(take_event(ctx, &ev));
  VKP_CUDA(cudaEventRecord(ev, ctx->stream));
  VKP_CUDA(cudaStreamWaitEvent(side, ev, 0));
  ctx->event_pool.push_back(ev);   // the wait has captured this recording; re-recording later is legal
  return VKP_OK;
}

extern "C" int vkp_upload_async(vkp_ctx* ctx, void* dst, const void* src_pinned, size_t bytes, vkp_job** job) {
  VKP_RANGE("vkp_upload_async");
  VKP_CHECK(ctx && dst && src_pinned && job, "vkp_upload_async: null argument");
  VKP_TRY(vkp_make_current(ctx));
  VKP_CHECK(is_page_locked(src_pinned), "vkp_upload_async: the source must be page-locked (vkp_host_alloc)");
  std::lock_guard<std::mutex> g(ctx->mu);
  auto it = ctx->blocks.find(dst);
  VKP_CHECK(it != ctx->blocks.end() && it->second->in_use, "vkp_upload_async: %p is not a live buffer", dst);
  vkp_block* b = it->second;
  VKP_CHECK(bytes <= b->bytes && bytes % 4 == 0, "vkp_upload_async: %zu bytes do not fit the buffer", bytes);
  VKP_TRY(ensure_copy_streams(ctx));
  const uint64_t busy = b->guard_seq > b->last_seq ? b->guard_seq : b->last_seq;
  if (busy > ctx->done_seq) VKP_TRY(order_after_compute(ctx, ctx->h2d_stream));
  if (b->d2h_ev) {   // a download still reads the block
    VKP_CUDA(cudaStreamWaitEvent(ctx->h2d_stream, b->d2h_ev, 0));
    ctx->event_pool.push_back(b->d2h_ev);
    b->d2h_ev = nullptr;
  }
  if (b->host_dirty) {
    VKP_CUDA(cudaMemPrefetchAsync(b->ptr, b->bytes, ctx->device, ctx->h2d_stream));
    b->host_dirty = false;
  }
  if (wants_bounce(ctx, dst, bytes))
    VKP_TRY(staged_copy(ctx, ctx->h2d_stream, 0, dst, src_pinned, bytes, true));
  else if (bytes)
    VKP_CUDA(cudaMemcpyAsync(dst, src_pinned, bytes, cudaMemcpyHostToDevice, ctx->h2d_stream));
  if (!b->h2d_ev) VKP_TRY(take_event(ctx, &b->h2d_ev));
  VKP_CUDA(cudaEventRecord(b->h2d_ev, ctx->h2d_stream));
  b->guard_seq = 0;
  cudaEvent_t jev;
  VKP_TRY(take_event(ctx, &jev));
  VKP_CUDA(cudaEventRecord(jev, ctx->h2d_stream));
  *job = new vkp_job{ctx, jev, 0, VKP_JOB_UPLOAD};
  return VKP_OK;
}

extern "C" int vkp_download_async(vkp_ctx* ctx, void* dst_pinned, const void* src, size_t bytes, vkp_job** job) {
  VKP_RANGE("vkp_download_async");
  VKP_CHECK(ctx && dst_pinned && src && job, "vkp_download_async: null argument");
  VKP_TRY(vkp_make_current(ctx));
  VKP_CHECK(is_page_locked(dst_pinned), "vkp_download_async: the destination must be page-locked (vkp_host_alloc)");
  std::lock_guard<std::mutex> g(ctx->mu);
  auto it = ctx->blocks.find(const_cast<void*>(src));
  VKP_CHECK(it != ctx->blocks.end() && it->second->in_use, "vkp_download_async: %p is not a live buffer", src);
  vkp_block* b = it->second;
  VKP_CHECK(bytes <= b->bytes && bytes % 4 == 0, "vkp_download_async: %zu bytes exceed the buffer", bytes);
  VKP_TRY(ensure_copy_streams(ctx));
  VKP_TRY(order_after_compute(ctx, ctx->d2h_stream));
  if (b->h2d_ev) VKP_CUDA(cudaStreamWaitEvent(ctx->d2h_stream, b->h2d_ev, 0));   // stays pending for compute
  if (wants_bounce(ctx, src, bytes))
    VKP_TRY(staged_copy(ctx, ctx->d2h_stream, 1, dst_pinned, src, bytes, false));
  else if (bytes)
    VKP_CUDA(cudaMemcpyAsync(dst_pinned, src, bytes, cudaMemcpyDeviceToHost, ctx->d2h_stream));
  if (!b->d2h_ev) VKP_TRY(take_event(ctx, &b->d2h_ev));
  VKP_CUDA(cudaEventRecord(b->d2h_ev, ctx->d2h_stream));
  cudaEvent_t jev;
  VKP_TRY(take_event(ctx, &jev));
  VKP_CUDA(cudaEventRecord(jev, ctx->d2h_stream));
  *job = new vkp_job{ctx, jev, ctx->seq, VKP_JOB_DOWNLOAD};
  return VKP_OK;
}

// Buffers are born in plain device memory.  The first time the host wants to look at one through
// its pointer (NumPy view), the contents move into a managed block and the buffer lives there from
// then on: `*out` replaces `ptr`, which is released.  Keeps managed memory -- a scarce, system-wide
// resource on multi-GPU boxes (scripts/micro/managed_limit.py: ~64 GiB over all processes) -- for the
// few arrays the host actually views, and lets everything else DMA and run from cudaMalloc memory.
extern "C" int vkp_host_view(vkp_ctx* ctx, void* ptr, void** out) {
  VKP_CHECK(ctx && ptr && out, "vkp_host_view: null argument");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  auto it = ctx->blocks.find(ptr);
  VKP_CHECK(it != ctx->blocks.end() && it->second->in_use, "vkp_host_view: %p is not a live buffer", ptr);
  vkp_block* b = it->second;
  if (b->managed) {
    *out = ptr;
    return VKP_OK;
  }
  void* np = nullptr;
  int rc = alloc_locked(ctx, b->bytes, &np, 0, 1);
  if (rc != VKP_OK) {
    const std::string why(vkp_last_error());   // vkp_set_error formats into the buffer this points at
    return vkp_set_error("host view of a %zu-byte buffer: no managed memory left (%s); use vkp_download / "
                         "Array.to_host() for bulk reads", b->bytes, why.c_str());
  }
  vkp_block* nb = ctx->blocks[np];
  void* bufs[2] = {ptr, np};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 2));   // also orders the copy after copy-engine transfers
  VKP_TRY(stage_launch(ctx, ctx->stream, np, ptr, b->bytes));
  ctx->seq++;
  nb->guard_seq = ctx->seq;                     // vkp_host_acquire waits for the move
  b->in_use = false;
  b->guard_seq = b->last_seq;
  ctx->live_bytes -= b->bytes;
  ctx->free_lists[0][b->bytes].push_back(b);
  *out = np;
  return VKP_OK;
}

extern "C" int vkp_host_acquire(vkp_ctx* ctx, void* ptr, size_t bytes, int mode) {
  VKP_CHECK(ctx && ptr, "vkp_host_acquire: null argument");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  auto it = ctx->blocks.find(ptr);
  VKP_CHECK(it != ctx->blocks.end(), "vkp_host_acquire: %p is not a buffer of this context", ptr);
  vkp_block* b = it->second;
  VKP_CHECK(b->managed, "vkp_host_acquire: %p is device memory; call vkp_host_view first", ptr);
  const bool prefetch = (mode & 1) != 0, writing = (mode & 2) != 0;
  if (b->h2d_ev) {   // copy-engine upload in flight: the host view must see it
    VKP_CUDA(cudaEventSynchronize(b->h2d_ev));
    ctx->event_pool.push_back(b->h2d_ev);
    b->h2d_ev = nullptr;
  }
  if (b->d2h_ev && (writing || prefetch)) {
    VKP_CUDA(cudaEventSynchronize(b->d2h_ev));
    ctx->event_pool.push_back(b->d2h_ev);
    b->d2h_ev = nullptr;
  }
  bool need_sync = b->guard_seq > ctx->done_seq;       // earlier tenant of a recycled block
  if (writing && ctx->seq > ctx->done_seq) need_sync = true;  // readers of this array in flight
  if (prefetch && !b->host_dirty && bytes > 0) {
    size_t nbytes = bytes < b->bytes ? bytes : b->bytes;
    VKP_CUDA(cudaMemPrefetchAsync(ptr, nbytes, cudaCpuDeviceId, ctx->stream));
    ctx->seq++;
    need_sync = true;
  }
  if (need_sync) VKP_TRY(sync_locked(ctx));
  b->guard_seq = 0;
  b->host_dirty = true;
  return VKP_OK;
}

extern "C" int vkp_host_alloc(size_t bytes, void** ptr) {
  VKP_CHECK(ptr, "vkp_host_alloc: null argument");
  VKP_CUDA(cudaMallocHost(ptr, bytes ? bytes : 1));
  return VKP_OK;
}

extern "C" int vkp_host_free(void* ptr) {
  if (ptr) VKP_CUDA(cudaFreeHost(ptr));
  return VKP_OK;
}

// ---- jobs -------------------------------------------------------------------------------
extern "C" int vkp_job_wait(vkp_job* job, uint64_t timeout_ns) {
  VKP_CHECK(job, "vkp_job_wait: null job");
  vkp_ctx* ctx = job->ctx;
  if (job->kind == VKP_JOB_COMPUTE && ctx->done_seq >= job->seq) return VKP_OK;
  VKP_TRY(vkp_make_current(ctx));
  if (timeout_ns == UINT64_MAX) {
    cudaError_t e = cudaEventSynchronize(job->ev);
    if (e != cudaSuccess) return vkp_set_error("Error at Command Wait: %s", cudaGetErrorString(e));
  } else {
    auto t0 = std::chrono::steady_clock::now();
    for (;;) {
      cudaError_t e = cudaEventQuery(job->ev);
      if (e == cudaSuccess) break;
      if (e != cudaErrorNotReady) {
        cudaGetLastError();
        return vkp_set_error("Error at Command Wait: %s", cudaGetErrorString(e));
      }
      auto dt = std::chrono::duration_cast<std::chrono::nanoseconds>(
                    std::chrono::steady_clock::now() - t0).count();
      if ((uint64_t)dt >= timeout_ns) {
        vkp_set_error("Timeout at Command Wait");
        return VKP_TIMEOUT;
      }
      std::this_thread::yield();
    }
  }
  std::lock_guard<std::mutex> g(ctx->mu);
  if (job->seq > ctx->done_seq) ctx->done_seq = job->seq;
  return VKP_OK;
}

extern "C" int vkp_job_done(vkp_job* job, int* done) {
  VKP_CHECK(job && done, "vkp_job_done: null argument");
  if (job->kind == VKP_JOB_COMPUTE && job->ctx->done_seq >= job->seq) { *done = 1; return VKP_OK; }
  cudaError_t e = cudaEventQuery(job->ev);
  if (e == cudaSuccess) { *done = 1; return VKP_OK; }
  if (e == cudaErrorNotReady) { *done = 0; return VKP_OK; }
  cudaGetLastError();
  return vkp_set_error("cudaEventQuery failed: %s", cudaGetErrorString(e));
}

extern "C" int vkp_job_release(vkp_job* job) {
  if (!job) return VKP_OK;
  vkp_ctx* ctx = job->ctx;
  {
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->event_pool.push_back(job->ev);  // re-recording a pooled event is legal
  }
  delete job;
  return VKP_OK;
}

// ---- timers -------------------------------------------------------------------------------
extern "C" int vkp_timer_create(vkp_ctx* ctx, vkp_timer** out) {
  VKP_CHECK(ctx && out, "vkp_timer_create: null argument");
  VKP_TRY(vkp_make_current(ctx));
  cudaEvent_t ev;
  VKP_CUDA(cudaEventCreate(&ev));
  *out = new vkp_timer{ctx, ev};
  return VKP_OK;
}

extern "C" int vkp_timer_record(vkp_timer* t) {
  VKP_CHECK(t, "vkp_timer_record: null timer");
  VKP_TRY(vkp_make_current(t->ctx));
  VKP_CUDA(cudaEventRecord(t->ev, t->ctx->stream));
  return VKP_OK;
}

extern "C" int vkp_timer_elapsed_ms(vkp_timer* start, vkp_timer* stop, float* ms) {
  VKP_CHECK(start && stop && ms, "vkp_timer_elapsed_ms: null argument");
  VKP_TRY(vkp_make_current(stop->ctx));
  VKP_CUDA(cudaEventSynchronize(stop->ev));
  VKP_CUDA(cudaEventElapsedTime(ms, start->ev, stop->ev));
  return VKP_OK;
}

extern "C" int vkp_timer_destroy(vkp_timer* t) {
  if (!t) return VKP_OK;
  cudaEventDestroy(t->ev);
  delete t;
  return VKP_OK;
}
