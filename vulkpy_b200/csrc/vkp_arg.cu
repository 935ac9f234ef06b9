// vkp_arg.cu -- argmax / argmin and random permutations on the device.
//
// The reference lists both as missing (README.md:73 "argmax, argmin", :77 "shuffle") and its own
// training example falls back to the host for them (example/02-nn.py:82 `rng.shuffle(idx)`,
// :96 `np.argmax(pred_y, axis=1)`).  Semantics follow NumPy, which is what that example calls:
// first occurrence wins, a NaN counts as the extreme value.  Indices are uint32 like every index
// in the reference (U32Array, vkarray.py:191-257).
//
// Layout is the reductions' [prev, axis, post] (vkarray.py:1398-1432).  post == 1: a CTA scans a
// segment of one row with coalesced (float4 where aligned) loads and reduces (value, index) pairs
// by warp shuffle; post > 1: threads run along `post`, each scanning its column of a segment.
// Long axes are cut into segments so that the grid fills the machine; the per-segment winners go
// through a second, deterministic pass.
//
// permutation(n): indices 0..n-1 stably sorted by n keys drawn from the Xoshiro128++ stream
// (vkp_rng_uint32), i.e. np.argsort(keys, kind="stable").  The radix sort is cub's
// DeviceRadixSort (CUDA toolkit header library); it is a helper, not one of the hot-path rows.
#include "vkp_common.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace {

struct VI {
  float v;
  uint32_t i;
};

// is (v, i) a better candidate than (bv, bi)?  MAX: larger value, NaN beats everything, ties and
// NaN-vs-NaN go to the lower index.
template <bool MAX>
__device__ __forceinline__ bool better(float v, uint32_t i, float bv, uint32_t bi) {
  const bool vn = v != v, bn = bv != bv;
  if (vn || bn) return vn && (!bn || i < bi);
  if (MAX ? (v > bv) : (v < bv)) return true;
  return v == bv && i < bi;
}

template <bool MAX>
__device__ __forceinline__ void take(VI& b, float v, uint32_t i) {
  if (better<MAX>(v, i, b.v, b.i)) {
    b.v = v;
    b.i = i;
  }
}

constexpr uint32_t NONE = 0xffffffffu;

// rows [rows, axis]; CTA (blockIdx.x = segment, blockIdx.y = row)
template <bool MAX>
__global__ void __launch_bounds__(256)
arg_rows_kernel(const float* __restrict__ a, float* __restrict__ pv, uint32_t* __restrict__ pi, uint32_t* __restrict__ out,
                uint32_t axis, uint32_t seg_len, uint32_t nseg, uint32_t elem0_mod4) {
  const uint32_t row = blockIdx.y, seg = blockIdx.x;
  const float* r = a + (size_t)row * axis;
  const uint32_t lo = seg * seg_len;
  const uint32_t hi = min(axis, lo + seg_len);
  VI b{0.0f, NONE};
  // float4 body when the row start is 16-byte aligned (seg_len is a multiple of 4)
  if (((elem0_mod4 + (size_t)row * axis) & 3) == 0) {
    const uint32_t nv = (hi - lo) >> 2;
    const float4* rv = reinterpret_cast<const float4*>(r + lo);
    auto fold4 = [&](const float4 x, uint32_t i) {
      if (b.i == NONE) { b.v = x.x; b.i = i; } else take<MAX>(b, x.x, i);
      take<MAX>(b, x.y, i + 1);
      take<MAX>(b, x.z, i + 2);
      take<MAX>(b, x.w, i + 3);
    };
    uint32_t v = threadIdx.x;
    for (; v + 768 < nv; v += 1024) {   // four independent 16-byte loads in flight
      const float4 x0 = rv[v], x1 = rv[v + 256], x2 = rv[v + 512], x3 = rv[v + 768];
      fold4(x0, lo + (v << 2)); fold4(x1, lo + ((v + 256) << 2)); fold4(x2, lo + ((v + 512) << 2)); fold4(x3, lo + ((v + 768) << 2));
    }
    for (; v < nv; v += 256) fold4(rv[v], lo + (v << 2));
    for (uint32_t i = lo + (nv << 2) + threadIdx.x; i < hi; i += 256) {
      if (b.i == NONE) { b.v = r[i]; b.i = i; } else take<MAX>(b, r[i], i);
    }
  } else {
    for (uint32_t i = lo + threadIdx.x; i < hi; i += 256) {
      if (b.i == NONE) { b.v = r[i]; b.i = i; } else take<MAX>(b, r[i], i);
    }
  }
  // lanes that saw nothing carry index NONE: any real candidate beats them (index order), and a
  // NaN value with index NONE cannot occur because b.v stays 0
  __shared__ float sv[8];
  __shared__ uint32_t si[8];
  auto merge = [](VI x, float v, uint32_t i) {
    if (i != NONE && (x.i == NONE || better<MAX>(v, i, x.v, x.i))) { x.v = v; x.i = i; }
    return x;
  };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v = __shfl_xor_sync(0xffffffffu, b.v, o);
    const uint32_t i = __shfl_xor_sync(0xffffffffu, b.i, o);
    b = merge(b, v, i);
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = b.v; si[threadIdx.x >> 5] = b.i; }
  __syncthreads();
  if (threadIdx.x < 32) {
    VI c{threadIdx.x < 8 ? sv[threadIdx.x] : 0.0f, threadIdx.x < 8 ? si[threadIdx.x] : NONE};
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      const float v = __shfl_xor_sync(0xffffffffu, c.v, o);
      const uint32_t i = __shfl_xor_sync(0xffffffffu, c.i, o);
      c = merge(c, v, i);
    }
    if (threadIdx.x == 0) {
      if (nseg == 1) out[row] = c.i;
      else { pv[(size_t)row * nseg + seg] = c.v; pi[(size_t)row * nseg + seg] = c.i; }
    }
  }
}

// small rows: one thread per row (axis <= 64), neighbouring threads read neighbouring rows
template <bool MAX>
__global__ void __launch_bounds__(256)
arg_rows_small_kernel(const float* __restrict__ a, uint32_t* __restrict__ out, uint32_t rows, uint32_t axis) {
  const uint32_t row = blockIdx.x * 256 + threadIdx.x;
  if (row >= rows) return;
  const float* r = a + (size_t)row * axis;
  VI b{r[0], 0};
  for (uint32_t i = 1; i < axis; i++) take<MAX>(b, r[i], i);
  out[row] = b.i;
}

// a [prev, axis, post], threads along post; blockIdx.y = segment, blockIdx.z = prev index.
// VEC = 4: a thread owns four neighbouring columns (16-byte loads, post % 4 == 0, aligned base),
// four rows in flight = 64 bytes per thread; VEC = 1 is the scalar form for any shape.
template <bool MAX, int VEC>
__global__ void __launch_bounds__(256)
arg_cols_kernel(const float* __restrict__ a, float* __restrict__ pv, uint32_t* __restrict__ pi, uint32_t* __restrict__ out,
                uint32_t axis, uint32_t post, uint32_t seg_len, uint32_t nseg) {
  const uint32_t q = (blockIdx.x * 256 + threadIdx.x) * VEC;
  if (q >= post) return;
  const uint32_t p = blockIdx.z, seg = blockIdx.y;
  const uint32_t lo = seg * seg_len, hi = min(axis, lo + seg_len);
  const float* col = a + (size_t)p * axis * post + q;
  VI b[VEC];
  auto load = [&](uint32_t i, float (&x)[VEC]) {
    if constexpr (VEC == 4) {
      const float4 v = *reinterpret_cast<const float4*>(col + (size_t)i * post);
      x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    } else {
      x[0] = col[(size_t)i * post];
    }
  };
  {
    float x[VEC];
    load(lo, x);
#pragma unroll
    for (int c = 0; c < VEC; c++) b[c] = VI{x[c], lo};
  }
  uint32_t i = lo + 1;
  for (; i + 3 < hi; i += 4) {   // four independent loads in flight
    float x0[VEC], x1[VEC], x2[VEC], x3[VEC];
    load(i, x0); load(i + 1, x1); load(i + 2, x2); load(i + 3, x3);
#pragma unroll
    for (int c = 0; c < VEC; c++) {
      take<MAX>(b[c], x0[c], i); take<MAX>(b[c], x1[c], i + 1); take<MAX>(b[c], x2[c], i + 2); take<MAX>(b[c], x3[c], i + 3);
    }
  }
  for (; i < hi; i++) {
    float x[VEC];
    load(i, x);
#pragma unroll
    for (int c = 0; c < VEC; c++) take<MAX>(b[c], x[c], i);
  }
#pragma unroll
  for (int c = 0; c < VEC; c++) {
    if (nseg == 1) out[(size_t)p * post + q + c] = b[c].i;
    else {
      const size_t o = ((size_t)p * nseg + seg) * post + q + c;
      pv[o] = b[c].v;
      pi[o] = b[c].i;
    }
  }
}

// winners of the segments, in segment order (so ties still go to the lowest index):
// partials [prev, nseg, post] -> out [prev, post]
template <bool MAX>
__global__ void __launch_bounds__(256)
arg_final_kernel(const float* __restrict__ pv, const uint32_t* __restrict__ pi, uint32_t* __restrict__ out,
                 uint64_t total, uint32_t post, uint32_t nseg) {
  const uint64_t t = blockIdx.x * (uint64_t)256 + threadIdx.x;
  if (t >= total) return;
  const uint64_t p = t / post, q = t % post;
  const size_t base = (size_t)p * nseg * post + q;
  VI b{pv[base], pi[base]};
  for (uint32_t s = 1; s < nseg; s++) take<MAX>(b, pv[base + (size_t)s * post], pi[base + (size_t)s * post]);
  out[t] = b.i;
}

// the same fold with one WARP per output, for many segments per output (a full reduction has 8 CTAs
// per SM = 1184 partials for its single output: a lone thread would walk them for ~0.1 ms)
template <bool MAX>
__global__ void __launch_bounds__(256)
arg_final_warp_kernel(const float* __restrict__ pv, const uint32_t* __restrict__ pi, uint32_t* __restrict__ out,
                      uint64_t total, uint32_t post, uint32_t nseg) {
  const uint64_t t = (blockIdx.x * (uint64_t)256 + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (t >= total) return;
  const uint64_t p = t / post, q = t % post;
  const size_t base = (size_t)p * nseg * post + q;
  VI b{0.0f, NONE};
  for (uint32_t s = lane; s < nseg; s += 32) {
    const float v = pv[base + (size_t)s * post];
    const uint32_t i = pi[base + (size_t)s * post];
    if (b.i == NONE) { b.v = v; b.i = i; } else take<MAX>(b, v, i);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v = __shfl_xor_sync(0xffffffffu, b.v, o);
    const uint32_t i = __shfl_xor_sync(0xffffffffu, b.i, o);
    if (i != NONE && (b.i == NONE || better<MAX>(v, i, b.v, b.i))) { b.v = v; b.i = i; }
  }
  if (lane == 0) out[t] = b.i;
}

template <bool MAX>
void launch_arg_final(vkp_ctx* ctx, const float* pv, const uint32_t* pi, uint32_t* out, uint64_t outputs, uint32_t post,
                      uint32_t nseg) {
  if (nseg >= 64 && outputs <= (1u << 20))
    arg_final_warp_kernel<MAX><<<(unsigned)((outputs * 32 + 255) / 256), 256, 0, ctx->stream>>>(pv, pi, out, outputs, post, nseg);
  else
    arg_final_kernel<MAX><<<(unsigned)((outputs + 255) / 256), 256, 0, ctx->stream>>>(pv, pi, out, outputs, post, nseg);
}

template <bool MAX>
int arg_launch(vkp_ctx* ctx, const float* in, uint32_t* out, uint32_t prev, uint32_t axis, uint32_t post) {
  const uint64_t outputs = (uint64_t)prev * post;
  // segments: enough CTAs to fill the machine, at least 4096 elements (rows) / 64 elements (columns) each
  const uint64_t want_ctas = (uint64_t)ctx->sms * 8;
  if (post == 1) {
    if (axis <= 64) {
      arg_rows_small_kernel<MAX><<<(unsigned)((prev + 255) / 256), 256, 0, ctx->stream>>>(in, out, prev, axis);
      return vkp_after_launch(ctx, "arg_rows_small");
    }
    uint32_t nseg = (uint32_t)std::min<uint64_t>((want_ctas + prev - 1) / prev, (axis + 4095) / 4096);
    if (nseg < 1) nseg = 1;
    uint32_t seg_len = (axis + nseg - 1) / nseg;
    seg_len = (seg_len + 3) & ~3u;
    nseg = (axis + seg_len - 1) / seg_len;
    VKP_CHECK(prev <= 65535u * 32768u, "argmax: too many rows");
    float* pv = nullptr;
    uint32_t* pi = nullptr;
    if (nseg > 1) {
      void* ws;
      VKP_TRY(vkp_workspace(ctx, 0, outputs * nseg * 8, &ws));
      pv = (float*)ws;
      pi = (uint32_t*)ws + outputs * nseg;
    }
    // blockIdx.y is limited to 65535: fold larger row counts into several launches
    for (uint64_t r0 = 0; r0 < prev; r0 += 65535) {
      const uint32_t nr = (uint32_t)std::min<uint64_t>(65535, prev - r0);
      arg_rows_kernel<MAX><<<dim3(nseg, nr), 256, 0, ctx->stream>>>(
          in + r0 * axis, pv ? pv + r0 * nseg : nullptr, pi ? pi + r0 * nseg : nullptr, out + r0, axis, seg_len, nseg,
          (uint32_t)((r0 * axis) & 3));
      VKP_TRY(vkp_after_launch(ctx, "arg_rows"));
    }
    if (nseg > 1) {
      launch_arg_final<MAX>(ctx, pv, pi, out, outputs, 1, nseg);
      VKP_TRY(vkp_after_launch(ctx, "arg_final"));
    }
    return VKP_OK;
  }
  const bool vec = post % 4 == 0 && ((((uintptr_t)in) & 15) == 0) && post >= 1024;
  const uint32_t bx = vec ? (post / 4 + 255) / 256 : (post + 255) / 256;
  uint32_t nseg = (uint32_t)std::min<uint64_t>((want_ctas + (uint64_t)bx * prev - 1) / ((uint64_t)bx * prev), (axis + 63) / 64);
  if (nseg < 1) nseg = 1;
  if (nseg > 65535) nseg = 65535;
  const uint32_t seg_len = (axis + nseg - 1) / nseg;
  nseg = (axis + seg_len - 1) / seg_len;
  float* pv = nullptr;
  uint32_t* pi = nullptr;
  if (nseg > 1) {
    void* ws;
    VKP_TRY(vkp_workspace(ctx, 0, outputs * nseg * 8, &ws));
    pv = (float*)ws;
    pi = (uint32_t*)ws + outputs * nseg;
  }
  for (uint64_t p0 = 0; p0 < prev; p0 += 65535) {
    const uint32_t np = (uint32_t)std::min<uint64_t>(65535, prev - p0);
    if (vec)
      arg_cols_kernel<MAX, 4><<<dim3(bx, nseg, np), 256, 0, ctx->stream>>>(
          in + p0 * axis * post, pv ? pv + p0 * nseg * post : nullptr, pi ? pi + p0 * nseg * post : nullptr,
          out + p0 * post, axis, post, seg_len, nseg);
    else
      arg_cols_kernel<MAX, 1><<<dim3(bx, nseg, np), 256, 0, ctx->stream>>>(
          in + p0 * axis * post, pv ? pv + p0 * nseg * post : nullptr, pi ? pi + p0 * nseg * post : nullptr,
          out + p0 * post, axis, post, seg_len, nseg);
    VKP_TRY(vkp_after_launch(ctx, "arg_cols"));
  }
  if (nseg > 1) {
    launch_arg_final<MAX>(ctx, pv, pi, out, outputs, post, nseg);
    VKP_TRY(vkp_after_launch(ctx, "arg_final"));
  }
  return VKP_OK;
}

__global__ void iota_kernel(uint32_t* out, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

}  // namespace

extern "C" int vkp_argreduce(vkp_ctx* ctx, int op, const float* in, uint32_t* out, uint32_t prev, uint32_t axis,
                             uint32_t post, vkp_job** job) {
  VKP_RANGE(__func__);
  VKP_CHECK(ctx && in && out, "vkp_argreduce: null argument");
  VKP_CHECK(op == 0 || op == 1, "vkp_argreduce: op must be 0 (max) or 1 (min)");
  VKP_CHECK(prev >= 1 && axis >= 1 && post >= 1, "attempt to get argmax of an empty sequence");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[2] = {const_cast<float*>(in), out};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 2));
  VKP_TRY(op == 0 ? arg_launch<true>(ctx, in, out, prev, axis, post) : arg_launch<false>(ctx, in, out, prev, axis, post));
  return vkp_finish_op(ctx, job);
}

// out[0..n) = indices 0..n-1 stably sorted by keys[0..n)  (keys are left untouched)
extern "C" int vkp_argsort_u32(vkp_ctx* ctx, const uint32_t* keys, uint32_t* out, uint32_t n, vkp_job** job) {
  VKP_RANGE(__func__);
  VKP_CHECK(ctx && (n == 0 || (keys && out)), "vkp_argsort_u32: null argument");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[2] = {const_cast<uint32_t*>(keys), out};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 2));
  if (n > 0) {
    size_t tmp_bytes = 0;
    VKP_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                             (uint32_t*)nullptr, (int)n, 0, 32, ctx->stream));
    const size_t words = ((size_t)n + 63) & ~(size_t)63;
    void* ws;
    VKP_TRY(vkp_workspace(ctx, 0, words * 8 + tmp_bytes + 256, &ws));
    uint32_t* keys_out = (uint32_t*)ws;
    uint32_t* iota = keys_out + words;
    void* tmp = iota + words;
    iota_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(iota, n);
    VKP_TRY(vkp_after_launch(ctx, "iota"));
    VKP_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys, keys_out, (const uint32_t*)iota, out, (int)n, 0, 32,
                                             ctx->stream));
    ctx->kernel_launches += 1;
  }
  return vkp_finish_op(ctx, job);
}
