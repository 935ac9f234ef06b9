// vkp_reduce.cu -- full, axis and axis+rebroadcast reductions (sum / prod / maximum / minimum).
//
// Replaces sum.comp / sum_v1.3.comp / sum_axis.comp / sum_axis_rebroadcast.comp and the prod,
// maximum, minimum variants (shader/sum.comp:18-30, sum_v1.3.comp:20-29, sum_axis.comp:20-32,
// sum_axis_rebroadcast.comp:20-35) plus the host loop of vkarray.py:1246-1274.
//
// The reference runs ONE thread per output element with a serial loop over the axis (and, for
// the full reduction, log64(n) dependent single-workgroup jobs that only cover n <= 4096,
// SURVEY Q2).  Here an array viewed as [prev, axis, post] is reduced by one of two kernels:
//   * post == 1  -> "row" kernels: lanes run along the contiguous axis with float4 loads;
//                   short rows use a lane group per row, long rows a CTA per (row, split);
//   * post  > 1  -> "column" kernels: threads run along `post`, the axis is split over thread
//                   groups and over CTAs.  Wide arrays (post >= 128, post % 4 == 0) take the
//                   TMA-staged kernel: 32-row x 128-column boxes of the strided axis land in a
//                   6-deep shared-memory ring (cp.async.bulk.tensor + mbarriers, one issuing
//                   thread), all 256 threads only add out of shared memory; other shapes use
//                   plain (float4 when post % 4 == 0) global loads.
// When the output alone cannot fill 148 SMs the axis is split across CTAs; partials go to a
// device workspace and a second launch of the same kernel folds them, in a fixed order
// (deterministic, no atomics).  Accumulation is float32; the order differs from the
// reference's serial k = 0..axis-1 order, the parity tolerance is stated in tests/.
#include "vkp_common.cuh"

#include <cuda.h>
#include <cstdlib>

int vkp_tma_map_3d(void* map_out, const float* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1);  // vkp_gemm_tc.cu

int vkp_broadcast_copy_3d(vkp_ctx* ctx, const float* src, float* dst, uint32_t prev, uint32_t axis,
                          uint32_t post);  // vkp_broadcast.cu

namespace {

template <int OP>
struct Red;
template <> struct Red<VKR_SUM> {
  static __device__ __forceinline__ float id() { return 0.f; }
  static __device__ __forceinline__ float op(float a, float b) { return a + b; }
};
template <> struct Red<VKR_PROD> {
  static __device__ __forceinline__ float id() { return 1.f; }
  static __device__ __forceinline__ float op(float a, float b) { return a * b; }
};
template <> struct Red<VKR_MAX> {
  static __device__ __forceinline__ float id() { return -INFINITY; }
  static __device__ __forceinline__ float op(float a, float b) { return fmaxf(a, b); }
};
template <> struct Red<VKR_MIN> {
  static __device__ __forceinline__ float id() { return INFINITY; }
  static __device__ __forceinline__ float op(float a, float b) { return fminf(a, b); }
};

template <int OP>
__device__ __forceinline__ float warp_reduce(float v, int width = 32) {
  for (int o = width >> 1; o > 0; o >>= 1) v = Red<OP>::op(v, __shfl_xor_sync(0xffffffffu, v, o, 32));
  return v;
}

constexpr int RB_BLOCK = 256;

// ---- long rows: one CTA per (row, split) ---------------------------------------------------
// in: [nrows, len] contiguous; out[row * nsplit + split]
template <int OP>
__global__ void __launch_bounds__(RB_BLOCK)
reduce_rows_block(const float* __restrict__ in, float* __restrict__ out, uint32_t nrows, uint64_t len,
                  uint32_t nsplit, uint64_t seg /* multiple of 4 */) {
  __shared__ float sm[RB_BLOCK / 32];
  for (uint64_t job = blockIdx.x; job < (uint64_t)nrows * nsplit; job += gridDim.x) {
    const uint64_t row = job / nsplit;
    const uint32_t split = (uint32_t)(job - row * nsplit);
    const uint64_t begin = (uint64_t)split * seg;
    const uint64_t end = (begin + seg < len) ? begin + seg : len;
    const float* p = in + row * len;
    float acc0 = Red<OP>::id(), acc1 = Red<OP>::id(), acc2 = Red<OP>::id(), acc3 = Red<OP>::id();
    if (begin < end) {
      // peel to a 16-byte boundary, then float4
      uint64_t i0 = begin;
      const uint64_t addr = (uint64_t)(p + begin);
      uint64_t head = ((16 - (addr & 15)) & 15) >> 2;
      if (head > end - begin) head = end - begin;
      if (threadIdx.x < head) acc0 = Red<OP>::op(acc0, p[begin + threadIdx.x]);
      i0 += head;
      const uint64_t nvec = (end - i0) >> 2;
      const float4* pv = reinterpret_cast<const float4*>(p + i0);
      uint64_t v = threadIdx.x;
      for (; v + 3 * RB_BLOCK < nvec; v += 4 * RB_BLOCK) {
        const float4 x0 = pv[v], x1 = pv[v + RB_BLOCK], x2 = pv[v + 2 * RB_BLOCK], x3 = pv[v + 3 * RB_BLOCK];
        acc0 = Red<OP>::op(acc0, Red<OP>::op(Red<OP>::op(x0.x, x0.y), Red<OP>::op(x0.z, x0.w)));
        acc1 = Red<OP>::op(acc1, Red<OP>::op(Red<OP>::op(x1.x, x1.y), Red<OP>::op(x1.z, x1.w)));
        acc2 = Red<OP>::op(acc2, Red<OP>::op(Red<OP>::op(x2.x, x2.y), Red<OP>::op(x2.z, x2.w)));
        acc3 = Red<OP>::op(acc3, Red<OP>::op(Red<OP>::op(x3.x, x3.y), Red<OP>::op(x3.z, x3.w)));
      }
      for (; v < nvec; v += RB_BLOCK) {
        const float4 x0 = pv[v];
        acc0 = Red<OP>::op(acc0, Red<OP>::op(Red<OP>::op(x0.x, x0.y), Red<OP>::op(x0.z, x0.w)));
      }
      const uint64_t t0 = i0 + (nvec << 2);
      if (t0 + threadIdx.x < end) acc1 = Red<OP>::op(acc1, p[t0 + threadIdx.x]);
    }
    float acc = Red<OP>::op(Red<OP>::op(acc0, acc1), Red<OP>::op(acc2, acc3));
    acc = warp_reduce<OP>(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
      float v2 = threadIdx.x < RB_BLOCK / 32 ? sm[threadIdx.x] : Red<OP>::id();
      v2 = warp_reduce<OP>(v2, RB_BLOCK / 32);
      if (threadIdx.x == 0) out[job] = v2;
    }
    __syncthreads();
  }
}

// ---- short rows: `lpr` lanes per row (power of two <= 32) -----------------------------------
template <int OP>
__global__ void __launch_bounds__(RB_BLOCK)
reduce_rows_group(const float* __restrict__ in, float* __restrict__ out, uint32_t nrows, uint32_t len,
                  int lpr) {
  const uint32_t groups_per_block = RB_BLOCK / lpr;
  const uint32_t g = threadIdx.x / lpr, l = threadIdx.x % lpr;
  for (uint64_t row0 = (uint64_t)blockIdx.x * groups_per_block; row0 < nrows;
       row0 += (uint64_t)gridDim.x * groups_per_block) {
    const uint64_t row = row0 + g;
    float acc = Red<OP>::id();
    if (row < nrows) {
      const float* p = in + row * len;
      for (uint32_t i = l; i < len; i += lpr) acc = Red<OP>::op(acc, p[i]);
    }
    acc = warp_reduce<OP>(acc, lpr);
    if (l == 0 && row < nrows) out[row] = acc;
  }
}

// ---- column kernel: in [prev, axis, post] -> out [prev, nsplit, post] ------------------------
// blockDim = (bx, by), bx * by == 256.  VEC = 4 needs post % 4 == 0.
template <int OP, int VEC>
__global__ void __launch_bounds__(256)
reduce_cols(const float* __restrict__ in, float* __restrict__ out, uint32_t prev, uint32_t axis,
            uint32_t post, uint32_t nsplit, uint32_t seg) {
  __shared__ float sm[256 * VEC];
  const uint32_t bx = blockDim.x, by = blockDim.y;
  const uint32_t ctile = bx * VEC;
  const uint32_t ntile = (post + ctile - 1) / ctile;
  const uint64_t njobs = (uint64_t)prev * nsplit * ntile;
  for (uint64_t job = blockIdx.x; job < njobs; job += gridDim.x) {
    const uint32_t tile = (uint32_t)(job % ntile);
    const uint64_t rest = job / ntile;
    const uint32_t split = (uint32_t)(rest % nsplit);
    const uint32_t i = (uint32_t)(rest / nsplit);
    const uint32_t col = tile * ctile + threadIdx.x * VEC;
    const uint32_t k0 = split * seg;
    const uint32_t k1 = (k0 + seg < axis) ? k0 + seg : axis;
    float acc[VEC];
#pragma unroll
    for (int c = 0; c < VEC; c++) acc[c] = Red<OP>::id();
    if (col < post) {
      const float* p = in + ((uint64_t)i * axis) * post + col;
      uint32_t k = k0 + threadIdx.y;
      if constexpr (VEC == 4) {
        for (; k + 3 * by < k1; k += 4 * by) {
          const float4 x0 = *reinterpret_cast<const float4*>(p + (uint64_t)k * post);
          const float4 x1 = *reinterpret_cast<const float4*>(p + (uint64_t)(k + by) * post);
          const float4 x2 = *reinterpret_cast<const float4*>(p + (uint64_t)(k + 2 * by) * post);
          const float4 x3 = *reinterpret_cast<const float4*>(p + (uint64_t)(k + 3 * by) * post);
          acc[0] = Red<OP>::op(Red<OP>::op(acc[0], x0.x), Red<OP>::op(Red<OP>::op(x1.x, x2.x), x3.x));
          acc[1] = Red<OP>::op(Red<OP>::op(acc[1], x0.y), Red<OP>::op(Red<OP>::op(x1.y, x2.y), x3.y));
          acc[2] = Red<OP>::op(Red<OP>::op(acc[2], x0.z), Red<OP>::op(Red<OP>::op(x1.z, x2.z), x3.z));
          acc[3] = Red<OP>::op(Red<OP>::op(acc[3], x0.w), Red<OP>::op(Red<OP>::op(x1.w, x2.w), x3.w));
        }
        for (; k < k1; k += by) {
          const float4 x0 = *reinterpret_cast<const float4*>(p + (uint64_t)k * post);
          acc[0] = Red<OP>::op(acc[0], x0.x);
          acc[1] = Red<OP>::op(acc[1], x0.y);
          acc[2] = Red<OP>::op(acc[2], x0.z);
          acc[3] = Red<OP>::op(acc[3], x0.w);
        }
      } else {
        for (; k + 3 * by < k1; k += 4 * by) {
          const float x0 = p[(uint64_t)k * post], x1 = p[(uint64_t)(k + by) * post];
          const float x2 = p[(uint64_t)(k + 2 * by) * post], x3 = p[(uint64_t)(k + 3 * by) * post];
          acc[0] = Red<OP>::op(Red<OP>::op(acc[0], x0), Red<OP>::op(Red<OP>::op(x1, x2), x3));
        }
        for (; k < k1; k += by) acc[0] = Red<OP>::op(acc[0], p[(uint64_t)k * post]);
      }
    }
    // fold threadIdx.y in shared memory (fixed order)
#pragma unroll
    for (int c = 0; c < VEC; c++) sm[(threadIdx.y * bx + threadIdx.x) * VEC + c] = acc[c];
    __syncthreads();
    if (threadIdx.y == 0 && col < post) {
      float r[VEC];
#pragma unroll
      for (int c = 0; c < VEC; c++) r[c] = acc[c];
      for (uint32_t y = 1; y < by; y++) {
#pragma unroll
        for (int c = 0; c < VEC; c++) r[c] = Red<OP>::op(r[c], sm[(y * bx + threadIdx.x) * VEC + c]);
      }
      float* o = out + ((uint64_t)i * nsplit + split) * post + col;
      if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(o) = make_float4(r[0], r[1], r[2], r[3]);
      } else {
        o[0] = r[0];
      }
    }
    __syncthreads();
  }
}

// ---- column kernel, TMA-staged: in [prev, axis, post] -> out [prev, nsplit, post] ------------
// One CTA per (prev index, 128-column tile, axis split).  Thread 0 keeps CT_STAGES boxes of
// CT_ROWS x CT_COLS floats in flight through a 3-D tensor map (dims post, axis, prev; rows of the
// strided axis are post * 4 bytes apart in HBM, 512 contiguous bytes each inside a box); a box
// row past `axis` is zero-filled by TMA and a row past the CTA's segment belongs to the next
// split, so consumers bound the rows they add by k1, never by the box.  Thread (tx, ty) owns
// columns 4 tx .. 4 tx + 3 and rows ty, ty + 8, ...; the 8 row groups are folded in order.
constexpr int CT_ROWS = 32, CT_COLS = 128, CT_STAGES = 6;
constexpr int CT_STAGE_BYTES = CT_ROWS * CT_COLS * 4;
constexpr int CT_SMEM = CT_STAGES * CT_STAGE_BYTES + 128 /*align*/ + 64 /*barriers*/;

__device__ __forceinline__ uint32_t ct_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ct_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; spin++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();   // a protocol bug must fail the launch, not hang the GPU
  }
}

template <int OP>
__global__ void __launch_bounds__(256)
reduce_cols_tma(const __grid_constant__ CUtensorMap tm, float* __restrict__ out, uint32_t prev, uint32_t axis,
                uint32_t post, uint32_t nsplit, uint32_t seg) {
  extern __shared__ uint8_t ct_raw[];
  uint8_t* smem = ct_raw + ((128 - (ct_smem_u32(ct_raw) & 127)) & 127);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + CT_STAGES * CT_STAGE_BYTES);
  const uint32_t ntile = (post + CT_COLS - 1) / CT_COLS;
  const uint32_t job = blockIdx.x;
  const uint32_t tile = job % ntile;
  const uint32_t rest = job / ntile;
  const uint32_t split = rest % nsplit;
  const uint32_t pi = rest / nsplit;
  const uint32_t k0 = split * seg;
  const uint32_t k1 = (k0 + seg < axis) ? k0 + seg : axis;
  const uint32_t nbox = (k1 - k0 + CT_ROWS - 1) / CT_ROWS;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm) : "memory");
    for (int s = 0; s < CT_STAGES; s++)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ct_smem_u32(&full[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](uint32_t b) {   // box b of this CTA into slot b % CT_STAGES (thread 0 only)
    const uint32_t slot = b % CT_STAGES;
    const uint32_t bar = ct_smem_u32(&full[slot]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)CT_STAGE_BYTES) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(ct_smem_u32(smem + slot * CT_STAGE_BYTES)), "l"(&tm), "r"(bar), "r"((int)(tile * CT_COLS)),
          "r"((int)(k0 + b * CT_ROWS)), "r"((int)pi)
        : "memory");
  };
  if (threadIdx.x == 0)
    for (uint32_t b = 0; b < nbox && b < CT_STAGES; b++) issue(b);

  float acc[4];
#pragma unroll
  for (int c = 0; c < 4; c++) acc[c] = Red<OP>::id();
  for (uint32_t b = 0; b < nbox; b++) {
    const uint32_t slot = b % CT_STAGES;
    ct_mbar_wait(ct_smem_u32(&full[slot]), (b / CT_STAGES) & 1);
    const float* st = reinterpret_cast<const float*>(smem + slot * CT_STAGE_BYTES);
    const uint32_t rows_here = min((uint32_t)CT_ROWS, k1 - (k0 + b * CT_ROWS));
#pragma unroll
    for (uint32_t r = 0; r < CT_ROWS / 8; r++) {
      const uint32_t row = ty + 8 * r;
      if (row < rows_here) {
        const float4 x = *reinterpret_cast<const float4*>(st + row * CT_COLS + tx * 4);
        acc[0] = Red<OP>::op(acc[0], x.x);
        acc[1] = Red<OP>::op(acc[1], x.y);
        acc[2] = Red<OP>::op(acc[2], x.z);
        acc[3] = Red<OP>::op(acc[3], x.w);
      }
    }
    __syncthreads();                                   // every thread is done with this slot
    if (threadIdx.x == 0 && b + CT_STAGES < nbox) issue(b + CT_STAGES);
  }
  // fold the 8 row groups in order through shared memory (slot 0 is free: all boxes consumed)
  float* fold = reinterpret_cast<float*>(smem);
#pragma unroll
  for (int c = 0; c < 4; c++) fold[(ty * 32 + tx) * 4 + c] = acc[c];
  __syncthreads();
  const uint32_t col = tile * CT_COLS + tx * 4;
  if (ty == 0 && col < post) {
    float r[4] = {acc[0], acc[1], acc[2], acc[3]};
    for (uint32_t y = 1; y < 8; y++) {
#pragma unroll
      for (int c = 0; c < 4; c++) r[c] = Red<OP>::op(r[c], fold[(y * 32 + tx) * 4 + c]);
    }
    *reinterpret_cast<float4*>(out + ((uint64_t)pi * nsplit + split) * post + col) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// ---- literal sum.comp semantics for sizeB > 1: out[i] = reduce_{j = i, i+sizeB, ...} a[j] -----
template <int OP>
__global__ void reduce_strided(const float* __restrict__ in, float* __restrict__ out, uint32_t sizeA,
                               uint32_t sizeB) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= sizeB) return;
  float acc = Red<OP>::id();
  for (uint64_t j = warp + (uint64_t)lane * sizeB; j < sizeA; j += 32ull * sizeB) acc = Red<OP>::op(acc, in[j]);
  acc = warp_reduce<OP>(acc);
  if (lane == 0) out[warp] = acc;
}

template <int OP>
int reduce_rows(vkp_ctx* ctx, const float* in, float* out, uint32_t nrows, uint64_t len, uint64_t ws_off = 0) {
  if (nrows == 0) return VKP_OK;
  if (len <= 1024) {
    int lpr = 1;
    while (lpr < 32 && (uint64_t)lpr * 4 < len) lpr <<= 1;
    const unsigned gpb = RB_BLOCK / lpr;
    const unsigned grid = (nrows + gpb - 1) / gpb;
    reduce_rows_group<OP><<<grid, RB_BLOCK, 0, ctx->stream>>>(in, out, nrows, (uint32_t)len, lpr);
    return vkp_after_launch(ctx, "reduce_rows_group");
  }
  // long rows: one CTA per <= 32 Ki-element segment (many small CTAs balance better than a
  // persistent grid: profiles/r01_micro_stream_variants.txt); rows are split only when there are
  // too few of them to fill the machine.  Partials of a split go to the workspace at `ws_off`
  // and are folded by a recursive call that uses the space behind them.
  const uint64_t target_ctas = (uint64_t)ctx->sms * 32;
  uint32_t nsplit = 1;
  if (nrows < target_ctas) {
    uint64_t want = (target_ctas + nrows - 1) / nrows;
    const uint64_t max_split = (len + 16383) / 16384;  // at least 16 Ki elements per CTA
    if (want > max_split) want = max_split;
    nsplit = (uint32_t)(want < 1 ? 1 : want);
  }
  uint64_t seg = (len + nsplit - 1) / nsplit;
  seg = (seg + 3) & ~3ull;
  nsplit = (uint32_t)((len + seg - 1) / seg);
  const uint64_t jobs = (uint64_t)nrows * nsplit;
  VKP_CHECK(jobs < (1ull << 31), "reduction too large");
  const unsigned grid = (unsigned)jobs;
  if (nsplit == 1) {
    reduce_rows_block<OP><<<grid, RB_BLOCK, 0, ctx->stream>>>(in, out, nrows, len, 1, seg);
    return vkp_after_launch(ctx, "reduce_rows_block");
  }
  void* ws;
  VKP_TRY(vkp_workspace(ctx, 0, (ws_off + jobs) * sizeof(float) * 2 + 4096, &ws));
  float* part = static_cast<float*>(ws) + ws_off;
  reduce_rows_block<OP><<<grid, RB_BLOCK, 0, ctx->stream>>>(in, part, nrows, len, nsplit, seg);
  VKP_TRY(vkp_after_launch(ctx, "reduce_rows_block"));
  return reduce_rows<OP>(ctx, part, out, nrows, nsplit, ws_off + ((jobs + 63) & ~63ull));
}

template <int OP>
int reduce_axis(vkp_ctx* ctx, const float* in, float* out, uint32_t prev, uint32_t axis, uint32_t post) {
  if ((uint64_t)prev * post == 0) return VKP_OK;
  if (post == 1) return reduce_rows<OP>(ctx, in, out, prev, axis);
  const bool vec = (post % 4 == 0) && ((((uintptr_t)in) & 15) == 0) && ((((uintptr_t)out) & 15) == 0);
  // TMA-staged kernel for wide arrays (VKP_REDUCE_TMA=0 keeps the plain-load kernel everywhere)
  static const bool tma_on = !(getenv("VKP_REDUCE_TMA") && getenv("VKP_REDUCE_TMA")[0] == '0');
  if (tma_on && vec && post >= CT_COLS && axis >= 16 * CT_ROWS) {   // short axes: too few boxes per CTA to pipeline (measured: 4.5 vs 6.7 TB/s at axis = 64)
    const uint32_t ntile = (post + CT_COLS - 1) / CT_COLS;
    const uint64_t base_jobs = (uint64_t)prev * ntile;
    const uint64_t target_ctas = (uint64_t)ctx->sms * 16;
    uint32_t nsplit = 1;
    if (base_jobs < target_ctas) {
      uint64_t want = target_ctas / base_jobs;          // floor: at most 8 full waves of 2 CTAs per SM
      const uint64_t max_split = (axis + 4 * CT_ROWS - 1) / (4 * CT_ROWS);   // >= 4 boxes per CTA
      if (want > max_split) want = max_split;
      nsplit = (uint32_t)(want < 1 ? 1 : want);
    }
    uint32_t seg = (axis + nsplit - 1) / nsplit;
    seg = (seg + CT_ROWS - 1) / CT_ROWS * CT_ROWS;      // whole boxes: no box is fetched by two CTAs
    nsplit = (axis + seg - 1) / seg;
    const uint64_t jobs = base_jobs * nsplit;
    if (jobs < (1ull << 31)) {
      CUtensorMap tm;
      VKP_TRY(vkp_tma_map_3d(&tm, in, post, axis, prev, CT_COLS, CT_ROWS));
      float* dst = out;
      if (nsplit > 1) {
        void* ws;
        VKP_TRY(vkp_workspace(ctx, 0, (uint64_t)prev * nsplit * post * sizeof(float), &ws));
        dst = (float*)ws;
      }
      static bool attr_set[64] = {};       // per device: the opt-in does not carry over to GPU(1)
      if (!attr_set[ctx->device & 63]) {
        VKP_CUDA(cudaFuncSetAttribute(reduce_cols_tma<OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, CT_SMEM));
        attr_set[ctx->device & 63] = true;
      }
      reduce_cols_tma<OP><<<(unsigned)jobs, 256, CT_SMEM, ctx->stream>>>(tm, dst, prev, axis, post, nsplit, seg);
      VKP_TRY(vkp_after_launch(ctx, "reduce_cols_tma"));
      if (nsplit > 1) {   // fold the split partials [prev, nsplit, post] with the plain-load kernel
        uint32_t bx2 = 64;
        while (bx2 > 8 && (uint64_t)prev * ((post + bx2 * 4 - 1) / (bx2 * 4)) < 2ull * ctx->sms) bx2 >>= 1;
        const unsigned grid2 = (unsigned)((uint64_t)prev * ((post + bx2 * 4 - 1) / (bx2 * 4)));
        reduce_cols<OP, 4><<<grid2, dim3(bx2, 256 / bx2), 0, ctx->stream>>>(dst, out, prev, nsplit, post, 1, nsplit);
        VKP_TRY(vkp_after_launch(ctx, "reduce_cols(pass2)"));
      }
      return VKP_OK;
    }
  }
  const uint32_t cols = vec ? post / 4 : post;
  // threads along `post`: up to VKP_COLS_BX (default 128) lanes wide so that a CTA row is a long
  // contiguous run (2 KiB with float4), the rest of the 256 threads split the axis
  static const uint32_t max_bx = getenv("VKP_COLS_BX") ? (uint32_t)atoi(getenv("VKP_COLS_BX")) : 64u;
  uint32_t bx = 1;
  while (bx < max_bx && bx < cols) bx <<= 1;
  const uint32_t by = 256 / bx;
  const uint32_t ctile = bx * (vec ? 4 : 1);
  const uint32_t ntile = (post + ctile - 1) / ctile;
  const uint64_t base_jobs = (uint64_t)prev * ntile;
  const uint64_t target_ctas = (uint64_t)ctx->sms * 32;
  uint32_t nsplit = 1;
  if (base_jobs < target_ctas) {
    uint64_t want = (target_ctas + base_jobs - 1) / base_jobs;
    const uint64_t max_split = (axis + by * 8 - 1) / (by * 8);  // >= 8 rows per thread
    if (want > max_split) want = max_split;
    nsplit = (uint32_t)(want < 1 ? 1 : want);
  }
  const uint32_t seg = (axis + nsplit - 1) / nsplit;
  nsplit = seg ? (axis + seg - 1) / seg : 1;
  if (nsplit < 1) nsplit = 1;
  const uint64_t jobs = base_jobs * nsplit;
  VKP_CHECK(jobs < (1ull << 31), "reduction too large");
  const unsigned grid = (unsigned)jobs;
  float* dst = out;
  if (nsplit > 1) {
    void* ws;
    VKP_TRY(vkp_workspace(ctx, 0, (uint64_t)prev * nsplit * post * sizeof(float), &ws));
    dst = (float*)ws;
  }
  dim3 block(bx, by);
  if (vec)
    reduce_cols<OP, 4><<<grid, block, 0, ctx->stream>>>(in, dst, prev, axis, post, nsplit, seg);
  else
    reduce_cols<OP, 1><<<grid, block, 0, ctx->stream>>>(in, dst, prev, axis, post, nsplit, seg);
  VKP_TRY(vkp_after_launch(ctx, "reduce_cols"));
  if (nsplit > 1) {
    // second pass over [prev, nsplit, post]; nsplit is small so it never splits again.  With the
    // first pass's wide tiles it would be a handful of CTAs whose threads walk nsplit / by rows
    // one dependent load after the other (20 us for [256, 1024] on 4 CTAs); narrower tiles give
    // more CTAs and more threads along the (short) axis.
    uint32_t bx2 = bx;
    while (bx2 > 8 && (uint64_t)prev * ((post + bx2 * (vec ? 4 : 1) - 1) / (bx2 * (vec ? 4 : 1))) < 2ull * ctx->sms) bx2 >>= 1;
    const uint32_t ctile2 = bx2 * (vec ? 4 : 1);
    const unsigned grid2 = (unsigned)((uint64_t)prev * ((post + ctile2 - 1) / ctile2));
    dim3 block2(bx2, 256 / bx2);
    if (vec)
      reduce_cols<OP, 4><<<grid2, block2, 0, ctx->stream>>>(dst, out, prev, nsplit, post, 1, nsplit);
    else
      reduce_cols<OP, 1><<<grid2, block2, 0, ctx->stream>>>(dst, out, prev, nsplit, post, 1, nsplit);
    VKP_TRY(vkp_after_launch(ctx, "reduce_cols(pass2)"));
  }
  return VKP_OK;
}

template <int OP>
int reduce_dispatch(vkp_ctx* ctx, int fam, void* const* bufs, int nbuf, const void* params, size_t pbytes) {
  switch (fam) {
    case VKF_REDUCE: {  // A, B ; MultiVector<2>{sizeA, sizeB}
      VKP_CHECK(nbuf == 2 && pbytes == sizeof(vkp_multivector2_params), "reduce: bad arguments");
      const auto* p = static_cast<const vkp_multivector2_params*>(params);
      if (p->size[1] == 0) return VKP_OK;
      if (p->size[1] == 1) return reduce_rows<OP>(ctx, (const float*)bufs[0], (float*)bufs[1], 1, p->size[0]);
      const unsigned blocks = (p->size[1] * 32u + 255u) / 256u;
      reduce_strided<OP><<<blocks, 256, 0, ctx->stream>>>((const float*)bufs[0], (float*)bufs[1], p->size[0], p->size[1]);
      return vkp_after_launch(ctx, "reduce_strided");
    }
    case VKF_REDUCE_SG: {  // A, B ; Vector{size}: whole array -> B[0]
      VKP_CHECK(nbuf == 2 && pbytes == sizeof(vkp_vector_params), "reduce(v1.3): bad arguments");
      const auto* p = static_cast<const vkp_vector_params*>(params);
      return reduce_rows<OP>(ctx, (const float*)bufs[0], (float*)bufs[1], 1, p->size);
    }
    case VKF_REDUCE_AXIS: {  // A [prev,axis,post], B [prev,post]
      VKP_CHECK(nbuf == 2 && pbytes == sizeof(vkp_axisreduction_params), "axis reduce: bad arguments");
      const auto* p = static_cast<const vkp_axisreduction_params*>(params);
      return reduce_axis<OP>(ctx, (const float*)bufs[0], (float*)bufs[1], p->prev_prod, p->axis_size, p->post_prod);
    }
    case VKF_REDUCE_AXIS_RB: {  // A [prev,axis,post], B same shape
      VKP_CHECK(nbuf == 2 && pbytes == sizeof(vkp_axisreduction_params), "axis reduce: bad arguments");
      const auto* p = static_cast<const vkp_axisreduction_params*>(params);
      const uint64_t nout = (uint64_t)p->prev_prod * p->post_prod;
      if (nout == 0 || p->axis_size == 0) return VKP_OK;
      void* ws;  // reduced values; slot 1 (slot 0 may hold split partials)
      VKP_TRY(vkp_workspace(ctx, 1, nout * sizeof(float), &ws));
      float* red = static_cast<float*>(ws);
      VKP_TRY(reduce_axis<OP>(ctx, (const float*)bufs[0], red, p->prev_prod, p->axis_size, p->post_prod));
      return vkp_broadcast_copy_3d(ctx, red, (float*)bufs[1], p->prev_prod, p->axis_size, p->post_prod);
    }
  }
  return vkp_set_error("vkp_launch_reduce: unknown family %d", fam);
}

}  // namespace

// [prev, axis, post] -> [prev, post] into any 16-byte aligned device pointer (vkp_comm.cu reduces a
// shard's partials straight into the peer mailbox slot the exchange kernel reads)
int vkp_reduce_axis_into(vkp_ctx* ctx, int op, const float* in, float* out, uint32_t prev, uint32_t axis, uint32_t post) {
  switch (op) {
    case VKR_SUM: return reduce_axis<VKR_SUM>(ctx, in, out, prev, axis, post);
    case VKR_PROD: return reduce_axis<VKR_PROD>(ctx, in, out, prev, axis, post);
    case VKR_MAX: return reduce_axis<VKR_MAX>(ctx, in, out, prev, axis, post);
    case VKR_MIN: return reduce_axis<VKR_MIN>(ctx, in, out, prev, axis, post);
  }
  return vkp_set_error("vkp_reduce_axis_into: unknown reduction %d", op);
}

int vkp_launch_reduce(vkp_ctx* ctx, int fam, int sub, void* const* bufs, int nbuf, const void* params,
                      size_t pbytes) {
  switch (sub) {
    case VKR_SUM: return reduce_dispatch<VKR_SUM>(ctx, fam, bufs, nbuf, params, pbytes);
    case VKR_PROD: return reduce_dispatch<VKR_PROD>(ctx, fam, bufs, nbuf, params, pbytes);
    case VKR_MAX: return reduce_dispatch<VKR_MAX>(ctx, fam, bufs, nbuf, params, pbytes);
    case VKR_MIN: return reduce_dispatch<VKR_MIN>(ctx, fam, bufs, nbuf, params, pbytes);
  }
  return vkp_set_error("vkp_launch_reduce: unknown reduction %d", sub);
}
