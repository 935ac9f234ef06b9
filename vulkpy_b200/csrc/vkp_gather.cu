// vkp_gather.cu -- gather / gather_axis (one-hot goes through gather_axis).
//
// Replaces shader/gather.comp:21-26 (c[i] = a[b[i]], no bounds check) and
// shader/gather_axis.comp:24-43 (c[k,i,j] = a[i, clamp(b[k],0,axis_size), j]).
// Indices are uint32 and copied bit-exactly; the payload is moved, never recomputed.
// Flat gather: four indices are fetched with one 16-byte load, the four dependent 4-byte
// table reads are issued back to back (they are sector-bound random reads), and the
// result leaves as one 16-byte store.  Axis gather: threads run along `post`
// (contiguous in both the table and the output) so reads and writes stay coalesced.
#include "vkp_common.cuh"

namespace {

constexpr int GA_BLOCK = 256;
// 4 index vectors (16 dependent 4-byte reads) in flight per thread and a grid of 16 CTAs per SM: 1.085 ms for
// 2^26 random indices into 256 MiB against 1.155 ms at 2 / 8 (profiles/r02_gather_gran.txt).  The random gather
// is DRAM-bound, not latency-bound: ncu counts 6.26 GB read from DRAM per launch (93 B per 4-byte element, the
// memory system's fetch granularity; cudaLimitMaxL2FetchGranularity 32/64/128 makes no difference), i.e.
// 5.6 TB/s of DRAM traffic (profiles/r02_ncu_rows.md).
constexpr int GA_UNROLL = 4;

__global__ void __launch_bounds__(GA_BLOCK)
gather_kernel(const float* __restrict__ a, const uint32_t* __restrict__ idx, float* __restrict__ c, size_t n) {
  const size_t nvec = n >> 2;
  const uint4* iv = reinterpret_cast<const uint4*>(idx);
  float4* cv = reinterpret_cast<float4*>(c);
  const size_t stride = (size_t)gridDim.x * GA_BLOCK * GA_UNROLL;
  for (size_t base = (size_t)blockIdx.x * GA_BLOCK * GA_UNROLL + threadIdx.x; base < nvec; base += stride) {
    uint4 ix[GA_UNROLL];
    float4 r[GA_UNROLL];
#pragma unroll
    for (int u = 0; u < GA_UNROLL; u++) {
      const size_t v = base + (size_t)u * GA_BLOCK;
      if (v < nvec) ix[u] = iv[v];
    }
#pragma unroll
    for (int u = 0; u < GA_UNROLL; u++) {
      const size_t v = base + (size_t)u * GA_BLOCK;
      if (v < nvec) {
        r[u].x = __ldg(a + ix[u].x);
        r[u].y = __ldg(a + ix[u].y);
        r[u].z = __ldg(a + ix[u].z);
        r[u].w = __ldg(a + ix[u].w);
      }
    }
#pragma unroll
    for (int u = 0; u < GA_UNROLL; u++) {
      const size_t v = base + (size_t)u * GA_BLOCK;
      if (v < nvec) cv[v] = r[u];
    }
  }
  if (blockIdx.x == 0) {
    const size_t i = (nvec << 2) + threadIdx.x;
    if (i < n) c[i] = a[idx[i]];
  }
}

// a [prev, axis, post], idx [nidx], c [nidx, prev, post]
__global__ void __launch_bounds__(GA_BLOCK)
gather_axis_kernel(const float* __restrict__ a, const uint32_t* __restrict__ idx, float* __restrict__ c,
                   uint32_t prev, uint32_t post, uint32_t axis, uint32_t nidx) {
  const uint64_t total = (uint64_t)nidx * prev * post;
  const uint64_t pp = (uint64_t)prev * post;
  for (uint64_t o = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; o < total;
       o += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = o / pp;
    const uint64_t rem = o - k * pp;
    const uint64_t i = rem / post;
    const uint64_t j = rem - i * post;
    uint32_t bk = idx[k];
    // The shader clamps to [0, axis_size] (inclusive: gather_axis.comp:32), i.e. an index equal
    // to axis_size reads one row past the axis.  Parity is defined for in-range indices only
    // (SURVEY Q10); out-of-range ones are clamped to the last valid row so no read leaves A.
    bk = bk >= axis ? axis - 1 : bk;
    c[o] = a[(i * axis + bk) * post + j];
  }
}

}  // namespace

int vkp_launch_gather(vkp_ctx* ctx, int fam, int sub, void* const* bufs, int nbuf, const void* params,
                      size_t pbytes) {
  (void)sub;
  if (fam == VKF_GATHER) {  // A, B (indices), C
    VKP_CHECK(nbuf == 3 && pbytes == sizeof(vkp_vector_params), "gather: bad arguments");
    const auto* p = static_cast<const vkp_vector_params*>(params);
    if (p->size == 0) return VKP_OK;
    const unsigned grid = vkp_grid_for(ctx, (p->size + 3) / 4, GA_BLOCK * GA_UNROLL, 16);
    gather_kernel<<<grid, GA_BLOCK, 0, ctx->stream>>>((const float*)bufs[0], (const uint32_t*)bufs[1],
                                                      (float*)bufs[2], p->size);
    return vkp_after_launch(ctx, "gather");
  }
  if (fam == VKF_GATHER_AXIS) {
    VKP_CHECK(nbuf == 3 && pbytes == sizeof(vkp_axisgather_params), "gather_axis: bad arguments");
    const auto* p = static_cast<const vkp_axisgather_params*>(params);
    const uint64_t total = (uint64_t)p->index_size * p->prev_prod * p->post_prod;
    if (total == 0) return VKP_OK;
    const unsigned grid = vkp_grid_for(ctx, total, GA_BLOCK * 4, 8);
    gather_axis_kernel<<<grid, GA_BLOCK, 0, ctx->stream>>>((const float*)bufs[0], (const uint32_t*)bufs[1],
                                                           (float*)bufs[2], p->prev_prod, p->post_prod,
                                                           p->axis_size, p->index_size);
    return vkp_after_launch(ctx, "gather_axis");
  }
  return vkp_set_error("vkp_launch_gather: unknown family %d", fam);
}
