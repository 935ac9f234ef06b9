// vkp_broadcast.cu -- NumPy-rule broadcasting fused with the binary op, and broadcast_to.
//
// Replaces add/sub/mul/div/max/min/pow_broadcast.comp (shader/add_broadcast.comp:25-45),
// their in-place forms (shader/iadd_broadcast.comp:22-41) and broadcast.comp (:25-44).
// The reference walks all ndim dimensions with a div/mod per output ELEMENT and reads
// the shape triple from a storage buffer.  Here the host collapses the aligned shapes
// into the fewest dimensions in which each operand is either contiguous or repeated
// (a size-1 dimension has stride 0, which is exactly `min(d, dim-1)` of the shader),
// and the kernel de-linearises once per 16-byte output vector: the innermost collapsed
// dimension is read as float4 (contiguous operand) or one scalar splat (repeated operand).
#include "vkp_common.cuh"
#include "vkp_math.cuh"

namespace {

constexpr int BC_MAXD = 8;
constexpr int BC_BLOCK = 256;
constexpr int BC_UNROLL = 4;
constexpr int BC_TILE = BC_BLOCK * BC_UNROLL;

struct BcastDesc {
  uint32_t inner;            // length of the innermost collapsed dimension (elements)
  uint32_t ia, ib;           // inner stride of A / B: 1 contiguous, 0 repeated
  int nouter;                // number of outer dimensions, innermost-outer first
  uint32_t shape[BC_MAXD];
  uint32_t sa[BC_MAXD];      // element strides of A / B per outer dimension (0 = repeated)
  uint32_t sb[BC_MAXD];
};

struct BAdd { __device__ float operator()(float a, float b) const { return a + b; } };
struct BSub { __device__ float operator()(float a, float b) const { return a - b; } };
struct BMul { __device__ float operator()(float a, float b) const { return a * b; } };
struct BDiv { __device__ float operator()(float a, float b) const { return a / b; } };
struct BMax { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct BMin { __device__ float operator()(float a, float b) const { return fminf(a, b); } };
struct BPow { __device__ float operator()(float a, float b) const { return vkpm::pow_f(a, b); } };
struct BFirst { __device__ float operator()(float a, float) const { return a; } };  // broadcast.comp

__device__ __forceinline__ void outer_offsets(const BcastDesc& d, uint32_t o, uint32_t& offa,
                                              uint32_t& offb) {
  offa = 0;
  offb = 0;
  for (int k = 0; k < d.nouter; k++) {
    const uint32_t q = o / d.shape[k];
    const uint32_t r = o - q * d.shape[k];
    offa += r * d.sa[k];
    offb += r * d.sb[k];
    o = q;
  }
}

// inner % 4 == 0: one float4 of output per step
template <class F, bool HAS_B>
__global__ void __launch_bounds__(BC_BLOCK)
bcast_vec_kernel(F f, const __grid_constant__ BcastDesc d, const float* A, const float* B, float* C,
                 uint32_t nvec) {
  const uint32_t lv = d.inner >> 2;
  {  // one tile per CTA (profiles/r01_micro_stream_variants.txt)
    const uint32_t base = blockIdx.x * BC_TILE + threadIdx.x;
    float4 a[BC_UNROLL], b[BC_UNROLL];
#pragma unroll
    for (int u = 0; u < BC_UNROLL; u++) {
      const uint32_t v = base + u * BC_BLOCK;
      if (v < nvec) {
        const uint32_t o = v / lv;
        const uint32_t iv = v - o * lv;
        uint32_t offa, offb;
        outer_offsets(d, o, offa, offb);
        if (d.ia) {
          a[u] = *reinterpret_cast<const float4*>(A + offa + (size_t)iv * 4);
        } else {
          const float s = A[offa];
          a[u] = make_float4(s, s, s, s);
        }
        b[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (HAS_B) {
          if (d.ib) {
            b[u] = *reinterpret_cast<const float4*>(B + offb + (size_t)iv * 4);
          } else {
            const float s = B[offb];
            b[u] = make_float4(s, s, s, s);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < BC_UNROLL; u++) {
      const uint32_t v = base + u * BC_BLOCK;
      if (v < nvec) {
        float4 r;
        r.x = f(a[u].x, b[u].x);
        r.y = f(a[u].y, b[u].y);
        r.z = f(a[u].z, b[u].z);
        r.w = f(a[u].w, b[u].w);
        reinterpret_cast<float4*>(C)[v] = r;
      }
    }
  }
}

// any shape: one element per step
template <class F, bool HAS_B>
__global__ void __launch_bounds__(BC_BLOCK)
bcast_scalar_kernel(F f, const __grid_constant__ BcastDesc d, const float* A, const float* B, float* C,
                    uint32_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t o = (uint32_t)i / d.inner;
    const uint32_t in = (uint32_t)i - o * d.inner;
    uint32_t offa, offb;
    outer_offsets(d, o, offa, offb);
    const float a = A[offa + in * d.ia];
    const float b = HAS_B ? B[offb + in * d.ib] : 0.f;
    C[i] = f(a, b);
  }
}

template <class F, bool HAS_B>
int launch_bcast(vkp_ctx* ctx, const char* name, const BcastDesc& d, const void* A, const void* B,
                 void* C, uint32_t n) {
  if (n == 0) return VKP_OK;
  if (d.inner % 4 == 0) {
    const uint32_t nvec = n / 4;
    const unsigned grid = (nvec + BC_TILE - 1) / BC_TILE;
    bcast_vec_kernel<F, HAS_B><<<grid, BC_BLOCK, 0, ctx->stream>>>(F(), d, (const float*)A, (const float*)B,
                                                                   (float*)C, nvec);
  } else {
    const unsigned grid = vkp_grid_for(ctx, n, BC_BLOCK * 4, 8);
    bcast_scalar_kernel<F, HAS_B><<<grid, BC_BLOCK, 0, ctx->stream>>>(F(), d, (const float*)A,
                                                                      (const float*)B, (float*)C, n);
  }
  return vkp_after_launch(ctx, name);
}

// Collapse [ndim] aligned shapes (row-major, outermost first) into a BcastDesc.
// A dimension of an operand whose extent is 1 (while the output's is not) is repeated.
int build_desc(const uint32_t* shA, const uint32_t* shB, const uint32_t* shC, uint32_t ndim,
               BcastDesc* out) {
  VKP_CHECK(ndim >= 1 && ndim <= 32, "broadcast: ndim %u out of range", ndim);
  // per-dimension (innermost first): extent, stride of A, stride of B
  uint32_t ext[32], sa[32], sb[32];
  int nd = 0;
  uint64_t accA = 1, accB = 1;
  for (int k = (int)ndim - 1; k >= 0; k--) {
    const uint32_t c = shC[k], a = shA[k], b = shB ? shB[k] : 1;
    VKP_CHECK((a == c || a == 1) && (b == c || b == 1), "broadcast: incompatible extents at axis %d", k);
    const uint32_t stA = (a == 1) ? 0u : (uint32_t)accA;
    const uint32_t stB = (b == 1) ? 0u : (uint32_t)accB;
    accA *= a;
    accB *= b;
    if (c == 1) continue;  // contributes nothing
    if (nd > 0) {
      // mergeable with the previous (inner) dimension if both operands continue the same pattern
      const uint32_t pe = ext[nd - 1];
      const bool mA = (sa[nd - 1] == 0 && stA == 0) || (sa[nd - 1] != 0 && stA == sa[nd - 1] * pe);
      const bool mB = (sb[nd - 1] == 0 && stB == 0) || (sb[nd - 1] != 0 && stB == sb[nd - 1] * pe);
      if (mA && mB) {
        ext[nd - 1] = pe * c;
        continue;
      }
    }
    ext[nd] = c;
    sa[nd] = stA;
    sb[nd] = stB;
    nd++;
  }
  if (nd == 0) {  // everything is size 1
    ext[0] = 1; sa[0] = 0; sb[0] = 0; nd = 1;
  }
  VKP_CHECK(nd - 1 <= BC_MAXD, "broadcast: more than %d non-mergeable dimensions", BC_MAXD + 1);
  BcastDesc d;
  memset(&d, 0, sizeof(d));
  d.inner = ext[0];
  d.ia = sa[0] ? 1 : 0;   // innermost stride is 1 when not repeated
  d.ib = sb[0] ? 1 : 0;
  d.nouter = nd - 1;
  for (int k = 1; k < nd; k++) {
    d.shape[k - 1] = ext[k];
    d.sa[k - 1] = sa[k];
    d.sb[k - 1] = sb[k];
  }
  *out = d;
  return VKP_OK;
}

template <bool HAS_B>
int dispatch(vkp_ctx* ctx, int sub, const BcastDesc& d, const void* A, const void* B, void* C, uint32_t n);

template <>
int dispatch<true>(vkp_ctx* ctx, int sub, const BcastDesc& d, const void* A, const void* B, void* C, uint32_t n) {
  switch (sub) {
    case VKB_ADD: return launch_bcast<BAdd, true>(ctx, "add_broadcast", d, A, B, C, n);
    case VKB_SUB: return launch_bcast<BSub, true>(ctx, "sub_broadcast", d, A, B, C, n);
    case VKB_MUL: return launch_bcast<BMul, true>(ctx, "mul_broadcast", d, A, B, C, n);
    case VKB_DIV: return launch_bcast<BDiv, true>(ctx, "div_broadcast", d, A, B, C, n);
    case VKB_MAX: return launch_bcast<BMax, true>(ctx, "max_broadcast", d, A, B, C, n);
    case VKB_MIN: return launch_bcast<BMin, true>(ctx, "min_broadcast", d, A, B, C, n);
    case VKB_POW: return launch_bcast<BPow, true>(ctx, "pow_broadcast", d, A, B, C, n);
  }
  return vkp_set_error("unknown broadcast op %d", sub);
}

}  // namespace

// dst[i, k, j] = src[i, j] for k < axis: the write-back half of *_axis_rebroadcast.comp:32-34
int vkp_broadcast_copy_3d(vkp_ctx* ctx, const float* src, float* dst, uint32_t prev, uint32_t axis,
                          uint32_t post) {
  const uint32_t shA[3] = {prev, 1, post}, shC[3] = {prev, axis, post};
  BcastDesc d;
  VKP_TRY(build_desc(shA, nullptr, shC, 3, &d));
  return launch_bcast<BFirst, false>(ctx, "rebroadcast", d, src, nullptr, dst, prev * axis * post);
}

int vkp_launch_broadcast(vkp_ctx* ctx, int fam, int sub, void* const* bufs, int nbuf,
                         const void* params, size_t pbytes) {
  BcastDesc d;
  switch (fam) {
    case VKF_BCAST: {  // A, B, C, shapeABC (host u32[3*ndim])
      VKP_CHECK(nbuf == 4 && pbytes == sizeof(vkp_multi3broadcast_params), "broadcast op: bad arguments");
      const auto* p = static_cast<const vkp_multi3broadcast_params*>(params);
      const uint32_t* sh = static_cast<const uint32_t*>(bufs[3]);
      VKP_CHECK(sh != nullptr, "broadcast op: null shape binding");
      VKP_TRY(build_desc(sh, sh + p->ndim, sh + 2 * p->ndim, p->ndim, &d));
      return dispatch<true>(ctx, sub, d, bufs[0], bufs[1], bufs[2], p->size[2]);
    }
    case VKF_IBCAST: {  // A (rw), B, shapeAB (host u32[2*ndim])
      VKP_CHECK(nbuf == 3 && pbytes == sizeof(vkp_broadcast_params), "in-place broadcast op: bad arguments");
      const auto* p = static_cast<const vkp_broadcast_params*>(params);
      const uint32_t* sh = static_cast<const uint32_t*>(bufs[2]);
      VKP_CHECK(sh != nullptr, "in-place broadcast op: null shape binding");
      VKP_TRY(build_desc(sh, sh + p->ndim, sh, p->ndim, &d));
      return dispatch<true>(ctx, sub, d, bufs[0], bufs[1], bufs[0], p->size[0]);
    }
    case VKF_BCAST_COPY: {  // A, B (out), shapeA (host), shapeB (host)
      VKP_CHECK(nbuf == 4 && pbytes == sizeof(vkp_broadcast_params), "broadcast: bad arguments");
      const auto* p = static_cast<const vkp_broadcast_params*>(params);
      const uint32_t* shA = static_cast<const uint32_t*>(bufs[2]);
      const uint32_t* shB = static_cast<const uint32_t*>(bufs[3]);
      VKP_CHECK(shA && shB, "broadcast: null shape binding");
      VKP_TRY(build_desc(shA, nullptr, shB, p->ndim, &d));
      return launch_bcast<BFirst, false>(ctx, "broadcast", d, bufs[0], nullptr, bufs[1], p->size[1]);
    }
  }
  return vkp_set_error("vkp_launch_broadcast: unknown family %d", fam);
}
