// vkp_elementwise.cu -- same-shape element-wise kernels (HBM-bound, 128-bit vectorised).
//
// Replaces the one-element-per-invocation shaders of the reference:
//   vec (+) vec      add/sub/mul/div/max/min/pow.comp          (shader/add.comp:21-26)
//   vec (+)= vec     iadd ... ipow.comp                         (shader/iadd.comp:18-23)
//   vec (+) scalar   *_scalar.comp, r*_scalar.comp              (shader/add_scalar.comp:19-24)
//   vec (+)= scalar  i*_scalar.comp                             (shader/iadd_scalar.comp:16-21)
//   unary            abs ... invsqrt.comp and i-forms           (shader/abs.comp:22)
//   clamp x8         clamp{,_sv,_vs,_ss}.comp, iclamp*.comp     (shader/clamp.comp:24-29)
//   nn_cross_entropy{,_backward}.comp                           (:25)
//   prng_box_muller.comp / prng_ibox_muller.comp / prng_randrange.comp
//
// Layout: flat contiguous float32.  Every thread moves EW_UNROLL independent 16-byte
// vectors per operand (coalesced: consecutive lanes -> consecutive float4), all loads
// issued before the first use so ~64 B per thread per operand are in flight; one 16 KiB
// tile per CTA and as many CTAs as tiles.
// Compiled with -fmad=false: one IEEE rounding per reference operation.
#include "vkp_common.cuh"
#include "vkp_math.cuh"
#include "vkp_tables.cuh"

namespace {

constexpr int EW_BLOCK = 256;
constexpr int EW_UNROLL = 4;
constexpr int EW_TILE_VEC = EW_BLOCK * EW_UNROLL;  // float4 per tile

// ---- functors ---------------------------------------------------------------------------
struct FAdd { __device__ float operator()(float a, float b) const { return a + b; } };
struct FSub { __device__ float operator()(float a, float b) const { return a - b; } };
struct FMul { __device__ float operator()(float a, float b) const { return a * b; } };
struct FDiv { __device__ float operator()(float a, float b) const { return a / b; } };
struct FMax { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct FMin { __device__ float operator()(float a, float b) const { return fminf(a, b); } };
struct FPow { __device__ float operator()(float a, float b) const { return vkpm::pow_f(a, b); } };

template <class B, bool REV>
struct FScalar {
  float s;
  __device__ float operator()(float a) const { return REV ? B()(s, a) : B()(a, s); }
};

struct USquare { __device__ float operator()(float a) const { return a * a; } };  // pow(x, 2.0), correctly rounded incl. ties
struct UAbs   { __device__ float operator()(float a) const { return fabsf(a); } };
struct USign  { __device__ float operator()(float a) const { return vkpm::sign_f(a); } };
struct USin   { __device__ float operator()(float a) const { return sinf(a); } };
struct UCos   { __device__ float operator()(float a) const { return cosf(a); } };
struct UTan   { __device__ float operator()(float a) const { return tanf(a); } };
struct UAsin  { __device__ float operator()(float a) const { return asinf(a); } };
struct UAcos  { __device__ float operator()(float a) const { return acosf(a); } };
struct UAtan  { __device__ float operator()(float a) const { return atanf(a); } };
struct USinh  { __device__ float operator()(float a) const { return sinhf(a); } };
struct UCosh  { __device__ float operator()(float a) const { return coshf(a); } };
struct UTanh  { __device__ float operator()(float a) const { return tanhf(a); } };
struct UAsinh { __device__ float operator()(float a) const { return vkpm::asinh_f(a); } };
struct UAcosh { __device__ float operator()(float a) const { return acoshf(a); } };
struct UAtanh { __device__ float operator()(float a) const { return atanhf(a); } };
struct UExp   { __device__ float operator()(float a) const { return vkpm::exp_f(a); } };
struct ULog   { __device__ float operator()(float a) const { return vkpm::log_f(a); } };
struct UExp2  { __device__ float operator()(float a) const { return vkpm::exp2_f(a); } };
struct ULog2  { __device__ float operator()(float a) const { return vkpm::log2_f(a); } };
struct USqrt  { __device__ float operator()(float a) const { return __fsqrt_rn(a); } };
struct UInvSqrt { __device__ float operator()(float a) const { return __fdiv_rn(1.0f, __fsqrt_rn(a)); } };

// GLSL clamp(x, lo, hi) = min(max(x, lo), hi)
struct CClampVV { __device__ float operator()(float a, float lo, float hi) const { return fminf(fmaxf(a, lo), hi); } };
struct CClampSV { float lo; __device__ float operator()(float a, float hi) const { return fminf(fmaxf(a, lo), hi); } };
struct CClampVS { float hi; __device__ float operator()(float a, float lo) const { return fminf(fmaxf(a, lo), hi); } };
struct CClampSS { float lo, hi; __device__ float operator()(float a) const { return fminf(fmaxf(a, lo), hi); } };

// L = -y * log(x + 1e-8)   (nn_cross_entropy.comp:25);  dx = -y / (x + 1e-8)  (..._backward.comp:25)
struct FCrossEntropy { __device__ float operator()(float x, float y) const { return (-y) * vkpm::log_f(x + 1e-8f); } };
struct FCrossEntropyBwd { __device__ float operator()(float x, float y) const { return (-y) / (x + 1e-8f); } };

// ---- generic vectorised kernel ------------------------------------------------------------
template <int NIN>
struct Apply;
template <> struct Apply<1> { template <class F> static __device__ float go(const F& f, float a, float, float) { return f(a); } };
template <> struct Apply<2> { template <class F> static __device__ float go(const F& f, float a, float b, float) { return f(a, b); } };
template <> struct Apply<3> { template <class F> static __device__ float go(const F& f, float a, float b, float c) { return f(a, b, c); } };

// One tile per CTA (no grid-stride loop): measured 5-15 % faster than a persistent grid of
// 148*k CTAs for streaming kernels on B200 (profiles/r01_micro_stream_variants.txt).
template <int NIN, class F>
__global__ void __launch_bounds__(EW_BLOCK)
ew_kernel(F f, const float* in0, const float* in1, const float* in2, float* out, size_t n) {
  const size_t nvec = n >> 2;
  const float4* v0 = reinterpret_cast<const float4*>(in0);
  const float4* v1 = reinterpret_cast<const float4*>(in1);
  const float4* v2 = reinterpret_cast<const float4*>(in2);
  float4* vo = reinterpret_cast<float4*>(out);
  const size_t base = (size_t)blockIdx.x * EW_TILE_VEC + threadIdx.x;
  float4 a[EW_UNROLL], b[EW_UNROLL], c[EW_UNROLL];
  if (base + (EW_UNROLL - 1) * EW_BLOCK < nvec) {  // whole tile in range for this thread
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
      const size_t i = base + (size_t)u * EW_BLOCK;
      a[u] = v0[i];
      if (NIN > 1) b[u] = v1[i];
      if (NIN > 2) c[u] = v2[i];
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
      float4 r;
      r.x = Apply<NIN>::go(f, a[u].x, b[u].x, c[u].x);
      r.y = Apply<NIN>::go(f, a[u].y, b[u].y, c[u].y);
      r.z = Apply<NIN>::go(f, a[u].z, b[u].z, c[u].z);
      r.w = Apply<NIN>::go(f, a[u].w, b[u].w, c[u].w);
      vo[base + (size_t)u * EW_BLOCK] = r;
    }
  } else {
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
      const size_t i = base + (size_t)u * EW_BLOCK;
      if (i < nvec) {
        a[u] = v0[i];
        if (NIN > 1) b[u] = v1[i];
        if (NIN > 2) c[u] = v2[i];
      }
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
      const size_t i = base + (size_t)u * EW_BLOCK;
      if (i < nvec) {
        float4 r;
        r.x = Apply<NIN>::go(f, a[u].x, b[u].x, c[u].x);
        r.y = Apply<NIN>::go(f, a[u].y, b[u].y, c[u].y);
        r.z = Apply<NIN>::go(f, a[u].z, b[u].z, c[u].z);
        r.w = Apply<NIN>::go(f, a[u].w, b[u].w, c[u].w);
        vo[i] = r;
      }
    }
  }
  // scalar tail (n % 4 elements)
  if (blockIdx.x == gridDim.x - 1) {
    const size_t i = (nvec << 2) + threadIdx.x;
    if (i < n) {
      const float a1 = in0[i];
      const float b1 = NIN > 1 ? in1[i] : 0.f;
      const float c1 = NIN > 2 ? in2[i] : 0.f;
      out[i] = Apply<NIN>::go(f, a1, b1, c1);
    }
  }
}

template <int NIN, class F>
int launch_ew(vkp_ctx* ctx, const char* name, F f, const void* in0, const void* in1,
              const void* in2, void* out, size_t n) {
  if (n == 0) return VKP_OK;
  const size_t nvec = n >> 2;
  const size_t tiles = (nvec + EW_TILE_VEC - 1) / EW_TILE_VEC;
  const unsigned grid = (unsigned)(tiles ? tiles : 1);
  ew_kernel<NIN, F><<<grid, EW_BLOCK, 0, ctx->stream>>>(
      f, (const float*)in0, (const float*)in1, (const float*)in2, (float*)out, n);
  return vkp_after_launch(ctx, name);
}

// ---- table-driven transcendental kernels (exp, exp2, pow) -----------------------------------
// The 32-entry tables of vkp_math.cuh live one entry per lane in registers; a lookup is a warp
// shuffle, so every lane of a warp evaluates the functor together: out-of-range lanes compute on
// dummy inputs and only their loads / stores are predicated off.
using vkpt::LaneTables;

// fast(): warp-collective, branch-free, ORs `special` for inputs it cannot handle; slow(): the
// careful scalar routine, only run for those inputs
struct TPow {
  __device__ float fast(const LaneTables& t, float a, float b, bool& sp) const { return vkpm::pow_core(a, b, t, sp); }
  __device__ float slow(float a, float b) const { return vkpm::pow_f(a, b); }
};
struct TExp {
  __device__ float fast(const LaneTables& t, float a, float, bool& sp) const { return vkpm::exp_core(a, t, sp); }
  __device__ float slow(float a, float) const { return vkpm::exp_f(a); }
};
struct TExp2 {
  __device__ float fast(const LaneTables& t, float a, float, bool& sp) const { return vkpm::exp2_core(a, t, sp); }
  __device__ float slow(float a, float) const { return vkpm::exp2_f(a); }
};
struct TLog {
  __device__ float fast(const LaneTables& t, float a, float, bool& sp) const { return vkpm::log_core(a, t, sp); }
  __device__ float slow(float a, float) const { return vkpm::log_f(a); }
};
struct TLog2 {
  __device__ float fast(const LaneTables& t, float a, float, bool& sp) const { return vkpm::log2_core(a, t, sp); }
  __device__ float slow(float a, float) const { return vkpm::log2_f(a); }
};
// L = -y * log(x + 1e-8)   (nn_cross_entropy.comp:25)
struct TCrossEntropy {
  __device__ float fast(const LaneTables& t, float x, float y, bool& sp) const { return (-y) * vkpm::log_core(x + 1e-8f, t, sp); }
  __device__ float slow(float x, float y) const { return (-y) * vkpm::log_f(x + 1e-8f); }
};
template <bool REV>
struct TPowScalar {
  float s;
  __device__ float fast(const LaneTables& t, float a, float, bool& sp) const {
    return REV ? vkpm::pow_core(s, a, t, sp) : vkpm::pow_core(a, s, t, sp);
  }
  __device__ float slow(float a, float) const { return REV ? vkpm::pow_f(s, a) : vkpm::pow_f(a, s); }
};
// a ** s, vector form: the domain tests leave the per-element path.  |s| < 2^19 is a per-thread test (it bounds
// |s log2 a| < 2^26, so k = rint(32 s log2 a) is exact); the bases must be positive normal floats and
// -126*32 <= k < 128*32 (the result is then a normal float: no flush needed) -- both tested through the
// min / max over the four elements.  Anything else re-evaluates the vector with pow_f.
struct TPowScalarV {
  float s;
  __device__ float fast(const LaneTables& t, float a, float, bool& sp) const { return vkpm::pow_core(a, s, t, sp); }
  __device__ float slow(float a, float) const { return vkpm::pow_f(a, s); }
  __device__ float4 fast4(const LaneTables& t, float4 a, float4, bool& sp) const {
    const double sd = (double)s;
    int k0, k1, k2, k3;
    float4 r;
    r.x = vkpm::pow_core_nc(a.x, sd, t, k0);
    r.y = vkpm::pow_core_nc(a.y, sd, t, k1);
    r.z = vkpm::pow_core_nc(a.z, sd, t, k2);
    r.w = vkpm::pow_core_nc(a.w, sd, t, k3);
    const uint32_t u0 = __float_as_uint(a.x), u1 = __float_as_uint(a.y), u2 = __float_as_uint(a.z),
                   u3 = __float_as_uint(a.w);
    const uint32_t umin = min(min(u0, u1), min(u2, u3)), umax = max(max(u0, u1), max(u2, u3));
    const int kmin = min(min(k0, k1), min(k2, k3)), kmax = max(max(k0, k1), max(k2, k3));
    sp |= (umin < 0x00800000u) | (umax >= 0x7f800000u) | (kmin < -126 * 32) | (kmax >= 128 * 32) |
          !(fabsf(s) < 524288.0f);
    return r;
  }
};

template <class F>
__device__ __noinline__ float4 redo_slow(const F f, float4 a, float4 b) {
  // re-evaluates a whole vector with the careful routine (bit-identical on ordinary inputs)
  return make_float4(f.slow(a.x, b.x), f.slow(a.y, b.y), f.slow(a.z, b.z), f.slow(a.w, b.w));
}

// minimum resident CTAs per SM (register cap): the two-input pow kernel is fastest at 48 registers
// (profiles/r02_pow_variants.txt: 84 -> 48 registers, 5610 -> 5830 GB/s), the others are left to the compiler
template <class F> struct TabMinBlocks { static constexpr int v = 1; };
template <> struct TabMinBlocks<TPow> { static constexpr int v = 5; };

// one vector (four elements) through the fast path; functors with a vector form (fast4) test their domain
// with min / max over the four elements instead of per element
template <class F, class = void> struct HasFast4 { static constexpr bool v = false; };
template <class F> struct HasFast4<F, decltype((void)&F::fast4)> { static constexpr bool v = true; };
template <class F>
__device__ __forceinline__ float4 tab_eval4(const F& f, const LaneTables& tab, float4 a, float4 b, bool& sp) {
  if constexpr (HasFast4<F>::v) {
    return f.fast4(tab, a, b, sp);
  } else {
    float4 r;
    r.x = f.fast(tab, a.x, b.x, sp);
    r.y = f.fast(tab, a.y, b.y, sp);
    r.z = f.fast(tab, a.z, b.z, sp);
    r.w = f.fast(tab, a.w, b.w, sp);
    return r;
  }
}

// FULL: CTA `blockIdx.x` owns one whole tile -- no bounds test anywhere, all loads issued before the first
// use.  (Tried in round 2: a software-pipelined persistent form -- resident CTAs walking over half-tiles with the
// next half-tile's loads in flight during the evaluation -- because ncu shows the pow kernel's warps mostly on
// the long scoreboard: +1.5 % for a**b at 32 CTAs per SM, -2 ... -25 % for the single-input kernels that already
// run at the HBM rate, profiles/r02_tab_pipe.txt.  Not kept.)  (One kernel with a CTA-uniform "whole tile?" branch compiled to predicated loads and a serialised
// evaluation: 0.58 ms against 0.48 ms for a**2.7 on 2^28 elements, scripts/micro/pow_variants.cu.)
// !FULL: ONE CTA for what is left after the whole tiles: the partial tile, predicated, plus the n % 4 tail.
template <int NIN, class F, bool FULL>
__global__ void __launch_bounds__(EW_BLOCK, TabMinBlocks<F>::v)
ew_tab_kernel(F f, const __grid_constant__ vkpm::MathCoef coef, const float* in0, const float* in1, float* out,
              size_t n, size_t first_tile) {
  const LaneTables tab(coef);
  const size_t nvec = n >> 2;
  const float4* v0 = reinterpret_cast<const float4*>(in0);
  const float4* v1 = reinterpret_cast<const float4*>(in1);
  float4* vo = reinterpret_cast<float4*>(out);
  const size_t base = (first_tile + blockIdx.x) * EW_TILE_VEC + threadIdx.x;
  float4 a[EW_UNROLL], b[EW_UNROLL];
  if (FULL) {
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
      a[u] = v0[base + (size_t)u * EW_BLOCK];
      if (NIN > 1) b[u] = v1[base + (size_t)u * EW_BLOCK];
    }
#pragma unroll
    for (int u = 0; u < EW_UNROLL; u++) {
      bool sp = false;
      float4 r = tab_eval4(f, tab, a[u], b[u], sp);
      if (sp) r = redo_slow(f, a[u], b[u]);   // rare: zero / negative / denormal / inf / nan / overflow
      vo[base + (size_t)u * EW_BLOCK] = r;
    }
    return;
  }
  // every lane of a warp evaluates the functor together (lookups are shuffles): out-of-range lanes compute on
  // dummy inputs and only their loads / stores are predicated off
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
#pragma unroll
  for (int u = 0; u < EW_UNROLL; u++) {
    const size_t i = base + (size_t)u * EW_BLOCK;
    a[u] = one;
    b[u] = one;
    if (i < nvec) {
      a[u] = v0[i];
      if (NIN > 1) b[u] = v1[i];
    }
  }
#pragma unroll
  for (int u = 0; u < EW_UNROLL; u++) {
    const size_t i = base + (size_t)u * EW_BLOCK;
    bool sp = false;
    float4 r = tab_eval4(f, tab, a[u], b[u], sp);
    if (sp) r = redo_slow(f, a[u], b[u]);
    if (i < nvec) vo[i] = r;
  }
  if (threadIdx.x < 32) {   // n % 4 tail, whole first warp participates
    const size_t i = (nvec << 2) + threadIdx.x;
    const float a1 = i < n ? in0[i] : 1.f;
    const float b1 = (NIN > 1 && i < n) ? in1[i] : 1.f;
    bool sp = false;
    float r = f.fast(tab, a1, b1, sp);
    if (sp) r = f.slow(a1, b1);
    if (i < n) out[i] = r;
  }
}

template <int NIN, class F>
int launch_ew_tab(vkp_ctx* ctx, const char* name, F f, const void* in0, const void* in1, void* out, size_t n) {
  if (n == 0) return VKP_OK;
  const size_t nvec = n >> 2;
  const size_t full_tiles = nvec / EW_TILE_VEC;
  const vkpm::MathCoef& coef = vkpt::host_coef();
  if (full_tiles) {
    ew_tab_kernel<NIN, F, true><<<(unsigned)full_tiles, EW_BLOCK, 0, ctx->stream>>>(
        f, coef, (const float*)in0, (const float*)in1, (float*)out, n, 0);
    VKP_TRY(vkp_after_launch(ctx, name));
  }
  if (full_tiles * EW_TILE_VEC * 4 < n) {
    ew_tab_kernel<NIN, F, false><<<1, EW_BLOCK, 0, ctx->stream>>>(
        f, coef, (const float*)in0, (const float*)in1, (float*)out, n, full_tiles);
    VKP_TRY(vkp_after_launch(ctx, name));
  }
  return VKP_OK;
}

// ---- a ** s with a launch-constant exponent: binomial-series kernel (vkp_math.cuh, pows_core) ----------
// Tables (rc_i^-s x 32, 2^(s e) x 256) are built on the device once per distinct s (pows_build_kernel,
// cached per device: same stream, so a rebuild is ordered after the last reader).  rc_i and rc_i^-s live one
// entry per lane (lookups are shuffles: every lane of a warp evaluates together), 2^(s e) in shared memory.
constexpr int POWS_SLOTS = 4;
constexpr size_t POWS_TAB_DOUBLES = 32 + 256;
constexpr size_t POWS_MIN_N = (size_t)1 << 15;   // below: the general kernel (one launch, no tables)

struct PowsLaneTables {
  const char* es_;           // shared memory, 256 doubles
  uint32_t rc_;              // high word of this lane's rc entry (the low word is zero)
  double cs_;                // this lane's rc^-s
  const vkpm::PowsCoef& c;
  __device__ double rc(uint32_t d) const { return __hiloint2double((int)vkpt::LaneTables::shfl5(rc_, d >> 18), 0); }
  __device__ double cs(uint32_t d) const {
    return __hiloint2double((int)vkpt::LaneTables::shfl5((uint32_t)__double2hiint(cs_), d >> 18),
                            (int)vkpt::LaneTables::shfl5((uint32_t)__double2loint(cs_), d >> 18));
  }
  __device__ double es(uint32_t d) const { return *reinterpret_cast<const double*>(es_ + ((d >> 20) & 0x7f8u)); }
  __device__ double b(int k) const { return c.b[k]; }
};

__global__ void pows_build_kernel(double* tab, float s) {
  const uint32_t t = threadIdx.x;             // 256 threads
  const double sd = (double)s;
  if (t < 32) tab[t] = pow(vkpt::g_tab_rc[t], -sd);
  tab[32 + t] = exp2(sd * (double)vkpm::pows_slot_exponent(t));   // s e is exact in binary64
}

__device__ __noinline__ float4 pows_slow4(float4 a, float s) {
  return make_float4(vkpm::pow_f(a.x, s), vkpm::pow_f(a.y, s), vkpm::pow_f(a.z, s), vkpm::pow_f(a.w, s));
}
__device__ __noinline__ float pows_slow1(float x, float s) { return vkpm::pow_f(x, s); }

// all lanes of a warp call this together (shuffles); vectors with a zero / subnormal / negative / inf / nan
// element are re-evaluated with pow_f afterwards (rare)
template <int D>
__device__ __forceinline__ float4 pows_eval4(const PowsLaneTables& t, float s, float4 a) {
  const uint32_t u0 = __float_as_uint(a.x), u1 = __float_as_uint(a.y), u2 = __float_as_uint(a.z),
                 u3 = __float_as_uint(a.w);
  const uint32_t umin = min(min(u0, u1), min(u2, u3)), umax = max(max(u0, u1), max(u2, u3));
  float4 r = make_float4(vkpm::pows_core<D>(u0, t), vkpm::pows_core<D>(u1, t), vkpm::pows_core<D>(u2, t),
                         vkpm::pows_core<D>(u3, t));
  if ((umin < 0x00800000u) | (umax >= 0x7f800000u)) r = pows_slow4(a, s);
  return r;
}

template <int D, bool FULL>
__global__ void __launch_bounds__(EW_BLOCK)
ew_pows_kernel(const double* __restrict__ tab, const __grid_constant__ vkpm::PowsCoef coef, float s,
               const float* in0, float* out, size_t n, size_t first_tile) {
  __shared__ __align__(16) double s_es[256];
  const size_t nvec = n >> 2;
  const float4* v0 = reinterpret_cast<const float4*>(in0);
  float4* vo = reinterpret_cast<float4*>(out);
  const size_t base = (first_tile + blockIdx.x) * EW_TILE_VEC + threadIdx.x;
  const double cs_lane = tab[threadIdx.x & 31];
  double2 tv = make_double2(0.0, 0.0);
  if (threadIdx.x < 128) tv = reinterpret_cast<const double2*>(tab + 32)[threadIdx.x];
  float4 a[EW_UNROLL];
#pragma unroll
  for (int u = 0; u < EW_UNROLL; u++) {
    const size_t i = base + (size_t)u * EW_BLOCK;
    if (FULL || i < nvec) a[u] = v0[i];
    else a[u] = make_float4(1.f, 1.f, 1.f, 1.f);
  }
  if (threadIdx.x < 128) reinterpret_cast<double2*>(s_es)[threadIdx.x] = tv;
  __syncthreads();
  vkpm::PowsCoef c;
#pragma unroll
  for (int k = 0; k <= D; k++) c.b[k] = vkpt::pin(coef.b[k]);
  const PowsLaneTables t{reinterpret_cast<const char*>(s_es),
                         (uint32_t)__double2hiint(vkpt::g_tab_rc[threadIdx.x & 31]), cs_lane, c};
#pragma unroll
  for (int u = 0; u < EW_UNROLL; u++) {
    const size_t i = base + (size_t)u * EW_BLOCK;
    const float4 r = pows_eval4<D>(t, s, a[u]);
    if (FULL || i < nvec) vo[i] = r;
  }
  if (!FULL && threadIdx.x < 32) {   // n % 4 tail, whole first warp participates
    const size_t i = (nvec << 2) + threadIdx.x;
    const float x = i < n ? in0[i] : 1.f;
    const uint32_t u = __float_as_uint(x);
    float r = vkpm::pows_core<D>(u, t);
    if (!((u - 0x00800000u) < 0x7f000000u)) r = pows_slow1(x, s);
    if (i < n) out[i] = r;
  }
}

struct PowsCache {
  double* tab[POWS_SLOTS] = {};
  uint32_t key[POWS_SLOTS] = {};
  uint64_t stamp[POWS_SLOTS] = {};
  uint64_t clock = 0;
};
std::mutex g_pows_mu;
PowsCache g_pows[64];   // per device (one context = one stream per device)

template <int D>
int launch_pows_d(vkp_ctx* ctx, const double* tab, const vkpm::PowsCoef& coef, float s, const void* in0, void* out,
                  size_t n) {
  const size_t nvec = n >> 2;
  const size_t full_tiles = nvec / EW_TILE_VEC;
  if (full_tiles) {
    ew_pows_kernel<D, true><<<(unsigned)full_tiles, EW_BLOCK, 0, ctx->stream>>>(tab, coef, s, (const float*)in0,
                                                                                (float*)out, n, 0);
    VKP_TRY(vkp_after_launch(ctx, "pow_scalar"));
  }
  if (full_tiles * EW_TILE_VEC * 4 < n) {
    ew_pows_kernel<D, false><<<1, EW_BLOCK, 0, ctx->stream>>>(tab, coef, s, (const float*)in0, (float*)out, n,
                                                              full_tiles);
    VKP_TRY(vkp_after_launch(ctx, "pow_scalar"));
  }
  return VKP_OK;
}

// returns VKP_OK and sets *done when the binomial kernel took the operation
int launch_pows(vkp_ctx* ctx, float s, const void* in0, void* out, size_t n, bool* done) {
  *done = false;
  static const bool off = getenv("VKP_POWS_BINOMIAL") && atoi(getenv("VKP_POWS_BINOMIAL")) == 0;
  if (off || n < POWS_MIN_N) return VKP_OK;
  vkpm::PowsCoef coef;
  const int D = vkpm::pows_plan(s, coef);
  if (!D) return VKP_OK;
  uint32_t key;
  memcpy(&key, &s, 4);
  const double* tab = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_pows_mu);
    PowsCache& pc = g_pows[ctx->device & 63];
    int slot = -1, lru = 0;
    for (int i = 0; i < POWS_SLOTS; i++) {
      if (pc.tab[i] && pc.key[i] == key) slot = i;
      if (pc.stamp[i] < pc.stamp[lru]) lru = i;
    }
    if (slot < 0) {
      slot = lru;
      if (!pc.tab[slot]) VKP_CUDA(cudaMalloc(&pc.tab[slot], POWS_TAB_DOUBLES * sizeof(double)));
      pows_build_kernel<<<1, 256, 0, ctx->stream>>>(pc.tab[slot], s);
      VKP_TRY(vkp_after_launch(ctx, "pow_scalar(tables)"));
      pc.key[slot] = key;
    }
    pc.stamp[slot] = ++pc.clock;
    tab = pc.tab[slot];
  }
  *done = true;
  switch (D) {
    case 6: return launch_pows_d<6>(ctx, tab, coef, s, in0, out, n);
    case 8: return launch_pows_d<8>(ctx, tab, coef, s, in0, out, n);
    default: return launch_pows_d<10>(ctx, tab, coef, s, in0, out, n);
  }
}

// ---- fused element-wise chains (lazy element-wise fusion) -------------------
// A chain is what the reference issues as a sequence of same-shape element-wise jobs whose
// intermediate arrays nobody else reads, e.g. Sigmoid.forward (nn/layers.py:239-243):
//     y = 0.0 - x;  y.exp(inplace);  y += 1.0;  y = 1.0 / y        4 jobs, 4 x 8 B per element
// Here it is ONE launch that moves 8 B per element: a running value `acc` starts as input 0 and
// every step applies one reference operation to it -- the very functor the stand-alone kernel of
// that shader uses, so each step rounds exactly once like the job it replaces and the result is
// bit-identical to the op-by-op sequence.  The second operand of a step is a scalar, another
// input array (up to 3 more, same shape) or `tmp`, a copy of `acc` saved by an earlier step
// (Huber: d.min(d ** 2.0)).  The program is a kernel parameter; all threads run the same step at
// the same time, so the warp-collective table lookups of exp / log / pow stay legal.
enum {
  CH_ADD = 0, CH_SUB, CH_MUL, CH_DIV, CH_MAX, CH_MIN, CH_POW,      // acc = acc op b
  CH_RSUB, CH_RDIV, CH_RPOW,                                      // acc = b op acc
  CH_SQUARE,                                                      // pow(acc, 2.0) (see dispatch_scalar)
  CH_UNARY0,                                                      // + VKU_* : acc = f(acc)
  CH_SAVE = CH_UNARY0 + VKU_COUNT,                                // tmp = acc
  CH_COUNT
};
enum { CH_SRC_SCALAR = 0, CH_SRC_IN1 = 1, CH_SRC_IN2 = 2, CH_SRC_IN3 = 3, CH_SRC_TMP = 4 };
constexpr int CH_MAX_STEPS = 16;
constexpr int CH_UNROLL = 2;
constexpr int CH_TILE_VEC = EW_BLOCK * CH_UNROLL;

struct ChainStep { uint8_t op, src, pad0, pad1; float s; };
struct ChainProg { ChainStep st[CH_MAX_STEPS]; int n; int nin; };

template <class B>
__device__ __forceinline__ float4 ch_bin(float4 a, float4 b) {
  return make_float4(B()(a.x, b.x), B()(a.y, b.y), B()(a.z, b.z), B()(a.w, b.w));
}
template <class U>
__device__ __forceinline__ float4 ch_un(float4 a) {
  return make_float4(U()(a.x), U()(a.y), U()(a.z), U()(a.w));
}
template <class T>
__device__ __forceinline__ float4 ch_tab(const LaneTables& tab, const T f, float4 a, float4 b) {
  bool sp = false;
  float4 r;
  r.x = f.fast(tab, a.x, b.x, sp);
  r.y = f.fast(tab, a.y, b.y, sp);
  r.z = f.fast(tab, a.z, b.z, sp);
  r.w = f.fast(tab, a.w, b.w, sp);
  if (sp) r = redo_slow(f, a, b);
  return r;
}

template <int NV>   // NV float4 per thread, all in registers
__device__ __forceinline__ void ch_step(const LaneTables& tab, const ChainStep st, float4 (&acc)[NV], float4 (&tmp)[NV],
                                        const float4 (&i1)[NV], const float4 (&i2)[NV], const float4 (&i3)[NV]) {
  float4 b[NV];
#pragma unroll
  for (int u = 0; u < NV; u++) {
    switch (st.src) {
      case CH_SRC_IN1: b[u] = i1[u]; break;
      case CH_SRC_IN2: b[u] = i2[u]; break;
      case CH_SRC_IN3: b[u] = i3[u]; break;
      case CH_SRC_TMP: b[u] = tmp[u]; break;
      default: b[u] = make_float4(st.s, st.s, st.s, st.s);
    }
  }
#define CH_EACH(EXPR)                  \
  _Pragma("unroll") for (int u = 0; u < NV; u++) acc[u] = (EXPR); \
  break
  switch (st.op) {
    case CH_ADD: CH_EACH(ch_bin<FAdd>(acc[u], b[u]));
    case CH_SUB: CH_EACH(ch_bin<FSub>(acc[u], b[u]));
    case CH_MUL: CH_EACH(ch_bin<FMul>(acc[u], b[u]));
    case CH_DIV: CH_EACH(ch_bin<FDiv>(acc[u], b[u]));
    case CH_MAX: CH_EACH(ch_bin<FMax>(acc[u], b[u]));
    case CH_MIN: CH_EACH(ch_bin<FMin>(acc[u], b[u]));
    case CH_POW: CH_EACH(ch_tab(tab, TPow(), acc[u], b[u]));
    case CH_RSUB: CH_EACH(ch_bin<FSub>(b[u], acc[u]));
    case CH_RDIV: CH_EACH(ch_bin<FDiv>(b[u], acc[u]));
    case CH_RPOW: CH_EACH(ch_tab(tab, TPow(), b[u], acc[u]));
    case CH_SQUARE: CH_EACH(ch_un<USquare>(acc[u]));
    case CH_UNARY0 + VKU_ABS: CH_EACH(ch_un<UAbs>(acc[u]));
    case CH_UNARY0 + VKU_SIGN: CH_EACH(ch_un<USign>(acc[u]));
    case CH_UNARY0 + VKU_SIN: CH_EACH(ch_un<USin>(acc[u]));
    case CH_UNARY0 + VKU_COS: CH_EACH(ch_un<UCos>(acc[u]));
    case CH_UNARY0 + VKU_TAN: CH_EACH(ch_un<UTan>(acc[u]));
    case CH_UNARY0 + VKU_ASIN: CH_EACH(ch_un<UAsin>(acc[u]));
    case CH_UNARY0 + VKU_ACOS: CH_EACH(ch_un<UAcos>(acc[u]));
    case CH_UNARY0 + VKU_ATAN: CH_EACH(ch_un<UAtan>(acc[u]));
    case CH_UNARY0 + VKU_SINH: CH_EACH(ch_un<USinh>(acc[u]));
    case CH_UNARY0 + VKU_COSH: CH_EACH(ch_un<UCosh>(acc[u]));
    case CH_UNARY0 + VKU_TANH: CH_EACH(ch_un<UTanh>(acc[u]));
    case CH_UNARY0 + VKU_ASINH: CH_EACH(ch_un<UAsinh>(acc[u]));
    case CH_UNARY0 + VKU_ACOSH: CH_EACH(ch_un<UAcosh>(acc[u]));
    case CH_UNARY0 + VKU_ATANH: CH_EACH(ch_un<UAtanh>(acc[u]));
    case CH_UNARY0 + VKU_EXP: CH_EACH(ch_tab(tab, TExp(), acc[u], b[u]));
    case CH_UNARY0 + VKU_LOG: CH_EACH(ch_tab(tab, TLog(), acc[u], b[u]));
    case CH_UNARY0 + VKU_EXP2: CH_EACH(ch_tab(tab, TExp2(), acc[u], b[u]));
    case CH_UNARY0 + VKU_LOG2: CH_EACH(ch_tab(tab, TLog2(), acc[u], b[u]));
    case CH_UNARY0 + VKU_SQRT: CH_EACH(ch_un<USqrt>(acc[u]));
    case CH_UNARY0 + VKU_INVSQRT: CH_EACH(ch_un<UInvSqrt>(acc[u]));
    case CH_SAVE:
#pragma unroll
      for (int u = 0; u < NV; u++) tmp[u] = acc[u];
      break;
  }
#undef CH_EACH
}

__global__ void __launch_bounds__(EW_BLOCK)
ew_chain_kernel(const __grid_constant__ ChainProg prog, const __grid_constant__ vkpm::MathCoef coef, const float* in0,
                const float* in1, const float* in2, const float* in3, float* out, size_t n) {
  const LaneTables tab(coef);
  const size_t nvec = n >> 2;
  const size_t base = (size_t)blockIdx.x * CH_TILE_VEC + threadIdx.x;
  const float4 one = make_float4(1.f, 1.f, 1.f, 1.f);
  float4 acc[CH_UNROLL], tmp[CH_UNROLL], i1[CH_UNROLL], i2[CH_UNROLL], i3[CH_UNROLL];
#pragma unroll
  for (int u = 0; u < CH_UNROLL; u++) {
    const size_t i = base + (size_t)u * EW_BLOCK;
    const bool ok = i < nvec;
    acc[u] = ok ? reinterpret_cast<const float4*>(in0)[i] : one;
    i1[u] = (ok && prog.nin > 1) ? reinterpret_cast<const float4*>(in1)[i] : one;
    i2[u] = (ok && prog.nin > 2) ? reinterpret_cast<const float4*>(in2)[i] : one;
    i3[u] = (ok && prog.nin > 3) ? reinterpret_cast<const float4*>(in3)[i] : one;
    tmp[u] = one;
  }
  for (int k = 0; k < prog.n; k++) ch_step<CH_UNROLL>(tab, prog.st[k], acc, tmp, i1, i2, i3);
#pragma unroll
  for (int u = 0; u < CH_UNROLL; u++) {
    const size_t i = base + (size_t)u * EW_BLOCK;
    if (i < nvec) reinterpret_cast<float4*>(out)[i] = acc[u];
  }
  if (blockIdx.x == gridDim.x - 1 && threadIdx.x < 32) {   // n % 4 tail: the whole first warp walks the program
    const size_t i = (nvec << 2) + threadIdx.x;
    const bool ok = i < n;
    float4 a1[1] = {make_float4(ok ? in0[i] : 1.f, 1.f, 1.f, 1.f)}, t1[1] = {one};
    const float4 b1[1] = {make_float4((ok && prog.nin > 1) ? in1[i] : 1.f, 1.f, 1.f, 1.f)};
    const float4 b2[1] = {make_float4((ok && prog.nin > 2) ? in2[i] : 1.f, 1.f, 1.f, 1.f)};
    const float4 b3[1] = {make_float4((ok && prog.nin > 3) ? in3[i] : 1.f, 1.f, 1.f, 1.f)};
    for (int k = 0; k < prog.n; k++) ch_step<1>(tab, prog.st[k], a1, t1, b1, b2, b3);
    if (ok) out[i] = a1[0].x;
  }
}

// ---- Box-Muller (prng_box_muller.comp:19-32, prng_ibox_muller.comp:16-27) -----------------
// One thread per pair.  Unlike the reference dispatch (floor(n/2) invocations rounded up to a
// workgroup, random.py:106-121) the last element of an odd-length output is always written.
template <bool FAST>
__global__ void __launch_bounds__(256)
box_muller_kernel(const float* a, float* b, size_t n, float mean, float stddev) {
  const size_t npair = (n + 1) >> 1;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npair; i += (size_t)gridDim.x * blockDim.x) {
    const size_t j = 2 * i, k = j + 1;
    const float2 u = *reinterpret_cast<const float2*>(a + j);  // a has an even number of elements
    float o0, o1;
    vkpm::box_muller_pair<FAST>(u.x, u.y, mean, stddev, o0, o1);
    if (k < n) *reinterpret_cast<float2*>(b + j) = make_float2(o0, o1);
    else b[j] = o0;
  }
}

// b[i] = low + uint(float(high - low + 1) * a[i])   (prng_randrange.comp:20-27)
__global__ void __launch_bounds__(256)
randrange_kernel(const float* a, uint32_t* b, size_t n, uint32_t low, uint32_t high) {
  const float range = __uint2float_rn(high - low + 1u);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    b[i] = low + __float2uint_rz(range * a[i]);
  }
}

__global__ void __launch_bounds__(256) fill_u32_kernel(uint32_t* dst, size_t n, uint32_t v) {
  const size_t nvec = n >> 2;
  uint4* dv = reinterpret_cast<uint4*>(dst);
  const uint4 vv = make_uint4(v, v, v, v);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < nvec;
       i += (size_t)gridDim.x * blockDim.x)
    dv[i] = vv;
  if (blockIdx.x == 0) {
    const size_t i = (nvec << 2) + threadIdx.x;
    if (i < n) dst[i] = v;
  }
}

// one draw per lane (prng_xoshiro128pp_uint32.comp:26-43 / _float.comp:26-44); the stream API
// in vkp_prng.cu is the fast path, this is the literal single-dispatch form.
template <bool AS_FLOAT>
__global__ void xoshiro_step_kernel(uint32_t* state, uint32_t* out, uint32_t shift, uint32_t size) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= size) return;
  uint4 s = reinterpret_cast<uint4*>(state)[i];
  const uint32_t sum = s.x + s.w;
  const uint32_t result = ((sum << 7) | (sum >> 25)) + s.x;
  out[i + shift] = AS_FLOAT ? __float_as_uint(__uint_as_float((result >> 9) | 0x3f800000u) - 1.0f) : result;
  const uint32_t t = s.y << 9;
  s.z ^= s.x; s.w ^= s.y; s.y ^= s.z; s.x ^= s.w;
  s.z ^= t;
  s.w = (s.w << 11) | (s.w >> 21);
  reinterpret_cast<uint4*>(state)[i] = s;
}

template <class B>
int launch_scalar(vkp_ctx* ctx, const char* name, bool rev, float s, const void* a, void* out, size_t n) {
  if (rev) return launch_ew<1>(ctx, name, FScalar<B, true>{s}, a, nullptr, nullptr, out, n);
  return launch_ew<1>(ctx, name, FScalar<B, false>{s}, a, nullptr, nullptr, out, n);
}

int dispatch_binary(vkp_ctx* ctx, int sub, const void* a, const void* b, void* out, size_t n) {
  switch (sub) {
    case VKB_ADD: return launch_ew<2>(ctx, "add", FAdd(), a, b, nullptr, out, n);
    case VKB_SUB: return launch_ew<2>(ctx, "sub", FSub(), a, b, nullptr, out, n);
    case VKB_MUL: return launch_ew<2>(ctx, "mul", FMul(), a, b, nullptr, out, n);
    case VKB_DIV: return launch_ew<2>(ctx, "div", FDiv(), a, b, nullptr, out, n);
    case VKB_MAX: return launch_ew<2>(ctx, "max", FMax(), a, b, nullptr, out, n);
    case VKB_MIN: return launch_ew<2>(ctx, "min", FMin(), a, b, nullptr, out, n);
    case VKB_POW: return launch_ew_tab<2>(ctx, "pow", TPow(), a, b, out, n);
  }
  return vkp_set_error("unknown binary op %d", sub);
}

int dispatch_scalar(vkp_ctx* ctx, int sub, float s, const void* a, void* out, size_t n) {
  switch (sub) {
    case VKB_ADD: return launch_scalar<FAdd>(ctx, "add_scalar", false, s, a, out, n);
    case VKB_SUB: return launch_scalar<FSub>(ctx, "sub_scalar", false, s, a, out, n);
    case VKB_MUL: return launch_scalar<FMul>(ctx, "mul_scalar", false, s, a, out, n);
    case VKB_DIV: return launch_scalar<FDiv>(ctx, "div_scalar", false, s, a, out, n);
    case VKB_MAX: return launch_scalar<FMax>(ctx, "max_scalar", false, s, a, out, n);
    case VKB_MIN: return launch_scalar<FMin>(ctx, "min_scalar", false, s, a, out, n);
    case VKB_POW:
      // x ** 2.0 (MSELoss, Ridge, Adam, AdaGrad: nn/losses.py:294-296, nn/optimizers.py:131,239) is the
      // exactly rounded square, also on ties and for negative x
      if (s == 2.0f) return launch_ew<1>(ctx, "pow_scalar(2)", USquare(), a, nullptr, nullptr, out, n);
      {
        bool done = false;
        VKP_TRY(launch_pows(ctx, s, a, out, n, &done));
        if (done) return VKP_OK;
      }
      return launch_ew_tab<1>(ctx, "pow_scalar", TPowScalarV{s}, a, nullptr, out, n);
    case VKB_RSUB: return launch_scalar<FSub>(ctx, "rsub_scalar", true, s, a, out, n);
    case VKB_RDIV: return launch_scalar<FDiv>(ctx, "rdiv_scalar", true, s, a, out, n);
    case VKB_RPOW: return launch_ew_tab<1>(ctx, "rpow_scalar", TPowScalar<true>{s}, a, nullptr, out, n);
  }
  return vkp_set_error("unknown scalar op %d", sub);
}

int dispatch_unary(vkp_ctx* ctx, int sub, const void* a, void* out, size_t n) {
#define U(ID, FN, NAME) case ID: return launch_ew<1>(ctx, NAME, FN(), a, nullptr, nullptr, out, n);
  switch (sub) {
    U(VKU_ABS, UAbs, "abs") U(VKU_SIGN, USign, "sign") U(VKU_SIN, USin, "sin") U(VKU_COS, UCos, "cos")
    U(VKU_TAN, UTan, "tan") U(VKU_ASIN, UAsin, "asin") U(VKU_ACOS, UAcos, "acos") U(VKU_ATAN, UAtan, "atan")
    U(VKU_SINH, USinh, "sinh") U(VKU_COSH, UCosh, "cosh") U(VKU_TANH, UTanh, "tanh")
    U(VKU_ASINH, UAsinh, "asinh") U(VKU_ACOSH, UAcosh, "acosh") U(VKU_ATANH, UAtanh, "atanh")
    case VKU_EXP: return launch_ew_tab<1>(ctx, "exp", TExp(), a, nullptr, out, n);
    case VKU_LOG: return launch_ew_tab<1>(ctx, "log", TLog(), a, nullptr, out, n);
    case VKU_EXP2: return launch_ew_tab<1>(ctx, "exp2", TExp2(), a, nullptr, out, n);
    case VKU_LOG2: return launch_ew_tab<1>(ctx, "log2", TLog2(), a, nullptr, out, n);
   
    U(VKU_SQRT, USqrt, "sqrt") U(VKU_INVSQRT, UInvSqrt, "invsqrt")
  }
#undef U
  return vkp_set_error("unknown unary op %d", sub);
}

}  // namespace

#define NEED(nb, T)                                                                           \
  VKP_CHECK(nbuf == (nb) && pbytes == sizeof(T), "%s: expected %d buffers and %zu parameter " \
            "bytes, got %d and %zu", __func__, (nb), sizeof(T), nbuf, pbytes);                \
  const T* p = static_cast<const T*>(params)

int vkp_launch_elementwise(vkp_ctx* ctx, int fam, int sub, void* const* bufs, int nbuf,
                           const void* params, size_t pbytes) {
  switch (fam) {
    case VKF_BIN: {   // bindings A, B, C
      NEED(3, vkp_vector_params);
      return dispatch_binary(ctx, sub, bufs[0], bufs[1], bufs[2], p->size);
    }
    case VKF_IBIN: {  // bindings A (rw), B
      NEED(2, vkp_vector_params);
      return dispatch_binary(ctx, sub, bufs[0], bufs[1], bufs[0], p->size);
    }
    case VKF_SCALAR: {  // bindings A, B (out)
      NEED(2, vkp_vectorscalar_params);
      return dispatch_scalar(ctx, sub, p->scalar, bufs[0], bufs[1], p->size);
    }
    case VKF_ISCALAR: {  // binding A (rw)
      NEED(1, vkp_vectorscalar_params);
      return dispatch_scalar(ctx, sub, p->scalar, bufs[0], bufs[0], p->size);
    }
    case VKF_UNARY: {
      NEED(2, vkp_vector_params);
      return dispatch_unary(ctx, sub, bufs[0], bufs[1], p->size);
    }
    case VKF_IUNARY: {
      NEED(1, vkp_vector_params);
      return dispatch_unary(ctx, sub, bufs[0], bufs[0], p->size);
    }
    case VKF_CLAMP: {
      const bool inplace = sub & 1;
      const int variant = sub >> 1;
      switch (variant) {
        case VKC_VV: {  // clamp.comp: A, B=min, C=max, D / iclamp.comp: A, B, C
          NEED(inplace ? 3 : 4, vkp_vector_params);
          return launch_ew<3>(ctx, "clamp", CClampVV(), bufs[0], bufs[1], bufs[2],
                              inplace ? bufs[0] : bufs[3], p->size);
        }
        case VKC_SV: {  // clamp_sv.comp: A, B=max, C ; scalar = min
          NEED(inplace ? 2 : 3, vkp_vectorscalar_params);
          return launch_ew<2>(ctx, "clamp_sv", CClampSV{p->scalar}, bufs[0], bufs[1], nullptr,
                              inplace ? bufs[0] : bufs[2], p->size);
        }
        case VKC_VS: {  // clamp_vs.comp: A, B=min, C ; scalar = max
          NEED(inplace ? 2 : 3, vkp_vectorscalar_params);
          return launch_ew<2>(ctx, "clamp_vs", CClampVS{p->scalar}, bufs[0], bufs[1], nullptr,
                              inplace ? bufs[0] : bufs[2], p->size);
        }
        case VKC_SS: {  // clamp_ss.comp: A, B ; scalars = [min, max]
          NEED(inplace ? 1 : 2, vkp_vectorscalar2_params);
          return launch_ew<1>(ctx, "clamp_ss", CClampSS{p->scalar[0], p->scalar[1]}, bufs[0],
                              nullptr, nullptr, inplace ? bufs[0] : bufs[1], p->size);
        }
      }
      return vkp_set_error("unknown clamp variant %d", variant);
    }
    case VKF_CE: {  // X, Y, L
      NEED(3, vkp_vector_params);
      return launch_ew_tab<2>(ctx, "nn_cross_entropy", TCrossEntropy(), bufs[0], bufs[1], bufs[2], p->size);
    }
    case VKF_CE_BWD: {  // X, Y, dX
      NEED(3, vkp_vector_params);
      return launch_ew<2>(ctx, "nn_cross_entropy_backward", FCrossEntropyBwd(), bufs[0], bufs[1], nullptr, bufs[2], p->size);
    }
    case VKF_BOX_MULLER: {  // A (uniform, n rounded up to even), B (out, n)
      NEED(2, vkp_vectorscalar2_params);
      if (p->size == 0) return VKP_OK;
      const unsigned grid = vkp_grid_for(ctx, (p->size + 1) / 2, 256, 8);
      if (vkp_normal_precise())
        box_muller_kernel<false><<<grid, 256, 0, ctx->stream>>>((const float*)bufs[0], (float*)bufs[1], p->size,
                                                               p->scalar[0], p->scalar[1]);
      else
        box_muller_kernel<true><<<grid, 256, 0, ctx->stream>>>((const float*)bufs[0], (float*)bufs[1], p->size,
                                                              p->scalar[0], p->scalar[1]);
      return vkp_after_launch(ctx, "prng_box_muller");
    }
    case VKF_IBOX_MULLER: {  // A (rw, even n)
      NEED(1, vkp_vectorscalar2_params);
      if (p->size == 0) return VKP_OK;
      VKP_CHECK(p->size % 2 == 0, "prng_ibox_muller needs an even element count (random.py:109-115)");
      const unsigned grid = vkp_grid_for(ctx, p->size / 2, 256, 8);
      if (vkp_normal_precise())
        box_muller_kernel<false><<<grid, 256, 0, ctx->stream>>>((const float*)bufs[0], (float*)bufs[0], p->size,
                                                               p->scalar[0], p->scalar[1]);
      else
        box_muller_kernel<true><<<grid, 256, 0, ctx->stream>>>((const float*)bufs[0], (float*)bufs[0], p->size,
                                                              p->scalar[0], p->scalar[1]);
      return vkp_after_launch(ctx, "prng_ibox_muller");
    }
    case VKF_RANDRANGE: {  // A ([0,1) floats), B (u32 out)
      NEED(2, vkp_vectorrange_params);
      if (p->size == 0) return VKP_OK;
      const unsigned grid = vkp_grid_for(ctx, p->size, 256, 8);
      randrange_kernel<<<grid, 256, 0, ctx->stream>>>((const float*)bufs[0], (uint32_t*)bufs[1], p->size,
                                                       p->low, p->high);
      return vkp_after_launch(ctx, "prng_randrange");
    }
    case VKF_PRNG_U32:
    case VKF_PRNG_F32: {  // A = state (4 words per lane), B = out
      NEED(2, vkp_shiftvector_params);
      if (p->size == 0) return VKP_OK;
      const unsigned grid = (p->size + 63) / 64;
      if (fam == VKF_PRNG_F32)
        xoshiro_step_kernel<true><<<grid, 64, 0, ctx->stream>>>((uint32_t*)bufs[0], (uint32_t*)bufs[1], p->shift, p->size);
      else
        xoshiro_step_kernel<false><<<grid, 64, 0, ctx->stream>>>((uint32_t*)bufs[0], (uint32_t*)bufs[1], p->shift, p->size);
      return vkp_after_launch(ctx, "prng_xoshiro128pp");
    }
  }
  return vkp_set_error("vkp_launch_elementwise: unknown family %d", fam);
}

extern "C" int vkp_fill_u32(vkp_ctx* ctx, void* dst, size_t count, uint32_t bits, vkp_job** job) {
  VKP_RANGE("vkp_fill_u32");
  VKP_CHECK(ctx && (dst || count == 0), "vkp_fill_u32: null argument");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[1] = {dst};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 1));
  if (count) {
    const unsigned grid = vkp_grid_for(ctx, (count + 3) / 4, 256, 8);
    fill_u32_kernel<<<grid, 256, 0, ctx->stream>>>((uint32_t*)dst, count, bits);
    VKP_TRY(vkp_after_launch(ctx, "fill"));
  }
  return vkp_finish_op(ctx, job);
}

// One launch for a chain of same-shape element-wise operations (see ew_chain_kernel).  ops[k] is a
// VKP_CHAIN_* code, srcs[k] selects the second operand (0 scalar, 1..3 input k, 4 the saved copy),
// scalars[k] its value when it is a scalar.  `out` may alias any input.
extern "C" int vkp_ew_chain(vkp_ctx* ctx, int n_in, const float* const* in, float* out, size_t count, int n_steps,
                            const int* ops, const int* srcs, const float* scalars, vkp_job** job) {
  VKP_RANGE("vkp_ew_chain");
  VKP_CHECK(ctx && in && out && ops && srcs && scalars, "vkp_ew_chain: null argument");
  VKP_CHECK(n_in >= 1 && n_in <= 4, "vkp_ew_chain: 1..4 inputs, got %d", n_in);
  VKP_CHECK(n_steps >= 1 && n_steps <= CH_MAX_STEPS, "vkp_ew_chain: 1..%d steps, got %d", CH_MAX_STEPS, n_steps);
  ChainProg prog;
  memset(&prog, 0, sizeof(prog));
  prog.n = n_steps;
  prog.nin = n_in;
  for (int k = 0; k < n_steps; k++) {
    int op = ops[k];
    const int src = srcs[k];
    VKP_CHECK(op >= 0 && op < CH_COUNT && op != CH_SQUARE, "vkp_ew_chain: bad op %d at step %d", op, k);
    VKP_CHECK(src >= 0 && src <= CH_SRC_TMP && (src == CH_SRC_SCALAR || src == CH_SRC_TMP || src < n_in),
              "vkp_ew_chain: step %d reads input %d of %d", k, src, n_in);
    if (op == CH_POW && src == CH_SRC_SCALAR && scalars[k] == 2.0f) op = CH_SQUARE;   // as dispatch_scalar does
    prog.st[k].op = (uint8_t)op;
    prog.st[k].src = (uint8_t)src;
    prog.st[k].s = scalars[k];
  }
  for (int i = 0; i < n_in; i++) VKP_CHECK(in[i] && ((uintptr_t)in[i] & 15) == 0, "vkp_ew_chain: input %d is null or unaligned", i);
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[5] = {(void*)in[0], n_in > 1 ? (void*)in[1] : nullptr, n_in > 2 ? (void*)in[2] : nullptr,
                   n_in > 3 ? (void*)in[3] : nullptr, (void*)out};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 5));
  if (count) {
    const size_t nvec = count >> 2;
    const size_t tiles = (nvec + CH_TILE_VEC - 1) / CH_TILE_VEC;
    const unsigned grid = (unsigned)(tiles ? tiles : 1);
    ew_chain_kernel<<<grid, EW_BLOCK, 0, ctx->stream>>>(prog, vkpt::host_coef(), in[0], n_in > 1 ? in[1] : in[0],
                                                        n_in > 2 ? in[2] : in[0], n_in > 3 ? in[3] : in[0], out, count);
    VKP_TRY(vkp_after_launch(ctx, "ew_chain"));
  }
  return vkp_finish_op(ctx, job);
}
