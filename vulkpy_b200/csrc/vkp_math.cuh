// vkp_math.cuh -- scalar float32 math used by the element-wise kernels.
//
// The reference computes these with the GLSL built-ins of whatever Vulkan driver runs
// the shader (vulkpy/shader/exp.comp:22, log.comp:22, exp2.comp:22, log2.comp:22,
// pow.comp:25, pow_scalar.comp:23, rpow_scalar.comp:23).  The reference's own tests pin
// them at rtol=1e-7 against float64 NumPy (test/test_vulkpy.py:551-702), which on the
// tested points leaves no slack around the correctly rounded float32 result, so a 1-2 ulp
// libm is not good enough.  exp / exp2 / log / log2 / pow are therefore evaluated in
// binary64 (B200 issues DFMA at half the FFMA rate) and rounded once to binary32:
// the result is the correctly rounded float32 except in the ~2^-20 fraction of inputs
// that sit within 2^-44 of a rounding boundary.
//
// Everything here is __host__ __device__ so the very same code is compiled with g++ by
// tests/test_math_host.py and checked against float64 libm on the CPU.
#pragma once

#include <cstdint>
#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define VKP_HD __host__ __device__ __forceinline__
#else
#define VKP_HD inline
#endif

namespace vkpm {

VKP_HD double dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}

VKP_HD uint64_t d2bits(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; std::memcpy(&u, &x, 8); return u;
#endif
}

VKP_HD double bits2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; std::memcpy(&x, &u, 8); return x;
#endif
}

VKP_HD uint32_t f2bits(float x) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(x);
#else
  uint32_t u; std::memcpy(&u, &x, 4); return u;
#endif
}

VKP_HD float bits2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float x; std::memcpy(&x, &u, 4); return x;
#endif
}

// Reciprocal seed with ~2^-22 relative error (MUFU.RCP on the device).
VKP_HD float rcp_seed(float d) {
#if defined(__CUDA_ARCH__)
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
  return r;
#else
  return 1.0f / d;
#endif
}

// binary64 -> binary32, round to nearest, subnormal results flushed to (signed) zero.  The
// reference's test-suite expects softmax([100, 0]) == [1, 0] exactly (test/test_nn.py:120-126),
// i.e. exp(-100) = 3.8e-44 must come out as 0: GPU Vulkan drivers flush float32 denormals.
VKP_HD float d2f_ftz(double x) {
#if defined(__CUDA_ARCH__)
  // cvt.rn.ftz.f32.f64 compiles to F2F + DSETP + a predicated FMUL; round first (subnormal results are
  // correctly rounded subnormals), then let an FTZ add of -0 flush them: F2F + FADD, nothing on the FP64 pipe
  float r, z;
  asm("cvt.rn.f32.f64 %0, %1;" : "=f"(r) : "d"(x));
  asm("add.ftz.f32 %0, %1, 0f80000000;" : "=f"(z) : "f"(r));
  return z;
#else
  const float r = (float)x;
  if (r < 1.17549435e-38f && r > -1.17549435e-38f) return std::copysign(0.0f, r);
  return r;
#endif
}

VKP_HD double drint(double x) {
#if defined(__CUDA_ARCH__)
  return rint(x);
#else
  return std::nearbyint(x);
#endif
}

// 2^t for |t| <= 200 (caller clamps; float32 saturates long before).  Relative error ~2^-37.
VKP_HD double exp2_core(double t) {
  const double k = drint(t);
  const double r = t - k;                     // exact, |r| <= 0.5
  const double z = r * 0.693147180559945309417232121458;   // |z| <= 0.3466
  // Taylor series of e^z, degree 9 (remainder z^10/10! < 2^-37).
  double p = 2.75573192239858906525573192e-06;             // 1/9!
  p = dfma(p, z, 2.48015873015873015873015873e-05);        // 1/8!
  p = dfma(p, z, 1.98412698412698412698412698e-04);        // 1/7!
  p = dfma(p, z, 1.38888888888888888888888889e-03);        // 1/6!
  p = dfma(p, z, 8.33333333333333333333333333e-03);        // 1/5!
  p = dfma(p, z, 4.16666666666666666666666667e-02);        // 1/4!
  p = dfma(p, z, 1.66666666666666666666666667e-01);        // 1/3!
  p = dfma(p, z, 0.5);
  p = dfma(p, z, 1.0);
  p = dfma(p, z, 1.0);
  // p in [0.70, 1.42]; add k to the exponent field (|k| <= 200 keeps it normal)
  return bits2d(d2bits(p) + ((uint64_t)(int64_t)(int)k << 52));
}

// log(m) and binary exponent e such that x = 2^e * m, m in [sqrt(1/2), sqrt(2)).
// x must be a positive, finite, normal double.  Relative error of log(m) ~2^-44.
VKP_HD double logm_core(double x, int& e) {
  uint64_t u = d2bits(x);
  int ex = (int)(u >> 52) - 1023;
  u = (u & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL;     // m in [1,2)
  if (u > 0x3ff6a09e667f3bccULL) {                             // m > sqrt(2)
    u -= 0x0010000000000000ULL;                                // m /= 2
    ex += 1;
  }
  e = ex;
  const double m = bits2d(u);
  const double f = m - 1.0;                                    // exact
  const double d = m + 1.0;                                    // = 2 + f, exact
  double r = (double)rcp_seed((float)d);
  r = dfma(r, dfma(-d, r, 1.0), r);                            // Newton: ~2^-44
  const double s = f * r;                                      // |s| <= 0.1716
  const double w = s * s;
  // atanh series: log(m) = 2s (1 + w/3 + w^2/5 + ... + w^7/15); remainder < 2^-44
  double p = 1.0 / 15.0;
  p = dfma(p, w, 1.0 / 13.0);
  p = dfma(p, w, 1.0 / 11.0);
  p = dfma(p, w, 1.0 / 9.0);
  p = dfma(p, w, 1.0 / 7.0);
  p = dfma(p, w, 1.0 / 5.0);
  p = dfma(p, w, 1.0 / 3.0);
  p = p * w;
  const double s2 = s + s;
  return dfma(s2, p, s2);
}

VKP_HD float exp_f(float x) {
  if (!(x == x)) return x;
  double t = (double)x * 1.44269504088896340735992468100;
  t = t < -200.0 ? -200.0 : (t > 200.0 ? 200.0 : t);
  return d2f_ftz(exp2_core(t));
}

VKP_HD float exp2_f(float x) {
  if (!(x == x)) return x;
  double t = (double)x;
  t = t < -200.0 ? -200.0 : (t > 200.0 ? 200.0 : t);
  return d2f_ftz(exp2_core(t));
}

// kind 0: natural log, 1: log2.  Returns binary64 so pow() can reuse it.
template <int LOG2>
VKP_HD double log_d(float x) {
  const uint32_t ub = f2bits(x);
  if (ub - 1u >= 0x7f7fffffu) {                 // 0, negative, inf or nan
    const float inf = bits2f(0x7f800000u);
    if ((ub << 1) == 0u) return -(double)inf;   // +-0 -> -inf
    if (ub == 0x7f800000u) return (double)inf;  // +inf
    return (double)bits2f(0x7fc00000u);         // negative or nan -> nan
  }
  int e;
  const double lm = logm_core((double)x, e);
  if (LOG2) return dfma(lm, 1.44269504088896340735992468100, (double)e);
  return dfma((double)e, 0.693147180559945309417232121458, lm);
}

VKP_HD float log_f(float x)  { return (float)log_d<0>(x); }
VKP_HD float log2_f(float x) { return (float)log_d<1>(x); }

// C99 powf semantics (GLSL leaves x<0 undefined: vulkpy/shader/pow.comp:25).
VKP_HD float pow_f(float x, float y) {
  const uint32_t ux = f2bits(x), uy = f2bits(y);
  const float ay = bits2f(uy & 0x7fffffffu);
  if ((uy << 1) == 0u || ux == 0x3f800000u) return 1.0f;       // y == 0 or x == 1
  if (!(x == x) || !(y == y)) return x + y;                    // nan
  float sign = 1.0f;
  float ax = x;
  if (ux >> 31) {                                              // x < 0 or -0
    ax = bits2f(ux & 0x7fffffffu);
    const bool y_int = (ay >= 8388608.0f) || (ay == (float)(int)ay);
    const bool y_odd = (ay < 16777216.0f) && y_int && (((int)ay) & 1);
    if (ax != 0.0f && !y_int) return bits2f(0x7fc00000u);
    if (y_odd) sign = -1.0f;
  }
  if (ax == 1.0f) return sign;                                 // (-1)^y, y integer or inf
  double t = (double)y * log_d<1>(ax);                         // +-inf handled by clamp
  t = t < -200.0 ? -200.0 : (t > 200.0 ? 200.0 : t);
  return sign * d2f_ftz(exp2_core(t));
}


// =============================================================================================
// Table-driven fast paths for exp / exp2 / pow (same binary64 accuracy, ~2.5x fewer instructions)
//
//   log2(x): x = 2^e * m with m in [2/3, 4/3), i = 5 bits of (bits(x) - bits(2/3)), rc_i = 1 / centre_i rounded to
//            21 significant bits (scripts/fit/gen_log_tables.py: its binary64 form has a zero low word, so the
//            device shuffles ONE word per lookup and needs no float -> double conversion):
//            r = m * rc_i - 1 is exact in binary64 (|r| <= 2^-5.6), log2(m) = l2_i + log2(1 + r)
//            with l2_i = -log2(rc_i) and a degree-8 series (truncation < 2^-56 absolute).
//   2^t    : k = rint(32 t), r = t - k/32 (|r| <= 1/64), 2^t = 2^(k>>5) * e2[k & 31] * 2^r with a
//            degree-4 series for 2^r (truncation < 2^-39 relative).
// The 32-entry tables are passed through an accessor `TA`: on the device each lane of a warp keeps
// ONE entry per table in registers and lookups are warp shuffles (no shared memory, no bank
// conflicts); on the host (tests) the accessor indexes plain arrays.
// =============================================================================================
#define VKPM_TABLE_RC \
  0x1.7b8d600000000p+0, 0x1.72f5600000000p+0, 0x1.6abed00000000p+0, 0x1.62e3600000000p+0, \
  0x1.5b5d300000000p+0, 0x1.5427000000000p+0, 0x1.4d3be00000000p+0, 0x1.4697600000000p+0, \
  0x1.4035600000000p+0, 0x1.3a12000000000p+0, 0x1.3429c00000000p+0, 0x1.2e79500000000p+0, \
  0x1.28fdb00000000p+0, 0x1.23b4100000000p+0, 0x1.1e99c00000000p+0, 0x1.19ac600000000p+0, \
  0x1.14e9a00000000p+0, 0x1.104f700000000p+0, 0x1.0bdbc00000000p+0, 0x1.078cb00000000p+0, \
  0x1.0360900000000p+0, 0x1.0000000000000p+0, 0x1.edfd700000000p-1, 0x1.df88200000000p-1, \
  0x1.d1e5500000000p-1, 0x1.c503900000000p-1, 0x1.b8d3400000000p-1, 0x1.ad46700000000p-1, \
  0x1.a250a00000000p-1, 0x1.97e6800000000p-1, 0x1.8dfdf00000000p-1, 0x1.848db00000000p-1, \

#define VKPM_TABLE_L2 \
  -0x1.22e52b8935c71p-1, -0x1.11fa756958969p-1, -0x1.0170c08b65159p-1, \
  -0x1.e287afcd8d8b6p-2, -0x1.c2df2e30c514dp-2, -0x1.a3e0cbffcf234p-2, \
  -0x1.85854c712481bp-2, -0x1.67c659ebc3ee4p-2, -0x1.4a9dd9bf8b49fp-2, \
  -0x1.2e05bd3e7f63dp-2, -0x1.11f8b10599dfep-2, -0x1.ece2af0237decp-3, \
  -0x1.b6d5effd2b9bfp-3, -0x1.81c1f6e5e8252p-3, -0x1.4d9d55c042cb5p-3, \
  -0x1.1a606a4e9af49p-3, -0x1.d005e6b84c9d8p-4, -0x1.6cfc415a83409p-4, \
  -0x1.0b94081853b9ep-4, -0x1.577e950487283p-5, -0x1.35c93d810bd45p-6, \
  0x0.0p+0, 0x1.a73711e8083d2p-5, 0x1.8324c9b914bc7p-4, \
  0x1.16cecdbed61d6p-3, 0x1.69a73368d7dc5p-3, 0x1.ba3d3e8ffde41p-3, \
  0x1.0457cb8fdd002p-2, 0x1.2a8d349814dfcp-2, 0x1.4fcc132d48bcbp-2, \
  0x1.742054cd53df9p-2, 0x1.97955a856c4c6p-2, \

#define VKPM_TABLE_E2 \
  0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, \
  0x1.11301d0125b51p+0, 0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, \
  0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0, 0x1.306fe0a31b715p+0, \
  0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0, \
  0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, \
  0x1.6247eb03a5585p+0, 0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, \
  0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0, 0x1.8ace5422aa0dbp+0, \
  0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0, \
  0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, \
  0x1.cb720dcef9069p+0, 0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, \
  0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0,

// Long polynomial coefficients.  A 64-bit literal costs two UMOVs at EVERY use (and nvcc folds
// initialised __constant__ data back into literals), so the device accessor receives them as a
// kernel parameter and pins them in registers; short constants (1, 32, -1/32, 1.5*2^52) encode
// as DFMA immediates and stay literals.
//   lc: log2(1+r) = r (lc5 + r (lc4 + ... r lc0)),  lc_k = -+ log2(e)/(6-k)   (degree 6,
//       truncation log2(e) r^7/7 < 2^-44 for |r| <= 2^-6)
//   ec: 2^r = 1 + r (ec3 + r (ec2 + r (ec1 + r ec0))), ec_k = ln2^(4-k)/(4-k)!  (degree 4,
//       truncation (r ln2)^5/5! < 2^-39 for |r| <= 1/64)
#define VKPM_LOG_COEF                                                                          \
  -0x1.ec709dc3a03fdp-3, 0x1.2776c50ef9bfep-2, -0x1.71547652b82fep-2, 0x1.ec709dc3a03fdp-2,  \
  -0x1.71547652b82fep-1, 0x1.71547652b82fep+0
#define VKPM_EXP_COEF \
  0x1.3b2ab6fba4e77p-7, 0x1.c6b08d704a0c0p-5, 0x1.ebfbdff82c58fp-3, 0x1.62e42fefa39efp-1
struct MathCoef {
  double lc[6];
  double ec[4];
  double log2e;
};
inline MathCoef make_math_coef() {
  const MathCoef c = {{VKPM_LOG_COEF}, {VKPM_EXP_COEF}, 1.44269504088896340735992468100};
  return c;
}

struct HostTables {
  MathCoef c_ = make_math_coef();
  double lc(int i) const { return c_.lc[i]; }
  double ec(int i) const { return c_.ec[i]; }
  double log2e() const { return c_.log2e; }
  // like a warp shuffle, the accessors look at the low 5 bits of the index only
  double rc(uint32_t i) const { static const double t[32] = {VKPM_TABLE_RC}; return t[i & 31u]; }
  double l2(uint32_t i) const { static const double t[32] = {VKPM_TABLE_L2}; return t[i & 31u]; }
  double e2(uint32_t i) const { static const double t[32] = {VKPM_TABLE_E2}; return t[i & 31u]; }
};

VKP_HD double hilo2d(uint32_t hi, uint32_t lo) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double((int)hi, (int)lo);
#else
  return bits2d(((uint64_t)hi << 32) | lo);
#endif
}
VKP_HD uint32_t hi32(double x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__double2hiint(x);
#else
  return (uint32_t)(d2bits(x) >> 32);
#endif
}
VKP_HD uint32_t lo32(double x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__double2loint(x);
#else
  return (uint32_t)d2bits(x);
#endif
}

// (double)e for |e| < 2^31 without the conversion unit (I2F.F64 runs on the quarter-rate XU pipe): the
// binary64 number with high word 0x43300000 and low word e + 2^31 is 2^52 + 2^31 + e exactly.
VKP_HD double small_int_to_double(int e) {
  return hilo2d(0x43300000u, (uint32_t)e ^ 0x80000000u) - 4503601774854144.0;
}

// log2 of the positive NORMAL float whose bits are u.  x = 2^e * m with m in [2/3, 4/3) so that
// inputs near 1 have e = 0; interval 21 contains 1.0 and has rc = 1, l2 = 0 exactly, hence
// log2(1) = 0, log2(2^n) = n and full RELATIVE accuracy around 1.  |r| <= 0.0209; truncation
// log2(e) r^7/7 < 2^-41 absolute and < 2^-36 relative to the result.
template <class TA>
VKP_HD double log2_tab(uint32_t u, const TA& ta) {
#if defined(VKPM_LOG_BIAS)
  // biased exponent: d = (u - bits(2/3)) + (128 << 23) never wraps for a finite positive u, so e + 128 is a
  // LOGICAL shift and the int -> double step needs no sign flip
  const uint32_t d = u + 0x00d55555u;
  const uint32_t mb = u + 0x40000000u - (d & 0xff800000u);  // float bits of m = x / 2^e
  const double ed = hilo2d(0x43300000u, d >> 23) - 4503599627370624.0;   // (2^52 + e + 128) - (2^52 + 128)
#else
  const uint32_t d = u - 0x3f2aaaabu;                       // bits of 2/3
  const int e = (int)d >> 23;
  const uint32_t mb = u - (d & 0xff800000u);                // float bits of m = x / 2^e
#if !defined(VKPM_LOG_NO_I2F)
  const double ed = (double)e;                              // one conversion-unit instruction
#else
  const double ed = small_int_to_double(e);
#endif
#endif
  const double md = (double)bits2f(mb);                     // exact (one conversion; no bit assembly)
  // interval index = bits 18..22 of d; the accessor uses only the low 5 bits of the index it is given
  const double r = dfma(md, ta.rc(d >> 18), -1.0);          // exact
  double p = ta.lc(0);
  p = dfma(p, r, ta.lc(1));
  p = dfma(p, r, ta.lc(2));
  p = dfma(p, r, ta.lc(3));
  p = dfma(p, r, ta.lc(4));
  p = dfma(p, r, ta.lc(5));
  return dfma(p, r, ed + ta.l2(d >> 18));
}

// 2^t; meaningful for |t| <= 150 (other inputs give garbage but never trap)
template <class TA>
VKP_HD double exp2_tab(double t, const TA& ta) {
  const double magic = 6755399441055744.0;   // 1.5 * 2^52: the low word of (x + magic) is rint(x)
  const double kd = dfma(t, 32.0, magic);
  const int k = (int)lo32(kd);
  const double r = dfma(kd - magic, -0.03125, t);   // t - k/32: exact, |r| <= 1/64
  double p = ta.ec(0);
  p = dfma(p, r, ta.ec(1));
  p = dfma(p, r, ta.ec(2));
  p = dfma(p, r, ta.ec(3));
  p = dfma(p, r, 1.0);
  const double v = p * ta.e2(k);             // in [0.98, 2.02]; the accessor uses the low 5 bits of k
#if !defined(VKPM_EXP_NO_MAD) && defined(__CUDA_ARCH__)
  uint32_t hi;                                                     // += (k >> 5) in the exponent field: LOP3 + IMAD
  asm("mad.lo.u32 %0, %1, 32768, %2;" : "=r"(hi) : "r"((uint32_t)(k & ~31)), "r"(hi32(v)));
  return hilo2d(hi, lo32(v));
#else
  return hilo2d(hi32(v) + ((uint32_t)(k & ~31) << 15), lo32(v));   // += (k >> 5) in the exponent field
#endif
}

// Fast cores: every lane of a warp calls them together on the device (lookups are shuffles).
// They never branch to a slow path; inputs outside the fast domain set `special` and yield an
// unspecified value that the caller replaces with exp_f / exp2_f / pow_f.
template <class TA>
VKP_HD float exp_core(float x, const TA& ta, bool& special) {
  special |= !((f2bits(x) & 0x7fffffffu) < 0x42b00000u);     // |x| >= 88, inf, nan
  return d2f_ftz(exp2_tab((double)x * ta.log2e(), ta));
}
template <class TA>
VKP_HD float exp2_core(float x, const TA& ta, bool& special) {
  special |= !((f2bits(x) & 0x7fffffffu) < 0x42fc0000u);     // |x| >= 126, inf, nan
  return d2f_ftz(exp2_tab((double)x, ta));
}
template <class TA>
VKP_HD float log2_core(float x, const TA& ta, bool& special) {
  const uint32_t ux = f2bits(x);
  special |= !((ux - 0x00800000u) < 0x7f000000u);            // not (positive, normal, finite): value unused
  return (float)log2_tab(ux, ta);
}
template <class TA>
VKP_HD float log_core(float x, const TA& ta, bool& special) {
  const uint32_t ux = f2bits(x);
  special |= !((ux - 0x00800000u) < 0x7f000000u);
  return (float)(log2_tab(ux, ta) * 0.693147180559945309417232121458);
}
template <class TA>
VKP_HD float pow_core(float x, float y, const TA& ta, bool& special) {
  // x = 1 needs no special case: log2_tab(1) = 0 exactly, t = y * 0 is 0 (-> 1) or, for y = inf / nan,
  // nan, which the |t| test sends to pow_f (C99: pow(1, anything) = 1)
  const uint32_t ux = f2bits(x);
  const double t = (double)y * log2_tab(ux, ta);
  special |= !((ux - 0x00800000u) < 0x7f000000u) | !((hi32(t) & 0x7fffffffu) < 0x4062c000u);   // x not positive
  return d2f_ftz(exp2_tab(t, ta));                                           // normal, or |t| >= 150 / nan
}

// pow without the domain tests: returns the value of pow_core and k = rint(32 y log2(x)) (only meaningful when
// |y log2 x| < 2^26: callers bound |y| < 2^19 and x to positive normal floats).  The caller tests
// -126*32 <= k < 128*32 -- inside that range the result is a normal float (no flush needed), outside it
// (or for x / y outside their domains) it must call pow_f.  Lets a kernel test min / max over a vector.
template <class TA>
VKP_HD float pow_core_nc(float x, double yd, const TA& ta, int& k_out) {
  const double t = yd * log2_tab(f2bits(x), ta);
  const double magic = 6755399441055744.0;
  const double kd = dfma(t, 32.0, magic);
  k_out = (int)lo32(kd);
  const double v = exp2_tab(t, ta);
#if defined(__CUDA_ARCH__)
  float r;
  asm("cvt.rn.f32.f64 %0, %1;" : "=f"(r) : "d"(v));
  return r;
#else
  return (float)v;
#endif
}

// =============================================================================================
// a ** s for a LAUNCH-CONSTANT exponent s (pow_scalar.comp:25): no logarithm, no exponential.
//   x = 2^e m, m in [2/3, 4/3), interval i and rc_i exactly as in log2_tab (r = m rc_i - 1 exact,
//   |r| <= 0.0209):
//       x^s = 2^(s e) * rc_i^(-s) * (1 + r)^s
//   2^(s e): 256-entry table indexed by the low byte of e;  rc_i^(-s): 32-entry table;  (1 + r)^s: the
//   binomial series sum_k C(s, k) r^k cut at degree D (host picks D = 6 / 8 / 10 so that the tail is
//   below 2^-42; s = 2.7 needs 6).  9 binary64 operations per element at D = 6 (pow_core: 17), two
//   conversions, no range test besides "x is a positive normal float": the tables are binary64, so
//   2^(s e) cannot overflow for |s| <= 7.75 and the final conversion produces inf / flushes by itself.
//   Relative error before the final rounding < 2^-41  ->  <= 0.5001 ulp, same contract as pow_core.
// A first version with 128 intervals (D = 4, 7 binary64 operations, tables in shared memory) halved the
// instruction count but ran no faster than pow_core: 6.3 shared-memory wavefronts per warp of elements
// (random 16-byte table reads conflict) kept that pipe 82 % busy (profiles/r02_ncu_rows_v3.md).  With 32
// intervals rc_i and rc_i^(-s) live one entry per lane and a lookup is a shuffle (3 wavefronts); only
// 2^(s e) stays in shared memory, where neighbouring elements mostly share their exponent (broadcast).
// Tables depend on s: built on the device by pows_build_kernel (vkp_elementwise.cu) or on the host
// (PowsHostTables, tests) with the same formulas; coefficients C(s, k) come from pows_plan.
// =============================================================================================
struct PowsCoef {
  double b[11];   // b[k] = C(s, k); b[0] = 1
};

// exponent e of x = 2^e m for the table slot jb = e & 255 (e runs over -126 .. 128)
VKP_HD int pows_slot_exponent(uint32_t jb) { return jb <= 128u ? (int)jb : (int)jb - 256; }

// Host: pick the series degree for exponent s (0: this path does not apply) and fill the coefficients.
inline int pows_plan(float s, PowsCoef& c) {
  if (!(std::fabs(s) <= 7.75f)) return 0;                  // also nan; keeps |s e| < 1000
  static double rmax = 0.0;
  if (rmax == 0.0) {
    static const double rc[32] = {VKPM_TABLE_RC};
    double m = 0.0;
    for (uint32_t i = 0; i < 32; i++) {
      const uint32_t first = 0x3f2aaaabu + (i << 18);
      m = std::fmax(m, std::fabs((double)bits2f(first) * rc[i] - 1.0));
      m = std::fmax(m, std::fabs((double)bits2f(first + 0x3ffffu) * rc[i] - 1.0));
    }
    rmax = m;
  }
  long double bk[21];
  bk[0] = 1.0L;
  for (int k = 1; k <= 20; k++) bk[k] = bk[k - 1] * ((long double)s - (k - 1)) / k;
  for (int k = 0; k <= 10; k++) c.b[k] = (double)bk[k];
  for (int D = 6; D <= 10; D += 2) {
    long double tail = 0.0L, rk = 1.0L;
    for (int k = 1; k <= 20; k++) {
      rk *= rmax;
      if (k > D) tail += std::fabs(bk[k]) * rk;
    }
    if (tail < 0x1p-42L) return D;
  }
  return 0;
}
struct PowsHostTables {
  double rc_[32], cs_[32], es_[256];
  PowsCoef c_;
  explicit PowsHostTables(float s, const PowsCoef& c) : c_(c) {
    static const double rc[32] = {VKPM_TABLE_RC};
    for (uint32_t i = 0; i < 32; i++) {
      rc_[i] = rc[i];
      cs_[i] = std::pow(rc[i], -(double)s);
    }
    for (uint32_t j = 0; j < 256; j++) es_[j] = std::exp2((double)s * pows_slot_exponent(j));
  }
  double rc(uint32_t d) const { return rc_[(d >> 18) & 31u]; }
  double cs(uint32_t d) const { return cs_[(d >> 18) & 31u]; }
  double es(uint32_t d) const { return es_[(d >> 23) & 255u]; }
  double b(int k) const { return c_.b[k]; }
};

// u = bits of a positive NORMAL finite float.  Accessors take d = u - bits(2/3) and pick their bit field.
template <int D, class PT>
VKP_HD float pows_core(uint32_t u, const PT& t) {
  const uint32_t d = u - 0x3f2aaaabu;
  const uint32_t mb = u - (d & 0xff800000u);                // float bits of m = x / 2^e
  const double md = (double)bits2f(mb);
  const double r = dfma(md, t.rc(d), -1.0);                 // exact
  double p = t.b(D);
#pragma unroll
  for (int k = D - 1; k >= 1; k--) p = dfma(p, r, t.b(k));
  p = dfma(p, r, 1.0);
  return d2f_ftz((p * t.cs(d)) * t.es(d));
}

// convenience wrappers (host tests, single-element callers)
template <class TA>
VKP_HD float exp_fast(float x, const TA& ta) {
  bool sp = false;
  const float r = exp_core(x, ta, sp);
  return sp ? exp_f(x) : r;
}
template <class TA>
VKP_HD float exp2_fast(float x, const TA& ta) {
  bool sp = false;
  const float r = exp2_core(x, ta, sp);
  return sp ? exp2_f(x) : r;
}
template <class TA>
VKP_HD float log_fast(float x, const TA& ta) {
  bool sp = false;
  const float r = log_core(x, ta, sp);
  return sp ? log_f(x) : r;
}
template <class TA>
VKP_HD float log2_fast(float x, const TA& ta) {
  bool sp = false;
  const float r = log2_core(x, ta, sp);
  return sp ? log2_f(x) : r;
}
template <class TA>
VKP_HD float pow_fast(float x, float y, const TA& ta) {
  bool sp = false;
  const float r = pow_core(x, y, ta, sp);
  return sp ? pow_f(x, y) : r;
}

// sin and cos of a float angle in [-8, 8] (Box-Muller feeds 2*pi*u, u in [0,1)): Cody-Waite
// reduction by pi/2 in two FMAs, degree-7 / degree-8 minimax polynomials on [-pi/4, pi/4]
// (about 1 ulp), quadrant fix-up.  ~20 FP32 instructions; the general sincosf carries a
// Payne-Hanek slow path that is never needed here.
VKP_HD float ffma(float a, float b, float c) {
#if defined(__CUDA_ARCH__)
  return __fmaf_rn(a, b, c);
#else
  return std::fmaf(a, b, c);
#endif
}
VKP_HD void sincos_small(float x, float& s, float& c) {
  const float kf = ffma(x, 0.636619772367581343f, 12582912.0f) - 12582912.0f;   // rint(x * 2/pi)
  const int k = (int)kf;
  float r = ffma(-kf, 1.57079637050628662109375f, x);       // pi/2 high part
  r = ffma(-kf, -4.37113900018624283e-8f, r);               // pi/2 low part
  const float z = r * r;
  float ps = ffma(z, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = ffma(ps, z, -1.6666654611e-1f);
  const float sr = ffma(ps * z, r, r);
  float pc = ffma(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = ffma(pc, z, 4.166664568298827e-2f);
  const float cr = ffma(pc * z, z, ffma(z, -0.5f, 1.0f));
  const float s0 = (k & 1) ? cr : sr;
  const float c0 = (k & 1) ? sr : cr;
  s = (k & 2) ? -s0 : s0;
  c = ((k + 1) & 2) ? -c0 : c0;
}

// Reciprocal square root seed with ~2^-22 relative error (MUFU.RSQ on the device).
VKP_HD float rsqrt_seed(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return 1.0f / std::sqrt(x);
#endif
}

// log(x) for a positive, finite, NORMAL float32 x, all in float32: x = 2^e m with m in
// [2/3, 4/3), log(x) = e ln2 + f + f^2 Q(f), f = m - 1; Q fitted by scripts/fit/fit_bm_log.py.
// About 0.9 ulp.  No checks for zero / negative / subnormal / inf: callers know their range.
VKP_HD float log_pos_normal(float x) {
  const uint32_t b = f2bits(x);
  const int32_t e = (int32_t)(b - 0x3f2aaaabu) >> 23;
  const float f = bits2f(b - ((uint32_t)e << 23)) - 1.0f;
  float q = -0.135309964f;
  q = ffma(q, f, 0.142100722f);
  q = ffma(q, f, -0.120142907f);
  q = ffma(q, f, 0.139542714f);
  q = ffma(q, f, -0.166959479f);
  q = ffma(q, f, 0.200140804f);
  q = ffma(q, f, -0.249992505f);
  q = ffma(q, f, 0.333331376f);
  q = ffma(q, f, -0.50000006f);
  return ffma((float)e, 0.693147180559945f, ffma(f, f * q, f));
}

// log(x) for x = k 2^-23, k = 1 .. 2^23 -- the only values 1 - u can take for a uniform
// u = (bits >> 9) 2^-23 (prng_xoshiro128pp_float.comp:44).  Worst error over all 2^23 inputs:
// 0.92 ulp (tests/test_math_host.py checks every one of them).
VKP_HD float bm_log(float x) { return log_pos_normal(x); }

// asinh(x) = sign(x) log1p(|x| + x^2 / (1 + sqrt(1 + x^2)))  (asinh.comp:22 leaves it to the driver).
// The log1p is log(u) + (w - (u - 1)) / u with u = 1 + w, which keeps full relative accuracy for
// tiny |x|.  About 40 float32 instructions, no branches below |x| = 1e18; error < 2.5 ulp.
VKP_HD float asinh_f(float x) {
  const float ax = bits2f(f2bits(x) & 0x7fffffffu);
  if (!(ax < 1e18f)) {                                  // x^2 would overflow; inf and nan land here too
    if (!(ax <= 3.4028234e38f)) return x;
    const float r = log_pos_normal(ax) + 0.693147180559945f;
    return bits2f(f2bits(r) | (f2bits(x) & 0x80000000u));
  }
  const float z = ax * ax;
  const float t = 1.0f + z;
  const float y = rsqrt_seed(t);
  const float s0 = t * y, h = 0.5f * y;
  const float s = ffma(ffma(-s0, s0, t), h, s0);       // sqrt(1 + x^2)
  const float w = ax + z * rcp_seed(1.0f + s);
  const float u = 1.0f + w;
  const float c = w - (u - 1.0f);
  const float r = log_pos_normal(u) + c * rcp_seed(u);
  return bits2f(f2bits(r) | (f2bits(x) & 0x80000000u));
}

// sqrt(x) for 0 <= x <= 32 (-2 log(2^-23) = 31.9): reciprocal-square-root seed and one
// Newton step, no range checks.  x = 0 gives 0 (the seed is clamped so 0 * y stays 0).
VKP_HD float bm_sqrt(float x) {
#if defined(__CUDA_ARCH__)
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  y = fminf(y, 1e18f);
  const float s = x * y, h = 0.5f * y;
  return ffma(ffma(-s, s, x), h, s);
#else
  return std::sqrt(x);
#endif
}

// One Box-Muller pair exactly in the shader's operation order (prng_box_muller.comp:26-31):
//   r = sqrt(-2 * log(1 - u0)) * stddev;  angle = 6.28318530718f * u1;
//   o0 = mean + r * sin(angle);  o1 = mean + r * cos(angle)
// `om` is 1 - u0 (exact in float32 for every u0 the generator produces).  The reference leaves
// log / sqrt / sin / cos to the Vulkan driver; here log <= 0.92 ulp, sqrt <= 1 ulp, sin / cos
// <= 1.6 ulp, every multiply and add rounded separately as GLSL does without `precise` fusing.
VKP_HD void box_muller_core(float om, float u1, float mean, float stddev, float& o0, float& o1) {
  const float r = bm_sqrt(fabsf(-2.0f * bm_log(om))) * stddev;
  float s, c;
  sincos_small(6.28318530718f * u1, s, c);
  o0 = mean + r * s;
  o1 = mean + r * c;
}

// The same pair through the special-function unit -- what the reference's shader turns into on this very GPU:
// a Vulkan driver compiles GLSL log / sqrt / sin / cos to the hardware approximations (lg2 * ln2, sqrt.approx,
// sin.approx / cos.approx), whose error the Vulkan spec allows to be far larger (sin / cos: 2^-11 absolute) than
// what is measured here over EVERY input the generator can produce (scripts/micro/mufu_error.cu, numbers in
// DESIGN.md).  lg2.approx has only ABSOLUTE accuracy near 1 (measured 2^-22), i.e. an error of 3.3e-7 / (2 r)
// in r = sqrt(L): below u0 = 1 - om = 2^-10 (r < 0.044, error > 3.7e-6) the logarithm is the series
// 2 u (1 + u/2 + u^2/3) instead (truncation < 2^-31 relative).  One lane in 1024 needs it, so callers test it
// once per group of pairs (bm_fast_L + bm_fast_fix) and the series stays off the common path.
// stddev = 1, mean = 0 (the default call) skips the multiply and the add: x * 1 and 0 + x are exact; only a
// result of -0 (u0 = 0, sin < 0) comes out as -0 where the three-operation form gives +0.
constexpr float BM_SERIES_BELOW = 0.0009765625f;   // 2^-10
VKP_HD float bm_fast_series(float u0) {
  float p = ffma(u0, 0.6666667f, 1.0f);
  p = ffma(p, u0, 2.0f);
  return p * u0;
}
#if defined(__CUDACC__)
__device__ __forceinline__ float bm_fast_L(float om) {      // -2 ln(om) >= 0 through lg2.approx
  float l2;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(om));
  return l2 * -1.3862943611198906f;
}
template <bool UNIT>
__device__ __forceinline__ void bm_fast_finish(float L, float u1, float mean, float stddev, float& o0, float& o1) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(L));
  const float angle = 6.28318530718f * u1;
  if (UNIT) {
    o0 = r * __sinf(angle);
    o1 = r * __cosf(angle);
  } else {
    r *= stddev;
    o0 = mean + r * __sinf(angle);
    o1 = mean + r * __cosf(angle);
  }
}
#endif
VKP_HD void box_muller_fast(float om, float u1, float mean, float stddev, float& o0, float& o1) {
#if defined(__CUDA_ARCH__)
  const float u0 = 1.0f - om;                          // exact
  float L = bm_fast_L(om);
  if (u0 < BM_SERIES_BELOW) L = bm_fast_series(u0);
  if (mean == 0.0f && stddev == 1.0f) bm_fast_finish<true>(L, u1, mean, stddev, o0, o1);
  else bm_fast_finish<false>(L, u1, mean, stddev, o0, o1);
#else
  box_muller_core(om, u1, mean, stddev, o0, o1);
#endif
}

template <bool FAST = false>
VKP_HD void box_muller_pair(float u0, float u1, float mean, float stddev, float& o0, float& o1) {
  // uniforms outside [0, 1) can only come from a caller-made buffer; keep log's domain
  float om = 1.0f - u0;
  om = om < 1.1920929e-7f ? 1.1920929e-7f : (om > 1.0f ? 1.0f : om);
  if (FAST) box_muller_fast(om, u1, mean, stddev, o0, o1);
  else box_muller_core(om, u1, mean, stddev, o0, o1);
}

VKP_HD float sign_f(float x) {
  return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);      // GLSL sign(): sign.comp:22
}

}  // namespace vkpm
