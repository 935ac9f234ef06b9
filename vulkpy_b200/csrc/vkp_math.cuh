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
  return (float)exp2_core(t);
}

VKP_HD float exp2_f(float x) {
  if (!(x == x)) return x;
  double t = (double)x;
  t = t < -200.0 ? -200.0 : (t > 200.0 ? 200.0 : t);
  return (float)exp2_core(t);
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
  return sign * (float)exp2_core(t);
}

VKP_HD float sign_f(float x) {
  return (x > 0.0f) ? 1.0f : ((x < 0.0f) ? -1.0f : 0.0f);      // GLSL sign(): sign.comp:22
}

}  // namespace vkpm
