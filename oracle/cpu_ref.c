/* cpu_ref.c -- C restatement of the reference's compute shaders and PRNG: TEST / BASELINE
 * INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Each function follows one shader of /root/reference/vulkpy/shader (cited per function) with
 * the shader's own structure: one "invocation" per output element, serial per-output loops,
 * float32 arithmetic, uint32 index math.  OpenMP spreads invocations over the host cores the
 * way a CPU Vulkan ICD (Mesa lavapipe, the reference's CI device: Dockerfile:1-12) spreads
 * workgroups.  Used (a) by tests/ to cross-check the NumPy oracle at sizes where Python loops
 * are too slow and (b) by bench.py as the `cpu_baseline` / `--impl reference` arm, because the
 * reference itself cannot run in this image (no libvulkan, glslc or ICD: SURVEY.md 8(c)).
 * GLSL built-ins map to the C float library (sinf, expf, powf ...), as an LLVM-based ICD does.
 *
 * Build: make -C oracle   (gcc -O3 -march=x86-64-v3 -fopenmp -ffp-contract=off -shared)
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define API __attribute__((visibility("default")))

API int ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

API void ref_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- binary ops: add/sub/mul/div/max/min/pow.comp (add.comp:21-26) ----------------------- */
enum { B_ADD, B_SUB, B_MUL, B_DIV, B_MAX, B_MIN, B_POW };

static inline float bin(int op, float a, float b) {
  switch (op) {
    case B_ADD: return a + b;
    case B_SUB: return a - b;
    case B_MUL: return a * b;
    case B_DIV: return a / b;
    case B_MAX: return a < b ? b : a;   /* GLSL max(x,y): y if x<y else x */
    case B_MIN: return b < a ? b : a;
    default: return powf(a, b);
  }
}

API void ref_binary(int op, const float* a, const float* b, float* c, uint32_t n) {
#pragma omp parallel for schedule(static)
  for (uint32_t i = 0; i < n; i++) c[i] = bin(op, a[i], b[i]);
}

/* vec (+) scalar and reversed forms (add_scalar.comp:19-24, rsub_scalar.comp:23) */
API void ref_scalar(int op, int reverse, const float* a, float s, float* b, uint32_t n) {
#pragma omp parallel for schedule(static)
  for (uint32_t i = 0; i < n; i++) b[i] = reverse ? bin(op, s, a[i]) : bin(op, a[i], s);
}

/* ---- unary ops: abs ... invsqrt.comp (abs.comp:22) ---------------------------------------- */
enum { U_ABS, U_SIGN, U_SIN, U_COS, U_TAN, U_ASIN, U_ACOS, U_ATAN, U_SINH, U_COSH, U_TANH, U_ASINH,
       U_ACOSH, U_ATANH, U_EXP, U_LOG, U_EXP2, U_LOG2, U_SQRT, U_INVSQRT };

static inline float un(int op, float a) {
  switch (op) {
    case U_ABS: return fabsf(a);
    case U_SIGN: return a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f);
    case U_SIN: return sinf(a);
    case U_COS: return cosf(a);
    case U_TAN: return tanf(a);
    case U_ASIN: return asinf(a);
    case U_ACOS: return acosf(a);
    case U_ATAN: return atanf(a);
    case U_SINH: return sinhf(a);
    case U_COSH: return coshf(a);
    case U_TANH: return tanhf(a);
    case U_ASINH: return asinhf(a);
    case U_ACOSH: return acoshf(a);
    case U_ATANH: return atanhf(a);
    case U_EXP: return expf(a);
    case U_LOG: return logf(a);
    case U_EXP2: return exp2f(a);
    case U_LOG2: return log2f(a);
    case U_SQRT: return sqrtf(a);
    default: return 1.0f / sqrtf(a);
  }
}

API void ref_unary(int op, const float* a, float* b, uint32_t n) {
#pragma omp parallel for schedule(static)
  for (uint32_t i = 0; i < n; i++) b[i] = un(op, a[i]);
}

/* clamp.comp:24-29 (lo/hi arrays), clamp_ss.comp:23 (lo/hi scalars) */
API void ref_clamp_vv(const float* a, const float* lo, const float* hi, float* d, uint32_t n) {
#pragma omp parallel for schedule(static)
  for (uint32_t i = 0; i < n; i++) d[i] = fminf(fmaxf(a[i], lo[i]), hi[i]);
}

API void ref_clamp_ss(const float* a, float lo, float hi, float* d, uint32_t n) {
#pragma omp parallel for schedule(static)
  for (uint32_t i = 0; i < n; i++) d[i] = fminf(fmaxf(a[i], lo), hi);
}

/* ---- broadcast: add_broadcast.comp:25-45 (per-element ndim-long div/mod walk) ---------------- */
API void ref_broadcast_binary(int op, const float* a, const float* b, float* c, const uint32_t* shapeABC,
                              uint32_t sizeA, uint32_t sizeB, uint32_t sizeC, uint32_t ndim) {
#pragma omp parallel for schedule(static)
  for (uint32_t ci = 0; ci < sizeC; ci++) {
    uint32_t sa = sizeA, sb = sizeB, sc = sizeC, ai = 0, bi = 0, rem = ci;
    for (uint32_t dim = 0; dim < ndim; dim++) {
      const uint32_t da = shapeABC[dim], db = shapeABC[dim + ndim], dc = shapeABC[dim + 2 * ndim];
      sa /= da; sb /= db; sc /= dc;
      const uint32_t d = rem / sc;
      ai += sa * (d < da - 1 ? d : da - 1);
      bi += sb * (d < db - 1 ? d : db - 1);
      rem = rem % sc;
    }
    c[ci] = bin(op, a[ai], b[bi]);
  }
}

/* broadcast.comp:25-44 */
API void ref_broadcast_copy(const float* a, float* b, const uint32_t* shapeA, const uint32_t* shapeB,
                            uint32_t sizeA, uint32_t sizeB, uint32_t ndim) {
#pragma omp parallel for schedule(static)
  for (uint32_t bi = 0; bi < sizeB; bi++) {
    uint32_t sa = sizeA, sb = sizeB, ai = 0, rem = bi;
    for (uint32_t dim = 0; dim < ndim; dim++) {
      sa /= shapeA[dim]; sb /= shapeB[dim];
      const uint32_t d = rem / sb;
      ai += sa * (d < shapeA[dim] - 1 ? d : shapeA[dim] - 1);
      rem = rem % sb;
    }
    b[bi] = a[ai];
  }
}

/* ---- reductions ---------------------------------------------------------------------------- */
enum { R_SUM, R_PROD, R_MAX, R_MIN };

static inline float red(int op, float acc, float x) {
  switch (op) {
    case R_SUM: return acc + x;
    case R_PROD: return acc * x;
    case R_MAX: return acc < x ? x : acc;
    default: return x < acc ? x : acc;
  }
}

/* sum_axis.comp:20-32 / sum_axis_rebroadcast.comp:20-35: one invocation per (i, j), serial k */
API void ref_reduce_axis(int op, const float* a, float* b, uint32_t prev, uint32_t axis, uint32_t post,
                         int rebroadcast) {
#pragma omp parallel for collapse(2) schedule(static)
  for (uint32_t i = 0; i < prev; i++) {
    for (uint32_t j = 0; j < post; j++) {
      const size_t ij = (size_t)i * axis * post + j;
      float acc = op == R_SUM ? 0.f : (op == R_PROD ? 1.f : a[ij]);
      for (uint32_t k = 0; k < axis; k++) acc = red(op, acc, a[(size_t)k * post + ij]);
      if (rebroadcast) {
        for (uint32_t k = 0; k < axis; k++) b[(size_t)k * post + ij] = acc;
      } else {
        b[(size_t)i * post + j] = acc;
      }
    }
  }
}

/* sum.comp:18-30 applied pass after pass like vkarray.py:1246-1274, but with every one of the
 * m = ceil(n/64) strided partials of a pass computed (the reference dispatches a single
 * 64-thread workgroup, which only covers n <= 4096: SURVEY Q2).  tmp needs ceil(n/64) floats. */
API float ref_reduce_full(int op, const float* a, uint32_t n, float* tmp) {
  const float* src = a;
  uint32_t cur = n;
  float* bufs[2] = {tmp, tmp + (n + 63) / 64};
  int which = 0;
  for (;;) {
    const uint32_t m = (cur + 63) / 64;
    float* dst = bufs[which];
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < m; i++) {
      float acc = op == R_SUM ? 0.f : (op == R_PROD ? 1.f : src[i]);
      for (uint32_t j = i; j < cur; j += m) acc = red(op, acc, src[j]);
      dst[i] = acc;
    }
    if (m == 1) return dst[0];
    src = dst;
    cur = m;
    which ^= 1;
  }
}

/* ---- gather.comp:21-26, gather_axis.comp:24-43 ---------------------------------------------- */
API void ref_gather(const float* a, const uint32_t* idx, float* c, uint32_t n) {
#pragma omp parallel for schedule(static)
  for (uint32_t i = 0; i < n; i++) c[i] = a[idx[i]];
}

API void ref_gather_axis(const float* a, const uint32_t* idx, float* c, uint32_t prev, uint32_t post,
                         uint32_t axis, uint32_t nidx) {
#pragma omp parallel for collapse(2) schedule(static)
  for (uint32_t k = 0; k < nidx; k++) {
    for (uint32_t i = 0; i < prev; i++) {
      uint32_t bk = idx[k];
      if (bk > axis) bk = axis; /* clamp(b[k], 0, axis_size): inclusive, as in the shader */
      for (uint32_t j = 0; j < post; j++)
        c[((size_t)k * prev + i) * post + j] = a[((size_t)i * axis + bk) * post + j];
    }
  }
}

/* ---- matmul.comp:23-33, batch_affine.comp:25-39: one invocation per output, serial k --------- */
API void ref_matmul(const float* a, const float* b, float* c, uint32_t M, uint32_t K, uint32_t N) {
#pragma omp parallel for collapse(2) schedule(static)
  for (uint32_t row = 0; row < M; row++) {
    for (uint32_t col = 0; col < N; col++) {
      float sum = 0.f;
      for (uint32_t s = 0; s < K; s++) sum += a[(size_t)row * K + s] * b[(size_t)s * N + col];
      c[(size_t)row * N + col] = sum;
    }
  }
}

API void ref_batch_affine(const float* w, const float* bias, const float* x, float* y, uint32_t batch,
                          uint32_t in, uint32_t out) {
#pragma omp parallel for collapse(2) schedule(static)
  for (uint32_t bi = 0; bi < batch; bi++) {
    for (uint32_t o = 0; o < out; o++) {
      float sum = 0.f;
      for (uint32_t i = 0; i < in; i++) sum += w[(size_t)o * in + i] * x[(size_t)bi * in + i];
      y[(size_t)bi * out + o] = sum + bias[o];
    }
  }
}

/* nn_cross_entropy.comp:25, nn_cross_entropy_backward.comp:25 */
API void ref_cross_entropy(const float* x, const float* y, float* L, uint32_t n, int backward) {
#pragma omp parallel for schedule(static)
  for (uint32_t i = 0; i < n; i++) L[i] = backward ? (-y[i] / (x[i] + 1e-8f)) : (-y[i] * logf(x[i] + 1e-8f));
}

/* ---- PRNG: _vkarray.cc:589-717, prng_xoshiro128pp_uint32.comp:26-43, ..._float.comp:26-44 ------ */
static inline uint32_t rotl(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

static inline uint32_t xo_next(uint32_t* s) {
  const uint32_t result = rotl(s[0] + s[3], 7) + s[0];
  const uint32_t t = s[1] << 9;
  s[2] ^= s[0];
  s[3] ^= s[1];
  s[1] ^= s[2];
  s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl(s[3], 11);
  return result;
}

static inline uint64_t splitmix64(uint64_t x) {
  uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

/* state: [size][4] (_vkarray.cc:653-674, jump with per-word write-back :597-621) */
API void ref_xoshiro_seed(uint32_t* state, uint32_t size, uint64_t seed) {
  static const uint32_t JUMP[4] = {0x8764000bu, 0xf542d2d3u, 0x6fa035c3u, 0x77f2db5bu};
  uint32_t s[4];
  for (int i = 0; i < 4; i++) {
    seed = splitmix64(seed);
    s[i] = (uint32_t)seed;
  }
  memcpy(state, s, 16);
  for (uint32_t l = 1; l < size; l++) {
    uint32_t acc[4] = {0, 0, 0, 0};
    for (int w = 0; w < 4; w++) {
      for (int b = 0; b < 32; b++) {
        if (JUMP[w] & (1u << b))
          for (int i = 0; i < 4; i++) acc[i] ^= s[i];
        xo_next(s);
      }
      memcpy(s, acc, 16);
    }
    memcpy(state + 4 * (size_t)l, s, 16);
  }
}

/* random(): chunks of `size`, chunk c -> out[c*size + lane]; lanes run in parallel inside a
 * chunk (one dispatch each, _vkarray.cc:697-717).  as_float selects the [0,1) mapping. */
API void ref_xoshiro_fill(uint32_t* state, uint32_t size, uint32_t* out, uint64_t n, int as_float) {
  for (uint64_t i = 0; i < n; i += size) {
    const uint32_t m = (n - i) < size ? (uint32_t)(n - i) : size;
#pragma omp parallel for schedule(static) if (m >= 4096)
    for (uint32_t l = 0; l < m; l++) {
      const uint32_t r = xo_next(state + 4 * (size_t)l);
      if (as_float) {
        const uint32_t bits = (r >> 9) | 0x3f800000u;
        float f;
        memcpy(&f, &bits, 4);
        f -= 1.0f;
        memcpy(out + i + l, &f, 4);
      } else {
        out[i + l] = r;
      }
    }
  }
}

/* prng_box_muller.comp:19-32 / prng_ibox_muller.comp:16-27 (a may alias b) */
API void ref_box_muller(const float* a, float* b, uint32_t n, float mean, float stddev) {
  const uint32_t npair = (n + 1) / 2;
#pragma omp parallel for schedule(static)
  for (uint32_t i = 0; i < npair; i++) {
    const uint32_t j = 2 * i, k = j + 1;
    const float u0 = a[j], u1 = a[k];
    const float r = sqrtf(-2 * logf(1.0f - u0)) * stddev;
    const float angle = 6.28318530718f * u1;
    b[j] = mean + r * sinf(angle);
    if (k < n) b[k] = mean + r * cosf(angle);
  }
}

/* prng_randrange.comp:20-27 */
API void ref_randrange(const float* a, uint32_t* b, uint32_t n, uint32_t low, uint32_t high) {
  const uint32_t range = high - low + 1;
#pragma omp parallel for schedule(static)
  for (uint32_t i = 0; i < n; i++) b[i] = low + (uint32_t)((float)range * a[i]);
}
