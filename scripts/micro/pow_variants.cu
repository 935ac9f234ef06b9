// Micro-benchmark: variants of the table-driven pow / log2 / exp kernels (vkp_math.cuh) on 2^28 floats.
// Build one binary per macro set (scripts/micro/pow_variants.sh):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I vulkpy_b200/csrc [-DVKPM_LOG_NO_I2F -DVKPM_EXP_NO_MAD -DVKPM_LOG_BIAS -DPV_MINB=5 ...] -o pow_vX
// Prints ms, GB/s (algorithmic bytes) and a checksum of the output bits (equal checksums = identical results).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "vkp_tables.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)

#ifndef PV_BLOCK
#define PV_BLOCK 256
#endif
#ifndef PV_UNROLL
#define PV_UNROLL 4
#endif
#ifndef PV_MINB
#define PV_MINB 1
#endif

using vkpt::LaneTables;

struct TPow {
  __device__ float fast(const LaneTables& t, float a, float b, bool& sp) const { return vkpm::pow_core(a, b, t, sp); }
  __device__ float slow(float a, float b) const { return vkpm::pow_f(a, b); }
};
struct TPowScalar {
  float s;
  __device__ float fast(const LaneTables& t, float a, float, bool& sp) const { return vkpm::pow_core(a, s, t, sp); }
  __device__ float slow(float a, float) const { return vkpm::pow_f(a, s); }
};
struct TLog2 {
  __device__ float fast(const LaneTables& t, float a, float, bool& sp) const { return vkpm::log2_core(a, t, sp); }
  __device__ float slow(float a, float) const { return vkpm::log2_f(a); }
};
struct TExp {
  __device__ float fast(const LaneTables& t, float a, float, bool& sp) const { return vkpm::exp_core(a, t, sp); }
  __device__ float slow(float a, float) const { return vkpm::exp_f(a); }
};

// scalar exponent, domain tests hoisted out of the per-element path: |s| < 2^19 is tested once per thread, x and
// k = rint(32 t) through min / max over the vector
struct TPowScalarMM {
  float s;
  __device__ float slow(float a, float) const { return vkpm::pow_f(a, s); }
  __device__ float4 fast4(const LaneTables& t, float4 a, float4, bool& sp) const {
    const double sd = (double)s;
    int k0, k1, k2, k3;
    float4 r;
    r.x = vkpm::pow_core_nc(a.x, sd, t, k0);
    r.y = vkpm::pow_core_nc(a.y, sd, t, k1);
    r.z = vkpm::pow_core_nc(a.z, sd, t, k2);
    r.w = vkpm::pow_core_nc(a.w, sd, t, k3);
    const uint32_t u0 = __float_as_uint(a.x), u1 = __float_as_uint(a.y), u2 = __float_as_uint(a.z), u3 = __float_as_uint(a.w);
    const uint32_t umin = min(min(u0, u1), min(u2, u3)), umax = max(max(u0, u1), max(u2, u3));
    const int kmin = min(min(k0, k1), min(k2, k3)), kmax = max(max(k0, k1), max(k2, k3));
    sp |= (umin < 0x00800000u) | (umax >= 0x7f800000u) | (kmin < -126 * 32) | (kmax >= 128 * 32) |
          !(fabsf(s) < 524288.0f);
    return r;
  }
};

template <class F, class = void> struct Has4 { static constexpr bool v = false; };
template <class F> struct Has4<F, decltype((void)&F::fast4)> { static constexpr bool v = true; };

template <class F>
__device__ __forceinline__ float4 eval4(const F& f, const LaneTables& tab, float4 a, float4 b, bool& sp) {
  if constexpr (Has4<F>::v) {
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

template <class F>
__device__ __noinline__ float4 redo_slow(const F f, float4 a, float4 b) {
  return make_float4(f.slow(a.x, b.x), f.slow(a.y, b.y), f.slow(a.z, b.z), f.slow(a.w, b.w));
}

template <int NIN, class F>
__global__ void __launch_bounds__(PV_BLOCK, PV_MINB)
ew_tab_kernel(F f, const __grid_constant__ vkpm::MathCoef coef, const float* in0, const float* in1, float* out, size_t n) {
  const LaneTables tab(coef);
  const float4* v0 = reinterpret_cast<const float4*>(in0);
  const float4* v1 = reinterpret_cast<const float4*>(in1);
  float4* vo = reinterpret_cast<float4*>(out);
  const size_t base = (size_t)blockIdx.x * (PV_BLOCK * PV_UNROLL) + threadIdx.x;
  float4 a[PV_UNROLL], b[PV_UNROLL];
#pragma unroll
  for (int u = 0; u < PV_UNROLL; u++) {          // the benchmark sizes are multiples of the tile
    a[u] = v0[base + (size_t)u * PV_BLOCK];
    if (NIN > 1) b[u] = v1[base + (size_t)u * PV_BLOCK];
  }
#pragma unroll
  for (int u = 0; u < PV_UNROLL; u++) {
    const size_t i = base + (size_t)u * PV_BLOCK;
    bool sp = false;
    float4 r = eval4(f, tab, a[u], b[u], sp);
    if (sp) r = redo_slow(f, a[u], b[u]);
    vo[i] = r;
  }
}

__global__ void fill_kernel(float* p, size_t n, float lo, float hi, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u + seed;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = lo + (hi - lo) * ((x >> 8) * (1.0f / 16777216.0f));
  }
}

__global__ void checksum_kernel(const uint32_t* p, size_t n, unsigned long long* out) {
  unsigned long long s = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    s += (unsigned long long)p[i] * (i % 1021 + 1);
  atomicAdd(out, s);
}

template <int NIN, class F>
void run(const char* name, F f, const float* a, const float* b, float* o, size_t n, int bytes_per_elem) {
  const vkpm::MathCoef coef = vkpm::make_math_coef();
  const unsigned grid = (unsigned)(n / 4 / (PV_BLOCK * PV_UNROLL));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) ew_tab_kernel<NIN, F><<<grid, PV_BLOCK>>>(f, coef, a, b, o, n);
  CK(cudaDeviceSynchronize());
  const int reps = 20;
  cudaEventRecord(e0);
  for (int i = 0; i < reps; i++) ew_tab_kernel<NIN, F><<<grid, PV_BLOCK>>>(f, coef, a, b, o, n);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
  unsigned long long* d; CK(cudaMalloc(&d, 8)); CK(cudaMemset(d, 0, 8));
  checksum_kernel<<<1184, 256>>>((const uint32_t*)o, n, d);
  unsigned long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost)); cudaFree(d);
  cudaFuncAttributes at; cudaFuncGetAttributes(&at, ew_tab_kernel<NIN, F>);
  printf("  %-8s %7.4f ms  %7.1f GB/s  regs %3d  checksum %016llx\n", name, ms, (double)n * bytes_per_elem / ms / 1e6, at.numRegs, h);
}

int main() {
  const size_t n = (size_t)1 << 28;
  float *a, *b, *o;
  CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMalloc(&o, n * 4));
  fill_kernel<<<1184, 256>>>(a, n, 0.5f, 2.0f, 1u);
  fill_kernel<<<1184, 256>>>(b, n, -2.0f, 2.0f, 2u);
  CK(cudaDeviceSynchronize());
  printf("block %d unroll %d minb %d\n", PV_BLOCK, PV_UNROLL, PV_MINB);
  run<1>("a**2.7", TPowScalar{2.7f}, a, nullptr, o, n, 8);
  run<1>("a**2.7mm", TPowScalarMM{2.7f}, a, nullptr, o, n, 8);
  run<2>("a**b", TPow(), a, b, o, n, 12);
  run<1>("log2(a)", TLog2(), a, nullptr, o, n, 8);
  run<1>("exp(b)", TExp(), b, nullptr, o, n, 8);
  return 0;
}
