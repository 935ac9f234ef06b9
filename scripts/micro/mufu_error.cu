// Exhaustive error of the special-function unit (MUFU) over the inputs Box-Muller can see:
//   om = k 2^-23, k = 1 .. 2^23      r  = sqrt(-2 ln om)      via lg2.approx + sqrt.approx
//   u1 = j 2^-23, j = 0 .. 2^23-1    sin / cos(6.28318530718f * u1)   via sin.approx / cos.approx
// against binary64.  Prints max abs error overall and per range, so the thresholds of the fast normal path
// (vkp_math.cuh box_muller_fast) are measured numbers, not the PTX manual's bounds.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_error scripts/micro/mufu_error.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cmath>

__device__ double atomicMaxD(double* addr, double v) {
  unsigned long long* a = (unsigned long long*)addr;
  unsigned long long old = *a, assumed;
  do {
    assumed = old;
    if (__longlong_as_double(assumed) >= v) break;
    old = atomicCAS(a, assumed, __double_as_longlong(v));
  } while (assumed != old);
  return __longlong_as_double(old);
}

// out[0..7]: max |r_fast - r| for om in (2^-(b+1), 2^-b] b=0 (0.5,1], 1, 2, 3.., out[8]: om > 1-2^-5, out[9]: om in (0.5, 1-2^-5],
// out[11]: om in (0.5, 1-2^-10] (the range the lg2 path serves since the series threshold moved to u0 < 2^-10)
__global__ void log_err(double* out, double* lg_abs) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (k > (1u << 23)) return;
  const float om = (float)k * 1.1920928955078125e-7f;
  float l2;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(om));
  const float L = l2 * -1.3862943611198906f;
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(L));
  const double rd = sqrt(-2.0 * log((double)om));
  const double e = fabs((double)r - rd);
  int b = 0;
  float t = om;
  while (t <= 0.5f && b < 7) { t *= 2.0f; b++; }
  if (b == 0) {
    atomicMaxD(out + (om > 0.96875f ? 8 : 9), e);
    if (om <= 1.0f - 0.0009765625f) atomicMaxD(out + 11, e);
    atomicMaxD(lg_abs, fabs((double)l2 - log2((double)om)));
  }
  atomicMaxD(out + b, e);
  // relative error of r outside (0.5, 1]
  if (b > 0) atomicMaxD(out + 10, e / rd);
}

// series branch: u0 < 2^-10 (vkp_math.cuh BM_SERIES_BELOW)
__global__ void series_err(double* out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;     // u0 = k 2^-23, k < 2^18
  if (k >= (1u << 13)) return;                                   // u0 < 2^-10
  const float u = (float)k * 1.1920928955078125e-7f;
  float p = fmaf(u, 0.6666667f, 1.0f);
  p = fmaf(p, u, 2.0f);
  const float L = p * u;                                         // 2 (u + u^2/2 + u^3/3)
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(L));
  const double rd = sqrt(-2.0 * log1p(-(double)u));
  atomicMaxD(out, fabs((double)r - rd));
  if (k) atomicMaxD(out + 1, fabs((double)r - rd) / rd);
}

__global__ void sincos_err(double* out) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= (1u << 23)) return;
  const float u1 = (float)j * 1.1920928955078125e-7f;
  const float angle = 6.28318530718f * u1;
  const float s = __sinf(angle), c = __cosf(angle);
  atomicMaxD(out + 0, fabs((double)s - sin((double)angle)));
  atomicMaxD(out + 1, fabs((double)c - cos((double)angle)));
  // against the exact angle 2 pi u1 (what a turn-based evaluation would see)
  const double ex = 6.283185307179586476925286766559 * (double)u1;
  atomicMaxD(out + 2, fabs((double)s - sin(ex)));
  atomicMaxD(out + 3, fabs((double)c - cos(ex)));
}

int main() {
  double *d, h[16];
  cudaMalloc(&d, 16 * 8);
  cudaMemset(d, 0, 16 * 8);
  log_err<<<(1 << 23) / 256, 256>>>(d, d + 12);
  cudaMemcpy(h, d, 16 * 8, cudaMemcpyDeviceToHost);
  for (int b = 0; b < 8; b++) printf("r = sqrt(-2 ln om): om in (2^-%d, 2^-%d]%s  max abs err %.3e\n", b + 1, b, b == 7 ? " and below" : "", h[b]);
  printf("   om in (1-2^-5, 1]: %.3e   om in (0.5, 1-2^-5]: %.3e   max rel err for om <= 0.5: %.3e\n", h[8], h[9], h[10]);
  printf("   lg2.approx max abs err on (0.5, 1]: %.3e (2^-22 = 2.38e-7)\n", h[12]);
  printf("   om in (0.5, 1-2^-10] (lg2 path as shipped): %.3e\n", h[11]);
  cudaMemset(d, 0, 16 * 8);
  series_err<<<(1 << 13) / 256, 256>>>(d);
  cudaMemcpy(h, d, 16 * 8, cudaMemcpyDeviceToHost);
  printf("series branch u0 < 2^-10: max abs err of r %.3e, max rel %.3e\n", h[0], h[1]);
  cudaMemset(d, 0, 16 * 8);
  sincos_err<<<(1 << 23) / 256, 256>>>(d);
  cudaMemcpy(h, d, 16 * 8, cudaMemcpyDeviceToHost);
  printf("sin.approx / cos.approx of angle = 6.28318530718f*u1: max abs err vs sin/cos(angle) %.3e / %.3e; vs sin/cos(2 pi u1) %.3e / %.3e\n",
         h[0], h[1], h[2], h[3]);
  return 0;
}
