// Micro-benchmark: random 4-byte gather (2^26 indices into a 256 MiB table) under different L2 fetch
// granularities (cudaLimitMaxL2FetchGranularity 32 / 64 / 128) and loads in flight per thread.
// ncu (profiles/r02_ncu_rows.md) shows the product kernel reading 93 B of DRAM per gathered element:
// the question is whether the sector over-fetch can be switched off.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_gran scripts/micro/gather_gran.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1);} } while (0)

template <int UNROLL, int MODE>
__global__ void __launch_bounds__(256) gather_k(const float* __restrict__ a, const uint4* __restrict__ iv, float4* __restrict__ cv, size_t nvec) {
  const size_t stride = (size_t)gridDim.x * 256 * UNROLL;
  for (size_t base = (size_t)blockIdx.x * 256 * UNROLL + threadIdx.x; base < nvec; base += stride) {
    uint4 ix[UNROLL];
    float4 r[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) ix[u] = iv[base + (size_t)u * 256];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      if (MODE == 0) {
        r[u].x = __ldg(a + ix[u].x); r[u].y = __ldg(a + ix[u].y); r[u].z = __ldg(a + ix[u].z); r[u].w = __ldg(a + ix[u].w);
      } else if (MODE == 1) {      // no L1 allocation, streaming
        asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r[u].x) : "l"(a + ix[u].x));
        asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r[u].y) : "l"(a + ix[u].y));
        asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r[u].z) : "l"(a + ix[u].z));
        asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r[u].w) : "l"(a + ix[u].w));
      } else {                     // evict-first in L2
        r[u].x = __ldcs(a + ix[u].x); r[u].y = __ldcs(a + ix[u].y); r[u].z = __ldcs(a + ix[u].z); r[u].w = __ldcs(a + ix[u].w);
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) cv[base + (size_t)u * 256] = r[u];
  }
}

__global__ void fill_idx(uint32_t* p, size_t n, uint32_t mask) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t x = (uint32_t)i * 2654435761u + 12345u;
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    p[i] = x & mask;
  }
}

template <int UNROLL, int MODE>
void run(const char* name, const float* a, const uint32_t* idx, float* c, size_t n, int grid_mult) {
  const unsigned grid = 148 * grid_mult;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; i++) gather_k<UNROLL, MODE><<<grid, 256>>>(a, (const uint4*)idx, (float4*)c, n / 4);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  for (int i = 0; i < 10; i++) gather_k<UNROLL, MODE><<<grid, 256>>>(a, (const uint4*)idx, (float4*)c, n / 4);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
  printf("  %-34s grid %5u  %7.4f ms  %7.1f GB/s algorithmic (12 B/elem)\n", name, grid, ms, (double)n * 12 / ms / 1e6);
}

int main() {
  const size_t n = (size_t)1 << 26, tn = (size_t)1 << 26;
  float *a, *c; uint32_t* idx;
  CK(cudaMalloc(&a, tn * 4)); CK(cudaMalloc(&c, n * 4)); CK(cudaMalloc(&idx, n * 4));
  CK(cudaMemset(a, 0, tn * 4));
  fill_idx<<<1184, 256>>>(idx, n, (uint32_t)(tn - 1));
  CK(cudaDeviceSynchronize());
  const int grans[4] = {0, 32, 64, 128};
  for (int g = 0; g < 4; g++) {
    if (grans[g]) {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, grans[g]);
      if (e != cudaSuccess) { printf("set limit %d: %s\n", grans[g], cudaGetErrorString(e)); continue; }
    }
    size_t v = 0; cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity);
    printf("cudaLimitMaxL2FetchGranularity = %zu%s\n", v, grans[g] ? "" : " (default)");
    run<2, 0>("ldg unroll 2 (product kernel)", a, idx, c, n, 8);
    run<4, 0>("ldg unroll 4", a, idx, c, n, 8);
    run<8, 0>("ldg unroll 8", a, idx, c, n, 4);
    run<4, 1>("L1::no_allocate unroll 4", a, idx, c, n, 8);
    run<4, 2>("ld.cs unroll 4", a, idx, c, n, 8);
    run<4, 0>("ldg unroll 4, 16 CTAs/SM", a, idx, c, n, 16);
  }
  return 0;
}
