// Micro-benchmark: streaming kernel variants for 1R1W (scale) and 2R1W (add) on 2^28 floats,
// cudaMalloc vs cudaMallocManaged(+prefetch).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <functional>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

template <int BLOCK, int UNROLL, int HINT>
__global__ void __launch_bounds__(BLOCK) k_scale(const float4* __restrict__ a, float4* __restrict__ o, size_t nvec, float s) {
  const size_t tile = (size_t)BLOCK * UNROLL;
  const size_t ntiles = (nvec + tile - 1) / tile;
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const size_t base = t * tile + threadIdx.x;
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) { size_t i = base + (size_t)u * BLOCK; if (i < nvec) v[u] = HINT ? __ldcs(a + i) : a[i]; }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      size_t i = base + (size_t)u * BLOCK;
      if (i < nvec) { float4 r = make_float4(v[u].x * s, v[u].y * s, v[u].z * s, v[u].w * s); if (HINT) __stcs(o + i, r); else o[i] = r; }
    }
  }
}

template <int BLOCK, int UNROLL, int HINT>
__global__ void __launch_bounds__(BLOCK) k_add(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ o, size_t nvec) {
  const size_t tile = (size_t)BLOCK * UNROLL;
  const size_t ntiles = (nvec + tile - 1) / tile;
  for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const size_t base = t * tile + threadIdx.x;
    float4 v[UNROLL], w[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) { size_t i = base + (size_t)u * BLOCK; if (i < nvec) { v[u] = HINT ? __ldcs(a + i) : a[i]; w[u] = HINT ? __ldcs(b + i) : b[i]; } }
#pragma unroll
    for (int u = 0; u < UNROLL; u++) {
      size_t i = base + (size_t)u * BLOCK;
      if (i < nvec) { float4 r = make_float4(v[u].x + w[u].x, v[u].y + w[u].y, v[u].z + w[u].z, v[u].w + w[u].w); if (HINT) __stcs(o + i, r); else o[i] = r; }
    }
  }
}

// one tile per CTA, no grid-stride loop
template <int BLOCK, int UNROLL>
__global__ void __launch_bounds__(BLOCK) k_scale_flat(const float4* __restrict__ a, float4* __restrict__ o, size_t nvec, float s) {
  const size_t base = (size_t)blockIdx.x * BLOCK * UNROLL + threadIdx.x;
  float4 v[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; u++) { size_t i = base + (size_t)u * BLOCK; if (i < nvec) v[u] = a[i]; }
#pragma unroll
  for (int u = 0; u < UNROLL; u++) { size_t i = base + (size_t)u * BLOCK; if (i < nvec) o[i] = make_float4(v[u].x * s, v[u].y * s, v[u].z * s, v[u].w * s); }
}

static float time_it(cudaStream_t st, int reps, const std::function<void()>& f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) f();
  CK(cudaStreamSynchronize(st));
  cudaEventRecord(e0, st);
  for (int i = 0; i < reps; i++) f();
  cudaEventRecord(e1, st);
  CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}

int main() {
  const size_t n = 1ull << 28, nvec = n / 4, bytes = n * 4;
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int managed = 0; managed < 2; managed++) {
    float *a, *b, *o;
    if (managed) {
      CK(cudaMallocManaged(&a, bytes)); CK(cudaMallocManaged(&b, bytes)); CK(cudaMallocManaged(&o, bytes));
      CK(cudaMemPrefetchAsync(a, bytes, 0, st)); CK(cudaMemPrefetchAsync(b, bytes, 0, st)); CK(cudaMemPrefetchAsync(o, bytes, 0, st));
    } else { CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&o, bytes)); }
    CK(cudaMemsetAsync(a, 0, bytes, st)); CK(cudaMemsetAsync(b, 0, bytes, st)); CK(cudaMemsetAsync(o, 0, bytes, st));
    CK(cudaStreamSynchronize(st));
    const char* mem = managed ? "managed" : "device ";
    auto rep = [&](const char* name, float ms, int nbuf) { printf("%s %-34s %.4f ms  %7.1f GB/s\n", mem, name, ms, nbuf * (double)bytes / ms / 1e6); };
    rep("cudaMemcpyAsync D2D", time_it(st, 20, [&] { cudaMemcpyAsync(o, a, bytes, cudaMemcpyDeviceToDevice, st); }), 2);
#define SCALE(B, U, H, G) rep("scale B" #B " U" #U " hint" #H " grid*" #G, time_it(st, 20, [&] { k_scale<B, U, H><<<sms * G, B, 0, st>>>((float4*)a, (float4*)o, nvec, 2.5f); }), 2)
#define ADD(B, U, H, G) rep("add   B" #B " U" #U " hint" #H " grid*" #G, time_it(st, 20, [&] { k_add<B, U, H><<<sms * G, B, 0, st>>>((float4*)a, (float4*)b, (float4*)o, nvec); }), 3)
    SCALE(256, 4, 0, 8); SCALE(256, 4, 1, 8); SCALE(256, 8, 0, 8); SCALE(256, 8, 1, 4); SCALE(256, 2, 0, 8); SCALE(256, 4, 0, 16);
    SCALE(512, 4, 0, 4); SCALE(512, 2, 0, 4); SCALE(1024, 2, 0, 2); SCALE(128, 4, 0, 16); SCALE(128, 8, 0, 16); SCALE(256, 4, 0, 32); SCALE(256, 1, 0, 8);
    rep("scale flat B256 U4", time_it(st, 20, [&] { k_scale_flat<256, 4><<<(unsigned)((nvec + 1023) / 1024), 256, 0, st>>>((float4*)a, (float4*)o, nvec, 2.5f); }), 2);
    rep("scale flat B256 U8", time_it(st, 20, [&] { k_scale_flat<256, 8><<<(unsigned)((nvec + 2047) / 2048), 256, 0, st>>>((float4*)a, (float4*)o, nvec, 2.5f); }), 2);
    rep("scale flat B512 U2", time_it(st, 20, [&] { k_scale_flat<512, 2><<<(unsigned)((nvec + 1023) / 1024), 512, 0, st>>>((float4*)a, (float4*)o, nvec, 2.5f); }), 2);
    rep("scale flat B1024 U1", time_it(st, 20, [&] { k_scale_flat<1024, 1><<<(unsigned)((nvec + 1023) / 1024), 1024, 0, st>>>((float4*)a, (float4*)o, nvec, 2.5f); }), 2);
    ADD(256, 4, 0, 8); ADD(256, 4, 1, 8); ADD(256, 8, 0, 4); ADD(256, 2, 0, 8); ADD(512, 2, 0, 4); ADD(128, 4, 0, 16); ADD(256, 4, 0, 16);
    cudaFree(a); cudaFree(b); cudaFree(o);
  }
  return 0;
}
