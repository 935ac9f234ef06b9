// Do H2D and D2H overlap?  cudaMalloc vs cudaMallocManaged destinations, two streams.
// nvcc -O2 -o /tmp/copy_duplex scripts/micro/copy_duplex.cu && /tmp/copy_duplex
#include <cstdio>
#include <chrono>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
static double now() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  const size_t B = 1ull << 30;
  void *hin, *hout; CK(cudaMallocHost(&hin, B)); CK(cudaMallocHost(&hout, B));
  cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
  for (int managed = 0; managed < 2; managed++) {
    void *d1, *d2;
    if (managed) { CK(cudaMallocManaged(&d1, B)); CK(cudaMallocManaged(&d2, B)); CK(cudaMemPrefetchAsync(d1, B, 0, s1)); CK(cudaMemPrefetchAsync(d2, B, 0, s1)); }
    else { CK(cudaMalloc(&d1, B)); CK(cudaMalloc(&d2, B)); }
    CK(cudaMemsetAsync(d1, 0, B, s1)); CK(cudaMemsetAsync(d2, 0, B, s1)); CK(cudaDeviceSynchronize());
    for (int rep = 0; rep < 2; rep++) {
      double t0 = now(); CK(cudaMemcpyAsync(d1, hin, B, cudaMemcpyHostToDevice, s1)); CK(cudaDeviceSynchronize()); double t1 = now();
      CK(cudaMemcpyAsync(hout, d2, B, cudaMemcpyDeviceToHost, s2)); CK(cudaDeviceSynchronize()); double t2 = now();
      CK(cudaMemcpyAsync(d1, hin, B, cudaMemcpyHostToDevice, s1)); CK(cudaMemcpyAsync(hout, d2, B, cudaMemcpyDeviceToHost, s2)); CK(cudaDeviceSynchronize()); double t3 = now();
      printf("%s rep %d: H2D %.2f ms (%.1f GB/s)  D2H %.2f ms (%.1f GB/s)  both %.2f ms\n", managed ? "managed" : "cudaMalloc", rep, t1 - t0, B / (t1 - t0) / 1e6, t2 - t1, B / (t2 - t1) / 1e6, t3 - t2);
    }
    cudaFree(d1); cudaFree(d2);
  }
  return 0;
}
