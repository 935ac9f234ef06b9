// vkp_gemm.cu -- float32 contraction behind `@` and nn.Dense.
//
// Replaces shader/matmul.comp:23-33 (C[M,N] = A[M,K] B[K,N], one thread per C element, serial
// fp32 loop over K) and shader/batch_affine.comp:25-39 (Y[B,out] = X[B,in] W[out,in]^T + b).
// Dense.backward of the reference builds a B x out x in temporary and sums it
// (nn/layers.py:126-141); here dW = dy^T x and dx = dy W are two more calls of the same GEMM.
//
// Two implementations behind one entry point (vkp_launch_gemm):
//   * gemm_simt: register-tiled fp32 FFMA kernel, any shape / any transposition.  It is the
//     path for small or unaligned problems and the on-device cross-check of the tensor path.
//   * gemm_tc (vkp_gemm_tc.cu): tcgen05 3xTF32 kernel with TMEM accumulators fed by TMA, used
//     when the shape is tile-aligned.
#include "vkp_common.cuh"

int vkp_gemm_tc_supported(int transA, int transB, uint32_t M, uint32_t N, uint32_t K, const float* A,
                          const float* B, float* C, int forced);
int vkp_gemm_tc(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K, const float* A,
                const float* B, float* C, const float* bias, int accumulate);

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

// C[m,n] (+)= sum_k a(m,k) b(k,n) + bias[n]
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                 const float* __restrict__ bias, uint32_t M, uint32_t N, uint32_t K, int accumulate,
                 uint32_t k_per_split) {
  // blockIdx.z selects a K slice (split-K for skinny problems); with more than one slice C is a
  // [splits][M][N] partial buffer and bias / accumulate are applied by simt_splitk_reduce.
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const uint32_t tid = threadIdx.x;
  const uint32_t tx = tid % 16, ty = tid / 16;
  const uint32_t m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  const uint32_t k_begin = blockIdx.z * k_per_split;
  const uint32_t k_end = (k_begin + k_per_split < K) ? k_begin + k_per_split : K;
  C += (size_t)blockIdx.z * M * N;
  for (uint32_t k0 = k_begin; k0 < k_end; k0 += TK) {
    // stage A tile (TM x TK) and B tile (TK x TN); the fast thread index follows the
    // contiguous direction of the source
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint32_t e = tid + r * 256;  // 0..1023
      uint32_t m, k;
      if (TA) { m = e % TM; k = e / TM; } else { k = e % TK; m = e / TK; }
      const uint32_t gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < k_end) v = TA ? A[(size_t)gk * M + gm] : A[(size_t)gm * K + gk];
      As[k][m] = v;
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const uint32_t e = tid + r * 256;
      uint32_t n, k;
      if (TB) { k = e % TK; n = e / TK; } else { n = e % TN; k = e / TN; }
      const uint32_t gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < k_end) v = TB ? B[(size_t)gn * K + gk] : B[(size_t)gk * N + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; k++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const uint32_t gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      float* c = C + (size_t)gm * N + gn;
      *c = accumulate ? (*c + v) : v;
    }
  }
}

__global__ void __launch_bounds__(256)
simt_splitk_reduce(const float* __restrict__ part, float* __restrict__ C, const float* __restrict__ bias, uint32_t M,
                   uint32_t N, uint32_t splits, int accumulate) {
  const size_t mn = (size_t)M * N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < mn; i += (size_t)gridDim.x * blockDim.x) {
    float acc = part[i];
    for (uint32_t s = 1; s < splits; s++) acc += part[s * mn + i];   // fixed order: deterministic
    if (bias) acc += bias[i % N];
    C[i] = accumulate ? (C[i] + acc) : acc;
  }
}

int gemm_simt(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K, const float* A,
              const float* B, float* C, const float* bias, int accumulate) {
  dim3 grid((N + TN - 1) / TN, (M + TM - 1) / TM, 1);
  VKP_CHECK(grid.y <= 65535, "gemm_simt: M too large for the fallback kernel");
  // skinny problems (few output tiles, long K): split K over blockIdx.z so that every SM has work
  uint32_t splits = 1;
  const uint64_t tiles = (uint64_t)grid.x * grid.y;
  if (tiles < 2ull * ctx->sms && K >= 512) {
    uint64_t want = (4ull * ctx->sms + tiles - 1) / tiles;
    if (want > K / 128) want = K / 128;
    if (want > 256) want = 256;
    splits = (uint32_t)(want < 1 ? 1 : want);
  }
  uint32_t k_per = (K + splits - 1) / splits;
  k_per = (k_per + TK - 1) / TK * TK;
  splits = K ? (K + k_per - 1) / k_per : 1;
  if (splits < 1) splits = 1;
  grid.z = splits;
  float* dst = C;
  if (splits > 1) {
    void* ws;
    VKP_TRY(vkp_workspace(ctx, 0, (size_t)splits * M * N * sizeof(float), &ws));
    dst = static_cast<float*>(ws);
  }
  const float* kb = splits > 1 ? nullptr : bias;
  const int ka = splits > 1 ? 0 : accumulate;
  if (!transA && !transB) gemm_simt_kernel<false, false><<<grid, 256, 0, ctx->stream>>>(A, B, dst, kb, M, N, K, ka, k_per);
  else if (!transA && transB) gemm_simt_kernel<false, true><<<grid, 256, 0, ctx->stream>>>(A, B, dst, kb, M, N, K, ka, k_per);
  else if (transA && !transB) gemm_simt_kernel<true, false><<<grid, 256, 0, ctx->stream>>>(A, B, dst, kb, M, N, K, ka, k_per);
  else gemm_simt_kernel<true, true><<<grid, 256, 0, ctx->stream>>>(A, B, dst, kb, M, N, K, ka, k_per);
  VKP_TRY(vkp_after_launch(ctx, "gemm_simt"));
  if (splits > 1) {
    simt_splitk_reduce<<<vkp_grid_for(ctx, (size_t)M * N, 256, 8), 256, 0, ctx->stream>>>(dst, C, bias, M, N, splits,
                                                                                        accumulate);
    VKP_TRY(vkp_after_launch(ctx, "gemm_simt_splitk_reduce"));
  }
  return VKP_OK;
}

}  // namespace

int vkp_launch_gemm(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K,
                    const float* A, const float* B, float* C, const float* bias, int flags) {
  if (M == 0 || N == 0) return VKP_OK;
  const int accumulate = (flags & VKP_GEMM_ACCUMULATE) ? 1 : 0;
  const bool force_simt = flags & VKP_GEMM_FORCE_SIMT, force_tc = flags & VKP_GEMM_FORCE_TC;
  const bool tc_ok = !force_simt && vkp_gemm_tc_supported(transA, transB, M, N, K, A, B, C, force_tc);
  VKP_CHECK(!(force_tc && !tc_ok), "vkp_gemm: tensor-core path forced but the shape (%u,%u,%u) is not supported", M, N, K);
  if (tc_ok) return vkp_gemm_tc(ctx, transA, transB, M, N, K, A, B, C, bias, accumulate);
  return gemm_simt(ctx, transA, transB, M, N, K, A, B, C, bias, accumulate);
}

extern "C" int vkp_gemm(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K,
                        const float* A, const float* B, float* C, const float* bias, int flags,
                        vkp_job** job) {
  VKP_CHECK(ctx && A && B && C, "vkp_gemm: null argument");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[4] = {(void*)A, (void*)B, (void*)C, (void*)bias};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 4));
  VKP_TRY(vkp_launch_gemm(ctx, transA, transB, M, N, K, A, B, C, bias, flags));
  return vkp_finish_op(ctx, job);
}
