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

#include <algorithm>
#include <cstdlib>

int vkp_gemm_tc_supported(int transA, int transB, uint32_t M, uint32_t N, uint32_t K, const float* A,
                          const float* B, float* C, int forced);
int vkp_gemm_tc(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K, const float* A,
                const float* B, float* C, const float* bias, int accumulate, vkp_gemm_post post);

#include "vkp_math.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

__device__ __forceinline__ float post1(float v, const vkp_gemm_post& p, size_t idx) {
  if (p.relu) v = fmaxf(v, 0.0f);
  if (p.mask) v = fmaxf(vkpm::sign_f(p.mask[idx]), 0.0f) * v;
  return v;
}

// C[m,n] (+)= sum_k a(m,k) b(k,n) + bias[n]
template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                 const float* __restrict__ bias, uint32_t M, uint32_t N, uint32_t K, int accumulate,
                 uint32_t k_per_split, vkp_gemm_post post) {
  // blockIdx.z selects a K slice (split-K for skinny problems); with more than one slice C is a
  // [splits][M][N] partial buffer and bias / accumulate are applied by simt_splitk_reduce.
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const uint32_t tid = threadIdx.x;
  const uint32_t tx = tid % 16, ty = tid / 16;
  // M tiles on grid.x (2^31 limit), N tiles on grid.y: tall matrices (M > 65535 * 64) stay legal
  const uint32_t m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;
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
      if (accumulate) v = *c + v;
      *c = post1(v, post, (size_t)gm * N + gn);
    }
  }
}

// many partials, few outputs (the skinny weight gradient: ~148 partials of a [16, 1024] result): 8 lanes per
// output element, lane g folds partials g, g + 8, ... in order, then the 8 lanes are folded in a fixed xor tree
__global__ void __launch_bounds__(256)
simt_splitk_reduce_wide(const float* __restrict__ part, float* __restrict__ C, const float* __restrict__ bias, uint32_t M,
                        uint32_t N, uint32_t splits, int accumulate, vkp_gemm_post post) {
  const size_t mn = (size_t)M * N;
  const uint32_t g = threadIdx.x & 7;
  for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 3; i < ((mn + 31) & ~(size_t)31);
       i += ((size_t)gridDim.x * blockDim.x) >> 3) {
    float acc = 0.f;
    if (i < mn)
      for (uint32_t s = g; s < splits; s += 8) acc += part[s * mn + i];
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if (g == 0 && i < mn) {
      if (bias) acc += bias[i % N];
      if (accumulate) acc = C[i] + acc;
      C[i] = post1(acc, post, i);
    }
  }
}

__global__ void __launch_bounds__(256)
simt_splitk_reduce(const float* __restrict__ part, float* __restrict__ C, const float* __restrict__ bias, uint32_t M,
                   uint32_t N, uint32_t splits, int accumulate, vkp_gemm_post post) {
  const size_t mn = (size_t)M * N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < mn; i += (size_t)gridDim.x * blockDim.x) {
    float acc = part[i];
    for (uint32_t s = 1; s < splits; s++) acc += part[s * mn + i];   // fixed order: deterministic
    if (bias) acc += bias[i % N];
    if (accumulate) acc = C[i] + acc;
    C[i] = post1(acc, post, i);
  }
}

// ---- skinny shapes of nn.Dense with few classes (config 5: 1024 -> 16) -------------------------
// The 64x64-tile kernel above wastes 3/4 of its tile on a 16-wide operand and the tensor-core
// kernel needs >= 128 on both sides; these three contractions are bandwidth problems (one pass
// over a [batch, 1024] matrix) and get one streaming kernel each
// (profiles/r01_launches_mlp_step_v2.md: 62 + 49 + 17 us of a 494 us step before).

// forward:  C[M, N] (+)= A[M, K] B[N, K]^T + bias[N],  N <= 32, K % 4 == 0, B (N*K floats) in shared memory.
// One warp per row of A at a time: lane l owns k = 4l, 4l+1, .. (float4, stride 128), N accumulators
// per lane, butterfly-reduced; lane n writes C[m, n].
template <int NMAX>
__global__ void __launch_bounds__(256)
gemm_skinny_n_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                     const float* __restrict__ bias, uint32_t M, uint32_t N, uint32_t K, int accumulate,
                     vkp_gemm_post post) {
  extern __shared__ float4 sk_b[];                  // [N][K/4]
  const uint32_t k4 = K / 4;
  for (uint32_t i = threadIdx.x; i < N * k4; i += 256) sk_b[i] = reinterpret_cast<const float4*>(B)[i];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t warps = gridDim.x * 8;
  for (uint32_t m = blockIdx.x * 8 + warp; m < M; m += warps) {
    const float4* a = reinterpret_cast<const float4*>(A + (size_t)m * K);
    float acc[NMAX];
#pragma unroll
    for (int n = 0; n < NMAX; n++) acc[n] = 0.f;
    for (uint32_t j = lane; j < k4; j += 32) {
      const float4 x = a[j];
#pragma unroll
      for (int n = 0; n < NMAX; n++) {
        if (n < (int)N) {
          const float4 w = sk_b[n * k4 + j];
          acc[n] = fmaf(x.x, w.x, acc[n]);
          acc[n] = fmaf(x.y, w.y, acc[n]);
          acc[n] = fmaf(x.z, w.z, acc[n]);
          acc[n] = fmaf(x.w, w.w, acc[n]);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NMAX; n++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc[n] += __shfl_xor_sync(0xffffffffu, acc[n], o);
    }
    // every lane now holds all N sums; lane n stores column n
    float v = 0.f;
#pragma unroll
    for (int n = 0; n < NMAX; n++)
      if ((int)lane == n) v = acc[n];
    if (lane < N) {
      if (bias) v += bias[lane];
      float* c = C + (size_t)m * N + lane;
      if (accumulate) v = *c + v;
      *c = post1(v, post, (size_t)m * N + lane);
    }
  }
}

// weight gradient:  part[split][M, N] = sum_{k in split} A[k, M]^T B[k, N],  M <= 16, N % 4 == 0.
// Block = 256 column threads (4 neighbouring columns of B each) x 2 k-groups; the block's slice of A
// ([k_per_split, M], a few KB) is staged in shared memory once and read back as broadcast float4s
// (the first version issued M scalar loads per k and thread: LSU-bound, 32 us for a 32 MB pass), B
// streams through 16-byte loads, the two k-groups are folded through shared memory, and one partial
// per block goes to the workspace; simt_splitk_reduce folds the partials (fixed order) and applies
// bias / accumulate.
template <int MMAX>
__global__ void __launch_bounds__(512)
gemm_skinny_m_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ part,
                     uint32_t M, uint32_t N, uint32_t K, uint32_t k_per_split) {
  extern __shared__ float sk_a[];                              // [k_per_split][MMAX], then 256 float4 of scratch
  const uint32_t tx = threadIdx.x & 255, ty = threadIdx.x >> 8;
  const uint32_t n4 = blockIdx.x * 256 + tx;                   // float4 column index
  const uint32_t k0 = blockIdx.y * k_per_split;
  const uint32_t k1 = (k0 + k_per_split < K) ? k0 + k_per_split : K;
  for (uint32_t i = threadIdx.x; i < (k1 - k0) * MMAX; i += 512) {
    const uint32_t r = i / MMAX, m = i % MMAX;
    sk_a[i] = m < M ? A[(size_t)(k0 + r) * M + m] : 0.f;
  }
  __syncthreads();
  float4 acc[MMAX];
#pragma unroll
  for (int m = 0; m < MMAX; m++) acc[m] = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool live = n4 * 4 < N;
  if (live) {
    for (uint32_t k = k0 + ty; k < k1; k += 2) {
      const float4 b = *reinterpret_cast<const float4*>(B + (size_t)k * N + n4 * 4);
      const float4* a4 = reinterpret_cast<const float4*>(sk_a + (size_t)(k - k0) * MMAX);
#pragma unroll
      for (int q = 0; q < MMAX / 4; q++) {
        const float4 a = a4[q];                                // same address for the whole warp: broadcast
        const float am[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
          float4& c = acc[4 * q + j];
          c.x = fmaf(am[j], b.x, c.x);
          c.y = fmaf(am[j], b.y, c.y);
          c.z = fmaf(am[j], b.z, c.z);
          c.w = fmaf(am[j], b.w, c.w);
        }
      }
    }
  }
  float4* red = reinterpret_cast<float4*>(sk_a + (size_t)k_per_split * MMAX);
  float* out = part + (size_t)blockIdx.y * M * N;
#pragma unroll
  for (int m = 0; m < MMAX; m++) {
    if (m < (int)M) {                                          // M is uniform: every thread takes the same path
      if (ty == 1) red[tx] = acc[m];
      __syncthreads();
      if (ty == 0 && live) {
        const float4 o = red[tx];
        *reinterpret_cast<float4*>(out + (size_t)m * N + n4 * 4) =
            make_float4(acc[m].x + o.x, acc[m].y + o.y, acc[m].z + o.z, acc[m].w + o.w);
      }
      __syncthreads();
    }
  }
}

// input gradient:  C[M, N] (+)= A[M, K] B[K, N],  K <= 16, N % 4 == 0.  Thread = 4 neighbouring
// columns; its K x 4 slice of B stays in registers while it walks the rows of A.
template <int KMAX>
__global__ void __launch_bounds__(256)
gemm_skinny_k_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
                     uint32_t M, uint32_t N, uint32_t K, int accumulate, uint32_t rows_per_block, vkp_gemm_post post) {
  const uint32_t n4 = blockIdx.x * 256 + threadIdx.x;
  if (n4 * 4 >= N) return;
  float4 b[KMAX];
#pragma unroll
  for (int k = 0; k < KMAX; k++)
    b[k] = (k < (int)K) ? *reinterpret_cast<const float4*>(B + (size_t)k * N + n4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const uint32_t m0 = blockIdx.y * rows_per_block;
  const uint32_t m1 = (m0 + rows_per_block < M) ? m0 + rows_per_block : M;
  const bool avec = (K % 4 == 0) && ((((uintptr_t)A) & 15) == 0);
  for (uint32_t m = m0; m < m1; m++) {
    const float* a = A + (size_t)m * K;                          // warp-uniform address: broadcast
    float ar[KMAX];
    if (avec) {                                                  // K / 4 16-byte loads instead of K scalar ones
#pragma unroll
      for (int q = 0; q < KMAX / 4; q++) {
        const float4 v = (4 * q < (int)K) ? __ldg(reinterpret_cast<const float4*>(a) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        ar[4 * q] = v.x; ar[4 * q + 1] = v.y; ar[4 * q + 2] = v.z; ar[4 * q + 3] = v.w;
      }
    } else {
#pragma unroll
      for (int k = 0; k < KMAX; k++) ar[k] = (k < (int)K) ? __ldg(a + k) : 0.f;
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < KMAX; k++) {
      if (k < (int)K) {
        const float ak = ar[k];
        acc.x = fmaf(ak, b[k].x, acc.x);
        acc.y = fmaf(ak, b[k].y, acc.y);
        acc.z = fmaf(ak, b[k].z, acc.z);
        acc.w = fmaf(ak, b[k].w, acc.w);
      }
    }
    float4* c = reinterpret_cast<float4*>(C + (size_t)m * N + n4 * 4);
    if (accumulate) {
      const float4 o = *c;
      acc.x = o.x + acc.x; acc.y = o.y + acc.y; acc.z = o.z + acc.z; acc.w = o.w + acc.w;
    }
    if (post.relu) { acc.x = fmaxf(acc.x, 0.f); acc.y = fmaxf(acc.y, 0.f); acc.z = fmaxf(acc.z, 0.f); acc.w = fmaxf(acc.w, 0.f); }
    if (post.mask) {    // N % 4 == 0 and 16-byte aligned like C
      const float4 y = *reinterpret_cast<const float4*>(post.mask + (size_t)m * N + n4 * 4);
      acc.x = fmaxf(vkpm::sign_f(y.x), 0.f) * acc.x; acc.y = fmaxf(vkpm::sign_f(y.y), 0.f) * acc.y;
      acc.z = fmaxf(vkpm::sign_f(y.z), 0.f) * acc.z; acc.w = fmaxf(vkpm::sign_f(y.w), 0.f) * acc.w;
    }
    *c = acc;
  }
}

// returns 1 and launches when one of the skinny kernels takes the problem, 0 otherwise, < 0 on error
int gemm_skinny(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K, const float* A,
                const float* B, float* C, const float* bias, int accumulate, vkp_gemm_post post) {
  static const bool off = getenv("VKP_DISABLE_SKINNY") != nullptr;
  if (off) return 0;
  const bool aligned = ((((uintptr_t)A) | ((uintptr_t)B) | ((uintptr_t)C) | ((uintptr_t)post.mask)) & 15) == 0;
  if (!aligned || (uint64_t)M * N * K < (1ull << 22)) return 0;       // small problems keep the tiled kernel
  if (!transA && transB && N <= 16 && K % 4 == 0 && K >= 128 && M >= 1024 && (size_t)N * K * 4 <= 96 * 1024) {
    const size_t smem = (size_t)N * K * 4;
    static bool attr[64] = {};      // the opt-in is per device (GPU(0) then GPU(1) in one process)
    if (!attr[ctx->device & 63]) {
      if (cudaFuncSetAttribute(gemm_skinny_n_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024) != cudaSuccess)
        return vkp_set_error("gemm_skinny: cannot raise the shared-memory limit") ? -1 : -1;
      attr[ctx->device & 63] = true;
    }
    const unsigned grid = (unsigned)std::min<uint64_t>((M + 7) / 8, (uint64_t)ctx->sms * 2);
    gemm_skinny_n_kernel<16><<<grid, 256, smem, ctx->stream>>>(A, B, C, bias, M, N, K, accumulate, post);
    return vkp_after_launch(ctx, "gemm_skinny_n") == VKP_OK ? 1 : -1;
  }
  if (transA && !transB && M <= 16 && N % 4 == 0 && N >= 256 && K >= 1024) {
    const uint32_t nblk = (N / 4 + 255) / 256;
    uint32_t splits = (uint32_t)std::max<uint64_t>(1, (uint64_t)ctx->sms / nblk);   // one 512-thread block per SM
    if (splits > K / 16) splits = K / 16;
    uint32_t k_per = (K + splits - 1) / splits;
    if (k_per > 512) k_per = 512;                      // A slice of at most 32 KiB in shared memory
    splits = (K + k_per - 1) / k_per;
    void* ws;
    if (vkp_workspace(ctx, 0, (size_t)splits * M * N * sizeof(float), &ws) != VKP_OK) return -1;
    float* part = static_cast<float*>(ws);
    const size_t smem_m = (size_t)k_per * 16 * sizeof(float) + 256 * sizeof(float4);
    gemm_skinny_m_kernel<16><<<dim3(nblk, splits), 512, smem_m, ctx->stream>>>(A, B, part, M, N, K, k_per);
    if (vkp_after_launch(ctx, "gemm_skinny_m") != VKP_OK) return -1;
    if (splits >= 16)
      simt_splitk_reduce_wide<<<vkp_grid_for(ctx, (size_t)M * N * 8, 256, 8), 256, 0, ctx->stream>>>(part, C, bias, M, N, splits, accumulate, post);
    else
      simt_splitk_reduce<<<vkp_grid_for(ctx, (size_t)M * N, 256, 8), 256, 0, ctx->stream>>>(part, C, bias, M, N, splits, accumulate, post);
    return vkp_after_launch(ctx, "gemm_skinny_m_reduce") == VKP_OK ? 1 : -1;
  }
  if (!transA && !transB && K <= 16 && N % 4 == 0 && N >= 256 && M >= 1024 && !bias) {
    const uint32_t nblk = (N / 4 + 255) / 256;
    uint32_t yb = (uint32_t)std::max<uint64_t>(1, (uint64_t)ctx->sms * 4 / nblk);
    uint32_t rows_per = (M + yb - 1) / yb;
    if (rows_per < 8) rows_per = 8;
    yb = (M + rows_per - 1) / rows_per;
    gemm_skinny_k_kernel<16><<<dim3(nblk, yb), 256, 0, ctx->stream>>>(A, B, C, M, N, K, accumulate, rows_per, post);
    return vkp_after_launch(ctx, "gemm_skinny_k") == VKP_OK ? 1 : -1;
  }
  return 0;
}

int gemm_simt(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K, const float* A,
              const float* B, float* C, const float* bias, int accumulate, vkp_gemm_post post) {
  dim3 grid((M + TM - 1) / TM, (N + TN - 1) / TN, 1);
  VKP_CHECK(grid.y <= 65535, "gemm_simt: N = %u is too wide for the fallback kernel (max %d columns)", N, 65535 * TN);
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
  const vkp_gemm_post kp = splits > 1 ? vkp_gemm_post{0, nullptr} : post;
  if (!transA && !transB) gemm_simt_kernel<false, false><<<grid, 256, 0, ctx->stream>>>(A, B, dst, kb, M, N, K, ka, k_per, kp);
  else if (!transA && transB) gemm_simt_kernel<false, true><<<grid, 256, 0, ctx->stream>>>(A, B, dst, kb, M, N, K, ka, k_per, kp);
  else if (transA && !transB) gemm_simt_kernel<true, false><<<grid, 256, 0, ctx->stream>>>(A, B, dst, kb, M, N, K, ka, k_per, kp);
  else gemm_simt_kernel<true, true><<<grid, 256, 0, ctx->stream>>>(A, B, dst, kb, M, N, K, ka, k_per, kp);
  VKP_TRY(vkp_after_launch(ctx, "gemm_simt"));
  if (splits > 1) {
    simt_splitk_reduce<<<vkp_grid_for(ctx, (size_t)M * N, 256, 8), 256, 0, ctx->stream>>>(dst, C, bias, M, N, splits,
                                                                                        accumulate, post);
    VKP_TRY(vkp_after_launch(ctx, "gemm_simt_splitk_reduce"));
  }
  return VKP_OK;
}

}  // namespace

int vkp_launch_gemm(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K,
                    const float* A, const float* B, float* C, const float* bias, int flags, vkp_gemm_post post) {
  if (M == 0 || N == 0) return VKP_OK;
  const int accumulate = (flags & VKP_GEMM_ACCUMULATE) ? 1 : 0;
  const bool force_simt = flags & VKP_GEMM_FORCE_SIMT, force_tc = flags & VKP_GEMM_FORCE_TC;
  const bool tc_ok = !force_simt && vkp_gemm_tc_supported(transA, transB, M, N, K, A, B, C, force_tc);
  VKP_CHECK(!(force_tc && !tc_ok), "vkp_gemm: tensor-core path forced but the shape (%u,%u,%u) is not supported", M, N, K);
  if (flags & VKP_GEMM_RELU) post.relu = 1;
  if (tc_ok) return vkp_gemm_tc(ctx, transA, transB, M, N, K, A, B, C, bias, accumulate, post);
  if (!force_simt) {
    const int r = gemm_skinny(ctx, transA, transB, M, N, K, A, B, C, bias, accumulate, post);
    if (r < 0) return VKP_ERR;
    if (r > 0) return VKP_OK;
  }
  return gemm_simt(ctx, transA, transB, M, N, K, A, B, C, bias, accumulate, post);
}

extern "C" int vkp_gemm_fused(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K,
                              const float* A, const float* B, float* C, const float* bias, const float* relu_mask,
                              int flags, vkp_job** job) {
  VKP_RANGE("vkp_gemm");
  VKP_CHECK(ctx && A && B && C, "vkp_gemm: null argument");
  VKP_CHECK(!relu_mask || (N % 4 == 0 && (((uintptr_t)relu_mask) & 15) == 0), "vkp_gemm_fused: mask needs N %% 4 == 0 and 16-byte alignment");
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[5] = {(void*)A, (void*)B, (void*)C, (void*)bias, (void*)relu_mask};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 5));
  VKP_TRY(vkp_launch_gemm(ctx, transA, transB, M, N, K, A, B, C, bias, flags, vkp_gemm_post{0, relu_mask}));
  return vkp_finish_op(ctx, job);
}

extern "C" int vkp_gemm(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K,
                        const float* A, const float* B, float* C, const float* bias, int flags,
                        vkp_job** job) {
  return vkp_gemm_fused(ctx, transA, transB, M, N, K, A, B, C, bias, nullptr, flags, job);
}
