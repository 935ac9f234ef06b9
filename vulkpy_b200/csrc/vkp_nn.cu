// vkp_nn.cu -- fused kernels for the vulkpy.nn compositions.
//
// The reference builds these from chains of single-op shaders (Adam: 13 jobs and 5 temporaries
// per tensor, nn/optimizers.py:235-253; Softmax.forward: 5 jobs, nn/layers.py:297-300; the
// activation backward passes: 3 jobs each, nn/layers.py:207-210,262-267,320-323).  Each kernel
// here performs the SAME float32 operations in the SAME order with one rounding per reference
// operation (the library is compiled with -fmad=false), so results are bit-identical to the
// op-by-op path; only the HBM round trips between the operations disappear.
#include "vkp_common.cuh"
#include "vkp_math.cuh"
#include "vkp_tables.cuh"

namespace {

constexpr int NB = 256;
#define VKP_NN_MAX_TENSORS 16

// m = m*b1 + (1-b1)*g ; v = v*b2 + (1-b2)*g^2 ; diff = (m/(1-b1t)) * (-lr) / (sqrt(v/(1-b2t)) + eps)
// scalars arrive already rounded to float32 exactly as they cross the reference's boundary
// (_vkarray.cc:841-845).  g^2 is the correctly rounded square (what pow(g, 2.0) returns).
__global__ void __launch_bounds__(NB)
adam_kernel(const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, float* __restrict__ diff,
            size_t n, float b1, float omb1, float b2, float omb2, float c1, float c2, float eps, float neg_lr) {
  for (size_t i = blockIdx.x * (size_t)NB + threadIdx.x; i < n; i += (size_t)gridDim.x * NB) {
    const float gi = g[i];
    float mi = m[i] * b1;          // self.m *= beta1
    mi = mi + omb1 * gi;           // self.m += (1 - beta1) * grad
    float vi = v[i] * b2;          // self.v *= beta2
    vi = vi + omb2 * (gi * gi);    // self.v += (1 - beta2) * grad ** 2
    m[i] = mi;
    v[i] = vi;
    float mh = mi / c1;            // mhat = m / (1 - beta1t)
    float vh = vi / c2;            // vhat = v / (1 - beta2t)
    vh = __fsqrt_rn(vh);           // vhat.sqrt(inplace=True)
    vh = vh + eps;                 // vhat += eps
    mh = mh * neg_lr;              // mhat *= -lr
    diff[i] = mh / vh;             // mhat /= vhat
  }
}

// The optimizer step of a whole model in one launch (blockIdx.y = tensor): per element exactly
// adam_kernel followed by Parameter.update's `value += diff` (nn/parameters.py:88-95), diff rounded
// to float32 before the add as the op-by-op path does.
struct AdamMany {
  const float* g[VKP_NN_MAX_TENSORS];
  float* m[VKP_NN_MAX_TENSORS];
  float* v[VKP_NN_MAX_TENSORS];
  float* value[VKP_NN_MAX_TENSORS];
  unsigned long long n[VKP_NN_MAX_TENSORS];
  float s[VKP_NN_MAX_TENSORS][8];   // b1, 1-b1, b2, 1-b2, 1-b1^t, 1-b2^t, eps, -lr
};
__global__ void __launch_bounds__(NB) adam_apply_many_kernel(const __grid_constant__ AdamMany a) {
  const int t = blockIdx.y;
  const float* __restrict__ g = a.g[t];
  float* __restrict__ m = a.m[t];
  float* __restrict__ v = a.v[t];
  float* __restrict__ val = a.value[t];
  const size_t n = a.n[t];
  const float b1 = a.s[t][0], omb1 = a.s[t][1], b2 = a.s[t][2], omb2 = a.s[t][3];
  const float c1 = a.s[t][4], c2 = a.s[t][5], eps = a.s[t][6], neg_lr = a.s[t][7];
  for (size_t i = blockIdx.x * (size_t)NB + threadIdx.x; i < n; i += (size_t)gridDim.x * NB) {
    const float gi = g[i];
    float mi = m[i] * b1;
    mi = mi + omb1 * gi;
    float vi = v[i] * b2;
    vi = vi + omb2 * (gi * gi);
    m[i] = mi;
    v[i] = vi;
    float mh = mi / c1;
    float vh = vi / c2;
    vh = __fsqrt_rn(vh);
    vh = vh + eps;
    mh = mh * neg_lr;
    const float diff = mh / vh;
    val[i] = val[i] + diff;        // self.value += diff
  }
}

struct FillMany {
  uint32_t* p[VKP_NN_MAX_TENSORS];
  unsigned long long n[VKP_NN_MAX_TENSORS];
};
__global__ void __launch_bounds__(NB) fill_many_kernel(const __grid_constant__ FillMany a, uint32_t bits) {
  uint32_t* __restrict__ p = a.p[blockIdx.y];
  const size_t n = a.n[blockIdx.y];
  for (size_t i = blockIdx.x * (size_t)NB + threadIdx.x; i < n; i += (size_t)gridDim.x * NB) p[i] = bits;
}

// dx = (sign(y) max 0) * dy   (ReLU.backward, nn/layers.py:207-210)
__global__ void __launch_bounds__(NB)
relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, size_t n) {
  for (size_t i = blockIdx.x * (size_t)NB + threadIdx.x; i < n; i += (size_t)gridDim.x * NB)
    dx[i] = fmaxf(vkpm::sign_f(y[i]), 0.0f) * dy[i];
}

// dx = ((1 - y) * y) * dy      (Sigmoid.backward / Softmax.backward, nn/layers.py:262-267,320-323)
__global__ void __launch_bounds__(NB)
ymul_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, size_t n) {
  for (size_t i = blockIdx.x * (size_t)NB + threadIdx.x; i < n; i += (size_t)gridDim.x * NB) {
    const float yi = y[i];
    dx[i] = ((1.0f - yi) * yi) * dy[i];
  }
}

// Softmax.forward over axis 1 of [rows, cols] (nn/layers.py:297-300): X = x - max(x); X = exp(X);
// X /= sum(X).  One warp per row; max and the sum run in the reference's serial order
// k = 0..cols-1 per row when cols <= 32 (each lane holds one element, lane 0 folds them through
// shuffles), longer rows use per-lane partials + a shuffle tree.
__global__ void __launch_bounds__(NB)
softmax_fwd_kernel(const __grid_constant__ vkpm::MathCoef coef, const float* __restrict__ x, float* __restrict__ y,
                   uint32_t rows, uint32_t cols) {
  const vkpt::LaneTables tab(coef);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps_per_grid = gridDim.x * (NB / 32);
  const uint32_t nwarp_rows = (rows + warps_per_grid - 1) / warps_per_grid;
  for (uint32_t it = 0; it < nwarp_rows; it++) {   // warp-uniform trip count (shuffles inside)
    const uint32_t row = (blockIdx.x * (NB / 32) + (threadIdx.x >> 5)) + it * warps_per_grid;
    const bool live_row = row < rows;
    const float* xr = x + (size_t)(live_row ? row : 0) * cols;
    float* yr = y + (size_t)(live_row ? row : 0) * cols;
    // max
    float mx = -INFINITY;
    for (uint32_t k = lane; k < cols; k += 32) mx = fmaxf(mx, xr[k]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // exp(x - max), kept in registers for short rows, written to y otherwise
    float sum = 0.f;
    const uint32_t nk = (cols + 31) / 32;
    for (uint32_t j = 0; j < nk; j++) {
      const uint32_t k = j * 32 + lane;
      const bool ok = k < cols;
      bool sp = false;
      const float d = ok ? (xr[k] - mx) : 0.f;
      float e = vkpm::exp_core(d, tab, sp);
      if (sp) e = vkpm::exp_f(d);
      if (!ok) e = 0.f;
      if (live_row && ok) yr[k] = e;
      // serial fold of this group of 32 in index order (lane 0 accumulates)
      for (uint32_t l = 0; l < 32; l++) {
        const float el = __shfl_sync(0xffffffffu, e, l);
        if (j * 32 + l < cols) sum = sum + el;
      }
    }
    __syncwarp();
    for (uint32_t k = lane; k < cols; k += 32)
      if (live_row) yr[k] = yr[k] / sum;
  }
}

// The tail of a classifier's training step in one launch (Sequence.train with a Softmax last layer and
// CrossEntropyLoss): per row
//   p  = softmax(z)                                 exactly softmax_fwd_kernel above        (nn/layers.py:297-300)
//   L  = (-t) * log(p + 1e-8)                       nn_cross_entropy.comp:25 (the loss still reduces L by its own jobs)
//   g  = (-t) / (p + 1e-8);  g *= scale             nn_cross_entropy_backward.comp:25, losses.py:58-68 (1/batch for "mean")
//   dz = ((1 - p) * p) * g                          Softmax.backward, nn/layers.py:320-323
// -- the float32 operations of the five launches it replaces, in their order, one rounding each (-fmad=false),
// so p, L and dz are bit-identical to the op-by-op path.  One warp per row.
__global__ void __launch_bounds__(NB)
softmax_ce_train_kernel(const __grid_constant__ vkpm::MathCoef coef, const float* __restrict__ z,
                        const float* __restrict__ t, float* __restrict__ p, float* __restrict__ L,
                        float* __restrict__ dz, uint32_t rows, uint32_t cols, float scale, int has_scale) {
  const vkpt::LaneTables tab(coef);
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warps_per_grid = gridDim.x * (NB / 32);
  const uint32_t nwarp_rows = (rows + warps_per_grid - 1) / warps_per_grid;
  for (uint32_t it = 0; it < nwarp_rows; it++) {   // warp-uniform trip count (shuffles inside)
    const uint32_t row = (blockIdx.x * (NB / 32) + (threadIdx.x >> 5)) + it * warps_per_grid;
    const bool live_row = row < rows;
    const size_t off = (size_t)(live_row ? row : 0) * cols;
    const float* xr = z + off;
    float* yr = p + off;
    float mx = -INFINITY;
    for (uint32_t k = lane; k < cols; k += 32) mx = fmaxf(mx, xr[k]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    const uint32_t nk = (cols + 31) / 32;
    for (uint32_t j = 0; j < nk; j++) {
      const uint32_t k = j * 32 + lane;
      const bool ok = k < cols;
      bool sp = false;
      const float d = ok ? (xr[k] - mx) : 0.f;
      float e = vkpm::exp_core(d, tab, sp);
      if (sp) e = vkpm::exp_f(d);
      if (!ok) e = 0.f;
      if (live_row && ok) yr[k] = e;
      for (uint32_t l = 0; l < 32; l++) {
        const float el = __shfl_sync(0xffffffffu, e, l);
        if (j * 32 + l < cols) sum = sum + el;
      }
    }
    __syncwarp();
    for (uint32_t j = 0; j < nk; j++) {          // every lane runs the table-driven log together
      const uint32_t k = j * 32 + lane;
      const bool ok = live_row && k < cols;
      const float pk = ok ? yr[k] / sum : 1.0f;
      const float tk = ok ? t[off + k] : 0.0f;
      const float q = pk + 1e-8f;
      bool sp = false;
      float lg = vkpm::log_core(q, tab, sp);
      if (sp) lg = vkpm::log_f(q);
      float g = (-tk) / q;
      if (has_scale) g = g * scale;
      if (ok) {
        yr[k] = pk;
        L[off + k] = (-tk) * lg;
        dz[off + k] = ((1.0f - pk) * pk) * g;
      }
    }
  }
}

}  // namespace

#define NN_PROLOGUE(...)                                   \
  VKP_CHECK(ctx, "vkp_nn: null context");                  \
  VKP_TRY(vkp_make_current(ctx));                          \
  std::lock_guard<std::mutex> g_(ctx->mu);                 \
  void* bufs_[] = {__VA_ARGS__};                           \
  VKP_TRY(vkp_prepare_buffers(ctx, bufs_, (int)(sizeof(bufs_) / sizeof(void*))))

extern "C" int vkp_nn_adam(vkp_ctx* ctx, const float* grad, float* m, float* v, float* diff, size_t n,
                           float beta1, float one_minus_beta1, float beta2, float one_minus_beta2,
                           float one_minus_beta1t, float one_minus_beta2t, float eps, float neg_lr,
                           vkp_job** job) {
  VKP_RANGE(__func__);
  NN_PROLOGUE((void*)grad, m, v, diff);
  if (n) {
    adam_kernel<<<vkp_grid_for(ctx, n, NB, 16), NB, 0, ctx->stream>>>(grad, m, v, diff, n, beta1, one_minus_beta1, beta2,
                                                                      one_minus_beta2, one_minus_beta1t,
                                                                      one_minus_beta2t, eps, neg_lr);
    VKP_TRY(vkp_after_launch(ctx, "nn_adam"));
  }
  return vkp_finish_op(ctx, job);
}

extern "C" int vkp_nn_adam_apply_many(vkp_ctx* ctx, int n_tensors, const float* const* grad, float* const* m,
                                      float* const* v, float* const* value, const size_t* count,
                                      const float* scalars /* [n_tensors][8] */, vkp_job** job) {
  VKP_RANGE(__func__);
  VKP_CHECK(ctx && grad && m && v && value && count && scalars, "vkp_nn_adam_apply_many: null argument");
  VKP_CHECK(n_tensors >= 1 && n_tensors <= VKP_NN_MAX_TENSORS, "vkp_nn_adam_apply_many: 1..%d tensors", VKP_NN_MAX_TENSORS);
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g_(ctx->mu);
  void* bufs[4 * VKP_NN_MAX_TENSORS];
  AdamMany a;
  size_t most = 0;
  for (int t = 0; t < n_tensors; t++) {
    bufs[4 * t] = (void*)grad[t]; bufs[4 * t + 1] = m[t]; bufs[4 * t + 2] = v[t]; bufs[4 * t + 3] = value[t];
    a.g[t] = grad[t]; a.m[t] = m[t]; a.v[t] = v[t]; a.value[t] = value[t]; a.n[t] = count[t];
    for (int k = 0; k < 8; k++) a.s[t][k] = scalars[8 * t + k];
    if (count[t] > most) most = count[t];
  }
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 4 * n_tensors));
  if (most) {
    dim3 grid(vkp_grid_for(ctx, most, NB, 16), n_tensors);
    adam_apply_many_kernel<<<grid, NB, 0, ctx->stream>>>(a);
    VKP_TRY(vkp_after_launch(ctx, "nn_adam_apply_many"));
  }
  return vkp_finish_op(ctx, job);
}

extern "C" int vkp_fill_many_u32(vkp_ctx* ctx, int n_tensors, void* const* ptr, const size_t* count, uint32_t bits,
                                 vkp_job** job) {
  VKP_CHECK(ctx && ptr && count, "vkp_fill_many_u32: null argument");
  VKP_CHECK(n_tensors >= 1 && n_tensors <= VKP_NN_MAX_TENSORS, "vkp_fill_many_u32: 1..%d tensors", VKP_NN_MAX_TENSORS);
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g_(ctx->mu);
  FillMany a;
  size_t most = 0;
  for (int t = 0; t < n_tensors; t++) {
    a.p[t] = static_cast<uint32_t*>(ptr[t]);
    a.n[t] = count[t];
    if (count[t] > most) most = count[t];
  }
  VKP_TRY(vkp_prepare_buffers(ctx, ptr, n_tensors));
  if (most) {
    dim3 grid(vkp_grid_for(ctx, most, NB * 4, 8), n_tensors);
    fill_many_kernel<<<grid, NB, 0, ctx->stream>>>(a, bits);
    VKP_TRY(vkp_after_launch(ctx, "fill_many"));
  }
  return vkp_finish_op(ctx, job);
}

extern "C" int vkp_nn_activation_backward(vkp_ctx* ctx, int kind, const float* y, const float* dy, float* dx,
                                          size_t n, vkp_job** job) {
  VKP_RANGE(__func__);
  NN_PROLOGUE((void*)y, (void*)dy, dx);
  VKP_CHECK(kind == 0 || kind == 1, "vkp_nn_activation_backward: kind must be 0 (relu) or 1 (y(1-y))");
  if (n) {
    const unsigned grid = vkp_grid_for(ctx, n, NB, 16);
    if (kind == 0) relu_bwd_kernel<<<grid, NB, 0, ctx->stream>>>(y, dy, dx, n);
    else ymul_bwd_kernel<<<grid, NB, 0, ctx->stream>>>(y, dy, dx, n);
    VKP_TRY(vkp_after_launch(ctx, "nn_activation_backward"));
  }
  return vkp_finish_op(ctx, job);
}

extern "C" int vkp_nn_softmax_forward(vkp_ctx* ctx, const float* x, float* y, uint32_t rows, uint32_t cols,
                                      vkp_job** job) {
  VKP_RANGE(__func__);
  NN_PROLOGUE((void*)x, y);
  if (rows && cols) {
    const unsigned grid = vkp_grid_for(ctx, rows, NB / 32, 16);
    softmax_fwd_kernel<<<grid, NB, 0, ctx->stream>>>(vkpt::host_coef(), x, y, rows, cols);
    VKP_TRY(vkp_after_launch(ctx, "nn_softmax_forward"));
  }
  return vkp_finish_op(ctx, job);
}

extern "C" int vkp_nn_softmax_ce_train(vkp_ctx* ctx, const float* z, const float* t, float* p, float* L, float* dz,
                                       uint32_t rows, uint32_t cols, float scale, int has_scale, vkp_job** job) {
  VKP_RANGE(__func__);
  NN_PROLOGUE((void*)z, (void*)t, p, L, dz);
  if (rows && cols) {
    const unsigned grid = vkp_grid_for(ctx, rows, NB / 32, 16);
    softmax_ce_train_kernel<<<grid, NB, 0, ctx->stream>>>(vkpt::host_coef(), z, t, p, L, dz, rows, cols, scale, has_scale);
    VKP_TRY(vkp_after_launch(ctx, "nn_softmax_ce_train"));
  }
  return vkp_finish_op(ctx, job);
}
