/* vulkpy_b200.h -- C ABI of the B200 backend that replaces vulkpy's Vulkan layer.
 *
 * Drop-in boundary: this header replaces the pybind11 module `vulkpy._vkarray`
 * (reference vulkpy/_vkarray.cc:756-898).  Every entry point names the reference
 * interface it stands in for.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; the message of the
 *     last failure on the calling thread is returned by vkp_last_error()
 *     (reference: C++ exceptions -> Python RuntimeError, _vkarray.cc:32,286,446-456,508,752).
 *   - all work is enqueued on ONE in-order CUDA stream per context, so read-after-write,
 *     write-after-read and write-after-write hazards between submitted ops are ordered by
 *     construction (the reference host-blocks on every dependency: _vkarray.cc:430-432).
 *   - sizes that cross the reference boundary are uint32 (_vkarray.cc:132-203); the
 *     parameter structs below keep that layout bit-for-bit.  Kernels index with 64 bits.
 *   - buffers are plain device memory (cudaMalloc, pooled); the first request for a host view
 *     (vkp_host_view: NumPy view of the reference, _vkarray.cc:61-72,805-814) moves a buffer into a
 *     managed block whose pointer is valid on the host and on the device.
 */
#ifndef VULKPY_B200_H
#define VULKPY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VKP_ABI_VERSION 1

#if defined(__GNUC__)
#define VKP_API __attribute__((visibility("default")))
#else
#define VKP_API
#endif

typedef struct vkp_ctx vkp_ctx;     /* replaces class GPU        (_vkarray.cc:460-574) */
typedef struct vkp_job vkp_job;     /* replaces class Job        (_vkarray.cc:392-457) */
typedef struct vkp_rng vkp_rng;     /* replaces PRNG::Xoshiro128pp (_vkarray.cc:577-719) */
typedef struct vkp_timer vkp_timer; /* CUDA event on the context stream (measurement only) */

/* ---- error / introspection ------------------------------------------------------- */
VKP_API int         vkp_abi_version(void);
VKP_API const char* vkp_last_error(void);
VKP_API int         vkp_device_count(int* count);

/* ---- context: createGPU(n, priority) (_vkarray.cc:759-763), GPU::wait (:560-562) -- */
VKP_API int vkp_ctx_create(int device, float priority, vkp_ctx** out);
VKP_API int vkp_ctx_destroy(vkp_ctx* ctx);
VKP_API int vkp_ctx_sync(vkp_ctx* ctx);                          /* GPU.wait()                  */
VKP_API int vkp_ctx_device(vkp_ctx* ctx, int* device);
VKP_API int vkp_ctx_sm_count(vkp_ctx* ctx, int* sms);
VKP_API int vkp_ctx_set_debug_sync(vkp_ctx* ctx, int enable);    /* util.enable_debug analogue  */
VKP_API int vkp_ctx_launch_count(vkp_ctx* ctx, uint64_t* kernels);/* kernels launched so far    */
VKP_API int vkp_ctx_mem_info(vkp_ctx* ctx, size_t* pooled_bytes, size_t* live_bytes);
VKP_API int vkp_ctx_trim(vkp_ctx* ctx);                          /* give cached blocks back     */

/* ---- buffers: GPU::createBuffer<T>/toBuffer<T> (_vkarray.cc:512-525), Buffer<T> (:38-130)
 * Every device pointer an entry point of this library takes is the BASE pointer vkp_alloc
 * returned (the reference binds whole buffers too: BufferInfo, _vkarray.cc:88-96); the pool
 * tracks in-flight work per block by that pointer. */
VKP_API int vkp_alloc(vkp_ctx* ctx, size_t bytes, void** ptr);
VKP_API int vkp_free(vkp_ctx* ctx, void* ptr);
/* stream-ordered host->buffer / buffer->host copies (Buffer::set, _vkarray.cc:98-108) */
VKP_API int vkp_upload(vkp_ctx* ctx, void* dst, const void* src_host, size_t bytes);
VKP_API int vkp_download(vkp_ctx* ctx, void* dst_host, const void* src, size_t bytes);
/* Copy-engine transfers on side streams, for callers that pipeline steps: they overlap the
 * compute stream and each other (PCIe is full duplex).  Host memory must be page-locked
 * (vkp_host_alloc).  Ordering is per buffer: the upload waits for enqueued work that still uses
 * `dst`, operations bound to `dst` wait for the upload; the download sees everything enqueued
 * before it, later operations bound to `src` wait for it.  `*job` completes with the transfer
 * (the source of an upload may be rewritten / the destination of a download read after that).
 * vkp_alloc_for_upload prefers a cached block no enqueued work can still touch and, when every cached
 * block is still in the compute stream's future, takes a fresh one (while the pool is below half of the
 * device memory) so that the upload never waits for compute.  The reference
 * has no counterpart: Buffer::set (_vkarray.cc:98-108) is a blocking memcpy into mapped memory. */
VKP_API int vkp_alloc_for_upload(vkp_ctx* ctx, size_t bytes, void** ptr);
VKP_API int vkp_upload_async(vkp_ctx* ctx, void* dst, const void* src_pinned, size_t bytes, vkp_job** job);
VKP_API int vkp_download_async(vkp_ctx* ctx, void* dst_pinned, const void* src, size_t bytes, vkp_job** job);
/* Buffers are born in plain device memory (cudaMalloc).  The reference maps every buffer
 * host-visible and NumPy views it zero-copy (_vkarray.cc:61-72, :797-833); here the first request
 * for such a view moves the contents into a managed block whose pointer is valid on host and
 * device: `*host_visible` replaces `ptr` for every later call (the old pointer is released).
 * A buffer that is already host-visible returns itself. */
VKP_API int vkp_host_view(vkp_ctx* ctx, void* ptr, void** host_visible);
/* Call before the host touches a host-visible `ptr` through its NumPy view.  Makes sure no earlier
 * user of a recycled block is still in flight and (prefetch!=0) migrates the pages to
 * host memory in one bulk transfer instead of page faults. */
VKP_API int vkp_host_acquire(vkp_ctx* ctx, void* ptr, size_t bytes, int prefetch);
/* pinned host staging memory for callers that want full-rate PCIe copies */
VKP_API int vkp_host_alloc(size_t bytes, void** ptr);
VKP_API int vkp_host_free(void* ptr);

/* ---- parameter blocks: namespace OpParams (_vkarray.cc:132-203), same field order -- */
typedef struct { uint32_t size; }                                   vkp_vector_params;        /* Vector            */
typedef struct { uint32_t size[2]; }                                vkp_multivector2_params;  /* MultiVector<2>    */
typedef struct { uint32_t shift, size; }                            vkp_shiftvector_params;   /* ShiftVector       */
typedef struct { uint32_t size, low, high; }                        vkp_vectorrange_params;   /* VectorRange       */
typedef struct { uint32_t size; float scalar; }                     vkp_vectorscalar_params;  /* VectorScalar<f32> */
typedef struct { uint32_t size; float scalar[2]; }                  vkp_vectorscalar2_params; /* VectorMultiScalar<f32,2> */
typedef struct { uint32_t rowA, contractSize, columnB; }            vkp_matmul_params;        /* MatMul            */
typedef struct { uint32_t batch_size, input_size, output_size; }    vkp_batchaffine_params;   /* BatchAffine       */
typedef struct { uint32_t prev_prod, axis_size, post_prod; }        vkp_axisreduction_params; /* AxisReduction     */
typedef struct { uint32_t size[2]; uint32_t ndim; }                 vkp_broadcast_params;     /* Broadcast         */
typedef struct { uint32_t size[3]; uint32_t ndim; }                 vkp_multi3broadcast_params;/* MultiBroadcast<3>*/
typedef struct { uint32_t prev_prod, post_prod, axis_size, index_size; } vkp_axisgather_params;/* AxisGather       */

/* ---- ops: GPU::submit(spv, x,y,z, infos, DataShape, Params, wait) (_vkarray.cc:527-548)
 * The kernel is named by the reference shader's base name ("add", "iadd_scalar",
 * "sum_axis_rebroadcast", "sum_v1.3", ... the 121 files of vulkpy/shader/ and setup.py:11-48);
 * vkp_op_id() turns the name into the integer id vkp_submit() takes (-1 if unknown).
 * `bufs` are the bindings in the shader's binding order; `params` is the matching struct
 * above.  Shape bindings of the broadcast family (add_broadcast.comp binding 3,
 * iadd_broadcast.comp binding 2, broadcast.comp bindings 2 and 3) are read on the HOST
 * at submit time (they are tiny and produced by the host), every other binding is a
 * buffer from vkp_alloc.  The workgroup/DataShape arguments of the reference are not
 * needed: grids are derived from `params`.  *job receives a waitable handle (may be NULL). */
VKP_API int vkp_op_id(const char* name);
VKP_API const char* vkp_op_name(int op);
VKP_API int vkp_op_count(void);
VKP_API int vkp_submit(vkp_ctx* ctx, int op, void* const* bufs, int nbuf,
               const void* params, size_t params_bytes, vkp_job** job);

/* extra device-side utilities with no shader counterpart in the reference */
VKP_API int vkp_fill_u32(vkp_ctx* ctx, void* dst, size_t count, uint32_t bits, vkp_job** job); /* `a[:] = v` host fills: vkarray.py:1540-1542, nn/parameters.py:81-86 */

/* Fused chain of same-shape element-wise shaders (lazy element-wise fusion).  What the reference
 * issues as n_steps dependent jobs over throw-away intermediates -- e.g. Sigmoid.forward, nn/layers.py:239-243:
 * `y = 0.0 - x; y.exp(inplace); y += 1.0; y = 1.0 / y` -- is ONE launch that reads each input once and writes
 * `out` once.  A running value starts as in[0]; step k applies the operation of ONE reference shader to it with
 * exactly that shader's rounding (results are bit-identical to the op-by-op sequence).  srcs[k] names the second
 * operand: 0 = scalars[k], 1..3 = in[1..3] (same element count), 4 = the copy saved by an earlier VKP_CHAIN_SAVE. */
#define VKP_CHAIN_ADD   0   /* acc = acc + b   (add.comp / add_scalar.comp) */
#define VKP_CHAIN_SUB   1
#define VKP_CHAIN_MUL   2
#define VKP_CHAIN_DIV   3
#define VKP_CHAIN_MAX   4
#define VKP_CHAIN_MIN   5
#define VKP_CHAIN_POW   6
#define VKP_CHAIN_RSUB  7   /* acc = b - acc   (rsub_scalar.comp) */
#define VKP_CHAIN_RDIV  8
#define VKP_CHAIN_RPOW  9
#define VKP_CHAIN_UNARY 11  /* + index of the unary shader: abs sign sin cos tan asin acos atan sinh cosh tanh asinh
                             *   acosh atanh exp log exp2 log2 sqrt invsqrt (0..19) */
#define VKP_CHAIN_SAVE  31  /* tmp = acc */
#define VKP_CHAIN_MAX_STEPS 16
VKP_API int vkp_ew_chain(vkp_ctx* ctx, int n_in, const float* const* in, float* out, size_t count, int n_steps,
                         const int* ops, const int* srcs, const float* scalars, vkp_job** job);

/* General fp32 GEMM behind "matmul"/"batch_affine": C[M,N] = op(A)·op(B) (+ bias[N]).
 * transA=0: A is [M,K] row-major, 1: A is [K,M].  transB=0: B is [K,N], 1: B is [N,K].
 * Used for Dense.backward (nn/layers.py:104-141) where the reference materialises
 * a B x out x in temporary instead.  flags: VKP_GEMM_* */
#define VKP_GEMM_AUTO      0   /* tcgen05 3xTF32 when the shape allows, SIMT otherwise */
#define VKP_GEMM_FORCE_SIMT 1
#define VKP_GEMM_FORCE_TC   2
#define VKP_GEMM_ACCUMULATE 4  /* C += instead of C = */
#define VKP_GEMM_RELU       8  /* C = max(C, 0) after bias: Dense.forward followed by ReLU.forward (x.max(0.0),
                                * nn/layers.py:186) in the GEMM epilogue, same float32 operation */
VKP_API int vkp_gemm(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K,
             const float* A, const float* B, float* C, const float* bias, int flags,
             vkp_job** job);

/* vkp_gemm with an activation gradient folded into the epilogue: relu_mask ([M, N] like C, or NULL) is the OUTPUT y of
 * the ReLU in front of this layer and every C element becomes max(sign(y), 0) * C -- ReLU.backward (nn/layers.py:207-210)
 * applied to the dx = dy W contraction of Dense.backward, bit-identical to the separate job. */
VKP_API int vkp_gemm_fused(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K,
                           const float* A, const float* B, float* C, const float* bias, const float* relu_mask,
                           int flags, vkp_job** job);

/* ---- fused vulkpy.nn steps: same float32 operations, order and roundings as the
 * reference's op-by-op compositions, one kernel each ------------------------------------------- */
/* AdamState.grad2diff (nn/optimizers.py:235-253): updates m, v in place, writes the parameter
 * update to diff.  All scalars are the float32 values the reference would pass op by op. */
VKP_API int vkp_nn_adam(vkp_ctx* ctx, const float* grad, float* m, float* v, float* diff, size_t n,
                        float beta1, float one_minus_beta1, float beta2, float one_minus_beta2,
                        float one_minus_beta1t, float one_minus_beta2t, float eps, float neg_lr,
                        vkp_job** job);
/* The optimizer step of up to 16 parameters in one launch: per element vkp_nn_adam followed by
 * Parameter.update's `value += diff` (nn/parameters.py:88-95).  scalars = [n_tensors][8]:
 * beta1, 1-beta1, beta2, 1-beta2, 1-beta1^t, 1-beta2^t, eps, -lr as float32. */
VKP_API int vkp_nn_adam_apply_many(vkp_ctx* ctx, int n_tensors, const float* const* grad, float* const* m,
                                   float* const* v, float* const* value, const size_t* count,
                                   const float* scalars, vkp_job** job);
/* `a[:] = scalar` on up to 16 arrays in one launch (Parameter.zero_grad of a whole model,
 * nn/parameters.py:81-86, nn/models.py:46-53) */
VKP_API int vkp_fill_many_u32(vkp_ctx* ctx, int n_tensors, void* const* ptr, const size_t* count, uint32_t bits,
                              vkp_job** job);
/* kind 0: ReLU.backward dx = max(sign(y),0)*dy (nn/layers.py:207-210);
 * kind 1: Sigmoid/Softmax.backward dx = ((1-y)*y)*dy (nn/layers.py:262-267,320-323) */
VKP_API int vkp_nn_activation_backward(vkp_ctx* ctx, int kind, const float* y, const float* dy, float* dx,
                                       size_t n, vkp_job** job);
/* Softmax.forward over axis 1 of [rows, cols] (nn/layers.py:297-300) */
VKP_API int vkp_nn_softmax_forward(vkp_ctx* ctx, const float* x, float* y, uint32_t rows, uint32_t cols,
                                   vkp_job** job);
/* Tail of a classifier's training step in one launch: p = softmax(z) (nn/layers.py:297-300),
 * L = -t log(p + 1e-8) (nn_cross_entropy.comp:25), dz = ((1-p) p) * ((-t / (p + 1e-8)) [* scale])
 * (nn_cross_entropy_backward.comp:25, nn/losses.py:58-68, nn/layers.py:320-323); the same float32 operations
 * in the same order as the five launches it replaces.  z, t, p, L, dz: [rows, cols]. */
VKP_API int vkp_nn_softmax_ce_train(vkp_ctx* ctx, const float* z, const float* t, float* p, float* L, float* dz,
                                    uint32_t rows, uint32_t cols, float scale, int has_scale, vkp_job** job);

/* ---- argmax / argmin / permutation.  The reference lists them as missing
 * (README.md:73 "argmax, argmin", :77 "shuffle") and its training example does them on the host
 * (example/02-nn.py:82 `rng.shuffle(idx)`, :96 `np.argmax(pred_y, axis=1)`); semantics are NumPy's:
 * first occurrence wins, NaN counts as the extreme.  `in` is [prev, axis, post] as in the axis
 * reductions (vkarray.py:1398-1432), `out` [prev, post] uint32.  op: 0 = argmax, 1 = argmin.
 * vkp_argsort_u32: out = indices 0..n-1 stably sorted by keys (np.argsort(keys, kind="stable"));
 * with keys from vkp_rng_uint32 that is a random permutation. */
VKP_API int vkp_argreduce(vkp_ctx* ctx, int op, const float* in, uint32_t* out, uint32_t prev, uint32_t axis,
                          uint32_t post, vkp_job** job);
VKP_API int vkp_argsort_u32(vkp_ctx* ctx, const uint32_t* keys, uint32_t* out, uint32_t n, vkp_job** job);

/* ---- jobs: Job::wait (_vkarray.cc:446-456, :876-879) ------------------------------ */
VKP_API int vkp_job_wait(vkp_job* job, uint64_t timeout_ns);   /* timeout_ns==UINT64_MAX: forever; 2 = timeout */
VKP_API int vkp_job_done(vkp_job* job, int* done);
VKP_API int vkp_job_release(vkp_job* job);

/* ---- PRNG: Xoshiro128pp(gpu, spv_u32, spv_f32, size[, seed]) (_vkarray.cc:643-679),
 *      random_uint32 / random_float (:681-717).  Bit-exact stream and state layout. ---- */
VKP_API int vkp_rng_create(vkp_ctx* ctx, uint32_t size, uint64_t seed, int has_seed, vkp_rng** out);
VKP_API int vkp_rng_destroy(vkp_rng* rng);
VKP_API int vkp_rng_uint32(vkp_rng* rng, uint32_t* out, uint32_t n, vkp_job** job);
VKP_API int vkp_rng_float(vkp_rng* rng, float* out, uint32_t n, vkp_job** job);
/* fused uniform -> Box-Muller (random.py:60-124 + prng_box_muller.comp / prng_ibox_muller.comp):
 * consumes n (even) or n+1 (odd) uniforms exactly like the reference */
VKP_API int vkp_rng_normal(vkp_rng* rng, float* out, uint32_t n, float mean, float stddev, vkp_job** job);
VKP_API int vkp_rng_state(vkp_rng* rng, uint32_t* host_out /* 4*size words */);
/* discard n draws exactly as vkp_rng_uint32(n) would consume them (GF(2) jump-ahead): lets every rank
 * of a sharded generator start at its own chunk of the single-GPU stream.  Whole multiples of the lane
 * count are recorded and applied lazily: consecutive skips merge and the next draw folds them into its
 * own jump-ahead (vkp_rng_state applies a pending skip first) */
VKP_API int vkp_rng_advance(vkp_rng* rng, uint64_t n);

/* ---- timing on the context stream (bench only) ------------------------------------ */
VKP_API int vkp_timer_create(vkp_ctx* ctx, vkp_timer** out);
VKP_API int vkp_timer_record(vkp_timer* t);
VKP_API int vkp_timer_elapsed_ms(vkp_timer* start, vkp_timer* stop, float* ms); /* syncs on stop */
VKP_API int vkp_timer_destroy(vkp_timer* t);

/* ---- multi-GPU (additive; the reference has no multi-device surface) -------------- */
#define VKP_COMM_ID_BYTES 128
VKP_API int vkp_comm_unique_id(void* id_out /* VKP_COMM_ID_BYTES */);
VKP_API int vkp_comm_init(vkp_ctx* ctx, int nranks, int rank, const void* id);
VKP_API int vkp_comm_destroy(vkp_ctx* ctx);
/* op: 0 sum, 1 prod, 2 max, 3 min (the four reductions of vkarray.py:1278-1396) */
VKP_API int vkp_comm_allreduce(vkp_ctx* ctx, const float* send, float* recv, size_t count, int op, vkp_job** job);
/* n <= 16 tensors all-reduced in place as one grouped NCCL launch, then x *= scale on all of them in
 * one kernel (the data-parallel gradient bucket: sum, then 1/world so reduce="mean" losses of
 * nn/losses.py:41-46 keep their meaning) */
VKP_API int vkp_comm_allreduce_multi(vkp_ctx* ctx, float* const* bufs, const size_t* counts, int n, int op,
                                     float scale, vkp_job** job);
VKP_API int vkp_comm_allgather(vkp_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank, vkp_job** job);
/* Small all-reduces (reduction partials, the gradient bucket: anything that fits an 8 MiB mailbox slot)
 * and barriers do not call NCCL: every rank owns a mailbox block mapped into all peers with CUDA IPC and
 * ONE kernel copies the operands in, raises a flag in every peer's mailbox, waits for the peers' flags and
 * folds the w slots in rank order with 16-byte loads over NVLink (bit-identical results on all ranks).
 * VKP_COMM_PEER=0, a failed IPC mapping on any rank, or a larger operand select NCCL on ALL ranks.
 * vkp_comm_reduce_allreduce fuses the local reduction of a shard with that exchange: in [prev, axis, post]
 * -> local partials written straight into the mailbox slot -> out [prev, post] reduced over the ranks
 * (full reduction: prev = 1, axis = n, post = 1).  The reference reduces one array on one device
 * (vkarray.py:1194-1276); these are its sharded forms. */
VKP_API int vkp_comm_reduce_allreduce(vkp_ctx* ctx, int op, const float* in, float* out, uint32_t prev,
                                      uint32_t axis, uint32_t post, vkp_job** job);
VKP_API int vkp_comm_barrier(vkp_ctx* ctx, vkp_job** job);
/* mode 0: NCCL for every later collective, 1: peer mailbox where it applies (default), -1: query.
 * *active = 1 when the mailbox is mapped and selected.  Collective (same point on every rank). */
VKP_API int vkp_comm_peer_mode(vkp_ctx* ctx, int mode, int* active);
/* Row-sharded matmul (vkarray.py:585-605 on a sharded pair): C[M,N] = A[M,K] @ B[K,N] with A, C the
 * local row blocks and B_shard the local [K/nranks, N] row block of B.  ONE tcgen05 GEMM kernel: it
 * starts on the local K range while its spare warps pull the peers' shards over NVLink (CUDA IPC
 * peer memory) and enters each further range as its flag is raised; collective, same shapes on
 * every rank, at most 16 ranks. */
VKP_API int vkp_comm_matmul_allgather(vkp_ctx* ctx, uint32_t M, uint32_t N, uint32_t K, const float* A,
                                      const float* B_shard, float* C, vkp_job** job);

#ifdef __cplusplus
}
#endif
#endif /* VULKPY_B200_H */
