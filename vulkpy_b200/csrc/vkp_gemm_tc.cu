// vkp_gemm_tc.cu -- float32 GEMM on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with a
// 3xTF32 split, accumulators in TMEM, operands staged by TMA.
//
//   C[M,N] (+)= A[M,K] * Bt[N,K]^T (+ bias[N])        both operands K-major in shared memory
//
// Every fp32 operand x is split as x = hi + lo with hi = x with the low 13 mantissa bits cleared
// (exactly a TF32 number) and lo = x - hi (exact in fp32, then cut to TF32), and
//   a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi      (the dropped a_lo*b_lo term is ~2^-22 |a b|)
// is accumulated in fp32 in TMEM, i.e. 3 tensor-core MMAs per k-step.  Non-finite inputs come out
// as NaN (inf * b_lo with b_lo = 0), where a plain fp32 product would keep the infinity.
//
// CTA = 12 warps, one CTA per SM, persistent over output tiles (BM=128 x BN):
//   warp 0      TMA producer: cp.async.bulk.tensor 128B-swizzled fp32 tiles of A and Bt (the "hi"
//               tiles are the raw fp32 data) into a STAGES-deep ring, completion on mbarriers
//   warps 8-11  converters: read each landed tile, rewrite hi (low bits cleared) and write the lo
//               tile at the same (swizzled) offsets, fence.proxy.async, signal the MMA warp
//   warp 1      MMA issuer: one elected lane issues 4 k-steps x 3 tcgen05.mma (M=128, N=BN, K=8)
//               per stage; tcgen05.commit releases the stage / publishes the accumulator
//   warp 2      TMEM allocator (2 x BN fp32 columns: double-buffered accumulators)
//   warps 4-7   epilogue: tcgen05.ld 32 lanes x 32 columns per warp, + bias / + C, 128-byte
//               row segments stored straight to global memory
// Operands that are not K-major (A given as [K,M], B given as [K,N]: the B of every `a @ b`, both
// operands of Dense.backward's dW = dy^T x) are read as they lie: their tiles are MN-major in
// shared memory (tcgen05 takes MN-major TF32 operands), fetched by one 3-D TMA per tile that
// views the [K, MN] matrix as {32 floats of MN, K, MN/32} so that every 32-wide chunk lands as
// its own slab of BK rows x 128 B, swizzled in 32-byte atoms (the only MN-major form tcgen05 takes
// for 32-bit data: SWIZZLE_128B_BASE32B, LBO = BK*128 B between chunks, SBO = 512 B between 4-row groups).  Only leading dimensions that are not a multiple of 32 still go
// through the tiled transpose kernel (VKP_TC_MN=0 forces it: A/B measurements).
//
// PRESPLIT variant (large problems): the converter warps saturate the shared-memory pipe (they read
// and write every tile once more: profiles/r01_ncu_gemm_tc_notes.md), so for problems where an
// extra pass over the operands is cheap next to the contraction the lo parts are produced once in
// HBM by a pre-pass (fused with the transpose where one is needed) and TMA brings four tiles per
// stage (A, A_lo, Bt, Bt_lo); the converter warps retire immediately.  Measured at 8192^3
// (profiles/r01_gemm_variants.txt): 5.07 ms -> 4.27 ms including the pre-pass.
#include "vkp_common.cuh"
#include "vkp_math.cuh"

#include <cuda.h>
#include <cstdlib>

namespace {

constexpr int BM = 128;

__device__ __forceinline__ float4 post4(float4 o, const vkp_gemm_post& p, size_t idx) {
  if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
  if (p.mask) {
    const float4 y = *reinterpret_cast<const float4*>(p.mask + idx);
    o.x = fmaxf(vkpm::sign_f(y.x), 0.f) * o.x; o.y = fmaxf(vkpm::sign_f(y.y), 0.f) * o.y;
    o.z = fmaxf(vkpm::sign_f(y.z), 0.f) * o.z; o.w = fmaxf(vkpm::sign_f(y.w), 0.f) * o.w;
  }
  return o;
}
constexpr int UMMA_K = 8;         // kind::tf32
constexpr int NUM_THREADS = 384;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// spin with a watchdog: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; spin++) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spin > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Wait until *flag == epoch (written by the stream that brings the K range in, after the data).
// Bounded: ~2 s of polling, then trap -- a lost flag must fail the call, not hang the GPU.
__device__ __forceinline__ void chunk_wait(const uint32_t* flag, uint32_t epoch) {
  uint32_t v;
  for (uint32_t spin = 0;; spin++) {
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v == epoch) break;
    __nanosleep(200);
    if (spin > (1u << 23)) __trap();
  }
  asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy acquire -> TMA (async proxy) reads
}

__device__ __forceinline__ uint4 ld_peer16(const float* p) {   // peer memory over NVLink: never cached
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t tf32_lo_bits(uint32_t xb) {
  const float h = __uint_as_float(xb & 0xffffe000u);            // what the tensor core reads of x
  return (__float_as_uint(__uint_as_float(xb) - h) + 0x1000u) & 0xffffe000u;
}

// The communication half of the row-sharded matmul, run by the four warps that have nothing to
// convert in the pre-split variant (tid 0..127 of every CTA): stream the peers' K ranges in ring
// order (first+1, first+2, ...: at any time every owner serves one reader) with 16-byte loads from
// the mapped peer pointers, 8 in flight per thread, store hi and lo locally, then count the CTA
// in; the last one publishes the range to every producer warp (release / acquire on the flag).
__device__ __forceinline__ void pull_ranges(const vkp_tc_pull& pl, const vkp_tc_chunks& ch, int tid) {
  constexpr int U = 16, PT = 128;   // 16 loads in flight per thread: with 8 the pulls sustained ~320 GB/s and the fixed-size product on 8 GPUs was pull-bound (0.77 ms against 0.51 ms for the same GEMM with nothing crossing NVLink, profiles/r02_mm_fused_probe_n8.txt)
  const uint32_t c4 = pl.kc / 4;
  const size_t n4 = (size_t)pl.rows * c4;
  const size_t stride = (size_t)gridDim.x * PT;
  for (uint32_t j = 1; j < ch.n_chunks; j++) {
    uint32_t s = ch.first + j;
    if (s >= ch.n_chunks) s -= ch.n_chunks;
    const float* src = pl.src[s] + (size_t)s * pl.kc;
    float* hi = pl.hi + (size_t)s * pl.kc;
    float* lo = pl.lo + (size_t)s * pl.kc;
    for (size_t i0 = blockIdx.x * (size_t)PT + tid; i0 < n4; i0 += stride * U) {
      uint4 v[U];
      size_t off[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t i = i0 + u * stride;
        const size_t r = i / c4;
        off[u] = r * pl.ld + (i - r * c4) * 4;
        if (i < n4) v[u] = ld_peer16(src + off[u]);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        if (i0 + u * stride < n4) {
          *reinterpret_cast<uint4*>(hi + off[u]) = v[u];
          *reinterpret_cast<uint4*>(lo + off[u]) =
              make_uint4(tf32_lo_bits(v[u].x), tf32_lo_bits(v[u].y), tf32_lo_bits(v[u].z), tf32_lo_bits(v[u].w));
        }
      }
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");     // the four pulling warps of this CTA
    if (tid == 0) {
      __threadfence();
      if (atomicAdd(pl.counters + s, 1u) == gridDim.x - 1) {
        pl.counters[s] = 0;                            // ready for the next call (stream-ordered)
        __threadfence();
        asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(ch.flags + s), "r"(ch.epoch) : "memory");
      }
    }
  }
}

// K-major, swizzled rows of BK fp32 (128 B -> SWIZZLE_128B, 64 B -> SWIZZLE_64B), 8-row groups
// 8*row bytes apart (SBO), version 1 (sm_100)
template <int BK>
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  static_assert(BK == 32 || BK == 16, "BK is one 128-byte or one 64-byte swizzle row");
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);        // start address, 16-byte units
  d |= (uint64_t)1 << 16;                            // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)((8 * BK * 4) >> 4) << 32;          // stride byte offset
  d |= (uint64_t)1 << 46;                            // descriptor version
  d |= (uint64_t)(BK == 32 ? 2 : 4) << 61;           // SWIZZLE_128B / SWIZZLE_64B
  return d;
}

// MN-major operand.  For 32-bit (TF32) data tcgen05 takes only the 128-byte swizzle with 32-byte atoms
// (layout type 1, SWIZZLE_128B_BASE32B; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): 32-float chunks of
// MN, each a slab of BK rows x 128 B in which the four 32-byte pieces of row r sit at piece ^ (r & 3).
// Canonical form ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units: groups of 4 k-rows are 512 B
// apart (SBO), consecutive chunks BK*128 B apart (LBO); one k-step of kind::tf32 (8 k) = 1024 B.
template <int BK>
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((BK * 128) >> 4) << 16;            // leading byte offset: next 32-wide chunk of MN
  d |= (uint64_t)(512 >> 4) << 32;                   // stride byte offset: next group of 4 k
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;                            // SWIZZLE_128B_BASE32B
  return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int BN, int BK>
struct Cfg {
  static constexpr int A_TILE_BYTES = BM * BK * 4;
  static constexpr int B_TILE_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;   // hi + lo of A and B
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;                 // BN=256: 2 (BK=32) / 4 (BK=16)
  static constexpr int TMEM_COLS = 2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512));   // two accumulators; allocations are powers of two
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

enum : int { MODE_CONVERT = 0, MODE_CONVERT_REWRITE_HI = 1, MODE_PRESPLIT = 2 };

// Split a landed fp32 tile into TF32 hi / lo parts at the same (swizzled) offsets.
//   REWRITE_HI = true : hi = RN_tf32(x) written back in place, lo = RN_tf32(x - hi)
//                       (|x - hi - lo| <= 2^-22 |x|, independent of how the tensor core treats the
//                       low 13 mantissa bits of an fp32 operand)
//   REWRITE_HI = false: the hi tile stays the raw fp32 data -- the tensor core ignores the low 13
//                       bits (truncation, verified by tests/test_gpu_gemm.py) -- and
//                       lo = RN_tf32(x - trunc_tf32(x)); one shared-memory store less per 16 bytes
//                       (the kernel is shared-memory-pipe bound: profiles/r01_ncu_gemm_tc.csv).
template <bool REWRITE_HI>
__device__ __forceinline__ void split_tile(uint8_t* hi, uint8_t* lo, int bytes, int tid, int nthreads) {
  for (int off = tid * 16; off < bytes; off += nthreads * 16) {
    uint4 x = *reinterpret_cast<uint4*>(hi + off);
    uint4 h, l;
    if (REWRITE_HI) {
      h.x = (x.x + 0x1000u) & 0xffffe000u; h.y = (x.y + 0x1000u) & 0xffffe000u;
      h.z = (x.z + 0x1000u) & 0xffffe000u; h.w = (x.w + 0x1000u) & 0xffffe000u;
    } else {
      h.x = x.x & 0xffffe000u; h.y = x.y & 0xffffe000u; h.z = x.z & 0xffffe000u; h.w = x.w & 0xffffe000u;
    }
    l.x = (__float_as_uint(__uint_as_float(x.x) - __uint_as_float(h.x)) + 0x1000u) & 0xffffe000u;
    l.y = (__float_as_uint(__uint_as_float(x.y) - __uint_as_float(h.y)) + 0x1000u) & 0xffffe000u;
    l.z = (__float_as_uint(__uint_as_float(x.z) - __uint_as_float(h.z)) + 0x1000u) & 0xffffe000u;
    l.w = (__float_as_uint(__uint_as_float(x.w) - __uint_as_float(h.w)) + 0x1000u) & 0xffffe000u;
    if (REWRITE_HI) *reinterpret_cast<uint4*>(hi + off) = h;
    *reinterpret_cast<uint4*>(lo + off) = l;
  }
}

template <int BN, int BK, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo,
               float* __restrict__ C, const float* __restrict__ bias, uint32_t M, uint32_t N, uint32_t K,
               int accumulate, uint32_t splits, uint32_t kb_per_split, vkp_tc_chunks ch,
               const __grid_constant__ vkp_tc_pull pl, int a_mn, int b_mn, vkp_gemm_post post) {
  // a_mn / b_mn: the operand lies as [K, MN] in memory (MN contiguous): 3-D tensor map, MN-major tiles
  // ch.n_chunks > 1 (row-sharded matmul, vkp_comm.cu): K is cut into n_chunks ranges that become
  // valid one after the other while this kernel runs -- the peers' shards of B, fetched over NVLink
  // by this kernel's own spare warps (pull_ranges); the producer walks them starting at ch.first
  // (the local shard) and, before the first load of another range, waits until
  // ch.flags[range] == ch.epoch.
  // splits > 1 (split-K for problems with fewer output tiles than SMs): work item = (tile, split),
  // C is then a [splits][M][N] partial buffer and bias / accumulate are applied by splitk_reduce.
  using cfg = Cfg<BN, BK>;
  constexpr int STAGES = cfg::STAGES;
  constexpr int A_TILE_BYTES = cfg::A_TILE_BYTES;
  constexpr bool PRESPLIT = MODE == MODE_PRESPLIT;
  constexpr bool REWRITE_HI = MODE == MODE_CONVERT_REWRITE_HI;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t* smem = smem_raw + ((1024 - (smem_u32(smem_raw) & 1023)) & 1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                  // [STAGES] TMA landed
  uint64_t* conv_bar = bars + STAGES;         // [STAGES] lo tiles written
  uint64_t* empty_bar = bars + 2 * STAGES;    // [STAGES] MMAs that read the stage retired
  uint64_t* tfull_bar = bars + 3 * STAGES;    // [2] accumulator complete
  uint64_t* tempty_bar = bars + 3 * STAGES + 2;  // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t m_blocks = (M + BM - 1) / BM, n_blocks = (N + BN - 1) / BN;
  const uint32_t num_tiles = m_blocks * n_blocks * splits;     // work items
  const uint32_t k_blocks = (K + BK - 1) / BK;
  auto k_range = [&](uint32_t work, uint32_t& kb0, uint32_t& kb1) {
    const uint32_t sp = work % splits;
    kb0 = sp * kb_per_split;
    kb1 = (kb0 + kb_per_split < k_blocks) ? kb0 + kb_per_split : k_blocks;
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (PRESPLIT) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAlo) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; s++) {
      mbar_init(smem_u32(&full_bar[s]), 1);
      mbar_init(smem_u32(&conv_bar[s]), 4);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    for (int a = 0; a < 2; a++) {
      mbar_init(smem_u32(&tfull_bar[a]), 1);
      mbar_init(smem_u32(&tempty_bar[a]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // GROUP_M consecutive m-blocks share their B tiles in L2
  constexpr uint32_t GROUP_M = 16;
  auto tile_coords = [&](uint32_t work, uint32_t& mb, uint32_t& nb) {
    const uint32_t tile = work / splits;
    const uint32_t per_group = GROUP_M * n_blocks;
    const uint32_t g = tile / per_group;
    const uint32_t first_m = g * GROUP_M;
    const uint32_t gsz = (m_blocks - first_m) < GROUP_M ? (m_blocks - first_m) : GROUP_M;
    const uint32_t r = tile - g * per_group;
    mb = first_m + r % gsz;
    nb = r / gsz;
  };

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        uint32_t mb, nb;
        tile_coords(tile, mb, nb);
        uint32_t kb0, kb1;
        k_range(tile, kb0, kb1);
        for (uint32_t it = kb0; it < kb1; it++) {
          uint32_t kb = it;
          if (ch.n_chunks > 1) {
            const uint32_t ci = it / ch.kb_per_chunk;
            uint32_t c = ch.first + ci;
            if (c >= ch.n_chunks) c -= ch.n_chunks;
            kb = c * ch.kb_per_chunk + (it - ci * ch.kb_per_chunk);
            if (ci != 0 && it == ci * ch.kb_per_chunk) chunk_wait(ch.flags + c, ch.epoch);
          }
          mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
          uint8_t* st = smem + stage * cfg::STAGE_BYTES;
          const uint32_t fb = smem_u32(&full_bar[stage]);
          mbar_arrive_expect_tx(fb, PRESPLIT ? cfg::STAGE_BYTES : A_TILE_BYTES + cfg::B_TILE_BYTES);
          // MN-major operand: mode 1 = one 3-D box {32, BK, chunks}; mode 2 = one 2-D box {32, BK} per
          // 32-wide chunk of the [K, MN] matrix, each landing as its own BK x 128 B slab
          auto load_a = [&](uint32_t dst, const CUtensorMap* m) {
            if (a_mn == 1) tma_load_3d(dst, m, fb, 0, (int)(kb * BK), (int)(mb * (BM / 32)));
            else if (a_mn == 2) {
              for (int c = 0; c < BM / 32; c++) tma_load_2d(dst + c * BK * 128, m, fb, (int)(mb * BM + c * 32), (int)(kb * BK));
            } else tma_load_2d(dst, m, fb, (int)(kb * BK), (int)(mb * BM));
          };
          auto load_b = [&](uint32_t dst, const CUtensorMap* m) {
            if (b_mn == 1) tma_load_3d(dst, m, fb, 0, (int)(kb * BK), (int)(nb * (BN / 32)));
            else if (b_mn == 2) {
              for (int c = 0; c < BN / 32; c++) tma_load_2d(dst + c * BK * 128, m, fb, (int)(nb * BN + c * 32), (int)(kb * BK));
            } else tma_load_2d(dst, m, fb, (int)(kb * BK), (int)(nb * BN));
          };
          load_a(smem_u32(st), &tmA);
          load_b(smem_u32(st + 2 * A_TILE_BYTES), &tmB);
          if (PRESPLIT) {
            load_a(smem_u32(st + A_TILE_BYTES), &tmAlo);
            load_b(smem_u32(st + 2 * A_TILE_BYTES + cfg::B_TILE_BYTES), &tmBlo);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=tf32, both K-major, N>>3 @17, M>>4 @24
      // (bit 15 / 16: A / B is MN-major)
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24) |
                             (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        uint32_t kb0, kb1;
        k_range(tile, kb0, kb1);
        for (uint32_t kb = kb0; kb < kb1; kb++) {
          mbar_wait(smem_u32(&full_bar[stage]), phase);
          if (!PRESPLIT) mbar_wait(smem_u32(&conv_bar[stage]), phase);
          tcgen05_fence_after();
          uint8_t* st = smem + stage * cfg::STAGE_BYTES;
          const uint32_t sa = smem_u32(st), sb = smem_u32(st + 2 * A_TILE_BYTES);
          const uint64_t a_hi = a_mn ? make_desc_mn<BK>(sa) : make_desc<BK>(sa);
          const uint64_t a_lo = a_mn ? make_desc_mn<BK>(sa + A_TILE_BYTES) : make_desc<BK>(sa + A_TILE_BYTES);
          const uint64_t b_hi = b_mn ? make_desc_mn<BK>(sb) : make_desc<BK>(sb);
          const uint64_t b_lo = b_mn ? make_desc_mn<BK>(sb + cfg::B_TILE_BYTES) : make_desc<BK>(sb + cfg::B_TILE_BYTES);
          // one k-step (8 k): 32 bytes along a K-major swizzle row, one 1024-byte row group of an MN-major slab
          const uint64_t step_a = a_mn ? (1024 >> 4) : ((UMMA_K * 4) >> 4);
          const uint64_t step_b = b_mn ? (1024 >> 4) : ((UMMA_K * 4) >> 4);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; k++) {
            const uint64_t adv_a = k * step_a, adv_b = k * step_b;
            umma_tf32(d_tmem, a_lo + adv_a, b_hi + adv_b, idesc, ((kb - kb0) | k) != 0);
            umma_tf32(d_tmem, a_hi + adv_a, b_lo + adv_b, idesc, 1);
            umma_tf32(d_tmem, a_hi + adv_a, b_hi + adv_b, idesc, 1);
          }
          tcgen05_commit(smem_u32(&empty_bar[stage]));
          if (kb == kb1 - 1) tcgen05_commit(smem_u32(&tfull_bar[acc]));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 8) {
    // ===================================== converters =======================================
    if (PRESPLIT) {                       // the lo tiles arrive by TMA: nothing to convert
      if (ch.n_chunks > 1 && pl.hi) pull_ranges(pl, ch, threadIdx.x - 256);
      goto teardown;
    }
    const int ctid = threadIdx.x - 256;   // 0..127
    uint32_t stage = 0, phase = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      uint32_t kb0, kb1;
      k_range(tile, kb0, kb1);
      for (uint32_t kb = kb0; kb < kb1; kb++) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        uint8_t* st = smem + stage * cfg::STAGE_BYTES;
        split_tile<REWRITE_HI>(st, st + A_TILE_BYTES, A_TILE_BYTES, ctid, 128);
        split_tile<REWRITE_HI>(st + 2 * A_TILE_BYTES, st + 2 * A_TILE_BYTES + cfg::B_TILE_BYTES, cfg::B_TILE_BYTES, ctid, 128);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> tensor-core reads
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&conv_bar[stage]));
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================================== epilogue =========================================
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    uint32_t acc = 0, acc_phase = 0;
    for (uint32_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      uint32_t mb, nb;
      tile_coords(tile, mb, nb);
      mbar_wait(smem_u32(&tfull_bar[acc]), acc_phase);
      tcgen05_fence_after();
      const uint32_t row = mb * BM + q * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c, v);
        const uint32_t col0 = nb * BN + c;
        if (row < M && col0 < N) {
          float* dst = C + (size_t)(tile % splits) * M * N + (size_t)row * N + col0;
          const int ncol = (N - col0) < 32u ? (int)(N - col0) : 32;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j < ncol) {   // N % 4 == 0 is required by the host wrapper
              float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3]));
              if (bias) {
                const float4 bv = *reinterpret_cast<const float4*>(bias + col0 + j);
                o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
              }
              if (accumulate) {
                const float4 cv = *reinterpret_cast<const float4*>(dst + j);
                o.x += cv.x; o.y += cv.y; o.z += cv.z; o.w += cv.w;
              }
              *reinterpret_cast<float4*>(dst + j) = post4(o, post, (size_t)row * N + col0 + j);
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&tempty_bar[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

teardown:
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)cfg::TMEM_COLS)
                 : "memory");
  }
}

// C = (accumulate ? C : 0) + bias + sum_s part[s]   (fixed order: deterministic)
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ part, float* __restrict__ C, const float* __restrict__ bias,
                     uint32_t M, uint32_t N, uint32_t splits, int accumulate, vkp_gemm_post post) {
  const size_t mn = (size_t)M * N;
  for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4; i < mn; i += (size_t)gridDim.x * blockDim.x * 4) {
    float4 acc = *reinterpret_cast<const float4*>(part + i);
    for (uint32_t s2 = 1; s2 < splits; s2++) {
      const float4 p = *reinterpret_cast<const float4*>(part + s2 * mn + i);
      acc.x += p.x; acc.y += p.y; acc.z += p.z; acc.w += p.w;
    }
    if (bias) {
      const float4 bv = *reinterpret_cast<const float4*>(bias + (i % N));
      acc.x += bv.x; acc.y += bv.y; acc.z += bv.z; acc.w += bv.w;
    }
    if (accumulate) {
      const float4 cv = *reinterpret_cast<const float4*>(C + i);
      acc.x += cv.x; acc.y += cv.y; acc.z += cv.z; acc.w += cv.w;
    }
    *reinterpret_cast<float4*>(C + i) = post4(acc, post, i);
  }
}

__device__ __forceinline__ float tf32_lo(float x) {
  const float h = __uint_as_float(__float_as_uint(x) & 0xffffe000u);   // what the tensor core reads of x
  return __uint_as_float((__float_as_uint(x - h) + 0x1000u) & 0xffffe000u);
}

// out[c, r] = in[r, c]   (in: rows x cols); with lo != nullptr also lo[c, r] = tf32_lo(in[r, c])
// (ldo = leading dimension of out / lo, >= rows: lets a K-range of a wider [N, K] matrix be filled)
// 64 x 64 tiles through shared memory, 16-byte global accesses on both sides when the shape allows
// (cols % 4 == 0 for the reads; rows % 4 == 0, ldo % 4 == 0 and 16-byte aligned bases for the
// writes -- checked on the host, VEC = false is the scalar fallback).
template <bool VEC>
__global__ void __launch_bounds__(256) transpose_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                        float* __restrict__ lo, uint32_t rows, uint32_t cols,
                                                        size_t ldo) {
  __shared__ float tile[64][65];
  const uint32_t bx = blockIdx.x * 64, by = blockIdx.y * 64;   // bx: first column, by: first row of the tile
  const uint32_t t = threadIdx.x;
  if (VEC) {
    const uint32_t c4 = (t & 15) * 4, r0 = t >> 4;             // 16 float4 per tile row, 16 rows per pass
#pragma unroll
    for (uint32_t p = 0; p < 4; p++) {
      const uint32_t r = r0 + 16 * p;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (by + r < rows && bx + c4 < cols) v = *reinterpret_cast<const float4*>(in + (size_t)(by + r) * cols + bx + c4);
      tile[r][c4] = v.x; tile[r][c4 + 1] = v.y; tile[r][c4 + 2] = v.z; tile[r][c4 + 3] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (uint32_t p = 0; p < 4; p++) {
      const uint32_t c = r0 + 16 * p;                          // output row = input column
      const uint32_t r4 = c4;                                  // 4 consecutive input rows = 16 output bytes
      if (bx + c < cols && by + r4 < rows) {
        const float4 v = make_float4(tile[r4][c], tile[r4 + 1][c], tile[r4 + 2][c], tile[r4 + 3][c]);
        const size_t o = (size_t)(bx + c) * ldo + by + r4;
        *reinterpret_cast<float4*>(out + o) = v;
        if (lo) *reinterpret_cast<float4*>(lo + o) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
      }
    }
  } else {
    const uint32_t tx = t & 63, ty = t >> 6;                   // 64 x 4
    for (uint32_t j = ty; j < 64; j += 4) {
      const uint32_t r = by + j, c = bx + tx;
      tile[j][tx] = (r < rows && c < cols) ? in[(size_t)r * cols + c] : 0.f;
    }
    __syncthreads();
    for (uint32_t j = ty; j < 64; j += 4) {
      const uint32_t c = bx + j, r = by + tx;
      if (c < cols && r < rows) {
        const float x = tile[tx][j];
        out[(size_t)c * ldo + r] = x;
        if (lo) lo[(size_t)c * ldo + r] = tf32_lo(x);
      }
    }
  }
}

static void launch_transpose(cudaStream_t stream, const float* in, float* out, float* lo, uint32_t rows, uint32_t cols,
                             size_t ldo) {
  dim3 g((cols + 63) / 64, (rows + 63) / 64);
  const bool vec = cols % 4 == 0 && rows % 4 == 0 && ldo % 4 == 0 &&
                   ((((uintptr_t)in) | ((uintptr_t)out) | ((uintptr_t)lo)) & 15) == 0;
  if (vec) transpose_kernel<true><<<g, 256, 0, stream>>>(in, out, lo, rows, cols, ldo);
  else transpose_kernel<false><<<g, 256, 0, stream>>>(in, out, lo, rows, cols, ldo);
}

// lo[i] = tf32_lo(in[i]): the low part of the 3xTF32 split for an operand that is already K-major
__global__ void __launch_bounds__(256) split_lo_kernel(const float4* __restrict__ in, float4* __restrict__ lo, size_t n4) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 x = in[i];
    lo[i] = make_float4(tf32_lo(x.x), tf32_lo(x.y), tf32_lo(x.z), tf32_lo(x.w));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major [rows, K] fp32 matrix, box = BK x box_rows, swizzle = one box row, zero fill out of bounds
int make_map(CUtensorMap* map, const float* ptr, uint32_t rows, uint32_t K, uint32_t box_rows, uint32_t bk) {
  EncodeTiledFn enc = get_encode();
  VKP_CHECK(enc, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {K, rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 4};
  cuuint32_t box[2] = {bk, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VKP_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return VKP_OK;
}

// [K, MN] fp32 matrix (MN contiguous, MN % 32 == 0) seen as {32 floats of MN, K, MN / 32}: one box brings
// `chunks` 128-byte-swizzled slabs of bk rows, i.e. an MN-major tile of 32 * chunks x bk
int make_map_mn(CUtensorMap* map, const float* ptr, uint32_t MN, uint32_t K, uint32_t chunks, uint32_t bk, int mode) {
  EncodeTiledFn enc = get_encode();
  VKP_CHECK(enc, "cuTensorMapEncodeTiled is not available from this driver");
  if (mode == 2) {   // plain 2-D view of the [K, MN] matrix, one 32-wide chunk per box
    cuuint64_t dims2[2] = {MN, K};
    cuuint64_t strides2[1] = {(cuuint64_t)MN * 4};
    cuuint32_t box2[2] = {32, bk};
    cuuint32_t estr2[2] = {1, 1};
    CUresult r2 = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims2, strides2, box2, estr2,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VKP_CHECK(r2 == CUDA_SUCCESS, "cuTensorMapEncodeTiled(MN-major, 2-D) failed with %d", (int)r2);
    return VKP_OK;
  }
  cuuint64_t dims[3] = {32, K, MN / 32};
  cuuint64_t strides[2] = {(cuuint64_t)MN * 4, 128};
  cuuint32_t box[3] = {32, bk, chunks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VKP_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(MN-major) failed with %d", (int)r);
  return VKP_OK;
}

// Alo / Btlo: pre-split low parts (MODE_PRESPLIT) or nullptr (converter warps split in shared memory)
// a_mn / b_mn: that operand (and its low part) lies as [K, MN], not [MN, K]
template <int BN, int BK>
int launch_tc(vkp_ctx* ctx, const float* A, const float* Bt, const float* Alo, const float* Btlo, float* C,
              const float* bias, uint32_t M, uint32_t N, uint32_t K, int accumulate,
              vkp_tc_chunks ch = vkp_tc_chunks{nullptr, 0, 0, 0, 1}, const vkp_tc_pull* pull = nullptr,
              int a_mn = 0, int b_mn = 0, vkp_gemm_post post = vkp_gemm_post{0, nullptr}) {
  using cfg = Cfg<BN, BK>;
  const bool presplit = Alo != nullptr;
  CUtensorMap tmA, tmB, tmAlo, tmBlo;
  if (a_mn) {
    VKP_TRY(make_map_mn(&tmA, A, M, K, BM / 32, BK, a_mn));
    VKP_TRY(make_map_mn(&tmAlo, presplit ? Alo : A, M, K, BM / 32, BK, a_mn));
  } else {
    VKP_TRY(make_map(&tmA, A, M, K, BM, BK));
    VKP_TRY(make_map(&tmAlo, presplit ? Alo : A, M, K, BM, BK));
  }
  if (b_mn) {
    VKP_TRY(make_map_mn(&tmB, Bt, N, K, BN / 32, BK, b_mn));
    VKP_TRY(make_map_mn(&tmBlo, presplit ? Btlo : Bt, N, K, BN / 32, BK, b_mn));
  } else {
    VKP_TRY(make_map(&tmB, Bt, N, K, BN, BK));
    VKP_TRY(make_map(&tmBlo, presplit ? Btlo : Bt, N, K, BN, BK));
  }
  static const bool rewrite_hi = getenv("VKP_TC_REWRITE_HI") != nullptr;
  auto kernel = presplit ? gemm_tc_kernel<BN, BK, MODE_PRESPLIT>
                         : (rewrite_hi ? gemm_tc_kernel<BN, BK, MODE_CONVERT_REWRITE_HI> : gemm_tc_kernel<BN, BK, MODE_CONVERT>);
  VKP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cfg::SMEM_BYTES));
  const uint32_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
  const uint32_t k_blocks = (K + BK - 1) / BK;
  // split-K when the output has fewer tiles than SMs and K is long: partials in workspace slot 0
  uint32_t splits = 1;
  if (ch.n_chunks > 1) ch.kb_per_chunk = (K / ch.n_chunks) / BK;   // caller guarantees divisibility
  if (ch.n_chunks <= 1 && tiles * 2 <= (uint32_t)ctx->sms && k_blocks >= 16) {
    splits = (uint32_t)ctx->sms / tiles;
    if (splits > k_blocks / 8) splits = k_blocks / 8;
    if (splits > 16) splits = 16;
    if (splits < 1) splits = 1;
  }
  uint32_t kb_per = (k_blocks + splits - 1) / splits;
  splits = (k_blocks + kb_per - 1) / kb_per;
  float* dst = C;
  if (splits > 1) {
    void* ws;
    VKP_TRY(vkp_workspace(ctx, 0, (size_t)splits * M * N * sizeof(float), &ws));
    dst = static_cast<float*>(ws);
  }
  vkp_tc_pull pl;
  memset(&pl, 0, sizeof(pl));
  if (pull) pl = *pull;
  const uint32_t work = tiles * splits;
  const unsigned grid = work < (uint32_t)ctx->sms ? work : (unsigned)ctx->sms;
  kernel<<<grid, NUM_THREADS, cfg::SMEM_BYTES, ctx->stream>>>(
      tmA, tmB, tmAlo, tmBlo, dst, splits > 1 ? nullptr : bias, M, N, K, splits > 1 ? 0 : accumulate, splits, kb_per, ch, pl,
      a_mn, b_mn, splits > 1 ? vkp_gemm_post{0, nullptr} : post);
  VKP_TRY(vkp_after_launch(ctx, "gemm_tc(tcgen05 3xTF32)"));
  if (splits > 1) {
    const unsigned rgrid = vkp_grid_for(ctx, ((size_t)M * N + 3) / 4, 256, 8);
    splitk_reduce_kernel<<<rgrid, 256, 0, ctx->stream>>>(dst, C, bias, M, N, splits, accumulate, post);
    VKP_TRY(vkp_after_launch(ctx, "gemm_splitk_reduce"));
  }
  return VKP_OK;
}

// 128x224 instead of 128x256 tiles when that saves whole waves: 1024 x 8192 outputs (a rank's rows of the
// fixed-size 8192^3 product on 8 GPUs) are 256 tiles = 1.73 waves of 148 CTAs -> 2, but 8 x 37 = 296 tiles of
// 224 columns = exactly 2 waves that are 12.5 % shorter (the last column tile is ragged: TMA zero-fills, the
// epilogue masks).  VKP_TC_BN224=0 keeps 256.
static bool prefer_bn224(const vkp_ctx* ctx, uint32_t M, uint32_t N) {
  static const bool off = getenv("VKP_TC_BN224") && getenv("VKP_TC_BN224")[0] == '0';
  if (off || N < 1024) return false;
  const uint64_t mt = (M + BM - 1) / BM, sms = (uint64_t)ctx->sms;
  const uint64_t t256 = mt * ((N + 255) / 256), t224 = mt * ((N + 223) / 224);
  const uint64_t c256 = (t256 + sms - 1) / sms * 256, c224 = (t224 + sms - 1) / sms * 224;
  return c224 * 100 <= c256 * 90;    // at least 10 % fewer column-steps: a 224-wide tile runs ~5 % below a 256-wide one per column (profiles/r02_bn224.txt)
}

}  // namespace

int vkp_gemm_tc_supported(int transA, int transB, uint32_t M, uint32_t N, uint32_t K, const float* A,
                          const float* B, float* C, int forced) {
  if (!forced) {
    if (getenv("VKP_DISABLE_TC")) return 0;
    // narrow outputs (nn.Dense with a handful of classes: forward X W^T + b, N <= 32): one 128 x 32 tile per
    // 128 rows, the pass over X is the cost; both operands must be K-major as they lie
    const bool narrow = !transA && transB && N >= 8 && N <= 32 && M >= 1024 && K >= 256 && !getenv("VKP_TC_NO_NARROW");
    if (!narrow && (M < 128 || N < 128 || K < 32)) return 0;  // small problems: SIMT kernel
    if ((uint64_t)M * N * K < (1ull << 22)) return 0;
  }
  if (K == 0) return 0;
  if (K % 4 || N % 4 || M % 4) return 0;                      // 16-byte global strides for TMA / float4 stores
  if (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) return 0;
  return 1;
}

// The pre-split variant pays one extra pass over the operands (8 B per element) to take the
// converter warps off the shared-memory pipe; worth it once the contraction is long next to that
// pass.  VKP_TC_PRESPLIT=0/1 forces the choice.
static bool use_presplit(uint32_t M, uint32_t N, uint32_t K) {
  static const char* env = getenv("VKP_TC_PRESPLIT");
  if (env) return env[0] == '1';
  (void)K;
  // pre-pass ~ (M + N) K bytes against a contraction ~ M N K: config 5's Dense GEMMs (8192 x 1024 x 1024 forward,
  // 1024 x 1024 x 8192 weight gradient) gain 30-40 us each over the converter variant (52-57 % tensor pipe)
  return M >= 512 && N >= 512 && (uint64_t)M * N >= 512ull * ((uint64_t)M + N);
}

int vkp_gemm_tc(vkp_ctx* ctx, int transA, int transB, uint32_t M, uint32_t N, uint32_t K, const float* A,
                const float* B, float* C, const float* bias, int accumulate, vkp_gemm_post post) {
  // A as [M,K] and B as [N,K] are K-major; the other two storage orders are taken as MN-major tiles
  // when their leading dimension allows the 3-D tensor map, else transposed once into the workspace
  static bool mn_ok = !(getenv("VKP_TC_MN") && getenv("VKP_TC_MN")[0] == '0');
  const bool presplit = use_presplit(M, N, K);
  for (int attempt = 0; attempt < 2; attempt++) {
    static const int mn_mode = getenv("VKP_TC_MN_TMA") ? atoi(getenv("VKP_TC_MN_TMA")) : 1;   // 1: one 3-D box per tile (default), 2: one 2-D box per 32-wide chunk
    const bool a_mn = transA && mn_ok && M % 32 == 0;
    const bool b_mn = !transB && mn_ok && N % 32 == 0;
    const bool a_tr = transA && !a_mn, b_tr = !transB && !b_mn;
    const float* Ak = A;
    const float* Bk = B;
    const float* Alo = nullptr;
    const float* Blo = nullptr;
    const size_t a_elems = (size_t)M * K, b_elems = (size_t)N * K;
    size_t need = 0;
    if (a_tr) need += a_elems * 4;
    if (b_tr) need += b_elems * 4;
    if (presplit) need += (a_elems + b_elems) * 4;
    if (need) {
      void* ws;
      VKP_TRY(vkp_workspace(ctx, 1, need, &ws));
      float* w = static_cast<float*>(ws);
      float* alo = nullptr;
      float* blo = nullptr;
      if (presplit) {
        alo = w; w += a_elems;
        blo = w; w += b_elems;
        Alo = alo; Blo = blo;
      }
      if (a_tr) {     // A stored [K, M] -> [M, K]
        launch_transpose(ctx->stream, A, w, alo, K, M, K);
        VKP_TRY(vkp_after_launch(ctx, "transpose(A)"));
        Ak = w;
        w += a_elems;
      } else if (presplit) {   // low parts in the operand's own layout (element-wise)
        split_lo_kernel<<<vkp_grid_for(ctx, a_elems / 4, 256, 8), 256, 0, ctx->stream>>>(
            reinterpret_cast<const float4*>(A), reinterpret_cast<float4*>(alo), a_elems / 4);
        VKP_TRY(vkp_after_launch(ctx, "split_lo(A)"));
      }
      if (b_tr) {     // B stored [K, N] -> [N, K]
        launch_transpose(ctx->stream, B, w, blo, K, N, K);
        VKP_TRY(vkp_after_launch(ctx, "transpose(B)"));
        Bk = w;
      } else if (presplit) {
        split_lo_kernel<<<vkp_grid_for(ctx, b_elems / 4, 256, 8), 256, 0, ctx->stream>>>(
            reinterpret_cast<const float4*>(B), reinterpret_cast<float4*>(blo), b_elems / 4);
        VKP_TRY(vkp_after_launch(ctx, "split_lo(B)"));
      }
    }
    if (bias && (((uintptr_t)bias) & 15)) return vkp_set_error("vkp_gemm_tc: bias must be 16-byte aligned");
    const bool wide = getenv("VKP_TC_BN128") == nullptr && (N % 256 == 0 || N >= 1024);
    // pre-split tiles need no converter pass, so shorter k-blocks (64-byte swizzle rows) buy a 4-deep
    // ring in the same shared memory: 8192^3 4.46 -> 4.27 ms; with converters BK=32 stays ahead
    static const char* bk_env = getenv("VKP_TC_BK16");
    const bool bk16 = bk_env ? bk_env[0] == '1' : presplit;
    const vkp_tc_chunks no_chunks{nullptr, 0, 0, 0, 1};
    int rc;
    const int am = a_mn ? mn_mode : 0, bm = b_mn ? mn_mode : 0;
    if (N <= 32) rc = launch_tc<32, 32>(ctx, Ak, Bk, Alo, Blo, C, bias, M, N, K, accumulate, no_chunks, nullptr, am, bm, post);
    else if (wide && bk16 && !bm && prefer_bn224(ctx, M, N)) rc = launch_tc<224, 16>(ctx, Ak, Bk, Alo, Blo, C, bias, M, N, K, accumulate, no_chunks, nullptr, am, bm, post);
    else if (wide && bk16) rc = launch_tc<256, 16>(ctx, Ak, Bk, Alo, Blo, C, bias, M, N, K, accumulate, no_chunks, nullptr, am, bm, post);
    else if (wide) rc = launch_tc<256, 32>(ctx, Ak, Bk, Alo, Blo, C, bias, M, N, K, accumulate, no_chunks, nullptr, am, bm, post);
    else rc = launch_tc<128, 32>(ctx, Ak, Bk, Alo, Blo, C, bias, M, N, K, accumulate, no_chunks, nullptr, am, bm, post);
    if (rc == VKP_OK || !(a_mn || b_mn) || !strstr(vkp_last_error(), "MN-major")) return rc;
    mn_ok = false;      // this driver refuses the 3-D map: transposes from now on
  }
  return VKP_ERR;
}

// ---- pieces of the row-sharded matmul (vkp_comm.cu): operands pre-split, K arriving in ranges ----
int vkp_tc_split_lo(vkp_ctx* ctx, cudaStream_t stream, const float* in, float* lo, size_t elems) {
  split_lo_kernel<<<vkp_grid_for(ctx, elems / 4, 256, 8), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(lo), elems / 4);
  return vkp_after_launch(ctx, "split_lo");
}

// in: [rows, cols] row-major  ->  hi[c * ldo + r] = in[r, c], lo[...] = its TF32 low part
int vkp_tc_transpose_split(vkp_ctx* ctx, cudaStream_t stream, const float* in, uint32_t rows, uint32_t cols,
                           float* hi, float* lo, size_t ldo) {
  launch_transpose(stream, in, hi, lo, rows, cols, ldo);
  return vkp_after_launch(ctx, "transpose_split");
}

int vkp_gemm_tc_chunked(vkp_ctx* ctx, uint32_t M, uint32_t N, uint32_t K, const float* A, const float* Alo,
                        const float* Bt, const float* Btlo, float* C, vkp_tc_chunks ch, const vkp_tc_pull* pull) {
  VKP_CHECK(ch.n_chunks >= 1 && K % ch.n_chunks == 0 && (K / ch.n_chunks) % 32 == 0,
            "vkp_gemm_tc_chunked: K = %u does not split into %u ranges of whole k-blocks", K, ch.n_chunks);
  if (prefer_bn224(ctx, M, N)) return launch_tc<224, 16>(ctx, A, Bt, Alo, Btlo, C, nullptr, M, N, K, 0, ch, pull);
  if (N % 256 == 0 || N >= 1024) return launch_tc<256, 16>(ctx, A, Bt, Alo, Btlo, C, nullptr, M, N, K, 0, ch, pull);
  return launch_tc<128, 32>(ctx, A, Bt, Alo, Btlo, C, nullptr, M, N, K, 0, ch, pull);
}

// 3-D tiled tensor map over a float32 [d2, d1, d0] array (d0 contiguous), no swizzle, zero fill out
// of bounds: used by the TMA-staged strided-axis reduction (vkp_reduce.cu).
int vkp_tma_map_3d(void* map_out, const float* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t b0, uint32_t b1) {
  EncodeTiledFn enc = get_encode();
  VKP_CHECK(enc, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {(cuuint64_t)d0 * 4, (cuuint64_t)d0 * d1 * 4};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(static_cast<CUtensorMap*>(map_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(ptr), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VKP_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with %d", (int)r);
  return VKP_OK;
}
