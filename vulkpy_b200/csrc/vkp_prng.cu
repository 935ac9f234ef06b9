// vkp_prng.cu -- xoshiro128++ streams, bit-exact with the reference.
//
// Replaces PRNG::Xoshiro128pp (vulkpy/_vkarray.cc:577-719), prng_xoshiro128pp_uint32.comp:26-43,
// prng_xoshiro128pp_float.comp:26-44 and, fused, random.py:60-124 + prng_box_muller.comp:19-32.
//
// Reference semantics that are reproduced exactly:
//   * seeding: four chained splitmix64 outputs, truncated to 32 bit, form lane 0; lane i is
//     lane i-1 advanced by the reference's jump(), which writes the accumulators back after
//     EACH of the four JUMP words (_vkarray.cc:597-621) -- not the canonical jump;
//   * `size` lanes; a request of n numbers is served in chunks of `size`: out[c*size + lane] is
//     draw c of that lane, the last chunk advances only the first n % size lanes, and the state
//     persists between calls (_vkarray.cc:697-717);
//   * [0,1) mapping: uintBitsToFloat((x >> 9) | 0x3f800000) - 1.0.
//
// The reference issues ceil(n/size) dependent dispatches with the state round-tripping through
// memory.  xoshiro's state transition T is linear over GF(2), so here a lane's sequential stream
// is cut into segments of L = 2^m draws; thread (lane group, segment p) starts from
// T^(p*L) * state, obtained by multiplying with precomputed 128x128 bit matrices T^(2^k)
// (k = m + set bits of p), keeps the state in registers for L draws and stores with 8/16-byte
// vectors (adjacent lanes are adjacent in memory).  States are double buffered so a segment that
// starts late never sees the final state another segment already wrote.
#include "vkp_common.cuh"
#include "vkp_math.cuh"
#include "vkp_tables.cuh"

#include <cstdlib>
#include <random>

namespace {

constexpr int JUMP_LEVELS = 34;  // T^(2^k), k = 0..33

struct HostState { uint32_t s[4]; };

inline uint32_t rotl32(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }

inline uint32_t next_host(uint32_t (&s)[4]) {
  const uint32_t result = rotl32(s[0] + s[3], 7) + s[0];
  const uint32_t t = s[1] << 9;
  s[2] ^= s[0];
  s[3] ^= s[1];
  s[1] ^= s[2];
  s[0] ^= s[3];
  s[2] ^= t;
  s[3] = rotl32(s[3], 11);
  return result;
}

inline uint64_t splitmix64(uint64_t x) {
  uint64_t z = (x += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}

// the reference's jump(): accumulators are NOT reset between JUMP words and are written back
// after each word
inline void jump_ref(uint32_t (&s)[4]) {
  static const uint32_t JUMP[4] = {0x8764000bu, 0xf542d2d3u, 0x6fa035c3u, 0x77f2db5bu};
  uint32_t s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  for (int w = 0; w < 4; w++) {
    for (int b = 0; b < 32; b++) {
      if (JUMP[w] & (1u << b)) {
        s0 ^= s[0]; s1 ^= s[1]; s2 ^= s[2]; s3 ^= s[3];
      }
      next_host(s);
    }
    s[0] = s0; s[1] = s1; s[2] = s2; s[3] = s3;
  }
}

// 128x128 GF(2) matrix stored by columns: col[c] = image of basis vector e_c (bit c of the
// 128-bit state, word c/32, bit c%32)
struct BitMat { uint32_t col[128][4]; };

inline void matvec_host(const BitMat& m, const uint32_t (&v)[4], uint32_t (&out)[4]) {
  uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  for (int w = 0; w < 4; w++)
    for (int b = 0; b < 32; b++)
      if ((v[w] >> b) & 1u) {
        const uint32_t* c = m.col[w * 32 + b];
        a0 ^= c[0]; a1 ^= c[1]; a2 ^= c[2]; a3 ^= c[3];
      }
  out[0] = a0; out[1] = a1; out[2] = a2; out[3] = a3;
}

std::once_flag g_jump_once;
BitMat* g_jump_host = nullptr;  // [JUMP_LEVELS]

// Nibble-sliced form of the same matrices for the device: tab[k][j][v] = XOR of the columns 4j + b of
// T^(2^k) over the set bits b of v, so that a mat-vec is 32 table look-ups (one per state nibble) instead of
// 128 masked column loads.  Round 1/2 history: byte slices (16 look-ups into 4 KiB sub-tables, 64 KiB per level)
// halved the instruction count again, but the 32 lanes of a warp then read 32 DIFFERENT cache lines per look-up
// and the L1 tag stage, not the issue slots, set the pace of every jump (ncu: long-scoreboard stalls,
// profiles/r02_prng_full_64_lanes.md).  A 16-entry sub-table is 256 B = two lines: a warp's look-up touches at most
// two, and a whole level is 8 KiB (all 34 levels: 272 KiB, L1/L2 resident).
struct ByteMat { uint32_t e[32][16][4]; };
ByteMat* g_jump_bytes = nullptr;  // [JUMP_LEVELS]

void build_byte_tables() {
  g_jump_bytes = new ByteMat[JUMP_LEVELS];
  for (int k = 0; k < JUMP_LEVELS; k++)
    for (int j = 0; j < 32; j++) {
      uint32_t (*t)[4] = g_jump_bytes[k].e[j];
      t[0][0] = t[0][1] = t[0][2] = t[0][3] = 0;
      for (int v = 1; v < 16; v++) {
        const int low = __builtin_ctz(v);
        const uint32_t* c = g_jump_host[k].col[4 * j + low];
        const uint32_t* r = t[v & (v - 1)];
        for (int w = 0; w < 4; w++) t[v][w] = r[w] ^ c[w];
      }
    }
}

void build_jump_tables() {
  g_jump_host = new BitMat[JUMP_LEVELS];
  for (int c = 0; c < 128; c++) {  // T itself
    uint32_t s[4] = {0, 0, 0, 0};
    s[c / 32] = 1u << (c % 32);
    next_host(s);
    for (int w = 0; w < 4; w++) g_jump_host[0].col[c][w] = s[w];
  }
  for (int k = 1; k < JUMP_LEVELS; k++)  // squaring: (M*M) e_c = M (M e_c)
    for (int c = 0; c < 128; c++) matvec_host(g_jump_host[k - 1], g_jump_host[k - 1].col[c], g_jump_host[k].col[c]);
  build_byte_tables();
}

// ---- device ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rotl_d(uint32_t x, int k) { return __funnelshift_l(x, x, k); }

__device__ __forceinline__ uint32_t next_dev(uint4& s) {
  const uint32_t result = rotl_d(s.x + s.w, 7) + s.x;
  const uint32_t t = s.y << 9;
  s.z ^= s.x;
  s.w ^= s.y;
  s.y ^= s.z;
  s.x ^= s.w;
  s.z ^= t;
  s.w = rotl_d(s.w, 11);
  return result;
}

// m: one level of the nibble-sliced tables, [32][16] uint4
__device__ __forceinline__ uint4 matvec_dev(const uint4* __restrict__ m, uint4 v) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const uint4 c = __ldg(m + (i * 8 + b) * 16 + ((w[i] >> (4 * b)) & 15u));
      acc.x ^= c.x; acc.y ^= c.y; acc.z ^= c.z; acc.w ^= c.w;
    }
  }
  return acc;
}

constexpr size_t LEVEL_STRIDE = 32 * 16;   // uint4 per level

__device__ __forceinline__ float u2f01(uint32_t r) { return __uint_as_float((r >> 9) | 0x3f800000u) - 1.0f; }

enum { MODE_U32 = 0, MODE_F32 = 1, MODE_NORMAL = 2 };

// The stream is written once and read by a later kernel: streaming (evict-first) stores keep the
// thousands of concurrent segment streams from ageing in L2 and reaching DRAM as scattered lines.
template <int LPT> struct VecStore;
template <> struct VecStore<1> { static __device__ void st(uint32_t* p, const uint32_t* v, bool cs) { if (cs) __stcs(p, v[0]); else p[0] = v[0]; } };
template <> struct VecStore<2> { static __device__ void st(uint32_t* p, const uint32_t* v, bool cs) {
  if (cs) __stcs(reinterpret_cast<uint2*>(p), make_uint2(v[0], v[1])); else *reinterpret_cast<uint2*>(p) = make_uint2(v[0], v[1]); } };
template <> struct VecStore<4> { static __device__ void st(uint32_t* p, const uint32_t* v, bool cs) {
  if (cs) __stcs(reinterpret_cast<uint4*>(p), make_uint4(v[0], v[1], v[2], v[3])); else *reinterpret_cast<uint4*>(p) = make_uint4(v[0], v[1], v[2], v[3]); } };

// Segment start states by doubling (pre-pass of the draw kernels when a lane's stream is cut into >= 8
// segments): one mat-vec per segment instead of one per set bit of the segment number in every thread -- with
// the default 64 lanes the per-thread jump-ahead was a third of a 2^28-sample launch (ncu: long-scoreboard
// stalls on the table reads, profiles/r02_prng_full_64_lanes.md).
//   entry e of a lane = segment base + e * stride;  round k copies every known entry 2^k entries ahead:
//   starts[e + 2^k] = T^(2^(level0 + k)) starts[e].
// Two launches: COARSE (stride 64: every 64th segment of all lanes, entry 0 = T^skip state) and FINE (blockIdx.y =
// coarse block, stride 1: the 63 segments after each coarse one).  A CTA owns `lpc` lanes of one block, so rounds
// are separated by __syncthreads only; the mat-vec issues its 32 table reads before the first use (a round
// costs one load latency), and all threads of a round walk the same 8 KiB matrix.
constexpr int STARTS_BLOCK = 256;
constexpr uint32_t STARTS_FINE = 64;        // segments per coarse block
__device__ __forceinline__ uint4 matvec_wide(const uint4* __restrict__ m, uint4 v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint4 c[32];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int b = 0; b < 8; b++) c[i * 8 + b] = __ldg(m + (i * 8 + b) * 16 + ((w[i] >> (4 * b)) & 15u));
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1)
#pragma unroll
    for (int i = 0; i < h; i++) {
      c[i].x ^= c[i + h].x; c[i].y ^= c[i + h].y; c[i].z ^= c[i + h].z; c[i].w ^= c[i + h].w;
    }
  return c[0];
}
// mat-vec against a level staged in shared memory ([32][16] uint4): 32 LDS.128, no global latency
__device__ __forceinline__ uint4 matvec_smem(const uint4* m, uint4 v) {
  uint4 acc = make_uint4(0, 0, 0, 0);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
#pragma unroll
    for (int b = 0; b < 8; b++) {
      const uint4 c = m[(i * 8 + b) * 16 + ((w[i] >> (4 * b)) & 15u)];
      acc.x ^= c.x; acc.y ^= c.y; acc.z ^= c.z; acc.w ^= c.w;
    }
  }
  return acc;
}
// A CTA keeps the levels of ALL its rounds (<= 7 x 8 KiB) and the <= 512 states it works on in shared memory and
// walks over several (lane block, coarse block) pairs.  History (profiles/r02_ncu_rows_v6.md): with tables and
// states in global memory a round cost 5-7 us (32 dependent-address reads in ~7 batches of one L2 latency), with
// one level staged per round 2-4 us (each round still waited for its level and for the previous round's stores).
constexpr uint32_t STARTS_MAX_ROUNDS = 7;
constexpr uint32_t STARTS_STATES = 512;
__global__ void __launch_bounds__(STARTS_BLOCK)
xoshiro_starts_kernel(const uint4* __restrict__ state_in, uint4* starts, const uint4* __restrict__ jump,
                      uint32_t size, uint32_t nseg, uint32_t stride, uint32_t level0, uint64_t skip, int from_state,
                      uint32_t lpc, uint32_t nby, uint32_t nrounds) {
  extern __shared__ uint4 s_starts[];
  uint4* tab = s_starts;                                   // [nrounds][LEVEL_STRIDE]
  uint4* st = s_starts + (size_t)nrounds * LEVEL_STRIDE;   // [count][nl]
  for (uint32_t i = threadIdx.x; i < nrounds * LEVEL_STRIDE; i += blockDim.x)
    tab[i] = __ldg(jump + (size_t)level0 * LEVEL_STRIDE + i);
  const uint32_t nbx = (size + lpc - 1) / lpc;
  for (uint32_t blk = blockIdx.x; blk < nbx * nby; blk += gridDim.x) {
    const uint32_t lbase = (blk % nbx) * lpc;
    const uint32_t nl = min(lpc, size - lbase);
    const uint32_t base = (blk / nbx) * STARTS_FINE;       // first segment of this block (0 for COARSE)
    const uint32_t left = nseg - base;
    const uint32_t count = stride == 1 ? min(STARTS_FINE, left) : (left + stride - 1) / stride;
    if (threadIdx.x < nl) {
      uint4 s;
      if (from_state) {
        s = state_in[lbase + threadIdx.x];
        for (uint32_t b = 0; (skip >> b) != 0; b++)
          if ((skip >> b) & 1ull) s = matvec_wide(jump + (size_t)b * LEVEL_STRIDE, s);
      } else {
        s = __ldcg(starts + (size_t)base * size + lbase + threadIdx.x);   // written by the COARSE launch
      }
      st[threadIdx.x] = s;
    }
    __syncthreads();       // also: the tables are in place
    for (uint32_t k = 0; (1u << k) < count; k++) {
      const uint32_t span = 1u << k;
      const uint32_t cnt = min(span, count - span);          // entries span .. span + cnt - 1 become known
      for (uint32_t i = threadIdx.x; i < cnt * nl; i += blockDim.x)
        st[span * nl + i] = matvec_smem(tab + (size_t)k * LEVEL_STRIDE, st[i]);
      __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < count * nl; i += blockDim.x)
      starts[(size_t)(base + (i / nl) * stride) * size + lbase + i % nl] = st[i];
    __syncthreads();
  }
}

// uniform streams (MODE_U32 / MODE_F32): n numbers, out[c*size + lane]
template <int LPT, int MODE>
__global__ void __launch_bounds__(128)
xoshiro_stream_kernel(const uint4* __restrict__ state_in, uint4* __restrict__ state_out,
                      uint32_t* __restrict__ out, const uint4* __restrict__ jump, uint32_t size,
                      uint64_t n_draw, uint32_t log2L, uint32_t nseg, bool cs, uint64_t skip,
                      const uint4* __restrict__ starts) {
  const uint32_t groups = size / LPT;
  const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (t >= (uint64_t)groups * nseg) return;
  const uint32_t p = (uint32_t)(t / groups);
  const uint32_t l0 = (uint32_t)(t % groups) * LPT;
  const uint64_t full = n_draw / size;          // draws every lane makes
  const uint32_t rem = (uint32_t)(n_draw % size);  // lanes < rem make one more
  const uint64_t start = (uint64_t)p << log2L;
  const uint64_t seg_end = start + (1ull << log2L);

  uint4 s[LPT];
  if (starts) {   // pre-computed by xoshiro_starts_kernel (skip included)
#pragma unroll
    for (int q = 0; q < LPT; q++) s[q] = starts[(size_t)p * size + l0 + q];
  } else {
#pragma unroll
    for (int q = 0; q < LPT; q++) s[q] = state_in[l0 + q];
    // jump ahead by skip + p * 2^log2L steps (skip: whole chunks a preceding vkp_rng_advance left pending)
    const uint64_t J = skip + start;
    for (uint32_t b = 0; (J >> b) != 0; b++) {
      if ((J >> b) & 1ull) {
        const uint4* m = jump + (size_t)b * LEVEL_STRIDE;
#pragma unroll
        for (int q = 0; q < LPT; q++) s[q] = matvec_dev(m, s[q]);
      }
    }
  }
  if (p == 0) {  // lanes that draw nothing keep their (skipped-ahead) state
#pragma unroll
    for (int q = 0; q < LPT; q++)
      if (full == 0 && l0 + q >= rem) state_out[l0 + q] = s[q];
  }

  const uint64_t e1 = seg_end < full ? seg_end : full;
  for (uint64_t c = start; c < e1; c++) {
    uint32_t r[LPT];
#pragma unroll
    for (int q = 0; q < LPT; q++) r[q] = next_dev(s[q]);
    if (MODE == MODE_F32) {
#pragma unroll
      for (int q = 0; q < LPT; q++) r[q] = __float_as_uint(u2f01(r[q]));
    }
    VecStore<LPT>::st(out + c * size + l0, r, cs);
  }
  // tail chunk: draw index `full`, lanes < rem only
  if (rem != 0 && full >= start && full < seg_end) {
    const uint64_t j = full * size + l0;
#pragma unroll
    for (int q = 0; q < LPT; q++)
      if (l0 + q < rem) {
        const uint32_t r = next_dev(s[q]);
        out[j + q] = (MODE == MODE_F32) ? __float_as_uint(u2f01(r)) : r;
      }
  }
  // the segment that contains a lane's last draw publishes the lane's new state
#pragma unroll
  for (int q = 0; q < LPT; q++) {
    const uint64_t cnt = full + ((l0 + q < rem) ? 1 : 0);
    if (cnt > start && cnt <= seg_end) state_out[l0 + q] = s[q];
  }
}

// Fused uniform -> Box-Muller (random.py:60-124 + prng_box_muller.comp:19-32): a thread owns the
// lane pair (l0, l0+1), whose draws are the adjacent outputs (2i, 2i+1); the uniforms never leave
// registers.  n_draw = n rounded up to even (the reference draws n+1 uniforms for odd n).
template <bool FAST, bool UNIT>
__global__ void __launch_bounds__(128)
xoshiro_normal_kernel(const uint4* __restrict__ state_in, uint4* __restrict__ state_out, float* __restrict__ out,
                      const uint4* __restrict__ jump, uint32_t size, uint64_t n_draw, uint64_t n_out,
                      uint32_t log2L, uint32_t nseg, float mean, float stddev, uint64_t skip,
                      const uint4* __restrict__ starts) {
  const uint32_t groups = size / 2;
  const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (t >= (uint64_t)groups * nseg) return;
  const uint32_t p = (uint32_t)(t / groups);
  const uint32_t l0 = (uint32_t)(t % groups) * 2;
  const uint64_t full = n_draw / size;
  const uint32_t rem = (uint32_t)(n_draw % size);   // even, like size and l0
  const uint64_t start = (uint64_t)p << log2L;
  const uint64_t seg_end = start + (1ull << log2L);

  uint4 s0, s1;
  if (starts) {
    s0 = starts[(size_t)p * size + l0];
    s1 = starts[(size_t)p * size + l0 + 1];
  } else {
    s0 = state_in[l0];
    s1 = state_in[l0 + 1];
    const uint64_t J = skip + start;
    for (uint32_t b = 0; (J >> b) != 0; b++) {
      if ((J >> b) & 1ull) {
        const uint4* m = jump + (size_t)b * LEVEL_STRIDE;
        s0 = matvec_dev(m, s0);
        s1 = matvec_dev(m, s1);
      }
    }
  }
  if (p == 0 && full == 0 && l0 >= rem) {
    state_out[l0] = s0;
    state_out[l0 + 1] = s1;
  }
  // 1 - u = 2 - f for f = 1.bits in [1, 2): both exact, so one subtract replaces two
#define VKP_BM_PAIR(o0, o1)                                                                          \
  {                                                                                                  \
    const float om = 2.0f - __uint_as_float((next_dev(s0) >> 9) | 0x3f800000u);                      \
    const float u1 = u2f01(next_dev(s1));                                                            \
    if (FAST) vkpm::box_muller_fast(om, u1, mean, stddev, o0, o1);                                   \
    else vkpm::box_muller_core(om, u1, mean, stddev, o0, o1);                                        \
  }
  const uint64_t e1 = seg_end < full ? seg_end : full;
  uint64_t iters = e1 > start ? e1 - start : 0;
  float* o = out + start * size + l0;
  // for an odd request the very last pair stores one value only; it is some thread's final emit
  const bool half_last = (n_out & 1) && rem == 0 && iters > 0 && e1 == full && l0 == size - 2;
  if (half_last) iters--;
  if (FAST) {
    // four pairs per trip: the logarithm's series (one lane in 1024, see vkpm::box_muller_fast) is tested once
    // per trip through the minimum of the four u0, so the common path carries no series instructions
    for (; iters >= 4; iters -= 4) {
      float om[4], u1[4], L[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        om[k] = 2.0f - __uint_as_float((next_dev(s0) >> 9) | 0x3f800000u);
        u1[k] = u2f01(next_dev(s1));
        L[k] = vkpm::bm_fast_L(om[k]);
      }
      if (fmaxf(fmaxf(om[0], om[1]), fmaxf(om[2], om[3])) > 1.0f - vkpm::BM_SERIES_BELOW) {
#pragma unroll
        for (int k = 0; k < 4; k++)
          if (1.0f - om[k] < vkpm::BM_SERIES_BELOW) L[k] = vkpm::bm_fast_series(1.0f - om[k]);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        float o0, o1;
        vkpm::bm_fast_finish<UNIT>(L[k], u1[k], mean, stddev, o0, o1);
        __stcs(reinterpret_cast<float2*>(o), make_float2(o0, o1));
        o += size;
      }
    }
  }
  while (iters > 0) {
    const uint32_t batch = iters > 0x40000000ull ? 0x40000000u : (uint32_t)iters;
    for (uint32_t i = 0; i < batch; i++) {
      float o0, o1;
      VKP_BM_PAIR(o0, o1);
      __stcs(reinterpret_cast<float2*>(o), make_float2(o0, o1));
      o += size;
    }
    iters -= batch;
  }
  if (half_last) {
    float o0, o1;
    VKP_BM_PAIR(o0, o1);
    o[0] = o0;
  }
  if (rem != 0 && full >= start && full < seg_end && l0 < rem) {   // tail chunk
    const uint64_t j = full * size + l0;
    float o0, o1;
    VKP_BM_PAIR(o0, o1);
    if (j + 1 < n_out) *reinterpret_cast<float2*>(out + j) = make_float2(o0, o1);
    else out[j] = o0;
  }
#undef VKP_BM_PAIR
  const uint64_t cnt = full + ((l0 < rem) ? 1 : 0);
  if (cnt > start && cnt <= seg_end) {
    state_out[l0] = s0;
    state_out[l0 + 1] = s1;
  }
}

// discard n draws exactly as random(n) would consume them: every lane jumps floor(n/size) steps
// ahead through the T^(2^k) matrices, the first n % size lanes take one more step
__global__ void xoshiro_advance_kernel(const uint4* __restrict__ state_in, uint4* __restrict__ state_out,
                                       const uint4* __restrict__ jump, uint32_t size, uint64_t full, uint32_t rem) {
  const uint32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= size) return;
  uint4 s = state_in[l];
  for (uint32_t b = 0; (full >> b) != 0; b++)
    if ((full >> b) & 1ull) s = matvec_dev(jump + (size_t)b * LEVEL_STRIDE, s);
  if (l < rem) next_dev(s);
  state_out[l] = s;
}

}  // namespace

struct vkp_rng {
  vkp_ctx* ctx;
  uint32_t size;
  uint4* state[2];  // device, double buffered; state[cur] is current
  int cur;
  uint4* jump;      // device copy of the T^(2^k) tables
  // whole chunks (every lane one step each) skipped by vkp_rng_advance and not yet applied: consecutive
  // advances merge, and the next draw folds the skip into its own jump-ahead (no extra launch, no extra
  // trip of the state through memory).  Applied before anything else reads the state.
  uint64_t pending = 0;
  uint4* starts = nullptr;   // [nseg][size] segment start states of the launch in flight (xoshiro_starts_kernel)
  size_t starts_cap = 0;     // in uint4
};

static int rng_flush_pending(vkp_rng* rng);

static std::mutex g_jump_dev_mu;
static uint4* g_jump_dev[64] = {nullptr};

extern "C" int vkp_rng_create(vkp_ctx* ctx, uint32_t size, uint64_t seed, int has_seed, vkp_rng** out) {
  VKP_CHECK(ctx && out, "vkp_rng_create: null argument");
  VKP_CHECK(size >= 1, "Xoshiro128pp: size must be positive");
  VKP_TRY(vkp_make_current(ctx));
  if (!has_seed) seed = std::random_device{}();  // _vkarray.cc:679
  std::call_once(g_jump_once, build_jump_tables);

  std::vector<HostState> st(size);
  uint32_t s[4];
  for (int i = 0; i < 4; i++) {  // _vkarray.cc:659-663
    seed = splitmix64(seed);
    s[i] = (uint32_t)seed;
  }
  memcpy(st[0].s, s, 16);
  for (uint32_t i = 1; i < size; i++) {  // _vkarray.cc:666-672
    jump_ref(s);
    memcpy(st[i].s, s, 16);
  }

  vkp_rng* r = new vkp_rng();
  r->ctx = ctx;
  r->size = size;
  r->cur = 0;
  std::lock_guard<std::mutex> g(ctx->mu);
  {
    std::lock_guard<std::mutex> gj(g_jump_dev_mu);
    if (!g_jump_dev[ctx->device]) {
      VKP_CUDA(cudaMalloc(&g_jump_dev[ctx->device], sizeof(ByteMat) * JUMP_LEVELS));
      VKP_CUDA(cudaMemcpyAsync(g_jump_dev[ctx->device], g_jump_bytes, sizeof(ByteMat) * JUMP_LEVELS,
                               cudaMemcpyHostToDevice, ctx->stream));
    }
    r->jump = g_jump_dev[ctx->device];
  }
  VKP_CUDA(cudaMalloc(&r->state[0], 16ull * size));
  VKP_CUDA(cudaMalloc(&r->state[1], 16ull * size));
  VKP_CUDA(cudaMemcpyAsync(r->state[0], st.data(), 16ull * size, cudaMemcpyHostToDevice, ctx->stream));
  VKP_CUDA(cudaStreamSynchronize(ctx->stream));  // `st` is pageable and about to go out of scope
  *out = r;
  return VKP_OK;
}

extern "C" int vkp_rng_destroy(vkp_rng* rng) {
  if (!rng) return VKP_OK;
  vkp_ctx* ctx = rng->ctx;
  if (vkp_make_current(ctx) == VKP_OK) {
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(rng->state[0]);
    cudaFree(rng->state[1]);
    if (rng->starts) cudaFree(rng->starts);
  }
  delete rng;
  return VKP_OK;
}

extern "C" int vkp_rng_state(vkp_rng* rng, uint32_t* host_out) {
  VKP_CHECK(rng && host_out, "vkp_rng_state: null argument");
  vkp_ctx* ctx = rng->ctx;
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  VKP_TRY(rng_flush_pending(rng));
  VKP_CUDA(cudaMemcpyAsync(host_out, rng->state[rng->cur], 16ull * rng->size, cudaMemcpyDeviceToHost, ctx->stream));
  VKP_CUDA(cudaStreamSynchronize(ctx->stream));
  return VKP_OK;
}

template <int MODE>
static int rng_generate(vkp_rng* rng, void* out, uint64_t n_out, float mean, float stddev, vkp_job** job) {
  VKP_RANGE(MODE == MODE_NORMAL ? "vkp_rng_normal" : (MODE == MODE_F32 ? "vkp_rng_float" : "vkp_rng_uint32"));
  VKP_CHECK(rng && (out || n_out == 0), "vkp_rng: null argument");
  vkp_ctx* ctx = rng->ctx;
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  void* bufs[1] = {out};
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, 1));
  if (n_out > 0) {
    const uint32_t size = rng->size;
    const uint64_t n_draw = (MODE == MODE_NORMAL) ? ((n_out + 1) & ~1ull) : n_out;
    int lpt;
    // lanes per thread: the widest store the lane count allows (16 bytes for the default 64 lanes:
    // half a warp per segment row; the table-driven jump-ahead makes mixed segments in a warp cheap)
    if (MODE == MODE_NORMAL) lpt = 2;  // caller guarantees an even size
    else lpt = (size % 4 == 0) ? 4 : ((size % 2 == 0) ? 2 : 1);
    const uint32_t groups = size / lpt;
    const uint64_t draws_per_lane = (n_draw + size - 1) / size;
    // segment length: power of two, >= 256 draws, giving about sms * 1024 threads (VKP_PRNG_THREADS_PER_SM)
    static const uint64_t tps = getenv("VKP_PRNG_THREADS_PER_SM") ? (uint64_t)atoll(getenv("VKP_PRNG_THREADS_PER_SM")) : 1024;
    // streaming stores: +3-4 % at both lane counts (profiles/r02_prng_variants_v3.txt); VKP_PRNG_STCS=0 for A/B
    static const bool cs = !(getenv("VKP_PRNG_STCS") && getenv("VKP_PRNG_STCS")[0] == '0');
    uint64_t want_seg = ((uint64_t)ctx->sms * tps + groups - 1) / groups;
    if (want_seg < 1) want_seg = 1;
    uint64_t L = (draws_per_lane + want_seg - 1) / want_seg;
    uint32_t log2L = 8;
    while ((1ull << log2L) < L) log2L++;
    const uint64_t nseg64 = (draws_per_lane + (1ull << log2L) - 1) >> log2L;
    VKP_CHECK(nseg64 >= 1 && nseg64 < (1ull << 24), "vkp_rng: request too large");
    const uint32_t nseg = (uint32_t)nseg64;
    {
      uint32_t top = 0;
      while ((nseg - 1) >> top) top++;
      VKP_CHECK(log2L + top <= JUMP_LEVELS, "vkp_rng: jump table too small for this request");
    }
    // fold the pending skip into this launch when skip + (nseg << log2L) still fits the jump tables
    if (rng->pending && ((rng->pending + ((uint64_t)nseg << log2L)) >> (JUMP_LEVELS - 1)) != 0) VKP_TRY(rng_flush_pending(rng));
    const uint64_t skip = rng->pending;
    rng->pending = 0;
    const uint64_t threads = (uint64_t)groups * nseg;
    const unsigned grid = (unsigned)((threads + 127) / 128);
    const uint4* sin_ = rng->state[rng->cur];
    uint4* sout = rng->state[rng->cur ^ 1];
    // start states by doubling (VKP_PRNG_STARTS=0: every thread jumps by itself); beyond 8192 segments per lane
    // the last rounds of a one-lane CTA get long and the per-thread jump wins
    static const bool starts_on = !(getenv("VKP_PRNG_STARTS") && getenv("VKP_PRNG_STARTS")[0] == '0');
    const uint4* starts = nullptr;
    const uint32_t ncoarse = (nseg + STARTS_FINE - 1) / STARTS_FINE;
    if (starts_on && nseg >= 8 && ncoarse <= (1u << STARTS_MAX_ROUNDS) && (size_t)nseg * size <= ((size_t)1 << 22)) {
      const size_t need = (size_t)nseg * size;
      if (need > rng->starts_cap) {
        if (rng->starts) {
          VKP_CUDA(cudaStreamSynchronize(ctx->stream));
          VKP_CUDA(cudaFree(rng->starts));
          rng->starts = nullptr;
          rng->starts_cap = 0;
        }
        VKP_CUDA(cudaMalloc(&rng->starts, need * sizeof(uint4)));
        rng->starts_cap = need;
      }
      auto rounds_for = [](uint32_t count) { uint32_t r = 0; while ((1u << r) < count) r++; return r; };
      auto smem_for = [](uint32_t rounds) { return (size_t)(rounds * LEVEL_STRIDE + STARTS_STATES) * sizeof(uint4); };
      VKP_CUDA(cudaFuncSetAttribute(xoshiro_starts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem_for(STARTS_MAX_ROUNDS)));
      if (ncoarse > 1) {   // every 64th segment of all lanes; entry 0 = T^skip state
        const uint32_t lpc = std::max<uint32_t>(1, std::min<uint32_t>(size, STARTS_STATES / ncoarse));
        const uint32_t nbx = (size + lpc - 1) / lpc, r = rounds_for(ncoarse);
        xoshiro_starts_kernel<<<std::min<uint32_t>(nbx, 4u * ctx->sms), STARTS_BLOCK, smem_for(r), ctx->stream>>>(
            sin_, rng->starts, rng->jump, size, nseg, STARTS_FINE, log2L + 6, skip, 1, lpc, 1, r);
        VKP_TRY(vkp_after_launch(ctx, "xoshiro128pp_starts(coarse)"));
      }
      const uint32_t lpc = std::min<uint32_t>(size, STARTS_STATES / STARTS_FINE);
      const uint32_t nbx = (size + lpc - 1) / lpc, r = rounds_for(std::min<uint32_t>(nseg, STARTS_FINE));
      xoshiro_starts_kernel<<<(unsigned)std::min<uint64_t>((uint64_t)nbx * ncoarse, 3u * ctx->sms), STARTS_BLOCK, smem_for(r),
                              ctx->stream>>>(sin_, rng->starts, rng->jump, size, nseg, 1, log2L, skip,
                                             ncoarse > 1 ? 0 : 1, lpc, ncoarse, r);
      VKP_TRY(vkp_after_launch(ctx, "xoshiro128pp_starts(fine)"));
      starts = rng->starts;
    }
#define LAUNCH(LPT)                                                                                  \
  xoshiro_stream_kernel<LPT, (MODE == MODE_NORMAL ? MODE_F32 : MODE)><<<grid, 128, 0, ctx->stream>>>(     \
      sin_, sout, (uint32_t*)out, rng->jump, size, n_draw, log2L, nseg, cs, skip, starts)
    if (MODE == MODE_NORMAL) {
      // VKP_NORMAL_PRECISE=1: log / sqrt / sin / cos as <= 2 ulp float32 routines instead of the special-function
      // unit (see vkpm::box_muller_fast); read per call so a test can flip it
      if (vkp_normal_precise())
        xoshiro_normal_kernel<false, false><<<grid, 128, 0, ctx->stream>>>(sin_, sout, (float*)out, rng->jump, size,
                                                                           n_draw, n_out, log2L, nseg, mean, stddev, skip, starts);
      else if (mean == 0.0f && stddev == 1.0f)
        xoshiro_normal_kernel<true, true><<<grid, 128, 0, ctx->stream>>>(sin_, sout, (float*)out, rng->jump, size,
                                                                         n_draw, n_out, log2L, nseg, mean, stddev, skip, starts);
      else
        xoshiro_normal_kernel<true, false><<<grid, 128, 0, ctx->stream>>>(sin_, sout, (float*)out, rng->jump, size,
                                                                          n_draw, n_out, log2L, nseg, mean, stddev, skip, starts);
    }
    else if (lpt == 4) { LAUNCH(4); }
    else if (lpt == 2) { LAUNCH(2); }
    else { LAUNCH(1); }
#undef LAUNCH
    VKP_TRY(vkp_after_launch(ctx, "xoshiro128pp_stream"));
    rng->cur ^= 1;
  }
  return vkp_finish_op(ctx, job);
}

extern "C" int vkp_rng_uint32(vkp_rng* rng, uint32_t* out, uint32_t n, vkp_job** job) {
  return rng_generate<MODE_U32>(rng, out, n, 0.f, 1.f, job);
}

extern "C" int vkp_rng_float(vkp_rng* rng, float* out, uint32_t n, vkp_job** job) {
  return rng_generate<MODE_F32>(rng, out, n, 0.f, 1.f, job);
}

extern "C" int vkp_rng_normal(vkp_rng* rng, float* out, uint32_t n, float mean, float stddev, vkp_job** job) {
  VKP_CHECK(rng, "vkp_rng_normal: null argument");
  VKP_CHECK(rng->size % 2 == 0,
            "vkp_rng_normal: fused Box-Muller needs an even lane count; use random()+prng_box_muller");
  return rng_generate<MODE_NORMAL>(rng, out, n, mean, stddev, job);
}

// runs the advance kernel for the pending whole-chunk skip (caller holds ctx->mu, context current)
static int rng_flush_pending(vkp_rng* rng) {
  if (!rng->pending) return VKP_OK;
  vkp_ctx* ctx = rng->ctx;
  const uint64_t full = rng->pending;
  rng->pending = 0;
  xoshiro_advance_kernel<<<(rng->size + 127) / 128, 128, 0, ctx->stream>>>(
      rng->state[rng->cur], rng->state[rng->cur ^ 1], rng->jump, rng->size, full, 0u);
  VKP_TRY(vkp_after_launch(ctx, "xoshiro128pp_advance"));
  rng->cur ^= 1;
  ctx->seq++;
  return VKP_OK;
}

extern "C" int vkp_rng_advance(vkp_rng* rng, uint64_t n) {
  VKP_CHECK(rng, "vkp_rng_advance: null argument");
  if (n == 0) return VKP_OK;
  vkp_ctx* ctx = rng->ctx;
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  if (n % rng->size == 0 && ((rng->pending + n / rng->size) >> (JUMP_LEVELS - 2)) == 0) {
    rng->pending += n / rng->size;      // applied by the next draw (or by whoever reads the state)
    return VKP_OK;
  }
  VKP_TRY(rng_flush_pending(rng));
  const uint64_t full = n / rng->size;
  VKP_CHECK((full >> (JUMP_LEVELS - 1)) == 0, "vkp_rng_advance: jump too long");
  xoshiro_advance_kernel<<<(rng->size + 127) / 128, 128, 0, ctx->stream>>>(
      rng->state[rng->cur], rng->state[rng->cur ^ 1], rng->jump, rng->size, full, (uint32_t)(n % rng->size));
  VKP_TRY(vkp_after_launch(ctx, "xoshiro128pp_advance"));
  rng->cur ^= 1;
  ctx->seq++;
  return VKP_OK;
}
