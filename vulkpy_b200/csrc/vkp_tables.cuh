// vkp_tables.cuh -- device-side accessor for the 32-entry tables of vkp_math.cuh.
//
// Each lane of a warp keeps ONE entry of every table in registers; a lookup is a warp shuffle
// (no shared memory, no bank conflicts), so every lane of a warp must evaluate the table-driven
// functions together.  Long polynomial coefficients arrive as a kernel parameter (MathCoef) and
// are pinned in registers (a 64-bit literal would cost two UMOVs at every use).
#pragma once

#include "vkp_math.cuh"

namespace vkpt {

__device__ const double g_tab_rc[32] = {VKPM_TABLE_RC};   // 21 significant bits: low word zero
__device__ const double g_tab_l2[32] = {VKPM_TABLE_L2};
__device__ const double g_tab_e2[32] = {VKPM_TABLE_E2};

__device__ __forceinline__ double pin(double x) {   // opaque to the optimiser: stays in a register pair
  asm("" : "+d"(x));
  return x;
}

struct LaneTables {
  double lc_[6], ec_[4], log2e_;
  uint32_t rc_;    // high word of the lane's rc entry
  double l2_, e2_;
  __device__ explicit LaneTables(const vkpm::MathCoef& c) {
#pragma unroll
    for (int k = 0; k < 6; k++) lc_[k] = pin(c.lc[k]);
#pragma unroll
    for (int k = 0; k < 4; k++) ec_[k] = pin(c.ec[k]);
    log2e_ = pin(c.log2e);
    const int lane = threadIdx.x & 31;
    rc_ = (uint32_t)__double2hiint(g_tab_rc[lane]);
    l2_ = g_tab_l2[lane];
    e2_ = g_tab_e2[lane];
  }
  __device__ double lc(int i) const { return lc_[i]; }
  __device__ double ec(int i) const { return ec_[i]; }
  __device__ double log2e() const { return log2e_; }
  // shfl.sync.idx with a full-width segment takes the source lane from the low 5 bits of its index
  // operand, so callers pass unmasked bit fields (the interval number sits in bits 0..4)
  static __device__ __forceinline__ uint32_t shfl5(uint32_t v, uint32_t i) {
    uint32_t r;
    asm volatile("shfl.sync.idx.b32 %0, %1, %2, 0x1f, 0xffffffff;" : "=r"(r) : "r"(v), "r"(i));
    return r;
  }
  __device__ double rc(uint32_t i) const { return __hiloint2double((int)shfl5(rc_, i), 0); }
  __device__ double l2(uint32_t i) const {
    return __hiloint2double((int)shfl5((uint32_t)__double2hiint(l2_), i), (int)shfl5((uint32_t)__double2loint(l2_), i));
  }
  __device__ double e2(uint32_t i) const {
    return __hiloint2double((int)shfl5((uint32_t)__double2hiint(e2_), i), (int)shfl5((uint32_t)__double2loint(e2_), i));
  }
};

inline const vkpm::MathCoef& host_coef() {
  static const vkpm::MathCoef c = vkpm::make_math_coef();
  return c;
}

}  // namespace vkpt
