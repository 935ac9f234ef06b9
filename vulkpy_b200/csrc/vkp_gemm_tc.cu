// vkp_gemm_tc.cu -- tcgen05 3xTF32 GEMM (placeholder until the kernel lands: reports "unsupported"
// so that every shape takes the SIMT kernel).
#include "vkp_common.cuh"

int vkp_gemm_tc_supported(int, int, uint32_t, uint32_t, uint32_t, const float*, const float*, float*) {
  return 0;
}

int vkp_gemm_tc(vkp_ctx*, int, int, uint32_t, uint32_t, uint32_t, const float*, const float*, float*,
                const float*, int) {
  return vkp_set_error("vkp_gemm_tc: not built");
}
