// vkp_submit.cu -- op table and the single submit entry point.
//
// Replaces GPU::submit<N, Params> + the per-path Op cache (vulkpy/_vkarray.cc:527-548) and the
// 11 pybind overloads that select on the parameter struct (:770-791).  The reference names a
// kernel by the path of its .spv file (vulkpy/util.py:58-72); here the base name of that file
// maps to an integer id once (vkp_op_id) and the id indexes a static table.
#include "vkp_common.cuh"

namespace {

struct OpEntry {
  const char* name;
  int fam;
  int sub;
};

#define BIN7(F, PRE, SUF)                                                          \
  {PRE "add" SUF, F, VKB_ADD}, {PRE "sub" SUF, F, VKB_SUB}, {PRE "mul" SUF, F, VKB_MUL}, \
  {PRE "div" SUF, F, VKB_DIV}, {PRE "max" SUF, F, VKB_MAX}, {PRE "min" SUF, F, VKB_MIN}, \
  {PRE "pow" SUF, F, VKB_POW}

#define UN20(F, PRE)                                                                              \
  {PRE "abs", F, VKU_ABS}, {PRE "sign", F, VKU_SIGN}, {PRE "sin", F, VKU_SIN}, {PRE "cos", F, VKU_COS}, \
  {PRE "tan", F, VKU_TAN}, {PRE "asin", F, VKU_ASIN}, {PRE "acos", F, VKU_ACOS},                  \
  {PRE "atan", F, VKU_ATAN}, {PRE "sinh", F, VKU_SINH}, {PRE "cosh", F, VKU_COSH},                \
  {PRE "tanh", F, VKU_TANH}, {PRE "asinh", F, VKU_ASINH}, {PRE "acosh", F, VKU_ACOSH},            \
  {PRE "atanh", F, VKU_ATANH}, {PRE "exp", F, VKU_EXP}, {PRE "log", F, VKU_LOG},                  \
  {PRE "exp2", F, VKU_EXP2}, {PRE "log2", F, VKU_LOG2}, {PRE "sqrt", F, VKU_SQRT},                \
  {PRE "invsqrt", F, VKU_INVSQRT}

#define RED4(F, SUF)                                                                     \
  {"sum" SUF, F, VKR_SUM}, {"prod" SUF, F, VKR_PROD}, {"maximum" SUF, F, VKR_MAX}, \
  {"minimum" SUF, F, VKR_MIN}

// the 121 shaders compiled by the reference's setup.py:11-48
const OpEntry g_ops[] = {
    BIN7(VKF_BIN, "", ""),
    BIN7(VKF_IBIN, "i", ""),
    BIN7(VKF_SCALAR, "", "_scalar"),
    {"rsub_scalar", VKF_SCALAR, VKB_RSUB},
    {"rdiv_scalar", VKF_SCALAR, VKB_RDIV},
    {"rpow_scalar", VKF_SCALAR, VKB_RPOW},
    BIN7(VKF_ISCALAR, "i", "_scalar"),
    BIN7(VKF_BCAST, "", "_broadcast"),
    BIN7(VKF_IBCAST, "i", "_broadcast"),
    {"broadcast", VKF_BCAST_COPY, 0},
    UN20(VKF_UNARY, ""),
    UN20(VKF_IUNARY, "i"),
    {"clamp", VKF_CLAMP, (VKC_VV << 1)},
    {"iclamp", VKF_CLAMP, (VKC_VV << 1) | 1},
    {"clamp_sv", VKF_CLAMP, (VKC_SV << 1)},
    {"iclamp_sv", VKF_CLAMP, (VKC_SV << 1) | 1},
    {"clamp_vs", VKF_CLAMP, (VKC_VS << 1)},
    {"iclamp_vs", VKF_CLAMP, (VKC_VS << 1) | 1},
    {"clamp_ss", VKF_CLAMP, (VKC_SS << 1)},
    {"iclamp_ss", VKF_CLAMP, (VKC_SS << 1) | 1},
    RED4(VKF_REDUCE, ""),
    RED4(VKF_REDUCE_SG, "_v1.3"),
    RED4(VKF_REDUCE_AXIS, "_axis"),
    RED4(VKF_REDUCE_AXIS_RB, "_axis_rebroadcast"),
    {"gather", VKF_GATHER, 0},
    {"gather_axis", VKF_GATHER_AXIS, 0},
    {"matmul", VKF_MATMUL, 0},
    {"batch_affine", VKF_BATCH_AFFINE, 0},
    {"nn_cross_entropy", VKF_CE, 0},
    {"nn_cross_entropy_backward", VKF_CE_BWD, 0},
    {"prng_box_muller", VKF_BOX_MULLER, 0},
    {"prng_ibox_muller", VKF_IBOX_MULLER, 0},
    {"prng_randrange", VKF_RANDRANGE, 0},
    {"prng_xoshiro128pp_uint32", VKF_PRNG_U32, 0},
    {"prng_xoshiro128pp_float", VKF_PRNG_F32, 0},
};
constexpr int g_nops = sizeof(g_ops) / sizeof(g_ops[0]);

}  // namespace

extern "C" int vkp_op_count(void) { return g_nops; }

extern "C" const char* vkp_op_name(int op) { return (op >= 0 && op < g_nops) ? g_ops[op].name : nullptr; }

extern "C" int vkp_op_id(const char* name) {
  if (!name) return -1;
  // accept "add", "add.spv" and ".../shader/add.spv"
  const char* base = strrchr(name, '/');
  base = base ? base + 1 : name;
  size_t len = strlen(base);
  if (len > 4 && strcmp(base + len - 4, ".spv") == 0) len -= 4;
  for (int i = 0; i < g_nops; i++)
    if (strlen(g_ops[i].name) == len && strncmp(g_ops[i].name, base, len) == 0) return i;
  return -1;
}

extern "C" int vkp_submit(vkp_ctx* ctx, int op, void* const* bufs, int nbuf, const void* params,
                          size_t params_bytes, vkp_job** job) {
  VKP_CHECK(ctx, "vkp_submit: null context");
  VKP_CHECK(op >= 0 && op < g_nops, "Unknown Operation");  // _vkarray.cc:752
  VKP_RANGE(g_ops[op].name);
  VKP_CHECK(bufs && params && nbuf >= 1 && nbuf <= 4, "vkp_submit: bad buffer list");
  for (int i = 0; i < nbuf; i++) VKP_CHECK(bufs[i], "vkp_submit(%s): binding %d is null", g_ops[op].name, i);
  VKP_TRY(vkp_make_current(ctx));
  std::lock_guard<std::mutex> g(ctx->mu);
  const OpEntry& e = g_ops[op];
  VKP_TRY(vkp_prepare_buffers(ctx, bufs, nbuf));  // host-side shape bindings are not pool blocks: ignored
  int r;
  switch (e.fam) {
    case VKF_BCAST:
    case VKF_IBCAST:
    case VKF_BCAST_COPY:
      r = vkp_launch_broadcast(ctx, e.fam, e.sub, bufs, nbuf, params, params_bytes);
      break;
    case VKF_REDUCE:
    case VKF_REDUCE_SG:
    case VKF_REDUCE_AXIS:
    case VKF_REDUCE_AXIS_RB:
      r = vkp_launch_reduce(ctx, e.fam, e.sub, bufs, nbuf, params, params_bytes);
      break;
    case VKF_GATHER:
    case VKF_GATHER_AXIS:
      r = vkp_launch_gather(ctx, e.fam, e.sub, bufs, nbuf, params, params_bytes);
      break;
    case VKF_MATMUL: {  // A [rowA, contract], B [contract, columnB], C
      VKP_CHECK(nbuf == 3 && params_bytes == sizeof(vkp_matmul_params), "matmul: bad arguments");
      const auto* p = static_cast<const vkp_matmul_params*>(params);
      r = vkp_launch_gemm(ctx, 0, 0, p->rowA, p->columnB, p->contractSize, (const float*)bufs[0],
                          (const float*)bufs[1], (float*)bufs[2], nullptr, VKP_GEMM_AUTO);
      break;
    }
    case VKF_BATCH_AFFINE: {  // W [out, in], b [out], X [batch, in], Y [batch, out]
      VKP_CHECK(nbuf == 4 && params_bytes == sizeof(vkp_batchaffine_params), "batch_affine: bad arguments");
      const auto* p = static_cast<const vkp_batchaffine_params*>(params);
      r = vkp_launch_gemm(ctx, 0, 1, p->batch_size, p->output_size, p->input_size, (const float*)bufs[2],
                          (const float*)bufs[0], (float*)bufs[3], (const float*)bufs[1], VKP_GEMM_AUTO);
      break;
    }
    default:
      r = vkp_launch_elementwise(ctx, e.fam, e.sub, bufs, nbuf, params, params_bytes);
  }
  if (r != VKP_OK) return r;
  return vkp_finish_op(ctx, job);
}
