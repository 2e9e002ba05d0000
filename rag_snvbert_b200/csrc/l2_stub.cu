#include "common.cuh"
#include "kernels.cuh"
namespace snv {
size_t l2_plan(L2SearchParams&) { set_error("L2 path not built"); return (size_t)-1; }
int l2_launch(const L2SearchParams&, cudaStream_t) { set_error("L2 path not built"); return SNV_ERR_UNSUPPORTED; }
int l2_prep_launch(const float*, int64_t, int64_t, int, bool, int, float*, float*, cudaStream_t) { set_error("L2 path not built"); return SNV_ERR_UNSUPPORTED; }
int l2_operand_depth(int64_t d, int mode) { int64_t k = mode == SNV_L2_TF32X3 ? 3 * d : d; return (int)((k + 31) / 32 * 32); }
}
