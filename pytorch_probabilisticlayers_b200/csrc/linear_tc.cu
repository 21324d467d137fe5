// tcgen05 / TMEM / TMA path of the BayesLinear contraction (PL_PREC_BF16) -- placeholder until the
// fused sample+GEMM kernels land; reports PL_ERR_UNSUPPORTED so callers fail loudly.
#include "common.cuh"

namespace pl {
size_t linear_tc_workspace_bytes(const pl_linear_args*) { return 0; }
int linear_fwd_tc_launch(const pl_linear_args*) {
  set_error("pl_linear_fwd: PL_PREC_BF16 tensor-core path not built yet");
  return PL_ERR_UNSUPPORTED;
}
int linear_bwd_tc_launch(const pl_linear_args*) {
  set_error("pl_linear_bwd: PL_PREC_BF16 tensor-core path not built yet");
  return PL_ERR_UNSUPPORTED;
}
}  // namespace pl
