// pl_linear_* entry points: argument validation + dispatch on the requested precision path.
#include "common.cuh"

namespace pl {
size_t linear_simt_workspace_bytes(const pl_linear_args* a);
int linear_fwd_simt_launch(const pl_linear_args* a);
int linear_bwd_simt_launch(const pl_linear_args* a);
size_t linear_tc_workspace_bytes(const pl_linear_args* a);
size_t linear_tc_state_bytes(const pl_linear_args* a);
int linear_fwd_tc_launch(const pl_linear_args* a);
int linear_bwd_tc_launch(const pl_linear_args* a);
}  // namespace pl

static inline bool is_tc(int p) {
  return p == PL_PREC_BF16 || p == PL_PREC_TF32 || p == PL_PREC_BF16_FAST || p == PL_PREC_TF32_FAST;
}

extern "C" size_t pl_linear_workspace_bytes(const pl_linear_args* a) {
  if (!a) return 0;
  if (is_tc(a->precision)) return pl::linear_tc_workspace_bytes(a);
  return pl::linear_simt_workspace_bytes(a);
}

extern "C" size_t pl_linear_state_bytes(const pl_linear_args* a) {
  if (!a || !is_tc(a->precision)) return 0;
  return pl::linear_tc_state_bytes(a);
}

extern "C" int pl_linear_fwd(const pl_linear_args* a) {
  PL_CHECK_ARG(a, "pl_linear_fwd: NULL args");
  switch (a->precision) {
    case PL_PREC_FP32: return pl::linear_fwd_simt_launch(a);
    case PL_PREC_BF16:
    case PL_PREC_TF32:
    case PL_PREC_BF16_FAST:
    case PL_PREC_TF32_FAST: return pl::linear_fwd_tc_launch(a);
    default:
      pl::set_error("pl_linear_fwd: precision %d not implemented", a->precision);
      return PL_ERR_UNSUPPORTED;
  }
}

extern "C" int pl_linear_bwd(const pl_linear_args* a) {
  PL_CHECK_ARG(a, "pl_linear_bwd: NULL args");
  switch (a->precision) {
    case PL_PREC_FP32: return pl::linear_bwd_simt_launch(a);
    case PL_PREC_BF16:
    case PL_PREC_TF32:
    case PL_PREC_BF16_FAST:
    case PL_PREC_TF32_FAST: return pl::linear_bwd_tc_launch(a);
    default:
      pl::set_error("pl_linear_bwd: precision %d not implemented", a->precision);
      return PL_ERR_UNSUPPORTED;
  }
}
