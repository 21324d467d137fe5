// tcgen05 / TMEM / TMA path of the BayesLinear contraction (PL_PREC_BF16).
//
//   forward   Y[mc]  = X[mc]  . W[mc]   + b[mc]          (src/ProbabilisticLayers.py:925-935)
//   backward  dX[mc] = dY[mc] . W[mc]^T                  (autograd of the above)
//             dW[mc] = X[mc]^T . dY[mc];  dmu = sum_mc dW[mc];  drho = sigmoid(rho) sum_mc dW[mc]*eps[mc]
//
// with W[mc] = mu + softplus(rho) * eps[mc] synthesised ON CHIP: the weight operand of the MMA is
// written by producer warps straight into the swizzled shared-memory layout tcgen05.mma reads
// (Philox counter -> Box-Muller -> FMA -> bf16), so sampled weights never exist in HBM, and the
// backward regenerates eps from the same counter instead of loading it.
//
// sampled_gemm_tc (forward and dX share it):
//   CTA tile  = [MT x 128 batch rows] x [128 output columns] for one MC sample; the generated
//               W tile (64 x 128 per k-block) is reused by all MT row sub-tiles, i.e. by the whole
//               batch (B = 512 -> MT = 4, 4 x 128 fp32 accumulator columns = the full 512-column TMEM)
//   warp 0    = TMA producer: activation tile (bf16, K-major, SWIZZLE_128B) -> smem ring A
//   warp 1    = TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, fp32 accumulate)
//   warps 2-9 = W-tile generators (ring B) during the K loop, then the TMEM -> HBM epilogue
//   rings are mbarrier pipelines; tcgen05.commit frees stages; ring depths are decoupled because
//   the activation stage is 4x larger than the generated stage.
//
// Roofline: tensor pipe (2*MC*B*I*O algorithmic FLOPs per launch); the secondary limit is the
// sampler's ALU/MUFU budget: every generated element feeds only 2*B flops.
#include "tc_kernels.cuh"

#ifndef PL_DW_2CTA_DEFAULT
#define PL_DW_2CTA_DEFAULT 0
#endif

namespace pl {

// forward->backward state: parameters as the kernels read them | x as bf16 [x_rows*I]
//   PL_PREC_BF16 (split operand):  mu16 [I][O] bf16 (TMA) | sig16 [I][O] bf16 (generators)
//   PL_PREC_TF32 (split operand):  mu_hi [I*O] f32 tf32-rounded (TMA) | sigma [I*O] f32
//   PL_PREC_BF16_FAST: packed {8 mu | 8 sigma} bf16 groups;  PL_PREC_TF32_FAST: sigma [I*O] f32
struct TcState {
  float* sigma;      // fast: the packed / sigma buffer; split: the sigma half
  float* mean;       // split: the rounded-mean matrix the TMA reads, else NULL
  __nv_bfloat16* x_bf16;
  size_t total;
};
static inline bool prec_tf32(int p) { return p == PL_PREC_TF32 || p == PL_PREC_TF32_FAST; }
static inline bool prec_split(int p) { return p == PL_PREC_BF16 || p == PL_PREC_TF32; }
// per-call scratch: [state, when the caller keeps none] | sampled bias [mc*O] | colsum [mc*O] |
// dy as bf16 [mc*B*O]
struct TcScratch {
  TcState st;
  float* bias;
  float* colsum;
  float* mean_mat;            // [B, O] fp32: X.round(mu) of an MC-broadcast input (split modes, forward)
  __nv_bfloat16* dy_bf16;
  size_t total;
};
// the MC-shared mean part is hoisted out of the per-sample launches when the input has no MC axis
static inline bool hoist_mean(const pl_linear_args* a) {
  return prec_split(a->precision) && a->x_mc_stride == 0 && a->mc >= 4 && !a->eps_w;
}

static TcState carve_state(const pl_linear_args* a, void* base) {
  const size_t I = a->in_features, O = a->out_features, MC = a->mc, B = a->batch;
  const size_t x_rows = a->x_mc_stride == 0 ? B : MC * B;
  uint8_t* b = (uint8_t*)(((uintptr_t)base + 255) & ~(uintptr_t)255);
  TcState s;
  size_t off = 0;
  s.mean = nullptr;
  if (prec_split(a->precision)) {
    const size_t half = align256(prec_tf32(a->precision) ? I * O * 4 : split16_bytes(I, O));
    s.mean = (float*)(b + off); off += half;
    s.sigma = (float*)(b + off); off += half;
  } else {
    s.sigma = (float*)(b + off); off += align256(std::max(I * O * 4, pack_bytes(I, O)));
  }
  s.x_bf16 = (__nv_bfloat16*)(b + off); off += align256(x_rows * I * 2);
  s.total = off + 256;
  return s;
}

static TcScratch carve_scratch(const pl_linear_args* a, void* base, bool with_state, bool backward) {
  const size_t O = a->out_features, MC = a->mc, B = a->batch;
  uint8_t* b = (uint8_t*)(((uintptr_t)base + 255) & ~(uintptr_t)255);
  TcScratch w;
  size_t off = 0;
  if (with_state) {
    w.st = carve_state(a, b);
    off += w.st.total;
  } else {
    w.st = TcState{nullptr, nullptr, nullptr, 0};
  }
  w.bias = (float*)(b + off); off += align256(MC * O * 4);
  w.colsum = (float*)(b + off); off += align256(MC * O * 4);
  w.mean_mat = (float*)(b + off);
  if (!backward && hoist_mean(a)) off += align256(B * O * 4);
  w.dy_bf16 = (__nv_bfloat16*)(b + off);
  if (backward) off += align256(MC * B * O * 2);
  w.total = off + 256;
  return w;
}

size_t linear_tc_state_bytes(const pl_linear_args* a) { return carve_state(a, nullptr).total; }

size_t linear_tc_workspace_bytes(const pl_linear_args* a) {
  if (!a) return 0;
  const bool backward = a->dy != nullptr;
  const bool need_state = a->fwd_state == nullptr;
  return carve_scratch(a, nullptr, need_state, backward).total;
}

static int tc_shape_ok(const pl_linear_args* a, const char* who) {
  if (a->param_mode != PL_PARAM_SHARED_RHO) {
    set_error("%s: the tensor-core precisions support PL_PARAM_SHARED_RHO only", who);
    return PL_ERR_UNSUPPORTED;
  }
  if (a->in_features % 8 || a->out_features % 8 || a->in_features < 8 || a->out_features < 8) {
    set_error("%s: PL_PREC_BF16 needs in/out features to be multiples of 8 (got %d, %d)", who,
              a->in_features, a->out_features);
    return PL_ERR_UNSUPPORTED;
  }
  if (a->x_mc_stride != 0 && a->x_mc_stride != (int64_t)a->batch * a->in_features) {
    set_error("%s: PL_PREC_BF16 needs a dense [mc,batch,in] input (or mc stride 0)", who);
    return PL_ERR_UNSUPPORTED;
  }
  return PL_OK;
}

int linear_bwd_simt_launch(const pl_linear_args* a);

// the caller hands the weight in operand form (pl_linear_args.w_lp): point the state's parameter views at it
static bool use_prepacked(const pl_linear_args* a, TcState& s) {
  if (!a->w_lp || prec_tf32(a->precision)) return false;
  if (prec_split(a->precision)) {
    s.mean = (float*)a->w_lp;
    s.sigma = (float*)((uint8_t*)a->w_lp + a->w_lp_half);
  } else {
    s.mean = nullptr;
    s.sigma = (float*)a->w_lp;
  }
  return true;
}

// picks the sampled_gemm_tc instantiation for (injected eps, tf32, split operand); `mean` = the rounded-mean matrix
// [w_rows][pitch] of the split kernels
template <bool kDx>
static int dispatch_gemm(const CUtensorMap& tm, const TcGemmParams& p, cudaStream_t st, const void* mean,
                         int64_t w_rows, int64_t w_cols, int64_t pitch, bool inj, bool tf32, bool split) {
  if (split) {
    CUtensorMap tmb;
    int rc = make_tmap_mean(&tmb, mean, (uint64_t)w_rows, (uint64_t)w_cols, (uint64_t)pitch, kDx, tf32);
    if (rc != PL_OK) return rc;
    if (tf32) return inj ? launch_gemm<kDx, true, true, true>(tm, p, st, &tmb) : launch_gemm<kDx, false, true, true>(tm, p, st, &tmb);
    return inj ? launch_gemm<kDx, true, false, true>(tm, p, st, &tmb) : launch_gemm<kDx, false, false, true>(tm, p, st, &tmb);
  }
  if (tf32) return inj ? launch_gemm<kDx, true, true>(tm, p, st) : launch_gemm<kDx, false, true>(tm, p, st);
  return inj ? launch_gemm<kDx, true>(tm, p, st) : launch_gemm<kDx, false>(tm, p, st);
}

static int check_buffers(const pl_linear_args* a, const char* who) {
  if (!a->workspace || a->workspace_bytes < linear_tc_workspace_bytes(a)) {
    set_error("%s: workspace %zu < %zu bytes", who, a->workspace_bytes, linear_tc_workspace_bytes(a));
    return PL_ERR_WORKSPACE;
  }
  if (a->fwd_state && a->fwd_state_bytes < linear_tc_state_bytes(a)) {
    set_error("%s: fwd_state %zu < %zu bytes", who, a->fwd_state_bytes, linear_tc_state_bytes(a));
    return PL_ERR_WORKSPACE;
  }
  return PL_OK;
}

int linear_fwd_tc_launch(const pl_linear_args* a) {
  PL_CHECK_ARG(a, "pl_linear_fwd: NULL args");
  if (a->mc == 0 || a->batch == 0) return PL_OK;
  PL_CHECK_ARG((a->x || a->x_lp) && a->w_mu && a->w_rho && (a->y || a->y_lp), "pl_linear_fwd: NULL pointer");
  PL_CHECK_ARG((a->b_mu == nullptr) == (a->b_rho == nullptr), "pl_linear_fwd: b_mu and b_rho must both be set or NULL");
  int rc = tc_shape_ok(a, "pl_linear_fwd");
  if (rc != PL_OK) return rc;
  PL_CHECK_ARG(a->mc <= 65535, "pl_linear_fwd: mc too large");
  pl_linear_args fa = *a;
  fa.dy = nullptr;  // sizing: forward scratch
  rc = check_buffers(&fa, "pl_linear_fwd");
  if (rc != PL_OK) return rc;
  cudaStream_t st = (cudaStream_t)a->stream;
  TcScratch w = carve_scratch(&fa, a->workspace, a->fwd_state == nullptr, false);
  TcState state = a->fwd_state ? carve_state(a, a->fwd_state) : w.st;
  const int64_t I = a->in_features, O = a->out_features, MC = a->mc, B = a->batch;
  const int64_t x_rows = a->x_mc_stride == 0 ? B : MC * B;
  const bool tf32 = prec_tf32(a->precision);   // fp32 operands straight from the caller's tensors
  const bool split = prec_split(a->precision);
  if (!use_prepacked(a, state)) {
    rc = split ? prep_params_split(a->w_mu, a->w_rho, state.mean, state.sigma, I, O, tf32, st)
               : prep_params(a->w_mu, a->w_rho, state.sigma, I, O, tf32, st);
    if (rc != PL_OK) return rc;
  }
  PL_CHECK_ARG(!tf32 || !(a->x_lp || a->y_lp), "pl_linear_fwd: bf16 activation chaining needs PL_PREC_BF16");
  if (!tf32 && !a->x_lp) {
    cast_bf16_kernel<<<ew_grid(x_rows * I / 8, 1024), 256, 0, st>>>(a->x, state.x_bf16, x_rows * I / 8);
    PL_LAUNCH_CHECK("cast_bf16_kernel");
  }
  if (a->b_mu) {
    sample_bias_kernel<<<ew_grid(MC * ((O + 3) / 4), 256), 256, 0, st>>>(
        a->b_mu, a->b_rho, a->eps_b, make_sampler(a->seed, a->layer_id, true, a->mc_global_offset, a->seed_dev),
        (int)MC, (int)O, w.bias);
    PL_LAUNCH_CHECK("sample_bias_kernel");
  }
  CUtensorMap tm;
  rc = tf32 ? make_tmap_2d(&tm, a->x, (uint64_t)x_rows, (uint64_t)I, 128, 32, true)
            : make_tmap_bf16(&tm, a->x_lp ? a->x_lp : (const void*)state.x_bf16, (uint64_t)x_rows,
                             (uint64_t)I, 128, kBK);
  if (rc != PL_OK) return rc;
  TcGemmParams p{};
  p.mc = (int)MC; p.rows = (int)B; p.ndim = (int)O; p.kdim = (int)I;
  p.w_in = (int)I; p.w_out = (int)O;
  p.a_mc_rows = a->x_mc_stride == 0 ? 0 : B;
  p.w_mu = a->w_mu; p.w_sigma = state.sigma; p.w_pack = (const uint4*)state.sigma; p.eps_w = a->eps_w;
  p.w_sig16 = (const uint4*)state.sigma;
  p.sw = make_sampler(a->seed, a->layer_id, false, a->mc_global_offset, a->seed_dev);
  p.bias = a->b_mu ? w.bias : nullptr;
  p.out = a->y;
  p.out_hw = 0;
  p.w_aligned = 1;
  p.out_lp = (__nv_bfloat16*)a->y_lp;
  p.act = a->y_act;
  p.act_slope = a->y_act_slope;
  if (hoist_mean(a)) {
    // X.round(mu) once (one "sample", no bias / activation), then the per-sample launches add it in their epilogue
    TcGemmParams pm = p;
    pm.mc = 1; pm.bias = nullptr; pm.out = w.mean_mat; pm.out_lp = nullptr; pm.act = 0; pm.mean_only = 1;
    rc = dispatch_gemm<false>(tm, pm, st, state.mean, I, O, O, false, tf32, true);
    if (rc != PL_OK) return rc;
    p.noise_only = 1;
    p.mean_mat = w.mean_mat;
  }
  return dispatch_gemm<false>(tm, p, st, state.mean, I, O, O, a->eps_w != nullptr, tf32, split);
}

int linear_bwd_tc_launch(const pl_linear_args* a) {
  PL_CHECK_ARG(a, "pl_linear_bwd: NULL args");
  if (a->mc == 0 || a->batch == 0) return linear_bwd_simt_launch(a);
  const bool have_state = a->fwd_state != nullptr;
  const bool tf32 = prec_tf32(a->precision);
  const bool split = prec_split(a->precision);
  PL_CHECK_ARG((a->x || a->x_lp || (have_state && !tf32)) && a->w_mu && a->w_rho && (a->dy || a->dy_lp),
               "pl_linear_bwd: NULL pointer");
  PL_CHECK_ARG(!tf32 || !(a->x_lp || a->dy_lp || a->dx_lp), "pl_linear_bwd: bf16 activation chaining needs PL_PREC_BF16");
  PL_CHECK_ARG(!a->dx_lp || a->x_lp, "pl_linear_bwd: dx_lp needs x_lp (the activation mask source)");
  PL_CHECK_ARG((a->db_mu == nullptr) == (a->db_rho == nullptr), "pl_linear_bwd: db_mu/db_rho must both be set or NULL");
  PL_CHECK_ARG(!a->db_mu || a->b_rho, "pl_linear_bwd: bias gradients requested without a bias");
  int rc = tc_shape_ok(a, "pl_linear_bwd");
  if (rc != PL_OK) return rc;
  rc = check_buffers(a, "pl_linear_bwd");
  if (rc != PL_OK) return rc;
  cudaStream_t st = (cudaStream_t)a->stream;
  TcScratch w = carve_scratch(a, a->workspace, !have_state, true);
  TcState state = have_state ? carve_state(a, a->fwd_state) : w.st;
  const int64_t I = a->in_features, O = a->out_features, MC = a->mc, B = a->batch;
  const int64_t x_rows = a->x_mc_stride == 0 ? B : MC * B;
  const bool prepacked = use_prepacked(a, state);
  if (!have_state) {
    if ((a->dx || a->dx_lp) && !prepacked) {
      rc = split ? prep_params_split(a->w_mu, a->w_rho, state.mean, state.sigma, I, O, tf32, st)
                 : prep_params(a->w_mu, a->w_rho, state.sigma, I, O, tf32, st);
      if (rc != PL_OK) return rc;
    }
    if (a->dw_mu && !tf32 && !a->x_lp) {
      cast_bf16_kernel<<<ew_grid(x_rows * I / 8, 1024), 256, 0, st>>>(a->x, state.x_bf16, x_rows * I / 8);
      PL_LAUNCH_CHECK("cast_bf16_kernel");
    }
  }
  const bool want_dx = a->dx || a->dx_lp;
  const __nv_bfloat16* dy_bf16 = a->dy_lp ? (const __nv_bfloat16*)a->dy_lp : w.dy_bf16;
  if (a->dy_lp) {
    if (a->db_mu) {   // bias column sums straight from the bf16 dY
      dim3 grid((unsigned)((O + 127) / 128), (unsigned)MC);
      colsum_bf16_kernel<<<grid, 256, 0, st>>>(dy_bf16, w.colsum, (int)B, (int)O);
      PL_LAUNCH_CHECK("colsum_bf16_kernel");
    }
  } else if ((!tf32 && (want_dx || a->dw_mu)) || a->db_mu) {
    // one pass over dY: bf16 operand copy + per-sample column sums for the bias gradients
    dim3 grid((unsigned)((O + 127) / 128), (unsigned)MC);
    cast_dy_colsum_kernel<<<grid, 256, 0, st>>>(a->dy, tf32 ? nullptr : w.dy_bf16,
                                                a->db_mu ? w.colsum : nullptr, (int)B, (int)O);
    PL_LAUNCH_CHECK("cast_dy_colsum_kernel");
  }
  if (a->db_mu) {
    bias_grad_tc_kernel<<<(unsigned)((O + 127) / 128), 128, 0, st>>>(
        w.colsum, a->b_rho, a->eps_b, make_sampler(a->seed, a->layer_id, true, a->mc_global_offset, a->seed_dev),
        (int)MC, (int)O, a->db_mu, a->db_rho);
    PL_LAUNCH_CHECK("bias_grad_tc_kernel");
  }
  if (want_dx) {
    CUtensorMap tm;
    rc = tf32 ? make_tmap_2d(&tm, a->dy, (uint64_t)(MC * B), (uint64_t)O, 128, 32, true)
              : make_tmap_bf16(&tm, dy_bf16, (uint64_t)(MC * B), (uint64_t)O, 128, kBK);
    if (rc != PL_OK) return rc;
    TcGemmParams p{};
    p.mc = (int)MC; p.rows = (int)B; p.ndim = (int)I; p.kdim = (int)O;
    p.w_in = (int)I; p.w_out = (int)O;
    p.a_mc_rows = B;
    p.w_mu = a->w_mu; p.w_sigma = state.sigma; p.w_pack = (const uint4*)state.sigma; p.eps_w = a->eps_w;
    p.w_sig16 = (const uint4*)state.sigma;
    p.sw = make_sampler(a->seed, a->layer_id, false, a->mc_global_offset, a->seed_dev);
    p.bias = nullptr;
    p.out = a->dx;
    p.out_hw = 0;
    p.w_aligned = 1;
    p.out_lp = (__nv_bfloat16*)a->dx_lp;
    if (a->dx_lp) {   // chain: dX * act'(x) for the previous layer, mask from this layer's bf16 input
      p.mask_src = (const __nv_bfloat16*)a->x_lp;
      p.mask_slope = a->x_act_slope;
    }
    rc = dispatch_gemm<true>(tm, p, st, state.mean, I, O, O, a->eps_w != nullptr, tf32, split);
    if (rc != PL_OK) return rc;
  }
  if (a->dw_mu || a->dw_scatter_world > 0) {
    PL_CHECK_ARG(a->dw_scatter_world > 0 || a->dw_rho, "pl_linear_bwd: dw_mu/dw_rho must both be set or NULL");
    CUtensorMap tmx, tmdy;
    const uint64_t mcx = a->x_mc_stride == 0 ? 1 : (uint64_t)MC;
    if (tf32) {
      rc = make_tmap_3d(&tmx, a->x, mcx, (uint64_t)B, (uint64_t)I, 32, 32, true);
      if (rc != PL_OK) return rc;
      rc = make_tmap_3d(&tmdy, a->dy, (uint64_t)MC, (uint64_t)B, (uint64_t)O, 32, 32, true);
    } else {
      rc = make_tmap_bf16_3d(&tmx, a->x_lp ? a->x_lp : (const void*)state.x_bf16, mcx, (uint64_t)B,
                             (uint64_t)I, kBK, 64);
      if (rc != PL_OK) return rc;
      rc = make_tmap_bf16_3d(&tmdy, dy_bf16, (uint64_t)MC, (uint64_t)B, (uint64_t)O, kBK, 64);
    }
    if (rc != PL_OK) return rc;
    TcDwParams p{};
    p.mc = (int)MC; p.batch = (int)B; p.w_in = (int)I; p.w_out = (int)O;
    p.x_broadcast = a->x_mc_stride == 0;
    p.dy_broadcast = 0;
    p.w_aligned = 1;
    p.w_rho = a->dw_rho_raw ? nullptr : a->w_rho; p.eps_w = a->eps_w;
    p.sw = make_sampler(a->seed, a->layer_id, false, a->mc_global_offset, a->seed_dev);
    p.dw_mu = a->dw_mu; p.dw_rho = a->dw_rho;
    const int tiles = (int)(((I + 127) / 128) * ((O + 127) / 128));
    int splits = 1;
    if (tiles < 148) {  // too few weight tiles to fill the GPU: split MC, combine with atomics
      splits = (2 * 148) / tiles;
      if (splits > MC) splits = (int)MC;
      if (splits < 1) splits = 1;
    }
    p.mc_per_split = (int)((MC + splits - 1) / splits);
    p.kb_splits = 1;
    p.kb_per_split = (int)((B + (tf32 ? 32 : kBK) - 1) / (tf32 ? 32 : kBK));
    splits = (int)((MC + p.mc_per_split - 1) / p.mc_per_split);
    p.atomic_out = splits > 1;
    if (a->dw_scatter_world > 0) {
      // reduce-scatter fused into the epilogue (pl_linear_args.dw_scatter_*): needs the one-CTA-per-tile path
      PL_CHECK_ARG(!p.atomic_out, "pl_linear_bwd: dw_scatter needs >= 148 weight tiles (no MC split)");
      PL_CHECK_ARG(a->dw_scatter_world <= 8 && a->dw_scatter_peers && a->dw_scatter_chunk > 0 && a->dw_scatter_chunk % 8 == 0,
                   "pl_linear_bwd: bad dw_scatter arguments");
      for (int r = 0; r < a->dw_scatter_world; ++r) p.sc_peers[r] = a->dw_scatter_peers[r];
      p.sc_mu_off = a->dw_scatter_mu_off; p.sc_rho_off = a->dw_scatter_rho_off; p.sc_chunk = a->dw_scatter_chunk;
      p.sc_rank = a->dw_scatter_rank; p.sc_world = a->dw_scatter_world; p.sc_bf16 = a->dw_scatter_bf16;
      p.w_rho = nullptr;           // the owners apply sigmoid(rho) after the cross-rank sum
    }
    if (p.atomic_out) {
      PL_CUDA(cudaMemsetAsync(a->dw_mu, 0, sizeof(float) * I * O, st));
      PL_CUDA(cudaMemsetAsync(a->dw_rho, 0, sizeof(float) * I * O, st));
    }
    // CTA pairs on a 256 x 128 weight tile (dw_tc2): 25 % fewer operand bytes per FLOP into each SM.  PL_DW_2CTA=0 / 1.
    static int dw_pairs = -1;
    if (dw_pairs < 0) {
      const char* e = getenv("PL_DW_2CTA");
      dw_pairs = e ? (e[0] == '1') : PL_DW_2CTA_DEFAULT;
    }
    if (dw_pairs && !tf32 && I >= 256 && O >= 128 && O % 8 == 0 && a->dw_scatter_world == 0) {
      dim3 grid2(2u * (unsigned)((I + 255) / 256), (unsigned)((O + 127) / 128), (unsigned)splits);
      if (a->eps_w) {
        PL_CUDA(cudaFuncSetAttribute(dw_tc2<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDw2SmemBytes));
        dw_tc2<true><<<grid2, kTcThreads, kDw2SmemBytes, st>>>(tmx, tmdy, p);
      } else {
        PL_CUDA(cudaFuncSetAttribute(dw_tc2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDw2SmemBytes));
        dw_tc2<false><<<grid2, kTcThreads, kDw2SmemBytes, st>>>(tmx, tmdy, p);
      }
      PL_LAUNCH_CHECK("dw_tc2");
      return PL_OK;
    }
    dim3 grid((unsigned)((O + 127) / 128), (unsigned)((I + 127) / 128), (unsigned)splits);
#define PL_DW_LAUNCH(INJ, TF)                                                                          \
  do {                                                                                                \
    PL_CUDA(cudaFuncSetAttribute(dw_tc<INJ, TF>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemBytes)); \
    dw_tc<INJ, TF><<<grid, kTcThreads, kDwSmemBytes, st>>>(tmx, tmdy, p);                             \
  } while (0)
    if (a->eps_w) {
      if (tf32) PL_DW_LAUNCH(true, true);
      else PL_DW_LAUNCH(true, false);
    } else {
      if (tf32) PL_DW_LAUNCH(false, true);
      else PL_DW_LAUNCH(false, false);
    }
#undef PL_DW_LAUNCH
    PL_LAUNCH_CHECK("dw_tc");
  }
  return PL_OK;
}

}  // namespace pl
