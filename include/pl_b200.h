/*
 * pl_b200.h -- C ABI of the B200-native MC-parallel variational layers.
 *
 * Drop-in boundary for the hot path of ludwigwinkler/pytorch_ProbabilisticLayers: the reference
 * has no FFI of its own (pure PyTorch), so each entry point below replaces the torch call sites
 * cited next to it (paths relative to the reference checkout).  All pointers are DEVICE pointers
 * owned by the caller (PyTorch's caching allocator in the Python host layer), all work is
 * enqueued asynchronously on `stream` (a cudaStream_t passed as void*), nothing is allocated or
 * synchronised inside the library, no C++ types or exceptions cross the boundary.
 *
 * Return value: PL_OK or a negative pl_status; the text of the last error of the calling thread
 * is available from pl_last_error().
 *
 * Tensor conventions (README.md:12 of the reference): dim 0 = Monte-Carlo sample, dim 1 = batch,
 * row-major fp32.  Linear weights are [in_features, out_features] (src/ProbabilisticLayers.py:858),
 * conv weights [Cout, Cin, k, k] (:993).
 *
 * eps: either injected (eps_w / eps_b non-NULL: the reference's own torch.randn draws, captured
 * at src/ProbabilisticLayers.py:835) or synthesised on chip from Philox4x32-10 with
 *     key = seed,  counter = (col/8, row, mc_global_offset + mc, 2*layer_id + is_bias)
 * (one call -> 8 normals: word j is a Box-Muller pair with a 20-bit radius and a 12-bit angle)
 * so that forward and backward regenerate the same numbers for any tiling and GPU count
 * (oracle/philox.py is the CPU restatement).
 */
#ifndef PL_B200_H_
#define PL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PL_B200_VERSION 100 /* major*10000 + minor*100 + patch */

typedef enum pl_status {
  PL_OK = 0,
  PL_ERR_BAD_ARG = -1,     /* NULL / negative / inconsistent argument                        */
  PL_ERR_UNSUPPORTED = -2, /* shape or mode not supported by the requested precision path    */
  PL_ERR_CUDA = -3,        /* a CUDA runtime/driver call failed; see pl_last_error()         */
  PL_ERR_WORKSPACE = -4    /* caller-provided workspace too small                            */
} pl_status;

typedef enum pl_precision {
  PL_PREC_FP32 = 0, /* fp32 FFMA kernels, any shape (exact-parity / tiny layers)              */
  /* Tensor-core modes (tcgen05, fp32 TMEM accumulators).  BF16 / TF32 keep the Monte-Carlo noise in its OWN MMA
   * operand: Y = X.round(mu) + X.round(sigma*eps), two MMAs per k-step into one accumulator (the mean tile by TMA,
   * the noise tile synthesised on chip).  The reference initialises sigma_w = sqrt(2/(I+O))/O
   * (src/ProbabilisticLayers.py:881), ~3e-4 of |mu|: a single operand round(mu + sigma*eps) is identical for all
   * MC samples until sigma has grown ~10x.  The *_FAST modes are that single rounded operand (half the MMAs):
   * valid once sigma >~ ulp(mu), opt-in. */
  PL_PREC_BF16 = 1,      /* kind::f16, bf16 operands, split mean / noise operands              */
  PL_PREC_TF32 = 2,      /* kind::tf32, fp32 inputs read as tf32, split mean / noise operands  */
  PL_PREC_BF16_FAST = 3, /* kind::f16, one operand bf16(mu + sigma*eps)                        */
  PL_PREC_TF32_FAST = 4  /* kind::tf32, one operand tf32(mu + sigma*eps)                       */
} pl_precision;

typedef enum pl_param_mode {
  PL_PARAM_SHARED_RHO = 0,  /* one (mu, rho) for all MC samples; sigma = softplus(rho);
                               gradients are reduced over MC (BayesLinear / BayesConv2d)      */
  PL_PARAM_PER_MC_SIGMA = 1 /* per-MC (mu, sigma) [MC, ...]; gradients keep the MC axis
                               (SIVILayer, src/ProbabilisticLayers.py:307-326)                 */
} pl_param_mode;

/* ---- library ---------------------------------------------------------------------------- */
int pl_version(void);
const char* pl_last_error(void);
/* Number of kernels this library has launched in this process (all threads); bench.py reports
 * the difference over its timed region as "gpu_launches". */
uint64_t pl_launch_count(void);
/* 1 if the tcgen05 path can run on the current device (compute capability 10.x), else 0. */
int pl_device_supported(void);
/* sizeof of the argument structs, for bindings to check their mirrors: 0 pl_linear_args, 1 pl_conv2d_args, 2 pl_ce_args,
 * 3 pl_bnpool_args, 4 pl_shard_seg, 5 pl_shard_ctx. */
size_t pl_struct_size(int32_t which);

/* ---- eps stream (test bridge; replaces torch.distributions.utils._standard_normal as used at
 *      src/ProbabilisticLayers.py:835) -----------------------------------------------------
 * out: float [num_mc, rows, cols] standard normals (raw == 0) or uint32 [num_mc, rows,
 * 4*ceil(cols/8)] raw Philox words (raw == 1) of stream `stream_id` (2*layer_id + is_bias). */
int pl_eps_dump(uint64_t seed, uint32_t stream_id, uint32_t mc_global_offset, int32_t num_mc,
                int32_t rows, int32_t cols, void* out, int32_t raw, void* stream);
/* The same with key = seed + *seed_dev (device-resident per-step offset, NULL = none): how SIVILayer draws its
 * hyper-network noise (torch.randn at src/ProbabilisticLayers.py:301-305) inside a captured CUDA graph. */
int pl_eps_dump_dev(uint64_t seed, const uint64_t* seed_dev, uint32_t stream_id, uint32_t mc_global_offset,
                    int32_t num_mc, int32_t rows, int32_t cols, void* out, int32_t raw, void* stream);

/* ---- KL(q||p) + entropy (replaces torch.distributions.kl_divergence(...).sum() and
 *      Normal.entropy().sum() at src/ProbabilisticLayers.py:955-956 and :1059) ---------------
 * out2[0] = sum 0.5[(s/sp)^2 + (mu/sp)^2 - 1 - log((s/sp)^2)],  out2[1] = sum 0.5+0.5log(2pi)+log s
 * prior_scale_elem (optional, per-element prior scale, n floats) overrides prior_scale.
 * workspace: pl_kl_workspace_bytes() bytes, ZERO-INITIALISED before its first use; every call
 * leaves it zeroed again, so one buffer can be reused by successive calls on the same stream.
 * Single launch, deterministic (fixed reduction tree). */
size_t pl_kl_workspace_bytes(void);
int pl_kl_fwd(const float* mu, const float* rho, int64_t n, float prior_scale,
              const float* prior_scale_elem, float* out2, void* workspace, size_t workspace_bytes,
              void* stream);
/* dmu (+)= g_kl * mu/sp^2 ; drho (+)= [g_kl*(s/sp^2 - 1/s) + g_ent/s] * sigmoid(rho), where
 * g_kl = host_scale * (*grad_kl or 1), g_ent = host_scale * (*grad_ent or 0) (device scalars,
 * may be NULL).  accumulate != 0 adds into dmu/drho. */
int pl_kl_bwd(const float* mu, const float* rho, int64_t n, float prior_scale,
              const float* prior_scale_elem, const float* grad_kl, const float* grad_ent,
              float host_scale, float* dmu, float* drho, int32_t accumulate, void* stream);

/* ---- fused optimiser pass (replaces torch.optim.Adam.step, experiments/BNN.py:349, plus the KL
 *      backward): Adam update of one parameter tensor with the analytic KL gradient added on the fly.
 * kind 0: plain Adam; 1: p is a posterior mean (grad += kl_scale * p/sp^2); 2: p is a posterior rho
 * (grad += kl_scale * (s/sp^2 - 1/s) sigmoid(p)).  kl_value (optional device double) += this tensor's
 * part of the KL VALUE from the pre-update parameters: kind 1: sum 0.5 p^2/sp^2, kind 2:
 * sum 0.5 s^2/sp^2 - log s; the caller adds n*(log sp - 0.5) per (mean, rho) pair.  step >= 1 is the
 * 1-based Adam step count.  Arithmetic identical to torch.optim.Adam (no amsgrad / weight decay). */
int pl_adam_kl_step(float* p, const float* g, float* m, float* v, int64_t n, int32_t kind,
                    float prior_scale, float kl_scale, float lr, float beta1, float beta2, float eps,
                    int32_t step, double* kl_value, void* stream);
/* The same pass with the step-dependent scalars read from device memory: sched_dev[0] = lr / (1 - beta1^t),
 * sched_dev[1] = 1 / sqrt(1 - beta2^t).  For CUDA-graph replays, where the caller advances t on the device. */
int pl_adam_kl_step_dev(float* p, const float* g, float* m, float* v, int64_t n, int32_t kind,
                        float prior_scale, float kl_scale, float beta1, float beta2, float eps,
                        const float* sched_dev, double* kl_value, void* stream);

/* ---- MC-first cross entropy + accuracy + gradient (replaces MC_CrossEntropyLoss.forward,
 *      src/ProbabilisticLayers.py:32-43, MC_Accuracy, :46-90, and the autograd backward of both) ----
 * pred [mc, batch, classes] fp32 logits, target [batch] int64 class ids shared by all MC samples.
 * out[0] = sum over (mc, batch) of -log softmax(pred)[target]   (the caller scales: the reference's value
 *          is out[0] / (mc*batch) * num_samples), out[1] = number of rows whose argmax equals the target
 *          (lowest index wins ties, as torch.argmax).
 * dpred (optional) [mc, batch, classes] = (softmax(pred) - onehot(target)) * grad_scale, i.e. the gradient of
 *          grad_scale * out[0]; rows with a target outside [0, classes) contribute no loss and no -onehot.
 * One launch, deterministic; workspace zero on first use and left zero (self-resetting ticket). */
typedef struct pl_ce_args {
  int32_t mc, batch, classes;
  float grad_scale;
  const float* pred;
  const int64_t* target;
  float* dpred;
  float* out;
  void* workspace;
  size_t workspace_bytes;
  void* stream;
} pl_ce_args;
size_t pl_mc_cross_entropy_workspace_bytes(void);
int pl_mc_cross_entropy(const pl_ce_args* a);

/* ---- MC_BatchNorm2D -> LeakyReLU -> MC_MaxPool2d(2, 2) fused (replaces, per conv block of the reference's
 *      'cbnn' net, experiments/BNN.py:232-247: MC_BatchNorm2D.forward, src/ProbabilisticLayers.py:736-762,
 *      LeakyReLU, MC_MaxPool2d.forward, :207-236, and their autograd) -----------------------------------
 * x: [n, channels, h, w] fp32 contiguous with n = MC*B (statistics are pooled over MC, B, H, W jointly, as
 * the reference's permute + F.batch_norm does), h and w even.
 * pl_bn_stats: sums[2*c] = sum x, sums[2*c+1] = sum x^2 over (n, hw) for each channel (fp64, zeroed by the call).
 * The caller derives mean / rstd = 1/sqrt(var + eps) (biased variance) and the running-statistics update.
 * fwd: y[n,c,h/2,w/2] = max over each 2x2 window of leaky_relu(gamma*(x-mean)*rstd + beta, slope);
 *      idx = which window element won (0..3, row-major, first maximum).
 * bwd: dx (optional), dgamma, dbeta (optional) from dy; batch_stats != 0: the statistics came from x itself
 *      (training: the mean(dz) and xhat*mean(dz*xhat) terms are included); 0: running statistics. */
typedef struct pl_bnpool_args {
  int32_t n, channels, h, w;
  float slope;
  int32_t batch_stats;
  const float* x;
  const float* gamma;
  const float* beta;
  const float* mean;
  const float* rstd;
  float* y;
  uint8_t* idx;
  const float* dy;
  float* dx;
  float* dgamma;
  float* dbeta;
  void* workspace;
  size_t workspace_bytes;
  void* stream;
} pl_bnpool_args;
int pl_bn_stats(const float* x, int32_t n, int32_t channels, int32_t hw, double* sums, void* stream);
size_t pl_bn_act_pool_workspace_bytes(int32_t channels);
int pl_bn_act_pool_fwd(const pl_bnpool_args* a);
int pl_bn_act_pool_bwd(const pl_bnpool_args* a);

/* ---- BayesLinear / SIVILayer contraction (replaces VariationalNormal.rsample + torch.baddbmm /
 *      torch.bmm at src/ProbabilisticLayers.py:925-935 and :321-326, and their autograd) ------ */
typedef struct pl_linear_args {
  int32_t mc, batch, in_features, out_features;
  int32_t precision;  /* pl_precision */
  int32_t param_mode; /* pl_param_mode */
  /* input x: [mc, batch, in]; x_mc_stride in elements, 0 => one [batch,in] matrix broadcast over
     MC (what MC_ExpansionLayer, src/ProbabilisticLayers.py:778-785, materialises by repeat) */
  const float* x;
  int64_t x_mc_stride;
  /* parameters: w_mu/w_rho [in,out] (+ leading MC axis in PER_MC_SIGMA mode, where w_rho and
     b_rho hold sigma itself); b_* may be NULL (bias=False) */
  const float* w_mu;
  const float* w_rho;
  const float* b_mu;
  const float* b_rho;
  /* sampler */
  uint64_t seed;
  uint32_t layer_id;
  uint32_t mc_global_offset; /* first global MC index owned by this rank (MC sharding) */
  const float* eps_w;        /* optional injected eps [mc,in,out] */
  const float* eps_b;        /* optional injected eps [mc,out]    */
  /* forward output [mc,batch,out] */
  float* y;
  /* backward: dy [mc,batch,out] in; outputs (each may be NULL to skip) */
  const float* dy;
  float* dx;     /* [mc,batch,in] */
  float* dw_mu;  /* [in,out]   (PER_MC_SIGMA: [mc,in,out]) */
  float* dw_rho; /* [in,out]   (PER_MC_SIGMA: d/dsigma, [mc,in,out]) */
  float* db_mu;  /* [out]      (PER_MC_SIGMA: [mc,out]) */
  float* db_rho; /* [out]      (PER_MC_SIGMA: [mc,out]) */
  /* scratch for the tensor-core path, see pl_linear_workspace_bytes */
  void* workspace;
  size_t workspace_bytes;
  void* stream;
  /* optional forward->backward state (pl_linear_state_bytes): pl_linear_fwd leaves
     sigma = softplus(rho) and the bf16 copy of x there; pl_linear_bwd called with the same
     buffer skips recomputing them (and then does not read x).  NULL = stateless. */
  void* fwd_state;
  size_t fwd_state_bytes;
  /* optional bf16 activation chaining between consecutive layers (PL_PREC_BF16 only): lets a caller
     that owns a whole BayesLinear -> LeakyReLU -> BayesLinear chain skip the fp32 round trips
     (y write, activation read/write, cast read/write) between the layers. */
  const void* x_lp;  /* fwd/bwd: x as bf16 [mc|1, batch, in]; replaces the cast pre-pass (x may be NULL) */
  void* y_lp;        /* fwd: also write act(y) as bf16 [mc,batch,out]; y may then be NULL */
  int32_t y_act;     /* 0: none, 1: LeakyReLU(y_act_slope) applied to y and y_lp */
  float y_act_slope;
  const void* dy_lp; /* bwd: dY as bf16 [mc,batch,out]; replaces the cast pre-pass (dy may be NULL) */
  void* dx_lp;       /* bwd: write dX * act'(x) as bf16 [mc,batch,in] (needs x_lp); dx may be NULL */
  float x_act_slope; /* slope of the LeakyReLU that produced x: act'(x) = x > 0 ? 1 : x_act_slope */
  const uint64_t* seed_dev; /* optional device pointer: the eps stream uses seed + *seed_dev (mod 2^64).  A captured
                               CUDA graph replays fixed launch parameters, so the per-step part of the seed lives in
                               device memory and is advanced inside the graph; NULL = seed as given. */
  /* MC-sharded training (pl_shard_*): the weight operand arrives ready-made from the owners' optimiser pass, and the
     rho gradient leaves un-scaled so that the owner applies sigmoid(rho) once after the cross-rank sum. */
  const void* w_lp;    /* tensor-core precisions: the weight in operand form (PL_PREC_BF16: mu16 [in*out] bf16 followed,
                          w_lp_half bytes later, by sig16; PL_PREC_BF16_FAST: packed {8 mu | 8 sigma} groups); replaces the
                          per-call parameter pre-pass, w_mu / w_rho are then not read by forward / dX.  NULL = pack here. */
  int64_t w_lp_half;
  int32_t dw_rho_raw;  /* != 0: dw_rho = sum_mc dW * eps (no sigmoid(rho) factor; w_rho is not read by the dW kernel) */
  /* Reduce-scatter fused into the dW epilogue (tensor-core precisions, >= 148 weight tiles): the tile of (dw_mu, raw dw_rho)
     is written straight into the staging rows  peers[owner] + *_off + (rank*chunk + j) * elem  of the ranks that own its
     elements (pl_shard_seg.g_off / g2_off / chunk of this weight), as fp32 or bf16 (dw_scatter_bf16, = ctx.wire_bf16).
     dw_mu / dw_rho may then be NULL and pl_shard_push skips this weight.  dw_scatter_world == 0: off. */
  void* const* dw_scatter_peers;   /* HOST array of dw_scatter_world arena base pointers */
  int64_t dw_scatter_mu_off, dw_scatter_rho_off, dw_scatter_chunk;
  int32_t dw_scatter_rank, dw_scatter_world, dw_scatter_bf16;
} pl_linear_args;

size_t pl_linear_workspace_bytes(const pl_linear_args* a);
size_t pl_linear_state_bytes(const pl_linear_args* a);
int pl_linear_fwd(const pl_linear_args* a);
int pl_linear_bwd(const pl_linear_args* a);

/* ---- BayesConv2d (replaces rsample + F.conv2d(groups=MC) at
 *      src/ProbabilisticLayers.py:1032-1057 and its autograd).  Square kernel, scalar
 *      stride/padding/dilation like the reference (:991).  x [mc,B,Cin,H,W] contiguous,
 *      y [mc,B,Cout,H',W'] contiguous.  PL_PREC_FP32: fp32 implicit GEMM; PL_PREC_BF16: tcgen05
 *      GEMMs over a bf16 im2col matrix (workspace: pl_conv2d_workspace_bytes, sized for the call's
 *      dy / dx pointers). ---------------------------------------------------------------------- */
typedef struct pl_conv2d_args {
  int32_t mc, batch, cin, cout, h, w, k, stride, padding, dilation;
  int32_t precision;
  const float* x;
  int64_t x_mc_stride; /* 0 => input broadcast over MC */
  const float* w_mu;   /* [cout,cin,k,k] */
  const float* w_rho;
  const float* b_mu;   /* [cout] */
  const float* b_rho;
  uint64_t seed;
  uint32_t layer_id;
  uint32_t mc_global_offset;
  const float* eps_w; /* optional [mc,cout,cin,k,k] */
  const float* eps_b; /* optional [mc,cout] */
  float* y;
  const float* dy;
  float* dx;
  float* dw_mu;
  float* dw_rho;
  float* db_mu;
  float* db_rho;
  void* workspace;
  size_t workspace_bytes;
  void* stream;
  /* optional forward->backward state (pl_conv2d_state_bytes, PL_PREC_BF16 only): pl_conv2d_fwd
     leaves the bf16 im2col matrix there and pl_conv2d_bwd given the same buffer does not rebuild it */
  void* fwd_state;
  size_t fwd_state_bytes;
  const uint64_t* seed_dev; /* as pl_linear_args.seed_dev */
} pl_conv2d_args;

size_t pl_conv2d_workspace_bytes(const pl_conv2d_args* a);
size_t pl_conv2d_state_bytes(const pl_conv2d_args* a);
int pl_conv2d_fwd(const pl_conv2d_args* a);
int pl_conv2d_bwd(const pl_conv2d_args* a);

/* ---- MC-sharded train step over NVLink peer memory (replaces, for the N-GPU step, Lightning DDP's gradient
 *      all-reduce + the replicated torch.optim.Adam.step: experiments/BNN.py:349, :399-400) --------------------------
 * Every rank owns an ARENA of identical layout (one device allocation, exported with pl_ipc_export and mapped by the
 * peers with pl_ipc_open): fp32 parameters | gradient staging | low-precision operand area | statistics | flags.
 * Rank r owns elements [r*chunk, (r+1)*chunk) of every parameter tensor (chunk a multiple of 8).  All *_off fields are
 * BYTE offsets from the arena base and are the same on every rank.
 *   pl_shard_push    scatters the local gradients (a flat fp32 buffer laid out like the arena's parameter area) into the
 *                    owners' staging:  peers[owner].staging[tensor][this rank][...]  (+ the 4 local statistics)
 *   pl_shard_barrier cross-GPU barrier (device-side epoch; graph-replayable; times out after ~4 s into the status word)
 *   pl_shard_adam    owner-side: sum of the staged partials in rank order (deterministic), sigmoid(rho) on raw drho sums,
 *                    analytic KL gradient, Adam on the shard (m, v hold only the shard), KL value, then the updated
 *                    parameters written into EVERY rank's arena in the form the forward consumes (gather below)
 *   pl_shard_finish  shares the KL partials, second barrier, stats_out4 = [LL, KL, ACC, LL+KL] summed over ranks
 * No NCCL call and no host synchronisation anywhere on this path. */
#define PL_SHARD_MAX_RANKS 8
#define PL_SHARD_MAX_SEGS 48

typedef struct pl_shard_seg {
  int32_t kind;      /* 0: plain Adam tensor; 1: (posterior mean, posterior rho) pair of one weight                   */
  int32_t gather;    /* 0: fp32 values into every rank's parameter area; 1: pair -> bf16 mu16 | sig16 (split operand,
                        PL_PREC_BF16) at lp_off / lp_off + lp_half; 2: pair -> packed {8 mu | 8 sigma} bf16 groups
                        (PL_PREC_BF16_FAST) at lp_off.  With 1 / 2 the fp32 values stay on the owner only.            */
  int32_t rho_raw;   /* pair: the staged rho gradient is sum_mc dW*eps; the owner multiplies by sigmoid(rho)          */
  float prior_scale; /* pair: scale of the N(0, scale) prior                                                          */
  int64_t numel;     /* elements of the tensor (of each tensor of a pair)                                             */
  int64_t chunk;     /* elements per rank, multiple of 8                                                              */
  int64_t p_off;     /* parameter (pair: mean) tensor in the arena's fp32 parameter area                              */
  int64_t p2_off;    /* pair: rho tensor                                                                              */
  int64_t g_off;     /* staging [world][chunk] floats of the tensor (pair: of the mean)                               */
  int64_t g2_off;    /* pair: staging of rho                                                                          */
  int64_t s_off;     /* ELEMENT offset into the shard's m / v state (pair: rho follows at s_off + chunk)              */
  int64_t lp_off;    /* gather 1 / 2: low-precision operand area of this weight                                       */
  int64_t lp_half;   /* gather 1: byte distance from mu16 to sig16                                                    */
} pl_shard_seg;

typedef struct pl_shard_ctx {
  int32_t rank, world, num_segs;
  int32_t wire_bf16;                 /* != 0: the partial gradients travel and are staged as bf16 (half the reduce-scatter
                                        bytes; the owner still sums them in fp32, in rank order)                         */
  void* peers[PL_SHARD_MAX_RANKS];   /* arena base of every rank as mapped in THIS process (peers[rank] = own arena)   */
  const pl_shard_seg* segs;          /* DEVICE pointer: num_segs entries                                               */
  void* mc_base;                     /* optional NVLS multicast mapping of the arenas (one store lands in every rank's
                                        arena, replicated by the NVSwitch): the all-gather then leaves each GPU once
                                        instead of once per peer; NULL = one store per peer                             */
  int64_t stats_off;                 /* float [2][8][4]: per-rank LL, -, ACC, - ; two slots alternating by step        */
  int64_t kl_off;                    /* double [2][8]: per-rank KL partial, two slots alternating by step              */
  int64_t flags_off;                 /* 4 channels x uint32 [16]: epoch, status, -, ..., then one arrival flag per rank  */
  int64_t push_units, adam_units;    /* totals of the two prefix arrays                                                */
  int64_t push_prefix[2 * PL_SHARD_MAX_SEGS + 1]; /* 8-element units of every tensor (2 slots per segment)            */
  int64_t adam_prefix[PL_SHARD_MAX_SEGS + 1];     /* 8-element units of THIS rank's chunk of every segment            */
} pl_shard_ctx;

/* CUDA IPC: handle (64 bytes) + byte offset of `ptr` inside its device allocation; pl_ipc_open maps a peer's allocation
 * into this process (peer access enabled lazily) and returns the pointer corresponding to `ptr`. */
int pl_ipc_export(const void* ptr, void* handle64, int64_t* offset_out);
int pl_ipc_open(const void* handle64, int64_t offset, void** ptr_out);
int pl_ipc_close(void* ptr, int64_t offset);

int pl_shard_push(const pl_shard_ctx* c, const float* grads, const float* stats4, void* stream);
/* The same for the 8-element units [unit_begin, unit_end) of push_prefix only (one layer's tensors, so that its transfer
 * can run on a side stream under the backward of the layers before it); stats4 may be NULL. */
int pl_shard_push_range(const pl_shard_ctx* c, const float* grads, const float* stats4, int64_t unit_begin,
                        int64_t unit_end, void* stream);
/* channel 0..3: independent flag sets, one per stream that issues barriers (their epochs must not interleave). */
int pl_shard_barrier(const pl_shard_ctx* c, int32_t channel, void* stream);
/* sched_dev (optional, device): [lr/(1-beta1^t), 1/sqrt(1-beta2^t)] for CUDA-graph replays, else `step` (1-based). */
int pl_shard_adam(const pl_shard_ctx* c, float* m, float* v, float kl_scale, float lr, float beta1, float beta2, float eps,
                  int32_t step, const float* sched_dev, double* kl_value, void* stream);
/* The same for segments [seg_begin, seg_end) only: a layer's optimiser pass + operand gather can then run on a side
 * stream (after its own push + barrier) under the backward of the layers before it.  kl_value accumulates. */
int pl_shard_adam_range(const pl_shard_ctx* c, int32_t seg_begin, int32_t seg_end, float* m, float* v, float kl_scale,
                        float lr, float beta1, float beta2, float eps, int32_t step, const float* sched_dev,
                        double* kl_value, void* stream);
int pl_shard_finish(const pl_shard_ctx* c, const double* kl_acc, double kl_const, double inv_num_samples,
                    float* stats_out4, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PL_B200_H_ */
