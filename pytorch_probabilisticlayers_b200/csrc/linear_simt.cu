// fp32 FFMA kernels for the BayesLinear / SIVILayer contraction (PL_PREC_FP32).
//
// Any shape, exact fp32 arithmetic: used for the launch-latency-bound configs (C1 regression MLP
// 1-49-51-52-1, C5 SIVI last layer 50->1), for shapes the tensor-core path does not tile, and as
// the on-device cross-check of the tcgen05 kernels.  Same fusion as the tensor-core path: the
// sampled weights W = mu + softplus(rho)*eps are synthesised per tile in shared memory from the
// Philox counter (or read from injected eps) and never written to HBM; backward regenerates eps.
//
// Replaces src/ProbabilisticLayers.py:925-935 (rsample + baddbmm/bmm), :321-326 (SIVI) and their
// autograd (closed forms in oracle/closed_form.py::linear_bwd).
#include "common.cuh"

namespace pl {

constexpr int kT = 64;         // output tile edge
constexpr int kK = 16;         // reduction chunk
constexpr int kThreads = 256;  // 16 x 16 threads, 4 x 4 outputs each
constexpr int kPad = 4;

struct LinearParams {
  int mc, batch, in, out;
  int per_mc;  // PL_PARAM_PER_MC_SIGMA
  int dw_mc_per_split, dw_atomic;  // dW kernel: MC samples per blockIdx.z, partial sums combined by atomics
  const float* x;
  int64_t x_mc_stride;
  const float* w_mu;
  const float* w_rho;
  const float* b_mu;
  const float* b_rho;
  const float* eps_w;
  const float* eps_b;
  Sampler sw, sb;
  float* y;
  const float* dy;
  float* dx;
  float* dw_mu;
  float* dw_rho;
  float* db_mu;
  float* db_rho;
  float* colsum;  // workspace [mc, out]
};

// 4 eps values for (mc, row, cols col..col+3), col % 4 == 0; injected or synthesised
__device__ __forceinline__ void load_eps4(const float* __restrict__ inj, const Sampler& s, int mc,
                                          int row, int col, int rows, int cols, float e[4]) {
  if (inj) {
    const float* p = inj + ((int64_t)mc * rows + row) * cols + col;
#pragma unroll
    for (int j = 0; j < 4; ++j) e[j] = (col + j < cols) ? __ldg(p + j) : 0.f;
  } else {
    const float4 n = eps4(s, mc, row, col >> 2);
    e[0] = n.x; e[1] = n.y; e[2] = n.z; e[3] = n.w;
  }
}

// sampled weight values W[mc][row][col..col+3]
__device__ __forceinline__ void sample_w4(const LinearParams& p, int mc, int row, int col, float w[4]) {
  float e[4];
  load_eps4(p.eps_w, p.sw, mc, row, col, p.in, p.out, e);
  const int64_t base = (p.per_mc ? (int64_t)mc * p.in * p.out : 0) + (int64_t)row * p.out + col;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (col + j < p.out) {
      const float m = __ldg(p.w_mu + base + j);
      const float r = __ldg(p.w_rho + base + j);
      const float sg = p.per_mc ? r : softplus_f(r);
      w[j] = fmaf(sg, e[j], m);
    } else {
      w[j] = 0.f;
    }
  }
}

__device__ __forceinline__ void fma_tile(const float (*As)[kT + kPad], const float (*Bs)[kT + kPad],
                                         int ty, int tx, float acc[4][4]) {
#pragma unroll
  for (int kk = 0; kk < kK; ++kk) {
    const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
    const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// ---------------------------------------------------------------------------------------------
// forward: Y[mc] = X[mc] . W[mc] + b[mc]      grid (ceil(O/64), ceil(B/64), MC)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) linear_fwd_simt(LinearParams p) {
  p.sw = live(p.sw);   // seed + device-side offset (CUDA-graph replays), see common.cuh
  p.sb = live(p.sb);
  __shared__ __align__(16) float As[kK][kT + kPad];  // [k][batch row]
  __shared__ __align__(16) float Bs[kK][kT + kPad];  // [k][out col]
  const int mc = blockIdx.z, b0 = blockIdx.y * kT, o0 = blockIdx.x * kT;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const float* xp = p.x + (int64_t)mc * p.x_mc_stride;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.in; k0 += kK) {
    {  // X tile: thread -> (row = t/4, 4 consecutive k)
      const int row = t >> 2, kq = (t & 3) * 4;
      const int b = b0 + row;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + kq + j;
        As[kq + j][row] = (b < p.batch && k < p.in) ? __ldg(xp + (int64_t)b * p.in + k) : 0.f;
      }
    }
    {  // W tile synthesised on chip: thread -> (k = t/16, one 4-group of out cols)
      const int kk = t >> 4, n = (t & 15) * 4;
      float w[4] = {0.f, 0.f, 0.f, 0.f};
      if (k0 + kk < p.in && o0 + n < p.out) sample_w4(p, mc, k0 + kk, o0 + n, w);
      *reinterpret_cast<float4*>(&Bs[kk][n]) = make_float4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
    fma_tile(As, Bs, ty, tx, acc);
    __syncthreads();
  }
  // epilogue: + sampled bias
  const int oc = o0 + tx * 4;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.b_mu && oc < p.out) {
    float e[4];
    load_eps4(p.eps_b, p.sb, mc, 0, oc, 1, p.out, e);
    const int64_t base = (p.per_mc ? (int64_t)mc * p.out : 0) + oc;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (oc + j < p.out) {
        const float r = __ldg(p.b_rho + base + j);
        bias[j] = fmaf(p.per_mc ? r : softplus_f(r), e[j], __ldg(p.b_mu + base + j));
      }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + ty * 4 + i;
    if (b >= p.batch) continue;
    float* yp = p.y + ((int64_t)mc * p.batch + b) * p.out + oc;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (oc + j < p.out) yp[j] = acc[i][j] + bias[j];
  }
}

// ---------------------------------------------------------------------------------------------
// dX[mc] = dY[mc] . W[mc]^T                   grid (ceil(I/64), ceil(B/64), MC)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) linear_dx_simt(LinearParams p) {
  p.sw = live(p.sw);   // seed + device-side offset (CUDA-graph replays), see common.cuh
  p.sb = live(p.sb);
  __shared__ __align__(16) float As[kK][kT + kPad];  // [o][batch row]
  __shared__ __align__(16) float Bs[kK][kT + kPad];  // [o][in col]
  const int mc = blockIdx.z, b0 = blockIdx.y * kT, i0 = blockIdx.x * kT;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const float* dyp = p.dy + (int64_t)mc * p.batch * p.out;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.out; k0 += kK) {
    {
      const int row = t >> 2, kq = (t & 3) * 4;
      const int b = b0 + row;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + kq + j;
        As[kq + j][row] = (b < p.batch && k < p.out) ? __ldg(dyp + (int64_t)b * p.out + k) : 0.f;
      }
    }
    {  // W^T tile: thread -> (in row = t/4, one 4-group of out cols inside the k chunk)
      const int il = t >> 2, g = (t & 3) * 4;
      float w[4] = {0.f, 0.f, 0.f, 0.f};
      if (i0 + il < p.in && k0 + g < p.out) sample_w4(p, mc, i0 + il, k0 + g, w);
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[g + j][il] = w[j];
    }
    __syncthreads();
    fma_tile(As, Bs, ty, tx, acc);
    __syncthreads();
  }
  const int ic = i0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int b = b0 + ty * 4 + i;
    if (b >= p.batch) continue;
    float* o = p.dx + ((int64_t)mc * p.batch + b) * p.in + ic;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (ic + j < p.in) o[j] = acc[i][j];
  }
}

// ---------------------------------------------------------------------------------------------
// dW: per (in-tile, out-tile) CTA, loop over MC; dW[mc] = X[mc]^T dY[mc] lives in registers,
// eps is regenerated from the counter, d_mu / d_rho are accumulated over MC on chip.
//                                               grid (ceil(O/64), ceil(I/64), mc_splits)
// Thin layers (e.g. 1200 -> 10: 19 tiles) leave most SMs idle, so blockIdx.z splits the MC range and the
// partial sums are combined with atomics into zero-initialised outputs (p.dw_atomic).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) linear_dw_simt(LinearParams p) {
  p.sw = live(p.sw);   // seed + device-side offset (CUDA-graph replays), see common.cuh
  p.sb = live(p.sb);
  __shared__ __align__(16) float As[kK][kT + kPad];  // [b][in row]
  __shared__ __align__(16) float Bs[kK][kT + kPad];  // [b][out col]
  const int i0 = blockIdx.y * kT, o0 = blockIdx.x * kT;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int kk = t >> 4, c4 = (t & 15) * 4;
  float acc_mu[4][4] = {}, acc_rho[4][4] = {};
  const int mc_begin = blockIdx.z * p.dw_mc_per_split, mc_end = min(p.mc, mc_begin + p.dw_mc_per_split);
  for (int mc = mc_begin; mc < mc_end; ++mc) {
    const float* xp = p.x + (int64_t)mc * p.x_mc_stride;
    const float* dyp = p.dy + (int64_t)mc * p.batch * p.out;
    float acc[4][4] = {};
    for (int b0 = 0; b0 < p.batch; b0 += kK) {
      const int b = b0 + kk;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + c4 + j, o = o0 + c4 + j;
        As[kk][c4 + j] = (b < p.batch && i < p.in) ? __ldg(xp + (int64_t)b * p.in + i) : 0.f;
        Bs[kk][c4 + j] = (b < p.batch && o < p.out) ? __ldg(dyp + (int64_t)b * p.out + o) : 0.f;
      }
      __syncthreads();
      fma_tile(As, Bs, ty, tx, acc);
      __syncthreads();
    }
    const int oc = o0 + tx * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = i0 + ty * 4 + i;
      if (row >= p.in || oc >= p.out) continue;
      float e[4];
      load_eps4(p.eps_w, p.sw, mc, row, oc, p.in, p.out, e);
      if (p.per_mc) {
        const int64_t base = ((int64_t)mc * p.in + row) * p.out + oc;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (oc + j < p.out) {
            p.dw_mu[base + j] = acc[i][j];
            p.dw_rho[base + j] = acc[i][j] * e[j];
          }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc_mu[i][j] += acc[i][j];
          acc_rho[i][j] = fmaf(acc[i][j], e[j], acc_rho[i][j]);
        }
      }
    }
  }
  if (p.per_mc) return;
  const int oc = o0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = i0 + ty * 4 + i;
    if (row >= p.in) continue;
    const int64_t base = (int64_t)row * p.out + oc;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (oc + j < p.out) {
        const float gr = acc_rho[i][j] * sigmoid_f(__ldg(p.w_rho + base + j));
        if (p.dw_atomic) {
          atomicAdd(p.dw_mu + base + j, acc_mu[i][j]);
          atomicAdd(p.dw_rho + base + j, gr);
        } else {
          p.dw_mu[base + j] = acc_mu[i][j];
          p.dw_rho[base + j] = gr;
        }
      }
  }
}

// ---------------------------------------------------------------------------------------------
// bias gradients: colsum[mc][o] = sum_b dY[mc][b][o], then a deterministic pass over MC
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bias_colsum_kernel(LinearParams p) {
  p.sw = live(p.sw);   // seed + device-side offset (CUDA-graph replays), see common.cuh
  p.sb = live(p.sb);
  __shared__ float sm[8][33];
  const int mc = blockIdx.y, o = blockIdx.x * 32 + threadIdx.x, ys = threadIdx.y;
  const float* dyp = p.dy + (int64_t)mc * p.batch * p.out;
  float s = 0.f;
  if (o < p.out)
    for (int b = ys; b < p.batch; b += 8) s += __ldg(dyp + (int64_t)b * p.out + o);
  sm[ys][threadIdx.x] = s;
  __syncthreads();
  if (ys == 0 && o < p.out) {
#pragma unroll
    for (int k = 1; k < 8; ++k) s += sm[k][threadIdx.x];
    if (p.per_mc) {
      float e[4];
      load_eps4(p.eps_b, p.sb, mc, 0, o & ~3, 1, p.out, e);
      p.db_mu[(int64_t)mc * p.out + o] = s;
      p.db_rho[(int64_t)mc * p.out + o] = s * e[o & 3];
    } else {
      p.colsum[(int64_t)mc * p.out + o] = s;
    }
  }
}

__global__ void __launch_bounds__(128) bias_grad_kernel(LinearParams p) {
  p.sw = live(p.sw);   // seed + device-side offset (CUDA-graph replays), see common.cuh
  p.sb = live(p.sb);
  const int o = blockIdx.x * 128 + threadIdx.x;
  if (o >= p.out) return;
  float am = 0.f, ar = 0.f;
  for (int mc = 0; mc < p.mc; ++mc) {
    const float s = p.colsum[(int64_t)mc * p.out + o];
    float e[4];
    load_eps4(p.eps_b, p.sb, mc, 0, o & ~3, 1, p.out, e);
    am += s;
    ar = fmaf(s, e[o & 3], ar);
  }
  p.db_mu[o] = am;
  p.db_rho[o] = ar * sigmoid_f(__ldg(p.b_rho + o));
}

static int fill_params(const pl_linear_args* a, LinearParams& p, const char* who) {
  PL_CHECK_ARG(a, "%s: NULL args", who);
  PL_CHECK_ARG(a->mc >= 0 && a->batch >= 0 && a->in_features >= 0 && a->out_features >= 0,
               "%s: negative dimension", who);
  const bool empty_x = (int64_t)a->mc * a->batch * a->in_features == 0;
  const bool empty_w = (int64_t)a->in_features * a->out_features == 0;
  PL_CHECK_ARG((a->x || empty_x) && ((a->w_mu && a->w_rho) || empty_w), "%s: NULL x / w_mu / w_rho", who);
  PL_CHECK_ARG((a->b_mu == nullptr) == (a->b_rho == nullptr), "%s: b_mu and b_rho must both be set or NULL", who);
  PL_CHECK_ARG(a->param_mode == PL_PARAM_SHARED_RHO || a->param_mode == PL_PARAM_PER_MC_SIGMA,
               "%s: bad param_mode %d", who, a->param_mode);
  PL_CHECK_ARG(a->x_mc_stride == 0 || a->x_mc_stride >= (int64_t)a->batch * a->in_features,
               "%s: x_mc_stride %lld overlaps samples", who, (long long)a->x_mc_stride);
  p.mc = a->mc; p.batch = a->batch; p.in = a->in_features; p.out = a->out_features;
  p.per_mc = a->param_mode == PL_PARAM_PER_MC_SIGMA;
  p.x = a->x; p.x_mc_stride = a->x_mc_stride;
  p.w_mu = a->w_mu; p.w_rho = a->w_rho; p.b_mu = a->b_mu; p.b_rho = a->b_rho;
  p.eps_w = a->eps_w; p.eps_b = a->eps_b;
  p.sw = make_sampler(a->seed, a->layer_id, false, a->mc_global_offset, a->seed_dev);
  p.sb = make_sampler(a->seed, a->layer_id, true, a->mc_global_offset, a->seed_dev);
  p.y = a->y; p.dy = a->dy; p.dx = a->dx;
  p.dw_mu = a->dw_mu; p.dw_rho = a->dw_rho; p.db_mu = a->db_mu; p.db_rho = a->db_rho;
  p.colsum = (float*)a->workspace;
  return PL_OK;
}

size_t linear_simt_workspace_bytes(const pl_linear_args* a) {
  return (size_t)a->mc * (size_t)a->out_features * sizeof(float);
}

int linear_fwd_simt_launch(const pl_linear_args* a) {
  LinearParams p;
  int rc = fill_params(a, p, "pl_linear_fwd");
  if (rc != PL_OK) return rc;
  if (p.mc == 0 || p.batch == 0 || p.out == 0) return PL_OK;
  PL_CHECK_ARG(a->y, "pl_linear_fwd: NULL y");
  PL_CHECK_ARG(p.mc <= 65535 && ceil_div64(p.batch, kT) <= 65535, "pl_linear_fwd: grid too large");
  dim3 grid((unsigned)ceil_div64(p.out, kT), (unsigned)ceil_div64(p.batch, kT), (unsigned)p.mc);
  linear_fwd_simt<<<grid, kThreads, 0, (cudaStream_t)a->stream>>>(p);
  PL_LAUNCH_CHECK("linear_fwd_simt");
  return PL_OK;
}

int linear_bwd_simt_launch(const pl_linear_args* a) {
  LinearParams p;
  int rc = fill_params(a, p, "pl_linear_bwd");
  if (rc != PL_OK) return rc;
  PL_CHECK_ARG(a->dy, "pl_linear_bwd: NULL dy");
  PL_CHECK_ARG((a->dw_mu == nullptr) == (a->dw_rho == nullptr), "pl_linear_bwd: dw_mu/dw_rho must both be set or NULL");
  PL_CHECK_ARG((a->db_mu == nullptr) == (a->db_rho == nullptr), "pl_linear_bwd: db_mu/db_rho must both be set or NULL");
  PL_CHECK_ARG(!a->db_mu || a->b_rho, "pl_linear_bwd: bias gradients requested without a bias");
  cudaStream_t st = (cudaStream_t)a->stream;
  if (p.mc == 0) {  // empty MC shard: gradients are zero
    if (a->dw_mu && !p.per_mc) {
      PL_CUDA(cudaMemsetAsync(a->dw_mu, 0, sizeof(float) * (size_t)p.in * p.out, st));
      PL_CUDA(cudaMemsetAsync(a->dw_rho, 0, sizeof(float) * (size_t)p.in * p.out, st));
    }
    if (a->db_mu && !p.per_mc) {
      PL_CUDA(cudaMemsetAsync(a->db_mu, 0, sizeof(float) * (size_t)p.out, st));
      PL_CUDA(cudaMemsetAsync(a->db_rho, 0, sizeof(float) * (size_t)p.out, st));
    }
    return PL_OK;
  }
  PL_CHECK_ARG(p.mc <= 65535 && ceil_div64(p.batch, kT) <= 65535 && ceil_div64(p.in, kT) <= 65535,
               "pl_linear_bwd: grid too large");
  if (a->dx && p.batch > 0 && p.in > 0) {
    dim3 grid((unsigned)ceil_div64(p.in, kT), (unsigned)ceil_div64(p.batch, kT), (unsigned)p.mc);
    linear_dx_simt<<<grid, kThreads, 0, st>>>(p);
    PL_LAUNCH_CHECK("linear_dx_simt");
  }
  if (a->dw_mu && p.in > 0 && p.out > 0) {
    const int64_t tiles = ceil_div64(p.out, kT) * ceil_div64(p.in, kT);
    int splits = 1;
    if (tiles < 148 && p.mc > 1) splits = (int)std::min<int64_t>(p.mc, (2 * 148 + tiles - 1) / tiles);
    p.dw_mc_per_split = (p.mc + splits - 1) / splits;
    splits = (p.mc + p.dw_mc_per_split - 1) / p.dw_mc_per_split;
    p.dw_atomic = (splits > 1 && !p.per_mc) ? 1 : 0;
    if (p.dw_atomic) {
      PL_CUDA(cudaMemsetAsync(a->dw_mu, 0, sizeof(float) * (size_t)p.in * p.out, st));
      PL_CUDA(cudaMemsetAsync(a->dw_rho, 0, sizeof(float) * (size_t)p.in * p.out, st));
    }
    dim3 grid((unsigned)ceil_div64(p.out, kT), (unsigned)ceil_div64(p.in, kT), (unsigned)splits);
    linear_dw_simt<<<grid, kThreads, 0, st>>>(p);
    PL_LAUNCH_CHECK("linear_dw_simt");
  }
  if (a->db_mu && p.out > 0) {
    if (!p.per_mc) {
      if (!a->workspace || a->workspace_bytes < linear_simt_workspace_bytes(a)) {
        set_error("pl_linear_bwd: workspace %zu < %zu bytes", a->workspace_bytes,
                  linear_simt_workspace_bytes(a));
        return PL_ERR_WORKSPACE;
      }
    }
    dim3 grid((unsigned)ceil_div64(p.out, 32), (unsigned)p.mc), block(32, 8);
    bias_colsum_kernel<<<grid, block, 0, st>>>(p);
    PL_LAUNCH_CHECK("bias_colsum_kernel");
    if (!p.per_mc) {
      bias_grad_kernel<<<(unsigned)ceil_div64(p.out, 128), 128, 0, st>>>(p);
      PL_LAUNCH_CHECK("bias_grad_kernel");
    }
  }
  return PL_OK;
}

}  // namespace pl
