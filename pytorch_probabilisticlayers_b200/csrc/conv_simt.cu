// fp32 implicit-GEMM kernels for BayesConv2d (PL_PREC_FP32).
//
// Per MC sample the convolution is a GEMM with M = B*H'*W' (output pixels), N = Cout,
// K = Cin*k*k; the im2col matrix is gathered on the fly, the sampled filter
// W[mc] = mu + softplus(rho)*eps[mc] is synthesised per tile from the Philox counter (row = cout,
// col = cin*k*k + kh*k + kw) and never written to HBM.  Input is read in the reference's logical
// [MC,B,Cin,H,W] layout and the output is written contiguous [MC,B,Cout,H',W'], which removes the
// two full permute copies of src/ProbabilisticLayers.py:1032-1033 and :1056-1057.
//
// Replaces rsample + F.conv2d(groups=MC) (src/ProbabilisticLayers.py:1038-1051) and its autograd
// (closed forms in oracle/closed_form.py::conv2d_bwd).
#include "common.cuh"

namespace pl {

constexpr int kCT = 64;
constexpr int kCK = 16;
constexpr int kCThreads = 256;
constexpr int kCPad = 4;

struct ConvParams {
  int mc, batch, cin, cout, h, w, k, stride, pad, dil, ho, wo;
  int kdim;  // cin*k*k
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
  float* colsum;
};

__device__ __forceinline__ void conv_eps4(const float* __restrict__ inj, const Sampler& s, int mc,
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

// W[mc][co][col..col+3], col % 4 == 0
__device__ __forceinline__ void conv_w4(const ConvParams& p, int mc, int co, int col, float w[4]) {
  float e[4];
  conv_eps4(p.eps_w, p.sw, mc, co, col, p.cout, p.kdim, e);
  const int64_t base = (int64_t)co * p.kdim + col;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    w[j] = (col + j < p.kdim)
               ? fmaf(softplus_f(__ldg(p.w_rho + base + j)), e[j], __ldg(p.w_mu + base + j))
               : 0.f;
}

// im2col element A[m][kidx] of sample mc
__device__ __forceinline__ float im2col_at(const ConvParams& p, const float* __restrict__ xp, int m,
                                           int kidx) {
  const int hw = p.ho * p.wo, kk = p.k * p.k;
  const int b = m / hw, pix = m - b * hw, oh = pix / p.wo, ow = pix - oh * p.wo;
  const int ci = kidx / kk, r = kidx - ci * kk, kh = r / p.k, kw = r - kh * p.k;
  const int ih = oh * p.stride - p.pad + kh * p.dil, iw = ow * p.stride - p.pad + kw * p.dil;
  if (ih < 0 || ih >= p.h || iw < 0 || iw >= p.w) return 0.f;
  return __ldg(xp + (((int64_t)b * p.cin + ci) * p.h + ih) * p.w + iw);
}

__device__ __forceinline__ void conv_fma_tile(const float (*As)[kCT + kCPad],
                                              const float (*Bs)[kCT + kCPad], int ty, int tx,
                                              float acc[4][4]) {
#pragma unroll
  for (int kk = 0; kk < kCK; ++kk) {
    const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
    const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
  }
}

// forward: grid (ceil(Cout/64), ceil(M/64), MC)
__global__ void __launch_bounds__(kCThreads) conv_fwd_simt(ConvParams p) {
  p.sw = live(p.sw);   // seed + device-side offset (CUDA-graph replays), see common.cuh
  p.sb = live(p.sb);
  __shared__ __align__(16) float As[kCK][kCT + kCPad];  // [kidx][pixel row]
  __shared__ __align__(16) float Bs[kCK][kCT + kCPad];  // [kidx][cout]
  const int mc = blockIdx.z, m0 = blockIdx.y * kCT, c0 = blockIdx.x * kCT;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int M = p.batch * p.ho * p.wo;
  const float* xp = p.x + (int64_t)mc * p.x_mc_stride;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < p.kdim; k0 += kCK) {
    {
      const int row = t >> 2, kq = (t & 3) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kidx = k0 + kq + j;
        As[kq + j][row] = (m0 + row < M && kidx < p.kdim) ? im2col_at(p, xp, m0 + row, kidx) : 0.f;
      }
    }
    {
      const int cl = t >> 2, g = (t & 3) * 4;
      float w[4] = {0.f, 0.f, 0.f, 0.f};
      if (c0 + cl < p.cout && k0 + g < p.kdim) conv_w4(p, mc, c0 + cl, k0 + g, w);
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[g + j][cl] = w[j];
    }
    __syncthreads();
    conv_fma_tile(As, Bs, ty, tx, acc);
    __syncthreads();
  }
  const int cc = c0 + tx * 4, hw = p.ho * p.wo;
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (cc < p.cout) {
    float e[4];
    conv_eps4(p.eps_b, p.sb, mc, 0, cc, 1, p.cout, e);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (cc + j < p.cout)
        bias[j] = fmaf(softplus_f(__ldg(p.b_rho + cc + j)), e[j], __ldg(p.b_mu + cc + j));
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int b = m / hw, pix = m - b * hw;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (cc + j < p.cout)
        p.y[(((int64_t)mc * p.batch + b) * p.cout + cc + j) * hw + pix] = acc[i][j] + bias[j];
  }
}

// dX: M = B*H*W input pixels, N = Cin, K = Cout*k*k.   grid (ceil(Cin/64), ceil(M/64), MC)
__global__ void __launch_bounds__(kCThreads) conv_dx_simt(ConvParams p) {
  p.sw = live(p.sw);   // seed + device-side offset (CUDA-graph replays), see common.cuh
  p.sb = live(p.sb);
  __shared__ __align__(16) float As[kCK][kCT + kCPad];  // [(co,kh,kw)][input pixel]
  __shared__ __align__(16) float Bs[kCK][kCT + kCPad];  // [(co,kh,kw)][cin]
  const int mc = blockIdx.z, m0 = blockIdx.y * kCT, c0 = blockIdx.x * kCT;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int M = p.batch * p.h * p.w, kk2 = p.k * p.k, K = p.cout * kk2, hwi = p.h * p.w,
            hwo = p.ho * p.wo;
  const float* dyp = p.dy + (int64_t)mc * p.batch * p.cout * hwo;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += kCK) {
    {
      const int row = t >> 2, kq = (t & 3) * 4;
      const int m = m0 + row;
      const int b = m / hwi, pix = m - b * hwi, ih = pix / p.w, iw = pix - ih * p.w;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kid = k0 + kq + j;
        float v = 0.f;
        if (m < M && kid < K) {
          const int co = kid / kk2, r = kid - co * kk2, kh = r / p.k, kw = r - kh * p.k;
          const int th = ih + p.pad - kh * p.dil, tw = iw + p.pad - kw * p.dil;
          if (th >= 0 && tw >= 0 && th % p.stride == 0 && tw % p.stride == 0) {
            const int oh = th / p.stride, ow = tw / p.stride;
            if (oh < p.ho && ow < p.wo)
              v = __ldg(dyp + ((int64_t)b * p.cout + co) * hwo + oh * p.wo + ow);
          }
        }
        As[kq + j][row] = v;
      }
    }
    {  // one weight per thread-slot: Bs[kid][cin]; thread -> (kid = t/16, 4 cin)
      const int kl = t >> 4, cl = (t & 15) * 4;
      const int kid = k0 + kl;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ci = c0 + cl + j;
        float v = 0.f;
        if (kid < K && ci < p.cin) {
          const int co = kid / kk2, r = kid - co * kk2;
          const int col = ci * kk2 + r;
          float w[4];
          conv_w4(p, mc, co, col & ~3, w);
          v = w[col & 3];
        }
        Bs[kl][cl + j] = v;
      }
    }
    __syncthreads();
    conv_fma_tile(As, Bs, ty, tx, acc);
    __syncthreads();
  }
  const int cc = c0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int b = m / hwi, pix = m - b * hwi;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (cc + j < p.cin)
        p.dx[(((int64_t)mc * p.batch + b) * p.cin + cc + j) * hwi + pix] = acc[i][j];
  }
}

// dW: tile (cout rows, kidx cols), loop MC and output pixels.  grid (ceil(Kdim/64), ceil(Cout/64))
__global__ void __launch_bounds__(kCThreads) conv_dw_simt(ConvParams p) {
  p.sw = live(p.sw);   // seed + device-side offset (CUDA-graph replays), see common.cuh
  p.sb = live(p.sb);
  __shared__ __align__(16) float As[kCK][kCT + kCPad];  // [pixel][cout]
  __shared__ __align__(16) float Bs[kCK][kCT + kCPad];  // [pixel][kidx]
  const int c0 = blockIdx.y * kCT, q0 = blockIdx.x * kCT;
  const int t = threadIdx.x, ty = t >> 4, tx = t & 15;
  const int kl = t >> 4, c4 = (t & 15) * 4;
  const int hwo = p.ho * p.wo, M = p.batch * hwo;
  float acc_mu[4][4] = {}, acc_rho[4][4] = {};
  for (int mc = 0; mc < p.mc; ++mc) {
    const float* xp = p.x + (int64_t)mc * p.x_mc_stride;
    const float* dyp = p.dy + (int64_t)mc * p.batch * p.cout * hwo;
    float acc[4][4] = {};
    for (int m0 = 0; m0 < M; m0 += kCK) {
      const int m = m0 + kl;
      const int b = m / hwo, pix = m - b * hwo;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = c0 + c4 + j, kidx = q0 + c4 + j;
        As[kl][c4 + j] = (m < M && co < p.cout) ? __ldg(dyp + ((int64_t)b * p.cout + co) * hwo + pix) : 0.f;
        Bs[kl][c4 + j] = (m < M && kidx < p.kdim) ? im2col_at(p, xp, m, kidx) : 0.f;
      }
      __syncthreads();
      conv_fma_tile(As, Bs, ty, tx, acc);
      __syncthreads();
    }
    const int qc = q0 + tx * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int co = c0 + ty * 4 + i;
      if (co >= p.cout || qc >= p.kdim) continue;
      float e[4];
      conv_eps4(p.eps_w, p.sw, mc, co, qc, p.cout, p.kdim, e);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        acc_mu[i][j] += acc[i][j];
        acc_rho[i][j] = fmaf(acc[i][j], e[j], acc_rho[i][j]);
      }
    }
  }
  const int qc = q0 + tx * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = c0 + ty * 4 + i;
    if (co >= p.cout) continue;
    const int64_t base = (int64_t)co * p.kdim + qc;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (qc + j < p.kdim) {
        p.dw_mu[base + j] = acc_mu[i][j];
        p.dw_rho[base + j] = acc_rho[i][j] * sigmoid_f(__ldg(p.w_rho + base + j));
      }
  }
}

// colsum[mc][co] = sum_{b,pix} dy ; one CTA per (co, mc)
__global__ void __launch_bounds__(256) conv_bias_colsum_kernel(ConvParams p) {
  p.sw = live(p.sw);   // seed + device-side offset (CUDA-graph replays), see common.cuh
  p.sb = live(p.sb);
  __shared__ float sm[8];
  const int co = blockIdx.x, mc = blockIdx.y, hwo = p.ho * p.wo;
  const float* dyp = p.dy + (int64_t)mc * p.batch * p.cout * hwo;
  float s = 0.f;
  if ((hwo & 3) == 0 && hwo >= 128 && (reinterpret_cast<uintptr_t>(p.dy) & 15) == 0) {
    // large planes: a warp per image, 16-byte loads, four independent partial sums per lane (the scalar loop below
    // is one dependent add per 4-byte load: 2.4 TB/s on C3's first layer)
    const int hwo4 = hwo >> 2, lane = threadIdx.x & 31;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = threadIdx.x >> 5; b < p.batch; b += 8) {
      const float4* src = reinterpret_cast<const float4*>(dyp + ((int64_t)b * p.cout + co) * hwo);
#pragma unroll 4
      for (int q = lane; q < hwo4; q += 32) {
        const float4 v = __ldg(src + q);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    s = (acc.x + acc.y) + (acc.z + acc.w);
  } else {
    for (int i = threadIdx.x; i < p.batch * hwo; i += 256) {
      const int b = i / hwo, pix = i - b * hwo;
      s += __ldg(dyp + ((int64_t)b * p.cout + co) * hwo + pix);
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sm[k];
    p.colsum[(int64_t)mc * p.cout + co] = s;
  }
}

__global__ void __launch_bounds__(128) conv_bias_grad_kernel(ConvParams p) {
  p.sw = live(p.sw);   // seed + device-side offset (CUDA-graph replays), see common.cuh
  p.sb = live(p.sb);
  const int o = blockIdx.x * 128 + threadIdx.x;
  if (o >= p.cout) return;
  float am = 0.f, ar = 0.f;
  for (int mc = 0; mc < p.mc; ++mc) {
    const float s = p.colsum[(int64_t)mc * p.cout + o];
    float e[4];
    conv_eps4(p.eps_b, p.sb, mc, 0, o & ~3, 1, p.cout, e);
    am += s;
    ar = fmaf(s, e[o & 3], ar);
  }
  p.db_mu[o] = am;
  p.db_rho[o] = ar * sigmoid_f(__ldg(p.b_rho + o));
}

static int fill_conv(const pl_conv2d_args* a, ConvParams& p, const char* who) {
  PL_CHECK_ARG(a, "%s: NULL args", who);
  PL_CHECK_ARG(a->mc >= 0 && a->batch >= 0 && a->cin > 0 && a->cout > 0 && a->h > 0 && a->w > 0,
               "%s: bad dimension", who);
  PL_CHECK_ARG(a->k > 0 && a->stride > 0 && a->padding >= 0 && a->dilation > 0,
               "%s: bad kernel_size/stride/padding/dilation", who);
  PL_CHECK_ARG(a->x && a->w_mu && a->w_rho && a->b_mu && a->b_rho, "%s: NULL tensor", who);
  p.mc = a->mc; p.batch = a->batch; p.cin = a->cin; p.cout = a->cout; p.h = a->h; p.w = a->w;
  p.k = a->k; p.stride = a->stride; p.pad = a->padding; p.dil = a->dilation;
  p.ho = (a->h + 2 * a->padding - a->dilation * (a->k - 1) - 1) / a->stride + 1;
  p.wo = (a->w + 2 * a->padding - a->dilation * (a->k - 1) - 1) / a->stride + 1;
  PL_CHECK_ARG(p.ho > 0 && p.wo > 0, "%s: empty output (%d x %d)", who, p.ho, p.wo);
  p.kdim = a->cin * a->k * a->k;
  PL_CHECK_ARG(a->x_mc_stride == 0 || a->x_mc_stride >= (int64_t)a->batch * a->cin * a->h * a->w,
               "%s: x_mc_stride overlaps samples", who);
  PL_CHECK_ARG((int64_t)a->batch * a->h * a->w < (1ll << 31) && (int64_t)a->batch * p.ho * p.wo < (1ll << 31),
               "%s: per-sample pixel count overflows int32", who);
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

size_t conv_tc_workspace_bytes(const pl_conv2d_args* a);
size_t conv_tc_state_bytes(const pl_conv2d_args* a);
int conv_fwd_tc_launch(const pl_conv2d_args* a);
int conv_bwd_tc_launch(const pl_conv2d_args* a);

}  // namespace pl

// tensor-core conv: bf16 operands only (split or single operand); there is no tf32 conv kernel
static inline bool conv_is_tc(int p) { return p == PL_PREC_BF16 || p == PL_PREC_BF16_FAST; }

extern "C" size_t pl_conv2d_workspace_bytes(const pl_conv2d_args* a) {
  if (!a) return 0;
  const size_t colsum = sizeof(float) * (size_t)(a->mc > 0 ? a->mc : 0) * (size_t)(a->cout > 0 ? a->cout : 0) + 256;
  if (conv_is_tc(a->precision)) {
    const size_t tc = pl::conv_tc_workspace_bytes(a);
    return tc > colsum ? tc : colsum;
  }
  return colsum;
}

extern "C" size_t pl_conv2d_state_bytes(const pl_conv2d_args* a) {
  if (!a || !conv_is_tc(a->precision)) return 0;
  return pl::conv_tc_state_bytes(a);
}

extern "C" int pl_conv2d_fwd(const pl_conv2d_args* a) {
  using namespace pl;
  PL_CHECK_ARG(a, "pl_conv2d_fwd: NULL args");
  if (conv_is_tc(a->precision)) return conv_fwd_tc_launch(a);
  ConvParams p;
  int rc = fill_conv(a, p, "pl_conv2d_fwd");
  if (rc != PL_OK) return rc;
  PL_CHECK_ARG(a->y, "pl_conv2d_fwd: NULL y");
  if (a->precision != PL_PREC_FP32) {
    set_error("pl_conv2d_fwd: precision %d not implemented", a->precision);
    return PL_ERR_UNSUPPORTED;
  }
  if (p.mc == 0 || p.batch == 0) return PL_OK;
  const int64_t M = (int64_t)p.batch * p.ho * p.wo;
  PL_CHECK_ARG(p.mc <= 65535 && ceil_div64(M, kCT) <= 65535, "pl_conv2d_fwd: grid too large");
  dim3 grid((unsigned)ceil_div64(p.cout, kCT), (unsigned)ceil_div64(M, kCT), (unsigned)p.mc);
  conv_fwd_simt<<<grid, kCThreads, 0, (cudaStream_t)a->stream>>>(p);
  PL_LAUNCH_CHECK("conv_fwd_simt");
  return PL_OK;
}

extern "C" int pl_conv2d_bwd(const pl_conv2d_args* a) {
  using namespace pl;
  PL_CHECK_ARG(a, "pl_conv2d_bwd: NULL args");
  if (conv_is_tc(a->precision) && a->mc > 0 && a->batch > 0) {
    // tensor-core kernels for dX / dW; the bias gradients are HBM-bound column sums (fp32 kernels)
    int rc = conv_bwd_tc_launch(a);
    if (rc != PL_OK || !a->db_mu) return rc;
    pl_conv2d_args b = *a;
    b.precision = PL_PREC_FP32;
    b.dx = nullptr;
    b.dw_mu = b.dw_rho = nullptr;
    return pl_conv2d_bwd(&b);
  }
  ConvParams p;
  int rc = fill_conv(a, p, "pl_conv2d_bwd");
  if (rc != PL_OK) return rc;
  PL_CHECK_ARG(a->dy, "pl_conv2d_bwd: NULL dy");
  PL_CHECK_ARG((a->dw_mu == nullptr) == (a->dw_rho == nullptr), "pl_conv2d_bwd: dw_mu/dw_rho must both be set or NULL");
  PL_CHECK_ARG((a->db_mu == nullptr) == (a->db_rho == nullptr), "pl_conv2d_bwd: db_mu/db_rho must both be set or NULL");
  if (a->precision != PL_PREC_FP32 && !conv_is_tc(a->precision)) {
    set_error("pl_conv2d_bwd: precision %d not implemented", a->precision);
    return PL_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)a->stream;
  const size_t wn = (size_t)p.cout * p.kdim;
  if (p.mc == 0 || p.batch == 0) {
    if (a->dw_mu) {
      PL_CUDA(cudaMemsetAsync(a->dw_mu, 0, sizeof(float) * wn, st));
      PL_CUDA(cudaMemsetAsync(a->dw_rho, 0, sizeof(float) * wn, st));
    }
    if (a->db_mu) {
      PL_CUDA(cudaMemsetAsync(a->db_mu, 0, sizeof(float) * p.cout, st));
      PL_CUDA(cudaMemsetAsync(a->db_rho, 0, sizeof(float) * p.cout, st));
    }
    return PL_OK;
  }
  if (a->dx) {
    const int64_t M = (int64_t)p.batch * p.h * p.w;
    PL_CHECK_ARG(p.mc <= 65535 && ceil_div64(M, kCT) <= 65535, "pl_conv2d_bwd: grid too large");
    dim3 grid((unsigned)ceil_div64(p.cin, kCT), (unsigned)ceil_div64(M, kCT), (unsigned)p.mc);
    conv_dx_simt<<<grid, kCThreads, 0, st>>>(p);
    PL_LAUNCH_CHECK("conv_dx_simt");
  }
  if (a->dw_mu) {
    dim3 grid((unsigned)ceil_div64(p.kdim, kCT), (unsigned)ceil_div64(p.cout, kCT));
    conv_dw_simt<<<grid, kCThreads, 0, st>>>(p);
    PL_LAUNCH_CHECK("conv_dw_simt");
  }
  if (a->db_mu) {
    const size_t need = sizeof(float) * (size_t)p.mc * p.cout;
    if (!a->workspace || a->workspace_bytes < need) {
      set_error("pl_conv2d_bwd: workspace %zu < %zu bytes", a->workspace_bytes, need);
      return PL_ERR_WORKSPACE;
    }
    dim3 grid((unsigned)p.cout, (unsigned)p.mc);
    conv_bias_colsum_kernel<<<grid, 256, 0, st>>>(p);
    PL_LAUNCH_CHECK("conv_bias_colsum_kernel");
    conv_bias_grad_kernel<<<(unsigned)ceil_div64(p.cout, 128), 128, 0, st>>>(p);
    PL_LAUNCH_CHECK("conv_bias_grad_kernel");
  }
  return PL_OK;
}
