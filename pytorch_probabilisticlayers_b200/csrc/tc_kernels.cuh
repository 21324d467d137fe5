// Tensor-core kernels (tcgen05 / TMEM / TMA) shared by the BayesLinear and BayesConv2d paths.
// Included by linear_tc.cu and conv_tc.cu (kernel templates + launch helpers).
#pragma once

#include <cuda.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "tc_common.cuh"

namespace pl {

using namespace tc;

constexpr int kBN = 128;               // output-column tile (UMMA N)
constexpr int kBK = 64;                // k-block (one 128-byte swizzle row of bf16)
constexpr int kGenWarps = 16;          // generator / epilogue warps (4 per SM sub-partition)
constexpr int kTcThreads = 64 + kGenWarps * 32;  // + TMA warp + MMA warp = 18 warps
constexpr int kSubTileBytes = 128 * kBK * 2;  // one [128 x 64] bf16 operand sub-tile = 16 KB

struct TcGemmParams {
  int mc, rows, ndim, kdim;     // rows = batch; C[mc] is [rows, ndim], reduction over kdim
  int w_in, w_out;              // weight matrix [I, O]
  int64_t a_mc_rows;            // rows per MC sample in the A tensor map (0 => broadcast input)
  const float* w_mu;            // [I, O]
  const float* w_sigma;         // softplus(rho), [I, O]                              (tf32 kernels)
  const uint4* w_pack;          // bf16 kernels: [I][ceil(O/8)] x {8 mu bf16 | 8 sigma bf16}, zero padded
  const uint4* w_sig16;         // split bf16 kernels: sigma as bf16 [I][ceil(O/8)] x 8, zero padded (the mean operand
                                // comes by TMA from a bf16 / tf32-rounded copy of mu: tmap_b of the kernel)
  const float* eps_w;           // optional injected eps [mc, I, O]
  Sampler sw;
  const float* bias;            // optional sampled bias [mc, ndim] (forward only)
  float* out;                   // [mc, rows, ndim]
  int out_hw;                   // 0: row-major out; >0: out[(mc*rows/out_hw + row/out_hw)][n][row%out_hw]
                                //    (conv: rows = pixels of all images, NCHW output)
  int w_aligned;                // weight rows 16-byte aligned (w_out % 4 == 0): 128-bit parameter loads
  // bf16 activation chaining between consecutive layers (row-major mode only):
  __nv_bfloat16* out_lp;        // also write the (activated / masked) result as bf16; `out` may be NULL
  int act;                      // 1: LeakyReLU(act_slope) on the result (forward)
  float act_slope;
  const __nv_bfloat16* mask_src;  // dX: multiply by act'(x): x > 0 ? 1 : mask_slope, x = mask_src[row][col]
  float mask_slope;
  // split kernels, MC-broadcast input (layer 1 of the reference's nets): X.round(mu) is the same for every MC sample, so
  // it is computed ONCE (mean_only launch, mc = 1) and the per-sample launches add it in the epilogue instead of
  // re-running the mean MMA for all MC samples (noise_only + mean_mat)
  int mean_only;                // generators idle, only the mean MMA: out = X.round(mu)
  int noise_only;               // mean tile neither loaded nor multiplied
  const float* mean_mat;        // [rows, ndim] fp32 added to every sample's result (row-major epilogue only)
};

// ---------------------------------------------------------------------------------------------
// HBM-bound pre-passes
// ---------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) softplus_kernel(const float* __restrict__ rho,
                                                       float* __restrict__ sigma, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    sigma[i] = softplus_f(rho[i]);
}

// bf16 kernels read the variational parameters as packed bf16: per weight row, per group of 8 columns
// (= one Philox call), 16 bytes of mu followed by 16 bytes of sigma = softplus(rho).  Halves the
// L2->SM traffic of the fused kernels (which run at the L2 slice throughput cap with fp32 parameters) and
// lets the generator finish W = mu + sigma*eps with one packed HFMA2.BF16 per two weights.
// Columns >= cols of the last group are zero (the operand tile must hold zeros there anyway).
static __global__ void __launch_bounds__(256) pack_mu_sigma_kernel(const float* __restrict__ mu,
                                                            const float* __restrict__ rho,
                                                            uint4* __restrict__ pack, int rows, int cols) {
  const int groups = (cols + 7) >> 3;
  const int64_t n = (int64_t)rows * groups;
  const bool vec = (cols & 3) == 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int r = (int)(i / groups), c = (int)(i - (int64_t)r * groups) * 8;
    const int64_t idx = (int64_t)r * cols + c;
    float m[8], sg[8];
    if (vec && c + 8 <= cols) {
      const float4 m0 = __ldg(reinterpret_cast<const float4*>(mu + idx)),
                   m1 = __ldg(reinterpret_cast<const float4*>(mu + idx) + 1),
                   r0 = __ldg(reinterpret_cast<const float4*>(rho + idx)),
                   r1 = __ldg(reinterpret_cast<const float4*>(rho + idx) + 1);
      m[0] = m0.x; m[1] = m0.y; m[2] = m0.z; m[3] = m0.w; m[4] = m1.x; m[5] = m1.y; m[6] = m1.z; m[7] = m1.w;
      sg[0] = softplus_f(r0.x); sg[1] = softplus_f(r0.y); sg[2] = softplus_f(r0.z); sg[3] = softplus_f(r0.w);
      sg[4] = softplus_f(r1.x); sg[5] = softplus_f(r1.y); sg[6] = softplus_f(r1.z); sg[7] = softplus_f(r1.w);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = c + j < cols;
        m[j] = ok ? __ldg(mu + idx + j) : 0.f;
        sg[j] = ok ? softplus_f(__ldg(rho + idx + j)) : 0.f;
      }
    }
    pack[2 * i] = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]),
                             pack_bf16x2(m[6], m[7]));
    pack[2 * i + 1] = make_uint4(pack_bf16x2(sg[0], sg[1]), pack_bf16x2(sg[2], sg[3]),
                                 pack_bf16x2(sg[4], sg[5]), pack_bf16x2(sg[6], sg[7]));
  }
}
static inline size_t pack_bytes(size_t rows, size_t cols) { return rows * ((cols + 7) / 8) * 32; }

// Split-operand (noise-preserving) kernels: the mean and the noise are separate MMA operands,
//   Y = X . round(mu) + X . round(sigma * eps),
// so sigma*eps is never added to mu in the operand precision.  (At the reference's initialisation sigma_w =
// sqrt(2/(I+O))/O is ~3e-4 of |mu| (src/ProbabilisticLayers.py:881), far below half a bf16 / tf32 ulp of mu: one
// rounded operand round(mu + sigma*eps) is the same number for every MC sample.)
// bf16: mu16 / sig16 = bf16 [rows][pitch], pitch = cols rounded up to 8, zero padded: mu16 is read by TMA
// (MN-major boxes forward, K-major boxes dX), sig16 by the generators (one 16-byte load per Philox group of 8).
static __global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ mu,
                                                                const float* __restrict__ rho,
                                                                uint4* __restrict__ mu16, uint4* __restrict__ sig16,
                                                                int rows, int cols) {
  const int groups = (cols + 7) >> 3;
  const int64_t n = (int64_t)rows * groups;
  const bool vec = (cols & 3) == 0;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const int r = (int)(i / groups), c = (int)(i - (int64_t)r * groups) * 8;
    const int64_t idx = (int64_t)r * cols + c;
    float m[8], sg[8];
    if (vec && c + 8 <= cols) {
      const float4 m0 = __ldg(reinterpret_cast<const float4*>(mu + idx)),
                   m1 = __ldg(reinterpret_cast<const float4*>(mu + idx) + 1),
                   r0 = __ldg(reinterpret_cast<const float4*>(rho + idx)),
                   r1 = __ldg(reinterpret_cast<const float4*>(rho + idx) + 1);
      m[0] = m0.x; m[1] = m0.y; m[2] = m0.z; m[3] = m0.w; m[4] = m1.x; m[5] = m1.y; m[6] = m1.z; m[7] = m1.w;
      sg[0] = softplus_f(r0.x); sg[1] = softplus_f(r0.y); sg[2] = softplus_f(r0.z); sg[3] = softplus_f(r0.w);
      sg[4] = softplus_f(r1.x); sg[5] = softplus_f(r1.y); sg[6] = softplus_f(r1.z); sg[7] = softplus_f(r1.w);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const bool ok = c + j < cols;
        m[j] = ok ? __ldg(mu + idx + j) : 0.f;
        sg[j] = ok ? softplus_f(__ldg(rho + idx + j)) : 0.f;
      }
    }
    mu16[i] = make_uint4(pack_bf16x2(m[0], m[1]), pack_bf16x2(m[2], m[3]), pack_bf16x2(m[4], m[5]),
                         pack_bf16x2(m[6], m[7]));
    sig16[i] = make_uint4(pack_bf16x2(sg[0], sg[1]), pack_bf16x2(sg[2], sg[3]), pack_bf16x2(sg[4], sg[5]),
                          pack_bf16x2(sg[6], sg[7]));
  }
}
// tf32: mu_hi = mu rounded to tf32 (fp32 word with the low 13 mantissa bits clear, read by TMA) and sigma, fp32 [rows][cols]
static __global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ mu,
                                                                const float* __restrict__ rho,
                                                                float* __restrict__ mu_hi, float* __restrict__ sigma,
                                                                int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    mu_hi[i] = __uint_as_float(to_tf32(mu[i]));
    sigma[i] = softplus_f(rho[i]);
  }
}
static inline size_t split16_bytes(size_t rows, size_t cols) { return rows * ((cols + 7) / 8) * 16; }

// fp32 -> bf16, 8 elements per thread per iteration (n % 8 == 0, 16-byte aligned)
static __global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src,
                                                        __nv_bfloat16* __restrict__ dst, int64_t n8) {
  const float4* s4 = reinterpret_cast<const float4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(dst);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n8; i += (int64_t)gridDim.x * 256) {
    const float4 a = s4[2 * i], b = s4[2 * i + 1];
    d4[i] = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y),
                       pack_bf16x2(b.z, b.w));
  }
}

// sampled bias b[mc][o] = b_mu + softplus(b_rho) * eps_b[mc][o]
static __global__ void __launch_bounds__(256) sample_bias_kernel(const float* __restrict__ b_mu,
                                                          const float* __restrict__ b_rho,
                                                          const float* __restrict__ eps_b, Sampler sb,
                                                          int mc, int out, float* __restrict__ dst) {
  sb = live(sb);
  const int groups = (out + 3) / 4;
  for (int t = blockIdx.x * 256 + threadIdx.x; t < mc * groups; t += gridDim.x * 256) {
    const int m = t / groups, g = t - m * groups;
    float e[4];
    if (eps_b) {
      for (int j = 0; j < 4; ++j) e[j] = (4 * g + j < out) ? eps_b[(int64_t)m * out + 4 * g + j] : 0.f;
    } else {
      const float4 n = eps4(sb, m, 0, g);
      e[0] = n.x; e[1] = n.y; e[2] = n.z; e[3] = n.w;
    }
    for (int j = 0; j < 4 && 4 * g + j < out; ++j) {
      const int o = 4 * g + j;
      dst[(int64_t)m * out + o] = fmaf(softplus_f(b_rho[o]), e[j], b_mu[o]);
    }
  }
}

// dY fp32 -> bf16 (operand of the dX / dW MMAs) fused with the per-sample column sums the bias
// gradients need: colsum[mc][o] = sum_b dY[mc][b][o].  One pass over dY instead of two.
// grid (ceil(O/128), MC), 256 threads = 16 row lanes x 16 column groups of 8.
static __global__ void __launch_bounds__(256) cast_dy_colsum_kernel(const float* __restrict__ dy,
                                                             __nv_bfloat16* __restrict__ dst,
                                                             float* __restrict__ colsum, int batch,
                                                             int out) {
  __shared__ float red[16][129];
  const int mc = blockIdx.y, cg = threadIdx.x & 15, ry = threadIdx.x >> 4;
  const int col = blockIdx.x * 128 + cg * 8;
  const bool ok = col < out;  // out % 8 == 0
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int64_t base = (int64_t)mc * batch * out;
  if (ok) {
    for (int b = ry; b < batch; b += 16) {
      const int64_t idx = base + (int64_t)b * out + col;
      const float4 v0 = *reinterpret_cast<const float4*>(dy + idx);
      const float4 v1 = *reinterpret_cast<const float4*>(dy + idx + 4);
      if (dst)   // tf32 path: dY is consumed as fp32 by TMA, only the column sums are needed
        *reinterpret_cast<uint4*>(dst + idx) =
            make_uint4(pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y),
                       pack_bf16x2(v1.z, v1.w));
      s[0] += v0.x; s[1] += v0.y; s[2] += v0.z; s[3] += v0.w;
      s[4] += v1.x; s[5] += v1.y; s[6] += v1.z; s[7] += v1.w;
    }
  }
  if (colsum == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ry][cg * 8 + j] = s[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) t += red[r][threadIdx.x];
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c < out) colsum[(int64_t)mc * out + c] = t;
  }
}

// column sums of a bf16 dY (activation-chained backward): colsum[mc][o] = sum_b dY[mc][b][o]
// grid (ceil(O/128), MC), 256 threads = 16 row lanes x 16 column groups of 8
static __global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ dy,
                                                                 float* __restrict__ colsum, int batch,
                                                                 int out) {
  __shared__ float red[16][129];
  const int mc = blockIdx.y, cg = threadIdx.x & 15, ry = threadIdx.x >> 4;
  const int col = blockIdx.x * 128 + cg * 8;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < out) {
    const int64_t base = (int64_t)mc * batch * out;
    for (int b = ry; b < batch; b += 16) {
      const uint4 q = *reinterpret_cast<const uint4*>(dy + base + (int64_t)b * out + col);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        s[2 * t] += __uint_as_float(w[t] << 16);
        s[2 * t + 1] += __uint_as_float(w[t] & 0xffff0000u);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[ry][cg * 8 + j] = s[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) t += red[r][threadIdx.x];
    const int c = blockIdx.x * 128 + threadIdx.x;
    if (c < out) colsum[(int64_t)mc * out + c] = t;
  }
}

// db_mu[o] = sum_mc colsum ; db_rho[o] = sigmoid(b_rho) sum_mc colsum * eps_b   (deterministic order)
static __global__ void __launch_bounds__(128) bias_grad_tc_kernel(const float* __restrict__ colsum,
                                                           const float* __restrict__ b_rho,
                                                           const float* __restrict__ eps_b, Sampler sb,
                                                           int mc, int out, float* __restrict__ db_mu,
                                                           float* __restrict__ db_rho) {
  sb = live(sb);
  const int o = blockIdx.x * 128 + threadIdx.x;
  if (o >= out) return;
  float am = 0.f, ar = 0.f;
  for (int m = 0; m < mc; ++m) {
    const float s = colsum[(int64_t)m * out + o];
    float e;
    if (eps_b) {
      e = eps_b[(int64_t)m * out + o];
    } else {
      const float4 n = eps4(sb, m, 0, o >> 2);
      const float v[4] = {n.x, n.y, n.z, n.w};
      e = v[o & 3];
    }
    am += s;
    ar = fmaf(s, e, ar);
  }
  db_mu[o] = am;
  db_rho[o] = ar * sigmoid_f(b_rho[o]);
}


// ---------------------------------------------------------------------------------------------
// W-tile generator shared by the 1-CTA and 2-CTA kernels.  A k-block of the weight operand is
// [128 n x kBKe k]; it is cut into groups of 8 consecutive o (the unit of one Philox call, 8 normals):
// along n for the forward (operand MN-major), along k for dX (operand K-major).  512 generator threads
// own 2 groups each per k-block in bf16 (64-k blocks) and 1 in tf32 (32-k blocks).  load() issues the
// mu / sigma (/ injected eps) loads of a k-block, emit() turns them into the swizzled smem operand.
// ---------------------------------------------------------------------------------------------
#ifndef PL_STAGES_A4
#define PL_STAGES_A4 2   // A-ring depth of the MT=4 kernels (64 KB per stage)
#endif
#ifndef PL_STAGES_B
#define PL_STAGES_B 4    // generated-operand ring depth (16 KB per stage)
#endif
#ifndef PL_PROBE_W
#define PL_PROBE_W 0     // tools/build_probe.sh builds instrumented copies with 1 / 2 (never the product)
#endif
template <bool kDx, bool kInjected, bool kTf32, bool kSplit = false>
struct WGen {
  static constexpr int kGr = kTf32 ? 1 : 2;
  static constexpr int kBKe = kTf32 ? 32 : 64;
  int wrow[kGr], wcol;   // weight coordinates of the thread's groups at k-block 0
  uint32_t soff[kGr];    // byte offset of the group inside the smem operand tile
  float4 mu[kTf32 ? kGr : 1][2], sg[kTf32 ? kGr : 1][2];   // tf32: fp32 parameters
  uint4 pk[kTf32 ? 1 : kGr][2];                            // bf16: packed {8 mu | 8 sigma}
  float4 ep[kInjected ? kGr : 1][2];
  int row[kGr], col;

  // nb0: first weight index along the operand's n axis handled by this CTA
  __device__ __forceinline__ void init(int gw, int lane, int nb0) {
#pragma unroll
    for (int r = 0; r < kGr; ++r) {
      if (!kDx) {
        // group = 8 n at one k.  bf16: k = 2 gw + lane/16 + 32 r, n = 8 (lane%16); MN-major SW128:
        //   [n/64]*8192 + [k/8]*1024 + (k%8)*128 + (((n%64)/8) ^ (k%8))*16
        // tf32: k = 2 gw + lane/16, n = 8 (lane%16); SWIZZLE_128B_BASE32B:
        //   [n/32]*4096 + [k/4]*512 + (k%4)*128 + (((n%32)/8) ^ (k%4))*32   (one 32-byte chunk)
        const int k = 2 * gw + (lane >> 4) + 32 * r, n8 = lane & 15;
        wrow[r] = k;
        if (kTf32)
          soff[r] = (uint32_t)((n8 >> 2) * 4096 + (k >> 2) * 512 + (k & 3) * 128 + (((n8 & 3) ^ (k & 3)) << 5));
        else
          soff[r] = (uint32_t)((n8 >> 3) * 8192 + (k >> 3) * 1024 + (k & 7) * 128 + (((n8 & 7) ^ (k & 7)) << 4));
      } else {
        // group = 8 k at one n.  K-major SW128: n*128 + ((16-byte chunk of k) ^ (n%8))*16
        // bf16: n = 4 gw + lane/8 + 64 r, k = 8 (lane%8) (one chunk);
        // tf32: n = 8 gw + lane/4, k = 8 (lane%4) (two chunks 2c, 2c+1)
        const int n = kTf32 ? 8 * gw + (lane >> 2) : 4 * gw + (lane >> 3) + 64 * r;
        wrow[r] = nb0 + n;
        soff[r] = (uint32_t)(n * 128);      // chunk term added in emit()
      }
    }
    wcol = kDx ? 8 * (kTf32 ? (lane & 3) : (lane & 7)) : nb0 + 8 * (lane & 15);
  }

  template <bool kFull>
  __device__ __forceinline__ void load(const TcGemmParams& p, int mc, int k0) {
    col = kDx ? k0 + wcol : wcol;
    const float* eps_base = kInjected ? p.eps_w + (int64_t)mc * p.w_in * p.w_out : nullptr;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    const int pack_groups = (p.w_out + 7) >> 3;
#pragma unroll
    for (int r = 0; r < kGr; ++r) {
      row[r] = kDx ? wrow[r] : k0 + wrow[r];
      const bool ok = kFull || (row[r] < p.w_in && col < p.w_out);
#if PL_PROBE_W == 1      // probe: all parameter loads hit one 64 KB window (L1/L2-resident)
      const int64_t idx = ((int64_t)row[r] * p.w_out + col) & 0x3FF8;
#else
      const int64_t idx = (int64_t)row[r] * p.w_out + col;
#endif
      if (!kTf32) {
        const uint4 zz = make_uint4(0, 0, 0, 0);
        if (kSplit) {
          // noise operand only: one aligned 16-byte load of 8 bf16 sigmas (zero beyond w_out)
          pk[r][1] = ok ? __ldg(p.w_sig16 + (int64_t)row[r] * pack_groups + (col >> 3)) : zz;
        } else {
        // packed parameters: always two aligned 16-byte loads; the pack is zero beyond w_out
        const uint4* src = p.w_pack + ((int64_t)row[r] * pack_groups + (col >> 3)) * 2;
#if PL_PROBE_W == 2      // probe: no parameter loads at all
        pk[r][0] = make_uint4(0x3dcc3dcc, 0x3dcc3dcc, 0x3dcc3dcc, 0x3dcc3dcc);
        pk[r][1] = make_uint4(0x3c233c23, 0x3c233c23, 0x3c233c23, 0x3c233c23);
#else
        pk[r][0] = ok ? __ldg(src) : zz;
        pk[r][1] = ok ? __ldg(src + 1) : zz;
#endif
        }
        if (kInjected) {
          float e8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) e8[j] = (ok && col + j < p.w_out) ? __ldg(eps_base + idx + j) : 0.f;
          ep[r][0] = make_float4(e8[0], e8[1], e8[2], e8[3]);
          ep[r][1] = make_float4(e8[4], e8[5], e8[6], e8[7]);
        }
        continue;
      }
      if (kFull || (p.w_aligned && col + 8 <= p.w_out)) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mu[r][h] = (ok && !kSplit) ? __ldg(reinterpret_cast<const float4*>(p.w_mu + idx) + h) : z;
          sg[r][h] = ok ? __ldg(reinterpret_cast<const float4*>(p.w_sigma + idx) + h) : z;
          if (kInjected) ep[r][h] = ok ? __ldg(reinterpret_cast<const float4*>(eps_base + idx) + h) : z;
        }
      } else {   // rows not 16-byte aligned / ragged last group (conv K = Cin*k*k): scalar loads
        float m8[8], s8[8], e8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool okj = ok && col + j < p.w_out;
          m8[j] = (okj && !kSplit) ? __ldg(p.w_mu + idx + j) : 0.f;
          s8[j] = okj ? __ldg(p.w_sigma + idx + j) : 0.f;
          e8[j] = (kInjected && okj) ? __ldg(eps_base + idx + j) : 0.f;
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mu[r][h] = make_float4(m8[4 * h], m8[4 * h + 1], m8[4 * h + 2], m8[4 * h + 3]);
          sg[r][h] = make_float4(s8[4 * h], s8[4 * h + 1], s8[4 * h + 2], s8[4 * h + 3]);
          if (kInjected) ep[r][h] = make_float4(e8[4 * h], e8[4 * h + 1], e8[4 * h + 2], e8[4 * h + 3]);
        }
      }
    }
  }

  // noise(): eps of the groups loaded last (Philox + Box-Muller; independent of the parameter loads).
  // bf16 keeps it as 4 packed bf16 pairs per group, tf32 as 8 floats.
  uint32_t nz[kTf32 ? 1 : kGr][4];
  float nf[kTf32 ? kGr : 1][8];
  // The Philox words of a k-block are drawn one iteration ahead (words()) so that, inside one basic block,
  // the integer-multiply chain of block kb+1 interleaves with the MUFU-bound Box-Muller of block kb.
  uint4 wd[kInjected ? 1 : kGr];
  __device__ __forceinline__ void words(const TcGemmParams& p, const Sampler& sw, int mc, int k0) {
    if (kInjected) return;
#pragma unroll
    for (int r = 0; r < kGr; ++r) {
      const int rown = kDx ? wrow[r] : k0 + wrow[r], coln = kDx ? k0 + wcol : wcol;
      wd[r] = philox4x32_10((uint32_t)(coln >> 3), (uint32_t)rown, sw.mc_offset + (uint32_t)mc, sw.stream,
                            sw.k0, sw.k1);
    }
  }
  __device__ __forceinline__ void noise() {
#pragma unroll
    for (int r = 0; r < kGr; ++r) {
      float e[8];
      if (kInjected) {
        e[0] = ep[r][0].x; e[1] = ep[r][0].y; e[2] = ep[r][0].z; e[3] = ep[r][0].w;
        e[4] = ep[r][1].x; e[5] = ep[r][1].y; e[6] = ep[r][1].z; e[7] = ep[r][1].w;
      } else {
        box_muller(wd[r].x, e[0], e[1]);
        box_muller(wd[r].y, e[2], e[3]);
        box_muller(wd[r].z, e[4], e[5]);
        box_muller(wd[r].w, e[6], e[7]);
      }
      if (kTf32) {
#pragma unroll
        for (int j = 0; j < 8; ++j) nf[r][j] = e[j];
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) nz[r][j] = pack_bf16x2(e[2 * j], e[2 * j + 1]);
      }
    }
  }
  // pin: noise and next words are complete before the (volatile) barrier traffic that follows in program order
  __device__ __forceinline__ void pin() {
#pragma unroll
    for (int r = 0; r < kGr; ++r) {
      if (!kTf32) {
#pragma unroll
        for (int j = 0; j < 4; ++j) asm volatile("" : "+r"(nz[r][j]));
      }
      if (!kInjected) asm volatile("" : "+r"(wd[r].x), "+r"(wd[r].y), "+r"(wd[r].z), "+r"(wd[r].w));
    }
  }

  __device__ __forceinline__ void store(uint32_t dst, int lane) {
#pragma unroll
    for (int r = 0; r < kGr; ++r) {
      float e[8];
      if (kTf32) {
#pragma unroll
        for (int j = 0; j < 8; ++j) e[j] = nf[r][j];
      }
      if (!kTf32) {
        // W = mu + sigma * eps, two weights per HFMA2.BF16 (eps rounded to bf16, one rounding of the result)
        uint32_t a0 = dst + soff[r];
        if (kDx) a0 += (uint32_t)(((lane & 7) ^ ((soff[r] >> 7) & 7)) << 4);
        if (kSplit)   // noise operand: sigma * eps alone (the mean operand is a separate TMA-loaded tile)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0),
                       "r"(mul_bf16x2(pk[r][1].x, nz[r][0])), "r"(mul_bf16x2(pk[r][1].y, nz[r][1])),
                       "r"(mul_bf16x2(pk[r][1].z, nz[r][2])), "r"(mul_bf16x2(pk[r][1].w, nz[r][3])) : "memory");
        else
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0),
                     "r"(fma_bf16x2(pk[r][1].x, nz[r][0], pk[r][0].x)),
                     "r"(fma_bf16x2(pk[r][1].y, nz[r][1], pk[r][0].y)),
                     "r"(fma_bf16x2(pk[r][1].z, nz[r][2], pk[r][0].z)),
                     "r"(fma_bf16x2(pk[r][1].w, nz[r][3], pk[r][0].w)) : "memory");
        continue;
      }
      const float w0 = fmaf(sg[r][0].x, e[0], mu[r][0].x), w1 = fmaf(sg[r][0].y, e[1], mu[r][0].y),
                  w2 = fmaf(sg[r][0].z, e[2], mu[r][0].z), w3 = fmaf(sg[r][0].w, e[3], mu[r][0].w),
                  w4 = fmaf(sg[r][1].x, e[4], mu[r][1].x), w5 = fmaf(sg[r][1].y, e[5], mu[r][1].y),
                  w6 = fmaf(sg[r][1].z, e[6], mu[r][1].z), w7 = fmaf(sg[r][1].w, e[7], mu[r][1].w);
      // tf32: 8 fp32 words = two 16-byte chunks.  forward: both inside one 32-byte chunk of the BASE32B
      // layout; dX: chunks 2c and 2c+1 of the K-major row, each XOR-swizzled with (n % 8)
      uint32_t a0 = dst + soff[r], a1 = a0 + 16;
      if (kDx) {
        const int n7 = (soff[r] >> 7) & 7, c = 2 * (lane & 3);
        a0 = dst + soff[r] + (uint32_t)((c ^ n7) << 4);
        a1 = dst + soff[r] + (uint32_t)(((c + 1) ^ n7) << 4);
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "r"(to_tf32(w0)), "r"(to_tf32(w1)),
                   "r"(to_tf32(w2)), "r"(to_tf32(w3)) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "r"(to_tf32(w4)), "r"(to_tf32(w5)),
                   "r"(to_tf32(w6)), "r"(to_tf32(w7)) : "memory");
    }
  }
};

// ---------------------------------------------------------------------------------------------
// fused sample + GEMM (forward: kDx = false, B operand MN-major; dX: kDx = true, B operand K-major)
// ---------------------------------------------------------------------------------------------
template <int MT, bool kSplit = false>
struct TcCfg {
  // split: every B stage holds the generated noise tile AND the TMA-loaded mean tile (2 x 16 KB)
  static constexpr int kStagesA = MT == 4 ? PL_STAGES_A4 : ((MT == 2 && kSplit) ? 3 : 4);
  static constexpr int kStagesB = kSplit ? (MT == 1 ? 4 : 3) : PL_STAGES_B;
  static constexpr int kABytes = MT * kSubTileBytes;
  static constexpr int kBBytes = (kSplit ? 2 : 1) * kSubTileBytes;
  static constexpr int kSmemBytes = kStagesA * kABytes + kStagesB * kBBytes + 1024 /*bias + barriers*/
                                    + 1024 /*alignment slack*/;
  static constexpr int kTmemCols = MT * kBN;  // 128, 256 or 512: powers of two
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
};

// kSplit: the weight operand is two tiles per k-block, accumulated into the same TMEM columns:
//   the mean tile  round(mu)          loaded by TMA (tmap_b) from a bf16 / tf32-rounded copy of mu, and
//   the noise tile round(sigma * eps) synthesised by the generator warps,
// i.e. twice the MMAs (the tensor pipe is ~40 % busy in the single-operand kernel) for a forward whose
// Monte-Carlo noise survives however small sigma is next to mu.
template <int MT, bool kDx, bool kInjected, bool kTf32, bool kSplit>
__global__ void __launch_bounds__(kTcThreads, 1)
sampled_gemm_tc(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const TcGemmParams p) {
  using Cfg = TcCfg<MT, kSplit>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + Cfg::kStagesA * Cfg::kABytes;
  const uint32_t sMisc = sB + Cfg::kStagesB * Cfg::kBBytes;
  float* bias_s = reinterpret_cast<float*>(smem + (sMisc - smem_base));  // 128 floats
  const uint32_t bar0 = sMisc + 512;
  // barrier map (8 bytes each)
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (Cfg::kStagesA + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * Cfg::kStagesA + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (2 * Cfg::kStagesA + Cfg::kStagesB + s); };
  const uint32_t acc_full = bar0 + 8u * (2 * Cfg::kStagesA + 2 * Cfg::kStagesB);
  const uint32_t tmem_slot = acc_full + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform role index
  const int n0 = blockIdx.x * kBN;
  const int m0 = blockIdx.y * (MT * 128);
  const int mc = blockIdx.z;
  // elements per 128-byte swizzle row = k-block: 64 (bf16) or 32 (tf32: fp32 words); one MMA
  // consumes 32 bytes of K (16 bf16 / 8 tf32), so a k-block is always 4 MMA k-steps
  constexpr int kBKe = kTf32 ? 32 : 64;
  const int num_kb = (p.kdim + kBKe - 1) / kBKe;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStagesA; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < Cfg::kStagesB; ++s) {
      mbar_init(b_full(s), kGenWarps + (kSplit ? 1 : 0));   // one elected arrival per generator warp (+ the mean tile's TMA)
      mbar_init(b_empty(s), 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_a);
    if (kSplit) tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::kTmemCols);
  if (warp >= 2) {  // stage the sampled bias of this column tile
    const int t = threadIdx.x - 64;
    if (t < kBN) bias_s[t] = (p.bias && n0 + t < p.ndim) ? p.bias[(int64_t)mc * p.ndim + n0 + t] : 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int row0 = (int)(mc * p.a_mc_rows) + m0;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (kSplit && p.noise_only) {
          // no mean tile: only this producer's arrival on the stage's full barrier
          const int sb = kb % Cfg::kStagesB, pb = (kb / Cfg::kStagesB) & 1;
          mbar_wait(b_empty(sb), pb ^ 1);
          mbar_arrive(b_full(sb));
        } else if (kSplit) {
          // mean tile of this k-block into the second half of B stage sb (freed by the same commit as the noise half)
          const int sb = kb % Cfg::kStagesB, pb = (kb / Cfg::kStagesB) & 1;
          mbar_wait(b_empty(sb), pb ^ 1);
          mbar_expect_tx(b_full(sb), kSubTileBytes);
          const uint32_t dst = sB + sb * Cfg::kBBytes + kSubTileBytes;
          if (kDx) {
            // K-major: [128 weight rows n0..] x [kBKe weight columns] = one 16 KB box
            tma_load_2d(dst, &tmap_b, b_full(sb), kb * kBKe, n0);
          } else {
            // MN-major: [kBKe weight rows] x [128 weight columns n0..] as 128-byte wide boxes
            constexpr int kBoxes = kTf32 ? 4 : 2, kCols = kTf32 ? 32 : 64;
#pragma unroll
            for (int g = 0; g < kBoxes; ++g)
              tma_load_2d(dst + g * (kSubTileBytes / kBoxes), &tmap_b, b_full(sb), n0 + g * kCols, kb * kBKe);
          }
        }
        const int s = kb % Cfg::kStagesA, ph = (kb / Cfg::kStagesA) & 1;
        mbar_wait(a_empty(s), ph ^ 1);
        mbar_expect_tx(a_full(s), Cfg::kABytes);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
          tma_load_2d(sA + s * Cfg::kABytes + mt * kSubTileBytes, &tmap_a, a_full(s), kb * kBKe,
                      row0 + mt * 128);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_for<kTf32>(128, kBN, /*a MN-major*/ 0, /*b MN-major*/ kDx ? 0 : 1);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int sa = kb % Cfg::kStagesA, pa = (kb / Cfg::kStagesA) & 1;
        const int sb = kb % Cfg::kStagesB, pb = (kb / Cfg::kStagesB) & 1;
        mbar_wait(a_full(sa), pa);
        mbar_wait(b_full(sb), pb);
        tc_fence_after_sync();
        const uint32_t a_addr = sA + sa * Cfg::kABytes;
        const uint32_t b_addr = sB + sb * Cfg::kBBytes;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          // A: K-major SW128 (8-row groups 1024 B apart); advancing K by 16 bf16 = 32 B
          // B: forward MN-major SW128 (64-col groups 8192 B apart = LBO, 8-k groups 1024 B = SBO,
          //    16 k = 2 groups = 2048 B); dX K-major SW128 like A
          // (tf32 MN-major: SWIZZLE_128B_BASE32B, 32-col groups 4096 B apart, 4-k groups 512 B apart,
          //  8 k = 1024 B per MMA)
          auto b_desc_at = [&](uint32_t base) {
            return kDx ? smem_desc_sw128(base + ks * 32, 16, 1024)
                       : (kTf32 ? smem_desc_sw128_base32(base + ks * 1024, 4096, 512)
                                : smem_desc_sw128(base + ks * 2048, 8192, 1024));
          };
          const bool do_noise = !(kSplit && p.mean_only), do_mean = kSplit && !p.noise_only;
          if (do_noise) {
            const uint64_t b_desc = b_desc_at(b_addr);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint64_t a_desc = smem_desc_sw128(a_addr + mt * kSubTileBytes + ks * 32, 16, 1024);
              mma_ss<kTf32>(tmem_base + mt * kBN, a_desc, b_desc, idesc, (kb | ks) != 0);
            }
          }
          if (do_mean) {   // the mean tile (same layout, second half of the stage) into the same accumulators
            const uint64_t m_desc = b_desc_at(b_addr + kSubTileBytes);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const uint64_t a_desc = smem_desc_sw128(a_addr + mt * kSubTileBytes + ks * 32, 16, 1024);
              mma_ss<kTf32>(tmem_base + mt * kBN, a_desc, m_desc, idesc, do_noise ? 1u : (uint32_t)((kb | ks) != 0));
            }
          }
        }
        mma_commit(a_empty(sa));   // frees the activation stage when these MMAs retire
        mma_commit(b_empty(sb));   // frees the generated-weight stage
      }
      mma_commit(acc_full);
    }
    __syncwarp();
  } else {
    // ===================== W-tile generators =====================
    // operand element (n, k): forward W[i = k0+k][o = n0+n]; dX W[i = n0+n][o = k0+k]
    const int gw = warp - 2;  // 0..15
    const Sampler sw = live(p.sw);
    WGen<kDx, kInjected, kTf32, kSplit> gen;
    gen.init(gw, lane, n0);
    // tiles fully inside the weight matrix take the predicate-free path
    const bool n_full = p.w_aligned && (kDx ? (n0 + kBN <= p.w_in) : (n0 + kBN <= p.w_out));
    const int k_extent = kDx ? p.w_out : p.w_in;
    auto load_block = [&](int kb) {
#if PL_PROBE_W != 3     // probe 3: barrier hand-shakes only (MMA + TMA pipeline floor)
      if (n_full && (kb + 1) * kBKe <= k_extent) gen.template load<true>(p, mc, kb * kBKe);
      else gen.template load<false>(p, mc, kb * kBKe);
#endif
    };
    if (kSplit && p.mean_only) {
      // nothing to synthesise: keep the stage hand-shake alive (the MMA thread only multiplies the TMA-loaded mean tiles)
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % Cfg::kStagesB, ph = (kb / Cfg::kStagesB) & 1;
        mbar_wait(b_empty(s), ph ^ 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(b_full(s));
      }
    } else {
    gen.words(p, sw, mc, 0);
    load_block(0);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % Cfg::kStagesB, ph = (kb / Cfg::kStagesB) & 1;
#if PL_PROBE_W != 3
      gen.noise();                            // Box-Muller of this block ...
      gen.words(p, sw, mc, (kb + 1) * kBKe);      // ... interleaved with the Philox chain of the next one
      gen.pin();
#endif
      // The previous k-block is published only now.  fence.proxy.async is MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC:
      // it waits for every outstanding memory operation of the warp, so it sits where the st.shared of the
      // previous block and the parameter loads of this one (issued a whole noise() ago) have completed.
      if (kb > 0) {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_full((kb - 1) % Cfg::kStagesB));
      }
      mbar_wait(b_empty(s), ph ^ 1);
#if PL_PROBE_W != 3
      gen.store(sB + s * Cfg::kBBytes, lane);
#endif
      if (kb + 1 < num_kb) load_block(kb + 1);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(b_full((num_kb - 1) % Cfg::kStagesB));
    }
    // ===================== epilogue: TMEM -> registers -> HBM =====================
    const int q = warp & 3;        // TMEM lane quadrant this warp may access
    const int col = (gw >> 2) * 32;  // which 32-column quarter of the tile
    // chained dX: the activation mask (this lane's 64 bytes of bf16 x) is fetched before the accumulator is
    // waited for, so its L2 round trip hides behind the last MMAs / the TMEM load instead of following them
    uint4 mk[4];
    auto fetch_mask = [&](int mt) {
      const int row = m0 + mt * 128 + q * 32 + lane;
      const __nv_bfloat16* src = p.mask_src + ((int64_t)mc * p.rows + row) * p.ndim + n0 + col;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        mk[c] = (row < p.rows && n0 + col + 8 * c < p.ndim) ? __ldg(reinterpret_cast<const uint4*>(src) + c)
                                                            : make_uint4(0, 0, 0, 0);
    };
    if (p.mask_src) fetch_mask(0);
    mbar_wait(acc_full, 0);
    tc_fence_after_sync();
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const int row = m0 + mt * 128 + q * 32 + lane;
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * kBN + col), v);
      tmem_ld_wait();
      if (row < p.rows) {
        if (p.out_hw == 0) {
          const int64_t obase = ((int64_t)mc * p.rows + row) * p.ndim + n0 + col;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (n0 + col + j < p.ndim) {   // ndim % 8 == 0 on this path
              float w[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) w[t] = __uint_as_float(v[j + t]) + bias_s[col + j + t];
              if (p.mean_mat) {   // the MC-shared mean part X.round(mu), computed once
                const float4* mm = reinterpret_cast<const float4*>(p.mean_mat + (int64_t)row * p.ndim + n0 + col + j);
                const float4 m0 = __ldg(mm), m1 = __ldg(mm + 1);
                w[0] += m0.x; w[1] += m0.y; w[2] += m0.z; w[3] += m0.w;
                w[4] += m1.x; w[5] += m1.y; w[6] += m1.z; w[7] += m1.w;
              }
              if (p.act) {
#pragma unroll
                for (int t = 0; t < 8; ++t) w[t] = w[t] > 0.f ? w[t] : p.act_slope * w[t];
              }
              if (p.mask_src) {
                const uint4 mkj = mk[j >> 3];
                const uint32_t mw[4] = {mkj.x, mkj.y, mkj.z, mkj.w};
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                  // bf16 sign / zero test on the raw bits: positive and non-zero
                  const uint32_t h = (t & 1) ? (mw[t >> 1] >> 16) : (mw[t >> 1] & 0xffffu);
                  const bool pos = (h & 0x8000u) == 0 && (h & 0x7fffu) != 0;
                  w[t] = pos ? w[t] : p.mask_slope * w[t];
                }
              }
              if (p.out) {
                *reinterpret_cast<float4*>(p.out + obase + j) = make_float4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<float4*>(p.out + obase + j + 4) = make_float4(w[4], w[5], w[6], w[7]);
              }
              if (p.out_lp)
                *reinterpret_cast<uint4*>(p.out_lp + obase + j) =
                    make_uint4(pack_bf16x2(w[0], w[1]), pack_bf16x2(w[2], w[3]), pack_bf16x2(w[4], w[5]),
                               pack_bf16x2(w[6], w[7]));
            }
          }
        } else {
          // channel-major output: consecutive lanes (rows = pixels) are contiguous in memory
          const int img = row / p.out_hw, pix = row - img * p.out_hw;
          const int64_t ooff = (((int64_t)mc * (p.rows / p.out_hw) + img) * p.ndim + n0 + col) * p.out_hw + pix;
          if (p.out_lp) {   // bf16 (the conv backward's dA intermediate)
            __nv_bfloat16* o = p.out_lp + ooff;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + col + j < p.ndim)
                o[(int64_t)j * p.out_hw] = __float2bfloat16(__uint_as_float(v[j]) + bias_s[col + j]);
          } else {
            float* o = p.out + ooff;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + col + j < p.ndim) o[(int64_t)j * p.out_hw] = __uint_as_float(v[j]) + bias_s[col + j];
          }
        }
      }
      if (p.mask_src && mt + 1 < MT) fetch_mask(mt + 1);   // in flight during the next TMEM load
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}


// ---------------------------------------------------------------------------------------------
// 2-CTA version of sampled_gemm_tc (cta_group::2): a CTA pair on one TPC owns a
// [MT2 x 256 batch rows] x [256 output columns] tile.  Each CTA stages its own 128 rows of every
// 256-row activation sub-tile (TMA) and synthesises its own 128 of the 256 weight columns; the
// leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256, N = 256) which reads both
// CTAs' shared memory and writes 128 accumulator rows into each CTA's TMEM.  Per SM this halves the
// shared-memory operand traffic of the 1-CTA kernel (8 KB per 128 MMA cycles instead of 8 KB per 64)
// and the activation bytes per k-block, at the same generated-elements-per-flop ratio.
//   full barriers live in the leader CTA (peer producers arrive remotely, TMA bytes of both CTAs
//   are credited to it); empty / accumulator barriers are multicast to both CTAs by tcgen05.commit.
// ---------------------------------------------------------------------------------------------
template <int MT2, bool kSplit = false>
struct Tc2Cfg {
  // split: every B stage = this CTA's generated noise tile + its TMA-loaded mean tile (2 x 16 KB)
  static constexpr int kStagesA = kSplit ? 3 : 4;
  static constexpr int kStagesB = kSplit ? 3 : 4;
  static constexpr int kABytes = MT2 * kSubTileBytes;          // this CTA's rows of one k-block
  static constexpr int kBBytes = (kSplit ? 2 : 1) * kSubTileBytes;
  static constexpr int kSmemBytes = kStagesA * kABytes + kStagesB * kBBytes + 2048 + 1024;
  static constexpr int kTmemCols = MT2 * 256;                   // 256 or 512
  static_assert(kSmemBytes <= 232448, "exceeds the 227 KB of dynamic shared memory per CTA");
};

template <int MT2, bool kDx, bool kInjected, bool kSplit>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
sampled_gemm_tc2(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const TcGemmParams p) {
  using Cfg = Tc2Cfg<MT2, kSplit>;
  constexpr int kTileN = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + Cfg::kStagesA * Cfg::kABytes;
  const uint32_t sMisc = sB + Cfg::kStagesB * Cfg::kBBytes;
  float* bias_s = reinterpret_cast<float*>(smem + (sMisc - smem_base));  // 256 floats
  const uint32_t bar0 = sMisc + 1024;
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (Cfg::kStagesA + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * Cfg::kStagesA + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (2 * Cfg::kStagesA + Cfg::kStagesB + s); };
  const uint32_t acc_full = bar0 + 8u * (2 * Cfg::kStagesA + 2 * Cfg::kStagesB);
  const uint32_t tmem_slot = acc_full + 8;
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform role index
  const uint32_t rank = cluster_ctarank();        // 0 = leader
  const int n0 = (blockIdx.x >> 1) * kTileN;      // pair tile: 256 output columns
  const int nb0 = n0 + (int)rank * 128;           // the 128 columns this CTA synthesises
  const int m0 = blockIdx.y * (MT2 * 256);
  const int mc = blockIdx.z;
  const int num_kb = (p.kdim + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStagesA; ++s) {
      mbar_init(a_full(s), 2);                    // leader producer (+tx) and peer producer
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < Cfg::kStagesB; ++s) {
      // one elected arrival per generator warp of both CTAs (+ split: the two mean-tile producers, like a_full)
      mbar_init(b_full(s), 2 * kGenWarps + (kSplit ? 2 : 0));
      mbar_init(b_empty(s), 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_a);
    if (kSplit) tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, Cfg::kTmemCols);
  if (warp >= 2) {
    const int t = threadIdx.x - 64;
    if (t < kTileN) bias_s[t] = (p.bias && n0 + t < p.ndim) ? p.bias[(int64_t)mc * p.ndim + n0 + t] : 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs, own rows / own weight columns) =====================
    if (lane == 0) {
      const int row0 = (int)(mc * p.a_mc_rows) + m0 + (int)rank * 128;
      for (int kb = 0; kb < num_kb; ++kb) {
        if (kSplit) {
          // this CTA's half of the mean tile, credited to the leader's b_full like the activation bytes
          const int sb = kb % Cfg::kStagesB, pb = (kb / Cfg::kStagesB) & 1;
          mbar_wait(b_empty(sb), pb ^ 1);
          if (rank == 0) mbar_expect_tx(b_full(sb), 2 * kSubTileBytes);
          else mbar_arrive_cluster(b_full(sb), 0);
          const uint32_t dst = sB + sb * Cfg::kBBytes + kSubTileBytes;
          if (kDx) {
            tma_load_2d_2cta(dst, &tmap_b, b_full(sb), kb * kBK, nb0);
          } else {
#pragma unroll
            for (int g = 0; g < 2; ++g)
              tma_load_2d_2cta(dst + g * (kSubTileBytes / 2), &tmap_b, b_full(sb), nb0 + g * 64, kb * kBK);
          }
        }
        const int s = kb % Cfg::kStagesA, ph = (kb / Cfg::kStagesA) & 1;
        mbar_wait(a_empty(s), ph ^ 1);
        if (rank == 0) mbar_expect_tx(a_full(s), 2 * Cfg::kABytes);
        else mbar_arrive_cluster(a_full(s), 0);
#pragma unroll
        for (int t = 0; t < MT2; ++t)
          tma_load_2d_2cta(sA + s * Cfg::kABytes + t * kSubTileBytes, &tmap_a, a_full(s), kb * kBK,
                           row0 + t * 256);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(256, kTileN, 0, kDx ? 0 : 1);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int sa = kb % Cfg::kStagesA, pa = (kb / Cfg::kStagesA) & 1;
        const int sb = kb % Cfg::kStagesB, pb = (kb / Cfg::kStagesB) & 1;
        mbar_wait_cluster(a_full(sa), pa);
        mbar_wait_cluster(b_full(sb), pb);
        tc_fence_after_sync();
        const uint32_t a_addr = sA + sa * Cfg::kABytes;
        const uint32_t b_addr = sB + sb * Cfg::kBBytes;
#pragma unroll
        for (int ks = 0; ks < kBK / 16; ++ks) {
          auto b_desc_at = [&](uint32_t base) {
            return kDx ? smem_desc_sw128(base + ks * 32, 16, 1024) : smem_desc_sw128(base + ks * 2048, 8192, 1024);
          };
          const uint64_t b_desc = b_desc_at(b_addr);
#pragma unroll
          for (int t = 0; t < MT2; ++t) {
            const uint64_t a_desc = smem_desc_sw128(a_addr + t * kSubTileBytes + ks * 32, 16, 1024);
            mma_bf16_ss_2cta(tmem_base + t * kTileN, a_desc, b_desc, idesc, (kb | ks) != 0);
          }
          if (kSplit) {   // the mean tiles of both CTAs into the same accumulators
            const uint64_t m_desc = b_desc_at(b_addr + kSubTileBytes);
#pragma unroll
            for (int t = 0; t < MT2; ++t) {
              const uint64_t a_desc = smem_desc_sw128(a_addr + t * kSubTileBytes + ks * 32, 16, 1024);
              mma_bf16_ss_2cta(tmem_base + t * kTileN, a_desc, m_desc, idesc, 1u);
            }
          }
        }
        mma_commit_2cta(a_empty(sa));
        mma_commit_2cta(b_empty(sb));
      }
      mma_commit_2cta(acc_full);
    }
    __syncwarp();
  } else {
    // ===================== W-tile generators (both CTAs, own 128 columns) =====================
    const int gw = warp - 2;  // 0..15
    const Sampler sw = live(p.sw);
    WGen<kDx, kInjected, false, kSplit> gen;
    gen.init(gw, lane, nb0);
    const bool n_full = p.w_aligned && (kDx ? (nb0 + 128 <= p.w_in) : (nb0 + 128 <= p.w_out));
    const int k_extent = kDx ? p.w_out : p.w_in;
    auto load_block = [&](int kb) {
      if (n_full && (kb + 1) * kBK <= k_extent) gen.template load<true>(p, mc, kb * kBK);
      else gen.template load<false>(p, mc, kb * kBK);
    };
    gen.words(p, sw, mc, 0);
    load_block(0);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % Cfg::kStagesB, ph = (kb / Cfg::kStagesB) & 1;
      gen.noise();
      gen.words(p, sw, mc, (kb + 1) * kBK);
      gen.pin();
      if (kb > 0) {   // deferred publish of the previous k-block (see the 1-CTA kernel)
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(b_full((kb - 1) % Cfg::kStagesB), 0);   // the leader's barrier
      }
      mbar_wait(b_empty(s), ph ^ 1);
      gen.store(sB + s * Cfg::kBBytes, lane);
      if (kb + 1 < num_kb) load_block(kb + 1);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(b_full((num_kb - 1) % Cfg::kStagesB), 0);
    // ===================== epilogue: this CTA's 128 rows of every sub-tile, all 256 columns ==========
    const int q = warp & 3;
    mbar_wait(acc_full, 0);
    tc_fence_after_sync();
#pragma unroll 1
    for (int t = 0; t < MT2; ++t) {
      const int row = m0 + t * 256 + (int)rank * 128 + q * 32 + lane;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int col = (gw >> 2) * 64 + cc * 32;
        uint4 mk[4];
        if (p.mask_src) {   // chained dX: this lane's 64 bytes of the bf16 activation, in flight during the TMEM load
          const __nv_bfloat16* src = p.mask_src + ((int64_t)mc * p.rows + row) * p.ndim + n0 + col;
#pragma unroll
          for (int c = 0; c < 4; ++c)
            mk[c] = (row < p.rows && n0 + col + 8 * c < p.ndim) ? __ldg(reinterpret_cast<const uint4*>(src) + c)
                                                                : make_uint4(0, 0, 0, 0);
        }
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * kTileN + col), v);
        tmem_ld_wait();
        if (row < p.rows) {
          if (p.out_hw == 0) {
            const int64_t obase = ((int64_t)mc * p.rows + row) * p.ndim + n0 + col;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (n0 + col + j < p.ndim) {   // ndim % 8 == 0 on this path
                float w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) w[e] = __uint_as_float(v[j + e]) + bias_s[col + j + e];
                if (p.act) {
#pragma unroll
                  for (int e = 0; e < 8; ++e) w[e] = w[e] > 0.f ? w[e] : p.act_slope * w[e];
                }
                if (p.mask_src) {
                  const uint4 mkj = mk[j >> 3];
                  const uint32_t mw[4] = {mkj.x, mkj.y, mkj.z, mkj.w};
#pragma unroll
                  for (int e = 0; e < 8; ++e) {
                    const uint32_t h = (e & 1) ? (mw[e >> 1] >> 16) : (mw[e >> 1] & 0xffffu);
                    const bool pos = (h & 0x8000u) == 0 && (h & 0x7fffu) != 0;
                    w[e] = pos ? w[e] : p.mask_slope * w[e];
                  }
                }
                if (p.out) {
                  *reinterpret_cast<float4*>(p.out + obase + j) = make_float4(w[0], w[1], w[2], w[3]);
                  *reinterpret_cast<float4*>(p.out + obase + j + 4) = make_float4(w[4], w[5], w[6], w[7]);
                }
                if (p.out_lp)
                  *reinterpret_cast<uint4*>(p.out_lp + obase + j) =
                      make_uint4(pack_bf16x2(w[0], w[1]), pack_bf16x2(w[2], w[3]), pack_bf16x2(w[4], w[5]),
                                 pack_bf16x2(w[6], w[7]));
              }
            }
          } else {
            const int img = row / p.out_hw, pix = row - img * p.out_hw;
            float* o = p.out + (((int64_t)mc * (p.rows / p.out_hw) + img) * p.ndim + n0 + col) * p.out_hw + pix;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + col + j < p.ndim) o[(int64_t)j * p.out_hw] = __uint_as_float(v[j]) + bias_s[col + j];
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();     // the peer's smem / TMEM must outlive the leader's MMAs
  if (warp == 1) tmem_dealloc_2cta(tmem_base, Cfg::kTmemCols);
}

// ---------------------------------------------------------------------------------------------
// fused weight-gradient kernel: one CTA owns a [128 in-rows x 128 out-cols] tile of the weight and
// walks over all MC samples.  Per sample, dW[mc] = X[mc]^T . dY[mc] (K = batch) is accumulated in
// TMEM by tcgen05.mma (both operands MN-major straight from the [batch, features] activations via
// TMA), the accumulator is double-buffered, and the epilogue warps regenerate eps[mc] for the tile
// from the Philox counter and fold the tile into register accumulators
//     acc_mu += dW[mc]        acc_rho += dW[mc] * eps[mc]
// while the tensor core already works on sample mc+1.  dW [MC,I,O] never exists in HBM.
// ---------------------------------------------------------------------------------------------
constexpr int kDwStages = 6;
constexpr int kDwStageBytes = 2 * kSubTileBytes;  // A [128 i x 64 b] + B [128 o x 64 b], bf16
constexpr int kDwSmemBytes = kDwStages * kDwStageBytes + 1024 + 1024;

struct TcDwParams {
  int mc, batch, w_in, w_out;
  int mc_per_split;        // MC samples per split
  int kb_splits;           // blockIdx.z = mc_split * kb_splits + kb_split
  int kb_per_split;        // reduction k-blocks (64 rows each) per split
  int x_broadcast;         // input has no MC axis
  int atomic_out;          // several MC splits add into zero-initialised outputs
  int dy_broadcast;        // the second operand has no MC axis (conv: im2col of an MC-broadcast input)
  int w_aligned;           // w_out % 4 == 0: 128-bit eps / rho loads and gradient stores
  const float* w_rho;      // [I, O]; NULL: leave drho un-scaled (sum dW*eps; the caller applies sigmoid(rho) later)
  const float* eps_w;      // optional injected eps [mc, I, O]
  Sampler sw;
  float* dw_mu;
  float* dw_rho;
  // MC-sharded training over peer memory (csrc/comm.cu): instead of writing [I, O] gradients locally, the tile goes straight
  // into the staging rows of the ranks that own its elements (reduce-scatter fused into the dW epilogue); sc_world == 0: off
  void* sc_peers[8];       // arena base of every rank
  int64_t sc_mu_off, sc_rho_off;   // byte offsets of this weight's mean / rho staging [world][chunk] in the arenas
  int64_t sc_chunk;        // elements per rank
  int sc_rank, sc_world, sc_bf16;
};

template <bool kInjected, bool kTf32>
__global__ void __launch_bounds__(kTcThreads, 1)
dw_tc(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
      const TcDwParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sStage = smem_base;
  const uint32_t bar0 = sStage + kDwStages * kDwStageBytes;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (kDwStages + s); };
  auto acc_full = [&](int b) { return bar0 + 8u * (2 * kDwStages + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (2 * kDwStages + 2 + b); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * kDwStages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;  // warp-uniform role index
  const int o0 = blockIdx.x * 128, i0 = blockIdx.y * 128;
  const int mc_begin = ((int)blockIdx.z / p.kb_splits) * p.mc_per_split;
  const int mc_end = min(p.mc, mc_begin + p.mc_per_split);
  const int kb_begin = ((int)blockIdx.z % p.kb_splits) * p.kb_per_split;
  constexpr int kBKe = kTf32 ? 32 : 64;   // reduction rows per stage (one 16 KB sub-tile per operand)
  const int kb_end = min((p.batch + kBKe - 1) / kBKe, kb_begin + p.kb_per_split);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDwStages; ++s) {
      mbar_init(full(s), 1);
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), kGenWarps);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_dy);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 256);
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int it = 0;
      for (int mc = mc_begin; mc < mc_end; ++mc) {
        const int xm = p.x_broadcast ? 0 : mc;
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % kDwStages, ph = (it / kDwStages) & 1;
          mbar_wait(empty(s), ph ^ 1);
          mbar_expect_tx(full(s), kDwStageBytes);
          const uint32_t dst = sStage + s * kDwStageBytes;
          // MN-major operand sub-tiles of [kBKe batch rows] x [128 B of features]: two 64-feature
          // boxes (bf16) or four 32-feature boxes (tf32) per operand
          const int ym = p.dy_broadcast ? 0 : mc;
          constexpr int kBoxes = kTf32 ? 4 : 2, kFeat = kTf32 ? 32 : 64, kBoxBytes = kSubTileBytes / kBoxes;
#pragma unroll
          for (int g = 0; g < kBoxes; ++g) {
            tma_load_3d(dst + g * kBoxBytes, &tmap_x, full(s), i0 + g * kFeat, kb * kBKe, xm);
            tma_load_3d(dst + kSubTileBytes + g * kBoxBytes, &tmap_dy, full(s), o0 + g * kFeat, kb * kBKe, ym);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_for<kTf32>(128, 128, 1, 1);
      int it = 0;
      for (int mc = mc_begin, n = 0; mc < mc_end; ++mc, ++n) {
        const int buf = n & 1;
        mbar_wait(acc_empty(buf), ((n >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        for (int kb = kb_begin; kb < kb_end; ++kb, ++it) {
          const int s = it % kDwStages, ph = (it / kDwStages) & 1;
          mbar_wait(full(s), ph);
          tc_fence_after_sync();
          const uint32_t a_addr = sStage + s * kDwStageBytes, b_addr = a_addr + kSubTileBytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            // 16 batch rows = two 8-row swizzle groups = 2048 B; 64-feature groups 8192 B apart
            // (tf32: 8 batch rows = one group = 1024 B per MMA; 32-feature groups 4096 B apart)
            constexpr uint32_t kStep = kTf32 ? 1024 : 2048;
            const uint64_t a_desc = kTf32 ? smem_desc_sw128_base32(a_addr + ks * kStep, 4096, 512)
                                          : smem_desc_sw128(a_addr + ks * kStep, 8192, 1024);
            const uint64_t b_desc = kTf32 ? smem_desc_sw128_base32(b_addr + ks * kStep, 4096, 512)
                                          : smem_desc_sw128(b_addr + ks * kStep, 8192, 1024);
            mma_ss<kTf32>(tmem_base + buf * 128, a_desc, b_desc, idesc, (kb > kb_begin) || ks != 0);
          }
          mma_commit(empty(s));
        }
        mma_commit(acc_full(buf));
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue: eps regeneration + MC reduction =====================
    const int gw = warp - 2, q = warp & 3, h = gw >> 2;
    const Sampler sw = live(p.sw);
    constexpr int kCols = 128 / (kGenWarps / 4);   // weight columns per thread (32)
    const int row = i0 + q * 32 + lane;          // weight row (in-feature) this thread owns
    const int col0 = o0 + h * kCols;             // first of its weight columns
    float acc_mu[kCols], acc_rho[kCols];
#pragma unroll
    for (int j = 0; j < kCols; ++j) acc_mu[j] = acc_rho[j] = 0.f;
    for (int mc = mc_begin, n = 0; mc < mc_end; ++mc, ++n) {
      const int buf = n & 1;
      mbar_wait(acc_full(buf), (n >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll
      for (int c = 0; c < kCols / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128 + h * kCols + c * 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int col = col0 + c * 16 + g * 8;
          float e[8];
          if (kInjected) {
#pragma unroll
            for (int t = 0; t < 8; ++t)
              e[t] = (row < p.w_in && col + t < p.w_out)
                         ? __ldg(p.eps_w + ((int64_t)mc * p.w_in + row) * p.w_out + col + t)
                         : 0.f;
          } else {
            eps8(sw, mc, row, col >> 3, e);
          }
          const int j = c * 16 + g * 8;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float d = __uint_as_float(v[g * 8 + t]);
            acc_mu[j + t] += d;
            acc_rho[j + t] = fmaf(d, e[t], acc_rho[j + t]);
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty(buf));
    }
    if (p.sc_world > 0) {
      // ---- reduce-scatter from the epilogue: tile -> shared memory -> full 256 / 512-byte rows into the owners' staging ----
      // (every MMA has retired -- the last acc_full was waited for -- so the operand stages are free to reuse; a thread's
      // own 32 columns would be 16-byte pieces 8 KB apart: partial-sector stores make poor NVLink packets)
      const int esz = p.sc_bf16 ? 2 : 4;
      const int pitch = 128 * esz + 16;                      // padded row: the 16-byte pieces of a warp spread over the banks
      uint8_t* t_mu = smem + (sStage - smem_base);
      uint8_t* t_rho = t_mu + 128 * pitch;
      const int rl = q * 32 + lane;                          // row inside the tile
#pragma unroll
      for (int j = 0; j < kCols; j += 8) {
        const int cl = h * kCols + j;                        // column inside the tile
        if (p.sc_bf16) {
          *reinterpret_cast<uint4*>(t_mu + rl * pitch + cl * 2) =
              make_uint4(pack_bf16x2(acc_mu[j], acc_mu[j + 1]), pack_bf16x2(acc_mu[j + 2], acc_mu[j + 3]),
                         pack_bf16x2(acc_mu[j + 4], acc_mu[j + 5]), pack_bf16x2(acc_mu[j + 6], acc_mu[j + 7]));
          *reinterpret_cast<uint4*>(t_rho + rl * pitch + cl * 2) =
              make_uint4(pack_bf16x2(acc_rho[j], acc_rho[j + 1]), pack_bf16x2(acc_rho[j + 2], acc_rho[j + 3]),
                         pack_bf16x2(acc_rho[j + 4], acc_rho[j + 5]), pack_bf16x2(acc_rho[j + 6], acc_rho[j + 7]));
        } else {
#pragma unroll
          for (int t = 0; t < 8; t += 4) {
            *reinterpret_cast<float4*>(t_mu + rl * pitch + (cl + t) * 4) =
                make_float4(acc_mu[j + t], acc_mu[j + t + 1], acc_mu[j + t + 2], acc_mu[j + t + 3]);
            *reinterpret_cast<float4*>(t_rho + rl * pitch + (cl + t) * 4) =
                make_float4(acc_rho[j + t], acc_rho[j + t + 1], acc_rho[j + t + 2], acc_rho[j + t + 3]);
          }
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kGenWarps * 32) : "memory");     // the 16 epilogue warps only
      const int per_piece = 16 / esz;                        // elements per 16-byte piece
      const int ppr = 128 / per_piece;                       // pieces per tile row
      const int et = threadIdx.x - 64;                       // 0 .. 511
      for (int pc = et; pc < 128 * ppr; pc += kGenWarps * 32) {
        const int r = pc / ppr, c = (pc - r * ppr) * per_piece;
        const int grow = i0 + r, gcol = o0 + c;
        if (grow >= p.w_in || gcol >= p.w_out) continue;     // w_out % 8 == 0: a piece is inside or outside
        const int64_t f = (int64_t)grow * p.w_out + gcol;
        const int owner = (int)(f / p.sc_chunk);             // chunk % 8 == 0: a piece never straddles two owners
        const int64_t jj = f - (int64_t)owner * p.sc_chunk + (int64_t)p.sc_rank * p.sc_chunk;
        uint8_t* base = (uint8_t*)p.sc_peers[owner];
        *reinterpret_cast<uint4*>(base + p.sc_mu_off + jj * esz) = *reinterpret_cast<const uint4*>(t_mu + r * pitch + c * esz);
        *reinterpret_cast<uint4*>(base + p.sc_rho_off + jj * esz) = *reinterpret_cast<const uint4*>(t_rho + r * pitch + c * esz);
      }
    } else
    if (row < p.w_in) {
#pragma unroll
      for (int j = 0; j < kCols; j += 4) {
        const int col = col0 + j;
        if (col < p.w_out) {
          const int64_t idx = (int64_t)row * p.w_out + col;
          if (!p.w_aligned) {   // ragged / unaligned rows (conv K = Cin*k*k): element-wise tail
#pragma unroll
            for (int t = 0; t < 4; ++t)
              if (col + t < p.w_out) {
                const float gmv = acc_mu[j + t];
                const float grv = acc_rho[j + t] * (p.w_rho ? sigmoid_f(__ldg(p.w_rho + idx + t)) : 1.f);
                if (p.atomic_out) {
                  atomicAdd(p.dw_mu + idx + t, gmv);
                  atomicAdd(p.dw_rho + idx + t, grv);
                } else {
                  p.dw_mu[idx + t] = gmv;
                  p.dw_rho[idx + t] = grv;
                }
              }
            continue;
          }
          float4 gm = make_float4(acc_mu[j], acc_mu[j + 1], acc_mu[j + 2], acc_mu[j + 3]);
          float4 gr = make_float4(acc_rho[j], acc_rho[j + 1], acc_rho[j + 2], acc_rho[j + 3]);
          if (p.w_rho) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.w_rho + idx));
            gr.x *= sigmoid_f(r.x); gr.y *= sigmoid_f(r.y); gr.z *= sigmoid_f(r.z); gr.w *= sigmoid_f(r.w);
          }
          if (p.atomic_out) {
            atomicAdd(p.dw_mu + idx + 0, gm.x); atomicAdd(p.dw_mu + idx + 1, gm.y);
            atomicAdd(p.dw_mu + idx + 2, gm.z); atomicAdd(p.dw_mu + idx + 3, gm.w);
            atomicAdd(p.dw_rho + idx + 0, gr.x); atomicAdd(p.dw_rho + idx + 1, gr.y);
            atomicAdd(p.dw_rho + idx + 2, gr.z); atomicAdd(p.dw_rho + idx + 3, gr.w);
          } else {
            *reinterpret_cast<float4*>(p.dw_mu + idx) = gm;
            *reinterpret_cast<float4*>(p.dw_rho + idx) = gr;
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------
// CTA-pair version of dw_tc (cta_group::2): the pair owns a [256 in-rows x 128 out-cols] weight tile.  Each CTA stages
// its own 128 in-features of X (16 KB per 64 batch rows) and HALF of the dY tile (64 out-features, 8 KB): 24 KB of operands
// per CTA and k-block instead of 32 KB -- dw_tc is bound by what the L2 -> SM fabric delivers (12 TB/s chip-wide), so
// fewer operand bytes per FLOP is the only lever.  The leader's MMA thread issues M = 256 x N = 128 MMAs that read both
// CTAs' shared memory; every CTA gets its 128 accumulator rows in its own TMEM and runs the same epilogue as dw_tc on
// them (eps regeneration, MC sums in registers), so the register budget is unchanged.
//   full barriers live in the leader (both CTAs' TMA bytes are credited to it), empty / acc_full are multicast commits,
//   acc_empty collects the arrivals of both CTAs' epilogue warps at the leader.
// ---------------------------------------------------------------------------------------------
constexpr int kDw2Stages = 8;
constexpr int kDw2StageBytes = kSubTileBytes + kSubTileBytes / 2;   // A [128 i x 64 b] + B [64 o x 64 b], bf16
constexpr int kDw2SmemBytes = kDw2Stages * kDw2StageBytes + 1024 + 1024;

template <bool kInjected>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
dw_tc2(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy, const TcDwParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sStage = smem_base;
  const uint32_t bar0 = sStage + kDw2Stages * kDw2StageBytes;
  auto full = [&](int s) { return bar0 + 8u * s; };
  auto empty = [&](int s) { return bar0 + 8u * (kDw2Stages + s); };
  auto acc_full = [&](int b) { return bar0 + 8u * (2 * kDw2Stages + b); };
  auto acc_empty = [&](int b) { return bar0 + 8u * (2 * kDw2Stages + 2 + b); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * kDw2Stages + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(smem + (tmem_slot - smem_base));

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();                   // 0 = leader
  const int i0 = ((int)blockIdx.x >> 1) * 256 + (int)rank * 128;   // this CTA's 128 weight rows (in-features)
  const int o0 = blockIdx.y * 128;                           // the pair's 128 weight columns
  const int ob0 = o0 + (int)rank * 64;                       // the 64 columns of dY this CTA stages
  const int mc_begin = (int)blockIdx.z * p.mc_per_split;
  const int mc_end = min(p.mc, mc_begin + p.mc_per_split);
  const int num_kb = (p.batch + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kDw2Stages; ++s) {
      mbar_init(full(s), 2);                                 // leader producer (+tx of both CTAs) and peer producer
      mbar_init(empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full(b), 1);
      mbar_init(acc_empty(b), 2 * kGenWarps);                // the epilogue warps of both CTAs
    }
    fence_mbar_init();
    tma_prefetch_desc(&tmap_x);
    tma_prefetch_desc(&tmap_dy);
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 256);
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after_sync();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int it = 0;
      for (int mc = mc_begin; mc < mc_end; ++mc) {
        const int xm = p.x_broadcast ? 0 : mc;
        const int ym = p.dy_broadcast ? 0 : mc;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kDw2Stages, ph = (it / kDw2Stages) & 1;
          mbar_wait(empty(s), ph ^ 1);
          if (rank == 0) mbar_expect_tx(full(s), 2 * kDw2StageBytes);
          else mbar_arrive_cluster(full(s), 0);
          const uint32_t dst = sStage + s * kDw2StageBytes;
#pragma unroll
          for (int g = 0; g < 2; ++g)
            tma_load_3d_2cta(dst + g * (kSubTileBytes / 2), &tmap_x, full(s), i0 + g * 64, kb * kBK, xm);
          tma_load_3d_2cta(dst + kSubTileBytes, &tmap_dy, full(s), ob0, kb * kBK, ym);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader only) =====================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(256, 128, 1, 1);
      int it = 0;
      for (int mc = mc_begin, n = 0; mc < mc_end; ++mc, ++n) {
        const int buf = n & 1;
        mbar_wait_cluster(acc_empty(buf), ((n >> 1) & 1) ^ 1);
        tc_fence_after_sync();
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kDw2Stages, ph = (it / kDw2Stages) & 1;
          mbar_wait_cluster(full(s), ph);
          tc_fence_after_sync();
          const uint32_t a_addr = sStage + s * kDw2StageBytes, b_addr = a_addr + kSubTileBytes;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            // both operands MN-major SW128: 16 batch rows = two 8-row swizzle groups = 2048 B per MMA k-step;
            // A: this CTA's 128 features = two 64-feature groups 8192 B apart; B: this CTA's 64 features = one group
            const uint64_t a_desc = smem_desc_sw128(a_addr + ks * 2048, 8192, 1024);
            const uint64_t b_desc = smem_desc_sw128(b_addr + ks * 2048, 8192, 1024);
            mma_bf16_ss_2cta(tmem_base + buf * 128, a_desc, b_desc, idesc, (kb > 0) || ks != 0);
          }
          mma_commit_2cta(empty(s));
        }
        mma_commit_2cta(acc_full(buf));
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (both CTAs, own 128 rows): eps regeneration + MC reduction =====================
    const int gw = warp - 2, q = warp & 3, h = gw >> 2;
    const Sampler sw = live(p.sw);
    constexpr int kCols = 128 / (kGenWarps / 4);   // weight columns per thread (32)
    const int row = i0 + q * 32 + lane;
    const int col0 = o0 + h * kCols;
    float acc_mu[kCols], acc_rho[kCols];
#pragma unroll
    for (int j = 0; j < kCols; ++j) acc_mu[j] = acc_rho[j] = 0.f;
    for (int mc = mc_begin, n = 0; mc < mc_end; ++mc, ++n) {
      const int buf = n & 1;
      mbar_wait(acc_full(buf), (n >> 1) & 1);
      tc_fence_after_sync();
#pragma unroll
      for (int c = 0; c < kCols / 16; ++c) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 128 + h * kCols + c * 16), v);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int col = col0 + c * 16 + g * 8;
          float e[8];
          if (kInjected) {
#pragma unroll
            for (int t = 0; t < 8; ++t)
              e[t] = (row < p.w_in && col + t < p.w_out)
                         ? __ldg(p.eps_w + ((int64_t)mc * p.w_in + row) * p.w_out + col + t)
                         : 0.f;
          } else {
            eps8(sw, mc, row, col >> 3, e);
          }
          const int j = c * 16 + g * 8;
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float d = __uint_as_float(v[g * 8 + t]);
            acc_mu[j + t] += d;
            acc_rho[j + t] = fmaf(d, e[t], acc_rho[j + t]);
          }
        }
      }
      tc_fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(acc_empty(buf), 0);     // the leader's MMA thread owns the accumulator hand-back
    }
    if (row < p.w_in) {
#pragma unroll
      for (int j = 0; j < kCols; j += 4) {
        const int col = col0 + j;
        if (col < p.w_out) {      // w_out % 8 == 0 on this path
          const int64_t idx = (int64_t)row * p.w_out + col;
          float4 gm = make_float4(acc_mu[j], acc_mu[j + 1], acc_mu[j + 2], acc_mu[j + 3]);
          float4 gr = make_float4(acc_rho[j], acc_rho[j + 1], acc_rho[j + 2], acc_rho[j + 3]);
          if (p.w_rho) {
            const float4 r = __ldg(reinterpret_cast<const float4*>(p.w_rho + idx));
            gr.x *= sigmoid_f(r.x); gr.y *= sigmoid_f(r.y); gr.z *= sigmoid_f(r.z); gr.w *= sigmoid_f(r.w);
          }
          if (p.atomic_out) {
            atomicAdd(p.dw_mu + idx + 0, gm.x); atomicAdd(p.dw_mu + idx + 1, gm.y);
            atomicAdd(p.dw_mu + idx + 2, gm.z); atomicAdd(p.dw_mu + idx + 3, gm.w);
            atomicAdd(p.dw_rho + idx + 0, gr.x); atomicAdd(p.dw_rho + idx + 1, gr.y);
            atomicAdd(p.dw_rho + idx + 2, gr.z); atomicAdd(p.dw_rho + idx + 3, gr.w);
          } else {
            *reinterpret_cast<float4*>(p.dw_mu + idx) = gm;
            *reinterpret_cast<float4*>(p.dw_rho + idx) = gr;
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();     // the peer's smem / TMEM must outlive the leader's MMAs
  if (warp == 1) tmem_dealloc_2cta(tmem_base, 256);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 row-major tensor [rows, cols], box [box_rows, 64 cols], SWIZZLE_128B
static int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                        uint32_t box_cols, bool f32);
static int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols,
                          uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_2d(m, base, rows, cols, box_rows, box_cols, false);
}
// 2-D row-major tensor [rows, cols] of bf16 or fp32, box [box_rows, box_cols] (box_cols * esize = 128 B)
static int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                        uint32_t box_cols, bool f32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return PL_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * (f32 ? 4 : 2)};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols);
    return PL_ERR_CUDA;
  }
  return PL_OK;
}

// 3-D bf16 tensor [mc, rows, cols] (cols innermost), box [1, box_rows, box_cols]: out-of-range rows
// of one MC sample are zero-filled instead of running into the next sample
static int make_tmap_3d(CUtensorMap* m, const void* base, uint64_t mc, uint64_t rows, uint64_t cols,
                        uint32_t box_rows, uint32_t box_cols, bool f32);
static int make_tmap_bf16_3d(CUtensorMap* m, const void* base, uint64_t mc, uint64_t rows, uint64_t cols,
                             uint32_t box_rows, uint32_t box_cols) {
  return make_tmap_3d(m, base, mc, rows, cols, box_rows, box_cols, false);
}
static int make_tmap_3d(CUtensorMap* m, const void* base, uint64_t mc, uint64_t rows, uint64_t cols,
                        uint32_t box_rows, uint32_t box_cols, bool f32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return PL_ERR_CUDA;
  }
  const uint64_t es = f32 ? 4 : 2;
  const cuuint64_t dims[3] = {cols, rows, mc};
  const cuuint64_t strides[2] = {cols * es, rows * cols * es};
  const cuuint32_t box[3] = {box_cols, box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  // fp32 here means an MN-major tf32 MMA operand: the 32-byte-atom flavour of the 128-byte swizzle
  CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                  const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  f32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed with CUresult %d", (int)r);
    return PL_ERR_CUDA;
  }
  return PL_OK;
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

static inline int ew_grid(int64_t n, int per_block) {
  int64_t g = (n + per_block - 1) / per_block;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

// parameters in the form the fused kernels read: bf16 -> packed {mu|sigma} groups (pack_bytes), tf32 -> sigma f32
static inline int prep_params(const float* mu, const float* rho, float* buf, int64_t rows, int64_t cols,
                              bool tf32, cudaStream_t st) {
  if (tf32) {
    softplus_kernel<<<ew_grid(rows * cols, 1024), 256, 0, st>>>(rho, buf, rows * cols);
    PL_LAUNCH_CHECK("softplus_kernel");
  } else {
    pack_mu_sigma_kernel<<<ew_grid(rows * ((cols + 7) / 8), 512), 256, 0, st>>>(mu, rho, (uint4*)buf, (int)rows,
                                                                                  (int)cols);
    PL_LAUNCH_CHECK("pack_mu_sigma_kernel");
  }
  return PL_OK;
}

// Kernel choice between one CTA per tile and CTA pairs (cta_group::2): see launch_gemm (PL_TC_2CTA=1 selects the pairs).
static int g_tc_pairs = -2;     // -1: default policy, 0: never, 1: always (PL_TC_2CTA)

template <bool kDx, bool kInj, bool kTf32 = false, bool kSplit = false>
static int launch_gemm(const CUtensorMap& tm, const TcGemmParams& p, cudaStream_t st, const CUtensorMap* tm_b = nullptr) {
  if (g_tc_pairs == -2) {
    const char* e = getenv("PL_TC_2CTA");
    g_tc_pairs = !e ? -1 : (e[0] == '1' ? 1 : 0);
  }
  if (kSplit && !tm_b) {
    set_error("launch_gemm: the split-operand kernel needs the mean operand's tensor map");
    return PL_ERR_BAD_ARG;
  }
  // CTA pairs (cta_group::2, M = 256 x N = 256 MMAs) halve the shared-memory operand traffic per MMA.  Measured equal to
  // the 1-CTA kernels for both operand forms (r02 C4: 11.50 vs 11.59 ms/step, and slower on the 1000-wide layer): the
  // fused kernels run at the board's power cap, where time follows the executed work, not its schedule.  Opt-in
  // (PL_TC_2CTA=1), parity-tested.
  const bool pair = g_tc_pairs == 1 && !p.mean_only && !p.noise_only;   // the hoisted-mean launches exist in the 1-CTA kernel only
  if (!kTf32 && pair && p.rows > 128 && p.ndim > 128 && (p.out_hw == 0 || (p.out && !p.out_lp))) {
    const int pair_rows = p.rows > 256 ? 512 : 256;
    dim3 grid(2u * (unsigned)((p.ndim + 255) / 256), (unsigned)((p.rows + pair_rows - 1) / pair_rows),
              (unsigned)p.mc);
    const CUtensorMap& tmb2 = tm_b ? *tm_b : tm;
    if (pair_rows == 512) {
      auto kern = sampled_gemm_tc2<2, kDx, kInj, kSplit>;
      PL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc2Cfg<2, kSplit>::kSmemBytes));
      kern<<<grid, kTcThreads, Tc2Cfg<2, kSplit>::kSmemBytes, st>>>(tm, tmb2, p);
    } else {
      auto kern = sampled_gemm_tc2<1, kDx, kInj, kSplit>;
      PL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Tc2Cfg<1, kSplit>::kSmemBytes));
      kern<<<grid, kTcThreads, Tc2Cfg<1, kSplit>::kSmemBytes, st>>>(tm, tmb2, p);
    }
    PL_LAUNCH_CHECK("sampled_gemm_tc2");
    return PL_OK;
  }
  // the hoisted mean launch has a single "sample": 128-row tiles give it 4x the CTAs (it is latency bound per CTA)
  const int mt_rows = p.mean_only ? 128 : (p.rows > 256 ? 512 : (p.rows > 128 ? 256 : 128));
  dim3 grid((unsigned)((p.ndim + kBN - 1) / kBN), (unsigned)((p.rows + mt_rows - 1) / mt_rows),
            (unsigned)p.mc);
  const CUtensorMap& tmb = tm_b ? *tm_b : tm;
#define PL_LAUNCH_MT(MT)                                                                            \
  do {                                                                                              \
    auto kern = sampled_gemm_tc<MT, kDx, kInj, kTf32, kSplit>;                                      \
    PL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,                 \
                                 TcCfg<MT, kSplit>::kSmemBytes));                                   \
    kern<<<grid, kTcThreads, TcCfg<MT, kSplit>::kSmemBytes, st>>>(tm, tmb, p);                       \
  } while (0)
  if (mt_rows == 512) PL_LAUNCH_MT(4);
  else if (mt_rows == 256) PL_LAUNCH_MT(2);
  else PL_LAUNCH_MT(1);
#undef PL_LAUNCH_MT
  PL_LAUNCH_CHECK("sampled_gemm_tc");
  return PL_OK;
}

// tensor map of the mean operand of the split kernels over mu_rounded [rows = I][pitch] (bf16, pitch = O rounded up to
// 8; or fp32 with tf32-rounded values, pitch = O): forward reads MN-major boxes [kBKe weight rows] x [128 B of columns],
// dX reads one K-major box [128 weight rows] x [128 B of columns]
static int make_tmap_mean(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t pitch, bool dx,
                          bool f32) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return PL_ERR_CUDA;
  }
  const uint64_t es = f32 ? 4 : 2;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {pitch * es};
  const cuuint32_t box[2] = {(cuuint32_t)(f32 ? 32 : 64), (cuuint32_t)(dx ? 128 : (f32 ? 32 : 64))};
  const cuuint32_t estr[2] = {1, 1};
  // MN-major fp32 (tf32) operands use the 32-byte-atom flavour of the 128-byte swizzle (see make_tmap_3d)
  const CUtensorMapSwizzle sw = (f32 && !dx) ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(mean operand) failed with CUresult %d (rows %llu cols %llu)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols);
    return PL_ERR_CUDA;
  }
  return PL_OK;
}

// split-operand parameters: bf16 -> mu16 | sig16 (split16_bytes each), tf32 -> mu_hi | sigma (fp32 each)
static inline int prep_params_split(const float* mu, const float* rho, void* mean_buf, void* sig_buf, int64_t rows,
                                    int64_t cols, bool tf32, cudaStream_t st) {
  if (tf32) {
    split_tf32_kernel<<<ew_grid(rows * cols, 1024), 256, 0, st>>>(mu, rho, (float*)mean_buf, (float*)sig_buf, rows * cols);
    PL_LAUNCH_CHECK("split_tf32_kernel");
  } else {
    split_bf16_kernel<<<ew_grid(rows * ((cols + 7) / 8), 512), 256, 0, st>>>(mu, rho, (uint4*)mean_buf, (uint4*)sig_buf,
                                                                              (int)rows, (int)cols);
    PL_LAUNCH_CHECK("split_bf16_kernel");
  }
  return PL_OK;
}


}  // namespace pl
