// MC_BatchNorm2D -> LeakyReLU -> MC_MaxPool2d(2, 2) as one HBM-bound pass each way.
//
// Replaces, per conv block of the reference's 'cbnn' net (experiments/BNN.py:232-247), the chain
//   MC_BatchNorm2D.forward (src/ProbabilisticLayers.py:736-762: two permutes + F.batch_norm over [B, C, H, W, MC],
//   i.e. per-channel statistics pooled over MC, B, H, W jointly), LeakyReLU, MC_MaxPool2d.forward (:207-236:
//   permute + flatten + F.max_pool2d + reshape + two .contiguous() copies)
// and their autograd backward.  On a B200 the ATen / cuDNN kernels behind those modules take ~11 ms of the
// 24 ms C3 step (cudnn bn_bw 4.5 ms, max_pool_backward 2.9 ms, ...: profiles/r01_launches_c3_final.csv).
//
// The input is the conv output [n = MC*B, C, H, W] fp32 contiguous, H and W even.  Forward: one statistics pass
// (sum, sum of squares per channel, fp64 across blocks) and one apply pass that normalises, applies LeakyReLU,
// takes the 2x2 maximum and records which of the 4 inputs won (2 bits in a byte).  Backward: one reduction
// pass (sum dz, sum dz*xhat per channel; dz is non-zero only at the winners) and one pass that writes dx.
// Algorithmic bytes (E = n*C*H*W*4): forward 2E read + E/4 + E/16 written; backward ~2E + E/2 read, E written.
#include "common.cuh"

namespace pl {

constexpr int kBnThreads = 256;

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = 0.f;
  if (warp == 0) {
    t = lane < kBnThreads / 32 ? sh[lane] : 0.f;
    t = warp_sum(t);
  }
  return t;   // valid in thread 0
}

// grid (C, chunks): sums[c] += (sum x, sum x^2) over this block's share of the n axis
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float* __restrict__ x, int n, int c_total, int hw,
                                                              double* __restrict__ sums) {
  __shared__ float sh[kBnThreads / 32];
  const int c = blockIdx.x;
  float s = 0.f, q = 0.f;
  // this block's planes are i = blockIdx.y + k * gridDim.y; the (plane, element) pairs are flattened so that
  // small planes (8 x 8, 4 x 4 in the last conv blocks) still keep all 256 threads busy
  const bool vec = (hw & 3) == 0;
  const int per = vec ? hw / 4 : hw;
  const int planes = n > (int)blockIdx.y ? (n - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y : 0;
  const int64_t items = (int64_t)planes * per;
  for (int64_t it = threadIdx.x; it < items; it += kBnThreads) {
    const int k = (int)(it / per), t = (int)(it - (int64_t)k * per);
    const float* p = x + ((int64_t)(blockIdx.y + k * gridDim.y) * c_total + c) * hw;
    if (vec) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p) + t);
      s += (v.x + v.y) + (v.z + v.w);
      q += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    } else {
      const float v = __ldg(p + t);
      s += v;
      q += v * v;
    }
  }
  const float bs = block_sum(s, sh);
  const float bq = block_sum(q, sh);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 2 * c, (double)bs);
    atomicAdd(sums + 2 * c + 1, (double)bq);
  }
}

__device__ __forceinline__ float bn_act(float x, float scale, float shift, float slope) {
  const float z = fmaf(x, scale, shift);
  return z > 0.f ? z : slope * z;
}

// one thread per pooled output; window elements in the order (0,0) (0,1) (1,0) (1,1), first maximum wins
// (torch's max_pool2d scans rows then columns and keeps the first maximum; NaN handling is not reproduced)
__global__ void __launch_bounds__(kBnThreads)
bn_act_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                       const float* __restrict__ mean, const float* __restrict__ rstd, float slope, int64_t total,
                       int c_total, int h, int w, float* __restrict__ y, uint8_t* __restrict__ idx) {
  const int ph_n = h >> 1, pw_n = w >> 1;
  for (int64_t o = (int64_t)blockIdx.x * kBnThreads + threadIdx.x; o < total; o += (int64_t)gridDim.x * kBnThreads) {
    const int pw = (int)(o % pw_n);
    const int64_t r = o / pw_n;
    const int ph = (int)(r % ph_n);
    const int64_t plane = r / ph_n;               // n * C + c
    const int c = (int)(plane % c_total);
    const float scale = __ldg(gamma + c) * __ldg(rstd + c);
    const float shift = fmaf(-__ldg(mean + c), scale, __ldg(beta + c));
    const float* p = x + (plane * h + 2 * ph) * w + 2 * pw;
    const float2 r0 = __ldg(reinterpret_cast<const float2*>(p));
    const float2 r1 = __ldg(reinterpret_cast<const float2*>(p + w));
    const float a0 = bn_act(r0.x, scale, shift, slope), a1 = bn_act(r0.y, scale, shift, slope);
    const float a2 = bn_act(r1.x, scale, shift, slope), a3 = bn_act(r1.y, scale, shift, slope);
    float best = a0;
    int bi = 0;
    if (a1 > best) { best = a1; bi = 1; }
    if (a2 > best) { best = a2; bi = 2; }
    if (a3 > best) { best = a3; bi = 3; }
    y[o] = best;
    idx[o] = (uint8_t)bi;
  }
}

// grid (C, chunks): sums[c] += (sum dz, sum dz * xhat) over the winners of this block's share of the n axis
__global__ void __launch_bounds__(kBnThreads)
bn_act_pool_bwd_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                              const uint8_t* __restrict__ idx, const float* __restrict__ gamma,
                              const float* __restrict__ beta, const float* __restrict__ mean,
                              const float* __restrict__ rstd, float slope, int n, int c_total, int h, int w,
                              double* __restrict__ sums) {
  __shared__ float sh[kBnThreads / 32];
  const int c = blockIdx.x;
  const int ph_n = h >> 1, pw_n = w >> 1, pooled = ph_n * pw_n;
  const float m = __ldg(mean + c), rs = __ldg(rstd + c), g = __ldg(gamma + c), b = __ldg(beta + c);
  float s1 = 0.f, s2 = 0.f;
  const int planes = n > (int)blockIdx.y ? (n - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y : 0;
  const int64_t items = (int64_t)planes * pooled;   // flattened (plane, pooled element) pairs, see bn_stats_kernel
  for (int64_t it = threadIdx.x; it < items; it += kBnThreads) {
    const int k = (int)(it / pooled), t = (int)(it - (int64_t)k * pooled);
    const int64_t plane = (int64_t)(blockIdx.y + k * gridDim.y) * c_total + c;
    const int ph = t / pw_n, pw = t - ph * pw_n;
    const int bi = idx[plane * pooled + t];
    const float xv = __ldg(x + (plane * h + 2 * ph + (bi >> 1)) * w + 2 * pw + (bi & 1));
    const float xhat = (xv - m) * rs;
    const float z = fmaf(g, xhat, b);
    const float dz = __ldg(dy + plane * pooled + t) * (z > 0.f ? 1.f : slope);
    s1 += dz;
    s2 = fmaf(dz, xhat, s2);
  }
  const float b1 = block_sum(s1, sh);
  const float b2 = block_sum(s2, sh);
  if (threadIdx.x == 0) {
    atomicAdd(sums + 2 * c, (double)b1);
    atomicAdd(sums + 2 * c + 1, (double)b2);
  }
}

// one thread per pooled output: dx of its 4 inputs.  batch_stats: dx = gamma*rstd*(dz - mean(dz) - xhat*mean(dz*xhat));
// otherwise (running statistics) dx = gamma*rstd*dz.  Block 0 also writes dgamma / dbeta from the sums.
__global__ void __launch_bounds__(kBnThreads)
bn_act_pool_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ dy, const uint8_t* __restrict__ idx,
                          const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ mean, const float* __restrict__ rstd, float slope,
                          int64_t total, int c_total, int h, int w, const double* __restrict__ sums,
                          double inv_count, int batch_stats, float* __restrict__ dx, float* __restrict__ dgamma,
                          float* __restrict__ dbeta) {
  if (blockIdx.x == 0) {
    for (int c = threadIdx.x; c < c_total; c += kBnThreads) {
      if (dbeta) dbeta[c] = (float)sums[2 * c];
      if (dgamma) dgamma[c] = (float)sums[2 * c + 1];
    }
  }
  if (!dx) return;
  const int ph_n = h >> 1, pw_n = w >> 1;
  for (int64_t o = (int64_t)blockIdx.x * kBnThreads + threadIdx.x; o < total; o += (int64_t)gridDim.x * kBnThreads) {
    const int pw = (int)(o % pw_n);
    const int64_t r = o / pw_n;
    const int ph = (int)(r % ph_n);
    const int64_t plane = r / ph_n;
    const int c = (int)(plane % c_total);
    const float m = __ldg(mean + c), rs = __ldg(rstd + c), g = __ldg(gamma + c), b = __ldg(beta + c);
    const float m1 = batch_stats ? (float)(sums[2 * c] * inv_count) : 0.f;
    const float m2 = batch_stats ? (float)(sums[2 * c + 1] * inv_count) : 0.f;
    const int64_t base = (plane * h + 2 * ph) * w + 2 * pw;
    const float2 r0 = __ldg(reinterpret_cast<const float2*>(x + base));
    const float2 r1 = __ldg(reinterpret_cast<const float2*>(x + base + w));
    const float xh[4] = {(r0.x - m) * rs, (r0.y - m) * rs, (r1.x - m) * rs, (r1.y - m) * rs};
    const int bi = idx[o];
    const float xb = bi == 0 ? xh[0] : (bi == 1 ? xh[1] : (bi == 2 ? xh[2] : xh[3]));
    const float zb = fmaf(g, xb, b);
    const float dzb = __ldg(dy + o) * (zb > 0.f ? 1.f : slope);
    const float k = g * rs;
    float d[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) d[e] = k * ((e == bi ? dzb : 0.f) - m1 - xh[e] * m2);
    *reinterpret_cast<float2*>(dx + base) = make_float2(d[0], d[1]);
    *reinterpret_cast<float2*>(dx + base + w) = make_float2(d[2], d[3]);
  }
}

static int bn_check(const pl_bnpool_args* a, const char* who) {
  if (!a) {
    set_error("%s: NULL args", who);
    return PL_ERR_BAD_ARG;
  }
  if (a->n < 0 || a->channels < 1 || a->h < 2 || a->w < 2 || (a->h & 1) || (a->w & 1)) {
    set_error("%s: needs n >= 0, channels >= 1 and even h, w >= 2 (got n=%d C=%d h=%d w=%d)", who, a->n,
              a->channels, a->h, a->w);
    return PL_ERR_UNSUPPORTED;
  }
  if (a->channels > 65535) {
    set_error("%s: more than 65535 channels", who);
    return PL_ERR_UNSUPPORTED;
  }
  return PL_OK;
}

static inline int chunks_for(int n, int channels) {
  int ch = (148 * 8 + channels - 1) / channels;
  if (ch > n) ch = n;
  if (ch > 65535) ch = 65535;
  return ch < 1 ? 1 : ch;
}

static inline int ew_blocks(int64_t total) {
  int64_t g = (total + kBnThreads - 1) / kBnThreads;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace pl

extern "C" int pl_bn_stats(const float* x, int32_t n, int32_t channels, int32_t hw, double* sums, void* stream) {
  using namespace pl;
  PL_CHECK_ARG(sums && channels >= 1 && channels <= 65535 && hw >= 1 && n >= 0, "pl_bn_stats: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  PL_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)channels, st));
  if (n == 0) return PL_OK;
  PL_CHECK_ARG(x, "pl_bn_stats: NULL pointer");
  dim3 grid((unsigned)channels, (unsigned)chunks_for(n, channels));
  bn_stats_kernel<<<grid, kBnThreads, 0, st>>>(x, n, channels, hw, sums);
  PL_LAUNCH_CHECK("bn_stats_kernel");
  return PL_OK;
}

extern "C" int pl_bn_act_pool_fwd(const pl_bnpool_args* a) {
  using namespace pl;
  int rc = bn_check(a, "pl_bn_act_pool_fwd");
  if (rc != PL_OK) return rc;
  if (a->n == 0) return PL_OK;
  PL_CHECK_ARG(a->x && a->gamma && a->beta && a->mean && a->rstd && a->y && a->idx, "pl_bn_act_pool_fwd: NULL pointer");
  const int64_t total = (int64_t)a->n * a->channels * (a->h / 2) * (a->w / 2);
  bn_act_pool_fwd_kernel<<<ew_blocks(total), kBnThreads, 0, (cudaStream_t)a->stream>>>(
      a->x, a->gamma, a->beta, a->mean, a->rstd, a->slope, total, a->channels, a->h, a->w, a->y, a->idx);
  PL_LAUNCH_CHECK("bn_act_pool_fwd_kernel");
  return PL_OK;
}

extern "C" size_t pl_bn_act_pool_workspace_bytes(int32_t channels) {
  return sizeof(double) * 2 * (size_t)(channels > 0 ? channels : 0) + 64;
}

extern "C" int pl_bn_act_pool_bwd(const pl_bnpool_args* a) {
  using namespace pl;
  int rc = bn_check(a, "pl_bn_act_pool_bwd");
  if (rc != PL_OK) return rc;
  PL_CHECK_ARG(a->workspace && a->workspace_bytes >= pl_bn_act_pool_workspace_bytes(a->channels),
               "pl_bn_act_pool_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)a->stream;
  double* sums = (double*)a->workspace;
  PL_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)a->channels, st));
  if (a->n > 0) {
    PL_CHECK_ARG(a->x && a->gamma && a->beta && a->mean && a->rstd && a->dy && a->idx,
                 "pl_bn_act_pool_bwd: NULL pointer");
    dim3 grid((unsigned)a->channels, (unsigned)chunks_for(a->n, a->channels));
    bn_act_pool_bwd_reduce_kernel<<<grid, kBnThreads, 0, st>>>(a->x, a->dy, a->idx, a->gamma, a->beta, a->mean,
                                                               a->rstd, a->slope, a->n, a->channels, a->h, a->w,
                                                               sums);
    PL_LAUNCH_CHECK("bn_act_pool_bwd_reduce_kernel");
  }
  const int64_t total = (int64_t)a->n * a->channels * (a->h / 2) * (a->w / 2);
  const double count = (double)a->n * a->h * a->w;
  bn_act_pool_bwd_dx_kernel<<<ew_blocks(total > 0 ? total : 1), kBnThreads, 0, st>>>(
      a->x, a->dy, a->idx, a->gamma, a->beta, a->mean, a->rstd, a->slope, total, a->channels, a->h, a->w, sums,
      count > 0 ? 1.0 / count : 0.0, a->batch_stats, a->n > 0 ? a->dx : nullptr, a->dgamma, a->dbeta);
  PL_LAUNCH_CHECK("bn_act_pool_bwd_dx_kernel");
  return PL_OK;
}
