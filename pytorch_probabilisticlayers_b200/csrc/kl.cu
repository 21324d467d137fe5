// KL(q||p) + entropy reduction and its analytic backward: HBM-bound streaming kernels.
//
// Replaces, per layer and per forward, ~10 ATen pointwise kernels + a reduce + two host syncs
// (torch.distributions.Normal argument validation) at src/ProbabilisticLayers.py:955-956/:1059.
//
// Roofline: HBM.  Algorithmic bytes: forward 8*n (mu, rho fp32; +4*n with a per-element prior),
// backward 16*n (24*n when accumulating).  One pass, 128-bit loads, 4 independent loads per thread
// in flight, grid = 4 CTAs x 148 SMs, warp-shuffle + one block tree reduction, fp64 across blocks
// in a fixed order (deterministic; the block that draws the last ticket finishes the tree, so the
// reduction is a single launch).
#include <math.h>

#include "common.cuh"

namespace pl {

constexpr int kKlThreads = 256;
#ifndef PL_KL_BLOCKS_PER_SM
#define PL_KL_BLOCKS_PER_SM 4
#endif
#ifndef PL_KL_UNROLL
#define PL_KL_UNROLL 2
#endif
constexpr int kKlBlocks = 148 * PL_KL_BLOCKS_PER_SM;
constexpr int kKlUnroll = PL_KL_UNROLL;  // float4 loads per array per thread per pipeline stage

// sigma = softplus(rho) and log(sigma) with MUFU-rate math that stays accurate for sigma ~ 1e-6
// (fresh-init rho ~ -12): series for log1p(e) below e = 1/16, lg2 above.
__device__ __forceinline__ void sigma_logsigma(float rho, float& sigma, float& logsigma, float& e) {
  e = __expf(fminf(rho, 20.f));
  const float poly = e * fmaf(e, fmaf(e, fmaf(e, fmaf(e, 0.2f, -0.25f), 0.33333334f), -0.5f), 1.0f);
  const float big = __logf(1.0f + e);
  sigma = e < 0.0625f ? poly : big;
  sigma = rho > 20.f ? rho : sigma;
  logsigma = __logf(sigma);
}

template <bool kElemPrior>
__device__ __forceinline__ void kl_elem(float mu, float rho, float ip2, float ps, float& acc_a,
                                        float& acc_l, float& acc_p) {
  float sigma, ls, e;
  sigma_logsigma(rho, sigma, ls, e);
  if (kElemPrior) {
    ip2 = __fdividef(1.f, ps * ps);
    acc_p += __logf(ps);
  }
  acc_a = fmaf(0.5f * ip2, fmaf(sigma, sigma, mu * mu), acc_a);
  acc_l += ls;
}

// Small-sigma regime (e = exp(rho) < 1/16, i.e. sigma < 0.06: every freshly initialised layer and
// most of training): sigma = e*P(e) and log sigma = rho + Q(e) with P, Q the series of log1p(e)/e
// and log(log1p(e)/e) -- ONE MUFU (ex2) per element instead of three, no selects.
// |truncation| < 2e-7 relative at e = 1/16.
__device__ __forceinline__ void kl_elem_small(float mu, float rho, float e, float hip2, float& acc_a,
                                              float& acc_l) {
  const float P = fmaf(e, fmaf(e, fmaf(e, fmaf(e, 0.2f, -0.25f), 0.33333334f), -0.5f), 1.0f);
  const float Q = e * fmaf(e, fmaf(e, fmaf(e, fmaf(e, -0.06597222f, 0.08715278f), -0.125f), 0.20833333f), -0.5f);
  const float sigma = e * P;
  acc_a = fmaf(hip2, fmaf(sigma, sigma, mu * mu), acc_a);
  acc_l += rho + Q;
}

// 4 elements: choice between the 1-MUFU series path and the general path
template <bool kElemPrior>
__device__ __forceinline__ void kl_elem4(const float4 m, const float4 r, float ip2, const float4 s,
                                         float& a, float& l, float& p) {
  if (!kElemPrior) {
    const float e0 = __expf(r.x), e1 = __expf(r.y), e2 = __expf(r.z), e3 = __expf(r.w);
    const float mx = fmaxf(fmaxf(e0, e1), fmaxf(e2, e3));
    if (mx < 0.0625f) {   // per thread; warps of a layer in one regime do not diverge
      const float hip2 = 0.5f * ip2;
      kl_elem_small(m.x, r.x, e0, hip2, a, l);
      kl_elem_small(m.y, r.y, e1, hip2, a, l);
      kl_elem_small(m.z, r.z, e2, hip2, a, l);
      kl_elem_small(m.w, r.w, e3, hip2, a, l);
      return;
    }
  }
  kl_elem<kElemPrior>(m.x, r.x, ip2, s.x, a, l, p);
  kl_elem<kElemPrior>(m.y, r.y, ip2, s.y, a, l, p);
  kl_elem<kElemPrior>(m.z, r.z, ip2, s.z, a, l, p);
  kl_elem<kElemPrior>(m.w, r.w, ip2, s.w, a, l, p);
}

__device__ __forceinline__ float4 ldg_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}

// partials[block] = (sum A, sum log sigma, sum log prior_scale_elem) as doubles; the block that
// takes the last ticket reduces the partials in a fixed order (deterministic), writes out2 and
// resets the ticket, so the whole reduction is ONE launch.
template <bool kElemPrior, bool kVec>
__global__ void __launch_bounds__(kKlThreads)
kl_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ rho,
              const float* __restrict__ pse, int64_t n, float ip2, double* __restrict__ partials,
              unsigned int* __restrict__ ticket, double log_ps_total, float* __restrict__ out2) {
  float a = 0.f, l = 0.f, p = 0.f;
  const int64_t tid = (int64_t)blockIdx.x * kKlThreads + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * kKlThreads;
  if (kVec) {
    const int64_t n4 = n >> 2;
    const float4* mu4 = reinterpret_cast<const float4*>(mu);
    const float4* rho4 = reinterpret_cast<const float4*>(rho);
    const float4* ps4 = reinterpret_cast<const float4*>(pse);
    // software pipeline: the loads of batch k+1 are in flight while batch k is reduced
    constexpr int kS = kElemPrior ? kKlUnroll : 1;   // prior-scale registers only when used
    float4 m[kKlUnroll], r[kKlUnroll], s[kS], mn[kKlUnroll], rn[kKlUnroll], sn[kS];
    s[0] = sn[0] = make_float4(1.f, 1.f, 1.f, 1.f);
    auto load = [&](int64_t i0, float4* mm, float4* rr, float4* ss) {
#pragma unroll
      for (int u = 0; u < kKlUnroll; ++u) {
        const int64_t i = i0 + u * nthreads;
        if (i < n4) {
          mm[u] = ldg_stream(mu4 + i);
          rr[u] = ldg_stream(rho4 + i);
          if (kElemPrior) ss[u] = ldg_stream(ps4 + i);
        }
      }
    };
    int64_t i = tid;
    load(i, m, r, s);
    for (; i < n4; i += kKlUnroll * nthreads) {
      const int64_t inext = i + kKlUnroll * nthreads;
      if (inext < n4) load(inext, mn, rn, sn);
#pragma unroll
      for (int u = 0; u < kKlUnroll; ++u)
        if (i + u * nthreads < n4) kl_elem4<kElemPrior>(m[u], r[u], ip2, s[kElemPrior ? u : 0], a, l, p);
#pragma unroll
      for (int u = 0; u < kKlUnroll; ++u) {
        m[u] = mn[u];
        r[u] = rn[u];
        if (kElemPrior) s[u] = sn[u];
      }
    }
    for (int64_t j = (n4 << 2) + tid; j < n; j += nthreads)
      kl_elem<kElemPrior>(mu[j], rho[j], ip2, kElemPrior ? pse[j] : 1.f, a, l, p);
  } else {
    for (int64_t j = tid; j < n; j += nthreads)
      kl_elem<kElemPrior>(mu[j], rho[j], ip2, kElemPrior ? pse[j] : 1.f, a, l, p);
  }
  // warp primitives, then one block-level tree in fp64
  double da = warp_sum((double)a), dl = warp_sum((double)l), dp = warp_sum((double)p);
  __shared__ double sm[3][kKlThreads / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    sm[0][warp] = da;
    sm[1][warp] = dl;
    sm[2][warp] = dp;
  }
  __syncthreads();
  if (warp == 0) {
    da = lane < kKlThreads / 32 ? sm[0][lane] : 0.0;
    dl = lane < kKlThreads / 32 ? sm[1][lane] : 0.0;
    dp = lane < kKlThreads / 32 ? sm[2][lane] : 0.0;
    da = warp_sum(da);
    dl = warp_sum(dl);
    dp = warp_sum(dp);
    if (lane == 0) {
      partials[3 * blockIdx.x + 0] = da;
      partials[3 * blockIdx.x + 1] = dl;
      partials[3 * blockIdx.x + 2] = dp;
    }
  }
  // ---- last block finishes ----
  __shared__ bool is_last;
  if (threadIdx.x == 0) {
    __threadfence();
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  double a2 = 0.0, l2 = 0.0, p2 = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kKlThreads) {
    a2 += __ldcg(partials + 3 * i + 0);
    l2 += __ldcg(partials + 3 * i + 1);
    p2 += __ldcg(partials + 3 * i + 2);
  }
  a2 = warp_sum(a2);
  l2 = warp_sum(l2);
  p2 = warp_sum(p2);
  __syncthreads();
  if (lane == 0) {
    sm[0][warp] = a2;
    sm[1][warp] = l2;
    sm[2][warp] = p2;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a2 = l2 = p2 = 0.0;
    for (int w = 0; w < kKlThreads / 32; ++w) {
      a2 += sm[0][w];
      l2 += sm[1][w];
      p2 += sm[2][w];
    }
    out2[0] = (float)(a2 - l2 + p2 + log_ps_total - 0.5 * (double)n);
    out2[1] = (float)(l2 + (double)n * (0.5 + 0.9189385332046727));
    *ticket = 0;   // leave the workspace ready for the next call
  }
}

template <bool kElemPrior, bool kAccumulate>
__device__ __forceinline__ void kl_bwd_elem(float mu, float rho, float ip2, float ps, float gk,
                                            float ge, float& dmu, float& drho) {
  float sigma, ls, e;
  sigma_logsigma(rho, sigma, ls, e);
  if (kElemPrior) ip2 = __fdividef(1.f, ps * ps);
  const float sig = rho > 20.f ? 1.f : __fdividef(e, 1.f + e);
  const float inv_sigma = __fdividef(1.f, sigma);
  const float gm = gk * mu * ip2;
  const float gr = (gk * fmaf(sigma, ip2, -inv_sigma) + ge * inv_sigma) * sig;
  dmu = kAccumulate ? dmu + gm : gm;
  drho = kAccumulate ? drho + gr : gr;
}

template <bool kElemPrior, bool kAccumulate, bool kVec>
__global__ void __launch_bounds__(kKlThreads)
kl_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ rho,
              const float* __restrict__ pse, int64_t n, float ip2, const float* __restrict__ grad_kl,
              const float* __restrict__ grad_ent, float host_scale, float* __restrict__ dmu,
              float* __restrict__ drho) {
  const float gk = host_scale * (grad_kl ? __ldg(grad_kl) : 1.f);
  const float ge = host_scale * (grad_ent ? __ldg(grad_ent) : 0.f);
  const int64_t tid = (int64_t)blockIdx.x * kKlThreads + threadIdx.x;
  const int64_t nthreads = (int64_t)gridDim.x * kKlThreads;
  if (kVec) {
    const int64_t n4 = n >> 2;
    const float4* mu4 = reinterpret_cast<const float4*>(mu);
    const float4* rho4 = reinterpret_cast<const float4*>(rho);
    const float4* ps4 = reinterpret_cast<const float4*>(pse);
    float4* dmu4 = reinterpret_cast<float4*>(dmu);
    float4* drho4 = reinterpret_cast<float4*>(drho);
    for (int64_t i = tid; i < n4; i += 2 * nthreads) {
      const int64_t i2 = i + nthreads;
      const bool has2 = i2 < n4;
      float4 m0 = ldg_stream(mu4 + i), r0 = ldg_stream(rho4 + i);
      float4 m1 = has2 ? ldg_stream(mu4 + i2) : m0, r1 = has2 ? ldg_stream(rho4 + i2) : r0;
      float4 s0 = make_float4(1.f, 1.f, 1.f, 1.f), s1 = s0, a0, b0, a1, b1;
      if (kElemPrior) {
        s0 = ldg_stream(ps4 + i);
        s1 = has2 ? ldg_stream(ps4 + i2) : s0;
      }
      if (kAccumulate) {
        a0 = dmu4[i];
        b0 = drho4[i];
        a1 = has2 ? dmu4[i2] : a0;
        b1 = has2 ? drho4[i2] : b0;
      }
      kl_bwd_elem<kElemPrior, kAccumulate>(m0.x, r0.x, ip2, s0.x, gk, ge, a0.x, b0.x);
      kl_bwd_elem<kElemPrior, kAccumulate>(m0.y, r0.y, ip2, s0.y, gk, ge, a0.y, b0.y);
      kl_bwd_elem<kElemPrior, kAccumulate>(m0.z, r0.z, ip2, s0.z, gk, ge, a0.z, b0.z);
      kl_bwd_elem<kElemPrior, kAccumulate>(m0.w, r0.w, ip2, s0.w, gk, ge, a0.w, b0.w);
      dmu4[i] = a0;
      drho4[i] = b0;
      if (has2) {
        kl_bwd_elem<kElemPrior, kAccumulate>(m1.x, r1.x, ip2, s1.x, gk, ge, a1.x, b1.x);
        kl_bwd_elem<kElemPrior, kAccumulate>(m1.y, r1.y, ip2, s1.y, gk, ge, a1.y, b1.y);
        kl_bwd_elem<kElemPrior, kAccumulate>(m1.z, r1.z, ip2, s1.z, gk, ge, a1.z, b1.z);
        kl_bwd_elem<kElemPrior, kAccumulate>(m1.w, r1.w, ip2, s1.w, gk, ge, a1.w, b1.w);
        dmu4[i2] = a1;
        drho4[i2] = b1;
      }
    }
    for (int64_t j = (n4 << 2) + tid; j < n; j += nthreads) {
      float a = kAccumulate ? dmu[j] : 0.f, b = kAccumulate ? drho[j] : 0.f;
      kl_bwd_elem<kElemPrior, kAccumulate>(mu[j], rho[j], ip2, kElemPrior ? pse[j] : 1.f, gk, ge, a, b);
      dmu[j] = a;
      drho[j] = b;
    }
  } else {
    for (int64_t j = tid; j < n; j += nthreads) {
      float a = kAccumulate ? dmu[j] : 0.f, b = kAccumulate ? drho[j] : 0.f;
      kl_bwd_elem<kElemPrior, kAccumulate>(mu[j], rho[j], ip2, kElemPrior ? pse[j] : 1.f, gk, ge, a, b);
      dmu[j] = a;
      drho[j] = b;
    }
  }
}


static inline int kl_grid(int64_t n) {
  const int64_t want = ceil_div64(n, (int64_t)kKlThreads * 4);
  return (int)(want < 1 ? 1 : (want > kKlBlocks ? kKlBlocks : want));
}

}  // namespace pl

extern "C" size_t pl_kl_workspace_bytes(void) { return sizeof(double) * 3 * pl::kKlBlocks + 64; }

extern "C" int pl_kl_fwd(const float* mu, const float* rho, int64_t n, float prior_scale,
                         const float* prior_scale_elem, float* out2, void* workspace,
                         size_t workspace_bytes, void* stream) {
  using namespace pl;
  PL_CHECK_ARG(mu && rho && out2 && workspace, "pl_kl_fwd: NULL pointer");
  PL_CHECK_ARG(n >= 0, "pl_kl_fwd: n < 0");
  PL_CHECK_ARG(prior_scale_elem || prior_scale > 0.f, "pl_kl_fwd: prior_scale must be > 0");
  if (workspace_bytes < pl_kl_workspace_bytes()) {
    set_error("pl_kl_fwd: workspace %zu < %zu bytes", workspace_bytes, pl_kl_workspace_bytes());
    return PL_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  double* partials = (double*)workspace;
  unsigned int* ticket = (unsigned int*)((char*)workspace + sizeof(double) * 3 * kKlBlocks);
  const int grid = kl_grid(n);
  const double log_ps_total = prior_scale_elem ? 0.0 : (double)n * log((double)prior_scale);
  const bool vec = aligned16(mu) && aligned16(rho) && (!prior_scale_elem || aligned16(prior_scale_elem));
  const float ip2 = prior_scale_elem ? 1.f : 1.f / (prior_scale * prior_scale);
#define PL_KL_FWD(EP, VEC)                                                                        \
  kl_fwd_kernel<EP, VEC><<<grid, kKlThreads, 0, st>>>(mu, rho, prior_scale_elem, n, ip2, partials, \
                                                      ticket, log_ps_total, out2)
  if (prior_scale_elem) {
    if (vec) PL_KL_FWD(true, true);
    else PL_KL_FWD(true, false);
  } else {
    if (vec) PL_KL_FWD(false, true);
    else PL_KL_FWD(false, false);
  }
#undef PL_KL_FWD
  PL_LAUNCH_CHECK("kl_fwd_kernel");
  return PL_OK;
}

extern "C" int pl_kl_bwd(const float* mu, const float* rho, int64_t n, float prior_scale,
                         const float* prior_scale_elem, const float* grad_kl, const float* grad_ent,
                         float host_scale, float* dmu, float* drho, int32_t accumulate, void* stream) {
  using namespace pl;
  PL_CHECK_ARG(mu && rho && dmu && drho, "pl_kl_bwd: NULL pointer");
  PL_CHECK_ARG(n >= 0, "pl_kl_bwd: n < 0");
  PL_CHECK_ARG(prior_scale_elem || prior_scale > 0.f, "pl_kl_bwd: prior_scale must be > 0");
  if (n == 0) return PL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t want = ceil_div64(n, (int64_t)kKlThreads * 8);
  const int grid = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  const bool vec = aligned16(mu) && aligned16(rho) && aligned16(dmu) && aligned16(drho) &&
                   (!prior_scale_elem || aligned16(prior_scale_elem));
  const float ip2 = prior_scale_elem ? 1.f : 1.f / (prior_scale * prior_scale);
#define PL_KL_BWD(EP, ACC, VEC)                                                               \
  kl_bwd_kernel<EP, ACC, VEC><<<grid, kKlThreads, 0, st>>>(mu, rho, prior_scale_elem, n, ip2, \
                                                           grad_kl, grad_ent, host_scale, dmu, drho)
  const int sel = (prior_scale_elem ? 4 : 0) | (accumulate ? 2 : 0) | (vec ? 1 : 0);
  switch (sel) {
    case 0: PL_KL_BWD(false, false, false); break;
    case 1: PL_KL_BWD(false, false, true); break;
    case 2: PL_KL_BWD(false, true, false); break;
    case 3: PL_KL_BWD(false, true, true); break;
    case 4: PL_KL_BWD(true, false, false); break;
    case 5: PL_KL_BWD(true, false, true); break;
    case 6: PL_KL_BWD(true, true, false); break;
    default: PL_KL_BWD(true, true, true); break;
  }
#undef PL_KL_BWD
  PL_LAUNCH_CHECK("kl_bwd_kernel");
  return PL_OK;
}
