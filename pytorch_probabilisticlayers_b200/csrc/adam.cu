// Fused optimiser pass of the ELBO train step: Adam update + analytic KL gradient + KL value in ONE
// streaming kernel over a parameter tensor.
//
// The KL(q||p) gradient of a factorised Gaussian posterior against N(0, sp) is an elementwise
// function of the parameter itself:  d/dmu = mu/sp^2,  d/drho = (sigma/sp^2 - 1/sigma) sigmoid(rho)
// (src/ProbabilisticLayers.py:955 through torch/distributions/kl.py:468-471), so it can be added to
// the likelihood gradient inside the optimiser pass instead of in separate passes over (mu, rho)
// (pl_kl_fwd: 8 B/element, pl_kl_bwd: 24 B/element).  The KL *value* (logged, never differentiated
// again) is accumulated from the pre-update parameters in the same pass.
//
// Replaces torch.optim.Adam.step (experiments/BNN.py:349) + the KL backward.  HBM-bound:
// 28 B/element (read p, g, m, v; write p, m, v).  Same arithmetic as torch's Adam (no amsgrad, no
// weight decay): step_size = lr/(1-b1^t), denom = sqrt(v)/sqrt(1-b2^t) + eps.
#include <math.h>

#include "common.cuh"

namespace pl {

__device__ __forceinline__ void sigma_parts(float rho, float& sigma, float& logsigma, float& sig) {
  const float e = __expf(fminf(rho, 20.f));
  const float poly = e * fmaf(e, fmaf(e, fmaf(e, fmaf(e, 0.2f, -0.25f), 0.33333334f), -0.5f), 1.0f);
  sigma = e < 0.0625f ? poly : __logf(1.0f + e);
  sigma = rho > 20.f ? rho : sigma;
  logsigma = __logf(sigma);
  sig = rho > 20.f ? 1.f : __fdividef(e, 1.f + e);
}

// kind: 0 = no KL term (bias, hyper-net, ...), 1 = posterior mean, 2 = posterior rho (logscale)
template <int kKind>
__global__ void __launch_bounds__(256)
adam_kl_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
               float* __restrict__ v, int64_t n, float ip2, float kl_scale, float step_size,
               float inv_sqrt_bc2, float beta1, float beta2, float eps, double* __restrict__ kl_out,
               const float* __restrict__ sched_dev) {
  if (sched_dev) {   // CUDA-graph replays: the step-dependent scalars live in device memory
    step_size = __ldg(sched_dev);
    inv_sqrt_bc2 = __ldg(sched_dev + 1);
  }
  float kl = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float w = p[i];
    float grad = g[i];
    if (kKind == 1) {
      grad = fmaf(kl_scale * ip2, w, grad);
      kl = fmaf(0.5f * ip2 * w, w, kl);
    } else if (kKind == 2) {
      float sigma, ls, sig;
      sigma_parts(w, sigma, ls, sig);
      grad = fmaf(kl_scale * sig, fmaf(sigma, ip2, -__fdividef(1.f, sigma)), grad);
      kl += fmaf(0.5f * ip2 * sigma, sigma, -ls);
    }
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * grad);
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * grad * grad);
    m[i] = mi;
    v[i] = vi;
    p[i] = w - step_size * __fdividef(mi, fmaf(sqrtf(vi), inv_sqrt_bc2, eps));
  }
  if (kKind != 0 && kl_out) {
    double d = warp_sum((double)kl);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
      d = 0.0;
      for (int k = 0; k < 8; ++k) d += sm[k];
      atomicAdd(kl_out, d);
    }
  }
}

}  // namespace pl

static int adam_kl_launch(const float* sched_dev, float* p, const float* g, float* m, float* v, int64_t n, int32_t kind,
                               float prior_scale, float kl_scale, float lr, float beta1, float beta2,
                               float eps, int32_t step, double* kl_value, void* stream) {
  using namespace pl;
  PL_CHECK_ARG(p && g && m && v, "pl_adam_kl_step: NULL pointer");
  PL_CHECK_ARG(n >= 0 && (sched_dev || step >= 1), "pl_adam_kl_step: bad n / step");
  PL_CHECK_ARG(kind >= 0 && kind <= 2, "pl_adam_kl_step: kind must be 0, 1 or 2");
  PL_CHECK_ARG(kind == 0 || prior_scale > 0.f, "pl_adam_kl_step: prior_scale must be > 0");
  if (n == 0) return PL_OK;
  if (sched_dev) step = 1;   // host-side scalars unused
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const float ip2 = kind ? 1.f / (prior_scale * prior_scale) : 0.f;
  int64_t blocks = ceil_div64(n, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (kind == 1 && kl_value) {
    // constant part of the KL value of this (mu, rho) pair is added with the rho tensor
  }
  switch (kind) {
    case 0: adam_kl_kernel<0><<<(int)blocks, 256, 0, st>>>(p, g, m, v, n, ip2, kl_scale, step_size, inv_sqrt_bc2, beta1, beta2, eps, kl_value, sched_dev); break;
    case 1: adam_kl_kernel<1><<<(int)blocks, 256, 0, st>>>(p, g, m, v, n, ip2, kl_scale, step_size, inv_sqrt_bc2, beta1, beta2, eps, kl_value, sched_dev); break;
    default: adam_kl_kernel<2><<<(int)blocks, 256, 0, st>>>(p, g, m, v, n, ip2, kl_scale, step_size, inv_sqrt_bc2, beta1, beta2, eps, kl_value, sched_dev); break;
  }
  PL_LAUNCH_CHECK("adam_kl_kernel");
  return PL_OK;
}

extern "C" int pl_adam_kl_step(float* p, const float* g, float* m, float* v, int64_t n, int32_t kind,
                               float prior_scale, float kl_scale, float lr, float beta1, float beta2,
                               float eps, int32_t step, double* kl_value, void* stream) {
  return adam_kl_launch(nullptr, p, g, m, v, n, kind, prior_scale, kl_scale, lr, beta1, beta2, eps, step, kl_value,
                        stream);
}

extern "C" int pl_adam_kl_step_dev(float* p, const float* g, float* m, float* v, int64_t n, int32_t kind,
                                   float prior_scale, float kl_scale, float beta1, float beta2, float eps,
                                   const float* sched_dev, double* kl_value, void* stream) {
  if (!sched_dev) {
    pl::set_error("pl_adam_kl_step_dev: NULL sched_dev");
    return PL_ERR_BAD_ARG;
  }
  return adam_kl_launch(sched_dev, p, g, m, v, n, kind, prior_scale, kl_scale, 1.0f, beta1, beta2, eps, 1, kl_value,
                        stream);
}
