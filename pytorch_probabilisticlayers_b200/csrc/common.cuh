// Shared device/host helpers for the pl_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pl_b200.h"

namespace pl {

// ---------------------------------------------------------------------------------------------
// error plumbing: negative pl_status + thread-local text, never throw / abort
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch();  // bumps the counter behind pl_launch_count()

#define PL_CHECK_ARG(cond, ...)              \
  do {                                       \
    if (!(cond)) {                           \
      ::pl::set_error(__VA_ARGS__);          \
      return PL_ERR_BAD_ARG;                 \
    }                                        \
  } while (0)

#define PL_CUDA(call)                                        \
  do {                                                       \
    cudaError_t e__ = (call);                                \
    if (e__ != cudaSuccess) return ::pl::cuda_fail(e__, #call); \
  } while (0)

#define PL_LAUNCH_CHECK(name)                                   \
  do {                                                          \
    cudaError_t e__ = cudaGetLastError();                       \
    if (e__ != cudaSuccess) return ::pl::cuda_fail(e__, name);  \
    ::pl::count_launch();                                       \
  } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// elementwise math shared by every kernel (must agree with oracle/closed_form.py)
// ---------------------------------------------------------------------------------------------
// softplus(beta=1, threshold=20) exactly like F.softplus (src/ProbabilisticLayers.py:826,836):
// accurate log1p so that sigma ~ 1e-6 (fresh-init rho ~ -12) keeps full precision.
__device__ __forceinline__ float softplus_f(float x) {
  return x > 20.f ? x : log1pf(__expf(x));
}
__device__ __forceinline__ float sigmoid_f(float x) {
  return x > 20.f ? 1.f : __fdividef(1.f, 1.f + __expf(-x));
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 eps stream (contract in include/pl_b200.h, CPU restatement oracle/philox.py)
// ---------------------------------------------------------------------------------------------
constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    // one IMAD.WIDE.U32 per product (hi and lo halves of the 64-bit result)
    const uint64_t p0 = (uint64_t)kPhiloxM0 * c0, p1 = (uint64_t)kPhiloxM1 * c2;
    const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += kPhiloxW0;
    k1 += kPhiloxW1;
  }
  return make_uint4(c0, c1, c2, c3);
}

// One Box-Muller pair from ONE 32-bit word: radius from the top 20 bits, angle from the low 12.
//   f = 2 - t20 in (0,1] (2^-20 steps, |eps| <= 5.27), r = sqrt(-2 ln f)
//   theta = 2 pi t12 - 3 pi in (-pi, pi) on a 4096-point mid-bin grid
// The angle grid is a midpoint rule over a smooth periodic integrand, so the marginals stay Gaussian
// to far below fp32 resolution; what is given up is 12 bits of each pair's direction, what is gained
// is 8 normals per Philox call instead of 4 (the 32x32->64 multiplies of Philox are the slowest part
// of the sampler on this chip: tools/sampler_bench.cu, profiles/r01_experiments.md).
// MUFU-rate math (lg2 / sqrt / sin / cos .approx): f is a normal number, no slow paths needed.
// a constant the compiler must keep in a register: (x & mask) | bits then folds into ONE 3-input LOP3
// instead of two immediate-operand LOP3s (the sampler is issue/power bound; 8 such pairs per Philox call)
__device__ __forceinline__ uint32_t reg_const(uint32_t c) {
  uint32_t r;
  asm("mov.b32 %0, %1;" : "=r"(r) : "r"(c));
  return r;
}
__device__ __forceinline__ void box_muller(uint32_t w, float& n_cos, float& n_sin) {
  // f: top 20 bits -> mantissa bits [22:3] of a float in [1,2); t: low 12 bits -> mantissa bits [22:11], mid-bin
  const float f = 2.0f - __uint_as_float(((w >> 9) & reg_const(0x007FFFF8u)) | reg_const(0x3F800000u));
  const float t = __uint_as_float(((w << 11) & reg_const(0x007FF800u)) | reg_const(0x3F800400u));
  float lg, r, sn, cs;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(f));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * lg));
  const float theta = fmaf(t, 6.283185307179586f, -9.42477796076938f);
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(theta));
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(theta));
  n_cos = r * cs;
  n_sin = r * sn;
}

struct Sampler {
  uint32_t k0, k1;     // seed
  uint32_t stream;     // 2*layer_id + is_bias
  uint32_t mc_offset;  // global MC index of local sample 0
  const uint64_t* seed_dev;   // optional device-resident offset added to the seed (CUDA-graph replays), else NULL
};

__host__ __device__ inline Sampler make_sampler(uint64_t seed, uint32_t layer_id, bool bias,
                                                uint32_t mc_offset, const uint64_t* seed_dev = nullptr) {
  Sampler s;
  s.k0 = (uint32_t)(seed & 0xffffffffu);
  s.k1 = (uint32_t)(seed >> 32);
  s.stream = 2u * layer_id + (bias ? 1u : 0u);
  s.mc_offset = mc_offset;
  s.seed_dev = seed_dev;
  return s;
}

// The sampler a kernel actually uses: key = seed + *seed_dev (mod 2^64) when a device-side offset is given.
// A captured CUDA graph replays the same launch parameters every step; the per-step part of the seed therefore
// lives in device memory and is advanced by one small kernel inside the graph.  Call once at kernel entry.
__device__ __forceinline__ Sampler live(const Sampler& s) {
  Sampler r = s;
  if (s.seed_dev) {
    const uint64_t key = (((uint64_t)s.k1 << 32) | (uint64_t)s.k0) + *s.seed_dev;
    r.k0 = (uint32_t)(key & 0xffffffffu);
    r.k1 = (uint32_t)(key >> 32);
  }
  return r;
}

// 8 standard normals for columns 8*group8 .. 8*group8+7 of `row` of MC sample `mc` (local index):
// Philox counter (group8, row, global mc, stream); word j -> columns 8*group8 + 2j, +2j+1
__device__ __forceinline__ void eps8(const Sampler& s, uint32_t mc, uint32_t row, uint32_t group8,
                                     float (&e)[8]) {
  const uint4 w = philox4x32_10(group8, row, s.mc_offset + mc, s.stream, s.k0, s.k1);
  box_muller(w.x, e[0], e[1]);
  box_muller(w.y, e[2], e[3]);
  box_muller(w.z, e[4], e[5]);
  box_muller(w.w, e[6], e[7]);
}

// 4 of them: columns 4*group4 .. +3 (the lower or upper half of an 8-group; fp32 / bias / dump paths)
__device__ __forceinline__ float4 eps4(const Sampler& s, uint32_t mc, uint32_t row, uint32_t group4) {
  const uint4 w = philox4x32_10(group4 >> 1, row, s.mc_offset + mc, s.stream, s.k0, s.k1);
  float4 n;
  box_muller((group4 & 1) ? w.z : w.x, n.x, n.y);
  box_muller((group4 & 1) ? w.w : w.y, n.z, n.w);
  return n;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace pl
