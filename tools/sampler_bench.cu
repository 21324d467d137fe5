// Standalone micro-benchmark of the on-chip weight sampler (Philox4x32-10 -> Box-Muller -> FMA ->
// bf16 -> st.shared), the bound of the fused tcgen05 kernels (profiles/r01_experiments.md).
// Build + run on a B200:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo \
//                         -I pytorch_probabilisticlayers_b200/csrc tools/sampler_bench.cu -o /tmp/sb && /tmp/sb
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace pl {
void set_error(const char*, ...) {}
int cuda_fail(cudaError_t, const char*) { return -3; }
void count_launch() {}
}  // namespace pl
using namespace pl;

__device__ __forceinline__ float unit_float(uint32_t u) { return __uint_as_float((u >> 9) | 0x3F800000u); }
// the round-1 v1 transform: two 32-bit words per Box-Muller pair (23-bit radius, 23-bit angle)
__device__ __forceinline__ void box_muller2(uint32_t u_r, uint32_t u_a, float& n_cos, float& n_sin) {
  const float f = 2.0f - unit_float(u_r);
  float lg, r, sn, cs;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(f));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * lg));
  const float theta = fmaf(unit_float(u_a), 6.283185307179586f, -9.42477796076938f);
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(theta));
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(theta));
  n_cos = r * cs;
  n_sin = r * sn;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// variant 1: separate mulhi / mullo
__device__ __forceinline__ uint4 philox_hilo(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                             uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(kPhiloxM0, c0), lo0 = kPhiloxM0 * c0;
    const uint32_t hi1 = __umulhi(kPhiloxM1, c2), lo1 = kPhiloxM1 * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += kPhiloxW0; k1 += kPhiloxW1;
  }
  return make_uint4(c0, c1, c2, c3);
}

template <int kVariant>
__device__ __forceinline__ float4 gen4(const Sampler& s, uint32_t mc, uint32_t row, uint32_t grp) {
  uint4 w;
  if (kVariant == 1) w = philox_hilo(grp, row, s.mc_offset + mc, s.stream, s.k0, s.k1);
  else w = philox4x32_10(grp, row, s.mc_offset + mc, s.stream, s.k0, s.k1);
  float4 n;
  if (kVariant == 2) {  // no Box-Muller at all: upper bound of what a cheaper transform could buy
    n.x = unit_float(w.x); n.y = unit_float(w.y); n.z = unit_float(w.z); n.w = unit_float(w.w);
  } else if (kVariant == 3) {  // no Philox: Box-Muller only
    box_muller2(grp * 2654435761u + row, row * 40503u + grp, n.x, n.y);
    box_muller2(grp * 2246822519u + mc, row * 3266489917u + grp, n.z, n.w);
    return n;
  } else {
    box_muller2(w.x, w.y, n.x, n.y);
    box_muller2(w.z, w.w, n.z, n.w);
  }
  return n;
}

// one CTA = one SM's generator team: WARPS warps produce `iters` tiles of 2048 groups (8192 elements)
template <int WARPS, int kVariant, bool kLoadParams>
__global__ void __launch_bounds__(WARPS * 32, 1)
sampler_kernel(const float* __restrict__ mu, const float* __restrict__ sg, int ld, int iters, Sampler s,
               uint32_t* sink) {
  extern __shared__ uint8_t smem[];
  constexpr int T = WARPS * 32;
  constexpr int G = (2048 + T - 1) / T;  // groups per thread per tile
  const int t = threadIdx.x;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    float4 m[G], g[G];
#pragma unroll
    for (int r = 0; r < G; ++r) {
      const int gi = t + r * T;
      const int k = gi >> 5, n4 = gi & 31;
      if (kLoadParams && gi < 2048) {
        const int64_t idx = (int64_t)(it * 64 + k) * ld + blockIdx.x % 8 * 128 + 4 * n4;
        m[r] = __ldg(reinterpret_cast<const float4*>(mu + idx));
        g[r] = __ldg(reinterpret_cast<const float4*>(sg + idx));
      } else {
        m[r] = make_float4(0.1f, 0.2f, 0.3f, 0.4f);
        g[r] = make_float4(1e-3f, 1e-3f, 1e-3f, 1e-3f);
      }
    }
#pragma unroll
    for (int r = 0; r < G; ++r) {
      const int gi = t + r * T;
      if (gi >= 2048) continue;
      const int k = gi >> 5, n4 = gi & 31;
      const float4 e = gen4<kVariant>(s, blockIdx.y, it * 64 + k, blockIdx.x * 32 + n4);
      const uint32_t lo = pack2(fmaf(g[r].x, e.x, m[r].x), fmaf(g[r].y, e.y, m[r].y));
      const uint32_t hi = pack2(fmaf(g[r].z, e.z, m[r].z), fmaf(g[r].w, e.w, m[r].w));
      const uint32_t off = (uint32_t)((n4 >> 4) * 8192 + (k >> 3) * 1024 + (k & 7) * 128 +
                                      ((((n4 & 15) >> 1) ^ (k & 7)) << 4) + ((n4 & 1) << 3));
      *reinterpret_cast<uint2*>(smem + (it & 3) * 16384 + off) = make_uint2(lo, hi);
      acc ^= lo;
    }
    __syncthreads();
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

// variant 8: one Philox call -> 8 normals (20-bit radius + 12-bit angle per pair), groups of 8 columns
__device__ __forceinline__ void bm_20_12(uint32_t w, float& n_cos, float& n_sin) {
  const float f = 2.0f - __uint_as_float(((w >> 12) << 3) | 0x3F800000u);        // top 20 bits -> (0,1]
  const float t = __uint_as_float(((w & 0xFFFu) << 11) | 0x3F800400u);           // low 12 bits -> [1,2), mid-bin
  float lg, r, sn, cs;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(f));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(-1.3862943611198906f * lg));
  const float theta = fmaf(t, 6.283185307179586f, -9.42477796076938f);
  asm("cos.approx.ftz.f32 %0, %1;" : "=f"(cs) : "f"(theta));
  asm("sin.approx.ftz.f32 %0, %1;" : "=f"(sn) : "f"(theta));
  n_cos = r * cs;
  n_sin = r * sn;
}

template <int WARPS, bool kLoadParams>
__global__ void __launch_bounds__(WARPS * 32, 1)
sampler8_kernel(const float* __restrict__ mu, const float* __restrict__ sg, int ld, int iters, Sampler s,
                uint32_t* sink) {
  extern __shared__ uint8_t smem[];
  constexpr int T = WARPS * 32;
  constexpr int G = (1024 + T - 1) / T;  // 8-element groups per thread per tile
  const int t = threadIdx.x;
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    float4 m[G][2], g[G][2];
#pragma unroll
    for (int r = 0; r < G; ++r) {
      const int gi = t + r * T;
      const int k = gi >> 4, n8 = gi & 15;
      if (kLoadParams && gi < 1024) {
        const int64_t idx = (int64_t)(it * 64 + k) * ld + blockIdx.x % 8 * 128 + 8 * n8;
        m[r][0] = __ldg(reinterpret_cast<const float4*>(mu + idx));
        m[r][1] = __ldg(reinterpret_cast<const float4*>(mu + idx + 4));
        g[r][0] = __ldg(reinterpret_cast<const float4*>(sg + idx));
        g[r][1] = __ldg(reinterpret_cast<const float4*>(sg + idx + 4));
      } else {
        m[r][0] = m[r][1] = make_float4(0.1f, 0.2f, 0.3f, 0.4f);
        g[r][0] = g[r][1] = make_float4(1e-3f, 1e-3f, 1e-3f, 1e-3f);
      }
    }
#pragma unroll
    for (int r = 0; r < G; ++r) {
      const int gi = t + r * T;
      if (gi >= 1024) continue;
      const int k = gi >> 4, n8 = gi & 15;
      const uint4 w = philox4x32_10(blockIdx.x * 16 + n8, it * 64 + k, s.mc_offset + blockIdx.y, s.stream, s.k0, s.k1);
      float e[8];
      bm_20_12(w.x, e[0], e[1]);
      bm_20_12(w.y, e[2], e[3]);
      bm_20_12(w.z, e[4], e[5]);
      bm_20_12(w.w, e[6], e[7]);
      const uint32_t p0 = pack2(fmaf(g[r][0].x, e[0], m[r][0].x), fmaf(g[r][0].y, e[1], m[r][0].y));
      const uint32_t p1 = pack2(fmaf(g[r][0].z, e[2], m[r][0].z), fmaf(g[r][0].w, e[3], m[r][0].w));
      const uint32_t p2 = pack2(fmaf(g[r][1].x, e[4], m[r][1].x), fmaf(g[r][1].y, e[5], m[r][1].y));
      const uint32_t p3 = pack2(fmaf(g[r][1].z, e[6], m[r][1].z), fmaf(g[r][1].w, e[7], m[r][1].w));
      const uint32_t off = (uint32_t)((n8 >> 3) * 8192 + (k >> 3) * 1024 + (k & 7) * 128 + (((n8 & 7) ^ (k & 7)) << 4));
      *reinterpret_cast<uint4*>(smem + (it & 3) * 16384 + off) = make_uint4(p0, p1, p2, p3);
      acc ^= p0;
    }
    __syncthreads();
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <int WARPS, bool kLoad>
static void run8(const char* name, const float* mu, const float* sg) {
  const int iters = 64, sms = 148;
  Sampler s = make_sampler(1234, 1, false, 0);
  uint32_t* sink;
  cudaMalloc(&sink, 4);
  auto kern = sampler8_kernel<WARPS, kLoad>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  dim3 grid(32, sms * 2 / 32 + 1);
  kern<<<grid, WARPS * 32, 65536>>>(mu, sg, 4096, iters, s, sink);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int rep = 0; rep < 5; ++rep) kern<<<grid, WARPS * 32, 65536>>>(mu, sg, 4096, iters, s, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 5;
  const double elems = (double)grid.x * grid.y * iters * 8192.0;
  const double waves = (double)grid.x * grid.y / sms;
  printf("%-46s %7.3f ms  %6.2f Gelem/s  (%.2f elem/ns/SM; err=%s)\n", name, ms, elems / ms * 1e-6,
         elems / ms * 1e-6 / sms * (waves / ceil(waves)), cudaGetErrorString(cudaGetLastError()));
  cudaFree(sink);
}

template <int WARPS, int kVariant, bool kLoad>
static void run(const char* name, const float* mu, const float* sg) {
  const int iters = 64, sms = 148;
  Sampler s = make_sampler(1234, 1, false, 0);
  uint32_t* sink;
  cudaMalloc(&sink, 4);
  auto kern = sampler_kernel<WARPS, kVariant, kLoad>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  dim3 grid(32, sms * 2 / 32 + 1);  // ~2 waves of one CTA per SM
  kern<<<grid, WARPS * 32, 65536>>>(mu, sg, 4096, iters, s, sink);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  for (int rep = 0; rep < 5; ++rep) kern<<<grid, WARPS * 32, 65536>>>(mu, sg, 4096, iters, s, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  ms /= 5;
  const double elems = (double)grid.x * grid.y * iters * 8192.0;
  const double waves = (double)grid.x * grid.y / sms;
  printf("%-46s %7.3f ms  %6.2f Gelem/s  (%.2f elem/ns/SM; err=%s)\n", name, ms, elems / ms * 1e-6,
         elems / ms * 1e-6 / sms * (waves / ceil(waves)), cudaGetErrorString(cudaGetLastError()));
  cudaFree(sink);
}

int main() {
  float *mu, *sg;
  cudaMalloc(&mu, 4096ull * 4096 * 4);
  cudaMalloc(&sg, 4096ull * 4096 * 4);
  cudaMemset(mu, 0, 4096ull * 4096 * 4);
  cudaMemset(sg, 0, 4096ull * 4096 * 4);
  printf("needed for 100%% of tensor peak at B=512: 8 elem/clk/SM = 15.7 elem/ns/SM at 1.965 GHz\n");
  run<16, 0, true>("16 warps, philox(mul.wide)+BM, param loads", mu, sg);
  run<16, 0, false>("16 warps, philox(mul.wide)+BM, no loads", mu, sg);
  run<16, 1, false>("16 warps, philox(mulhi+mullo)+BM, no loads", mu, sg);
  run<16, 2, false>("16 warps, philox only (no Box-Muller)", mu, sg);
  run<16, 3, false>("16 warps, Box-Muller only (no philox)", mu, sg);
  run<8, 0, false>("8 warps, philox+BM, no loads", mu, sg);
  run<12, 0, false>("12 warps, philox+BM, no loads", mu, sg);
  run<20, 0, false>("20 warps, philox+BM, no loads", mu, sg);
  run<24, 0, false>("24 warps, philox+BM, no loads", mu, sg);
  run<32, 0, false>("32 warps, philox+BM, no loads", mu, sg);
  run<32, 0, true>("32 warps, philox+BM, param loads", mu, sg);
  run8<16, false>("16 warps, 8 normals/philox (20+12 bit), no loads", mu, sg);
  run8<16, true>("16 warps, 8 normals/philox (20+12 bit), loads", mu, sg);
  run8<8, true>("8 warps, 8 normals/philox (20+12 bit), loads", mu, sg);
  run8<32, true>("32 warps, 8 normals/philox (20+12 bit), loads", mu, sg);
  return 0;
}
