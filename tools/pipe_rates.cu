// Measures issue cost (cycles per warp-instruction per SM sub-partition) of the instruction kinds the
// weight sampler is made of, alone and mixed, on one SM.  Used to build the sampler cost model in
// profiles/r01_experiments.md.   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/pipe_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int kIters = 2048;
constexpr int kChains = 8;

template <int kKind>
__global__ void rate_kernel(uint32_t* sink, long long* cycles, uint32_t seed) {
  uint32_t a[kChains];
  float f[kChains];
  uint64_t w[kChains];
#pragma unroll
  for (int c = 0; c < kChains; ++c) {
    a[c] = seed * (c + 1) + threadIdx.x;
    f[c] = 1.0f + (float)(threadIdx.x + c) * 1e-3f;
    w[c] = a[c];
  }
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
      if (kKind == 0) {  // IMAD.WIDE.U32
        w[c] = (uint64_t)(uint32_t)w[c] * 0xD2511F53u + (w[c] >> 32);
      } else if (kKind == 1) {  // LOP3
        a[c] = a[c] ^ (a[(c + 1) % kChains] & seed);
      } else if (kKind == 2) {  // FFMA
        f[c] = fmaf(f[c], 1.0001f, f[(c + 1) % kChains]);
      } else if (kKind == 3) {  // MUFU.LG2
        asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
      } else if (kKind == 4) {  // MUFU.SIN (with its range-reduction FMUL)
        asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
      } else if (kKind == 5) {  // MUFU.SQRT
        asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
      } else if (kKind == 6) {  // IMAD (32-bit lo)
        a[c] = a[c] * 0xD2511F53u + seed;
      } else if (kKind == 7) {  // IMAD.HI
        a[c] = __umulhi(a[c], 0xD2511F53u) + 1;
      } else if (kKind == 8) {  // Philox half-round: IMAD.WIDE + LOP3
        w[c] = (uint64_t)(uint32_t)w[c] * 0xD2511F53u;
        w[c] = (w[c] & 0xffffffff00000000ull) | (uint32_t)((uint32_t)(w[c] >> 32) ^ (uint32_t)w[c] ^ seed);
      } else if (kKind == 9) {  // IMAD.WIDE + MUFU.LG2 (independent streams)
        w[c] = (uint64_t)(uint32_t)w[c] * 0xD2511F53u + (w[c] >> 32);
        asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
      } else if (kKind == 10) {  // HFMA2.BF16
        asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(seed), "r"(a[(c + 1) % kChains]));
      } else if (kKind == 11) {  // F2FP pack
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(a[c]) : "f"(f[c]), "f"(__uint_as_float(a[c])));
      } else if (kKind == 12) {  // mul.lo + mul.hi separately (2 instr)
        const uint32_t lo = a[c] * 0xD2511F53u, hi = __umulhi(a[c], 0xD2511F53u);
        a[c] = lo ^ hi;
      } else if (kKind == 13) {  // FFMA + LOP3 (two pipes)
        f[c] = fmaf(f[c], 1.0001f, f[(c + 1) % kChains]);
        a[c] = a[c] ^ (a[(c + 1) % kChains] & seed);
      } else if (kKind == 14) {  // MUFU.COS
        asm volatile("cos.approx.ftz.f32 %0, %0;" : "+f"(f[c]));
      } else if (kKind == 15) {  // FFMA + MUFU
        f[c] = fmaf(f[c], 1.0001f, f[(c + 1) % kChains]);
        float g = __uint_as_float(a[c]);
        asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(g));
        a[c] = __float_as_uint(g);
      }
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int c = 0; c < kChains; ++c) acc ^= a[c] ^ __float_as_uint(f[c]) ^ (uint32_t)w[c] ^ (uint32_t)(w[c] >> 32);
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int kKind>
void run(const char* name, int per_iter_instrs) {
  uint32_t* sink; long long* cyc;
  cudaMalloc(&sink, 4 * 1024 * 4); cudaMalloc(&cyc, 8 * 4);
  for (int warps : {4, 16, 32}) {
    rate_kernel<kKind><<<1, warps * 32>>>(sink, cyc, 12345u);
    rate_kernel<kKind><<<1, warps * 32>>>(sink, cyc, 12345u);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double warp_instr_per_smsp = (double)kIters * kChains * per_iter_instrs * (warps / 4.0);
    printf("%-28s warps/SMSP=%d  cycles/warp-instr/SMSP = %.2f\n", name, warps / 4, c / warp_instr_per_smsp);
  }
  cudaFree(sink); cudaFree(cyc);
}

int main() {
  run<0>("IMAD.WIDE.U32", 1);
  run<1>("LOP3", 1);
  run<2>("FFMA", 1);
  run<3>("MUFU.LG2", 1);
  run<4>("sin.approx (FMUL+MUFU.SIN)", 2);
  run<14>("cos.approx (FMUL+MUFU.COS)", 2);
  run<5>("MUFU.SQRT", 1);
  run<6>("IMAD lo", 1);
  run<7>("IMAD.HI (+IADD)", 2);
  run<8>("IMAD.WIDE + LOP3", 2);
  run<9>("IMAD.WIDE + MUFU.LG2", 2);
  run<10>("HFMA2.BF16", 1);
  run<11>("F2FP.BF16 pack", 1);
  run<12>("IMAD lo + IMAD.HI + LOP3", 3);
  run<13>("FFMA + LOP3", 2);
  run<15>("FFMA + MUFU.LG2", 2);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
