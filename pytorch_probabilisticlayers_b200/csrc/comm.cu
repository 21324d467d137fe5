// MC-sharded train step without NCCL on the data path: gradient reduce-scatter, sharded Adam + KL, and the all-gather
// of the updated parameters as plain loads / stores over NVLink peer mappings (CUDA IPC), fused into the kernels
// that produce and consume the data.
//
// Every rank owns one ARENA with the same layout (symmetric): [fp32 parameters | gradient staging G | low-precision
// operand area | statistics staging | flags].  Rank r owns elements [r*chunk, (r+1)*chunk) of every parameter tensor.
//   push     each rank scatters its local gradient of tensor t, owner o's part, into  peers[o].G[t][rank][...]
//            (reduce-scatter by peer stores: the owner later sums the N staged partials in a FIXED order, so the
//            result is deterministic and the same for every N-way launch)
//   barrier  flags in peer memory, st.release.sys / ld.acquire.sys, epoch counter on the device (graph-replayable)
//   adam     the owner sums the staged partials, applies sigmoid(rho) to raw drho sums, adds the analytic KL gradient,
//            does the Adam update on ITS shard only (m, v exist only for the shard), accumulates its part of the KL
//            value, and writes the updated parameters in the form the forward consumes -- the bf16 {mu | sigma}
//            operand of the tcgen05 kernels (4 B/weight) or fp32 for small tensors -- straight into EVERY rank's arena
//            (all-gather by peer stores)
//   finish   shares the KL partials, second barrier, sums the statistics
// Replaces ncclAllReduce(301 MB fp32) + a replicated optimiser pass over all 75 M parameters (r01: ~1.5 ms of a 3.2 ms
// step at N = 8).  Reference being matched: Lightning DDP's bucketed all-reduce + Adam, experiments/BNN.py:349,399-400.
#include <cuda.h>
#include <math.h>
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pl {

using tc::pack_bf16x2;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// 16-byte store through the NVLS multicast mapping: the switch replicates it into every rank's arena
__device__ __forceinline__ void multimem_st16(void* mc_addr, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "f"(__uint_as_float(v.x)),
               "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w))
               : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One cross-GPU barrier, executed by the first `world` threads of ONE block.  epoch = ++(*epoch_dev) (thread 0).
// Never hangs: after ~4 s of spinning it raises status[0] and proceeds (the caller checks the status word).
// `channel` (0..3) selects an independent set of flags: barriers issued on different streams must not share an epoch counter.
__device__ void grid_barrier_block(const pl_shard_ctx& c, uint32_t* s_epoch, int channel = 0) {
  uint8_t* base = (uint8_t*)c.peers[c.rank];
  const int64_t foff = c.flags_off + 64 * channel;
  uint32_t* epoch_dev = (uint32_t*)(base + foff);                 // [0] epoch, [1] status, [8 + q] flag of rank q
  if (threadIdx.x == 0) *s_epoch = ++(*epoch_dev);
  __syncthreads();
  const uint32_t e = *s_epoch;
  if ((int)threadIdx.x < c.world) {
    __threadfence_system();       // everything this GPU wrote before (this and earlier kernels) is ordered before the flag
    uint32_t* remote = (uint32_t*)((uint8_t*)c.peers[threadIdx.x] + foff) + 8 + c.rank;
    st_release_sys(remote, e);
    const uint32_t* mine = epoch_dev + 8 + threadIdx.x;
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(mine) - e) < 0) {
      if (clock64() - t0 > (1ll << 33)) {
        ((uint32_t*)(base + c.flags_off))[1] = 1u;        // timeout: a peer never arrived (status word of channel 0)
        break;
      }
      __nanosleep(64);
    }
    __threadfence_system();
  }
  __syncthreads();
}

__global__ void shard_barrier_kernel(const pl_shard_ctx c, int channel) {
  __shared__ uint32_t s_epoch;
  grid_barrier_block(c, &s_epoch, channel);
}

// ---- push: local gradients -> owners' staging --------------------------------------------------------------------
// One thread per float4: consecutive threads move consecutive 16-byte pieces, so every warp-level store is 512 contiguous
// bytes (full 32-byte sectors -- partial-sector stores become separate byte-enabled NVLink packets).
// push_prefix[s] = first 8-element unit of tensor slot s (2 slots per segment: [mean | rho]); 2 float4 per unit.
__global__ void __launch_bounds__(256) shard_push_kernel(const pl_shard_ctx c, const float* __restrict__ grads,
                                                         const float* __restrict__ stats4, int64_t u0, int64_t u1) {
  const pl_shard_seg* segs = c.segs;
  const int64_t total = 2 * u1;
  int slot = 0;
  if (c.wire_bf16) {
    // bf16 wire format: one thread per 8-element unit, 16-byte stores of 8 bf16 partials (consecutive threads contiguous)
    for (int64_t u = u0 + (int64_t)blockIdx.x * 256 + threadIdx.x; u < u1; u += (int64_t)gridDim.x * 256) {
      while (u >= c.push_prefix[slot + 1]) ++slot;
      const pl_shard_seg& sg = segs[slot >> 1];
      const bool second = slot & 1;
      const int64_t f0 = (u - c.push_prefix[slot]) * 8;
      const int64_t numel = sg.numel, chunk = sg.chunk;
      const int owner = (int)(f0 / chunk);
      const int64_t j = f0 - (int64_t)owner * chunk;
      const float* src = grads + (second ? sg.p2_off : sg.p_off) / 4 + f0;
      float v[8];
      if (f0 + 8 <= numel) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = f0 + t < numel ? __ldg(src + t) : 0.f;
      }
      uint16_t* dst = (uint16_t*)((uint8_t*)c.peers[owner] + (second ? sg.g2_off : sg.g_off)) + (int64_t)c.rank * chunk + j;
      *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                                                  pack_bf16x2(v[6], v[7]));
    }
  } else
  for (int64_t q = 2 * u0 + (int64_t)blockIdx.x * 256 + threadIdx.x; q < total; q += (int64_t)gridDim.x * 256) {
    const int64_t u = q >> 1;
    while (u >= c.push_prefix[slot + 1]) ++slot;
    const pl_shard_seg& sg = segs[slot >> 1];
    const bool second = slot & 1;
    const int64_t f0 = (u - c.push_prefix[slot]) * 8 + (q & 1) * 4;
    const int64_t numel = sg.numel, chunk = sg.chunk;
    if (f0 >= numel) continue;
    const int owner = (int)(f0 / chunk);            // chunk % 8 == 0: a float4 never straddles two owners
    const int64_t j = f0 - (int64_t)owner * chunk;
    const float* src = grads + (second ? sg.p2_off : sg.p_off) / 4 + f0;
    float* dst = (float*)((uint8_t*)c.peers[owner] + (second ? sg.g2_off : sg.g_off)) + (int64_t)c.rank * chunk + j;
    if (f0 + 4 <= numel) {
      *reinterpret_cast<float4*>(dst) = __ldg(reinterpret_cast<const float4*>(src));
    } else {
      for (int t = 0; t < 4 && f0 + t < numel; ++t) dst[t] = __ldg(src + t);
    }
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < c.world && stats4) {
    // two statistics slots, alternating by step: a rank that is already pushing step k+1 cannot overwrite what a slower
    // peer is still summing for step k (channel-0 epoch: two barriers per step, so bit 1 is the step parity)
    const uint32_t epoch = *(const uint32_t*)((const uint8_t*)c.peers[c.rank] + c.flags_off);
    const int slot = (epoch >> 1) & 1;
    float* dst = (float*)((uint8_t*)c.peers[threadIdx.x] + c.stats_off) + 4 * (slot * PL_SHARD_MAX_RANKS + c.rank);
#pragma unroll
    for (int t = 0; t < 4; ++t) dst[t] = stats4[t];
  }
}

// ---- sharded Adam + KL + all-gather -------------------------------------------------------------------------------
__device__ __forceinline__ void sigma_parts2(float rho, float& sigma, float& logsigma, float& sig) {
  const float e = __expf(fminf(rho, 20.f));
  const float poly = e * fmaf(e, fmaf(e, fmaf(e, fmaf(e, 0.2f, -0.25f), 0.33333334f), -0.5f), 1.0f);
  sigma = e < 0.0625f ? poly : __logf(1.0f + e);
  sigma = rho > 20.f ? rho : sigma;
  logsigma = __logf(sigma);
  sig = rho > 20.f ? 1.f : __fdividef(e, 1.f + e);
}

struct AdamHyper {
  float kl_scale, step_size, inv_sqrt_bc2, beta1, beta2, eps;
  const float* sched_dev;
};

__device__ __forceinline__ float adam_update(float w, float grad, float& m, float& v, const AdamHyper& h) {
  m = fmaf(h.beta1, m, (1.f - h.beta1) * grad);
  v = fmaf(h.beta2, v, (1.f - h.beta2) * grad * grad);
  return w - h.step_size * __fdividef(m, fmaf(sqrtf(v), h.inv_sqrt_bc2, h.eps));
}

// 8 consecutive parameter / m / v values: 128-bit accesses on the full path (all offsets are multiples of 8 elements)
__device__ __forceinline__ void load8(const float* P, const float* M, const float* V, int nval, float (&w)[8], float (&m)[8],
                                      float (&v)[8]) {
  if (nval == 8) {
    const float4 a = *reinterpret_cast<const float4*>(P), b = *reinterpret_cast<const float4*>(P + 4);
    const float4 c = *reinterpret_cast<const float4*>(M), d = *reinterpret_cast<const float4*>(M + 4);
    const float4 e = *reinterpret_cast<const float4*>(V), f = *reinterpret_cast<const float4*>(V + 4);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    m[0] = c.x; m[1] = c.y; m[2] = c.z; m[3] = c.w; m[4] = d.x; m[5] = d.y; m[6] = d.z; m[7] = d.w;
    v[0] = e.x; v[1] = e.y; v[2] = e.z; v[3] = e.w; v[4] = f.x; v[5] = f.y; v[6] = f.z; v[7] = f.w;
  } else {
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const bool ok = t < nval;
      w[t] = ok ? P[t] : 0.f;
      m[t] = ok ? M[t] : 0.f;
      v[t] = ok ? V[t] : 0.f;
    }
  }
}
__device__ __forceinline__ void store8(float* P, float* M, float* V, int nval, const float (&w)[8], const float (&m)[8],
                                       const float (&v)[8]) {
  if (nval == 8) {
    reinterpret_cast<float4*>(P)[0] = make_float4(w[0], w[1], w[2], w[3]);
    reinterpret_cast<float4*>(P)[1] = make_float4(w[4], w[5], w[6], w[7]);
    reinterpret_cast<float4*>(M)[0] = make_float4(m[0], m[1], m[2], m[3]);
    reinterpret_cast<float4*>(M)[1] = make_float4(m[4], m[5], m[6], m[7]);
    reinterpret_cast<float4*>(V)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(V)[1] = make_float4(v[4], v[5], v[6], v[7]);
  } else {
#pragma unroll
    for (int t = 0; t < 8; ++t)       // fully unrolled + predicated: the arrays must stay in registers
      if (t < nval) {
        P[t] = w[t];
        M[t] = m[t];
        V[t] = v[t];
      }
  }
}

// sum over the ranks' staged partials of 8 consecutive elements, in rank order (deterministic); fp32 or bf16 wire format
__device__ __forceinline__ void staged_sum8(const uint8_t* stage, int64_t chunk, int64_t j, int world, bool bf16, float (&g)[8]) {
#pragma unroll
  for (int t = 0; t < 8; ++t) g[t] = 0.f;
  if (bf16) {
    const uint16_t* G = reinterpret_cast<const uint16_t*>(stage) + j;
    for (int s = 0; s < world; ++s) {
      const uint4 q = *reinterpret_cast<const uint4*>(G + (int64_t)s * chunk);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        g[2 * t] += __uint_as_float(w[t] << 16);
        g[2 * t + 1] += __uint_as_float(w[t] & 0xffff0000u);
      }
    }
  } else {
    const float* G = reinterpret_cast<const float*>(stage) + j;
    for (int s = 0; s < world; ++s) {
      const float4 a = *reinterpret_cast<const float4*>(G + (int64_t)s * chunk);
      const float4 b = *reinterpret_cast<const float4*>(G + (int64_t)s * chunk + 4);
      g[0] += a.x; g[1] += a.y; g[2] += a.z; g[3] += a.w; g[4] += b.x; g[5] += b.y; g[6] += b.z; g[7] += b.w;
    }
  }
}

// unit = 8 consecutive elements of this rank's chunk of one segment (both tensors of a pair); adam_prefix[s] = first unit
__global__ void __launch_bounds__(256, 2) shard_adam_kernel(const pl_shard_ctx c, float* __restrict__ mstate,
                                                         float* __restrict__ vstate, AdamHyper h,
                                                         double* __restrict__ kl_out, int64_t unit_begin, int64_t unit_end) {
  if (h.sched_dev) {
    h.step_size = __ldg(h.sched_dev);
    h.inv_sqrt_bc2 = __ldg(h.sched_dev + 1);
  }
  uint8_t* base = (uint8_t*)c.peers[c.rank];
  const int world = c.world, rank = c.rank;
  const int64_t total = unit_end;
  // the segment table is read once per 8 elements: keep it in shared memory
  __shared__ pl_shard_seg s_segs[PL_SHARD_MAX_SEGS];
  for (int i = threadIdx.x; i < c.num_segs * (int)(sizeof(pl_shard_seg) / 8); i += 256)
    reinterpret_cast<int64_t*>(s_segs)[i] = reinterpret_cast<const int64_t*>(c.segs)[i];
  __syncthreads();
  float kl = 0.f;
  int si = 0;
  for (int64_t u = unit_begin + (int64_t)blockIdx.x * 256 + threadIdx.x; u < total; u += (int64_t)gridDim.x * 256) {
    while (u >= c.adam_prefix[si + 1]) ++si;
    const pl_shard_seg& sg = s_segs[si];
    const int64_t j = (u - c.adam_prefix[si]) * 8;            // offset inside this rank's chunk
    const int64_t f0 = (int64_t)rank * sg.chunk + j;          // element index inside the tensor
    const int nval = (int)min((int64_t)8, sg.numel - f0);     // >= 1 by construction of the prefix
    const float ip2 = sg.kind ? 1.f / (sg.prior_scale * sg.prior_scale) : 0.f;
    float w[8], g[8], m[8], v[8];
    // ---- first tensor (plain parameter, or the posterior mean of a pair) ----
    {
      staged_sum8(base + sg.g_off, sg.chunk, j, world, c.wire_bf16 != 0, g);
      float* P = (float*)(base + sg.p_off) + f0;
      float* M = mstate + sg.s_off + j;
      float* V = vstate + sg.s_off + j;
      load8(P, M, V, nval, w, m, v);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const bool ok = t < nval;
        float grad = g[t];
        if (sg.kind == 1) {
          grad = fmaf(h.kl_scale * ip2, w[t], grad);
          if (ok) kl = fmaf(0.5f * ip2 * w[t], w[t], kl);
        }
        w[t] = adam_update(w[t], grad, m[t], v[t], h);
      }
      store8(P, M, V, nval, w, m, v);
      if (sg.gather == 0 && c.mc_base && nval == 8 && world > 1) {
        uint8_t* R = (uint8_t*)c.mc_base + sg.p_off + f0 * 4;
        multimem_st16(R, make_uint4(__float_as_uint(w[0]), __float_as_uint(w[1]), __float_as_uint(w[2]), __float_as_uint(w[3])));
        multimem_st16(R + 16, make_uint4(__float_as_uint(w[4]), __float_as_uint(w[5]), __float_as_uint(w[6]), __float_as_uint(w[7])));
      } else if (sg.gather == 0) {        // fp32 all-gather: the updated values into every other rank's parameter area
        for (int q = 0; q < world; ++q) {
          if (q == rank) continue;
          float* R = (float*)((uint8_t*)c.peers[q] + sg.p_off) + f0;
          if (nval == 8) {
            reinterpret_cast<float4*>(R)[0] = make_float4(w[0], w[1], w[2], w[3]);
            reinterpret_cast<float4*>(R)[1] = make_float4(w[4], w[5], w[6], w[7]);
          } else {
#pragma unroll
            for (int t = 0; t < 8; ++t)
              if (t < nval) R[t] = w[t];
          }
        }
      }
    }
    if (sg.kind != 1) continue;
    // ---- second tensor: the posterior rho of the pair ----
    const uint4 mu16 = make_uint4(pack_bf16x2(w[0], w[1]), pack_bf16x2(w[2], w[3]), pack_bf16x2(w[4], w[5]),
                                  pack_bf16x2(w[6], w[7]));
    float sg8[8];
    {
      staged_sum8(base + sg.g2_off, sg.chunk, j, world, c.wire_bf16 != 0, g);
      float* P = (float*)(base + sg.p2_off) + f0;
      float* M = mstate + sg.s_off + sg.chunk + j;
      float* V = vstate + sg.s_off + sg.chunk + j;
      load8(P, M, V, nval, w, m, v);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const bool ok = t < nval;
        const float rho = w[t];
        float sigma, ls, sig;
        sigma_parts2(rho, sigma, ls, sig);
        float grad = sg.rho_raw ? g[t] * sig : g[t];
        grad = fmaf(h.kl_scale * sig, fmaf(sigma, ip2, -__fdividef(1.f, sigma)), grad);
        if (ok) kl += fmaf(0.5f * ip2 * sigma, sigma, -ls);
        w[t] = adam_update(rho, grad, m[t], v[t], h);
        sg8[t] = ok ? softplus_f(w[t]) : 0.f;
      }
      store8(P, M, V, nval, w, m, v);
    }
    if (sg.gather == 0 && c.mc_base && nval == 8 && world > 1) {
      uint8_t* R = (uint8_t*)c.mc_base + sg.p2_off + f0 * 4;
      multimem_st16(R, make_uint4(__float_as_uint(w[0]), __float_as_uint(w[1]), __float_as_uint(w[2]), __float_as_uint(w[3])));
      multimem_st16(R + 16, make_uint4(__float_as_uint(w[4]), __float_as_uint(w[5]), __float_as_uint(w[6]), __float_as_uint(w[7])));
    } else if (sg.gather == 0) {
      for (int q = 0; q < world; ++q) {
        if (q == rank) continue;
        float* R = (float*)((uint8_t*)c.peers[q] + sg.p2_off) + f0;
        if (nval == 8) {
          reinterpret_cast<float4*>(R)[0] = make_float4(w[0], w[1], w[2], w[3]);
          reinterpret_cast<float4*>(R)[1] = make_float4(w[4], w[5], w[6], w[7]);
        } else {
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (t < nval) R[t] = w[t];
        }
      }
    } else {
      // the tcgen05 kernels' operand form, 4 B/weight, into EVERY rank's arena (the owner's own included):
      //   gather 1 (split):  mu16 [numel] bf16 at lp_off, sig16 [numel] bf16 at lp_off + lp_half
      //   gather 2 (packed): {8 mu bf16 | 8 sigma bf16} groups of 32 bytes at lp_off
      const uint4 s16 = make_uint4(pack_bf16x2(sg8[0], sg8[1]), pack_bf16x2(sg8[2], sg8[3]), pack_bf16x2(sg8[4], sg8[5]),
                                   pack_bf16x2(sg8[6], sg8[7]));
      const int64_t grp = f0 >> 3;
      if (c.mc_base && world > 1) {       // one multicast store per 16 bytes: the NVSwitch fans it out to all arenas
        uint8_t* L = (uint8_t*)c.mc_base + sg.lp_off;
        if (sg.gather == 1) {
          multimem_st16(L + grp * 16, mu16);
          multimem_st16(L + sg.lp_half + grp * 16, s16);
        } else {
          multimem_st16(L + grp * 32, mu16);
          multimem_st16(L + grp * 32 + 16, s16);
        }
      } else
      for (int q = 0; q < world; ++q) {
        uint8_t* L = (uint8_t*)c.peers[q] + sg.lp_off;
        if (sg.gather == 1) {
          reinterpret_cast<uint4*>(L)[grp] = mu16;
          reinterpret_cast<uint4*>(L + sg.lp_half)[grp] = s16;
        } else {
          reinterpret_cast<uint4*>(L)[2 * grp] = mu16;
          reinterpret_cast<uint4*>(L)[2 * grp + 1] = s16;
        }
      }
    }
  }
  if (kl_out) {
    double d = warp_sum((double)kl);
    __shared__ double sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
      d = 0.0;
      for (int k = 0; k < 8; ++k) d += sm[k];
      if (d != 0.0) atomicAdd(kl_out, d);
    }
  }
}

// ---- finish: share the KL partials, barrier, sum the statistics ---------------------------------------------------
// stats_out[0] = sum_r LL_r, [1] = (sum_r kl_r + kl_const) * inv_ns, [2] = sum_r ACC_r, [3] = [0] + [1]
__global__ void shard_finish_kernel(const pl_shard_ctx c, const double* __restrict__ kl_acc, double kl_const, double inv_ns,
                                    float* __restrict__ stats_out) {
  __shared__ uint32_t s_epoch;
  // step parity from the channel-0 epoch BEFORE this kernel's barrier (odd here: the step's first barrier is behind us)
  const uint32_t epoch0 = *(const uint32_t*)((const uint8_t*)c.peers[c.rank] + c.flags_off);
  const int slot = (epoch0 >> 1) & 1;
  if ((int)threadIdx.x < c.world) {
    double* dst = (double*)((uint8_t*)c.peers[threadIdx.x] + c.kl_off) + slot * PL_SHARD_MAX_RANKS + c.rank;
    *dst = kl_acc ? *kl_acc : 0.0;
  }
  __syncthreads();
  grid_barrier_block(c, &s_epoch);
  if (threadIdx.x == 0 && stats_out) {
    const uint8_t* base = (const uint8_t*)c.peers[c.rank];
    const float* st = (const float*)(base + c.stats_off) + 4 * slot * PL_SHARD_MAX_RANKS;
    const double* kls = (const double*)(base + c.kl_off) + slot * PL_SHARD_MAX_RANKS;
    float ll = 0.f, acc = 0.f;
    double kl = 0.0;
    for (int r = 0; r < c.world; ++r) {       // fixed order
      ll += st[4 * r];
      acc += st[4 * r + 2];
      kl += kls[r];
    }
    const float klv = (float)((kl + kl_const) * inv_ns);
    stats_out[0] = ll;
    stats_out[1] = klv;
    stats_out[2] = acc;
    stats_out[3] = ll + klv;
  }
}

static int check_ctx(const pl_shard_ctx* c, const char* who) {
  PL_CHECK_ARG(c, "%s: NULL context", who);
  PL_CHECK_ARG(c->world >= 1 && c->world <= PL_SHARD_MAX_RANKS && c->rank >= 0 && c->rank < c->world,
               "%s: bad rank / world (%d / %d)", who, c->rank, c->world);
  PL_CHECK_ARG(c->num_segs >= 0 && c->num_segs <= PL_SHARD_MAX_SEGS, "%s: too many segments (%d)", who, c->num_segs);
  for (int r = 0; r < c->world; ++r) PL_CHECK_ARG(c->peers[r], "%s: NULL arena pointer of rank %d", who, r);
  PL_CHECK_ARG(c->segs || c->num_segs == 0, "%s: NULL segment table", who);
  return PL_OK;
}

}  // namespace pl

// ---- CUDA IPC plumbing ------------------------------------------------------------------------------------------
extern "C" int pl_ipc_export(const void* ptr, void* handle64, int64_t* offset_out) {
  using namespace pl;
  PL_CHECK_ARG(ptr && handle64 && offset_out, "pl_ipc_export: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CUdeviceptr base = 0;
  size_t size = 0;
  typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) {
    set_error("pl_ipc_export: cuMemGetAddressRange entry point not available");
    return PL_ERR_CUDA;
  }
  if (((RangeFn)fn)(&base, &size, (CUdeviceptr)ptr) != CUDA_SUCCESS) {
    set_error("pl_ipc_export: cuMemGetAddressRange failed (not a device allocation?)");
    return PL_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  PL_CUDA(cudaIpcGetMemHandle(&h, (void*)base));
  memcpy(handle64, &h, 64);
  *offset_out = (int64_t)((CUdeviceptr)ptr - base);
  return PL_OK;
}

extern "C" int pl_ipc_open(const void* handle64, int64_t offset, void** ptr_out) {
  using namespace pl;
  PL_CHECK_ARG(handle64 && ptr_out, "pl_ipc_open: NULL argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* base = nullptr;
  PL_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = (uint8_t*)base + offset;
  return PL_OK;
}

extern "C" int pl_ipc_close(void* ptr, int64_t offset) {
  using namespace pl;
  if (!ptr) return PL_OK;
  PL_CUDA(cudaIpcCloseMemHandle((uint8_t*)ptr - offset));
  return PL_OK;
}

// ---- sharded step ----------------------------------------------------------------------------------------------
extern "C" int pl_shard_push_range(const pl_shard_ctx* c, const float* grads, const float* stats4, int64_t unit_begin,
                                   int64_t unit_end, void* stream) {
  using namespace pl;
  int rc = check_ctx(c, "pl_shard_push");
  if (rc != PL_OK) return rc;
  PL_CHECK_ARG(grads, "pl_shard_push: NULL gradients");
  PL_CHECK_ARG(unit_begin >= 0 && unit_begin <= unit_end && unit_end <= c->push_units, "pl_shard_push: bad unit range");
  if (unit_begin == unit_end && !stats4) return PL_OK;
  int64_t blocks = ceil_div64(unit_end > unit_begin ? 2 * (unit_end - unit_begin) : 1, 256 * 4);
  if (blocks > 148 * 8) blocks = 148 * 8;
  shard_push_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(*c, grads, stats4, unit_begin, unit_end);
  PL_LAUNCH_CHECK("shard_push_kernel");
  return PL_OK;
}

extern "C" int pl_shard_push(const pl_shard_ctx* c, const float* grads, const float* stats4, void* stream) {
  if (!c) {
    pl::set_error("pl_shard_push: NULL context");
    return PL_ERR_BAD_ARG;
  }
  return pl_shard_push_range(c, grads, stats4, 0, c->push_units, stream);
}

extern "C" int pl_shard_barrier(const pl_shard_ctx* c, int32_t channel, void* stream) {
  using namespace pl;
  int rc = check_ctx(c, "pl_shard_barrier");
  if (rc != PL_OK) return rc;
  PL_CHECK_ARG(channel >= 0 && channel < 4, "pl_shard_barrier: channel must be 0..3");
  shard_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*c, channel);
  PL_LAUNCH_CHECK("shard_barrier_kernel");
  return PL_OK;
}

extern "C" int pl_shard_adam(const pl_shard_ctx* c, float* m, float* v, float kl_scale, float lr, float beta1, float beta2,
                             float eps, int32_t step, const float* sched_dev, double* kl_value, void* stream) {
  if (!c) {
    pl::set_error("pl_shard_adam: NULL context");
    return PL_ERR_BAD_ARG;
  }
  return pl_shard_adam_range(c, 0, c->num_segs, m, v, kl_scale, lr, beta1, beta2, eps, step, sched_dev, kl_value, stream);
}

extern "C" int pl_shard_adam_range(const pl_shard_ctx* c, int32_t seg_begin, int32_t seg_end, float* m, float* v,
                                   float kl_scale, float lr, float beta1, float beta2, float eps, int32_t step,
                                   const float* sched_dev, double* kl_value, void* stream) {
  using namespace pl;
  int rc = check_ctx(c, "pl_shard_adam");
  if (rc != PL_OK) return rc;
  PL_CHECK_ARG(m && v, "pl_shard_adam: NULL optimiser state");
  PL_CHECK_ARG(sched_dev || step >= 1, "pl_shard_adam: step must be >= 1");
  PL_CHECK_ARG(seg_begin >= 0 && seg_begin <= seg_end && seg_end <= c->num_segs, "pl_shard_adam: bad segment range");
  const int64_t u0 = c->adam_prefix[seg_begin], u1 = c->adam_prefix[seg_end];
  if (u1 == u0) return PL_OK;
  AdamHyper h;
  if (sched_dev) step = 1;
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  h.kl_scale = kl_scale;
  h.step_size = (float)((double)lr / bc1);
  h.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
  h.beta1 = beta1; h.beta2 = beta2; h.eps = eps;
  h.sched_dev = sched_dev;
  int64_t blocks = ceil_div64(u1 - u0, 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  shard_adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(*c, m, v, h, kl_value, u0, u1);
  PL_LAUNCH_CHECK("shard_adam_kernel");
  return PL_OK;
}

extern "C" int pl_shard_finish(const pl_shard_ctx* c, const double* kl_acc, double kl_const, double inv_num_samples,
                               float* stats_out4, void* stream) {
  using namespace pl;
  int rc = check_ctx(c, "pl_shard_finish");
  if (rc != PL_OK) return rc;
  shard_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(*c, kl_acc, kl_const, inv_num_samples, stats_out4);
  PL_LAUNCH_CHECK("shard_finish_kernel");
  return PL_OK;
}
