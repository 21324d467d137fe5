// MC-first cross entropy + accuracy + gradient in one pass over the logits.
//
// Replaces MC_CrossEntropyLoss.forward (src/ProbabilisticLayers.py:32-43: cross_entropy of the
// [MC*B, C] logits against target.repeat(MC), 'mean', times num_samples), MC_Accuracy (:46-90: argmax
// + .item() host sync) and the autograd backward of both (log_softmax, nll_loss and their backward
// kernels: ~10 launches and five passes over the [MC,B,C] tensor).
//
// Roofline: HBM.  Algorithmic bytes: 4*MC*B*C read (+ 4*MC*B*C written when the gradient is wanted); the
// second sweep over a row re-reads it from L1 (a row is C*4 bytes, one warp owns it).
// One warp per row, 128-bit loads when C % 4 == 0, the row kept in registers when C <= 1024, warp shuffles only;
// per-block partial sums in fp64, summed in block order by the block that draws the last ticket
// (deterministic, single launch; the workspace is zero on entry and left zero).
#include "common.cuh"

namespace pl {

constexpr int kCeThreads = 256;
constexpr int kCeWarps = kCeThreads / 32;
constexpr int kCeMaxBlocks = 148 * 8;

constexpr int kCeRegVec = 8;   // rows of up to 32*4*8 = 1024 classes are held in registers between the sweeps

__device__ __forceinline__ void argmax_push(float x, int i, float& best, int& best_i) {
  if (x > best) {     // indices arrive in increasing order within a lane: the lowest index wins ties
    best = x;
    best_i = i;
  }
}

// kMode 0: scalar loads, any C (three sweeps: max/argmax, sum-exp, gradient; sweeps 2-3 hit L1)
// kMode 1: 128-bit loads, any C % 4 == 0 (same three sweeps)
// kMode 2: 128-bit loads, C % 4 == 0 and C <= 1024: the row lives in registers, one read of the logits
template <int kMode>
__global__ void __launch_bounds__(kCeThreads)
ce_kernel(const float* __restrict__ pred, const int64_t* __restrict__ target, int64_t rows, int batch,
          int classes, float grad_scale, float* __restrict__ dpred, double* __restrict__ partials,
          unsigned int* __restrict__ ticket, float* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n4 = classes >> 2;
  double loss_acc = 0.0, hit_acc = 0.0;
  for (int64_t row = (int64_t)blockIdx.x * kCeWarps + warp; row < rows; row += (int64_t)gridDim.x * kCeWarps) {
    const float* x = pred + row * classes;
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const int tgt = (int)target[row % batch];
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    float4 v[kMode == 2 ? kCeRegVec : 1];
    // ---- sweep 1: max + argmax ----
    if (kMode == 2) {
#pragma unroll
      for (int j = 0; j < kCeRegVec; ++j) {
        const int i = lane + 32 * j;
        v[j] = i < n4 ? __ldg(x4 + i) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        argmax_push(v[j].x, 4 * i, best, best_i);
        argmax_push(v[j].y, 4 * i + 1, best, best_i);
        argmax_push(v[j].z, 4 * i + 2, best, best_i);
        argmax_push(v[j].w, 4 * i + 3, best, best_i);
      }
    } else if (kMode == 1) {
      for (int i = lane; i < n4; i += 32) {
        const float4 t = __ldg(x4 + i);
        argmax_push(t.x, 4 * i, best, best_i);
        argmax_push(t.y, 4 * i + 1, best, best_i);
        argmax_push(t.z, 4 * i + 2, best, best_i);
        argmax_push(t.w, 4 * i + 3, best, best_i);
      }
    } else {
      for (int i = lane; i < classes; i += 32) argmax_push(__ldg(x + i), i, best, best_i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float b = __shfl_xor_sync(0xffffffffu, best, o);
      const int bi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (b > best || (b == best && bi < best_i)) {
        best = b;
        best_i = bi;
      }
    }
    const float m = best;
    // ---- sweep 2: sum of exp(x - max) ----
    float sum = 0.f;
    if (kMode == 2) {
#pragma unroll
      for (int j = 0; j < kCeRegVec; ++j) {   // exp(-inf) = 0 for the padding
        v[j].x = __expf(v[j].x - m); v[j].y = __expf(v[j].y - m);
        v[j].z = __expf(v[j].z - m); v[j].w = __expf(v[j].w - m);
        sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
      }
    } else if (kMode == 1) {
      for (int i = lane; i < n4; i += 32) {
        const float4 t = x4[i];
        sum += (__expf(t.x - m) + __expf(t.y - m)) + (__expf(t.z - m) + __expf(t.w - m));
      }
    } else {
      for (int i = lane; i < classes; i += 32) sum += __expf(x[i] - m);
    }
    sum = warp_sum(sum);
    const bool tgt_ok = tgt >= 0 && tgt < classes;
    if (lane == 0) {
      loss_acc += tgt_ok ? (double)(m + __logf(sum) - __ldg(x + tgt)) : 0.0;
      hit_acc += (best_i == tgt) ? 1.0 : 0.0;
    }
    // ---- sweep 3: (softmax - onehot) * grad_scale ----
    if (dpred) {
      float* d = dpred + row * classes;
      float4* d4 = reinterpret_cast<float4*>(d);
      const float inv = __fdividef(grad_scale, sum);
      if (kMode == 2) {
#pragma unroll
        for (int j = 0; j < kCeRegVec; ++j) {
          const int i = lane + 32 * j;
          if (i < n4) {
            float4 g = make_float4(v[j].x * inv, v[j].y * inv, v[j].z * inv, v[j].w * inv);
            if ((tgt >> 2) == i) {
              const int c = tgt & 3;
              if (c == 0) g.x -= grad_scale; else if (c == 1) g.y -= grad_scale;
              else if (c == 2) g.z -= grad_scale; else g.w -= grad_scale;
            }
            d4[i] = g;
          }
        }
      } else if (kMode == 1) {
        for (int i = lane; i < n4; i += 32) {
          const float4 t = x4[i];
          float4 g = make_float4(__expf(t.x - m) * inv, __expf(t.y - m) * inv, __expf(t.z - m) * inv,
                                 __expf(t.w - m) * inv);
          if ((tgt >> 2) == i) {
            const int c = tgt & 3;
            if (c == 0) g.x -= grad_scale; else if (c == 1) g.y -= grad_scale;
            else if (c == 2) g.z -= grad_scale; else g.w -= grad_scale;
          }
          d4[i] = g;
        }
      } else {
        for (int i = lane; i < classes; i += 32)
          d[i] = __expf(x[i] - m) * inv - (i == tgt ? grad_scale : 0.f);
      }
    }
  }
  // block partials -> ordered final sum by the last block
  __shared__ double sh[2][kCeWarps];
  __shared__ bool last;
  if (lane == 0) {
    sh[0][warp] = loss_acc;
    sh[1][warp] = hit_acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double l = 0.0, h = 0.0;
    for (int w = 0; w < kCeWarps; ++w) {
      l += sh[0][w];
      h += sh[1][w];
    }
    partials[2 * blockIdx.x] = l;
    partials[2 * blockIdx.x + 1] = h;
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && warp == 0) {
    __threadfence();
    double l = 0.0, h = 0.0;
    for (int b = lane; b < (int)gridDim.x; b += 32) {   // fixed assignment and order: deterministic
      l += partials[2 * b];
      h += partials[2 * b + 1];
    }
    l = warp_sum(l);
    h = warp_sum(h);
    if (lane == 0) {
      out[0] = (float)l;
      out[1] = (float)h;
      *ticket = 0u;   // self-resetting: the next launch on this workspace starts from zero again
    }
  }
}

}  // namespace pl

extern "C" size_t pl_mc_cross_entropy_workspace_bytes(void) {
  return sizeof(double) * 2 * pl::kCeMaxBlocks + 64;
}

extern "C" int pl_mc_cross_entropy(const pl_ce_args* a) {
  using namespace pl;
  PL_CHECK_ARG(a, "pl_mc_cross_entropy: NULL args");
  PL_CHECK_ARG(a->mc >= 0 && a->batch >= 0 && a->classes >= 1, "pl_mc_cross_entropy: bad sizes");
  PL_CHECK_ARG(a->out && a->workspace, "pl_mc_cross_entropy: NULL pointer");
  if (a->workspace_bytes < pl_mc_cross_entropy_workspace_bytes()) {
    set_error("pl_mc_cross_entropy: workspace %zu < %zu bytes", a->workspace_bytes,
              pl_mc_cross_entropy_workspace_bytes());
    return PL_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)a->stream;
  const int64_t rows = (int64_t)a->mc * a->batch;
  if (rows == 0) {
    PL_CUDA(cudaMemsetAsync(a->out, 0, 2 * sizeof(float), st));
    return PL_OK;
  }
  PL_CHECK_ARG(a->pred && a->target, "pl_mc_cross_entropy: NULL pointer");
  double* partials = (double*)a->workspace;
  unsigned int* ticket = (unsigned int*)((char*)a->workspace + sizeof(double) * 2 * kCeMaxBlocks);
  int64_t g = (rows + kCeWarps - 1) / kCeWarps;
  const int grid = (int)(g > kCeMaxBlocks ? kCeMaxBlocks : g);
  const bool vec = (a->classes % 4 == 0) && aligned16(a->pred) && (!a->dpred || aligned16(a->dpred));
#define PL_CE_LAUNCH(MODE)                                                                              \
  ce_kernel<MODE><<<grid, kCeThreads, 0, st>>>(a->pred, a->target, rows, a->batch, a->classes, a->grad_scale, \
                                               a->dpred, partials, ticket, a->out)
  if (vec && a->classes <= 128 * kCeRegVec) PL_CE_LAUNCH(2);
  else if (vec) PL_CE_LAUNCH(1);
  else PL_CE_LAUNCH(0);
#undef PL_CE_LAUNCH
  PL_LAUNCH_CHECK("ce_kernel");
  return PL_OK;
}
