// tcgen05 / TMEM / TMA path of the BayesLinear contraction (PL_PREC_BF16).
//
//   forward   Y[mc]  = X[mc]  . W[mc]   + b[mc]          (src/ProbabilisticLayers.py:925-935)
//   backward  dX[mc] = dY[mc] . W[mc]^T                  (autograd of the above)
//             dW[mc] = X[mc]^T . dY[mc];  dmu = sum_mc dW[mc];  drho = sigmoid(rho) sum_mc dW[mc]*eps[mc]
//
// with W[mc] = mu + softplus(rho) * eps[mc] synthesised ON CHIP: the weight operand of the MMA is
// written by producer warps straight into the swizzled shared-memory layout tcgen05.mma reads
// (Philox counter -> Box-Muller -> FMA -> bf16), so sampled weights never exist in HBM, and the
// backward regenerates eps from the same counter instead of loading it.
//
// sampled_gemm_tc (forward and dX share it):
//   CTA tile  = [MT x 128 batch rows] x [128 output columns] for one MC sample; the generated
//               W tile (64 x 128 per k-block) is reused by all MT row sub-tiles, i.e. by the whole
//               batch (B = 512 -> MT = 4, 4 x 128 fp32 accumulator columns = the full 512-column TMEM)
//   warp 0    = TMA producer: activation tile (bf16, K-major, SWIZZLE_128B) -> smem ring A
//   warp 1    = TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, fp32 accumulate)
//   warps 2-9 = W-tile generators (ring B) during the K loop, then the TMEM -> HBM epilogue
//   rings are mbarrier pipelines; tcgen05.commit frees stages; ring depths are decoupled because
//   the activation stage is 4x larger than the generated stage.
//
// Roofline: tensor pipe (2*MC*B*I*O algorithmic FLOPs per launch); the secondary limit is the
// sampler's ALU/MUFU budget: every generated element feeds only 2*B flops.
#include <cuda.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace pl {

using namespace tc;

constexpr int kBN = 128;               // output-column tile (UMMA N)
constexpr int kBK = 64;                // k-block (one 128-byte swizzle row of bf16)
constexpr int kTcThreads = 320;        // 10 warps
constexpr int kGenWarps = 8;
constexpr int kSubTileBytes = 128 * kBK * 2;  // one [128 x 64] bf16 operand sub-tile = 16 KB

struct TcGemmParams {
  int mc, rows, ndim, kdim;     // rows = batch; C[mc] is [rows, ndim], reduction over kdim
  int w_in, w_out;              // weight matrix [I, O]
  int64_t a_mc_rows;            // rows per MC sample in the A tensor map (0 => broadcast input)
  const float* w_mu;            // [I, O]
  const float* w_sigma;         // softplus(rho), [I, O]
  const float* eps_w;           // optional injected eps [mc, I, O]
  Sampler sw;
  const float* bias;            // optional sampled bias [mc, ndim] (forward only)
  float* out;                   // [mc, rows, ndim]
};

// ---------------------------------------------------------------------------------------------
// HBM-bound pre-passes
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softplus_kernel(const float* __restrict__ rho,
                                                       float* __restrict__ sigma, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256)
    sigma[i] = softplus_f(rho[i]);
}

// fp32 -> bf16, 8 elements per thread per iteration (n % 8 == 0, 16-byte aligned)
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src,
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
__global__ void __launch_bounds__(256) sample_bias_kernel(const float* __restrict__ b_mu,
                                                          const float* __restrict__ b_rho,
                                                          const float* __restrict__ eps_b, Sampler sb,
                                                          int mc, int out, float* __restrict__ dst) {
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

// ---------------------------------------------------------------------------------------------
// fused sample + GEMM (forward: kDx = false, B operand MN-major; dX: kDx = true, B operand K-major)
// ---------------------------------------------------------------------------------------------
template <int MT>
struct TcCfg {
  static constexpr int kStagesA = MT == 4 ? 2 : 4;
  static constexpr int kStagesB = 4;
  static constexpr int kABytes = MT * kSubTileBytes;
  static constexpr int kSmemBytes = kStagesA * kABytes + kStagesB * kSubTileBytes + 1024 /*bias + barriers*/
                                    + 1024 /*alignment slack*/;
  static constexpr int kTmemCols = MT * kBN;  // 128, 256 or 512: powers of two
};

template <int MT, bool kDx, bool kInjected>
__global__ void __launch_bounds__(kTcThreads, 1)
sampled_gemm_tc(const __grid_constant__ CUtensorMap tmap_a, const TcGemmParams p) {
  using Cfg = TcCfg<MT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base;
  const uint32_t sB = sA + Cfg::kStagesA * Cfg::kABytes;
  const uint32_t sMisc = sB + Cfg::kStagesB * kSubTileBytes;
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kBN;
  const int m0 = blockIdx.y * (MT * 128);
  const int mc = blockIdx.z;
  const int num_kb = (p.kdim + kBK - 1) / kBK;

  if (threadIdx.x == 0) {
    for (int s = 0; s < Cfg::kStagesA; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < Cfg::kStagesB; ++s) {
      mbar_init(b_full(s), kGenWarps * 32);
      mbar_init(b_empty(s), 1);
    }
    mbar_init(acc_full, 1);
    fence_mbar_init();
    tma_prefetch_desc(&tmap_a);
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
        const int s = kb % Cfg::kStagesA, ph = (kb / Cfg::kStagesA) & 1;
        mbar_wait(a_empty(s), ph ^ 1);
        mbar_expect_tx(a_full(s), Cfg::kABytes);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
          tma_load_2d(sA + s * Cfg::kABytes + mt * kSubTileBytes, &tmap_a, a_full(s), kb * kBK,
                      row0 + mt * 128);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = idesc_bf16(128, kBN, /*a MN-major*/ 0, /*b MN-major*/ kDx ? 0 : 1);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int sa = kb % Cfg::kStagesA, pa = (kb / Cfg::kStagesA) & 1;
        const int sb = kb % Cfg::kStagesB, pb = (kb / Cfg::kStagesB) & 1;
        mbar_wait(a_full(sa), pa);
        mbar_wait(b_full(sb), pb);
        tc_fence_after_sync();
        const uint32_t a_addr = sA + sa * Cfg::kABytes;
        const uint32_t b_addr = sB + sb * kSubTileBytes;
#pragma unroll
        for (int ks = 0; ks < kBK / 16; ++ks) {
          // A: K-major SW128 (8-row groups 1024 B apart); advancing K by 16 bf16 = 32 B
          // B: forward MN-major SW128 (64-col groups 8192 B apart = LBO, 8-k groups 1024 B = SBO,
          //    16 k = 2 groups = 2048 B); dX K-major SW128 like A
          const uint64_t b_desc = kDx ? smem_desc_sw128(b_addr + ks * 32, 16, 1024)
                                      : smem_desc_sw128(b_addr + ks * 2048, 8192, 1024);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            const uint64_t a_desc = smem_desc_sw128(a_addr + mt * kSubTileBytes + ks * 32, 16, 1024);
            mma_bf16_ss(tmem_base + mt * kBN, a_desc, b_desc, idesc, (kb | ks) != 0);
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
    const int gw = warp - 2;  // 0..7
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % Cfg::kStagesB, ph = (kb / Cfg::kStagesB) & 1;
      const int k0 = kb * kBK;
      float4 mu[8], sg[8], ep[8];
      int wrow[8], wcol[8];
      uint32_t soff[8];
      // operand element (n, k): forward W[i = k0+k][o = n0+n]; dX W[i = n0+n][o = k0+k].
      // Each thread owns 8 groups of 4 consecutive o (one Philox call each).
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        if (!kDx) {
          const int k = gw + 8 * r, n = 4 * lane;          // k row, 4 consecutive n
          wrow[r] = k0 + k;
          wcol[r] = n0 + n;
          // MN-major SW128: [n/64]*8192 + [k/8]*1024 + (k%8)*128 + (((n%64)/8) ^ (k%8))*16 + (n%8)*2
          soff[r] = (uint32_t)((lane >> 4) * 8192 + r * 1024 + gw * 128 +
                               ((((lane & 15) >> 1) ^ gw) << 4) + ((lane & 1) << 3));
        } else {
          const int n = 16 * gw + 2 * r + (lane >> 4), k = 4 * (lane & 15);  // n row, 4 consecutive k
          wrow[r] = n0 + n;
          wcol[r] = k0 + k;
          // K-major SW128: n*128 + (((k/8) ^ (n%8)) * 16) + (k%8)*2
          soff[r] = (uint32_t)(n * 128 + ((((lane & 15) >> 1) ^ (n & 7)) << 4) + ((lane & 1) << 3));
        }
        const bool ok = wrow[r] < p.w_in && wcol[r] < p.w_out;
        const int64_t idx = (int64_t)wrow[r] * p.w_out + wcol[r];
        mu[r] = ok ? __ldg(reinterpret_cast<const float4*>(p.w_mu + idx)) : make_float4(0.f, 0.f, 0.f, 0.f);
        sg[r] = ok ? __ldg(reinterpret_cast<const float4*>(p.w_sigma + idx)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (kInjected)
          ep[r] = ok ? __ldg(reinterpret_cast<const float4*>(p.eps_w + (int64_t)mc * p.w_in * p.w_out + idx))
                     : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait(b_empty(s), ph ^ 1);
      const uint32_t dst = sB + s * kSubTileBytes;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float4 e;
        if (kInjected) e = ep[r];
        else e = eps4(p.sw, mc, wrow[r], wcol[r] >> 2);
        const uint32_t lo = pack_bf16x2(fmaf(sg[r].x, e.x, mu[r].x), fmaf(sg[r].y, e.y, mu[r].y));
        const uint32_t hi = pack_bf16x2(fmaf(sg[r].z, e.z, mu[r].z), fmaf(sg[r].w, e.w, mu[r].w));
        asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(dst + soff[r]), "r"(lo), "r"(hi) : "memory");
      }
      fence_proxy_async_smem();
      mbar_arrive(b_full(s));
    }
    // ===================== epilogue: TMEM -> registers -> HBM =====================
    mbar_wait(acc_full, 0);
    tc_fence_after_sync();
    const int q = warp & 3;        // TMEM lane quadrant this warp may access
    const int h = gw >> 2;         // which 64-column half of the tile
#pragma unroll 1
    for (int mt = 0; mt < MT; ++mt) {
      const int row = m0 + mt * 128 + q * 32 + lane;
#pragma unroll 1
      for (int cc = 0; cc < 2; ++cc) {
        const int col = h * 64 + cc * 32;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * kBN + col), v);
        tmem_ld_wait();
        if (row < p.rows) {
          float* o = p.out + ((int64_t)mc * p.rows + row) * p.ndim + n0 + col;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (n0 + col + j < p.ndim) {
              float4 w;
              w.x = __uint_as_float(v[j + 0]) + bias_s[col + j + 0];
              w.y = __uint_as_float(v[j + 1]) + bias_s[col + j + 1];
              w.z = __uint_as_float(v[j + 2]) + bias_s[col + j + 2];
              w.w = __uint_as_float(v[j + 3]) + bias_s[col + j + 3];
              *reinterpret_cast<float4*>(o + j) = w;
            }
          }
        }
      }
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, Cfg::kTmemCols);
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
static int make_tmap_bf16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols,
                          uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return PL_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {cols * 2};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box,
                  estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows %llu cols %llu)", (int)r,
              (unsigned long long)rows, (unsigned long long)cols);
    return PL_ERR_CUDA;
  }
  return PL_OK;
}

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct TcWorkspace {
  float* sigma;            // [I*O]
  __nv_bfloat16* a_bf16;   // x (forward) / dy (backward) as bf16
  float* bias;             // [mc*O]
  __nv_bfloat16* x_bf16;   // backward: x as bf16 (dW)
  size_t total;
};

static TcWorkspace carve(const pl_linear_args* a, void* base) {
  const size_t I = a->in_features, O = a->out_features, MC = a->mc, B = a->batch;
  const size_t x_rows = a->x_mc_stride == 0 ? B : MC * B;
  TcWorkspace w;
  size_t off = 0;
  uint8_t* b = (uint8_t*)base;
  w.sigma = (float*)(b + off); off += align256(I * O * 4);
  w.bias = (float*)(b + off); off += align256(MC * O * 4);
  // forward: x bf16; backward: dy bf16 (+ x bf16 for dW)
  const size_t big = MC * B * (I > O ? I : O) * 2;
  w.a_bf16 = (__nv_bfloat16*)(b + off); off += align256(big);
  w.x_bf16 = (__nv_bfloat16*)(b + off); off += align256(x_rows * I * 2);
  w.total = off;
  return w;
}

size_t linear_tc_workspace_bytes(const pl_linear_args* a) {
  if (!a) return 0;
  return carve(a, nullptr).total + 256;
}

static int tc_shape_ok(const pl_linear_args* a, const char* who) {
  if (a->param_mode != PL_PARAM_SHARED_RHO) {
    set_error("%s: PL_PREC_BF16 supports PL_PARAM_SHARED_RHO only", who);
    return PL_ERR_UNSUPPORTED;
  }
  if (a->in_features % 8 || a->out_features % 8 || a->in_features < 8 || a->out_features < 8) {
    set_error("%s: PL_PREC_BF16 needs in/out features to be multiples of 8 (got %d, %d)", who,
              a->in_features, a->out_features);
    return PL_ERR_UNSUPPORTED;
  }
  if (a->x_mc_stride != 0 && a->x_mc_stride != (int64_t)a->batch * a->in_features) {
    set_error("%s: PL_PREC_BF16 needs a dense [mc,batch,in] input (or mc stride 0)", who);
    return PL_ERR_UNSUPPORTED;
  }
  return PL_OK;
}

static inline int ew_grid(int64_t n, int per_block) {
  int64_t g = (n + per_block - 1) / per_block;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

template <bool kDx, bool kInj>
static int launch_gemm(const CUtensorMap& tm, const TcGemmParams& p, cudaStream_t st) {
  const int mt_rows = p.rows > 256 ? 512 : (p.rows > 128 ? 256 : 128);
  dim3 grid((unsigned)((p.ndim + kBN - 1) / kBN), (unsigned)((p.rows + mt_rows - 1) / mt_rows),
            (unsigned)p.mc);
#define PL_LAUNCH_MT(MT)                                                                            \
  do {                                                                                              \
    auto kern = sampled_gemm_tc<MT, kDx, kInj>;                                                     \
    PL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,                 \
                                 TcCfg<MT>::kSmemBytes));                                           \
    kern<<<grid, kTcThreads, TcCfg<MT>::kSmemBytes, st>>>(tm, p);                                    \
  } while (0)
  if (mt_rows == 512) PL_LAUNCH_MT(4);
  else if (mt_rows == 256) PL_LAUNCH_MT(2);
  else PL_LAUNCH_MT(1);
#undef PL_LAUNCH_MT
  PL_LAUNCH_CHECK("sampled_gemm_tc");
  return PL_OK;
}

int linear_bwd_simt_launch(const pl_linear_args* a);

int linear_fwd_tc_launch(const pl_linear_args* a) {
  PL_CHECK_ARG(a && a->x && a->w_mu && a->w_rho && a->y, "pl_linear_fwd: NULL pointer");
  PL_CHECK_ARG((a->b_mu == nullptr) == (a->b_rho == nullptr), "pl_linear_fwd: b_mu and b_rho must both be set or NULL");
  int rc = tc_shape_ok(a, "pl_linear_fwd");
  if (rc != PL_OK) return rc;
  if (a->mc == 0 || a->batch == 0) return PL_OK;
  PL_CHECK_ARG(a->mc <= 65535, "pl_linear_fwd: mc too large");
  if (!a->workspace || a->workspace_bytes < linear_tc_workspace_bytes(a)) {
    set_error("pl_linear_fwd: workspace %zu < %zu bytes", a->workspace_bytes, linear_tc_workspace_bytes(a));
    return PL_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)a->stream;
  TcWorkspace w = carve(a, (void*)(((uintptr_t)a->workspace + 255) & ~(uintptr_t)255));
  const int64_t I = a->in_features, O = a->out_features, MC = a->mc, B = a->batch;
  const int64_t x_rows = a->x_mc_stride == 0 ? B : MC * B;
  softplus_kernel<<<ew_grid(I * O, 1024), 256, 0, st>>>(a->w_rho, w.sigma, I * O);
  PL_LAUNCH_CHECK("softplus_kernel");
  cast_bf16_kernel<<<ew_grid(x_rows * I / 8, 1024), 256, 0, st>>>(a->x, w.a_bf16, x_rows * I / 8);
  PL_LAUNCH_CHECK("cast_bf16_kernel");
  if (a->b_mu) {
    sample_bias_kernel<<<ew_grid(MC * ((O + 3) / 4), 256), 256, 0, st>>>(
        a->b_mu, a->b_rho, a->eps_b, make_sampler(a->seed, a->layer_id, true, a->mc_global_offset),
        (int)MC, (int)O, w.bias);
    PL_LAUNCH_CHECK("sample_bias_kernel");
  }
  CUtensorMap tm;
  rc = make_tmap_bf16(&tm, w.a_bf16, (uint64_t)x_rows, (uint64_t)I, 128, kBK);
  if (rc != PL_OK) return rc;
  TcGemmParams p;
  p.mc = (int)MC; p.rows = (int)B; p.ndim = (int)O; p.kdim = (int)I;
  p.w_in = (int)I; p.w_out = (int)O;
  p.a_mc_rows = a->x_mc_stride == 0 ? 0 : B;
  p.w_mu = a->w_mu; p.w_sigma = w.sigma; p.eps_w = a->eps_w;
  p.sw = make_sampler(a->seed, a->layer_id, false, a->mc_global_offset);
  p.bias = a->b_mu ? w.bias : nullptr;
  p.out = a->y;
  return a->eps_w ? launch_gemm<false, true>(tm, p, st) : launch_gemm<false, false>(tm, p, st);
}

int linear_bwd_tc_launch(const pl_linear_args* a) {
  PL_CHECK_ARG(a && a->x && a->w_mu && a->w_rho && a->dy, "pl_linear_bwd: NULL pointer");
  int rc = tc_shape_ok(a, "pl_linear_bwd");
  if (rc != PL_OK) return rc;
  if (a->mc == 0 || a->batch == 0) return linear_bwd_simt_launch(a);
  if (!a->workspace || a->workspace_bytes < linear_tc_workspace_bytes(a)) {
    set_error("pl_linear_bwd: workspace %zu < %zu bytes", a->workspace_bytes, linear_tc_workspace_bytes(a));
    return PL_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)a->stream;
  TcWorkspace w = carve(a, (void*)(((uintptr_t)a->workspace + 255) & ~(uintptr_t)255));
  const int64_t I = a->in_features, O = a->out_features, MC = a->mc, B = a->batch;
  if (a->dx) {
    softplus_kernel<<<ew_grid(I * O, 1024), 256, 0, st>>>(a->w_rho, w.sigma, I * O);
    PL_LAUNCH_CHECK("softplus_kernel");
    cast_bf16_kernel<<<ew_grid(MC * B * O / 8, 1024), 256, 0, st>>>(a->dy, w.a_bf16, MC * B * O / 8);
    PL_LAUNCH_CHECK("cast_bf16_kernel");
    CUtensorMap tm;
    rc = make_tmap_bf16(&tm, w.a_bf16, (uint64_t)(MC * B), (uint64_t)O, 128, kBK);
    if (rc != PL_OK) return rc;
    TcGemmParams p;
    p.mc = (int)MC; p.rows = (int)B; p.ndim = (int)I; p.kdim = (int)O;
    p.w_in = (int)I; p.w_out = (int)O;
    p.a_mc_rows = B;
    p.w_mu = a->w_mu; p.w_sigma = w.sigma; p.eps_w = a->eps_w;
    p.sw = make_sampler(a->seed, a->layer_id, false, a->mc_global_offset);
    p.bias = nullptr;
    p.out = a->dx;
    rc = a->eps_w ? launch_gemm<true, true>(tm, p, st) : launch_gemm<true, false>(tm, p, st);
    if (rc != PL_OK) return rc;
  }
  if (a->dw_mu || a->db_mu) {
    // weight / bias gradients: fp32 kernels for now (tcgen05 dW kernel is the next step)
    pl_linear_args b = *a;
    b.dx = nullptr;
    b.precision = PL_PREC_FP32;
    rc = linear_bwd_simt_launch(&b);
    if (rc != PL_OK) return rc;
  }
  return PL_OK;
}

}  // namespace pl
