// BayesConv2d on the tensor-core kernels (PL_PREC_BF16).
//
// Per MC sample the convolution is the GEMM  Y[M = B*H'*W', Cout] = A[M, K = Cin*k*k] . Wc[Cout, K]^T
// with Wc[mc] = mu + softplus(rho)*eps[mc] synthesised on chip (Philox row = cout, col = cin*k*k +
// kh*k + kw, the same stream the fp32 kernels and the oracle use).  This round the im2col matrix A is
// materialised once per call as bf16 (an HBM-bound pre-pass; TMA-im2col = next step) and the three
// GEMMs reuse the Linear kernels of tc_kernels.cuh with the roles mapped as
//     forward  Y   = A   . Wc^T      -> sampled_gemm_tc<kDx = true>   (w_in = Cout, w_out = K),
//                                       epilogue writes NCHW directly (out_hw = H'*W')
//     dWc      dWc = dYt^T . A       -> dw_tc (x-operand = dY pixel-major, dy-operand = A), reduction
//                                       over pixels split across CTAs when Cout*K has few tiles
//     dX       dA  = dYt . Wc        -> sampled_gemm_tc<kDx = false>, written K-major, then col2im
// Replaces src/ProbabilisticLayers.py:1032-1057 (F.conv2d(groups=MC) + two permute copies) and its
// autograd; parity: tests/test_gpu_tc.py::test_tc_conv_*.
#include "tc_kernels.cuh"

namespace pl {

struct ConvGeom {
  int mc, mcx, batch, cin, cout, h, w, k, stride, pad, dil, ho, wo;
  int K, Kp, Coutp;  // K = cin*k*k; Kp, Coutp = rounded up to 8 (16-byte bf16 rows for TMA)
  int64_t M;         // batch*ho*wo output pixels per MC sample
};

// A[mcx][m][Kp] (bf16) = im2col(x); one thread per 8 consecutive K entries (16-byte store)
__global__ void __launch_bounds__(256) im2col_bf16_kernel(const float* __restrict__ x,
                                                          int64_t x_mc_stride,
                                                          __nv_bfloat16* __restrict__ A, ConvGeom g) {
  const int chunks = g.Kp >> 3, kk2 = g.k * g.k, hw = g.ho * g.wo;
  const int64_t total = (int64_t)g.mcx * g.M * chunks;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int kc = (int)(t % chunks);
    const int64_t mm = t / chunks;
    const int m = (int)(mm % g.M), mc = (int)(mm / g.M);
    const int b = m / hw, pix = m - b * hw, oh = pix / g.wo, ow = pix - oh * g.wo;
    const int ih0 = oh * g.stride - g.pad, iw0 = ow * g.stride - g.pad;
    const float* xb = x + (int64_t)mc * x_mc_stride + (int64_t)b * g.cin * g.h * g.w;
    int kidx = kc << 3;
    int c = kidx / kk2, r = kidx - c * kk2, kh = r / g.k, kw = r - kh * g.k;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float val = 0.f;
      if (kidx + j < g.K) {
        const int ih = ih0 + kh * g.dil, iw = iw0 + kw * g.dil;
        if (ih >= 0 && ih < g.h && iw >= 0 && iw < g.w) val = __ldg(xb + ((int64_t)c * g.h + ih) * g.w + iw);
      }
      v[j] = val;
      if (++kw == g.k) {
        kw = 0;
        if (++kh == g.k) {
          kh = 0;
          ++c;
        }
      }
    }
    *reinterpret_cast<uint4*>(A + t * 8) = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]),
                                                      pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
  }
}

// Same matrix, one CTA per (image, `rpb` consecutive output rows).  The (rpb-1)*stride + (k-1)*dil + 1 input rows
// those output rows need are staged in shared memory WITH their zero padding ([Cin*rin][ws] fp32, ws = w + 2*pad;
// 16-byte global loads when the rows allow it), followed by a zero region that serves the K..Kp tail, next to two
// tables: off[kidx] = byte offset of (c, kh, kw) in the staging, pshift[px] = byte shift of output pixel px.
// A lane keeps the eight offsets of its 16-byte output chunk in registers and walks over the CTA's pixels: per
// chunk one broadcast table read, eight (add, shared load) pairs, four packs and one 16-byte store - 25 instructions
// where the first version (recomputing (c, kh, kw) and the bounds per element) needed ~430 and the second ~70 plus
// a scalar staging loop that cost as much again (profiles/r01_experiments.md, r02_experiments.md).
// grid (ceil(ho / rpb), mcx*batch); dynamic smem = im2col_rows_smem().
// kK: compile-time kernel size (divisions by k and k*k become multiply-shift); 0 = run-time k.
template <int kK>
__global__ void __launch_bounds__(256) im2col_rows_bf16_kernel(const float* __restrict__ x,
                                                               int64_t x_mc_stride,
                                                               __nv_bfloat16* __restrict__ A, ConvGeom g, int rpb,
                                                               int vec4) {
  extern __shared__ __align__(16) uint8_t im2col_smem[];
  const int k = kK ? kK : g.k;
  const int ws = g.w + 2 * g.pad;
  const int rin = (rpb - 1) * g.stride + (k - 1) * g.dil + 1;   // staged input rows per channel
  const int nrow = g.cin * rin;
  const int zr = (rpb - 1) * g.stride * ws + (g.wo - 1) * g.stride + 1;   // zero region: any pixel shift stays inside
  const int nfloat = (nrow * ws + zr + 3) & ~3;
  uint32_t* off = reinterpret_cast<uint32_t*>(im2col_smem);                                   // [Kp]
  uint32_t* pshift = off + ((g.Kp + 3) & ~3);                                                // [rpb * wo]
  float* rows = reinterpret_cast<float*>(pshift + ((rpb * g.wo + 3) & ~3));                   // [nfloat]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int oh0 = blockIdx.x * rpb, img = blockIdx.y, mc = img / g.batch, b = img - mc * g.batch;
  const float* xb = x + (int64_t)mc * x_mc_stride + (int64_t)b * g.cin * g.h * g.w;
  const int ih0 = oh0 * g.stride - g.pad;
  for (int t = threadIdx.x; t < (nfloat >> 2); t += 256) reinterpret_cast<uint4*>(rows)[t] = make_uint4(0u, 0u, 0u, 0u);
  const int kk2 = k * k;
  for (int kidx = threadIdx.x; kidx < g.Kp; kidx += 256) {
    int o = nrow * ws;                                   // the zero region
    if (kidx < g.K) {
      const int c = kidx / kk2, r = kidx - c * kk2, kh = r / k, kw = r - kh * k;
      o = (c * rin + kh * g.dil) * ws + kw * g.dil;
    }
    off[kidx] = (uint32_t)o * 4u;
  }
  for (int px = threadIdx.x; px < rpb * g.wo; px += 256) {
    const int j = px / g.wo, ow = px - j * g.wo;
    pshift[px] = (uint32_t)(j * g.stride * ws + ow * g.stride) * 4u;
  }
  __syncthreads();
  if (vec4) {   // rows of w % 4 == 0 floats, 16-byte aligned: one 16-byte load per four columns
    const int wq = g.w >> 2;
    for (int u = threadIdx.x; u < nrow * wq; u += 256) {
      const int rw = u / wq, q = u - rw * wq;
      const int c = rw / rin, ih = ih0 + (rw - c * rin);
      if (ih < 0 || ih >= g.h) continue;
      const float4 v = __ldg(reinterpret_cast<const float4*>(xb + ((int64_t)c * g.h + ih) * g.w) + q);
      float* d = rows + rw * ws + g.pad + (q << 2);
      d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
  } else {
    for (int u = threadIdx.x; u < nrow * g.w; u += 256) {
      const int rw = u / g.w, col = u - rw * g.w;
      const int c = rw / rin, ih = ih0 + (rw - c * rin);
      if (ih < 0 || ih >= g.h) continue;
      rows[rw * ws + g.pad + col] = __ldg(xb + ((int64_t)c * g.h + ih) * g.w + col);
    }
  }
  __syncthreads();
  const int chunks = g.Kp >> 3;
  // lanes per pixel: 32 when a row has >= 32 chunks, else the next power of two (narrow rows share a warp pass)
  int cpl = 32;
  while ((cpl >> 1) >= chunks) cpl >>= 1;
  const int ppw = 32 / cpl, sub = lane / cpl, kc0 = lane - sub * cpl;
  const int npix = min(rpb, g.ho - oh0) * g.wo;           // output pixels of this CTA (last block may be short)
  __nv_bfloat16* dst0 = A + (((int64_t)mc * g.batch + b) * g.ho * g.wo + (int64_t)oh0 * g.wo) * (int64_t)g.Kp;
  const uint8_t* rb = reinterpret_cast<const uint8_t*>(rows);
  for (int kc = kc0; kc < chunks; kc += cpl) {
    const uint4 oa = *reinterpret_cast<const uint4*>(off + (kc << 3));
    const uint4 ob = *reinterpret_cast<const uint4*>(off + (kc << 3) + 4);
    __nv_bfloat16* dst = dst0 + (kc << 3);
#pragma unroll 2
    for (int px = warp * ppw + sub; px < npix; px += 8 * ppw) {
      const uint8_t* r = rb + pshift[px];
      const float v0 = *reinterpret_cast<const float*>(r + oa.x), v1 = *reinterpret_cast<const float*>(r + oa.y);
      const float v2 = *reinterpret_cast<const float*>(r + oa.z), v3 = *reinterpret_cast<const float*>(r + oa.w);
      const float v4 = *reinterpret_cast<const float*>(r + ob.x), v5 = *reinterpret_cast<const float*>(r + ob.y);
      const float v6 = *reinterpret_cast<const float*>(r + ob.z), v7 = *reinterpret_cast<const float*>(r + ob.w);
      *reinterpret_cast<uint4*>(dst + (int64_t)px * g.Kp) =
          make_uint4(pack_bf16x2(v0, v1), pack_bf16x2(v2, v3), pack_bf16x2(v4, v5), pack_bf16x2(v6, v7));
    }
  }
}

static size_t im2col_rows_smem(const ConvGeom& g, int rpb) {
  const size_t ws = (size_t)(g.w + 2 * g.pad);
  const size_t rin = (size_t)(rpb - 1) * g.stride + (size_t)(g.k - 1) * g.dil + 1;
  const size_t zr = (size_t)(rpb - 1) * g.stride * ws + (size_t)(g.wo - 1) * g.stride + 1;
  const size_t nfloat = (g.cin * rin * ws + zr + 3) & ~(size_t)3;
  return ((((size_t)g.Kp + 3) & ~(size_t)3) + (((size_t)rpb * g.wo + 3) & ~(size_t)3) + nfloat) * 4;
}

// output rows per CTA: as many (<= 4) as keep the staging + tables under ~100 KB (two CTAs per SM)
static int im2col_rpb(const ConvGeom& g) {
  for (int rpb = 4; rpb > 1; rpb >>= 1)
    if (im2col_rows_smem(g, rpb) <= 100 * 1024 && g.ho >= rpb) return rpb;
  return 1;
}

static int launch_im2col(const pl_conv2d_args* a, const ConvGeom& g, __nv_bfloat16* A, cudaStream_t st) {
  const int rpb = im2col_rpb(g);
  const size_t smem = im2col_rows_smem(g, rpb);
  // the row kernel needs the staged rows + tables in shared memory and the standard output width;
  // PL_IM2COL_GENERIC=1 (read per call: a test flips it) forces the element-wise kernel
  const char* generic = getenv("PL_IM2COL_GENERIC");
  const bool rows_ok = smem <= 200 * 1024 && (int64_t)g.mcx * g.batch <= 65535 &&
                       (g.wo - 1) * g.stride + (g.k - 1) * g.dil < g.w + 2 * g.pad &&
                       !(generic && generic[0] == '1');
  if (rows_ok) {
    const int vec4 = (g.w % 4 == 0) && ((uintptr_t)a->x % 16 == 0) && (a->x_mc_stride % 4 == 0);
    dim3 grid((unsigned)((g.ho + rpb - 1) / rpb), (unsigned)(g.mcx * g.batch));
#define PL_IM2COL(KK)                                                                                      \
  do {                                                                                                    \
    if (smem > 48 * 1024)   /* per call: the attribute is per device, a process-wide cache would miss GPU #2 */ \
      PL_CUDA(cudaFuncSetAttribute(im2col_rows_bf16_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                   (int)smem));                                                           \
    im2col_rows_bf16_kernel<KK><<<grid, 256, smem, st>>>(a->x, a->x_mc_stride, A, g, rpb, vec4);          \
  } while (0)
    if (g.k == 5) PL_IM2COL(5);
    else if (g.k == 3) PL_IM2COL(3);
    else PL_IM2COL(0);
#undef PL_IM2COL
    PL_LAUNCH_CHECK("im2col_rows_bf16_kernel");
  } else {
    im2col_bf16_kernel<<<ew_grid((int64_t)g.mcx * g.M * (g.Kp / 8), 256), 256, 0, st>>>(a->x, a->x_mc_stride, A, g);
    PL_LAUNCH_CHECK("im2col_bf16_kernel");
  }
  return PL_OK;
}

// dy [mc*B, Cout, HW] fp32 -> dYt [mc*B*HW, Coutp] bf16 (pixel-major rows = GEMM operand), zero padded
// grid (mc*B, ceil(Coutp/32), ceil(HW/32)), block (32, 8)   (the image index rides on grid.x: no 65535 limit)
__global__ void __launch_bounds__(256) dy_pixel_major_kernel(const float* __restrict__ dy,
                                                             __nv_bfloat16* __restrict__ dyt, int cout,
                                                             int coutp, int hw) {
  __shared__ float tile[32][33];
  const int img = blockIdx.x, p0 = blockIdx.z * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = c0 + ty + 8 * i, pix = p0 + tx;
    tile[ty + 8 * i][tx] = (co < cout && pix < hw) ? __ldg(dy + ((int64_t)img * cout + co) * hw + pix) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pix = p0 + ty + 8 * i, co = c0 + tx;
    if (pix < hw && co < coutp)
      dyt[((int64_t)img * hw + pix) * coutp + co] = __float2bfloat16(tile[tx][ty + 8 * i]);
  }
}

// dx[mc][b][c][h][w] = sum_{kh,kw} dA[mc][c*k*k + kh*k + kw][b*HWo + oh*Wo + ow]   (gather, no atomics)
// dA is bf16: it is written once by the GEMM epilogue and read once here, and its k*k overlapping contributions
// per input pixel are summed in fp32 (halves the largest transient of the conv backward: 5 GB at C3's layer 2)
template <bool kUnit>   // kUnit: stride == 1 and dilation == 1 (no divisibility tests / divisions)
__global__ void __launch_bounds__(256) col2im_kernel(const __nv_bfloat16* __restrict__ dA, float* __restrict__ dx,
                                                     ConvGeom g) {
  const int64_t per_mc = (int64_t)g.batch * g.cin * g.h * g.w, total = (int64_t)g.mc * per_mc;
  const int kk2 = g.k * g.k, hwo = g.ho * g.wo;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < total; t += (int64_t)gridDim.x * 256) {
    const int iw = (int)(t % g.w);
    int64_t r = t / g.w;
    const int ih = (int)(r % g.h);
    r /= g.h;
    const int c = (int)(r % g.cin);
    r /= g.cin;
    const int b = (int)(r % g.batch), mc = (int)(r / g.batch);
    const __nv_bfloat16* base = dA + ((int64_t)mc * g.K + (int64_t)c * kk2) * g.M + (int64_t)b * hwo;
    float s = 0.f;
    for (int kh = 0; kh < g.k; ++kh) {
      const int th = ih + g.pad - (kUnit ? kh : kh * g.dil);
      if (th < 0 || (!kUnit && th % g.stride)) continue;
      const int oh = kUnit ? th : th / g.stride;
      if (oh >= g.ho) continue;
      for (int kw = 0; kw < g.k; ++kw) {
        const int tw = iw + g.pad - (kUnit ? kw : kw * g.dil);
        if (tw < 0 || (!kUnit && tw % g.stride)) continue;
        const int ow = kUnit ? tw : tw / g.stride;
        if (ow >= g.wo) continue;
        s += __bfloat162float(base[(int64_t)(kh * g.k + kw) * g.M + oh * g.wo + ow]);
      }
    }
    dx[t] = s;
  }
}

// Pipelined col2im for stride 1 / dilation 1.  A CTA owns `cpc` consecutive input channels of `nb` consecutive images
// of one MC sample and walks over the channels: the k*k dA rows of a channel (each nb*ho*wo bf16, contiguous in the
// K-major dA) arrive by 1-D bulk copies (cp.async.bulk + mbarrier) in one of two shared-memory stages while the
// threads gather the previous channel from the other, so loads never stop (the first tiled version loaded, synced
// and gathered in turn; the element-wise kernel above reads dA with 2-byte loads).  CTA order (channel- or
// image-major), compile-time k and mask-predicated taps all measured the same (profiles/r02_experiments.md).
// Several images per CTA keep the copied pieces >= 512 bytes on small planes.  Same summation order (kh, kw
// ascending) as col2im_kernel: bit-identical results.
// grid (ceil(cin / cpc), ceil(batch / bpc), mc); dynamic smem = 2 * k*k * bpc*ho*wo * 2 + 16 bytes;
// needs (ho*wo) % 8 == 0 (16-byte aligned pieces; M = batch*ho*wo is then a multiple of 8 as well).
// One tap of the unrolled gather: returns the bf16 at shared address `addr` as fp32 bits when bit 0 of `mask` is set,
// else +0; then shifts the mask and advances the address (inside the statement, so the compiler does not precompute
// 25 addresses and predicates and spill them).  Only the load is predicated; every temporary is defined
// unconditionally.  The branchy C loop costs ~18 instructions per tap (ncu: 700 thread instructions per output,
// issue slots 92 % busy).
__device__ __forceinline__ float lds_bf16_tap(uint32_t& addr, uint32_t& mask, uint32_t step) {
  uint32_t bits;
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b16 h;\n\t.reg .b32 w;\n\t"
      "and.b32 w, %2, 1;\n\t"
      "setp.ne.u32 p, w, 0;\n\t"
      "mov.b16 h, 0;\n\t"
      "@p ld.shared.u16 h, [%1];\n\t"
      "cvt.u32.u16 w, h;\n\t"
      "shl.b32 %0, w, 16;\n\t"
      "shr.u32 %2, %2, 1;\n\t"
      "add.u32 %1, %1, %3;\n\t}"
      : "=r"(bits), "+r"(addr), "+r"(mask)
      : "r"(step));
  return __uint_as_float(bits);
}

// kK: compile-time kernel size (3, 5): the k*k taps unroll into predicated (load, convert, add) triples driven by a
// k*k-bit validity mask, addresses advance by constant steps; 0 = run-time k, plain loops.
template <int kK>
__global__ void __launch_bounds__(256, 6) col2im_pipe_kernel(const __nv_bfloat16* __restrict__ dA,
                                                             float* __restrict__ dx, ConvGeom g, int cpc, int bpc) {
  extern __shared__ __align__(16) uint8_t col2im_smem[];
  const int k = kK ? kK : g.k;
  const int kk2 = k * k, hwo = g.ho * g.wo, hw = g.h * g.w;
  const int c0 = blockIdx.x * cpc, b0 = blockIdx.y * bpc, mc = blockIdx.z;
  const int nc = min(cpc, g.cin - c0), nb = min(bpc, g.batch - b0);
  const int piece = nb * hwo;                                   // bf16 elements per (channel, tap) row of this CTA
  const uint32_t stage_bytes = (uint32_t)kk2 * piece * 2u;
  __nv_bfloat16* planes = reinterpret_cast<__nv_bfloat16*>(col2im_smem);          // [2][kk2][piece]
  uint64_t* bars = reinterpret_cast<uint64_t*>(col2im_smem + 2 * (size_t)kk2 * bpc * hwo * 2);
  const uint32_t bar0 = smem_u32(bars), planes0 = smem_u32(planes);
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const __nv_bfloat16* base = dA + ((int64_t)mc * g.K + (int64_t)c0 * kk2) * g.M + (int64_t)b0 * hwo;
  auto issue = [&](int ci, int s) {     // warp 0: the k*k rows of channel c0 + ci -> stage s
    if (threadIdx.x == 0) mbar_expect_tx(bar0 + 8 * s, stage_bytes);
    __syncwarp();
    for (int r = threadIdx.x; r < kk2; r += 32)
      bulk_load_1d(planes0 + (uint32_t)s * stage_bytes + (uint32_t)r * piece * 2u,
                   base + ((int64_t)ci * kk2 + r) * g.M, (uint32_t)piece * 2u, bar0 + 8 * s);
  };
  if (threadIdx.x < 32) issue(0, 0);
#pragma unroll 1
  for (int ci = 0; ci < nc; ++ci) {
    const int s = ci & 1;
    if (threadIdx.x < 32 && ci + 1 < nc) issue(ci + 1, s ^ 1);   // stage s^1 was drained before the sync below
    mbar_wait(bar0 + 8 * s, (ci >> 1) & 1);
#pragma unroll 1
    for (int t = threadIdx.x; t < nb * hw; t += 256) {
      const int bl = t / hw, p = t - bl * hw, ih = p / g.w, iw = p - ih * g.w;
      float sum = 0.f;
      if (kK) {
        // tap (kh, kw) reads plane[kh*k + kw][ih + pad - kh][iw + pad - kw]: valid = row bit kh & column bit kw
        uint32_t cm = 0, m = 0;
#pragma unroll
        for (int j = 0; j < kK; ++j) cm |= ((unsigned)(iw + g.pad - j) < (unsigned)g.wo ? 1u : 0u) << j;
#pragma unroll
        for (int j = 0; j < kK; ++j)
          if ((unsigned)(ih + g.pad - j) < (unsigned)g.ho) m |= cm << (j * kK);
        uint32_t addr = planes0 + (uint32_t)s * stage_bytes +
                        (uint32_t)(bl * hwo + (ih + g.pad) * g.wo + iw + g.pad) * 2u;      // tap (0, 0)
        const uint32_t step_kw = (uint32_t)(piece - 1) * 2u;                                // next plane, ow - 1
        const uint32_t step_kh = (uint32_t)(kK - g.wo) * 2u;   // after k kw-steps: back k columns, up one row (mod 2^32)
#pragma unroll
        for (int kh = 0; kh < kK; ++kh) {
#pragma unroll
          for (int kw = 0; kw < kK; ++kw)   // the last tap of a row also moves back k columns and up one row
            sum += lds_bf16_tap(addr, m, kw == kK - 1 ? step_kw + step_kh : step_kw);
        }
      } else {
        const __nv_bfloat16* pb = planes + (size_t)s * kk2 * piece + bl * hwo;
        for (int kh = 0; kh < k; ++kh) {
          const int oh = ih + g.pad - kh;
          if (oh < 0 || oh >= g.ho) continue;
          for (int kw = 0; kw < k; ++kw) {
            const int ow = iw + g.pad - kw;
            if (ow < 0 || ow >= g.wo) continue;
            sum += __bfloat162float(pb[(kh * k + kw) * piece + oh * g.wo + ow]);
          }
        }
      }
      dx[(((int64_t)mc * g.batch + b0 + bl) * g.cin + c0 + ci) * hw + p] = sum;
    }
    __syncthreads();
  }
}

static int conv_geom(const pl_conv2d_args* a, ConvGeom& g, const char* who) {
  PL_CHECK_ARG(a, "%s: NULL args", who);
  PL_CHECK_ARG(a->mc >= 0 && a->batch >= 0 && a->cin > 0 && a->cout > 0 && a->h > 0 && a->w > 0,
               "%s: bad dimension", who);
  PL_CHECK_ARG(a->k > 0 && a->stride > 0 && a->padding >= 0 && a->dilation > 0,
               "%s: bad kernel_size/stride/padding/dilation", who);
  g.mc = a->mc; g.batch = a->batch; g.cin = a->cin; g.cout = a->cout; g.h = a->h; g.w = a->w;
  g.k = a->k; g.stride = a->stride; g.pad = a->padding; g.dil = a->dilation;
  g.ho = (a->h + 2 * a->padding - a->dilation * (a->k - 1) - 1) / a->stride + 1;
  g.wo = (a->w + 2 * a->padding - a->dilation * (a->k - 1) - 1) / a->stride + 1;
  PL_CHECK_ARG(g.ho > 0 && g.wo > 0, "%s: empty output", who);
  g.mcx = a->x_mc_stride == 0 ? 1 : a->mc;
  g.K = a->cin * a->k * a->k;
  g.Kp = (g.K + 7) & ~7;
  g.Coutp = (g.cout + 7) & ~7;
  g.M = (int64_t)a->batch * g.ho * g.wo;
  PL_CHECK_ARG(g.M < (1ll << 31) && (int64_t)a->batch * a->h * a->w < (1ll << 31),
               "%s: per-sample pixel count overflows int32", who);
  PL_CHECK_ARG(a->x_mc_stride == 0 || a->x_mc_stride == (int64_t)a->batch * a->cin * a->h * a->w,
               "%s: PL_PREC_BF16 needs a dense [mc,B,C,H,W] input (or mc stride 0)", who);
  return PL_OK;
}

struct ConvScratch {
  float* sigma;            // packed {mu|sigma} bf16 groups of the [cout, K] weight matrix
  float* bias;             // [mc*cout] sampled bias (fwd) / column sums (bwd)
  __nv_bfloat16* A;        // [mcx*M*Kp]
  __nv_bfloat16* dyt;      // [mc*M*Coutp]        (bwd)
  __nv_bfloat16* dA;       // [mc*K*M] bf16       (bwd, when dx is wanted)
  size_t total;
};

static ConvScratch conv_carve(const ConvGeom& g, void* base, bool backward, bool want_dx, void* state) {
  uint8_t* b = (uint8_t*)(((uintptr_t)base + 255) & ~(uintptr_t)255);
  ConvScratch w;
  size_t off = 0;
  w.sigma = (float*)(b + off); off += align256(pack_bytes((size_t)g.cout, (size_t)g.K));
  w.bias = (float*)(b + off); off += align256((size_t)g.mc * g.cout * 4);
  if (state) {   // the im2col matrix lives in the caller's forward->backward state buffer
    w.A = (__nv_bfloat16*)(((uintptr_t)state + 255) & ~(uintptr_t)255);
  } else {
    w.A = (__nv_bfloat16*)(b + off); off += align256((size_t)g.mcx * g.M * g.Kp * 2);
  }
  w.dyt = (__nv_bfloat16*)(b + off);
  if (backward) off += align256((size_t)g.mc * g.M * g.Coutp * 2);
  w.dA = (__nv_bfloat16*)(b + off);
  if (backward && want_dx) off += align256((size_t)g.mc * g.K * g.M * 2);
  w.total = off + 256;
  return w;
}

size_t conv_tc_workspace_bytes(const pl_conv2d_args* a) {
  ConvGeom g;
  if (conv_geom(a, g, "pl_conv2d_workspace_bytes") != PL_OK) return 0;
  return conv_carve(g, nullptr, a->dy != nullptr, a->dx != nullptr, a->fwd_state).total;
}

size_t conv_tc_state_bytes(const pl_conv2d_args* a) {
  ConvGeom g;
  if (conv_geom(a, g, "pl_conv2d_state_bytes") != PL_OK) return 0;
  return align256((size_t)g.mcx * g.M * g.Kp * 2) + 512;
}

int conv_fwd_tc_launch(const pl_conv2d_args* a) {
  ConvGeom g;
  int rc = conv_geom(a, g, "pl_conv2d_fwd");
  if (rc != PL_OK) return rc;
  if (g.mc == 0 || g.batch == 0) return PL_OK;
  PL_CHECK_ARG(a->x && a->w_mu && a->w_rho && a->b_mu && a->b_rho && a->y, "pl_conv2d_fwd: NULL tensor");
  PL_CHECK_ARG(g.mc <= 65535, "pl_conv2d_fwd: mc too large");
  pl_conv2d_args fa = *a;
  fa.dy = nullptr;
  fa.dx = nullptr;
  if (!a->workspace || a->workspace_bytes < conv_tc_workspace_bytes(&fa)) {
    set_error("pl_conv2d_fwd: workspace %zu < %zu bytes", a->workspace_bytes, conv_tc_workspace_bytes(&fa));
    return PL_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)a->stream;
  if (a->fwd_state && a->fwd_state_bytes < conv_tc_state_bytes(a)) {
    set_error("pl_conv2d_fwd: fwd_state %zu < %zu bytes", a->fwd_state_bytes, conv_tc_state_bytes(a));
    return PL_ERR_WORKSPACE;
  }
  ConvScratch w = conv_carve(g, a->workspace, false, false, a->fwd_state);
  const int64_t nw = (int64_t)g.cout * g.K;
  // split operand (PL_PREC_BF16): the buffer holds mu16 [cout][Kp] then sig16 [cout][Kp] (half of pack_bytes each)
  const bool split = a->precision == PL_PREC_BF16;
  void* sig16 = (uint8_t*)w.sigma + split16_bytes((size_t)g.cout, (size_t)g.K);
  rc = split ? prep_params_split(a->w_mu, a->w_rho, w.sigma, sig16, g.cout, g.K, false, st)
             : prep_params(a->w_mu, a->w_rho, w.sigma, g.cout, g.K, false, st);
  if (rc != PL_OK) return rc;
  sample_bias_kernel<<<ew_grid((int64_t)g.mc * ((g.cout + 3) / 4), 256), 256, 0, st>>>(
      a->b_mu, a->b_rho, a->eps_b, make_sampler(a->seed, a->layer_id, true, a->mc_global_offset, a->seed_dev), g.mc,
      g.cout, w.bias);
  PL_LAUNCH_CHECK("sample_bias_kernel");
  rc = launch_im2col(a, g, w.A, st);
  if (rc != PL_OK) return rc;
  CUtensorMap tm;
  rc = make_tmap_bf16(&tm, w.A, (uint64_t)g.mcx * g.M, (uint64_t)g.Kp, 128, kBK);
  if (rc != PL_OK) return rc;
  TcGemmParams p{};
  p.mc = g.mc; p.rows = (int)g.M; p.ndim = g.cout; p.kdim = g.K;
  p.w_in = g.cout; p.w_out = g.K;
  p.a_mc_rows = g.mcx == 1 ? 0 : g.M;
  p.w_mu = a->w_mu; p.w_sigma = w.sigma; p.w_pack = (const uint4*)w.sigma; p.eps_w = a->eps_w;
  p.sw = make_sampler(a->seed, a->layer_id, false, a->mc_global_offset, a->seed_dev);
  p.bias = w.bias;
  p.out = a->y;
  p.out_hw = g.ho * g.wo;
  p.w_aligned = (g.K % 4) == 0;
  if (split) {
    p.w_sig16 = (const uint4*)sig16;
    CUtensorMap tmb;
    rc = make_tmap_mean(&tmb, w.sigma, (uint64_t)g.cout, (uint64_t)g.K, (uint64_t)g.Kp, true, false);
    if (rc != PL_OK) return rc;
    return a->eps_w ? launch_gemm<true, true, false, true>(tm, p, st, &tmb)
                    : launch_gemm<true, false, false, true>(tm, p, st, &tmb);
  }
  return a->eps_w ? launch_gemm<true, true>(tm, p, st) : launch_gemm<true, false>(tm, p, st);
}

int conv_bwd_tc_launch(const pl_conv2d_args* a) {
  ConvGeom g;
  int rc = conv_geom(a, g, "pl_conv2d_bwd");
  if (rc != PL_OK) return rc;
  PL_CHECK_ARG(a->x && a->w_mu && a->w_rho && a->b_rho && a->dy, "pl_conv2d_bwd: NULL tensor");
  PL_CHECK_ARG((a->dw_mu == nullptr) == (a->dw_rho == nullptr), "pl_conv2d_bwd: dw_mu/dw_rho must both be set or NULL");
  if (!a->workspace || a->workspace_bytes < conv_tc_workspace_bytes(a)) {
    set_error("pl_conv2d_bwd: workspace %zu < %zu bytes", a->workspace_bytes, conv_tc_workspace_bytes(a));
    return PL_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)a->stream;
  if (a->fwd_state && a->fwd_state_bytes < conv_tc_state_bytes(a)) {
    set_error("pl_conv2d_bwd: fwd_state %zu < %zu bytes", a->fwd_state_bytes, conv_tc_state_bytes(a));
    return PL_ERR_WORKSPACE;
  }
  ConvScratch w = conv_carve(g, a->workspace, true, a->dx != nullptr, a->fwd_state);
  const int hwo = g.ho * g.wo;
  const int64_t nw = (int64_t)g.cout * g.K;
  if (a->dx || a->dw_mu) {
    dim3 grid((unsigned)(g.mc * g.batch), (unsigned)((g.Coutp + 31) / 32), (unsigned)((hwo + 31) / 32));
    PL_CHECK_ARG((int64_t)g.mc * g.batch < (1ll << 31) && (hwo + 31) / 32 <= 65535 && (g.Coutp + 31) / 32 <= 65535,
                 "pl_conv2d_bwd: transpose grid too large");
    dy_pixel_major_kernel<<<grid, dim3(32, 8), 0, st>>>(a->dy, w.dyt, g.cout, g.Coutp, hwo);
    PL_LAUNCH_CHECK("dy_pixel_major_kernel");
  }
  if (a->dw_mu) {
    if (!a->fwd_state) {   // stateless call: rebuild the im2col matrix
      rc = launch_im2col(a, g, w.A, st);
      if (rc != PL_OK) return rc;
    }
    CUtensorMap tmx, tmdy;
    rc = make_tmap_bf16_3d(&tmx, w.dyt, (uint64_t)g.mc, (uint64_t)g.M, (uint64_t)g.Coutp, kBK, 64);
    if (rc != PL_OK) return rc;
    rc = make_tmap_bf16_3d(&tmdy, w.A, (uint64_t)g.mcx, (uint64_t)g.M, (uint64_t)g.Kp, kBK, 64);
    if (rc != PL_OK) return rc;
    TcDwParams p{};
    p.mc = g.mc; p.batch = (int)g.M; p.w_in = g.cout; p.w_out = g.K;
    p.x_broadcast = 0;
    p.dy_broadcast = g.mcx == 1;
    p.w_aligned = (g.K % 4) == 0;
    p.w_rho = a->w_rho; p.eps_w = a->eps_w;
    p.sw = make_sampler(a->seed, a->layer_id, false, a->mc_global_offset, a->seed_dev);
    p.dw_mu = a->dw_mu; p.dw_rho = a->dw_rho;
    // few weight tiles, long pixel reduction: one CTA per (tile, MC sample, pixel chunk)
    const int tiles = ((g.cout + 127) / 128) * ((g.K + 127) / 128);
    const int num_kb = (int)((g.M + kBK - 1) / kBK);
    p.mc_per_split = 1;
    int want = (4 * 148) / (tiles * g.mc);
    if (want < 1) want = 1;
    if (want > num_kb) want = num_kb;
    p.kb_per_split = (num_kb + want - 1) / want;
    p.kb_splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
    p.atomic_out = (g.mc * p.kb_splits) > 1;
    if (p.atomic_out) {
      PL_CUDA(cudaMemsetAsync(a->dw_mu, 0, sizeof(float) * nw, st));
      PL_CUDA(cudaMemsetAsync(a->dw_rho, 0, sizeof(float) * nw, st));
    }
    PL_CHECK_ARG((int64_t)g.mc * p.kb_splits <= 65535, "pl_conv2d_bwd: too many reduction splits");
    dim3 grid((unsigned)((g.K + 127) / 128), (unsigned)((g.cout + 127) / 128), (unsigned)(g.mc * p.kb_splits));
    if (a->eps_w) {
      PL_CUDA(cudaFuncSetAttribute(dw_tc<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemBytes));
      dw_tc<true, false><<<grid, kTcThreads, kDwSmemBytes, st>>>(tmx, tmdy, p);
    } else {
      PL_CUDA(cudaFuncSetAttribute(dw_tc<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmemBytes));
      dw_tc<false, false><<<grid, kTcThreads, kDwSmemBytes, st>>>(tmx, tmdy, p);
    }
    PL_LAUNCH_CHECK("dw_tc");
  }
  if (a->dx) {
    const bool split = a->precision == PL_PREC_BF16;
    void* sig16 = (uint8_t*)w.sigma + split16_bytes((size_t)g.cout, (size_t)g.K);
    rc = split ? prep_params_split(a->w_mu, a->w_rho, w.sigma, sig16, g.cout, g.K, false, st)
               : prep_params(a->w_mu, a->w_rho, w.sigma, g.cout, g.K, false, st);
    if (rc != PL_OK) return rc;
    CUtensorMap tm;
    rc = make_tmap_bf16(&tm, w.dyt, (uint64_t)g.mc * g.M, (uint64_t)g.Coutp, 128, kBK);
    if (rc != PL_OK) return rc;
    TcGemmParams p{};
    p.mc = g.mc; p.rows = (int)g.M; p.ndim = g.K; p.kdim = g.cout;
    p.w_in = g.cout; p.w_out = g.K;
    p.a_mc_rows = g.M;
    p.w_mu = a->w_mu; p.w_sigma = w.sigma; p.w_pack = (const uint4*)w.sigma; p.eps_w = a->eps_w;
    p.sw = make_sampler(a->seed, a->layer_id, false, a->mc_global_offset, a->seed_dev);
    p.bias = nullptr;
    p.out = nullptr;
    p.out_lp = w.dA;              // bf16 [mc][K][M]: K-major so that col2im reads contiguous pixels
    p.out_hw = (int)g.M;
    p.w_aligned = (g.K % 4) == 0;
    if (split) {
      p.w_sig16 = (const uint4*)sig16;
      CUtensorMap tmb;
      rc = make_tmap_mean(&tmb, w.sigma, (uint64_t)g.cout, (uint64_t)g.K, (uint64_t)g.Kp, false, false);
      if (rc != PL_OK) return rc;
      rc = a->eps_w ? launch_gemm<false, true, false, true>(tm, p, st, &tmb)
                    : launch_gemm<false, false, false, true>(tm, p, st, &tmb);
    } else {
      rc = a->eps_w ? launch_gemm<false, true>(tm, p, st) : launch_gemm<false, false>(tm, p, st);
    }
    if (rc != PL_OK) return rc;
    const int64_t n = (int64_t)g.mc * g.batch * g.cin * g.h * g.w;
    const int hwo = g.ho * g.wo, hw = g.h * g.w;
    // images per CTA: pieces of >= 256 pixels (512 bytes) where the batch allows it; 8 channels per CTA
    int bpc = hwo >= 256 ? 1 : (256 + hwo - 1) / hwo;
    if (bpc > g.batch) bpc = g.batch;
    const int cpc = g.cin < 8 ? g.cin : 8;
    const size_t stage = (size_t)g.k * g.k * bpc * hwo * 2;
    const size_t tile_smem = 2 * stage + 16;
    const char* generic = getenv("PL_COL2IM_GENERIC");   // =1 (read per call: a test flips it): element-wise kernel
    const int force = generic ? atoi(generic) : 0;
    const unsigned gc = (unsigned)((g.cin + cpc - 1) / cpc), gb = (unsigned)((g.batch + bpc - 1) / bpc);
    if (g.stride == 1 && g.dil == 1 && (hwo & 7) == 0 && tile_smem <= 200 * 1024 && stage < (1u << 20) &&
        gb <= 65535 && gc <= 65535 && g.mc <= 65535 && force != 1) {
#define PL_COL2IM(KK)                                                                                       \
  do {                                                                                                     \
    if (tile_smem > 48 * 1024)   /* per call: the attribute is per device */                               \
      PL_CUDA(cudaFuncSetAttribute(col2im_pipe_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                   (int)tile_smem));                                                       \
    col2im_pipe_kernel<KK><<<dim3(gc, gb, (unsigned)g.mc), 256, tile_smem, st>>>(w.dA, a->dx, g, cpc, bpc); \
  } while (0)
      if (g.k == 5) PL_COL2IM(5);
      else if (g.k == 3) PL_COL2IM(3);
      else PL_COL2IM(0);
#undef PL_COL2IM
    } else if (g.stride == 1 && g.dil == 1) {
      col2im_kernel<true><<<ew_grid(n, 256), 256, 0, st>>>(w.dA, a->dx, g);
    } else {
      col2im_kernel<false><<<ew_grid(n, 256), 256, 0, st>>>(w.dA, a->dx, g);
    }
    PL_LAUNCH_CHECK("col2im_kernel");
  }
  return PL_OK;
}

}  // namespace pl
