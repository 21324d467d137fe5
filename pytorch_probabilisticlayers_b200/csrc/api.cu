// Library-level entry points: version, error text, device probe, eps dump (test bridge).
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace pl {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return PL_ERR_CUDA;
}

// out[mc][row][col] (normals) or out[mc][row][4*ceil(cols/8)] (raw Philox words, one call per 8
// columns); one thread per 4 columns
__global__ void eps_dump_kernel(Sampler s, int num_mc, int rows, int cols, void* out, int raw) {
  s = live(s);
  const int groups = 2 * ((cols + 7) / 8);
  const int64_t total = (int64_t)num_mc * rows * groups;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(t % groups);
    const int r = (int)((t / groups) % rows);
    const int m = (int)(t / ((int64_t)groups * rows));
    if (raw) {
      if (g & 1) continue;   // one Philox call per 8 columns
      const uint4 w = philox4x32_10(g >> 1, r, s.mc_offset + m, s.stream, s.k0, s.k1);
      uint32_t* o = (uint32_t*)out + ((int64_t)m * rows + r) * (2 * groups) + 2 * g;
      o[0] = w.x; o[1] = w.y; o[2] = w.z; o[3] = w.w;
    } else {
      const float4 n = eps4(s, m, r, g);
      float* o = (float*)out + ((int64_t)m * rows + r) * cols + 4 * g;
      const float v[4] = {n.x, n.y, n.z, n.w};
      for (int j = 0; j < 4 && 4 * g + j < cols; ++j) o[j] = v[j];
    }
  }
}

}  // namespace pl

extern "C" int pl_version(void) { return PL_B200_VERSION; }

extern "C" const char* pl_last_error(void) { return pl::g_err; }

extern "C" uint64_t pl_launch_count(void) { return pl::g_launches.load(std::memory_order_relaxed); }

extern "C" size_t pl_struct_size(int32_t which) {
  switch (which) {
    case 0: return sizeof(pl_linear_args);
    case 1: return sizeof(pl_conv2d_args);
    case 2: return sizeof(pl_ce_args);
    case 3: return sizeof(pl_bnpool_args);
    case 4: return sizeof(pl_shard_seg);
    case 5: return sizeof(pl_shard_ctx);
    default: return 0;
  }
}

extern "C" int pl_device_supported(void) {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
  return major == 10 ? 1 : 0;
}

extern "C" int pl_eps_dump(uint64_t seed, uint32_t stream_id, uint32_t mc_global_offset,
                           int32_t num_mc, int32_t rows, int32_t cols, void* out, int32_t raw,
                           void* stream) {
  return pl_eps_dump_dev(seed, nullptr, stream_id, mc_global_offset, num_mc, rows, cols, out, raw, stream);
}

extern "C" int pl_eps_dump_dev(uint64_t seed, const uint64_t* seed_dev, uint32_t stream_id, uint32_t mc_global_offset,
                               int32_t num_mc, int32_t rows, int32_t cols, void* out, int32_t raw,
                               void* stream) {
  using namespace pl;
  PL_CHECK_ARG(out, "pl_eps_dump: NULL out");
  PL_CHECK_ARG(num_mc >= 0 && rows >= 0 && cols >= 0, "pl_eps_dump: negative size");
  const int64_t total = (int64_t)num_mc * rows * (2 * ((cols + 7) / 8));
  if (total == 0) return PL_OK;
  Sampler s;
  s.k0 = (uint32_t)(seed & 0xffffffffu);
  s.k1 = (uint32_t)(seed >> 32);
  s.stream = stream_id;
  s.mc_offset = mc_global_offset;
  s.seed_dev = seed_dev;
  const int threads = 256;
  int64_t blocks = ceil_div64(total, threads);
  if (blocks > 148 * 16) blocks = 148 * 16;
  eps_dump_kernel<<<(int)blocks, threads, 0, (cudaStream_t)stream>>>(s, num_mc, rows, cols, out, raw);
  PL_LAUNCH_CHECK("eps_dump_kernel");
  return PL_OK;
}
