// libmcn runtime: thread-local error text, launch accounting, version, the per-device workspace
// and the two reduction helpers every deterministic kernel shares (xsum.cuh).
#include <cstring>

#include "mcn_common.cuh"
#include "xsum.cuh"

namespace mcn {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCN_PDL");
    v = (e && e[0] == '1') ? 1 : 0;   // off by default: no gain measured inside a CUDA graph (23.3 vs 23.1 ms)
  }
  return v != 0;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static Workspace g_ws[64];

Workspace current_workspace() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return Workspace{nullptr, 0};
  return g_ws[dev];
}

XsScratch xs_scratch(int n, const char* who) {
  const Workspace w = current_workspace();
  if (w.base == nullptr) {
    set_error("%s: no workspace registered for this device (mcn_set_workspace)", who);
    return XsScratch{nullptr, nullptr};
  }
  if (n > kWsXsMax) {
    set_error("%s: %d sums exceed the workspace's limb area (%d)", who, n, kWsXsMax);
    return XsScratch{nullptr, nullptr};
  }
  return XsScratch{reinterpret_cast<long long*>(w.base + kWsXsOff),
                   reinterpret_cast<unsigned int*>(w.base + kWsCounterOff)};
}

namespace {

// 256 threads = 32 float4 columns x Y split lanes.  Lane y sums splits y, y+Y, y+2Y, ... in that
// order (several loads in flight), the Y partial sums are then combined in lane order: a fixed
// summation tree, so the result is bit-reproducible, and the depth of the serial load chain is
// splits/Y instead of splits (the 148-way splits of the small early layers were latency-bound).
template <int Y>
__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ ws, long long stride, int splits, long long n4,
                     float* __restrict__ dw) {
  MCN_PDL_PROLOGUE();
  constexpr int X = 256 / Y;
  __shared__ float4 part[Y][X];
  const int tx = threadIdx.x % X, ty = threadIdx.x / X;
  for (long long v0 = (long long)blockIdx.x * X; v0 < n4; v0 += (long long)gridDim.x * X) {
    const long long v = v0 + tx;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (v < n4) {
      int s = ty;
      for (; s + 3 * Y < splits; s += 4 * Y) {
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          p[u] = __ldcs(reinterpret_cast<const float4*>(ws + (s + u * Y) * stride) + v);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          acc.x += p[u].x;
          acc.y += p[u].y;
          acc.z += p[u].z;
          acc.w += p[u].w;
        }
      }
      for (; s < splits; s += Y) {
        const float4 p = __ldcs(reinterpret_cast<const float4*>(ws + s * stride) + v);
        acc.x += p.x;
        acc.y += p.y;
        acc.z += p.z;
        acc.w += p.w;
      }
    }
    if (Y > 1) {
      part[ty][tx] = acc;
      __syncthreads();
    }
    if (ty == 0 && v < n4) {
      float4 o = *(reinterpret_cast<const float4*>(dw) + v);
      if (Y > 1) {
#pragma unroll
        for (int y = 1; y < Y; ++y) {
          acc.x += part[y][tx].x;
          acc.y += part[y][tx].y;
          acc.z += part[y][tx].z;
          acc.w += part[y][tx].w;
        }
      }
      o.x += acc.x;
      o.y += acc.y;
      o.z += acc.z;
      o.w += acc.w;
      *(reinterpret_cast<float4*>(dw) + v) = o;
    }
    if (Y > 1) __syncthreads();
  }
}

__global__ void splitk_reduce_scalar_kernel(const float* __restrict__ ws, long long stride, int splits,
                                            long long i0, long long n, float* __restrict__ dw) {
  MCN_PDL_PROLOGUE();
  for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += ws[s * stride + i];
    dw[i] += acc;
  }
}

__global__ void xsum_decode_kernel(const long long* __restrict__ limbs, int n, float* out_f32,
                                   double* out_f64, int accumulate) {
  MCN_PDL_PROLOGUE();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double v = xs::read(limbs, n, i);
  if (out_f32) out_f32[i] = static_cast<float>(accumulate ? static_cast<double>(out_f32[i]) + v : v);
  if (out_f64) out_f64[i] = accumulate ? out_f64[i] + v : v;
}

}  // namespace

int launch_splitk_reduce(const float* ws, long long stride, int splits, long long n, float* dw,
                         cudaStream_t st) {
  // vector path needs 16-byte aligned slices
  const bool vec = (stride % 4 == 0) && (reinterpret_cast<uintptr_t>(ws) % 16 == 0) &&
                   (reinterpret_cast<uintptr_t>(dw) % 16 == 0);
  const long long n4 = vec ? n / 4 : 0;
  if (n4 > 0) {
    const int y = splits <= 4 ? 1 : (splits <= 8 ? 2 : (splits <= 32 ? 4 : 8));
    const int x = 256 / y;
    const int grid = (int)std::max<long long>(1, std::min<long long>((n4 + x - 1) / x, 16LL * num_sms()));
    if (y == 1) ::mcn::launch(splitk_reduce_kernel<1>, grid, 256, 0, st, ws, stride, splits, n4, dw);
    else if (y == 2) ::mcn::launch(splitk_reduce_kernel<2>, grid, 256, 0, st, ws, stride, splits, n4, dw);
    else if (y == 4) ::mcn::launch(splitk_reduce_kernel<4>, grid, 256, 0, st, ws, stride, splits, n4, dw);
    else ::mcn::launch(splitk_reduce_kernel<8>, grid, 256, 0, st, ws, stride, splits, n4, dw);
    const int rc = after_launch("splitk_reduce");
    if (rc) return rc;
  }
  if (n4 * 4 < n) {
    const long long rest = n - n4 * 4;
    const int grid = (int)std::max<long long>(1, std::min<long long>((rest + 255) / 256, 8LL * num_sms()));
    ::mcn::launch(splitk_reduce_scalar_kernel, grid, 256, 0, st, ws, stride, splits, n4 * 4, n, dw);
    return after_launch("splitk_reduce");
  }
  return MCN_OK;
}

}  // namespace mcn

extern "C" const char* mcn_last_error(void) { return mcn::g_err; }
extern "C" int mcn_version(void) { return 200; }
extern "C" long long mcn_launch_count(void) {
  return mcn::g_launches.load(std::memory_order_relaxed);
}

extern "C" int mcn_set_workspace(void* ptr, long long bytes) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    mcn::set_error("set_workspace: no current CUDA device");
    return MCN_ECUDA;
  }
  if (ptr != nullptr && (bytes < mcn::kWsMinBytes || reinterpret_cast<uintptr_t>(ptr) % 256 != 0)) {
    mcn::set_error("set_workspace: need a 256-byte aligned buffer of at least %lld bytes",
                   mcn::kWsMinBytes);
    return MCN_EINVAL;
  }
  mcn::g_ws[dev] = mcn::Workspace{static_cast<unsigned char*>(ptr), ptr ? bytes : 0};
  return MCN_OK;
}
extern "C" long long mcn_workspace_min_bytes(void) { return mcn::kWsMinBytes; }

extern "C" int mcn_xsum_decode(const long long* limbs, int n, float* out_f32, double* out_f64,
                               int accumulate, void* stream) {
  MCN_REQUIRE(limbs && n > 0 && (out_f32 || out_f64), "xsum_decode: bad argument");
  ::mcn::launch(mcn::xsum_decode_kernel, (n + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), 
      limbs, n, out_f32, out_f64, accumulate);
  return mcn::after_launch("xsum_decode");
}
