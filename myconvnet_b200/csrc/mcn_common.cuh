// Shared host/device helpers for libmcn: error reporting, launch accounting, dtype dispatch,
// vector load/store and warp/block reductions.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/mcn.h"

namespace mcn {

void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;

// Call right after a kernel launch: records launch errors, counts the launch.
inline int after_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return MCN_ECUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MCN_OK;
}

// Programmatic dependent launch: every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, so the NEXT kernel's blocks may be scheduled
// (and run their private prologue) while this one drains.  MCN_PDL_PROLOGUE() is the first thing a
// kernel does before it touches global memory: it waits until the previous grid has completed and
// its writes are visible.  Correctness never depends
// on it (without the attribute the instruction is a no-op).  MCN_PDL=1 switches the attribute ON: it
// is off by default because replayed CUDA graphs showed no gain from it (profiles/README.md).
// (An explicit early griddepcontrol.launch_dependents was measured SLOWER, 25.1 vs 22.8 ms/step:
// dependent blocks then sit resident next to the primary for its whole run.  The implicit trigger at
// block exit keeps co-residency to the primary's tail.)
#define MCN_PDL_PROLOGUE() asm volatile("griddepcontrol.wait;" ::: "memory")

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline void launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                   Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

#define MCN_REQUIRE(cond, ...)         \
  do {                                 \
    if (!(cond)) {                     \
      ::mcn::set_error(__VA_ARGS__);   \
      return MCN_EINVAL;               \
    }                                  \
  } while (0)

// Dispatch a templated launch on the activation dtype.
#define MCN_DISPATCH_DTYPE(dtype, T, ...)                  \
  do {                                                     \
    if ((dtype) == MCN_F32) {                              \
      using T = float;                                     \
      __VA_ARGS__;                                         \
    } else if ((dtype) == MCN_BF16) {                      \
      using T = __nv_bfloat16;                             \
      __VA_ARGS__;                                         \
    } else {                                               \
      ::mcn::set_error("unsupported dtype %d", (dtype));   \
      return MCN_EINVAL;                                   \
    }                                                      \
  } while (0)

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// 16-byte vectors: 4 floats or 8 bf16.
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  float4 raw;
  __device__ __forceinline__ float get(int i) const { return (&raw.x)[i]; }
  __device__ __forceinline__ void set(int i, float v) { (&raw.x)[i] = v; }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  uint4 raw;
  __device__ __forceinline__ float get(int i) const {
    uint32_t w = (&raw.x)[i >> 1];
    uint32_t bits = (i & 1) ? (w & 0xFFFF0000u) : (w << 16);
    return __uint_as_float(bits);
  }
  __device__ __forceinline__ void set(int i, float v) {
    uint32_t b = static_cast<uint32_t>(__bfloat16_as_ushort(__float2bfloat16_rn(v)));
    uint32_t& w = (&raw.x)[i >> 1];
    w = (i & 1) ? ((w & 0x0000FFFFu) | (b << 16)) : ((w & 0xFFFF0000u) | b);
  }
};
template <typename T>
__device__ __forceinline__ Vec16<T> ld_vec(const T* p) {
  Vec16<T> v;
  v.raw = *reinterpret_cast<const decltype(v.raw)*>(p);
  return v;
}
// streaming (read-once) variant: bypass L1 allocation
template <typename T>
__device__ __forceinline__ Vec16<T> ld_vec_stream(const T* p) {
  Vec16<T> v;
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  v.raw = *reinterpret_cast<decltype(v.raw)*>(&r);
  return v;
}
template <typename T>
__device__ __forceinline__ void st_vec(T* p, const Vec16<T>& v) {
  *reinterpret_cast<decltype(v.raw)*>(p) = v.raw;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Activation forward / derivative.  The relu family's derivative is expressed through the
// OUTPUT y (what the fused kernels keep); the smooth ones through the INPUT x.
__device__ __forceinline__ float act_fwd(int act, float x, float alpha) {
  switch (act) {
    case MCN_ACT_RELU: return fmaxf(x, 0.f);
    case MCN_ACT_RELU6: return fminf(fmaxf(x, 0.f), 6.f);
    case MCN_ACT_LRELU: return fmaxf(x, alpha * x);
    case MCN_ACT_TANH: return tanhf(x);
    case MCN_ACT_SIGMOID: return 1.f / (1.f + __expf(-x));
    case MCN_ACT_SWISH: return x / (1.f + __expf(-x));
    default: return x;
  }
}
__device__ __forceinline__ float act_grad_from_x(int act, float x, float alpha) {
  switch (act) {
    case MCN_ACT_RELU: return x > 0.f ? 1.f : 0.f;
    case MCN_ACT_RELU6: return (x > 0.f && x < 6.f) ? 1.f : 0.f;
    case MCN_ACT_LRELU: return x > 0.f ? 1.f : alpha;
    case MCN_ACT_TANH: { float t = tanhf(x); return 1.f - t * t; }
    case MCN_ACT_SIGMOID: { float s = 1.f / (1.f + __expf(-x)); return s * (1.f - s); }
    case MCN_ACT_SWISH: { float s = 1.f / (1.f + __expf(-x)); return s * (1.f + x * (1.f - s)); }
    default: return 1.f;
  }
}
__device__ __forceinline__ float act_grad_from_y(int act, float y, float alpha) {
  switch (act) {
    case MCN_ACT_RELU: return y > 0.f ? 1.f : 0.f;
    case MCN_ACT_RELU6: return (y > 0.f && y < 6.f) ? 1.f : 0.f;
    case MCN_ACT_LRELU: return y > 0.f ? 1.f : alpha;
    case MCN_ACT_TANH: return 1.f - y * y;
    case MCN_ACT_SIGMOID: return y * (1.f - y);
    default: return 1.f;
  }
}
// true when the derivative can be rebuilt from the output alone
__host__ __device__ inline bool act_grad_uses_output(int act) { return act != MCN_ACT_SWISH; }

}  // namespace mcn
