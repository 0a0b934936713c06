// Random train-time ops: tf.nn.dropout (reference resnet_v1_5.py:75, efficientnet.py:117,
// deeplabv3plus.py:54) and ConvNet.stochastic_depth (convnet.py:2500-2512), plus the nearest-
// neighbour resize of upsampling_2d_layer (convnet.py:2393-2395).
//
// Randomness is counter-based (Philox4x32-10, restated here; no cuRAND state): a keep decision is a
// pure function of (seed, step, layer id, element index), so
//   - the backward pass recomputes the mask instead of storing it,
//   - a captured CUDA graph draws new masks every replay: seed and step live in the
//     hyper-parameter vector on the device (ints at float slots 12 and 13),
//   - results do not depend on the launch configuration (bit-reproducible), and the parity tests
//     regenerate the same masks on the host (oracle/philox.py).
// u = (bits >> 8) * 2^-24 in [0,1); keep <=> u >= rate; kept values are scaled by 1/(1-rate)
// (SURVEY Appendix A.10).
#include "mcn_common.cuh"

namespace mcn {
namespace {

constexpr int kHpSeedSlot = 12, kHpStepSlot = 13;

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0;
  c[1] = lo1;
  c[2] = n2;
  c[3] = lo0;
}
// counter = (idx_lo, idx_hi, step, 0), key = (seed, layer)
__device__ __forceinline__ void philox4x32_10(unsigned long long idx, uint32_t step, uint32_t seed,
                                              uint32_t layer, uint32_t (&out)[4]) {
  uint32_t c[4] = {static_cast<uint32_t>(idx), static_cast<uint32_t>(idx >> 32), step, 0u};
  uint32_t k0 = seed, k1 = layer;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}
__device__ __forceinline__ float u01(uint32_t bits) { return static_cast<float>(bits >> 8) * 0x1p-24f; }

// y[i] = keep(i) ? x[i]/(1-rate) : 0 ; the same kernel is the backward pass (dx from dy).
template <typename T>
__global__ void dropout_kernel(const T* __restrict__ x, long long n, float rate, const float* __restrict__ hp,
                               uint32_t layer, T* __restrict__ y) {
  MCN_PDL_PROLOGUE();
  const uint32_t seed = reinterpret_cast<const uint32_t*>(hp)[kHpSeedSlot];
  const uint32_t step = reinterpret_cast<const uint32_t*>(hp)[kHpStepSlot];
  const float scale = 1.f / (1.f - rate);
  const long long quads = (n + 3) / 4;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < quads;
       q += (long long)gridDim.x * blockDim.x) {
    uint32_t r[4];
    philox4x32_10(static_cast<unsigned long long>(q), step, seed, layer, r);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long i = q * 4 + j;
      if (i < n) y[i] = from_f32<T>(u01(r[j]) >= rate ? to_f32(x[i]) * scale : 0.f);
    }
  }
}

__device__ __forceinline__ float survive_factor(int sample, float rate, uint32_t seed, uint32_t step,
                                                uint32_t layer) {
  uint32_t r[4];
  philox4x32_10(static_cast<unsigned long long>(sample), step, seed, layer, r);
  return u01(r[0]) >= rate ? 1.f / (1.f - rate) : 0.f;
}

// y = act(a * survived[n] + b): per-sample Bernoulli keep of the residual branch
template <typename T>
__global__ void sd_add_fwd_kernel(const T* __restrict__ a, const T* __restrict__ b, int N, long long per,
                                  float rate, const float* __restrict__ hp, uint32_t layer, int act,
                                  float alpha, T* __restrict__ y) {
  MCN_PDL_PROLOGUE();
  const uint32_t seed = reinterpret_cast<const uint32_t*>(hp)[kHpSeedSlot];
  const uint32_t step = reinterpret_cast<const uint32_t*>(hp)[kHpStepSlot];
  const long long total = (long long)N * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float s = survive_factor(static_cast<int>(i / per), rate, seed, step, layer);
    y[i] = from_f32<T>(act_fwd(act, to_f32(a[i]) * s + to_f32(b[i]), alpha));
  }
}
// dz = dy * act'(y); da = dz * survived[n]; db = dz   (da / db may be NULL)
template <typename T>
__global__ void sd_add_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ y, int N, long long per,
                                  float rate, const float* __restrict__ hp, uint32_t layer, int act,
                                  float alpha, T* __restrict__ da, T* __restrict__ db) {
  MCN_PDL_PROLOGUE();
  const uint32_t seed = reinterpret_cast<const uint32_t*>(hp)[kHpSeedSlot];
  const uint32_t step = reinterpret_cast<const uint32_t*>(hp)[kHpStepSlot];
  const long long total = (long long)N * per;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float s = survive_factor(static_cast<int>(i / per), rate, seed, step, layer);
    const float dz = to_f32(dy[i]) * (act == MCN_ACT_NONE ? 1.f : act_grad_from_y(act, to_f32(y[i]), alpha));
    if (da) da[i] = from_f32<T>(dz * s);
    if (db) db[i] = from_f32<T>(dz);
  }
}

// ---- nearest-neighbour resize (tf.image.resize_nearest_neighbor)
// mode 0: src = floor(dst*in/out); 1 (align_corners): src = round(dst*(in-1)/(out-1));
// 2 (half_pixel_centers): src = floor((dst+0.5)*in/out); all clamped to in-1.
__device__ __forceinline__ int nearest_src(int dst, int in, int out, int mode) {
  int s;
  if (mode == 1) {
    const float sc = out > 1 ? static_cast<float>(in - 1) / static_cast<float>(out - 1) : 0.f;
    s = static_cast<int>(roundf(dst * sc));
  } else if (mode == 2) {
    s = static_cast<int>(floorf((dst + 0.5f) * (static_cast<float>(in) / static_cast<float>(out))));
  } else {
    s = static_cast<int>(floorf(dst * (static_cast<float>(in) / static_cast<float>(out))));
  }
  return min(max(s, 0), in - 1);
}
template <typename T>
__global__ void resize_nearest_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C, int Ho,
                                          int Wo, int mode, T* __restrict__ y) {
  MCN_PDL_PROLOGUE();
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long r = i / C;
    const int q = static_cast<int>(r % Wo);
    r /= Wo;
    const int p = static_cast<int>(r % Ho);
    const int n = static_cast<int>(r / Ho);
    y[i] = x[(((long long)n * H + nearest_src(p, H, Ho, mode)) * W + nearest_src(q, W, Wo, mode)) * C + c];
  }
}
// gather form of the scatter-add gradient (deterministic): every input pixel sums the output
// pixels that read it.  The source index is monotone in the destination index, so the candidates
// form a contiguous range around dst ~ src*out/in.
template <typename T>
__global__ void resize_nearest_bwd_kernel(const T* __restrict__ dy, int N, int H, int W, int C, int Ho,
                                          int Wo, int mode, T* __restrict__ dx) {
  MCN_PDL_PROLOGUE();
  const long long total = (long long)N * H * W * C;
  const int rh = Ho / max(H, 1) + 2, rw = Wo / max(W, 1) + 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long r = i / C;
    const int w = static_cast<int>(r % W);
    r /= W;
    const int h = static_cast<int>(r % H);
    const int n = static_cast<int>(r / H);
    const int pc = static_cast<int>((long long)h * Ho / H), qc = static_cast<int>((long long)w * Wo / W);
    float acc = 0.f;
    for (int p = max(pc - rh, 0); p <= min(pc + rh, Ho - 1); ++p) {
      if (nearest_src(p, H, Ho, mode) != h) continue;
      for (int q = max(qc - rw, 0); q <= min(qc + rw, Wo - 1); ++q)
        if (nearest_src(q, W, Wo, mode) == w) acc += to_f32(dy[(((long long)n * Ho + p) * Wo + q) * C + c]);
    }
    dx[i] = from_f32<T>(acc);
  }
}

inline int grid_for(long long n, int block) {
  return (int)std::max<long long>(1, std::min<long long>((n + block - 1) / block, 16LL * num_sms()));
}

}  // namespace
}  // namespace mcn

using namespace mcn;

extern "C" int mcn_dropout(int dtype, const void* x, long long n, float rate, const float* hp, int layer,
                           void* y, void* stream) {
  MCN_REQUIRE(x && y && hp && n >= 0 && rate >= 0.f && rate < 1.f, "dropout: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    ::mcn::launch(dropout_kernel<T>, grid_for((n + 3) / 4, 256), 256, 0, st, static_cast<const T*>(x), n, rate, hp,
                                                              static_cast<uint32_t>(layer), static_cast<T*>(y));
  });
  return after_launch("dropout");
}

extern "C" int mcn_sd_add_fwd(int dtype, const void* a, const void* b, int N, long long per_sample,
                              float rate, const float* hp, int layer, int act, float alpha, void* y,
                              void* stream) {
  MCN_REQUIRE(a && b && y && hp && N > 0 && per_sample > 0 && rate >= 0.f && rate < 1.f,
              "sd_add_fwd: bad argument");
  MCN_REQUIRE(act == MCN_ACT_NONE || act == MCN_ACT_RELU || act == MCN_ACT_RELU6 || act == MCN_ACT_LRELU,
              "sd_add_fwd: only relu-family activations can be fused (derivative from the output)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    ::mcn::launch(sd_add_fwd_kernel<T>, grid_for((long long)N * per_sample, 256), 256, 0, st, 
        static_cast<const T*>(a), static_cast<const T*>(b), N, per_sample, rate, hp,
        static_cast<uint32_t>(layer), act, alpha, static_cast<T*>(y));
  });
  return after_launch("sd_add_fwd");
}
extern "C" int mcn_sd_add_bwd(int dtype, const void* dy, const void* y, int N, long long per_sample,
                              float rate, const float* hp, int layer, int act, float alpha, void* da,
                              void* db, void* stream) {
  MCN_REQUIRE(dy && hp && N > 0 && per_sample > 0 && (act == MCN_ACT_NONE || y), "sd_add_bwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    ::mcn::launch(sd_add_bwd_kernel<T>, grid_for((long long)N * per_sample, 256), 256, 0, st, 
        static_cast<const T*>(dy), static_cast<const T*>(y), N, per_sample, rate, hp,
        static_cast<uint32_t>(layer), act, alpha, static_cast<T*>(da), static_cast<T*>(db));
  });
  return after_launch("sd_add_bwd");
}

extern "C" int mcn_resize_nearest_fwd(int dtype, const void* x, int N, int H, int W, int C, int Ho,
                                      int Wo, int mode, void* y, void* stream) {
  MCN_REQUIRE(x && y && mode >= 0 && mode <= 2, "resize_nearest_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    ::mcn::launch(resize_nearest_fwd_kernel<T>, grid_for((long long)N * Ho * Wo * C, 256), 256, 0, st, 
        static_cast<const T*>(x), N, H, W, C, Ho, Wo, mode, static_cast<T*>(y));
  });
  return after_launch("resize_nearest_fwd");
}
extern "C" int mcn_resize_nearest_bwd(int dtype, const void* dy, int N, int H, int W, int C, int Ho,
                                      int Wo, int mode, void* dx, void* stream) {
  MCN_REQUIRE(dy && dx && mode >= 0 && mode <= 2, "resize_nearest_bwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    ::mcn::launch(resize_nearest_bwd_kernel<T>, grid_for((long long)N * H * W * C, 256), 256, 0, st, 
        static_cast<const T*>(dy), N, H, W, C, Ho, Wo, mode, static_cast<T*>(dx));
  });
  return after_launch("resize_nearest_bwd");
}
