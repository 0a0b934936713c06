// Pooling, NHWC, bandwidth-bound.
// Replaces tf.nn.max_pool (reference convnet.py:1509), tf.nn.avg_pool (convnet.py:1548) and the
// tf.reduce_mean(axis=[1,2]) global pooling of resnet_v1_5.py:73 / efficientnet.py:108,183.
// Each thread owns one output pixel x one 16-byte channel vector, so every load and store is a
// coalesced 16-byte access along C.
#include <cfloat>

#include "mcn_common.cuh"

namespace mcn {
namespace {

// Max pooling.  Padding acts as -inf; ties resolve to the first maximum in row-major window
// order (TF CPU rule).  argmax = (h*W + w)*C + c inside the image (max_pool_with_argmax
// convention without the batch term).
template <typename T, int V>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C, int kh,
                                   int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo,
                                   T* __restrict__ y, int32_t* __restrict__ argmax) {
  const int cv = C / V;
  const long long total = (long long)N * Ho * Wo * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c0 = (int)(i % cv) * V;
    long long r = i / cv;
    int q = (int)(r % Wo);
    r /= Wo;
    int p = (int)(r % Ho);
    int n = (int)(r / Ho);
    float best[V];
    int bidx[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      best[e] = -FLT_MAX;
      bidx[e] = -1;
    }
    for (int a = 0; a < kh; ++a) {
      int h = p * sh + a - pad_t;
      if (h < 0 || h >= H) continue;
      for (int b = 0; b < kw; ++b) {
        int w = q * sw + b - pad_l;
        if (w < 0 || w >= W) continue;
        const T* src = x + (((long long)n * H + h) * W + w) * C + c0;
        float v[V];
        if (V == 1) {
          v[0] = to_f32(src[0]);
        } else {
          Vec16<T> t = ld_vec(src);
#pragma unroll
          for (int e = 0; e < V; ++e) v[e] = t.get(e);
        }
#pragma unroll
        for (int e = 0; e < V; ++e)
          if (v[e] > best[e] || bidx[e] < 0) {
            best[e] = v[e];
            bidx[e] = (h * W + w) * C + c0 + e;
          }
      }
    }
    long long o = (((long long)n * Ho + p) * Wo + q) * C + c0;
    if (V == 1) {
      y[o] = from_f32<T>(best[0]);
      if (argmax) argmax[o] = bidx[0];
    } else {
      // one 16-byte store for the values, 16-byte stores for the indices
      Vec16<T> ov;
#pragma unroll
      for (int e = 0; e < V; ++e) ov.set(e, best[e]);
      st_vec(y + o, ov);
      if (argmax) {
#pragma unroll
        for (int e = 0; e < V; e += 4)
          *reinterpret_cast<int4*>(argmax + o + e) = make_int4(bidx[e], bidx[e + 1], bidx[e + 2], bidx[e + 3]);
      }
    }
  }
}

// Backward: each input position gathers from the (few) windows that cover it — deterministic,
// no atomics.
template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ dy, const int32_t* __restrict__ argmax,
                                   int N, int H, int W, int C, int kh, int kw, int sh, int sw,
                                   int pad_t, int pad_l, int Ho, int Wo, T* __restrict__ dx) {
  const long long total = (long long)N * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long r = i / C;
    int w = (int)(r % W);
    r /= W;
    int h = (int)(r % H);
    int n = (int)(r / H);
    const int self = (h * W + w) * C + c;
    // windows p with p*sh - pad_t <= h <= p*sh - pad_t + kh - 1
    int p_lo = (h + pad_t - kh + 1 + sh - 1);
    p_lo = p_lo <= 0 ? 0 : p_lo / sh;
    int p_hi = min((h + pad_t) / sh, Ho - 1);
    int q_lo = (w + pad_l - kw + 1 + sw - 1);
    q_lo = q_lo <= 0 ? 0 : q_lo / sw;
    int q_hi = min((w + pad_l) / sw, Wo - 1);
    float acc = 0.f;
    for (int p = p_lo; p <= p_hi; ++p)
      for (int q = q_lo; q <= q_hi; ++q) {
        long long o = (((long long)n * Ho + p) * Wo + q) * C + c;
        if (argmax[o] == self) acc += to_f32(dy[o]);
      }
    dx[i] = from_f32<T>(acc);
  }
}

// Vectorised backward: one thread per input pixel x 8-channel (16-byte) vector.
template <typename T, int V>
__global__ void maxpool_bwd_vec_kernel(const T* __restrict__ dy, const int32_t* __restrict__ argmax,
                                       int N, int H, int W, int C, int kh, int kw, int sh, int sw,
                                       int pad_t, int pad_l, int Ho, int Wo, T* __restrict__ dx) {
  const int cv = C / V;
  // 32-bit index arithmetic (the host falls back to the scalar kernel beyond 2^31 vectors): the
  // 64-bit divisions were a third of this kernel's instructions
  const uint32_t total = (uint32_t)N * H * W * cv;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c0 = (int)(i % (uint32_t)cv) * V;
    uint32_t r = i / (uint32_t)cv;
    const int w = (int)(r % (uint32_t)W);
    r /= (uint32_t)W;
    const int h = (int)(r % (uint32_t)H);
    const int n = (int)(r / (uint32_t)H);
    const int self = (h * W + w) * C + c0;
    int p_lo = (h + pad_t - kh + 1 + sh - 1);
    p_lo = p_lo <= 0 ? 0 : p_lo / sh;
    const int p_hi = min((h + pad_t) / sh, Ho - 1);
    int q_lo = (w + pad_l - kw + 1 + sw - 1);
    q_lo = q_lo <= 0 ? 0 : q_lo / sw;
    const int q_hi = min((w + pad_l) / sw, Wo - 1);
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    for (int p = p_lo; p <= p_hi; ++p)
      for (int q = q_lo; q <= q_hi; ++q) {
        const long long o = (((long long)n * Ho + p) * Wo + q) * C + c0;
        Vec16<T> g = ld_vec(dy + o);
        int idx[V];
#pragma unroll
        for (int e = 0; e < V; e += 4) {
          int4 a = *reinterpret_cast<const int4*>(argmax + o + e);
          idx[e] = a.x;
          idx[e + 1] = a.y;
          idx[e + 2] = a.z;
          idx[e + 3] = a.w;
        }
#pragma unroll
        for (int e = 0; e < V; ++e)
          if (idx[e] == self + e) acc[e] += g.get(e);
      }
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < V; ++e) o.set(e, acc[e]);
    st_vec(dx + (((long long)n * H + h) * W + w) * C + c0, o);
  }
}

// Average pooling; SAME divides by the number of in-bounds elements.
template <typename T>
__global__ void avgpool_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C, int kh,
                                   int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo,
                                   T* __restrict__ y) {
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long r = i / C;
    int q = (int)(r % Wo);
    r /= Wo;
    int p = (int)(r % Ho);
    int n = (int)(r / Ho);
    float acc = 0.f;
    int cnt = 0;
    for (int a = 0; a < kh; ++a) {
      int h = p * sh + a - pad_t;
      if (h < 0 || h >= H) continue;
      for (int b = 0; b < kw; ++b) {
        int w = q * sw + b - pad_l;
        if (w < 0 || w >= W) continue;
        acc += to_f32(x[(((long long)n * H + h) * W + w) * C + c]);
        ++cnt;
      }
    }
    y[i] = from_f32<T>(acc / (float)max(cnt, 1));
  }
}
template <typename T>
__global__ void avgpool_bwd_kernel(const T* __restrict__ dy, int N, int H, int W, int C, int kh,
                                   int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo,
                                   T* __restrict__ dx) {
  const long long total = (long long)N * H * W * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long r = i / C;
    int w = (int)(r % W);
    r /= W;
    int h = (int)(r % H);
    int n = (int)(r / H);
    int p_lo = (h + pad_t - kh + 1 + sh - 1);
    p_lo = p_lo <= 0 ? 0 : p_lo / sh;
    int p_hi = min((h + pad_t) / sh, Ho - 1);
    int q_lo = (w + pad_l - kw + 1 + sw - 1);
    q_lo = q_lo <= 0 ? 0 : q_lo / sw;
    int q_hi = min((w + pad_l) / sw, Wo - 1);
    float acc = 0.f;
    for (int p = p_lo; p <= p_hi; ++p) {
      int h0 = max(p * sh - pad_t, 0), h1 = min(p * sh - pad_t + kh, H);
      for (int q = q_lo; q <= q_hi; ++q) {
        int w0 = max(q * sw - pad_l, 0), w1 = min(q * sw - pad_l + kw, W);
        int cnt = (h1 - h0) * (w1 - w0);
        acc += to_f32(dy[(((long long)n * Ho + p) * Wo + q) * C + c]) / (float)max(cnt, 1);
      }
    }
    dx[i] = from_f32<T>(acc);
  }
}

// Global average pool: one block per (image, 32-channel-vector slab); threads split HW.
template <typename T, typename TO>
__global__ void gap_fwd_kernel(const T* __restrict__ x, int HW, int C, TO* __restrict__ y) {
  // blockDim = (32, 8): x -> channel, y -> spatial lane
  __shared__ float sh[8][33];
  int n = blockIdx.y;
  int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C)
    for (int i = threadIdx.y; i < HW; i += 8) acc += to_f32(x[((long long)n * HW + i) * C + c]);
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += sh[j][threadIdx.x];
    y[(long long)n * C + c] = from_f32<TO>(s / (float)HW);
  }
}
template <typename T, typename TI>
__global__ void gap_bwd_kernel(const TI* __restrict__ dy, int HW, int C, long long total,
                               T* __restrict__ dx) {
  const float inv = 1.f / (float)HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long n = i / ((long long)HW * C);
    dx[i] = from_f32<T>(to_f32(dy[n * C + c]) * inv);
  }
}

inline int grid_for(long long n, int block) {
  return (int)std::max<long long>(1, std::min<long long>((n + block - 1) / block, 16LL * num_sms()));
}

}  // namespace
}  // namespace mcn

using namespace mcn;

extern "C" int mcn_maxpool_fwd(int dtype, const void* x, int N, int H, int W, int C, int kh,
                               int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo,
                               void* y, int32_t* argmax, void* stream) {
  MCN_REQUIRE(x && y, "maxpool_fwd: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec16<T>::N;
    if (C % V == 0) {
      long long total = (long long)N * Ho * Wo * (C / V);
      maxpool_fwd_kernel<T, V><<<grid_for(total, 256), 256, 0, st>>>(
          static_cast<const T*>(x), N, H, W, C, kh, kw, sh, sw, pad_t, pad_l, Ho, Wo,
          static_cast<T*>(y), argmax);
    } else {
      long long total = (long long)N * Ho * Wo * C;
      maxpool_fwd_kernel<T, 1><<<grid_for(total, 256), 256, 0, st>>>(
          static_cast<const T*>(x), N, H, W, C, kh, kw, sh, sw, pad_t, pad_l, Ho, Wo,
          static_cast<T*>(y), argmax);
    }
  });
  return after_launch("maxpool_fwd");
}

extern "C" int mcn_maxpool_bwd(int dtype, const void* dy, const int32_t* argmax, int N, int H,
                               int W, int C, int kh, int kw, int sh, int sw, int pad_t, int pad_l,
                               int Ho, int Wo, void* dx, void* stream) {
  MCN_REQUIRE(dy && argmax && dx, "maxpool_bwd: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec16<T>::N;
    if (C % V == 0 && (long long)N * H * W * (C / V) < (1LL << 31) - (1 << 24) &&
        (long long)N * Ho * Wo * C < (1LL << 62)) {
      long long total = (long long)N * H * W * (C / V);
      maxpool_bwd_vec_kernel<T, V><<<grid_for(total, 256), 256, 0, st>>>(
          static_cast<const T*>(dy), argmax, N, H, W, C, kh, kw, sh, sw, pad_t, pad_l, Ho, Wo,
          static_cast<T*>(dx));
    } else {
      long long total = (long long)N * H * W * C;
      maxpool_bwd_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(
          static_cast<const T*>(dy), argmax, N, H, W, C, kh, kw, sh, sw, pad_t, pad_l, Ho, Wo,
          static_cast<T*>(dx));
    }
  });
  return after_launch("maxpool_bwd");
}

extern "C" int mcn_avgpool_fwd(int dtype, const void* x, int N, int H, int W, int C, int kh,
                               int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo,
                               void* y, void* stream) {
  MCN_REQUIRE(x && y, "avgpool_fwd: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    long long total = (long long)N * Ho * Wo * C;
    avgpool_fwd_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(
        static_cast<const T*>(x), N, H, W, C, kh, kw, sh, sw, pad_t, pad_l, Ho, Wo,
        static_cast<T*>(y));
  });
  return after_launch("avgpool_fwd");
}
extern "C" int mcn_avgpool_bwd(int dtype, const void* dy, int N, int H, int W, int C, int kh,
                               int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo,
                               void* dx, void* stream) {
  MCN_REQUIRE(dy && dx, "avgpool_bwd: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    long long total = (long long)N * H * W * C;
    avgpool_bwd_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(
        static_cast<const T*>(dy), N, H, W, C, kh, kw, sh, sw, pad_t, pad_l, Ho, Wo,
        static_cast<T*>(dx));
  });
  return after_launch("avgpool_bwd");
}

extern "C" int mcn_gap_fwd(int dtype, const void* x, int N, int HW, int C, void* y, int y_dtype,
                           void* stream) {
  MCN_REQUIRE(x && y, "gap_fwd: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((C + 31) / 32, N), block(32, 8);
  MCN_DISPATCH_DTYPE(dtype, T, {
    if (y_dtype == MCN_F32)
      gap_fwd_kernel<T, float><<<grid, block, 0, st>>>(static_cast<const T*>(x), HW, C,
                                                       static_cast<float*>(y));
    else
      gap_fwd_kernel<T, __nv_bfloat16><<<grid, block, 0, st>>>(
          static_cast<const T*>(x), HW, C, static_cast<__nv_bfloat16*>(y));
  });
  return after_launch("gap_fwd");
}
extern "C" int mcn_gap_bwd(int dtype, const void* dy, int dy_dtype, int N, int HW, int C, void* dx,
                           void* stream) {
  MCN_REQUIRE(dy && dx, "gap_bwd: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long total = (long long)N * HW * C;
  MCN_DISPATCH_DTYPE(dtype, T, {
    if (dy_dtype == MCN_F32)
      gap_bwd_kernel<T, float><<<grid_for(total, 256), 256, 0, st>>>(
          static_cast<const float*>(dy), HW, C, total, static_cast<T*>(dx));
    else
      gap_bwd_kernel<T, __nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(
          static_cast<const __nv_bfloat16*>(dy), HW, C, total, static_cast<T*>(dx));
  });
  return after_launch("gap_bwd");
}
