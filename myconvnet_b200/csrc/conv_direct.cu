// CUDA-core convolution kernels: the exact-fp32 path (reference config 1 is fp32), the fallback
// for geometries the TMA path cannot address (RGB inputs/outputs), depthwise convolution
// (tf.nn.depthwise_conv2d, reference convnet.py:1645) and the explicit im2col used by the
// 3-channel stems.  Dense call sites: convnet.py:1659 and its autodiff, :2463.
// Depthwise is bandwidth-bound: threads run along the channel axis so every access is coalesced.
#include "mcn_common.cuh"
#include "xsum.cuh"

namespace mcn {
namespace {

inline int grid_for(long long n, int block) {
  return (int)std::max<long long>(1, std::min<long long>((n + block - 1) / block, 32LL * num_sms()));
}

// y[n,p,q,co..co+CO) for one thread; x broadcast across the co threads, w coalesced along co.
template <typename T, typename TW, int CO>
__global__ void conv_fprop_direct_kernel(mcn_conv_desc d, const T* __restrict__ x,
                                         const TW* __restrict__ w, const float* __restrict__ bias,
                                         T* __restrict__ y) {
  const int cog = d.Cout / CO;
  const long long total = (long long)d.N * d.Ho * d.Wo * cog;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int co = (int)(i % cog) * CO;
    long long r = i / cog;
    int q = (int)(r % d.Wo);
    r /= d.Wo;
    int p = (int)(r % d.Ho);
    int n = (int)(r / d.Ho);
    float acc[CO];
#pragma unroll
    for (int e = 0; e < CO; ++e) acc[e] = bias ? bias[co + e] : 0.f;
    for (int a = 0; a < d.kh; ++a) {
      int h = p * d.sh + a * d.dh - d.pad_t;
      if (h < 0 || h >= d.H) continue;
      for (int b = 0; b < d.kw; ++b) {
        int ww = q * d.sw + b * d.dw - d.pad_l;
        if (ww < 0 || ww >= d.W) continue;
        const T* xp = x + (((long long)n * d.H + h) * d.W + ww) * d.Cin;
        const TW* wp = w + ((long long)(a * d.kw + b) * d.Cin) * d.Cout + co;
        for (int ci = 0; ci < d.Cin; ++ci) {
          float xv = to_f32(xp[ci]);
#pragma unroll
          for (int e = 0; e < CO; ++e) acc[e] = fmaf(xv, to_f32(wp[(long long)ci * d.Cout + e]), acc[e]);
        }
      }
    }
    T* o = y + (((long long)n * d.Ho + p) * d.Wo + q) * d.Cout + co;
#pragma unroll
    for (int e = 0; e < CO; ++e) o[e] = from_f32<T>(acc[e]);
  }
}

// dx[n,h,w,ci] = sum over taps whose window covers (h,w) and over co.
template <typename T, typename TW>
__global__ void conv_dgrad_direct_kernel(mcn_conv_desc d, const T* __restrict__ dy,
                                         const TW* __restrict__ w, T* __restrict__ dx) {
  const long long total = (long long)d.N * d.H * d.W * d.Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int ci = (int)(i % d.Cin);
    long long r = i / d.Cin;
    int ww = (int)(r % d.W);
    r /= d.W;
    int h = (int)(r % d.H);
    int n = (int)(r / d.H);
    float acc = 0.f;
    for (int a = 0; a < d.kh; ++a) {
      int hp = h + d.pad_t - a * d.dh;
      if (hp < 0 || hp % d.sh != 0) continue;
      int p = hp / d.sh;
      if (p >= d.Ho) continue;
      for (int b = 0; b < d.kw; ++b) {
        int wq = ww + d.pad_l - b * d.dw;
        if (wq < 0 || wq % d.sw != 0) continue;
        int q = wq / d.sw;
        if (q >= d.Wo) continue;
        const T* g = dy + (((long long)n * d.Ho + p) * d.Wo + q) * d.Cout;
        const TW* wp = w + ((long long)(a * d.kw + b) * d.Cin + ci) * d.Cout;
        for (int co = 0; co < d.Cout; ++co) acc = fmaf(to_f32(g[co]), to_f32(wp[co]), acc);
      }
    }
    dx[i] = from_f32<T>(acc);
  }
}

// dW[tap,ci,co] over a chunk of pixels; threads along co (coalesced dy reads).  Each pixel chunk
// writes its own slice (slice_stride elements apart); splitk_reduce adds the slices in order.
template <typename T>
__global__ void conv_wgrad_direct_kernel(mcn_conv_desc d, const T* __restrict__ x,
                                         const T* __restrict__ dy, float* __restrict__ dw,
                                         long long slice_stride) {
  // grid.x: (tap*Cin + ci) * ceil(Cout/blockDim.x) ; grid.y: pixel chunks
  const int cob = (d.Cout + blockDim.x - 1) / blockDim.x;
  const int co = (blockIdx.x % cob) * blockDim.x + threadIdx.x;
  const int tc = blockIdx.x / cob;
  const int ci = tc % d.Cin;
  const int tap = tc / d.Cin;
  const int a = tap / d.kw, b = tap % d.kw;
  if (co >= d.Cout) return;
  const long long pixels = (long long)d.N * d.Ho * d.Wo;
  const long long m0 = pixels * blockIdx.y / gridDim.y, m1 = pixels * (blockIdx.y + 1) / gridDim.y;
  float acc = 0.f;
  for (long long m = m0; m < m1; ++m) {
    int q = (int)(m % d.Wo);
    long long r = m / d.Wo;
    int p = (int)(r % d.Ho);
    int n = (int)(r / d.Ho);
    int h = p * d.sh + a * d.dh - d.pad_t, ww = q * d.sw + b * d.dw - d.pad_l;
    if (h < 0 || h >= d.H || ww < 0 || ww >= d.W) continue;
    float xv = to_f32(x[(((long long)n * d.H + h) * d.W + ww) * d.Cin + ci]);
    acc = fmaf(xv, to_f32(dy[m * d.Cout + co]), acc);
  }
  dw[blockIdx.y * slice_stride + ((long long)tap * d.Cin + ci) * d.Cout + co] = acc;
}

// ---------------------------------------------------------------- depthwise
template <typename T, typename TW>
__global__ void dwconv_fwd_kernel(mcn_conv_desc d, int mult, const T* __restrict__ x,
                                  const TW* __restrict__ w, T* __restrict__ y) {
  const int Co = d.Cin * mult;
  const long long total = (long long)d.N * d.Ho * d.Wo * Co;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int oc = (int)(i % Co);
    int c = oc / mult, m = oc % mult;
    long long r = i / Co;
    int q = (int)(r % d.Wo);
    r /= d.Wo;
    int p = (int)(r % d.Ho);
    int n = (int)(r / d.Ho);
    float acc = 0.f;
    for (int a = 0; a < d.kh; ++a) {
      int h = p * d.sh + a * d.dh - d.pad_t;
      if (h < 0 || h >= d.H) continue;
      for (int b = 0; b < d.kw; ++b) {
        int ww = q * d.sw + b * d.dw - d.pad_l;
        if (ww < 0 || ww >= d.W) continue;
        acc = fmaf(to_f32(x[(((long long)n * d.H + h) * d.W + ww) * d.Cin + c]),
                   to_f32(w[((long long)(a * d.kw + b) * d.Cin + c) * mult + m]), acc);
      }
    }
    y[i] = from_f32<T>(acc);
  }
}
template <typename T, typename TW>
__global__ void dwconv_bwd_data_kernel(mcn_conv_desc d, int mult, const T* __restrict__ dy,
                                       const TW* __restrict__ w, T* __restrict__ dx) {
  const int Co = d.Cin * mult;
  const long long total = (long long)d.N * d.H * d.W * d.Cin;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % d.Cin);
    long long r = i / d.Cin;
    int ww = (int)(r % d.W);
    r /= d.W;
    int h = (int)(r % d.H);
    int n = (int)(r / d.H);
    float acc = 0.f;
    for (int a = 0; a < d.kh; ++a) {
      int hp = h + d.pad_t - a * d.dh;
      if (hp < 0 || hp % d.sh != 0) continue;
      int p = hp / d.sh;
      if (p >= d.Ho) continue;
      for (int b = 0; b < d.kw; ++b) {
        int wq = ww + d.pad_l - b * d.dw;
        if (wq < 0 || wq % d.sw != 0) continue;
        int q = wq / d.sw;
        if (q >= d.Wo) continue;
        for (int m = 0; m < mult; ++m)
          acc = fmaf(to_f32(dy[(((long long)n * d.Ho + p) * d.Wo + q) * Co + c * mult + m]),
                     to_f32(w[((long long)(a * d.kw + b) * d.Cin + c) * mult + m]), acc);
      }
    }
    dx[i] = from_f32<T>(acc);
  }
}
// blockDim (32 channels, 8 pixel lanes); grid (ceil(Co/32), taps, pixel chunks)
template <typename T>
__global__ void dwconv_bwd_filter_kernel(mcn_conv_desc d, int mult, const T* __restrict__ x,
                                         const T* __restrict__ dy, float* __restrict__ dw,
                                         long long slice_stride) {
  __shared__ float sh[8][33];
  const int Co = d.Cin * mult;
  const int oc = blockIdx.x * 32 + threadIdx.x;
  const int tap = blockIdx.y;
  const int a = tap / d.kw, b = tap % d.kw;
  const long long pixels = (long long)d.N * d.Ho * d.Wo;
  const long long m0 = pixels * blockIdx.z / gridDim.z, m1 = pixels * (blockIdx.z + 1) / gridDim.z;
  float acc = 0.f;
  if (oc < Co) {
    const int c = oc / mult;
    for (long long m = m0 + threadIdx.y; m < m1; m += 8) {
      int q = (int)(m % d.Wo);
      long long r = m / d.Wo;
      int p = (int)(r % d.Ho);
      int n = (int)(r / d.Ho);
      int h = p * d.sh + a * d.dh - d.pad_t, ww = q * d.sw + b * d.dw - d.pad_l;
      if (h < 0 || h >= d.H || ww < 0 || ww >= d.W) continue;
      acc = fmaf(to_f32(x[(((long long)n * d.H + h) * d.W + ww) * d.Cin + c]),
                 to_f32(dy[m * Co + oc]), acc);
    }
  }
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && oc < Co) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += sh[j][threadIdx.x];
    dw[blockIdx.z * slice_stride + (long long)tap * Co + oc] = s;  // [tap][c][m] == [tap][oc]
  }
}

// ---------------------------------------------------------------- explicit im2col (bf16 out)
// Row m of `col` is the receptive field of output pixel m laid out [kh][kw][Cin] and zero-padded
// to kpad.  With unit dilation along W the kw*Cin elements of one filter row are CONTIGUOUS in
// the NHWC input, so a thread builds 8 consecutive k (one 16-byte store) from at most two short
// contiguous runs of the input; only image-border masking is needed.
template <typename T>
__global__ void im2col_rows_kernel(mcn_conv_desc d, const T* __restrict__ x,
                                   __nv_bfloat16* __restrict__ col, int kpad) {
  const int K = d.kh * d.kw * d.Cin;
  const int RL = d.kw * d.Cin;        // run length of one filter row
  const int WL = d.W * d.Cin;         // elements in one input row
  const int groups = kpad / 8;
  const long long total = (long long)d.N * d.Ho * d.Wo * groups;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    const long long m = i / groups;
    const int q = (int)(m % d.Wo);
    long long rr = m / d.Wo;
    const int p = (int)(rr % d.Ho);
    const int n = (int)(rr / d.Ho);
    const int h0 = p * d.sh - d.pad_t;
    const int lin0 = (q * d.sw - d.pad_l) * d.Cin;
    int k = g * 8;
    int r = k / RL;
    int rem = k - r * RL;
    uint32_t pk[4] = {0, 0, 0, 0};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = 0.f;
      if (k < K) {
        const int h = h0 + r * d.dh;
        const int lin = lin0 + rem;
        if (h >= 0 && h < d.H && lin >= 0 && lin < WL)
          v = to_f32(x[((long long)n * d.H + h) * WL + lin]);
      }
      const uint32_t b = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
      pk[e >> 1] |= (e & 1) ? (b << 16) : b;
      ++k;
      if (++rem == RL) {
        rem = 0;
        ++r;
      }
    }
    *reinterpret_cast<uint4*>(col + m * kpad + g * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// general fallback (dilated along W)
template <typename T>
__global__ void im2col_kernel(mcn_conv_desc d, const T* __restrict__ x,
                              __nv_bfloat16* __restrict__ col, int kpad) {
  const int K = d.kh * d.kw * d.Cin;
  const long long total = (long long)d.N * d.Ho * d.Wo * kpad;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int k = (int)(i % kpad);
    long long m = i / kpad;
    float v = 0.f;
    if (k < K) {
      int ci = k % d.Cin;
      int tap = k / d.Cin;
      int a = tap / d.kw, b = tap % d.kw;
      int q = (int)(m % d.Wo);
      long long r = m / d.Wo;
      int p = (int)(r % d.Ho);
      int n = (int)(r / d.Ho);
      int h = p * d.sh + a * d.dh - d.pad_t, ww = q * d.sw + b * d.dw - d.pad_l;
      if (h >= 0 && h < d.H && ww >= 0 && ww < d.W)
        v = to_f32(x[(((long long)n * d.H + h) * d.W + ww) * d.Cin + ci]);
    }
    col[i] = __float2bfloat16_rn(v);
  }
}

}  // namespace
}  // namespace mcn

using namespace mcn;

#define MCN_DISPATCH_W(wdtype, TW, ...)                 \
  do {                                                  \
    if ((wdtype) == MCN_F32) {                          \
      using TW = float;                                 \
      __VA_ARGS__;                                      \
    } else {                                            \
      using TW = __nv_bfloat16;                         \
      __VA_ARGS__;                                      \
    }                                                   \
  } while (0)

extern "C" int mcn_conv2d_fprop_direct(const mcn_conv_desc* d, int dtype, const void* x,
                                       int wdtype, const void* w, const float* bias, void* y,
                                       void* stream) {
  MCN_REQUIRE(d && x && w && y, "fprop_direct: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, MCN_DISPATCH_W(wdtype, TW, {
    if (d->Cout % 4 == 0) {
      long long total = (long long)d->N * d->Ho * d->Wo * (d->Cout / 4);
      conv_fprop_direct_kernel<T, TW, 4><<<grid_for(total, 128), 128, 0, st>>>(
          *d, static_cast<const T*>(x), static_cast<const TW*>(w), bias, static_cast<T*>(y));
    } else {
      long long total = (long long)d->N * d->Ho * d->Wo * d->Cout;
      conv_fprop_direct_kernel<T, TW, 1><<<grid_for(total, 128), 128, 0, st>>>(
          *d, static_cast<const T*>(x), static_cast<const TW*>(w), bias, static_cast<T*>(y));
    }
  }));
  return after_launch("conv_fprop_direct");
}

extern "C" int mcn_conv2d_dgrad_direct(const mcn_conv_desc* d, int dtype, const void* dy,
                                       int wdtype, const void* w, void* dx, void* stream) {
  MCN_REQUIRE(d && dy && w && dx, "dgrad_direct: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long total = (long long)d->N * d->H * d->W * d->Cin;
  MCN_DISPATCH_DTYPE(dtype, T, MCN_DISPATCH_W(wdtype, TW, {
    conv_dgrad_direct_kernel<T, TW><<<grid_for(total, 128), 128, 0, st>>>(
        *d, static_cast<const T*>(dy), static_cast<const TW*>(w), static_cast<T*>(dx));
  }));
  return after_launch("conv_dgrad_direct");
}

extern "C" int mcn_conv2d_wgrad_direct(const mcn_conv_desc* d, int dtype, const void* x,
                                       const void* dy, float* dw, void* stream) {
  MCN_REQUIRE(d && x && dy && dw, "wgrad_direct: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int block = d->Cout >= 128 ? 128 : (d->Cout >= 64 ? 64 : 32);
  const int cob = (d->Cout + block - 1) / block;
  const long long gx = (long long)d->kh * d->kw * d->Cin * cob;
  const long long pixels = (long long)d->N * d->Ho * d->Wo;
  long long chunks = std::max<long long>(1, std::min<long long>(pixels / 256, (8LL * num_sms() + gx - 1) / gx));
  chunks = std::min<long long>(chunks, 65535);
  // deterministic: one workspace slice per pixel chunk, summed in chunk order afterwards
  const long long n = (long long)d->kh * d->kw * d->Cin * d->Cout;
  const long long stride = (n + 63) / 64 * 64;
  const Workspace w = current_workspace();
  MCN_REQUIRE(w.base != nullptr, "wgrad_direct: no workspace registered (mcn_set_workspace)");
  const long long cap = (w.bytes - kWsSplitOff) / (stride * 4);
  MCN_REQUIRE(cap >= 1, "wgrad_direct: workspace too small (%lld bytes, one slice needs %lld)",
              w.bytes, kWsSplitOff + stride * 4);
  chunks = std::min(chunks, cap);
  float* slices = reinterpret_cast<float*>(w.base + kWsSplitOff);
  dim3 grid((unsigned)gx, (unsigned)chunks);
  MCN_DISPATCH_DTYPE(dtype, T, {
    conv_wgrad_direct_kernel<T><<<grid, block, 0, st>>>(*d, static_cast<const T*>(x),
                                                        static_cast<const T*>(dy), slices, stride);
  });
  int rc = after_launch("conv_wgrad_direct");
  if (rc) return rc;
  return launch_splitk_reduce(slices, stride, (int)chunks, n, dw, st);
}

extern "C" int mcn_dwconv2d_fwd(const mcn_conv_desc* d, int mult, int dtype, const void* x,
                                int wdtype, const void* w, void* y, void* stream) {
  MCN_REQUIRE(d && x && w && y && mult >= 1, "dwconv_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long total = (long long)d->N * d->Ho * d->Wo * d->Cin * mult;
  MCN_DISPATCH_DTYPE(dtype, T, MCN_DISPATCH_W(wdtype, TW, {
    dwconv_fwd_kernel<T, TW><<<grid_for(total, 256), 256, 0, st>>>(
        *d, mult, static_cast<const T*>(x), static_cast<const TW*>(w), static_cast<T*>(y));
  }));
  return after_launch("dwconv_fwd");
}
extern "C" int mcn_dwconv2d_bwd_data(const mcn_conv_desc* d, int mult, int dtype, const void* dy,
                                     int wdtype, const void* w, void* dx, void* stream) {
  MCN_REQUIRE(d && dy && w && dx && mult >= 1, "dwconv_bwd_data: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long total = (long long)d->N * d->H * d->W * d->Cin;
  MCN_DISPATCH_DTYPE(dtype, T, MCN_DISPATCH_W(wdtype, TW, {
    dwconv_bwd_data_kernel<T, TW><<<grid_for(total, 256), 256, 0, st>>>(
        *d, mult, static_cast<const T*>(dy), static_cast<const TW*>(w), static_cast<T*>(dx));
  }));
  return after_launch("dwconv_bwd_data");
}
extern "C" int mcn_dwconv2d_bwd_filter(const mcn_conv_desc* d, int mult, int dtype, const void* x,
                                       const void* dy, float* dw, void* stream) {
  MCN_REQUIRE(d && x && dy && dw && mult >= 1, "dwconv_bwd_filter: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int Co = d->Cin * mult;
  const long long pixels = (long long)d->N * d->Ho * d->Wo;
  const long long gx = (Co + 31) / 32, taps = (long long)d->kh * d->kw;
  long long chunks = std::max<long long>(1, std::min<long long>(pixels / 64, (8LL * num_sms() + gx * taps - 1) / (gx * taps)));
  chunks = std::min<long long>(chunks, 65535);
  const long long n = taps * Co;
  const long long stride = (n + 63) / 64 * 64;
  const Workspace w = current_workspace();
  MCN_REQUIRE(w.base != nullptr, "dwconv_bwd_filter: no workspace registered (mcn_set_workspace)");
  const long long cap = (w.bytes - kWsSplitOff) / (stride * 4);
  MCN_REQUIRE(cap >= 1, "dwconv_bwd_filter: workspace too small");
  chunks = std::min(chunks, cap);
  float* slices = reinterpret_cast<float*>(w.base + kWsSplitOff);
  dim3 grid((unsigned)gx, (unsigned)taps, (unsigned)chunks), block(32, 8);
  MCN_DISPATCH_DTYPE(dtype, T, {
    dwconv_bwd_filter_kernel<T><<<grid, block, 0, st>>>(*d, mult, static_cast<const T*>(x),
                                                        static_cast<const T*>(dy), slices, stride);
  });
  int rc = after_launch("dwconv_bwd_filter");
  if (rc) return rc;
  return launch_splitk_reduce(slices, stride, (int)chunks, n, dw, st);
}

extern "C" int mcn_im2col(const mcn_conv_desc* d, int dtype, const void* x, void* col, int kpad,
                          void* stream) {
  MCN_REQUIRE(d && x && col && kpad >= d->kh * d->kw * d->Cin && kpad % 8 == 0,
              "im2col: bad argument (kpad must be >= kh*kw*Cin and a multiple of 8)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    if (d->dw == 1) {
      long long total = (long long)d->N * d->Ho * d->Wo * (kpad / 8);
      im2col_rows_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(
          *d, static_cast<const T*>(x), static_cast<__nv_bfloat16*>(col), kpad);
    } else {
      long long total = (long long)d->N * d->Ho * d->Wo * kpad;
      im2col_kernel<T><<<grid_for(total, 256), 256, 0, st>>>(*d, static_cast<const T*>(x),
                                                            static_cast<__nv_bfloat16*>(col), kpad);
    }
  });
  return after_launch("im2col");
}
