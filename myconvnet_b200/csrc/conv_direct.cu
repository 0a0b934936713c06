// CUDA-core convolution kernels: the exact-fp32 path (reference config 1 is fp32), the fallback
// for geometries the TMA path cannot address (RGB inputs/outputs), depthwise convolution
// (tf.nn.depthwise_conv2d, reference convnet.py:1645) and the explicit im2col used by the
// 3-channel stems.  Dense call sites: convnet.py:1659 and its autodiff, :2463.
// Depthwise is bandwidth-bound: threads run along the channel axis so every access is coalesced.
#include <cstdlib>

#include "mcn_common.cuh"
#include "xsum.cuh"

namespace mcn {
// specialised depthwise kernels (dwconv.cu): multiplier 1, 3x3 / 5x5, stride 1 / 2, vector-aligned C
bool dw_fast_eligible(const mcn_conv_desc* d, int mult, int dtype, int wdtype);
int dw_fast_fwd(const mcn_conv_desc* d, int dtype, const void* x, const float* w, void* y, cudaStream_t st);
int dw_fast_bwd_data(const mcn_conv_desc* d, int dtype, const void* dy, const float* w, void* dx, cudaStream_t st);
int dw_fast_bwd_filter(const mcn_conv_desc* d, int dtype, const void* x, const void* dy, float* dw, cudaStream_t st);

namespace {

inline bool dw_fast_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCN_DW_FAST");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

inline int grid_for(long long n, int block) {
  return (int)std::max<long long>(1, std::min<long long>((n + block - 1) / block, 32LL * num_sms()));
}

// y[n,p,q,co..co+CO) for one thread; x broadcast across the co threads, w coalesced along co.
template <typename T, typename TW, int CO>
__global__ void conv_fprop_direct_kernel(mcn_conv_desc d, const T* __restrict__ x,
                                         const TW* __restrict__ w, const float* __restrict__ bias,
                                         T* __restrict__ y) {
  MCN_PDL_PROLOGUE();
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
  MCN_PDL_PROLOGUE();
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
  MCN_PDL_PROLOGUE();
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

// ---------------------------------------------------------------- tiled implicit GEMM (CUDA cores)
// The three kernels above walk one output element per thread; fine for a handful of tiny layers,
// 1.1 s per step for the fp32 ResNet-50 of BASELINE config 1 and 34 ms for one 256->21 logits wgrad
// of DeepLab.  This is the classic shared-memory tiled GEMM on the FP32 pipe (exact fp32 FMA: the
// fp32 config stays a true fp32 path, no TF32), with the convolution's gather folded into the tile
// loads.  C[M x Nn] = A[M x K] * B[K x Nn]:
//   fprop: M = pixels of y,  Nn = Cout, K = taps*Cin ; A = x gathered, B = w[k][co]
//   dgrad: M = pixels of dx, Nn = Cin,  K = taps*Cout; A = dy gathered (exact stride divisions), B = w
//   wgrad: M = taps*Cin,     Nn = Cout, K = pixels   ; A = x gathered, B = dy; K is split over
//          gridDim.z chunks, each writing its own slice (summed in order by splitk_reduce).
// 256 threads, BM = 64 rows x BN = 16*TN columns x BK = 16; a thread owns a 4 x TN micro-tile.
constexpr int kIgBM = 64, kIgBK = 16;
enum { kIgFprop = 0, kIgDgrad = 1, kIgWgrad = 2 };

template <typename T, typename TW, int MODE, int TN>
__global__ void __launch_bounds__(256)
igemm_kernel(mcn_conv_desc d, const T* __restrict__ a_src, const void* __restrict__ b_src,
             const float* __restrict__ bias, void* __restrict__ out, long long slice_stride) {
  MCN_PDL_PROLOGUE();
  constexpr int BN = 16 * TN;
  __shared__ float As[kIgBK][kIgBM + 4];
  __shared__ float Bs[kIgBK][BN + 4];
  const int taps = d.kh * d.kw;
  const long long pix_out = (long long)d.N * d.Ho * d.Wo, pix_in = (long long)d.N * d.H * d.W;
  const long long M = MODE == kIgFprop ? pix_out : (MODE == kIgDgrad ? pix_in : (long long)taps * d.Cin);
  const int Nn = MODE == kIgDgrad ? d.Cin : d.Cout;
  const long long Kall = MODE == kIgFprop ? (long long)taps * d.Cin
                                          : (MODE == kIgDgrad ? (long long)taps * d.Cout : pix_out);
  const long long k_begin = MODE == kIgWgrad ? Kall * blockIdx.z / gridDim.z : 0;
  const long long k_end = MODE == kIgWgrad ? Kall * (blockIdx.z + 1) / gridDim.z : Kall;
  const long long m0 = (long long)blockIdx.x * kIgBM;
  const int n0 = blockIdx.y * BN;
  const int t = threadIdx.x;
  const int tx = t % 16, ty = t / 16;
  // ---- A-tile load pattern: the memory-contiguous index varies fastest across threads
  constexpr bool kAlongK = MODE != kIgWgrad;          // fprop/dgrad: k = channels contiguous
  const int a_k = kAlongK ? t % kIgBK : t / kIgBM;    // + 4*i for wgrad
  const int a_m = kAlongK ? t / kIgBK : t % kIgBM;    // + 16*i for fprop/dgrad
  // per-thread decode of the rows it loads (constant over the K loop)
  int r_n[4], r_h[4], r_w[4];
  bool r_ok[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + (kAlongK ? a_m + 16 * i : a_m);
    r_ok[i] = m < M;
    r_n[i] = r_h[i] = r_w[i] = 0;
    if (!r_ok[i]) continue;
    if (MODE == kIgFprop) {
      const int q = (int)(m % d.Wo);
      const long long r = m / d.Wo;
      r_h[i] = (int)(r % d.Ho) * d.sh - d.pad_t;
      r_w[i] = q * d.sw - d.pad_l;
      r_n[i] = (int)(r / d.Ho);
    } else if (MODE == kIgDgrad) {
      const int w = (int)(m % d.W);
      const long long r = m / d.W;
      r_h[i] = (int)(r % d.H) + d.pad_t;
      r_w[i] = w + d.pad_l;
      r_n[i] = (int)(r / d.H);
    } else {   // wgrad: one row per thread: (tap, ci)
      const int tap = (int)(m / d.Cin);
      r_n[i] = (int)(m % d.Cin);                       // ci
      r_h[i] = (tap / d.kw) * d.dh - d.pad_t;          // tap offset
      r_w[i] = (tap % d.kw) * d.dw - d.pad_l;
    }
  }
  // ---- B-tile load pattern
  constexpr bool kBAlongN = MODE != kIgDgrad;         // fprop: w[k][co]; wgrad: dy[pix][co]
  const int b_n = kBAlongN ? t % BN : t / kIgBK;      // dgrad: + 16*i
  const int b_k = kBAlongN ? t / BN : t % kIgBK;      // fprop/wgrad: + (256/BN)*i
  constexpr int kBRep = (kIgBK * BN) / 256;           // elements per thread (4 for BN 64, 1 for BN 16)
  float acc[4][TN];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (long long k0 = k_begin; k0 < k_end; k0 += kIgBK) {
    // ---- A tile
    if (kAlongK) {
      const long long k = k0 + a_k;
      const bool kok = k < k_end;
      const int cdim = MODE == kIgFprop ? d.Cin : d.Cout;
      const int tap = kok ? (int)(k / cdim) : 0, c = kok ? (int)(k % cdim) : 0;
      const int ta = tap / d.kw, tb = tap % d.kw;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        if (kok && r_ok[i]) {
          if (MODE == kIgFprop) {
            const int h = r_h[i] + ta * d.dh, w = r_w[i] + tb * d.dw;
            if (h >= 0 && h < d.H && w >= 0 && w < d.W)
              v = to_f32(a_src[(((long long)r_n[i] * d.H + h) * d.W + w) * d.Cin + c]);
          } else {
            const int hp = r_h[i] - ta * d.dh, wq = r_w[i] - tb * d.dw;
            if (hp >= 0 && wq >= 0 && hp % d.sh == 0 && wq % d.sw == 0) {
              const int p = hp / d.sh, q = wq / d.sw;
              if (p < d.Ho && q < d.Wo)
                v = to_f32(a_src[(((long long)r_n[i] * d.Ho + p) * d.Wo + q) * d.Cout + c]);
            }
          }
        }
        As[a_k][a_m + 16 * i] = v;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long k = k0 + a_k + 4 * i;          // output pixel
        float v = 0.f;
        if (k < k_end && r_ok[0]) {
          const int q = (int)(k % d.Wo);
          const long long r = k / d.Wo;
          const int p = (int)(r % d.Ho);
          const int n = (int)(r / d.Ho);
          const int h = p * d.sh + r_h[0], w = q * d.sw + r_w[0];
          if (h >= 0 && h < d.H && w >= 0 && w < d.W)
            v = to_f32(a_src[(((long long)n * d.H + h) * d.W + w) * d.Cin + r_n[0]]);
        }
        As[a_k + 4 * i][a_m] = v;
      }
    }
    // ---- B tile
#pragma unroll
    for (int i = 0; i < kBRep; ++i) {
      float v = 0.f;
      if (kBAlongN) {
        const long long k = k0 + b_k + (256 / BN) * i;
        const int n = n0 + b_n;
        if (k < k_end && n < Nn) {
          if (MODE == kIgFprop) v = to_f32(static_cast<const TW*>(b_src)[k * d.Cout + n]);
          else v = to_f32(static_cast<const T*>(b_src)[k * d.Cout + n]);
        }
        Bs[b_k + (256 / BN) * i][b_n] = v;
      } else {
        const long long k = k0 + b_k;                  // (tap, co)
        const int n = n0 + b_n + 16 * i;               // ci
        if (k < k_end && n < Nn) {
          const int tap = (int)(k / d.Cout), co = (int)(k % d.Cout);
          v = to_f32(static_cast<const TW*>(b_src)[((long long)tap * d.Cin + n) * d.Cout + co]);
        }
        if (b_n + 16 * i < BN) Bs[b_k][b_n + 16 * i] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kIgBK; ++kk) {
      float av[4], bv[TN];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) bv[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n >= Nn) continue;
      if (MODE == kIgWgrad)
        static_cast<float*>(out)[(long long)blockIdx.z * slice_stride + m * d.Cout + n] = acc[i][j];
      else
        static_cast<T*>(out)[m * Nn + n] = from_f32<T>(acc[i][j] + (MODE == kIgFprop && bias ? bias[n] : 0.f));
    }
  }
}

template <typename T, typename TW, int MODE>
void launch_igemm(const mcn_conv_desc* d, const void* a, const void* b, const float* bias, void* out,
                  long long slice_stride, int zchunks, cudaStream_t st) {
  const int taps = d->kh * d->kw;
  const long long M = MODE == kIgFprop ? (long long)d->N * d->Ho * d->Wo
                                       : (MODE == kIgDgrad ? (long long)d->N * d->H * d->W : (long long)taps * d->Cin);
  const int Nn = MODE == kIgDgrad ? d->Cin : d->Cout;
  if (Nn <= 16) {
    dim3 grid((unsigned)((M + kIgBM - 1) / kIgBM), (unsigned)((Nn + 15) / 16), (unsigned)zchunks);
    ::mcn::launch(igemm_kernel<T, TW, MODE, 1>, grid, 256, 0, st, *d, static_cast<const T*>(a), b, bias, out, slice_stride);
  } else {
    dim3 grid((unsigned)((M + kIgBM - 1) / kIgBM), (unsigned)((Nn + 63) / 64), (unsigned)zchunks);
    ::mcn::launch(igemm_kernel<T, TW, MODE, 4>, grid, 256, 0, st, *d, static_cast<const T*>(a), b, bias, out, slice_stride);
  }
}

inline bool igemm_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCN_DIRECT_IGEMM");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// ---------------------------------------------------------------- depthwise
template <typename T, typename TW>
__global__ void dwconv_fwd_kernel(mcn_conv_desc d, int mult, const T* __restrict__ x,
                                  const TW* __restrict__ w, T* __restrict__ y) {
  MCN_PDL_PROLOGUE();
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
  MCN_PDL_PROLOGUE();
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
  MCN_PDL_PROLOGUE();
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
  MCN_PDL_PROLOGUE();
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
  MCN_PDL_PROLOGUE();
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
  if (igemm_enabled() && d->Cout >= 8) {
    MCN_DISPATCH_DTYPE(dtype, T, MCN_DISPATCH_W(wdtype, TW,
                                                (launch_igemm<T, TW, kIgFprop>(d, x, w, bias, y, 0, 1, st))));
    return after_launch("conv_fprop_direct");
  }
  MCN_DISPATCH_DTYPE(dtype, T, MCN_DISPATCH_W(wdtype, TW, {
    if (d->Cout % 4 == 0) {
      long long total = (long long)d->N * d->Ho * d->Wo * (d->Cout / 4);
      ::mcn::launch(conv_fprop_direct_kernel<T, TW, 4>, grid_for(total, 128), 128, 0, st, 
          *d, static_cast<const T*>(x), static_cast<const TW*>(w), bias, static_cast<T*>(y));
    } else {
      long long total = (long long)d->N * d->Ho * d->Wo * d->Cout;
      ::mcn::launch(conv_fprop_direct_kernel<T, TW, 1>, grid_for(total, 128), 128, 0, st, 
          *d, static_cast<const T*>(x), static_cast<const TW*>(w), bias, static_cast<T*>(y));
    }
  }));
  return after_launch("conv_fprop_direct");
}

extern "C" int mcn_conv2d_dgrad_direct(const mcn_conv_desc* d, int dtype, const void* dy,
                                       int wdtype, const void* w, void* dx, void* stream) {
  MCN_REQUIRE(d && dy && w && dx, "dgrad_direct: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (igemm_enabled() && d->Cin >= 8) {
    MCN_DISPATCH_DTYPE(dtype, T, MCN_DISPATCH_W(wdtype, TW,
                                                (launch_igemm<T, TW, kIgDgrad>(d, dy, w, nullptr, dx, 0, 1, st))));
    return after_launch("conv_dgrad_direct");
  }
  long long total = (long long)d->N * d->H * d->W * d->Cin;
  MCN_DISPATCH_DTYPE(dtype, T, MCN_DISPATCH_W(wdtype, TW, {
    ::mcn::launch(conv_dgrad_direct_kernel<T, TW>, grid_for(total, 128), 128, 0, st, 
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
  if (igemm_enabled()) {
    // tiled implicit GEMM: enough K chunks to fill the GPU about twice
    const long long tiles = ((long long)d->kh * d->kw * d->Cin + kIgBM - 1) / kIgBM * ((d->Cout + 63) / 64);
    long long z = std::max<long long>(1, std::min<long long>((2LL * num_sms() + tiles - 1) / tiles, pixels / 256 + 1));
    z = std::min(z, std::min<long long>(cap, 65535));
    MCN_DISPATCH_DTYPE(dtype, T, (launch_igemm<T, float, kIgWgrad>(d, x, dy, nullptr, slices, stride, (int)z, st)));
    int rc0 = after_launch("conv_wgrad_direct");
    if (rc0) return rc0;
    return launch_splitk_reduce(slices, stride, (int)z, n, dw, st);
  }
  dim3 grid((unsigned)gx, (unsigned)chunks);
  MCN_DISPATCH_DTYPE(dtype, T, {
    ::mcn::launch(conv_wgrad_direct_kernel<T>, grid, block, 0, st, *d, static_cast<const T*>(x),
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
  if (dw_fast_enabled() && dw_fast_eligible(d, mult, dtype, wdtype))
    return dw_fast_fwd(d, dtype, x, static_cast<const float*>(w), y, st);
  long long total = (long long)d->N * d->Ho * d->Wo * d->Cin * mult;
  MCN_DISPATCH_DTYPE(dtype, T, MCN_DISPATCH_W(wdtype, TW, {
    ::mcn::launch(dwconv_fwd_kernel<T, TW>, grid_for(total, 256), 256, 0, st, 
        *d, mult, static_cast<const T*>(x), static_cast<const TW*>(w), static_cast<T*>(y));
  }));
  return after_launch("dwconv_fwd");
}
extern "C" int mcn_dwconv2d_bwd_data(const mcn_conv_desc* d, int mult, int dtype, const void* dy,
                                     int wdtype, const void* w, void* dx, void* stream) {
  MCN_REQUIRE(d && dy && w && dx && mult >= 1, "dwconv_bwd_data: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dw_fast_enabled() && dw_fast_eligible(d, mult, dtype, wdtype))
    return dw_fast_bwd_data(d, dtype, dy, static_cast<const float*>(w), dx, st);
  long long total = (long long)d->N * d->H * d->W * d->Cin;
  MCN_DISPATCH_DTYPE(dtype, T, MCN_DISPATCH_W(wdtype, TW, {
    ::mcn::launch(dwconv_bwd_data_kernel<T, TW>, grid_for(total, 256), 256, 0, st, 
        *d, mult, static_cast<const T*>(dy), static_cast<const TW*>(w), static_cast<T*>(dx));
  }));
  return after_launch("dwconv_bwd_data");
}
extern "C" int mcn_dwconv2d_bwd_filter(const mcn_conv_desc* d, int mult, int dtype, const void* x,
                                       const void* dy, float* dw, void* stream) {
  MCN_REQUIRE(d && x && dy && dw && mult >= 1, "dwconv_bwd_filter: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dw_fast_enabled() && dw_fast_eligible(d, mult, dtype, MCN_F32))
    return dw_fast_bwd_filter(d, dtype, x, dy, dw, st);
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
    ::mcn::launch(dwconv_bwd_filter_kernel<T>, grid, block, 0, st, *d, mult, static_cast<const T*>(x),
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
      ::mcn::launch(im2col_rows_kernel<T>, grid_for(total, 256), 256, 0, st, 
          *d, static_cast<const T*>(x), static_cast<__nv_bfloat16*>(col), kpad);
    } else {
      long long total = (long long)d->N * d->Ho * d->Wo * kpad;
      ::mcn::launch(im2col_kernel<T>, grid_for(total, 256), 256, 0, st, *d, static_cast<const T*>(x),
                                                            static_cast<__nv_bfloat16*>(col), kpad);
    }
  });
  return after_launch("im2col");
}
