// Element-wise and data-movement kernels (all bandwidth-bound, 16-byte vectorised where the
// shape allows).  Replaces the Eigen element-wise ops behind reference convnet.py:2500-2556
// (activations, stochastic_depth add), :1694 (bias_add), :452/:466/:471 (input zero-centre,
// scale, cast), efficientnet.py:163 (SE excite), tf.concat (deeplabv3plus.py:100,110) and
// tf.image.resize_bilinear (convnet.py:2397).
#include <cmath>

#include "mcn_common.cuh"
#include "xsum.cuh"

namespace mcn {
namespace {

inline int grid_for(long long n, int block) {
  return (int)std::max<long long>(1, std::min<long long>((n + block - 1) / block, 16LL * num_sms()));
}

// Generic vectorised map over up to two inputs.  F: float(float a, float b).
template <typename T, bool kTwo, typename F>
__global__ void map_kernel(const T* __restrict__ a, const T* __restrict__ b, long long n,
                           T* __restrict__ y, F f) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  const long long nvec = n / V;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
    Vec16<T> va = ld_vec(a + v * V), vb, o;
    if (kTwo) vb = ld_vec(b + v * V);
#pragma unroll
    for (int i = 0; i < V; ++i) o.set(i, f(va.get(i), kTwo ? vb.get(i) : 0.f));
    st_vec(y + v * V, o);
  }
  // tail
  for (long long i = nvec * V + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += stride)
    y[i] = from_f32<T>(f(to_f32(a[i]), kTwo ? to_f32(b[i]) : 0.f));
}

template <typename T, bool kTwo, typename F>
int launch_map(const void* a, const void* b, long long n, void* y, cudaStream_t st, F f,
               const char* what) {
  constexpr int V = Vec16<T>::N;
  ::mcn::launch(map_kernel<T, kTwo, F>, grid_for((n + V - 1) / V, 256), 256, 0, st, 
      static_cast<const T*>(a), static_cast<const T*>(b), n, static_cast<T*>(y), f);
  return after_launch(what);
}

struct ActF {
  int act;
  float alpha;
  __device__ float operator()(float a, float) const { return act_fwd(act, a, alpha); }
};
struct ActB {  // a = dy, b = x (pre-activation)
  int act;
  float alpha;
  __device__ float operator()(float a, float b) const { return a * act_grad_from_x(act, b, alpha); }
};
struct AddActF {
  int act;
  float alpha;
  __device__ float operator()(float a, float b) const { return act_fwd(act, a + b, alpha); }
};
struct AddActB {  // a = dy, b = y (output)
  int act;
  float alpha;
  __device__ float operator()(float a, float b) const { return a * act_grad_from_y(act, b, alpha); }
};
struct AddF {
  __device__ float operator()(float a, float b) const { return a + b; }
};

template <typename T>
__global__ void scale_bcast_fwd_kernel(const T* __restrict__ x, const T* __restrict__ m, int HW,
                                       int C, long long total, T* __restrict__ y) {
  MCN_PDL_PROLOGUE();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long n = i / ((long long)HW * C);
    y[i] = from_f32<T>(to_f32(x[i]) * to_f32(m[n * C + c]));
  }
}
// dx = dy*m; dm[n,c] = sum_hw dy*x.  One block per (n, 32 channels).
template <typename T>
__global__ void scale_bcast_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ x,
                                       const T* __restrict__ m, int HW, int C,
                                       T* __restrict__ dx, float* __restrict__ dm) {
  MCN_PDL_PROLOGUE();
  __shared__ float sh[8][33];
  int n = blockIdx.y;
  int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C) {
    float mv = to_f32(m[(long long)n * C + c]);
    for (int i = threadIdx.y; i < HW; i += 8) {
      long long o = ((long long)n * HW + i) * C + c;
      float g = to_f32(dy[o]);
      acc = fmaf(g, to_f32(x[o]), acc);
      dx[o] = from_f32<T>(g * mv);
    }
  }
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += sh[j][threadIdx.x];
    dm[(long long)n * C + c] = s;
  }
}

// 16-byte channel vectors (C % V == 0): one index computation per V channels, 16-byte accesses.
template <typename T>
__global__ void __launch_bounds__(256)
scale_bcast_fwd_vec_kernel(const T* __restrict__ x, const T* __restrict__ m, int HW, int cv, long long nvec,
                           T* __restrict__ y) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  const long long per_img = (long long)HW * cv;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += 4 * stride) {
    Vec16<T> a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long v = v0 + u * stride;
      if (v < nvec) {
        const long long n = v / per_img;
        const int c = (int)(v % cv);
        a[u] = ld_vec_stream(x + v * V);
        b[u] = ld_vec(m + (n * cv + c) * V);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long v = v0 + u * stride;
      if (v < nvec) {
        Vec16<T> o;
#pragma unroll
        for (int e = 0; e < V; ++e) o.set(e, a[u].get(e) * b[u].get(e));
        st_vec(y + v * V, o);
      }
    }
  }
}
// dx = dy*m; dm[n,c] = sum_hw dy*x.  One block per (image, slab of kSlab channel vectors): 256 threads =
// kSlab vector lanes x 256/kSlab row lanes, four rows in flight per thread, fixed-order reduction over the
// row lanes in shared memory (no atomics).
template <typename T, int kSlab>
__global__ void __launch_bounds__(256)
scale_bcast_bwd_vec_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ m, int HW,
                           int cv, T* __restrict__ dx, float* __restrict__ dm) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  constexpr int kRows = 256 / kSlab;
  __shared__ float sh[kRows][kSlab * V + 1];
  const int n = blockIdx.y;
  const int sv = threadIdx.x % kSlab, rl = threadIdx.x / kSlab;
  const int vec = blockIdx.x * kSlab + sv;
  float acc[V];
#pragma unroll
  for (int e = 0; e < V; ++e) acc[e] = 0.f;
  if (vec < cv) {
    const Vec16<T> mv = ld_vec(m + ((long long)n * cv + vec) * V);
    const long long base = (long long)n * HW * cv + vec;
    for (int i0 = rl; i0 < HW; i0 += 4 * kRows) {
      Vec16<T> g[4], a[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kRows;
        if (i < HW) {
          const long long o = (base + (long long)i * cv) * V;
          g[u] = ld_vec_stream(dy + o);
          a[u] = ld_vec_stream(x + o);
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kRows;
        if (i < HW) {
          Vec16<T> o;
#pragma unroll
          for (int e = 0; e < V; ++e) {
            acc[e] = fmaf(g[u].get(e), a[u].get(e), acc[e]);
            o.set(e, g[u].get(e) * mv.get(e));
          }
          st_vec(dx + (base + (long long)i * cv) * V, o);
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < V; ++e) sh[rl][sv * V + e] = acc[e];
  __syncthreads();
  for (int j = threadIdx.x; j < kSlab * V; j += 256) {
    const int c = blockIdx.x * kSlab * V + j;
    if (c < cv * V) {
      float sum = 0.f;
      for (int r = 0; r < kRows; ++r) sum += sh[r][j];
      dm[(long long)n * cv * V + c] = sum;
    }
  }
}

template <typename T>
__global__ void bias_add_kernel(T* __restrict__ y, long long total, int C,
                                const float* __restrict__ bias) {
  MCN_PDL_PROLOGUE();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    y[i] = from_f32<T>(to_f32(y[i]) + bias[i % C]);
}
// db[c] += sum_rows dy[r, c]; blockDim (32, 8), grid (C/32, row chunks)
template <typename T>
__global__ void bias_grad_kernel(const T* __restrict__ dy, long long rows, int C,
                                 float* __restrict__ db, XsScratch xsc) {
  MCN_PDL_PROLOGUE();
  __shared__ float sh[8][33];
  int c = blockIdx.x * 32 + threadIdx.x;
  long long r0 = rows * blockIdx.y / gridDim.y, r1 = rows * (blockIdx.y + 1) / gridDim.y;
  float acc = 0.f;
  if (c < C)
    for (long long r = r0 + threadIdx.y; r < r1; r += 8) acc += to_f32(dy[r * C + c]);
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += sh[j][threadIdx.x];
    xs::add(xsc.limbs, C, c, s);
  }
  // exact accumulation + decode by the last block: the gradient does not depend on block order
  if (xs::block_is_last(xsc.counter, gridDim.x * gridDim.y)) {
    const int tid = threadIdx.y * 32 + threadIdx.x;
    for (int i = tid; i < C; i += 256)
      db[i] = static_cast<float>(static_cast<double>(db[i]) + xs::read_clear(xsc.limbs, C, i));
    if (tid == 0) xs::release(xsc.counter);
  }
}

template <typename TS, typename TD>
__global__ void cast_kernel(const TS* __restrict__ s, TD* __restrict__ d, long long n) {
  MCN_PDL_PROLOGUE();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    d[i] = from_f32<TD>(to_f32(s[i]));
}
// Network input prologue (reference convnet.py:449-471): images arrive as fp32 in [0,1] or as raw
// uint8 (then /255 first), [N, Hi, Wi, C]; centre crop to [N, H, W, C] with offsets (Hi-H)//2,
// (Wi-W)//2 (convnet.py:1137-1149; a no-op when the sizes agree), zero-centre, scale, cast.
template <typename TS, typename TD>
__global__ void input_prep_kernel(const TS* __restrict__ x, int N, int Hi, int Wi, int H, int W, int C,
                                  float mean, float scale, TD* __restrict__ y) {
  MCN_PDL_PROLOGUE();
  const int oh = (Hi - H) / 2, ow = (Wi - W) / 2;
  const long long row = (long long)W * C;
  const long long total = (long long)N * H * row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / row;              // n*H + h
    const long long e = i - r * row;
    const int h = static_cast<int>(r % H);
    const long long n = r / H;
    const long long src = ((n * Hi + h + oh) * Wi + ow) * C + e;
    float v;
    if (sizeof(TS) == 1) v = static_cast<float>(x[src]) * (1.f / 255.f);
    else v = static_cast<float>(x[src]);
    y[i] = from_f32<TD>((v - mean) * scale);
  }
}

template <typename T>
__global__ void copy_channels_kernel(const T* __restrict__ src, long long rows, int Csrc,
                                     int src_off, T* __restrict__ dst, int Cdst, int dst_off,
                                     int Ccopy, int accumulate) {
  MCN_PDL_PROLOGUE();
  const long long total = rows * Ccopy;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % Ccopy);
    long long r = i / Ccopy;
    T v = src[r * Csrc + src_off + c];
    T* d = dst + r * Cdst + dst_off + c;
    *d = accumulate ? from_f32<T>(to_f32(*d) + to_f32(v)) : v;
  }
}

// Bilinear resize source coordinate (TF semantics).
__device__ __forceinline__ float src_coord(int dst, int in, int out, int mode) {
  if (mode == 1) return out > 1 ? dst * (float)(in - 1) / (float)(out - 1) : 0.f;
  if (mode == 2) {
    float s = ((float)dst + 0.5f) * ((float)in / (float)out) - 0.5f;
    return s;
  }
  return dst * ((float)in / (float)out);
}
__device__ __forceinline__ void lerp_idx(int dst, int in, int out, int mode, int* lo, int* hi,
                                         float* frac) {
  float s = src_coord(dst, in, out, mode);
  float fl = floorf(s);
  *lo = max((int)fl, 0);
  *hi = min((int)ceilf(s), in - 1);
  if (mode == 2) {
    // half-pixel: tf clamps the indices, keeps the fractional part of the unclamped coordinate
    *lo = min(max((int)fl, 0), in - 1);
    *hi = min(max((int)fl + 1, 0), in - 1);
  }
  *frac = s - fl;
}
template <typename T>
__global__ void resize_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C, int Ho,
                                  int Wo, int mode, T* __restrict__ y) {
  MCN_PDL_PROLOGUE();
  const long long total = (long long)N * Ho * Wo * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long r = i / C;
    int q = (int)(r % Wo);
    r /= Wo;
    int p = (int)(r % Ho);
    int n = (int)(r / Ho);
    int h0, h1, w0, w1;
    float fh, fw;
    lerp_idx(p, H, Ho, mode, &h0, &h1, &fh);
    lerp_idx(q, W, Wo, mode, &w0, &w1, &fw);
    const T* b = x + (long long)n * H * W * C + c;
    float tl = to_f32(b[((long long)h0 * W + w0) * C]), tr = to_f32(b[((long long)h0 * W + w1) * C]);
    float bl = to_f32(b[((long long)h1 * W + w0) * C]), br = to_f32(b[((long long)h1 * W + w1) * C]);
    float top = tl + (tr - tl) * fw, bot = bl + (br - bl) * fw;
    y[i] = from_f32<T>(top + (bot - top) * fh);
  }
}
// backward scatters with fp32 atomics into an fp32 scratch-free path: dx must be fp32-zeroed by
// the caller when T is float; for bf16 we gather instead (deterministic): each input pixel scans
// the output rows/cols that can touch it.
template <typename T>
__global__ void resize_bwd_kernel(const T* __restrict__ dy, int N, int H, int W, int C, int Ho,
                                  int Wo, int mode, T* __restrict__ dx) {
  MCN_PDL_PROLOGUE();
  const long long total = (long long)N * H * W * C;
  // conservative footprint of one input pixel in output space
  const int rh = (int)ceilf((float)Ho / (float)max(H - (mode == 1 ? 1 : 0), 1)) + 1;
  const int rw = (int)ceilf((float)Wo / (float)max(W - (mode == 1 ? 1 : 0), 1)) + 1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long r = i / C;
    int w = (int)(r % W);
    r /= W;
    int h = (int)(r % H);
    int n = (int)(r / H);
    float sh_ = (mode == 1) ? (H > 1 ? (float)(Ho - 1) / (float)(H - 1) : 0.f) : (float)Ho / (float)H;
    float sw_ = (mode == 1) ? (W > 1 ? (float)(Wo - 1) / (float)(W - 1) : 0.f) : (float)Wo / (float)W;
    int pc = (int)(h * sh_), qc = (int)(w * sw_);
    float acc = 0.f;
    for (int p = max(pc - rh, 0); p <= min(pc + rh, Ho - 1); ++p) {
      int h0, h1;
      float fh;
      lerp_idx(p, H, Ho, mode, &h0, &h1, &fh);
      float wh = (h0 == h ? (1.f - fh) : 0.f) + (h1 == h ? fh : 0.f);
      if (wh == 0.f) continue;
      for (int q = max(qc - rw, 0); q <= min(qc + rw, Wo - 1); ++q) {
        int w0, w1;
        float fw;
        lerp_idx(q, W, Wo, mode, &w0, &w1, &fw);
        float ww = (w0 == w ? (1.f - fw) : 0.f) + (w1 == w ? fw : 0.f);
        if (ww == 0.f) continue;
        acc = fmaf(wh * ww, to_f32(dy[(((long long)n * Ho + p) * Wo + q) * C + c]), acc);
      }
    }
    dx[i] = from_f32<T>(acc);
  }
}

// Table-driven backward (what the training step uses when the tables fit): the weight with which output
// row p reaches input row h depends on (p, h) only, so every block first builds, in shared memory, for
// each input row / column the first contributing output index and the weights of the kFoot outputs
// from there (zero outside the footprint), then each element is a kFoot x kFoot weighted gather —
// no coordinate arithmetic in the inner loops (the kernel above evaluates two lerp_idx per candidate
// pair: ~2k instructions per element at 4x upsampling; DeepLab: 1.1 + 0.9 ms -> see profiles).
// The summation order (p ascending, q ascending) is that of the kernel above.  V channels per thread.
template <typename T, int V>
__global__ void __launch_bounds__(256)
resize_bwd_table_kernel(const T* __restrict__ dy, int N, int H, int W, int C, int Ho, int Wo, int mode,
                        int foot, T* __restrict__ dx) {
  MCN_PDL_PROLOGUE();
  extern __shared__ float tab[];
  // per input index: [first output index (as int bits)] [foot weights]
  const int pitch = foot + 1;
  float* th = tab;
  float* tw = tab + (size_t)H * pitch;
  for (int which = 0; which < 2; ++which) {
    const int in = which ? W : H, out = which ? Wo : Ho;
    float* t = which ? tw : th;
    const float sc = (mode == 1) ? (in > 1 ? (float)(out - 1) / (float)(in - 1) : 0.f) : (float)out / (float)in;
    const int reach = (int)ceilf((float)out / (float)max(in - (mode == 1 ? 1 : 0), 1)) + 1;
    for (int i = threadIdx.x; i < in; i += blockDim.x) {
      const int centre = (int)(i * sc);
      int first = -1, n = 0;
      for (int p = max(centre - reach, 0); p <= min(centre + reach, out - 1); ++p) {
        int lo, hi;
        float fr;
        lerp_idx(p, in, out, mode, &lo, &hi, &fr);
        const float wgt = (lo == i ? (1.f - fr) : 0.f) + (hi == i ? fr : 0.f);
        if (first < 0 && wgt == 0.f) continue;
        if (first < 0) first = p;
        if (n < foot) t[i * pitch + 1 + n] = wgt;
        ++n;
      }
      for (int k = max(n, 0); k < foot; ++k) t[i * pitch + 1 + k] = 0.f;
      t[i * pitch] = __int_as_float(max(first, 0));
    }
  }
  __syncthreads();
  const int cv = C / V;
  const long long total = (long long)N * H * W * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * V;
    long long r = i / cv;
    const int w = (int)(r % W);
    r /= W;
    const int h = (int)(r % H);
    const int n = (int)(r / H);
    const float* ph = th + h * pitch;
    const float* pw = tw + w * pitch;
    const int p0 = __float_as_int(ph[0]), q0 = __float_as_int(pw[0]);
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    const T* img = dy + (long long)n * Ho * Wo * C + c0;
    for (int kp = 0; kp < foot; ++kp) {
      const float wh = ph[1 + kp];
      const int p = p0 + kp;
      if (wh == 0.f || p >= Ho) continue;
      const T* rowp = img + (long long)p * Wo * C;
      for (int kq = 0; kq < foot; ++kq) {
        const float ww = pw[1 + kq];
        const int q = q0 + kq;
        if (ww == 0.f || q >= Wo) continue;
        const float wgt = wh * ww;
        if (V == 1) {
          acc[0] = fmaf(wgt, to_f32(rowp[(long long)q * C]), acc[0]);
        } else {
          const Vec16<T> v = ld_vec(rowp + (long long)q * C);
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] = fmaf(wgt, v.get(e), acc[e]);
        }
      }
    }
    T* o = dx + (((long long)n * H + h) * W + w) * C + c0;
    if (V == 1) {
      o[0] = from_f32<T>(acc[0]);
    } else {
      Vec16<T> ov;
#pragma unroll
      for (int e = 0; e < V; ++e) ov.set(e, acc[e]);
      st_vec(o, ov);
    }
  }
}

// forward, 16-byte channel vectors (C % V == 0): one coordinate computation per V channels
template <typename T>
__global__ void __launch_bounds__(256)
resize_fwd_vec_kernel(const T* __restrict__ x, int N, int H, int W, int C, int Ho, int Wo, int mode,
                      T* __restrict__ y) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  const int cv = C / V;
  const long long total = (long long)N * Ho * Wo * cv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * V;
    long long r = i / cv;
    const int q = (int)(r % Wo);
    r /= Wo;
    const int p = (int)(r % Ho);
    const int n = (int)(r / Ho);
    int h0, h1, w0, w1;
    float fh, fw;
    lerp_idx(p, H, Ho, mode, &h0, &h1, &fh);
    lerp_idx(q, W, Wo, mode, &w0, &w1, &fw);
    const T* b = x + (long long)n * H * W * C + c0;
    const Vec16<T> tl = ld_vec(b + ((long long)h0 * W + w0) * C), tr = ld_vec(b + ((long long)h0 * W + w1) * C);
    const Vec16<T> bl = ld_vec(b + ((long long)h1 * W + w0) * C), br = ld_vec(b + ((long long)h1 * W + w1) * C);
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const float top = tl.get(e) + (tr.get(e) - tl.get(e)) * fw, bot = bl.get(e) + (br.get(e) - bl.get(e)) * fw;
      o.set(e, top + (bot - top) * fh);
    }
    st_vec(y + i * V, o);
  }
}

__global__ void fill_kernel(float* p, long long n, float v) {
  MCN_PDL_PROLOGUE();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}
__global__ void scale_kernel(float* p, long long n, float s) {
  MCN_PDL_PROLOGUE();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    p[i] *= s;
}

}  // namespace
}  // namespace mcn

using namespace mcn;

extern "C" int mcn_act_fwd(int dtype, const void* x, long long n, int act, float alpha, void* y,
                           void* stream) {
  MCN_REQUIRE(x && y && n >= 0, "act_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, return (launch_map<T, false>(x, nullptr, n, y, st, ActF{act, alpha}, "act_fwd")));
  return MCN_OK;
}
extern "C" int mcn_act_bwd(int dtype, const void* dy, const void* x, long long n, int act,
                           float alpha, void* dx, void* stream) {
  MCN_REQUIRE(dy && x && dx, "act_bwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, return (launch_map<T, true>(dy, x, n, dx, st, ActB{act, alpha}, "act_bwd")));
  return MCN_OK;
}
extern "C" int mcn_add_act_fwd(int dtype, const void* a, const void* b, long long n, int act,
                               float alpha, void* y, void* stream) {
  MCN_REQUIRE(a && b && y, "add_act_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, return (launch_map<T, true>(a, b, n, y, st, AddActF{act, alpha}, "add_act_fwd")));
  return MCN_OK;
}
extern "C" int mcn_add_act_bwd(int dtype, const void* dy, const void* y, long long n, int act,
                               float alpha, void* dz, void* stream) {
  MCN_REQUIRE(dy && y && dz, "add_act_bwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, return (launch_map<T, true>(dy, y, n, dz, st, AddActB{act, alpha}, "add_act_bwd")));
  return MCN_OK;
}
extern "C" int mcn_accumulate(int dtype, void* a, const void* b, long long n, void* stream) {
  MCN_REQUIRE(a && b, "accumulate: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, return (launch_map<T, true>(a, b, n, a, st, AddF{}, "accumulate")));
  return MCN_OK;
}

extern "C" int mcn_scale_bcast_fwd(int dtype, const void* x, const void* m, int N, int HW, int C,
                                   void* y, void* stream) {
  MCN_REQUIRE(x && m && y, "scale_bcast_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long total = (long long)N * HW * C;
  MCN_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec16<T>::N;
    const bool vec = C % V == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(m) % 16 == 0;
    if (vec)
      ::mcn::launch(scale_bcast_fwd_vec_kernel<T>, grid_for((total / V + 3) / 4, 256), 256, 0, st,
                    static_cast<const T*>(x), static_cast<const T*>(m), HW, C / V, total / V, static_cast<T*>(y));
    else
      ::mcn::launch(scale_bcast_fwd_kernel<T>, grid_for(total, 256), 256, 0, st, 
          static_cast<const T*>(x), static_cast<const T*>(m), HW, C, total, static_cast<T*>(y));
  });
  return after_launch("scale_bcast_fwd");
}
extern "C" int mcn_scale_bcast_bwd(int dtype, const void* dy, const void* x, const void* m, int N,
                                   int HW, int C, void* dx, float* dm, void* stream) {
  MCN_REQUIRE(dy && x && m && dx && dm, "scale_bcast_bwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid((C + 31) / 32, N), block(32, 8);
  MCN_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec16<T>::N;
    constexpr int kSlab = 4;
    const bool vec = C % V == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(dy) % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(dx) % 16 == 0 && reinterpret_cast<uintptr_t>(m) % 16 == 0;
    if (vec) {
      const int cv = C / V;
      dim3 vgrid((cv + kSlab - 1) / kSlab, N);
      ::mcn::launch(scale_bcast_bwd_vec_kernel<T, kSlab>, vgrid, 256, 0, st, static_cast<const T*>(dy),
                    static_cast<const T*>(x), static_cast<const T*>(m), HW, cv, static_cast<T*>(dx), dm);
    } else {
      ::mcn::launch(scale_bcast_bwd_kernel<T>, grid, block, 0, st, 
          static_cast<const T*>(dy), static_cast<const T*>(x), static_cast<const T*>(m), HW, C,
          static_cast<T*>(dx), dm);
    }
  });
  return after_launch("scale_bcast_bwd");
}

extern "C" int mcn_bias_add(int dtype, void* y, long long rows, int C, const float* bias,
                            void* stream) {
  MCN_REQUIRE(y && bias, "bias_add: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long total = rows * C;
  MCN_DISPATCH_DTYPE(dtype, T, {
    ::mcn::launch(bias_add_kernel<T>, grid_for(total, 256), 256, 0, st, static_cast<T*>(y), total, C, bias);
  });
  return after_launch("bias_add");
}
extern "C" int mcn_bias_grad(int dtype, const void* dy, long long rows, int C, float* db,
                             void* stream) {
  MCN_REQUIRE(dy && db, "bias_grad: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const XsScratch xsc = xs_scratch(C, "bias_grad");
  if (xsc.limbs == nullptr) return MCN_EINVAL;
  int chunks = (int)std::max<long long>(1, std::min<long long>(rows / 64, 2LL * num_sms()));
  dim3 grid((C + 31) / 32, chunks), block(32, 8);
  MCN_DISPATCH_DTYPE(dtype, T, {
    ::mcn::launch(bias_grad_kernel<T>, grid, block, 0, st, static_cast<const T*>(dy), rows, C, db, xsc);
  });
  return after_launch("bias_grad");
}

extern "C" int mcn_cast(int src_dtype, const void* src, int dst_dtype, void* dst, long long n,
                        void* stream) {
  MCN_REQUIRE(src && dst, "cast: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int g = grid_for(n, 256);
  if (src_dtype == MCN_F32 && dst_dtype == MCN_BF16)
    ::mcn::launch(cast_kernel<float, __nv_bfloat16>, g, 256, 0, st, (const float*)src, (__nv_bfloat16*)dst, n);
  else if (src_dtype == MCN_BF16 && dst_dtype == MCN_F32)
    ::mcn::launch(cast_kernel<__nv_bfloat16, float>, g, 256, 0, st, (const __nv_bfloat16*)src, (float*)dst, n);
  else if (src_dtype == MCN_F32 && dst_dtype == MCN_F32)
    ::mcn::launch(cast_kernel<float, float>, g, 256, 0, st, (const float*)src, (float*)dst, n);
  else if (src_dtype == MCN_BF16 && dst_dtype == MCN_BF16)
    ::mcn::launch(cast_kernel<__nv_bfloat16, __nv_bfloat16>, g, 256, 0, st, (const __nv_bfloat16*)src,
                                                                 (__nv_bfloat16*)dst, n);
  else {
    set_error("cast: unsupported dtypes %d -> %d", src_dtype, dst_dtype);
    return MCN_EINVAL;
  }
  return after_launch("cast");
}

// No crop: a flat element-wise map, 16 input elements per thread (one 16-byte load of uint8).
template <typename TD>
__global__ void input_prep_u8_flat_kernel(const uint4* __restrict__ x, long long n16, float mean, float scale,
                                          TD* __restrict__ y) {
  MCN_PDL_PROLOGUE();
  const float k = scale * (1.f / 255.f), b = -mean * scale;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n16;
       i += (long long)gridDim.x * blockDim.x) {
    const uint4 v = x[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    Vec16<TD> o[16 / Vec16<TD>::N];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int idx = j * 4 + e;
        o[idx / Vec16<TD>::N].set(idx % Vec16<TD>::N,
                                  fmaf(static_cast<float>((w[j] >> (8 * e)) & 255u), k, b));
      }
#pragma unroll
    for (int q = 0; q < 16 / Vec16<TD>::N; ++q) st_vec(y + i * 16 + q * Vec16<TD>::N, o[q]);
  }
}

extern "C" int mcn_input_prep(const void* x, int src_dtype, int N, int Hi, int Wi, int H, int W, int C,
                              float mean, float scale, int dst_dtype, void* y, void* stream) {
  MCN_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && Hi >= H && Wi >= W, "input_prep: bad argument");
  MCN_REQUIRE(src_dtype == MCN_F32 || src_dtype == MCN_U8, "input_prep: images are fp32 or uint8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)N * H * W * C;
  const int g = grid_for(n, 256);
  if (src_dtype == MCN_F32) {
    if (dst_dtype == MCN_F32)
      ::mcn::launch(input_prep_kernel<float, float>, g, 256, 0, st, (const float*)x, N, Hi, Wi, H, W, C, mean, scale, (float*)y);
    else
      ::mcn::launch(input_prep_kernel<float, __nv_bfloat16>, g, 256, 0, st, (const float*)x, N, Hi, Wi, H, W, C, mean, scale,
                                                               (__nv_bfloat16*)y);
  } else if (Hi == H && Wi == W && n % 16 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
             reinterpret_cast<uintptr_t>(y) % 16 == 0 && dst_dtype == MCN_BF16) {
    ::mcn::launch(input_prep_u8_flat_kernel<__nv_bfloat16>, grid_for(n / 16, 256), 256, 0, st, 
        (const uint4*)x, n / 16, mean, scale, (__nv_bfloat16*)y);
  } else {
    if (dst_dtype == MCN_F32)
      ::mcn::launch(input_prep_kernel<uint8_t, float>, g, 256, 0, st, (const uint8_t*)x, N, Hi, Wi, H, W, C, mean, scale,
                                                         (float*)y);
    else
      ::mcn::launch(input_prep_kernel<uint8_t, __nv_bfloat16>, g, 256, 0, st, (const uint8_t*)x, N, Hi, Wi, H, W, C, mean,
                                                                 scale, (__nv_bfloat16*)y);
  }
  return after_launch("input_prep");
}

// RGB -> 4-channel pixels (the stem convolution's input format): 4 pixels per thread, 24 bytes in
// (three aligned 8-byte loads), 32 bytes out; the 4th channel is written as zero.
__global__ void pad3to4_kernel(const uint2* __restrict__ x, long long quads, uint4* __restrict__ y) {
  MCN_PDL_PROLOGUE();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < quads;
       i += (long long)gridDim.x * blockDim.x) {
    const uint2 a = x[3 * i], b = x[3 * i + 1], c = x[3 * i + 2];
    // elements (16-bit): a.x = p0c0 p0c1 | a.y = p0c2 p1c0 | b.x = p1c1 p1c2 | b.y = p2c0 p2c1 | c.x = p2c2 p3c0 | c.y = p3c1 p3c2
    uint4 o0, o1;
    o0.x = a.x;
    o0.y = a.y & 0xFFFFu;
    o0.z = (a.y >> 16) | (b.x << 16);
    o0.w = b.x >> 16;
    o1.x = b.y;
    o1.y = c.x & 0xFFFFu;
    o1.z = (c.x >> 16) | (c.y << 16);
    o1.w = c.y >> 16;
    y[2 * i] = o0;
    y[2 * i + 1] = o1;
  }
}
__global__ void pad3to4_tail_kernel(const __nv_bfloat16* __restrict__ x, long long p0, long long p1,
                                    __nv_bfloat16* __restrict__ y) {
  MCN_PDL_PROLOGUE();
  for (long long p = p0 + blockIdx.x * blockDim.x + threadIdx.x; p < p1; p += (long long)gridDim.x * blockDim.x) {
    y[4 * p] = x[3 * p];
    y[4 * p + 1] = x[3 * p + 1];
    y[4 * p + 2] = x[3 * p + 2];
    y[4 * p + 3] = __float2bfloat16_rn(0.f);
  }
}

extern "C" int mcn_pad_rgb4(const void* x_bf16, long long pixels, void* y_bf16, void* stream) {
  MCN_REQUIRE(x_bf16 && y_bf16 && pixels >= 0, "pad_rgb4: bad argument");
  MCN_REQUIRE(reinterpret_cast<uintptr_t>(x_bf16) % 8 == 0 && reinterpret_cast<uintptr_t>(y_bf16) % 16 == 0,
              "pad_rgb4: misaligned tensor");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long quads = pixels / 4;
  if (quads > 0)
    ::mcn::launch(pad3to4_kernel, grid_for(quads, 256), 256, 0, st, static_cast<const uint2*>(x_bf16), quads,
                                                         static_cast<uint4*>(y_bf16));
  if (quads * 4 < pixels)
    ::mcn::launch(pad3to4_tail_kernel, 1, 32, 0, st, static_cast<const __nv_bfloat16*>(x_bf16), quads * 4, pixels,
                                          static_cast<__nv_bfloat16*>(y_bf16));
  return after_launch("pad_rgb4");
}

extern "C" int mcn_copy_channels(int dtype, const void* src, long long rows, int Csrc, int src_off,
                                 void* dst, int Cdst, int dst_off, int Ccopy, int accumulate,
                                 void* stream) {
  MCN_REQUIRE(src && dst && src_off + Ccopy <= Csrc && dst_off + Ccopy <= Cdst,
              "copy_channels: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    ::mcn::launch(copy_channels_kernel<T>, grid_for(rows * Ccopy, 256), 256, 0, st, 
        static_cast<const T*>(src), rows, Csrc, src_off, static_cast<T*>(dst), Cdst, dst_off, Ccopy,
        accumulate);
  });
  return after_launch("copy_channels");
}

extern "C" int mcn_resize_bilinear_fwd(int dtype, const void* x, int N, int H, int W, int C,
                                       int Ho, int Wo, int mode, void* y, void* stream) {
  MCN_REQUIRE(x && y && mode >= 0 && mode <= 2, "resize_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long total = (long long)N * Ho * Wo * C;
  MCN_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec16<T>::N;
    if (C % V == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0)
      ::mcn::launch(resize_fwd_vec_kernel<T>, grid_for(total / V, 256), 256, 0, st, static_cast<const T*>(x), N, H,
                    W, C, Ho, Wo, mode, static_cast<T*>(y));
    else
      ::mcn::launch(resize_fwd_kernel<T>, grid_for(total, 256), 256, 0, st, static_cast<const T*>(x), N, H, W, C,
                                                              Ho, Wo, mode, static_cast<T*>(y));
  });
  return after_launch("resize_fwd");
}
extern "C" int mcn_resize_bilinear_bwd(int dtype, const void* dy, int N, int H, int W, int C,
                                       int Ho, int Wo, int mode, void* dx, void* stream) {
  MCN_REQUIRE(dy && dx && mode >= 0 && mode <= 2, "resize_bwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  long long total = (long long)N * H * W * C;
  // footprint (outputs that can reach one input index, per axis) of the table-driven kernel
  auto reach = [&](int in, int out) {
    return 2 * ((int)std::ceil((float)out / (float)std::max(in - (mode == 1 ? 1 : 0), 1)) + 1) + 1;
  };
  const int foot = std::max(reach(H, Ho), reach(W, Wo));
  const size_t tab_bytes = (size_t)(H + W) * (foot + 1) * sizeof(float);
  const char* env_tab = getenv("MCN_RESIZE_TABLE");      // 0: the coordinate-per-candidate kernel (A/B, tests)
  const bool use_tab = !(env_tab && env_tab[0] == '0') && tab_bytes <= 40 * 1024;
  MCN_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec16<T>::N;
    if (use_tab && C % V == 0 && reinterpret_cast<uintptr_t>(dy) % 16 == 0 && reinterpret_cast<uintptr_t>(dx) % 16 == 0)
      ::mcn::launch(resize_bwd_table_kernel<T, V>, grid_for(total / V, 256), 256, tab_bytes, st,
                    static_cast<const T*>(dy), N, H, W, C, Ho, Wo, mode, foot, static_cast<T*>(dx));
    else if (use_tab)
      ::mcn::launch(resize_bwd_table_kernel<T, 1>, grid_for(total, 256), 256, tab_bytes, st,
                    static_cast<const T*>(dy), N, H, W, C, Ho, Wo, mode, foot, static_cast<T*>(dx));
    else
      ::mcn::launch(resize_bwd_kernel<T>, grid_for(total, 256), 256, 0, st, static_cast<const T*>(dy), N, H, W,
                                                                C, Ho, Wo, mode, static_cast<T*>(dx));
  });
  return after_launch("resize_bwd");
}

extern "C" int mcn_fill_f32(float* p, long long n, float v, void* stream) {
  MCN_REQUIRE(p, "fill: null");
  ::mcn::launch(fill_kernel, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), p, n, v);
  return after_launch("fill");
}
extern "C" int mcn_scale_f32(float* p, long long n, float s, void* stream) {
  MCN_REQUIRE(p, "scale: null");
  ::mcn::launch(scale_kernel, grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream), p, n, s);
  return after_launch("scale");
}
