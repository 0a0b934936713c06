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
  MCN_PDL_PROLOGUE();
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
  MCN_PDL_PROLOGUE();
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
  MCN_PDL_PROLOGUE();
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

// ---- compact-argmax max pooling (what the training step uses).
// The int32 TF-style argmax above costs 4 bytes per output element — twice the bf16 tensor itself
// (ResNet-50 stem pool at batch 256: 205 MB of 719 MB).  The winning position is one of kh*kw <= 255
// taps of the window, so the step stores the TAP index a*kw + b as one byte (51 MB); the TF index is
// (p*sh + a - pad_t)*W + (q*sw + b - pad_l))*C + c, recoverable exactly (maxpool_tap_to_argmax).
// K / S > 0 are compile-time square kernel / stride (3/2, 2/2, 3/1: every window load is issued
// before the first compare); K == 0 is the run-time general case.
template <typename T, int V, int K, int S>
__global__ void __launch_bounds__(256)
maxpool_fwd_tap_kernel(const T* __restrict__ x, int N, int H, int W, int C, int kh_, int kw_, int sh_, int sw_,
                       int pad_t, int pad_l, int Ho, int Wo, T* __restrict__ y, uint8_t* __restrict__ tap) {
  MCN_PDL_PROLOGUE();
  const int kh = K ? K : kh_, kw = K ? K : kw_, sh = S ? S : sh_, sw = S ? S : sw_;
  const int cv = C / V;
  const uint32_t total = (uint32_t)N * Ho * Wo * cv;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c0 = (int)(i % (uint32_t)cv) * V;
    uint32_t r = i / (uint32_t)cv;
    const int q = (int)(r % (uint32_t)Wo);
    r /= (uint32_t)Wo;
    const int p = (int)(r % (uint32_t)Ho);
    const int n = (int)(r / (uint32_t)Ho);
    const int h0 = p * sh - pad_t, w0 = q * sw - pad_l;
    const T* img = x + (long long)n * H * W * C + c0;
    float best[V];
    int bt[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      best[e] = -FLT_MAX;
      bt[e] = -1;
    }
    if (K > 0) {
      constexpr int KK = K ? K * K : 1;
      Vec16<T> t[KK];
      bool ok[KK];
#pragma unroll
      for (int a = 0; a < K; ++a)
#pragma unroll
        for (int b = 0; b < K; ++b) {
          const int h = h0 + a, w = w0 + b;
          ok[a * K + b] = h >= 0 && h < H && w >= 0 && w < W;
          if (ok[a * K + b]) t[a * K + b] = ld_vec_stream(img + ((long long)h * W + w) * C);
        }
#pragma unroll
      for (int k = 0; k < K * K; ++k)
        if (ok[k]) {
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const float v = t[k].get(e);
            if (v > best[e] || bt[e] < 0) {
              best[e] = v;
              bt[e] = k;
            }
          }
        }
    } else {
      for (int a = 0; a < kh; ++a) {
        const int h = h0 + a;
        if (h < 0 || h >= H) continue;
        for (int b = 0; b < kw; ++b) {
          const int w = w0 + b;
          if (w < 0 || w >= W) continue;
          const Vec16<T> t = ld_vec(img + ((long long)h * W + w) * C);
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const float v = t.get(e);
            if (v > best[e] || bt[e] < 0) {
              best[e] = v;
              bt[e] = a * kw + b;
            }
          }
        }
      }
    }
    const long long o = (((long long)n * Ho + p) * Wo + q) * C + c0;
    Vec16<T> ov;
#pragma unroll
    for (int e = 0; e < V; ++e) ov.set(e, best[e]);
    st_vec(y + o, ov);
    uint32_t pk[V / 4];
#pragma unroll
    for (int e = 0; e < V; e += 4)
      pk[e / 4] = (uint32_t)(bt[e] & 255) | ((uint32_t)(bt[e + 1] & 255) << 8) |
                  ((uint32_t)(bt[e + 2] & 255) << 16) | ((uint32_t)(bt[e + 3] & 255) << 24);
    if (V == 8) *reinterpret_cast<uint2*>(tap + o) = make_uint2(pk[0], pk[1]);
    else *reinterpret_cast<uint32_t*>(tap + o) = pk[0];
  }
}

// Backward, gather form: an input pixel (h, w) lies in window p at tap row a = h + pad_t - p*sh;
// it receives dy[p, q] when the stored tap equals (a, b).  One thread per input pixel x 16-byte
// channel vector; with K, S known every candidate window's dy / tap load is issued first.
template <typename T, int V, int K, int S>
__global__ void __launch_bounds__(256)
maxpool_bwd_tap_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ tap, int N, int H, int W, int C,
                       int kh_, int kw_, int sh_, int sw_, int pad_t, int pad_l, int Ho, int Wo,
                       T* __restrict__ dx) {
  MCN_PDL_PROLOGUE();
  const int kh = K ? K : kh_, kw = K ? K : kw_, sh = S ? S : sh_, sw = S ? S : sw_;
  constexpr int MAXW = K ? (K + S - 1) / S : 1;      // windows covering a pixel along one axis
  const int cv = C / V;
  const uint32_t total = (uint32_t)N * H * W * cv;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c0 = (int)(i % (uint32_t)cv) * V;
    uint32_t r = i / (uint32_t)cv;
    const int w = (int)(r % (uint32_t)W);
    r /= (uint32_t)W;
    const int h = (int)(r % (uint32_t)H);
    const int n = (int)(r / (uint32_t)H);
    const int p_hi = (h + pad_t) / sh, q_hi = (w + pad_l) / sw;      // window with the smallest tap row/col
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    const T* gimg = dy + (long long)n * Ho * Wo * C + c0;
    const uint8_t* timg = tap + (long long)n * Ho * Wo * C + c0;
    if (K > 0) {
      Vec16<T> g[MAXW * MAXW];
      uint2 tp[MAXW * MAXW];
      int want[MAXW * MAXW];
#pragma unroll
      for (int j = 0; j < MAXW; ++j)
#pragma unroll
        for (int k = 0; k < MAXW; ++k) {
          const int p = p_hi - j, q = q_hi - k;
          const int a = h + pad_t - p * S, b = w + pad_l - q * S;
          const bool ok = p >= 0 && p < Ho && q >= 0 && q < Wo && a < K && b < K;
          want[j * MAXW + k] = ok ? a * K + b : -1;
          if (ok) {
            const long long o = ((long long)p * Wo + q) * C;
            g[j * MAXW + k] = ld_vec_stream(gimg + o);
            if (V == 8) tp[j * MAXW + k] = *reinterpret_cast<const uint2*>(timg + o);
            else tp[j * MAXW + k] = make_uint2(*reinterpret_cast<const uint32_t*>(timg + o), 0u);
          }
        }
      // ascending (p, q): the summation order of the int32-argmax kernel, so both agree bit for bit
#pragma unroll
      for (int m = MAXW * MAXW - 1; m >= 0; --m)
        if (want[m] >= 0) {
#pragma unroll
          for (int e = 0; e < V; ++e) {
            const uint32_t word = e < 4 ? tp[m].x : tp[m].y;
            if ((int)((word >> (8 * (e & 3))) & 255u) == want[m]) acc[e] += g[m].get(e);
          }
        }
    } else {
      int p_lo = h + pad_t - kh + 1 + sh - 1;
      p_lo = p_lo <= 0 ? 0 : p_lo / sh;
      int q_lo = w + pad_l - kw + 1 + sw - 1;
      q_lo = q_lo <= 0 ? 0 : q_lo / sw;
      for (int p = p_lo; p <= min(p_hi, Ho - 1); ++p) {
        const int a = h + pad_t - p * sh;
        for (int q = q_lo; q <= min(q_hi, Wo - 1); ++q) {
          const int want = a * kw + (w + pad_l - q * sw);
          const long long o = ((long long)p * Wo + q) * C;
          const Vec16<T> g = ld_vec(gimg + o);
#pragma unroll
          for (int e = 0; e < V; ++e)
            if ((int)timg[o + e] == want) acc[e] += g.get(e);
        }
      }
    }
    Vec16<T> ov;
#pragma unroll
    for (int e = 0; e < V; ++e) ov.set(e, acc[e]);
    st_vec(dx + (((long long)n * H + h) * W + w) * C + c0, ov);
  }
}

// ---- 3x3 / stride 2 (the ResNet stem pool, 112^2 -> 56^2 of 411 MB at batch 256).
// The per-output kernels above fetch every window tap on their own with L1-bypassing loads: 9 vector
// loads per output of which 5 re-read a neighbour's data (2.25 L2 reads per input element, 925 MB in
// 185 us = the L2 -> SM streaming rate), and the per-input backward fetches 2.25 (dy, tap) pairs per
// element written.  Forward: 2-D output tiles per block with L1-allocating loads.  Backward: a thread
// owns one 2-pixel-wide input column pair of one 16-byte channel vector and walks down a strip of
// rows, carrying the window row shared with the next pixel pair in registers (293 -> 148 us).
// Tie-breaking and summation order are those of the kernels above (first maximum in row-major window
// order; ascending (p, q)): results are bit-identical.
template <typename T, int V>
__global__ void __launch_bounds__(256)
maxpool_fwd_tap_tile32_kernel(const T* __restrict__ x, int N, int H, int W, int C, int pad_t, int pad_l, int Ho,
                              int Wo, int TP, int tiles_p, int tiles_q, T* __restrict__ y,
                              uint8_t* __restrict__ tap) {
  MCN_PDL_PROLOGUE();
  // forward: a block owns a TP x 4 tile of outputs (256 threads = outputs x channel vectors) and loads
  // with L1 allocation: the taps shared by neighbouring windows hit L1 instead of crossing L2 again
  const int cv = C / V;
  const int c0 = (int)(threadIdx.x % (unsigned)cv) * V;
  const int o = (int)(threadIdx.x / (unsigned)cv);
  const int dq = o & 3, dp = o >> 2;
  const int total = N * tiles_p * tiles_q;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tq = tile % tiles_q;
    const int r = tile / tiles_q;
    const int tp = r % tiles_p, n = r / tiles_p;
    const int p = tp * TP + dp, q = tq * 4 + dq;
    if (p >= Ho || q >= Wo) continue;
    const int h0 = p * 2 - pad_t, w0 = q * 2 - pad_l;
    const T* img = x + (long long)n * H * W * C + c0;
    Vec16<T> t[9];
    bool ok[9];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int h = h0 + a, w = w0 + b;
        ok[a * 3 + b] = h >= 0 && h < H && w >= 0 && w < W;
        if (ok[a * 3 + b]) t[a * 3 + b] = ld_vec(img + ((long long)h * W + w) * C);
      }
    float best[V];
    int bt[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      best[e] = -FLT_MAX;
      bt[e] = -1;
    }
#pragma unroll
    for (int k = 0; k < 9; ++k)
      if (ok[k]) {
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float v = t[k].get(e);
          if (v > best[e] || bt[e] < 0) {
            best[e] = v;
            bt[e] = k;
          }
        }
      }
    const long long oo = (((long long)n * Ho + p) * Wo + q) * C + c0;
    Vec16<T> ov;
#pragma unroll
    for (int e = 0; e < V; ++e) ov.set(e, best[e]);
    st_vec(y + oo, ov);
    uint32_t pk[V / 4];
#pragma unroll
    for (int e = 0; e < V; e += 4)
      pk[e / 4] = (uint32_t)(bt[e] & 255) | ((uint32_t)(bt[e + 1] & 255) << 8) |
                  ((uint32_t)(bt[e + 2] & 255) << 16) | ((uint32_t)(bt[e + 3] & 255) << 24);
    if (V == 8) *reinterpret_cast<uint2*>(tap + oo) = make_uint2(pk[0], pk[1]);
    else *reinterpret_cast<uint32_t*>(tap + oo) = pk[0];
  }
}

// bf16 forward on packed pairs.  ncu on the kernel above: ALU pipe 80 % busy, DRAM 36 % — the per-channel
// unpack / compare / two selects (~500 instructions per output vector) bound it, not memory.  bf16 values
// compare exactly as packed pairs: one HSET2 mask + two LOP3 (value, tap id) per PAIR of channels and
// tap; outputs and tap bytes leave without any float conversion.  Same rule as above: the first valid
// tap initialises, later taps win only when strictly greater (padding acts as -inf).
__global__ void __launch_bounds__(256)
maxpool_fwd_tap_tile32_bf16_kernel(const __nv_bfloat16* __restrict__ x, int N, int H, int W, int C, int pad_t,
                                   int pad_l, int Ho, int Wo, int TP, int tiles_p, int tiles_q,
                                   __nv_bfloat16* __restrict__ y, uint8_t* __restrict__ tap) {
  MCN_PDL_PROLOGUE();
  const int cv = C / 8;
  const int c0 = (int)(threadIdx.x % (unsigned)cv) * 8;
  const int o = (int)(threadIdx.x / (unsigned)cv);
  const int dq = o & 3, dp = o >> 2;
  const int total = N * tiles_p * tiles_q;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int tq = tile % tiles_q;
    const int r = tile / tiles_q;
    const int tp = r % tiles_p, n = r / tiles_p;
    const int p = tp * TP + dp, q = tq * 4 + dq;
    if (p >= Ho || q >= Wo) continue;
    const int h0 = p * 2 - pad_t, w0 = q * 2 - pad_l;
    const __nv_bfloat16* img = x + (long long)n * H * W * C + c0;
    uint4 t[9];
    uint32_t first = 255u;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int h = h0 + a, w = w0 + b;
        const bool ok = h >= 0 && h < H && w >= 0 && w < W;
        if (ok) {
          t[a * 3 + b] = *reinterpret_cast<const uint4*>(img + ((long long)h * W + w) * C);
          if (first == 255u) first = (uint32_t)(a * 3 + b);
        } else {
          t[a * 3 + b] = make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);      // -inf
        }
      }
    uint32_t best[4], bt[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      best[j] = 0xFF80FF80u;
      bt[j] = first | (first << 16);
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const uint32_t w4[4] = {t[k].x, t[k].y, t[k].z, t[k].w};
      const uint32_t kk = (uint32_t)k | ((uint32_t)k << 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t m = __hgt2_mask(*reinterpret_cast<const __nv_bfloat162*>(&w4[j]),
                                       *reinterpret_cast<const __nv_bfloat162*>(&best[j]));
        best[j] = (w4[j] & m) | (best[j] & ~m);
        bt[j] = (kk & m) | (bt[j] & ~m);
      }
    }
    const long long oo = (((long long)n * Ho + p) * Wo + q) * C + c0;
    *reinterpret_cast<uint4*>(y + oo) = make_uint4(best[0], best[1], best[2], best[3]);
    // tap ids sit in the low byte of each 16-bit half
    *reinterpret_cast<uint2*>(tap + oo) = make_uint2(__byte_perm(bt[0], bt[1], 0x6420), __byte_perm(bt[2], bt[3], 0x6420));
  }
}

// Backward: with u = h + pad_t, v = w + pad_l the pixels (2j | 2j+1, 2k | 2k+1) are covered by the
// windows (j-1 | j, k-1 | k) only; a thread walks j down a strip, carrying window row j-1.
template <typename T, int V>
__device__ __forceinline__ void maxpool_take(const Vec16<T>& g, const uint2& tp, bool ok, int want, float* acc) {
  if (!ok) return;
#pragma unroll
  for (int e = 0; e < V; ++e) {
    const uint32_t word = e < 4 ? tp.x : tp.y;
    if ((int)((word >> (8 * (e & 3))) & 255u) == want) acc[e] += g.get(e);
  }
}

template <typename T, int V>
__global__ void __launch_bounds__(256)
maxpool_bwd_tap_strip32_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ tap, int N, int H, int W,
                               int C, int pad_t, int pad_l, int Ho, int Wo, int nj, int nk, int strip_rows,
                               int strips, T* __restrict__ dx) {
  MCN_PDL_PROLOGUE();
  const int cv = C / V;
  const uint32_t total = (uint32_t)N * strips * nk * cv;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c0 = (int)(i % (uint32_t)cv) * V;
    uint32_t r = i / (uint32_t)cv;
    const int k = (int)(r % (uint32_t)nk);
    r /= (uint32_t)nk;
    const int sidx = (int)(r % (uint32_t)strips);
    const int n = (int)(r / (uint32_t)strips);
    const int j0 = sidx * strip_rows, j1 = min(nj, j0 + strip_rows);
    const T* gimg = dy + (long long)n * Ho * Wo * C + c0;
    const uint8_t* timg = tap + (long long)n * Ho * Wo * C + c0;
    T* ximg = dx + (long long)n * H * W * C + c0;
    const bool okq[2] = {k - 1 >= 0 && k - 1 < Wo, k < Wo};      // window columns k-1, k
    const int w_a = 2 * k - pad_l, w_b = w_a + 1;                 // the two pixel columns
    const bool okwa = w_a >= 0 && w_a < W, okwb = w_b >= 0 && w_b < W;
    Vec16<T> gp[2], gc[2];
    uint2 tp[2], tc[2];
    bool okp[2], okc[2];
    auto load_row = [&](int p, Vec16<T>* g, uint2* t, bool* ok) {
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        ok[m] = p >= 0 && p < Ho && okq[m];
        if (ok[m]) {
          const long long o = ((long long)p * Wo + (k - 1 + m)) * C;
          g[m] = ld_vec(gimg + o);
          if (V == 8) t[m] = *reinterpret_cast<const uint2*>(timg + o);
          else t[m] = make_uint2(*reinterpret_cast<const uint32_t*>(timg + o), 0u);
        }
      }
    };
    load_row(j0 - 1, gp, tp, okp);
    for (int j = j0; j < j1; ++j) {
      load_row(j, gc, tc, okc);
      const int h_a = 2 * j - pad_t, h_b = h_a + 1;
      float acc[V];
      Vec16<T> ov;
      if (h_a >= 0 && h_a < H) {
        if (okwa) {     // (2j, 2k): windows (j-1,k-1) tap (2,2), (j-1,k) (2,0), (j,k-1) (0,2), (j,k) (0,0)
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] = 0.f;
          maxpool_take<T, V>(gp[0], tp[0], okp[0], 8, acc);
          maxpool_take<T, V>(gp[1], tp[1], okp[1], 6, acc);
          maxpool_take<T, V>(gc[0], tc[0], okc[0], 2, acc);
          maxpool_take<T, V>(gc[1], tc[1], okc[1], 0, acc);
#pragma unroll
          for (int e = 0; e < V; ++e) ov.set(e, acc[e]);
          st_vec(ximg + ((long long)h_a * W + w_a) * C, ov);
        }
        if (okwb) {     // (2j, 2k+1): windows (j-1,k) tap (2,1), (j,k) (0,1)
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] = 0.f;
          maxpool_take<T, V>(gp[1], tp[1], okp[1], 7, acc);
          maxpool_take<T, V>(gc[1], tc[1], okc[1], 1, acc);
#pragma unroll
          for (int e = 0; e < V; ++e) ov.set(e, acc[e]);
          st_vec(ximg + ((long long)h_a * W + w_b) * C, ov);
        }
      }
      if (h_b >= 0 && h_b < H) {
        if (okwa) {     // (2j+1, 2k): windows (j,k-1) tap (1,2), (j,k) (1,0)
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] = 0.f;
          maxpool_take<T, V>(gc[0], tc[0], okc[0], 5, acc);
          maxpool_take<T, V>(gc[1], tc[1], okc[1], 3, acc);
#pragma unroll
          for (int e = 0; e < V; ++e) ov.set(e, acc[e]);
          st_vec(ximg + ((long long)h_b * W + w_a) * C, ov);
        }
        if (okwb) {     // (2j+1, 2k+1): window (j,k) tap (1,1)
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] = 0.f;
          maxpool_take<T, V>(gc[1], tc[1], okc[1], 4, acc);
#pragma unroll
          for (int e = 0; e < V; ++e) ov.set(e, acc[e]);
          st_vec(ximg + ((long long)h_b * W + w_b) * C, ov);
        }
      }
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        gp[m] = gc[m];
        tp[m] = tc[m];
        okp[m] = okc[m];
      }
    }
  }
}

__global__ void maxpool_tap_to_argmax_kernel(const uint8_t* __restrict__ tap, long long total, int W, int C,
                                             int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo,
                                             int32_t* __restrict__ argmax) {
  MCN_PDL_PROLOGUE();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    long long r = i / C;
    const int q = (int)(r % Wo);
    r /= Wo;
    const int p = (int)(r % Ho);
    const int t = tap[i];
    argmax[i] = ((p * sh + t / kw - pad_t) * W + (q * sw + t % kw - pad_l)) * C + c;
  }
}

// Average pooling; SAME divides by the number of in-bounds elements.
template <typename T>
__global__ void avgpool_fwd_kernel(const T* __restrict__ x, int N, int H, int W, int C, int kh,
                                   int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo,
                                   T* __restrict__ y) {
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
  MCN_PDL_PROLOGUE();
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
  MCN_PDL_PROLOGUE();
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
  MCN_PDL_PROLOGUE();
  const float inv = 1.f / (float)HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    long long n = i / ((long long)HW * C);
    dx[i] = from_f32<T>(to_f32(dy[n * C + c]) * inv);
  }
}

// 16-byte stores: one thread writes V consecutive channels of one pixel
template <typename T, typename TI>
__global__ void gap_bwd_vec_kernel(const TI* __restrict__ dy, int HW, int C, long long nvec,
                                   T* __restrict__ dx) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  const float inv = 1.f / (float)HW;
  const int cv = C / V;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cv) * V;
    const long long n = i / ((long long)HW * cv);
    Vec16<T> o;
#pragma unroll
    for (int e = 0; e < V; ++e) o.set(e, to_f32(dy[n * C + c0 + e]) * inv);
    st_vec(dx + i * V, o);
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
      ::mcn::launch(maxpool_fwd_kernel<T, V>, grid_for(total, 256), 256, 0, st, 
          static_cast<const T*>(x), N, H, W, C, kh, kw, sh, sw, pad_t, pad_l, Ho, Wo,
          static_cast<T*>(y), argmax);
    } else {
      long long total = (long long)N * Ho * Wo * C;
      ::mcn::launch(maxpool_fwd_kernel<T, 1>, grid_for(total, 256), 256, 0, st, 
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
      ::mcn::launch(maxpool_bwd_vec_kernel<T, V>, grid_for(total, 256), 256, 0, st, 
          static_cast<const T*>(dy), argmax, N, H, W, C, kh, kw, sh, sw, pad_t, pad_l, Ho, Wo,
          static_cast<T*>(dx));
    } else {
      long long total = (long long)N * H * W * C;
      ::mcn::launch(maxpool_bwd_kernel<T>, grid_for(total, 256), 256, 0, st, 
          static_cast<const T*>(dy), argmax, N, H, W, C, kh, kw, sh, sw, pad_t, pad_l, Ho, Wo,
          static_cast<T*>(dx));
    }
  });
  return after_launch("maxpool_bwd");
}

template <typename T, int V>
static void launch_maxpool_tap(bool fwd, const void* in, const uint8_t* tap_in, int N, int H, int W, int C, int kh,
                               int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo, void* out,
                               uint8_t* tap_out, cudaStream_t st) {
  const long long total = fwd ? (long long)N * Ho * Wo * (C / V) : (long long)N * H * W * (C / V);
  // persistent grid-stride blocks: one short-lived block per 256 outputs (100 k blocks for the ResNet
  // stem's backward) spent its time in block scheduling; MCN_POOL_BLOCKS_PER_SM overrides (A/B)
  static int bps = 0;
  if (!bps) {
    const char* e = getenv("MCN_POOL_BLOCKS_PER_SM");
    bps = e ? std::max(1, atoi(e)) : 16;   // measured: 0.54 (one block per 256 outputs) / 0.50 (8) / 0.48 ms (16)
  }
  const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)bps * num_sms()));
#define MCN_POOL_CASE(KK, SS)                                                                        \
  if (fwd)                                                                                           \
    ::mcn::launch(maxpool_fwd_tap_kernel<T, V, KK, SS>, grid, 256, 0, st, static_cast<const T*>(in), N, H, W, C, kh, kw, sh, \
                                                             sw, pad_t, pad_l, Ho, Wo, static_cast<T*>(out), tap_out); \
  else                                                                                               \
    ::mcn::launch(maxpool_bwd_tap_kernel<T, V, KK, SS>, grid, 256, 0, st, static_cast<const T*>(in), tap_in, N, H, W, C, kh, \
                                                             kw, sh, sw, pad_t, pad_l, Ho, Wo, static_cast<T*>(out))
  const bool sq = kh == kw && sh == sw;
  static int strip = -1;
  if (strip < 0) {
    const char* e = getenv("MCN_POOL_STRIP");      // 0: the per-output kernels (A/B)
    strip = (e && e[0] == '0') ? 0 : 1;
  }
  if (strip && sq && kh == 3 && sh == 2 && pad_t >= 0 && pad_t <= 2 && pad_l >= 0 && pad_l <= 2) {
    const int cv = C / V;
    if (fwd) {
      if (256 % cv == 0 && 256 / cv >= 4) {
        const int TP = 256 / cv / 4, tiles_p = (Ho + TP - 1) / TP, tiles_q = (Wo + 3) / 4;
        const long long tiles = (long long)N * tiles_p * tiles_q;
        const int tgrid = (int)std::max<long long>(1, std::min<long long>(tiles, (long long)bps * num_sms()));
        static int packed = -1;
        if (packed < 0) {
          const char* e = getenv("MCN_POOL_PACKED");      // 0: the per-channel fp32 compare kernel (A/B)
          packed = (e && e[0] == '0') ? 0 : 1;
        }
        if (sizeof(T) == 2 && packed)
          ::mcn::launch(maxpool_fwd_tap_tile32_bf16_kernel, tgrid, 256, 0, st,
                        static_cast<const __nv_bfloat16*>(in), N, H, W, C, pad_t, pad_l, Ho, Wo, TP, tiles_p, tiles_q,
                        static_cast<__nv_bfloat16*>(out), tap_out);
        else
          ::mcn::launch(maxpool_fwd_tap_tile32_kernel<T, V>, tgrid, 256, 0, st, static_cast<const T*>(in), N, H, W, C,
                        pad_t, pad_l, Ho, Wo, TP, tiles_p, tiles_q, static_cast<T*>(out), tap_out);
        return;
      }
    } else {
      // strips of rows: enough threads to fill the chip several times over, strips as long as that allows
      const int rows = (H + pad_t + 1) / 2, cols = (W + pad_l + 1) / 2;
      const long long per_strip = (long long)N * cols * cv;
      const long long want = (long long)num_sms() * 2048 * 2;
      int strips = (int)std::max<long long>(1, std::min<long long>(rows, (want + per_strip - 1) / per_strip));
      int strip_rows = (rows + strips - 1) / strips;
      if (const char* e = getenv("MCN_POOL_STRIP_ROWS")) strip_rows = std::max(1, std::min(rows, atoi(e)));   // tests
      strips = (rows + strip_rows - 1) / strip_rows;
      const long long threads = per_strip * strips;
      const int sgrid = (int)std::max<long long>(1, std::min<long long>((threads + 255) / 256, (long long)bps * num_sms()));
      ::mcn::launch(maxpool_bwd_tap_strip32_kernel<T, V>, sgrid, 256, 0, st, static_cast<const T*>(in), tap_in, N, H,
                    W, C, pad_t, pad_l, Ho, Wo, rows, cols, strip_rows, strips, static_cast<T*>(out));
      return;
    }
  }
  if (sq && kh == 3 && sh == 2) { MCN_POOL_CASE(3, 2); }
  else if (sq && kh == 2 && sh == 2) { MCN_POOL_CASE(2, 2); }
  else if (sq && kh == 3 && sh == 1) { MCN_POOL_CASE(3, 1); }
  else { MCN_POOL_CASE(0, 0); }
#undef MCN_POOL_CASE
}

static int maxpool_tap_impl(bool fwd, int dtype, const void* in, const uint8_t* tap_in, int N, int H, int W, int C,
                            int kh, int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo, void* out,
                            uint8_t* tap_out, void* stream) {
  MCN_REQUIRE(kh * kw <= 255, "maxpool (tap form): window of %d x %d taps does not fit a byte", kh, kw);
  MCN_REQUIRE((long long)N * H * W * C < (1LL << 31) * 4 && (long long)N * H * W * (C / 4) < (1LL << 32) - (1 << 24),
              "maxpool (tap form): tensor too large for 32-bit vector indexing");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    constexpr int V = Vec16<T>::N;
    MCN_REQUIRE(C % V == 0, "maxpool (tap form): C=%d must be a multiple of %d", C, V);
    launch_maxpool_tap<T, V>(fwd, in, tap_in, N, H, W, C, kh, kw, sh, sw, pad_t, pad_l, Ho, Wo, out, tap_out, st);
  });
  return after_launch(fwd ? "maxpool_fwd_tap" : "maxpool_bwd_tap");
}

extern "C" int mcn_maxpool_fwd_tap(int dtype, const void* x, int N, int H, int W, int C, int kh, int kw, int sh,
                                   int sw, int pad_t, int pad_l, int Ho, int Wo, void* y, uint8_t* tap,
                                   void* stream) {
  MCN_REQUIRE(x && y && tap, "maxpool_fwd_tap: null argument");
  return maxpool_tap_impl(true, dtype, x, nullptr, N, H, W, C, kh, kw, sh, sw, pad_t, pad_l, Ho, Wo, y, tap, stream);
}
extern "C" int mcn_maxpool_bwd_tap(int dtype, const void* dy, const uint8_t* tap, int N, int H, int W, int C,
                                   int kh, int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo,
                                   void* dx, void* stream) {
  MCN_REQUIRE(dy && dx && tap, "maxpool_bwd_tap: null argument");
  return maxpool_tap_impl(false, dtype, dy, tap, N, H, W, C, kh, kw, sh, sw, pad_t, pad_l, Ho, Wo, dx, nullptr,
                          stream);
}
extern "C" int mcn_maxpool_tap_to_argmax(const uint8_t* tap, int N, int H, int W, int C, int kh, int kw, int sh,
                                         int sw, int pad_t, int pad_l, int Ho, int Wo, int32_t* argmax,
                                         void* stream) {
  MCN_REQUIRE(tap && argmax && kw > 0, "maxpool_tap_to_argmax: bad argument");
  (void)H;
  (void)kh;
  const long long total = (long long)N * Ho * Wo * C;
  ::mcn::launch(maxpool_tap_to_argmax_kernel, grid_for(total, 256), 256, 0, static_cast<cudaStream_t>(stream), 
      tap, total, W, C, kw, sh, sw, pad_t, pad_l, Ho, Wo, argmax);
  return after_launch("maxpool_tap_to_argmax");
}

extern "C" int mcn_avgpool_fwd(int dtype, const void* x, int N, int H, int W, int C, int kh,
                               int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo,
                               void* y, void* stream) {
  MCN_REQUIRE(x && y, "avgpool_fwd: null argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    long long total = (long long)N * Ho * Wo * C;
    ::mcn::launch(avgpool_fwd_kernel<T>, grid_for(total, 256), 256, 0, st, 
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
    ::mcn::launch(avgpool_bwd_kernel<T>, grid_for(total, 256), 256, 0, st, 
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
      ::mcn::launch(gap_fwd_kernel<T, float>, grid, block, 0, st, static_cast<const T*>(x), HW, C,
                                                       static_cast<float*>(y));
    else
      ::mcn::launch(gap_fwd_kernel<T, __nv_bfloat16>, grid, block, 0, st, 
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
    constexpr int V = Vec16<T>::N;
    if (C % V == 0 && reinterpret_cast<uintptr_t>(dx) % 16 == 0) {
      const long long nvec = total / V;
      if (dy_dtype == MCN_F32)
        ::mcn::launch(gap_bwd_vec_kernel<T, float>, grid_for(nvec, 256), 256, 0, st,
                      static_cast<const float*>(dy), HW, C, nvec, static_cast<T*>(dx));
      else
        ::mcn::launch(gap_bwd_vec_kernel<T, __nv_bfloat16>, grid_for(nvec, 256), 256, 0, st,
                      static_cast<const __nv_bfloat16*>(dy), HW, C, nvec, static_cast<T*>(dx));
    } else if (dy_dtype == MCN_F32)
      ::mcn::launch(gap_bwd_kernel<T, float>, grid_for(total, 256), 256, 0, st, 
          static_cast<const float*>(dy), HW, C, total, static_cast<T*>(dx));
    else
      ::mcn::launch(gap_bwd_kernel<T, __nv_bfloat16>, grid_for(total, 256), 256, 0, st, 
          static_cast<const __nv_bfloat16*>(dy), HW, C, total, static_cast<T*>(dx));
  });
  return after_launch("gap_bwd");
}
