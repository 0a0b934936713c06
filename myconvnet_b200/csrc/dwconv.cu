// Depthwise convolution, NHWC, channel multiplier 1 — the EfficientNet path
// (tf.nn.depthwise_conv2d, reference convnet.py:1634-1650; efficientnet.py:126-197: 16 layers,
// 3x3 / 5x5, stride 1 / 2, 32..1152 channels).  35 MMAC per image against ~20 MB of activations:
// pure bandwidth, graded on (|x| + |y|) * s bytes for every pass.
//
// A thread owns one 16-byte channel vector (8 bf16 / 4 fp32 channels) and TW consecutive output
// columns of one output row, everything else is unrolled at compile time (K, S, TW template
// parameters):
//   * each filter row's K weight vectors are loaded once per thread and reused for the TW pixels,
//   * the input row segment the TW windows cover ((TW-1)*S + K vectors) is loaded once and every
//     vector is reused by up to K taps — 5x5 stride 1: 10 activation loads per output vector
//     instead of 25, which moves the kernel from the L1-load limit back to the HBM limit,
//   * threads of a warp are adjacent channel vectors: every access is a coalesced 16-byte load.
// bwd-data with stride 1 is the same kernel on dy with mirrored taps; with stride 2 each input
// pixel gathers only the taps of matching parity.  bwd-filter: a thread owns (channel vector,
// filter row), keeps K x V fp32 sums in registers while it walks its pixels (TW per step), the
// block adds its pixel lanes in shared memory in a fixed order and writes ONE slice per block row
// chunk; splitk_reduce adds the slices in order (no atomics: bit-reproducible).
// Shapes the templates do not cover (multiplier > 1, channels not a multiple of the vector, other
// kernel sizes, dilation) use the generic kernels of conv_direct.cu.
#include "mcn_common.cuh"
#include "xsum.cuh"

namespace mcn {
namespace {

template <typename T>
struct WVec {
  static constexpr int V = Vec16<T>::N;
  float w[V];
};
// weights are the fp32 masters [kh][kw][C] (mult == 1)
template <int V>
__device__ __forceinline__ void load_w(const float* __restrict__ p, float (&w)[V]) {
#pragma unroll
  for (int i = 0; i < V; i += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
    w[i] = t.x;
    w[i + 1] = t.y;
    w[i + 2] = t.z;
    w[i + 3] = t.w;
  }
}

// Forward (kFlip = false) and stride-1 backward-data (kFlip = true: x := dy, taps mirrored, the
// caller passes pad := K-1-pad).  in: [N,H,W,C], out: [N,Ho,Wo,C].
template <typename T, int K, int S, int TW, bool kFlip>
__global__ void __launch_bounds__(256)
dw_fwd_kernel(const T* __restrict__ in, const float* __restrict__ wt, int N, int H, int W, int C, int Ho, int Wo,
              int pad_t, int pad_l, T* __restrict__ out) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  constexpr int SEG = (TW - 1) * S + K;
  const int cv = C / V;
  const int qt = (Wo + TW - 1) / TW;
  const uint32_t total = (uint32_t)N * Ho * qt * cv;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c0 = (int)(i % (uint32_t)cv) * V;
    uint32_t r = i / (uint32_t)cv;
    const int q0 = (int)(r % (uint32_t)qt) * TW;
    r /= (uint32_t)qt;
    const int p = (int)(r % (uint32_t)Ho);
    const int n = (int)(r / (uint32_t)Ho);
    float acc[TW][V];
#pragma unroll
    for (int t = 0; t < TW; ++t)
#pragma unroll
      for (int e = 0; e < V; ++e) acc[t][e] = 0.f;
    const int w0 = q0 * S - pad_l;
    const T* img = in + (long long)n * H * W * C + c0;
#pragma unroll
    for (int a = 0; a < K; ++a) {
      const int h = p * S + a - pad_t;
      if (h < 0 || h >= H) continue;
      Vec16<T> seg[SEG];
#pragma unroll
      for (int j = 0; j < SEG; ++j) {
        const int w = w0 + j;
        if (w >= 0 && w < W) seg[j] = ld_vec(img + ((long long)h * W + w) * C);
        else seg[j].raw = decltype(seg[j].raw){};
      }
#pragma unroll
      for (int b = 0; b < K; ++b) {
        float wv[V];
        const int tap = kFlip ? (K - 1 - a) * K + (K - 1 - b) : a * K + b;
        load_w<V>(wt + (long long)tap * C + c0, wv);
#pragma unroll
        for (int t = 0; t < TW; ++t)
#pragma unroll
          for (int e = 0; e < V; ++e) acc[t][e] = fmaf(seg[t * S + b].get(e), wv[e], acc[t][e]);
      }
    }
    T* o = out + (((long long)n * Ho + p) * Wo + q0) * C + c0;
#pragma unroll
    for (int t = 0; t < TW; ++t)
      if (q0 + t < Wo) {
        Vec16<T> ov;
#pragma unroll
        for (int e = 0; e < V; ++e) ov.set(e, acc[t][e]);
        st_vec(o + (long long)t * C, ov);
      }
  }
}

// Backward-data for stride S > 1: input pixel (h, w) receives dy[(h+pad_t-a)/S, (w+pad_l-b)/S] * w[a,b]
// for the taps where both divisions are exact: a = a0, a0+S, ... with a0 = (h+pad_t) % S.
template <typename T, int K, int S>
__global__ void __launch_bounds__(256)
dw_bwd_data_strided_kernel(const T* __restrict__ dy, const float* __restrict__ wt, int N, int H, int W, int C,
                           int Ho, int Wo, int pad_t, int pad_l, T* __restrict__ dx) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  constexpr int NT = (K + S - 1) / S;          // taps of one parity along an axis
  const int cv = C / V;
  const uint32_t total = (uint32_t)N * H * W * cv;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c0 = (int)(i % (uint32_t)cv) * V;
    uint32_t r = i / (uint32_t)cv;
    const int w = (int)(r % (uint32_t)W);
    r /= (uint32_t)W;
    const int h = (int)(r % (uint32_t)H);
    const int n = (int)(r / (uint32_t)H);
    const int a0 = (h + pad_t) % S, b0 = (w + pad_l) % S;
    const int p0 = (h + pad_t) / S, q0 = (w + pad_l) / S;        // window of tap a0 / b0
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    const T* g = dy + (long long)n * Ho * Wo * C + c0;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int a = a0 + j * S, p = p0 - j;
      if (a >= K || p < 0 || p >= Ho) continue;
#pragma unroll
      for (int k = 0; k < NT; ++k) {
        const int b = b0 + k * S, q = q0 - k;
        if (b >= K || q < 0 || q >= Wo) continue;
        const Vec16<T> gv = ld_vec(g + ((long long)p * Wo + q) * C);
        float wv[V];
        load_w<V>(wt + (long long)(a * K + b) * C + c0, wv);
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = fmaf(gv.get(e), wv[e], acc[e]);
      }
    }
    Vec16<T> ov;
#pragma unroll
    for (int e = 0; e < V; ++e) ov.set(e, acc[e]);
    st_vec(dx + (((long long)n * H + h) * W + w) * C + c0, ov);
  }
}

// Backward-filter.  blockDim = (CVB channel vectors, K filter rows, PL pixel lanes); grid =
// (channel-vector groups, row chunks).  Thread (cv, a, lane) sums dw[a][0..K)[8 channels] over the
// output rows of its chunk (rows lane, lane+PL, ...) in TW-pixel steps; lanes are added in lane
// order through shared memory; the block writes its [K][K][CVB*V] partial into slice blockIdx.y.
template <typename T, int K, int S, int TW>
__global__ void __launch_bounds__(256)
dw_bwd_filter_kernel(const T* __restrict__ x, const T* __restrict__ dy, int N, int H, int W, int C, int Ho,
                     int Wo, int pad_t, int pad_l, int cvb, int pl, float* __restrict__ slices,
                     long long slice_stride) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  constexpr int SEG = (TW - 1) * S + K;
  extern __shared__ float sh[];                       // [pl][K][cvb][K*V]
  const int cvi = threadIdx.x, a = threadIdx.y, lane = threadIdx.z;
  const int cvg = blockIdx.x * cvb + cvi;
  const int cv = C / V;
  const bool live = cvg < cv;
  const int c0 = cvg * V;
  float acc[K][V];
#pragma unroll
  for (int b = 0; b < K; ++b)
#pragma unroll
    for (int e = 0; e < V; ++e) acc[b][e] = 0.f;
  const long long rows_total = (long long)N * Ho;
  const long long r0 = rows_total * blockIdx.y / gridDim.y, r1 = rows_total * (blockIdx.y + 1) / gridDim.y;
  if (live) {
    for (long long row = r0 + lane; row < r1; row += pl) {
      const int n = (int)(row / Ho), p = (int)(row % Ho);
      const int h = p * S + a - pad_t;
      if (h < 0 || h >= H) continue;
      const T* xr = x + (((long long)n * H + h) * W) * C + c0;
      const T* gr = dy + (((long long)n * Ho + p) * Wo) * C + c0;
      for (int q0 = 0; q0 < Wo; q0 += TW) {
        Vec16<T> seg[SEG], g[TW];
        const int w0 = q0 * S - pad_l;
#pragma unroll
        for (int j = 0; j < SEG; ++j) {
          const int w = w0 + j;
          if (w >= 0 && w < W) seg[j] = ld_vec_stream(xr + (long long)w * C);
          else seg[j].raw = decltype(seg[j].raw){};
        }
#pragma unroll
        for (int t = 0; t < TW; ++t) {
          if (q0 + t < Wo) g[t] = ld_vec_stream(gr + (long long)(q0 + t) * C);
          else g[t].raw = decltype(g[t].raw){};
        }
#pragma unroll
        for (int b = 0; b < K; ++b)
#pragma unroll
          for (int t = 0; t < TW; ++t)
#pragma unroll
            for (int e = 0; e < V; ++e) acc[b][e] = fmaf(seg[t * S + b].get(e), g[t].get(e), acc[b][e]);
      }
    }
  }
  // block reduction over the pixel lanes, fixed order
  float* mine = sh + (((size_t)lane * K + a) * cvb + cvi) * (K * V);
#pragma unroll
  for (int b = 0; b < K; ++b)
#pragma unroll
    for (int e = 0; e < V; ++e) mine[b * V + e] = acc[b][e];
  __syncthreads();
  if (lane == 0 && live) {
    float* o = slices + (long long)blockIdx.y * slice_stride;
#pragma unroll
    for (int b = 0; b < K; ++b)
#pragma unroll
      for (int e = 0; e < V; ++e) {
        float s = 0.f;
        for (int l = 0; l < pl; ++l) s += sh[(((size_t)l * K + a) * cvb + cvi) * (K * V) + b * V + e];
        o[(long long)(a * K + b) * C + c0 + e] = s;
      }
  }
}

inline int grid_for(long long n, int block) {
  return (int)std::max<long long>(1, std::min<long long>((n + block - 1) / block, 32LL * num_sms()));
}

template <typename T, int K, int S, int TW, bool kFlip>
void launch_fwd(const void* in, const float* w, int N, int H, int W, int C, int Ho, int Wo, int pad_t, int pad_l,
                void* out, cudaStream_t st) {
  const long long total = (long long)N * Ho * ((Wo + TW - 1) / TW) * (C / Vec16<T>::N);
  ::mcn::launch(dw_fwd_kernel<T, K, S, TW, kFlip>, grid_for(total, 256), 256, 0, st, 
      static_cast<const T*>(in), w, N, H, W, C, Ho, Wo, pad_t, pad_l, static_cast<T*>(out));
}

}  // namespace

// Returns true when a specialised kernel handled the call.
bool dw_fast_eligible(const mcn_conv_desc* d, int mult, int dtype, int wdtype) {
  const int V = dtype == MCN_BF16 ? 8 : 4;
  if (mult != 1 || wdtype != MCN_F32 || d->Cin % V != 0) return false;
  if (d->kh != d->kw || d->sh != d->sw || d->dh != 1 || d->dw != 1) return false;
  if (!((d->kh == 3 || d->kh == 5) && (d->sh == 1 || d->sh == 2))) return false;
  const long long big = std::max((long long)d->N * d->H * d->W, (long long)d->N * d->Ho * d->Wo) * (d->Cin / V);
  return big < (1LL << 31);
}

#define MCN_DW_DISPATCH_KS(K_, S_, ...)            \
  do {                                             \
    if (K_ == 3 && S_ == 1) { constexpr int K = 3, S = 1; __VA_ARGS__; }      \
    else if (K_ == 3 && S_ == 2) { constexpr int K = 3, S = 2; __VA_ARGS__; } \
    else if (K_ == 5 && S_ == 1) { constexpr int K = 5, S = 1; __VA_ARGS__; } \
    else { constexpr int K = 5, S = 2; __VA_ARGS__; }                         \
  } while (0)

int dw_fast_fwd(const mcn_conv_desc* d, int dtype, const void* x, const float* w, void* y, cudaStream_t st) {
  MCN_DISPATCH_DTYPE(dtype, T, MCN_DW_DISPATCH_KS(d->kh, d->sh, {
    if (d->Wo >= 4) launch_fwd<T, K, S, 4, false>(x, w, d->N, d->H, d->W, d->Cin, d->Ho, d->Wo, d->pad_t, d->pad_l, y, st);
    else launch_fwd<T, K, S, 1, false>(x, w, d->N, d->H, d->W, d->Cin, d->Ho, d->Wo, d->pad_t, d->pad_l, y, st);
  }));
  return after_launch("dwconv_fwd");
}

int dw_fast_bwd_data(const mcn_conv_desc* d, int dtype, const void* dy, const float* w, void* dx, cudaStream_t st) {
  MCN_DISPATCH_DTYPE(dtype, T, MCN_DW_DISPATCH_KS(d->kh, d->sh, {
    if (S == 1) {
      // correlation of dy with the mirrored taps; dx[h] uses dy rows h + pad - a  ->  pad' = K-1-pad
      if (d->W >= 4)
        launch_fwd<T, K, 1, 4, true>(dy, w, d->N, d->Ho, d->Wo, d->Cin, d->H, d->W, K - 1 - d->pad_t, K - 1 - d->pad_l, dx, st);
      else
        launch_fwd<T, K, 1, 1, true>(dy, w, d->N, d->Ho, d->Wo, d->Cin, d->H, d->W, K - 1 - d->pad_t, K - 1 - d->pad_l, dx, st);
    } else {
      const long long total = (long long)d->N * d->H * d->W * (d->Cin / Vec16<T>::N);
      ::mcn::launch(dw_bwd_data_strided_kernel<T, K, (S == 1 ? 2 : S)>, grid_for(total, 256), 256, 0, st, 
          static_cast<const T*>(dy), w, d->N, d->H, d->W, d->Cin, d->Ho, d->Wo, d->pad_t, d->pad_l,
          static_cast<T*>(dx));
    }
  }));
  return after_launch("dwconv_bwd_data");
}

int dw_fast_bwd_filter(const mcn_conv_desc* d, int dtype, const void* x, const void* dy, float* dw, cudaStream_t st) {
  const int V = dtype == MCN_BF16 ? 8 : 4;
  const int cv = d->Cin / V, K0 = d->kh;
  // block: cvb channel vectors x K rows x pl lanes <= 256 threads, shared memory pl*K*cvb*K*V floats <= 48 KB
  int cvb = std::min(cv, 16);
  int pl = std::max(1, 256 / (cvb * K0));
  while (pl > 1 && (size_t)pl * K0 * cvb * K0 * V * 4 > 48 * 1024) --pl;
  const long long rows = (long long)d->N * d->Ho;
  const int groups = (cv + cvb - 1) / cvb;
  long long chunks = std::max<long long>(1, std::min<long long>(rows / std::max(pl, 1), (4LL * num_sms() + groups - 1) / groups));
  const long long n = (long long)K0 * K0 * d->Cin;
  const long long stride = (n + 63) / 64 * 64;
  const Workspace ws = current_workspace();
  MCN_REQUIRE(ws.base != nullptr, "dwconv_bwd_filter: no workspace registered (mcn_set_workspace)");
  const long long cap = (ws.bytes - kWsSplitOff) / (stride * 4);
  MCN_REQUIRE(cap >= 1, "dwconv_bwd_filter: workspace too small");
  chunks = std::min(chunks, cap);
  float* slices = reinterpret_cast<float*>(ws.base + kWsSplitOff);
  const size_t smem = (size_t)pl * K0 * cvb * K0 * V * sizeof(float);
  dim3 grid((unsigned)groups, (unsigned)chunks), block((unsigned)cvb, (unsigned)K0, (unsigned)pl);
  MCN_DISPATCH_DTYPE(dtype, T, MCN_DW_DISPATCH_KS(d->kh, d->sh, {
    if (d->Wo >= 4)
      ::mcn::launch(dw_bwd_filter_kernel<T, K, S, 4>, grid, block, smem, st, 
          static_cast<const T*>(x), static_cast<const T*>(dy), d->N, d->H, d->W, d->Cin, d->Ho, d->Wo, d->pad_t,
          d->pad_l, cvb, pl, slices, stride);
    else
      ::mcn::launch(dw_bwd_filter_kernel<T, K, S, 1>, grid, block, smem, st, 
          static_cast<const T*>(x), static_cast<const T*>(dy), d->N, d->H, d->W, d->Cin, d->Ho, d->Wo, d->pad_t,
          d->pad_l, cvb, pl, slices, stride);
  }));
  int rc = after_launch("dwconv_bwd_filter");
  if (rc) return rc;
  return launch_splitk_reduce(slices, stride, (int)chunks, n, dw, st);
}

}  // namespace mcn
