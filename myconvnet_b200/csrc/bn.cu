// Batch normalisation, forward and backward, NHWC, bandwidth-bound.
// Replaces tf.nn.fused_batch_norm at reference convnet.py:1883-1896,1916 and the hand-written
// moving-statistics update at convnet.py:1898-1901; activation (convnet.py:2536-2556) and the
// residual add (convnet.py:2509-2511) are fused into the same pass.
//
// Access pattern: the tensor is walked as a flat array of 16-byte vectors.  The launch makes the
// total thread count a multiple of (C / vector width), so a thread always lands on the same
// channel group: its per-channel constants live in registers and every warp reads contiguous
// 512-byte segments.  Reductions go thread partial (fp32, a few rows) -> shared memory -> one
// add per channel per block into the exact fixed-point accumulator of xsum.cuh (order-
// independent, so the step is bit-reproducible); the last block decodes it into the output.
#include <cstdlib>

#include "mcn_common.cuh"
#include "xsum.cuh"

namespace mcn {
namespace {

constexpr int kUnroll = 4;

// MCN_BN_RUNS=0 selects the grid-stride kernels everywhere (A/B switch)
// the run-based backward-apply kernel measured slower than the grid-stride one in the training
// step (three read streams): off unless MCN_BN_BWD_RUNS=1
inline bool bn_bwd_use_runs() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCN_BN_BWD_RUNS");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}
// rows (reduce) / vectors (apply) in flight per thread of the backward kernels: MCN_BN_BWD_UNROLL=2|4
inline int bn_bwd_unroll() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCN_BN_BWD_UNROLL");
    v = (e && e[0] == '4') ? 4 : 2;   // measured: 4 spills and costs occupancy (25.6 -> 28.1 ms/step)
  }
  return v;
}
// the same for the variants without a y stream (two read streams): default 4, MCN_BN_NOY_UNROLL=2|4
inline int bn_noy_unroll() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCN_BN_NOY_UNROLL");
    v = (e && e[0] == '2') ? 2 : 4;
  }
  return v;
}
inline bool bn_use_runs() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCN_BN_RUNS");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

struct ChanLaunch {
  int cv;      // channel vectors per row
  int block;   // threads per block, multiple of cv
  int grid;
};

template <typename T>
bool plan(long long rows, int C, ChanLaunch* L, int blocks_per_sm, int max_block = 512) {
  constexpr int V = Vec16<T>::N;
  if (C % V != 0) return false;
  int cv = C / V;
  if (cv > 512) return false;
  if (cv > max_block) max_block = 512;
  L->cv = cv;
  L->block = (max_block / cv) * cv;
  long long nvec = rows * cv;
  long long want = (nvec + (long long)L->block * kUnroll - 1) / ((long long)L->block * kUnroll);
  long long cap = (long long)num_sms() * blocks_per_sm;
  L->grid = (int)std::max<long long>(1, std::min(want, cap));
  return true;
}

// ---------------------------------------------------------------- reductions over rows
// Grid = (channel slabs, row chunks).  A slab is up to 16 channel vectors (256 B of a row), so a
// warp reads two 256-byte row segments per load; a block owns one slab for a contiguous chunk of
// rows.  Only the blocks of the same slab ever add into the same channel's accumulator, which
// keeps the number of contended global atomics per address at the row-chunk count (tens), not the
// whole grid (hundreds) — that contention was the fixed cost of the small late-layer tensors.
struct SlabLaunch {
  int cv, slab_v, rowlanes;
  dim3 grid;
};

template <typename T>
bool plan_slab(long long rows, int C, SlabLaunch* L, int blocks_total) {
  constexpr int V = Vec16<T>::N;
  if (C % V != 0) return false;
  L->cv = C / V;
  L->slab_v = std::min(L->cv, 16);
  L->rowlanes = 256 / L->slab_v;
  int slabs = (L->cv + L->slab_v - 1) / L->slab_v;
  long long per_block_rows = (long long)L->rowlanes * 8;   // >= 8 rows per thread
  long long chunks = std::max<long long>(1, std::min<long long>((rows + per_block_rows - 1) / per_block_rows,
                                                                std::max(1, blocks_total / slabs)));
  chunks = std::min<long long>(chunks, 65535);
  L->grid = dim3((unsigned)slabs, (unsigned)chunks);
  return true;
}

// Last block of a channel group (ticket counter per group): limbs [3][2C] -> out1[c] += sum 0,
// out2[c] += sum 1 for the group's channels [c0, c1), limbs and counter cleared for the next
// launch.  All limb loads are issued before anything is consumed (one L2 round trip).
template <typename TAcc>
__device__ __forceinline__ void xs_decode_pair(const XsScratch& xsc, int group, int C, int c0, int c1,
                                               TAcc* __restrict__ out1, TAcc* __restrict__ out2) {
  const int nthreads = blockDim.x * blockDim.y;
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  const int w = c1 - c0;
  for (int i = tid; i < 2 * w; i += nthreads) {
    const int which = i >= w ? 1 : 0;
    const int c = c0 + i - which * w;
    TAcc* o = which ? out2 + c : out1 + c;
    const TAcc prev = *o;
    const double v = xs::read_clear(xsc.limbs, 2 * C, which * C + c);
    *o = static_cast<TAcc>(static_cast<double>(prev) + v);
  }
  if (tid == 0) xs::release(xsc.counter + group);
}

// Block-level finish shared by the reduction kernels: sh holds [2][rowlanes][slab_v*V] partials.
// Only the blocks of one slab (blockIdx.x) add into the same channels, so the ticket is per slab:
// the last block of a slab decodes just that slab's channels — no grid-wide serial tail.
template <int V, typename TAcc>
__device__ __forceinline__ void slab_finish(float* sh, int slab_v, int rowlanes, int slab, int C,
                                            TAcc* out1, TAcc* out2, const XsScratch& xsc) {
  __syncthreads();
  const int width = slab_v * V;
  for (int t = threadIdx.x; t < 2 * width; t += blockDim.x) {
    const int which = t / width, e = t - which * width;
    const float* src = sh + (size_t)which * rowlanes * width + e;
    float acc = 0.f;
    for (int r = 0; r < rowlanes; ++r) acc += src[(size_t)r * width];
    const int c = slab * width + e;
    if (c < C) xs::add(xsc.limbs, 2 * C, which * C + c, acc);
  }
  if (xs::block_is_last(xsc.counter + slab, gridDim.y))
    xs_decode_pair<TAcc>(xsc, slab, C, slab * width, min(C, (slab + 1) * width), out1, out2);
}

template <typename T>
__global__ void __launch_bounds__(256)
bn_stats_kernel(const T* __restrict__ x, long long rows, int C, int slab_v, int rowlanes,
                double* __restrict__ sums, XsScratch xsc) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  extern __shared__ float sh[];
  const int sv = threadIdx.x % slab_v, rl = threadIdx.x / slab_v;
  const int vec = blockIdx.x * slab_v + sv;
  const bool active = rl < rowlanes && vec * V < C;
  const long long r0 = rows * blockIdx.y / gridDim.y, r1 = rows * (blockIdx.y + 1) / gridDim.y;
  float s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0.f;
  if (active) {
    const T* base = x + (long long)vec * V;
    long long r = r0 + rl;
    for (; r + (kUnroll - 1) * rowlanes < r1; r += (long long)kUnroll * rowlanes) {
      Vec16<T> a[kUnroll];
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) a[u] = ld_vec_stream(base + (r + (long long)u * rowlanes) * C);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u)
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float f = a[u].get(i);
          s1[i] += f;
          s2[i] = fmaf(f, f, s2[i]);
        }
    }
    for (; r < r1; r += rowlanes) {
      Vec16<T> a = ld_vec_stream(base + r * C);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        float f = a.get(i);
        s1[i] += f;
        s2[i] = fmaf(f, f, s2[i]);
      }
    }
  }
  if (rl < rowlanes) {
    const int width = slab_v * V;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      sh[(size_t)rl * width + sv * V + i] = s1[i];
      sh[(size_t)(rowlanes + rl) * width + sv * V + i] = s2[i];
    }
  }
  slab_finish<V, double>(sh, slab_v, rowlanes, blockIdx.x, C, sums, sums + C, xsc);
}

// scalar fallback (C not a multiple of the vector width)
template <typename T>
__global__ void bn_stats_scalar_kernel(const T* __restrict__ x, long long rows, int C,
                                       double* __restrict__ sums, XsScratch xsc) {
  MCN_PDL_PROLOGUE();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    long long r0 = rows * blockIdx.y / gridDim.y, r1 = rows * (blockIdx.y + 1) / gridDim.y;
    float s1 = 0.f, s2 = 0.f;
    for (long long r = r0; r < r1; ++r) {
      float f = to_f32(x[r * C + c]);
      s1 += f;
      s2 = fmaf(f, f, s2);
    }
    xs::add(xsc.limbs, 2 * C, c, s1);
    xs::add(xsc.limbs, 2 * C, C + c, s2);
  }
  if (xs::block_is_last(xsc.counter + blockIdx.x, gridDim.y))
    xs_decode_pair<double>(xsc, blockIdx.x, C, blockIdx.x * blockDim.x, min(C, (int)((blockIdx.x + 1) * blockDim.x)),
                           sums, sums + C);
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, int C, float eps,
                                   float momentum, float* __restrict__ mean,
                                   float* __restrict__ invstd, float* __restrict__ moving_mean,
                                   float* __restrict__ moving_var) {
  MCN_PDL_PROLOGUE();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double m = sums[c] / count;
  double var = sums[C + c] / count - m * m;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)m;
  invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (moving_mean != nullptr) {
    // tf.nn.fused_batch_norm returns the Bessel-corrected variance; convnet.py:1900-1901
    double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    moving_mean[c] = momentum * moving_mean[c] + (1.f - momentum) * (float)m;
    moving_var[c] = momentum * moving_var[c] + (1.f - momentum) * (float)unbiased;
  }
}

// Frozen batch-norm (update off: blocks outside blocks_to_train or update_batch_norm=False): the
// reference normalises with the STORED moving statistics even while training
// (fused_batch_norm(is_training=False), convnet.py:1916-1924).  This fills the saved mean / invstd
// the apply and backward kernels read, so both passes use the same constants.
__global__ void bn_frozen_stats_kernel(const float* __restrict__ moving_mean,
                                       const float* __restrict__ moving_var, int C, float eps,
                                       float* __restrict__ mean, float* __restrict__ invstd) {
  MCN_PDL_PROLOGUE();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  mean[c] = moving_mean[c];
  invstd[c] = static_cast<float>(1.0 / sqrt(static_cast<double>(moving_var[c]) + static_cast<double>(eps)));
}

// ---------------------------------------------------------------- forward apply
// Per-channel scale/shift of a thread's V channels are held in registers; the tensor is then
// streamed once.  Three sources for the normalisation constants:
//   kMode 0: saved mean / invstd            (mcn_bn_apply)
//   kMode 1: moving mean / variance + eps   (mcn_bn_infer)
//   kMode 2: the fp64 sums [sum x | sum x^2] of this step (mcn_bn_apply_stats): every thread
//            finalises its own channels (mean, biased variance; the cancellation-prone
//            E[x^2] - mean^2 stays in fp64), and block 0 also writes the saved mean / invstd for
//            the backward pass and the reference's moving-statistics update
//            (convnet.py:1898-1901, Bessel-corrected variance) — no separate finalize launch.
struct BnSumsArgs {
  const double* sums;
  double inv_count, bessel;
  float eps, momentum;
  float* save_mean;
  float* save_invstd;
  float* moving_mean;
  float* moving_var;
};

template <int V, int kMode>
__device__ __forceinline__ void bn_apply_coefs(int c0, int C, int cv, const float* __restrict__ mean,
                                               const float* __restrict__ invstd_or_var, float eps,
                                               const float* __restrict__ gamma,
                                               const float* __restrict__ beta, const BnSumsArgs& fs,
                                               float* sc, float* sf) {
#pragma unroll
  for (int i = 0; i < V; ++i) {
    float mu, is;
    if (kMode == 2) {
      const double m = fs.sums[c0 + i] * fs.inv_count;
      double var = fs.sums[C + c0 + i] * fs.inv_count - m * m;
      if (var < 0.0) var = 0.0;
      mu = static_cast<float>(m);
      // 1/sqrt(var + eps) correctly rounded from the fp64 variance (as bn_finalize_kernel does) at
      // the cost of four fp64 operations: fp32 seed, one Newton step in fp64 (error ~1e-14).  A ReLU
      // network amplifies even 1-ulp differences of invstd into mask flips, so this matters for
      // step-to-step reproducibility against the oracle.
      const double ve = var + static_cast<double>(fs.eps);
      double yd = static_cast<double>(rsqrtf(static_cast<float>(ve)));
      yd = yd * (1.5 - 0.5 * ve * yd * yd);
      is = static_cast<float>(yd);
      if (blockIdx.x == 0 && threadIdx.x < cv) {
        fs.save_mean[c0 + i] = mu;
        fs.save_invstd[c0 + i] = is;
        if (fs.moving_mean != nullptr) {
          const float unbiased = static_cast<float>(var * fs.bessel);
          fs.moving_mean[c0 + i] = fs.momentum * fs.moving_mean[c0 + i] + (1.f - fs.momentum) * mu;
          fs.moving_var[c0 + i] = fs.momentum * fs.moving_var[c0 + i] + (1.f - fs.momentum) * unbiased;
        }
      }
    } else {
      mu = mean[c0 + i];
      is = (kMode == 1) ? rsqrtf(invstd_or_var[c0 + i] + eps) : invstd_or_var[c0 + i];
    }
    const float g = gamma ? gamma[c0 + i] : 1.f;
    const float b = beta ? beta[c0 + i] : 0.f;
    sc[i] = g * is;
    sf[i] = b - mu * sc[i];
  }
}

template <typename T, int kMode>
__global__ void __launch_bounds__(512, 2)   // two blocks per SM: the fp64 finalize must not cost occupancy
bn_apply_kernel(const T* __restrict__ x, long long nvec, int cv, const float* __restrict__ mean,
                const float* __restrict__ invstd_or_var, float eps,
                const float* __restrict__ gamma, const float* __restrict__ beta,
                const T* __restrict__ residual, int act, float alpha, T* __restrict__ y,
                BnSumsArgs fs) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  const int c0 = (threadIdx.x % cv) * V;
  float sc[V], sf[V];
  bn_apply_coefs<V, kMode>(c0, cv * V, cv, mean, invstd_or_var, eps, gamma, beta, fs, sc, sf);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec;
       v += kUnroll * stride) {
    Vec16<T> a[kUnroll], r[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      long long vv = v + u * stride;
      if (vv < nvec) {
        a[u] = ld_vec_stream(x + vv * V);
        if (residual) r[u] = ld_vec_stream(residual + vv * V);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      long long vv = v + u * stride;
      if (vv < nvec) {
        Vec16<T> o;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float f = fmaf(a[u].get(i), sc[i], sf[i]);
          if (residual) f += r[u].get(i);
          o.set(i, act_fwd(act, f, alpha));
        }
        st_vec(y + vv * V, o);
      }
    }
  }
}

// Non-persistent variant ("runs"): a block of 256 threads owns contiguous runs of 256*U vectors
// (32 KB of bf16 at U = 8) and issues all U loads of a run before touching the data — measured
// 10-15 % faster than the grid-stride walk on 25-411 MB tensors (profiles/r01_stream_sweep.txt).
// Needs 256 % cv == 0 so that a thread keeps its channel group from run to run.  The loads of the
// first run are in flight while the per-channel constants are derived.
template <typename T, int kMode, int U, bool kRes>
__global__ void __launch_bounds__(256)
bn_apply_runs_kernel(const T* __restrict__ x, long long nvec, int cv, const float* __restrict__ mean,
                     const float* __restrict__ invstd_or_var, float eps,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     const T* __restrict__ residual, int act, float alpha, T* __restrict__ y,
                     BnSumsArgs fs, int runs_per_block) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  const int c0 = (threadIdx.x % cv) * V;
  long long v = (long long)blockIdx.x * runs_per_block * (256 * U) + threadIdx.x;
  Vec16<T> a[U], r[U];
  auto load = [&](long long vbase) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long vv = vbase + u * 256;
      if (vv < nvec) {
        a[u] = ld_vec_stream(x + vv * V);
        if (kRes) r[u] = ld_vec_stream(residual + vv * V);
      }
    }
  };
  load(v);
  float sc[V], sf[V];
  bn_apply_coefs<V, kMode>(c0, cv * V, cv, mean, invstd_or_var, eps, gamma, beta, fs, sc, sf);
  for (int run = 0; run < runs_per_block; ++run, v += 256 * U) {
    if (run > 0) load(v);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long vv = v + u * 256;
      if (vv < nvec) {
        Vec16<T> o;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float f = fmaf(a[u].get(i), sc[i], sf[i]);
          if (kRes) f += r[u].get(i);
          o.set(i, act_fwd(act, f, alpha));
        }
        st_vec(y + vv * V, o);
      }
    }
  }
}

template <typename T, bool kVarInput>
__global__ void bn_apply_scalar_kernel(const T* __restrict__ x, long long n, int C,
                                       const float* __restrict__ mean,
                                       const float* __restrict__ invstd_or_var, float eps,
                                       const float* __restrict__ gamma,
                                       const float* __restrict__ beta,
                                       const T* __restrict__ residual, int act, float alpha,
                                       T* __restrict__ y) {
  MCN_PDL_PROLOGUE();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float is = kVarInput ? rsqrtf(invstd_or_var[c] + eps) : invstd_or_var[c];
    float sc = (gamma ? gamma[c] : 1.f) * is;
    float f = fmaf(to_f32(x[i]), sc, (beta ? beta[c] : 0.f) - mean[c] * sc);
    if (residual) f += to_f32(residual[i]);
    y[i] = from_f32<T>(act_fwd(act, f, alpha));
  }
}

// ---------------------------------------------------------------- backward
// dz = dy * act'(.).  With `y` given the derivative comes from the forward output (relu family,
// also valid with a fused residual); otherwise the pre-activation is rebuilt from x.
template <typename T>
__device__ __forceinline__ float dz_of(float dy, float xv, float yv, bool have_y, float sc,
                                       float sf, int act, float alpha) {
  if (act == MCN_ACT_NONE) return dy;
  if (have_y) return dy * act_grad_from_y(act, yv, alpha);
  return dy * act_grad_from_x(act, fmaf(xv, sc, sf), alpha);
}

// U rows per thread are in flight at once (U*2 or U*3 16-byte loads): at 768 threads per SM two rows
// keep ~49 KB in flight, about what 6.5 TB/s x ~1 us of loaded latency needs per SM and no more
// (measured 0.65 of the copy peak); four rows double that.
// kMask: `y` is the ReLU bit mask written by the forward pass (bn_apply_pipe_kernel), one byte per
// 8-element vector, instead of the output tensor.
template <typename T, int U, bool kHaveY, bool kMask = false, int kAct = -1>
__global__ void __launch_bounds__(256, (U > 2 && kHaveY && !kMask) ? 2 : 3)
bn_bwd_reduce_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                     long long rows, int C, int slab_v, int rowlanes,
                     const float* __restrict__ mean, const float* __restrict__ invstd,
                     const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                     float alpha, float* __restrict__ sum_dz, float* __restrict__ sum_dz_xhat,
                     XsScratch xsc) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  extern __shared__ float sh[];
  const int sv = threadIdx.x % slab_v, rl = threadIdx.x / slab_v;
  const int vec = blockIdx.x * slab_v + sv;
  const bool active = rl < rowlanes && vec * V < C;
  const long long r0 = rows * blockIdx.y / gridDim.y, r1 = rows * (blockIdx.y + 1) / gridDim.y;
  float s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0.f;
  if (active) {
    const int c0 = vec * V;
    float sc[V], sf[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      sc[i] = (gamma ? gamma[c0 + i] : 1.f) * invstd[c0 + i];
      sf[i] = (beta ? beta[c0 + i] : 0.f) - mean[c0 + i] * sc[i];
    }
    constexpr bool have_y = kHaveY && !kMask;
    const uint8_t* mask = reinterpret_cast<const uint8_t*>(y);
    for (long long r = r0 + rl; r < r1; r += (long long)U * rowlanes) {
      Vec16<T> g[U], a[U], o[have_y ? U : 1];
      uint32_t mb[kMask ? U : 1];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        long long rr = r + (long long)u * rowlanes;
        if (rr < r1) {
          long long off = rr * C + c0;
          g[u] = ld_vec_stream(dy + off);
          a[u] = ld_vec_stream(x + off);
          if (have_y) o[u] = ld_vec_stream(y + off);
          if (kMask) mb[u] = __ldg(mask + (off >> 3));
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        long long rr = r + (long long)u * rowlanes;
        if (rr < r1) {
#pragma unroll
          for (int i = 0; i < V; ++i) {
            float xv = a[u].get(i);
            float dz;
            if (kMask) dz = ((mb[kMask ? u : 0] >> i) & 1u) ? g[u].get(i) : 0.f;
            else dz = dz_of<T>(g[u].get(i), xv, have_y ? o[have_y ? u : 0].get(i) : 0.f, have_y, sc[i], sf[i],
                               kAct >= 0 ? kAct : act, alpha);
            s1[i] += dz;
            s2[i] = fmaf(dz, xv, s2[i]);          // sum dz*x; turned into sum dz*xhat below
          }
        }
      }
    }
    // sum dz*xhat = invstd * (sum dz*x - mean * sum dz)
#pragma unroll
    for (int i = 0; i < V; ++i) s2[i] = invstd[c0 + i] * (s2[i] - mean[c0 + i] * s1[i]);
  }
  if (rl < rowlanes) {
    const int width = slab_v * V;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      sh[(size_t)rl * width + sv * V + i] = s1[i];
      sh[(size_t)(rowlanes + rl) * width + sv * V + i] = s2[i];
    }
  }
  slab_finish<V, float>(sh, slab_v, rowlanes, blockIdx.x, C, sum_dz, sum_dz_xhat, xsc);
}

template <typename T, int U, bool kHaveY>
__global__ void __launch_bounds__(sizeof(T) == 2 ? 256 : 512, sizeof(T) == 2 ? 3 : 1)   // fp32: cv up to 512
bn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                    long long nvec, int cv, const float* __restrict__ mean,
                    const float* __restrict__ invstd, const float* __restrict__ gamma,
                    const float* __restrict__ beta, int act, float alpha,
                    const float* __restrict__ sum_dz, const float* __restrict__ sum_dz_xhat,
                    float inv_count, T* __restrict__ dx, T* __restrict__ d_residual) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  const int c0 = (threadIdx.x % cv) * V;
  // dx = sc*(dz - k1 - xhat*k2) with xhat = (x-mu)*is  ==  A*dz + B*x + Cc
  float A[V], B[V], Cc[V], sf[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float mu = mean[c0 + i], is = invstd[c0 + i];
    const float k1 = sum_dz[c0 + i] * inv_count, k2 = sum_dz_xhat[c0 + i] * inv_count;
    A[i] = (gamma ? gamma[c0 + i] : 1.f) * is;
    sf[i] = (beta ? beta[c0 + i] : 0.f) - mu * A[i];
    B[i] = -A[i] * k2 * is;
    Cc[i] = A[i] * (k2 * mu * is - k1);
  }
  constexpr bool have_y = kHaveY;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec;
       v += U * stride) {
    Vec16<T> g[U], a[U], o[kHaveY ? U : 1];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      long long vv = v + u * stride;
      if (vv < nvec) {
        g[u] = ld_vec_stream(dy + vv * V);
        a[u] = ld_vec_stream(x + vv * V);
        if (have_y) o[u] = ld_vec_stream(y + vv * V);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      long long vv = v + u * stride;
      if (vv < nvec) {
        Vec16<T> ox, orr;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          float xv = a[u].get(i);
          float dz = dz_of<T>(g[u].get(i), xv, have_y ? o[kHaveY ? u : 0].get(i) : 0.f, have_y, A[i], sf[i],
                              act, alpha);
          ox.set(i, fmaf(A[i], dz, fmaf(B[i], xv, Cc[i])));
          orr.set(i, dz);
        }
        st_vec(dx + vv * V, ox);
        if (d_residual) st_vec(d_residual + vv * V, orr);
      }
    }
  }
}

// Run-based variant of the backward apply pass (see bn_apply_runs_kernel): 256 threads own
// contiguous runs of 256*U vectors, all loads of a run are issued before any arithmetic.
template <typename T, int U, bool kHaveY, bool kRes>
__global__ void __launch_bounds__(256)
bn_bwd_apply_runs_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                         long long nvec, int cv, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ gamma,
                         const float* __restrict__ beta, int act, float alpha,
                         const float* __restrict__ sum_dz, const float* __restrict__ sum_dz_xhat,
                         float inv_count, T* __restrict__ dx, T* __restrict__ d_residual,
                         int runs_per_block) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  const int c0 = (threadIdx.x % cv) * V;
  long long v = (long long)blockIdx.x * runs_per_block * (256 * U) + threadIdx.x;
  Vec16<T> g[U], a[U], o[U];
  auto load = [&](long long vbase) {
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long vv = vbase + u * 256;
      if (vv < nvec) {
        g[u] = ld_vec_stream(dy + vv * V);
        a[u] = ld_vec_stream(x + vv * V);
        if (kHaveY) o[u] = ld_vec_stream(y + vv * V);
      }
    }
  };
  load(v);
  float A[V], B[V], Cc[V], sf[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float mu = mean[c0 + i], is = invstd[c0 + i];
    const float k1 = sum_dz[c0 + i] * inv_count, k2 = sum_dz_xhat[c0 + i] * inv_count;
    A[i] = (gamma ? gamma[c0 + i] : 1.f) * is;
    sf[i] = (beta ? beta[c0 + i] : 0.f) - mu * A[i];
    B[i] = -A[i] * k2 * is;
    Cc[i] = A[i] * (k2 * mu * is - k1);
  }
  for (int run = 0; run < runs_per_block; ++run, v += 256 * U) {
    if (run > 0) load(v);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long vv = v + u * 256;
      if (vv < nvec) {
        Vec16<T> ox, orr;
#pragma unroll
        for (int i = 0; i < V; ++i) {
          const float xv = a[u].get(i);
          const float dz = dz_of<T>(g[u].get(i), xv, kHaveY ? o[u].get(i) : 0.f, kHaveY, A[i], sf[i],
                                    act, alpha);
          ox.set(i, fmaf(A[i], dz, fmaf(B[i], xv, Cc[i])));
          if (kRes) orr.set(i, dz);
        }
        st_vec(dx + vv * V, ox);
        if (kRes) st_vec(d_residual + vv * V, orr);
      }
    }
  }
}

// ---------------------------------------------------------------- asynchronous prefetch rings
// The register-staged kernels above keep 50-100 KB of loads in flight per SM, which is what a
// 400 MB tensor needs and far too little for the 25-150 MB tensors of the later blocks: those ran
// at 1.2-3.7 TB/s because the whole tensor is only a few "rounds" of the in-flight window and every
// round costs a loaded memory latency (~2.5 us).  The *_pipe kernels keep the same thread ->
// channel mapping but fetch with cp.async.cg (16 bytes, global -> shared, no registers held): each
// thread owns a private ring of D slots per input stream, issues D items ahead and consumes them
// in order with cp.async.wait_group.  S*D = 12 slots x 256 threads x 16 B = 48 KB per block, three
// blocks per SM: 144 KB in flight per SM whatever the arithmetic needs in registers.  A slot is
// only ever read by the thread that filled it, so no block synchronisation is involved.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
template <typename T>
__device__ __forceinline__ Vec16<T> lds_vec(uint32_t addr) {
  Vec16<T> v;
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "r"(addr)
               : "memory");
  v.raw = *reinterpret_cast<decltype(v.raw)*>(&r);
  return v;
}
constexpr int kPipeSlots = 12;                       // S * D
constexpr int kPipeBytes = kPipeSlots * 256 * 16;    // per 256-thread block

inline bool bn_use_pipe() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCN_BN_PIPE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// The slab reductions measured SLOWER with the ring (bn_bwd_reduce 3.67 -> 3.82 ms, bn_stats 0.67 ->
// 0.90 ms per step: their tail — block reduce, exact adds, ticket — not the fetch, is the fixed
// cost), so only the element-wise passes use it by default; MCN_BN_REDUCE_PIPE=1 selects it there too.
inline bool bn_reduce_pipe() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MCN_BN_REDUCE_PIPE");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v != 0 && bn_use_pipe();
}

// forward apply (+ residual): S = 1 or 2 streams
// kAct >= 0: the activation is a compile-time constant (ReLU / none: the per-element run-time switch
// over seven activations costs the streaming loop measurably), -1: run-time `act`.
template <typename T, int kMode, bool kRes, bool kMaskOut = false, int kAct = -1>
__global__ void __launch_bounds__(256, 3)
bn_apply_pipe_kernel(const T* __restrict__ x, long long nvec, int cv, const float* __restrict__ mean,
                     const float* __restrict__ invstd_or_var, float eps,
                     const float* __restrict__ gamma, const float* __restrict__ beta,
                     const T* __restrict__ residual, int act, float alpha, T* __restrict__ y,
                     BnSumsArgs fs, uint8_t* __restrict__ relu_mask) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  constexpr int S = kRes ? 2 : 1, D = kPipeSlots / S;
  extern __shared__ uint4 pipe_smem[];
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(pipe_smem)) + threadIdx.x * 16u;
  constexpr uint32_t kSlot = 256u * 16u;
  // (256 / cv) * cv threads work: a thread keeps its channel group from item to item for ANY cv <= 256
  const int nthr = (256 / cv) * cv;
  if (static_cast<int>(threadIdx.x) >= nthr) return;
  const long long stride = (long long)gridDim.x * nthr;
  long long v = (long long)blockIdx.x * nthr + threadIdx.x;
  // the loads do not depend on the coefficients: start them first
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const long long vv = v + d * stride;
    if (vv < nvec) {
      cp_async16(base + d * kSlot, x + vv * V);
      if (kRes) cp_async16(base + (D + d) * kSlot, residual + vv * V);
    }
    cp_async_commit();
  }
  const int c0 = (threadIdx.x % cv) * V;
  float sc[V], sf[V];
  bn_apply_coefs<V, kMode>(c0, cv * V, cv, mean, invstd_or_var, eps, gamma, beta, fs, sc, sf);
  int d = 0;
  for (; v < nvec; v += stride) {
    cp_async_wait<D - 1>();
    const Vec16<T> a = lds_vec<T>(base + d * kSlot);
    Vec16<T> r;
    if (kRes) r = lds_vec<T>(base + (D + d) * kSlot);
    Vec16<T> o;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float f = fmaf(a.get(i), sc[i], sf[i]);
      if (kRes) f += r.get(i);
      o.set(i, act_fwd(kAct >= 0 ? kAct : act, f, alpha));
    }
    st_vec(y + v * V, o);
    if (kMaskOut) {
      // one bit per element of the STORED output (bit i of byte v <=> y[v*8 + i] > 0): what the backward
      // passes need from y, in 1/16 of its bytes (bf16 only: V == 8)
      uint32_t m = 0;
#pragma unroll
      for (int i = 0; i < V; ++i) m |= (o.get(i) > 0.f ? 1u : 0u) << i;
      relu_mask[v] = static_cast<uint8_t>(m);
    }
    const long long vn = v + D * stride;
    if (vn < nvec) {
      cp_async16(base + d * kSlot, x + vn * V);
      if (kRes) cp_async16(base + (D + d) * kSlot, residual + vn * V);
    }
    cp_async_commit();
    d = (d + 1 == D) ? 0 : d + 1;
  }
  cp_async_wait<0>();
}

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
constexpr int kMaskRingBytes = (kPipeSlots / 2) * 256 * 4;   // bit-mask ring of the masked backward apply

// backward apply: S = 2 (dy, x) or 3 (+ y).  kY: 0 no y, 1 y is the output tensor, 2 y is the ReLU bit
// mask of the forward pass (one byte per vector): two full rings (D = 6) plus a ring of 4-byte words —
// each thread fetches the aligned word that holds its byte.
template <typename T, int kY, bool kRes, int kAct = -1>
__global__ void __launch_bounds__(256, 3)
bn_bwd_apply_pipe_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                         long long nvec, int cv, const float* __restrict__ mean,
                         const float* __restrict__ invstd, const float* __restrict__ gamma,
                         const float* __restrict__ beta, int act, float alpha,
                         const float* __restrict__ sum_dz, const float* __restrict__ sum_dz_xhat,
                         float inv_count, T* __restrict__ dx, T* __restrict__ d_residual) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  constexpr bool kHaveY = kY == 1, kMask = kY == 2;
  constexpr int S = kHaveY ? 3 : 2, D = kPipeSlots / S;
  extern __shared__ uint4 pipe_smem[];
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(pipe_smem)) + threadIdx.x * 16u;
  constexpr uint32_t kSlot = 256u * 16u;
  const uint32_t mbase = static_cast<uint32_t>(__cvta_generic_to_shared(pipe_smem)) + kPipeBytes + threadIdx.x * 4u;
  const uint8_t* mask = reinterpret_cast<const uint8_t*>(y);
  // (256 / cv) * cv threads work: a thread keeps its channel group from item to item for ANY cv <= 256
  const int nthr = (256 / cv) * cv;
  if (static_cast<int>(threadIdx.x) >= nthr) return;
  const long long stride = (long long)gridDim.x * nthr;
  long long v = (long long)blockIdx.x * nthr + threadIdx.x;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const long long vv = v + d * stride;
    if (vv < nvec) {
      cp_async16(base + d * kSlot, dy + vv * V);
      cp_async16(base + (D + d) * kSlot, x + vv * V);
      if (kHaveY) cp_async16(base + (2 * D + d) * kSlot, y + vv * V);
      if (kMask) cp_async4(mbase + d * 1024u, mask + (vv & ~3LL));
    }
    cp_async_commit();
  }
  const int c0 = (threadIdx.x % cv) * V;
  // dx = sc*(dz - k1 - xhat*k2) with xhat = (x-mu)*is  ==  A*dz + B*x + Cc
  float A[V], B[V], Cc[V], sf[V];
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const float mu = mean[c0 + i], is = invstd[c0 + i];
    const float k1 = sum_dz[c0 + i] * inv_count, k2 = sum_dz_xhat[c0 + i] * inv_count;
    A[i] = (gamma ? gamma[c0 + i] : 1.f) * is;
    sf[i] = (beta ? beta[c0 + i] : 0.f) - mu * A[i];
    B[i] = -A[i] * k2 * is;
    Cc[i] = A[i] * (k2 * mu * is - k1);
  }
  int d = 0;
  for (; v < nvec; v += stride) {
    cp_async_wait<D - 1>();
    const Vec16<T> g = lds_vec<T>(base + d * kSlot);
    const Vec16<T> a = lds_vec<T>(base + (D + d) * kSlot);
    Vec16<T> o;
    if (kHaveY) o = lds_vec<T>(base + (2 * D + d) * kSlot);
    uint32_t mb = 0;
    if (kMask) {
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(mb) : "r"(mbase + d * 1024u) : "memory");
      mb >>= 8u * (static_cast<uint32_t>(v) & 3u);
    }
    Vec16<T> ox, orr;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const float xv = a.get(i);
      float dz;
      if (kMask) dz = ((mb >> i) & 1u) ? g.get(i) : 0.f;
      else dz = dz_of<T>(g.get(i), xv, kHaveY ? o.get(i) : 0.f, kHaveY, A[i], sf[i], kAct >= 0 ? kAct : act, alpha);
      ox.set(i, fmaf(A[i], dz, fmaf(B[i], xv, Cc[i])));
      if (kRes) orr.set(i, dz);
    }
    st_vec(dx + v * V, ox);
    if (kRes) st_vec(d_residual + v * V, orr);
    const long long vn = v + D * stride;
    if (vn < nvec) {
      cp_async16(base + d * kSlot, dy + vn * V);
      cp_async16(base + (D + d) * kSlot, x + vn * V);
      if (kHaveY) cp_async16(base + (2 * D + d) * kSlot, y + vn * V);
      if (kMask) cp_async4(mbase + d * 1024u, mask + (vn & ~3LL));
    }
    cp_async_commit();
    d = (d + 1 == D) ? 0 : d + 1;
  }
  cp_async_wait<0>();
}

// row reductions on the slab grid of bn_stats_kernel / bn_bwd_reduce_kernel.
//   kStats: S = 1, sums of x and x^2 (fp64 outputs); otherwise S = 2 or 3, sums of dz and dz*xhat
template <typename T, bool kStats, bool kHaveY, typename TAcc>
__global__ void __launch_bounds__(256, 3)
bn_reduce_pipe_kernel(const T* __restrict__ dy, const T* __restrict__ x, const T* __restrict__ y,
                      long long rows, int C, int slab_v, int rowlanes,
                      const float* __restrict__ mean, const float* __restrict__ invstd,
                      const float* __restrict__ gamma, const float* __restrict__ beta, int act,
                      float alpha, TAcc* __restrict__ out1, TAcc* __restrict__ out2, XsScratch xsc) {
  MCN_PDL_PROLOGUE();
  constexpr int V = Vec16<T>::N;
  constexpr int S = kStats ? 1 : (kHaveY ? 3 : 2), D = kPipeSlots / S;
  extern __shared__ uint4 pipe_smem[];
  float* sh = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(pipe_smem) + kPipeBytes);
  const uint32_t base = static_cast<uint32_t>(__cvta_generic_to_shared(pipe_smem)) + threadIdx.x * 16u;
  constexpr uint32_t kSlot = 256u * 16u;
  const int sv = threadIdx.x % slab_v, rl = threadIdx.x / slab_v;
  const int vec = blockIdx.x * slab_v + sv;
  const bool active = rl < rowlanes && vec * V < C;
  const long long r0 = rows * blockIdx.y / gridDim.y, r1 = rows * (blockIdx.y + 1) / gridDim.y;
  const int c0 = vec * V;
  float s1[V], s2[V];
#pragma unroll
  for (int i = 0; i < V; ++i) s1[i] = s2[i] = 0.f;
  if (active) {
    long long r = r0 + rl;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const long long rr = r + (long long)d * rowlanes;
      if (rr < r1) {
        const long long off = rr * C + c0;
        cp_async16(base + d * kSlot, x + off);
        if (!kStats) cp_async16(base + (D + d) * kSlot, dy + off);
        if (!kStats && kHaveY) cp_async16(base + (2 * D + d) * kSlot, y + off);
      }
      cp_async_commit();
    }
    float sc[V], sf[V];
    if (!kStats) {
#pragma unroll
      for (int i = 0; i < V; ++i) {
        sc[i] = (gamma ? gamma[c0 + i] : 1.f) * invstd[c0 + i];
        sf[i] = (beta ? beta[c0 + i] : 0.f) - mean[c0 + i] * sc[i];
      }
    }
    int d = 0;
    for (; r < r1; r += rowlanes) {
      cp_async_wait<D - 1>();
      const Vec16<T> a = lds_vec<T>(base + d * kSlot);
      Vec16<T> g, o;
      if (!kStats) g = lds_vec<T>(base + (D + d) * kSlot);
      if (!kStats && kHaveY) o = lds_vec<T>(base + (2 * D + d) * kSlot);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float xv = a.get(i);
        if (kStats) {
          s1[i] += xv;
          s2[i] = fmaf(xv, xv, s2[i]);
        } else {
          const float dz = dz_of<T>(g.get(i), xv, kHaveY ? o.get(i) : 0.f, kHaveY, sc[i], sf[i], act, alpha);
          s1[i] += dz;
          s2[i] = fmaf(dz, xv, s2[i]);          // sum dz*x; turned into sum dz*xhat below
        }
      }
      const long long rn = r + (long long)D * rowlanes;
      if (rn < r1) {
        const long long off = rn * C + c0;
        cp_async16(base + d * kSlot, x + off);
        if (!kStats) cp_async16(base + (D + d) * kSlot, dy + off);
        if (!kStats && kHaveY) cp_async16(base + (2 * D + d) * kSlot, y + off);
      }
      cp_async_commit();
      d = (d + 1 == D) ? 0 : d + 1;
    }
    cp_async_wait<0>();
    if (!kStats) {
      // sum dz*xhat = invstd * (sum dz*x - mean * sum dz)
#pragma unroll
      for (int i = 0; i < V; ++i) s2[i] = invstd[c0 + i] * (s2[i] - mean[c0 + i] * s1[i]);
    }
  }
  if (rl < rowlanes) {
    const int width = slab_v * V;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      sh[(size_t)rl * width + sv * V + i] = s1[i];
      sh[(size_t)(rowlanes + rl) * width + sv * V + i] = s2[i];
    }
  }
  slab_finish<V, TAcc>(sh, slab_v, rowlanes, blockIdx.x, C, out1, out2, xsc);
}

// one-time opt-in to > 48 KB of dynamic shared memory for a pipe kernel instantiation
template <typename K>
static bool pipe_smem_ok(K kernel, size_t bytes) {
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
}

// scalar fallbacks for odd channel counts
template <typename T>
__global__ void bn_bwd_reduce_scalar_kernel(const T* dy, const T* x, const T* y, long long rows,
                                            int C, const float* mean, const float* invstd,
                                            const float* gamma, const float* beta, int act,
                                            float alpha, float* sum_dz, float* sum_dz_xhat,
                                            XsScratch xsc) {
  MCN_PDL_PROLOGUE();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    long long r0 = rows * blockIdx.y / gridDim.y, r1 = rows * (blockIdx.y + 1) / gridDim.y;
    float is = invstd[c], mu = mean[c];
    float sc = (gamma ? gamma[c] : 1.f) * is, sf = (beta ? beta[c] : 0.f) - mu * sc;
    float s1 = 0.f, s2 = 0.f;
    for (long long r = r0; r < r1; ++r) {
      float xv = to_f32(x[r * C + c]);
      float dz = dz_of<T>(to_f32(dy[r * C + c]), xv, y ? to_f32(y[r * C + c]) : 0.f, y != nullptr,
                          sc, sf, act, alpha);
      s1 += dz;
      s2 = fmaf(dz, (xv - mu) * is, s2);
    }
    xs::add(xsc.limbs, 2 * C, c, s1);
    xs::add(xsc.limbs, 2 * C, C + c, s2);
  }
  if (xs::block_is_last(xsc.counter + blockIdx.x, gridDim.y))
    xs_decode_pair<float>(xsc, blockIdx.x, C, blockIdx.x * blockDim.x, min(C, (int)((blockIdx.x + 1) * blockDim.x)),
                          sum_dz, sum_dz_xhat);
}
template <typename T>
__global__ void bn_bwd_apply_scalar_kernel(const T* dy, const T* x, const T* y, long long n, int C,
                                           const float* mean, const float* invstd,
                                           const float* gamma, const float* beta, int act,
                                           float alpha, const float* sum_dz,
                                           const float* sum_dz_xhat, float inv_count, T* dx,
                                           T* d_residual) {
  MCN_PDL_PROLOGUE();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    float is = invstd[c], mu = mean[c];
    float sc = (gamma ? gamma[c] : 1.f) * is, sf = (beta ? beta[c] : 0.f) - mu * sc;
    float xv = to_f32(x[i]);
    float dz = dz_of<T>(to_f32(dy[i]), xv, y ? to_f32(y[i]) : 0.f, y != nullptr, sc, sf, act, alpha);
    float xhat = (xv - mu) * is;
    dx[i] = from_f32<T>(sc * (dz - sum_dz[c] * inv_count - xhat * sum_dz_xhat[c] * inv_count));
    if (d_residual) d_residual[i] = from_f32<T>(dz);
  }
}

}  // namespace
}  // namespace mcn

using namespace mcn;

extern "C" int mcn_bn_stats(int dtype, const void* x, long long rows, int C, double* sums,
                            void* stream) {
  MCN_REQUIRE(x && sums && rows > 0 && C > 0, "bn_stats: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XsScratch xsc = xs_scratch(2 * C, "bn_stats");
  if (xsc.limbs == nullptr) return MCN_EINVAL;
  MCN_DISPATCH_DTYPE(dtype, T, {
    SlabLaunch L;
    if (plan_slab<T>(rows, C, &L, 3 * num_sms())) {
      size_t smem = 2 * (size_t)L.rowlanes * L.slab_v * Vec16<T>::N * sizeof(float);
      static bool pipe_ok = pipe_smem_ok(bn_reduce_pipe_kernel<T, true, false, double>, kPipeBytes + 16384);
      if (bn_reduce_pipe() && pipe_ok && xsc.limbs != nullptr)
        ::mcn::launch(bn_reduce_pipe_kernel<T, true, false, double>, L.grid, 256, kPipeBytes + smem, st,
                      static_cast<const T*>(nullptr), static_cast<const T*>(x), static_cast<const T*>(nullptr),
                      rows, C, L.slab_v, L.rowlanes, static_cast<const float*>(nullptr),
                      static_cast<const float*>(nullptr), static_cast<const float*>(nullptr),
                      static_cast<const float*>(nullptr), 0, 0.f, sums, sums + C, xsc);
      else
        ::mcn::launch(bn_stats_kernel<T>, L.grid, 256, smem, st, static_cast<const T*>(x), rows, C, L.slab_v,
                      L.rowlanes, sums, xsc);
    } else {
      MCN_REQUIRE((C + 127) / 128 <= kWsCounters, "bn_stats: too many channels (%d)", C);
      dim3 grid((C + 127) / 128, (unsigned)std::min<long long>(rows, 4LL * num_sms()));
      ::mcn::launch(bn_stats_scalar_kernel<T>, grid, 128, 0, st, static_cast<const T*>(x), rows, C, sums, xsc);
    }
  });
  return after_launch("bn_stats");
}

extern "C" int mcn_bn_frozen_stats(const float* moving_mean, const float* moving_var, int C, float eps,
                                   float* mean, float* invstd, void* stream) {
  MCN_REQUIRE(moving_mean && moving_var && mean && invstd && C > 0, "bn_frozen_stats: bad argument");
  ::mcn::launch(bn_frozen_stats_kernel, (C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), 
      moving_mean, moving_var, C, eps, mean, invstd);
  return after_launch("bn_frozen_stats");
}

extern "C" int mcn_bn_finalize(const double* sums, double count, int C, float eps, float momentum,
                               float* mean, float* invstd, float* moving_mean, float* moving_var,
                               void* stream) {
  MCN_REQUIRE(sums && mean && invstd && count > 0, "bn_finalize: bad argument");
  ::mcn::launch(bn_finalize_kernel, (C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), 
      sums, count, C, eps, momentum, mean, invstd, moving_mean, moving_var);
  return after_launch("bn_finalize");
}

template <int kMode>
static int bn_apply_impl(int dtype, const void* x, long long rows, int C, const float* mean,
                         const float* is_or_var, float eps, const float* gamma, const float* beta,
                         const void* residual, int act, float alpha, void* y, const BnSumsArgs& fs,
                         void* stream, uint8_t* relu_mask = nullptr) {
  MCN_REQUIRE(x && y && rows > 0, "bn_apply: bad argument");
  MCN_REQUIRE(kMode == 2 || (mean && is_or_var), "bn_apply: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  MCN_DISPATCH_DTYPE(dtype, T, {
    ChanLaunch L;
    constexpr int V = Vec16<T>::N;
    const int cvr = (C % V == 0) ? C / V : 0;
    // measured in the training step (profiles/r01_bn_runs_ab.txt): the run-based kernel wins for
    // the single-stream case (170 -> 140 us on 411 MB) and loses once a residual stream is added
    if (cvr > 0 && cvr <= 256 && bn_use_pipe()) {
      const long long nvec = rows * cvr;
      const int nthr = (256 / cvr) * cvr;
      const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((nvec + nthr - 1) / nthr, 3LL * num_sms()));
      auto go = [&](auto kern) {
        ::mcn::launch(kern, grid, 256, kPipeBytes, st, static_cast<const T*>(x), nvec, cvr, mean, is_or_var, eps,
                      gamma, beta, static_cast<const T*>(residual), act, alpha, static_cast<T*>(y), fs, relu_mask);
      };
      // the training pass (kMode 2) gets activation-specialised instantiations for ReLU / none
      constexpr bool kSpec = kMode == 2;
      if (relu_mask != nullptr) {
        if (residual == nullptr || sizeof(T) != 2 || act != MCN_ACT_RELU) {
          set_error("bn_apply_stats_mask: bf16 tensors with a fused residual and a ReLU only");
          return MCN_EINVAL;
        }
        go(bn_apply_pipe_kernel<T, kMode, true, true, MCN_ACT_RELU>);
      } else if (residual != nullptr) {
        if (kSpec && act == MCN_ACT_RELU) go(bn_apply_pipe_kernel<T, kMode, true, false, kSpec ? MCN_ACT_RELU : -1>);
        else go(bn_apply_pipe_kernel<T, kMode, true>);
      } else {
        if (kSpec && act == MCN_ACT_RELU) go(bn_apply_pipe_kernel<T, kMode, false, false, kSpec ? MCN_ACT_RELU : -1>);
        else if (kSpec && act == MCN_ACT_NONE) go(bn_apply_pipe_kernel<T, kMode, false, false, kSpec ? MCN_ACT_NONE : -1>);
        else go(bn_apply_pipe_kernel<T, kMode, false>);
      }
    } else if (relu_mask != nullptr) {
      set_error("bn_apply_stats_mask: needs bf16, C %% 8 == 0 and C <= 2048 (the pipe kernel)");
      return MCN_EINVAL;
    } else if (cvr > 0 && cvr <= 256 && 256 % cvr == 0 && bn_use_runs() && residual == nullptr) {
      const long long nvec = rows * cvr;
      const bool res = residual != nullptr;
      const int U = res ? 4 : 8;
      const long long runs = (nvec + 256LL * U - 1) / (256LL * U);
      // several runs per block amortise the constants' derivation once the grid is >= 4 waves
      int rpb = (int)std::max<long long>(1, std::min<long long>(8, runs / (4LL * 3 * num_sms())));
      const unsigned grid = (unsigned)((runs + rpb - 1) / rpb);
      if (res)
        ::mcn::launch(bn_apply_runs_kernel<T, kMode, 4, true>, grid, 256, 0, st, 
            static_cast<const T*>(x), nvec, cvr, mean, is_or_var, eps, gamma, beta,
            static_cast<const T*>(residual), act, alpha, static_cast<T*>(y), fs, rpb);
      else
        ::mcn::launch(bn_apply_runs_kernel<T, kMode, 8, false>, grid, 256, 0, st, 
            static_cast<const T*>(x), nvec, cvr, mean, is_or_var, eps, gamma, beta, nullptr, act, alpha,
            static_cast<T*>(y), fs, rpb);
    } else if (plan<T>(rows, C, &L, 2)) {
      ::mcn::launch(bn_apply_kernel<T, kMode>, L.grid, L.block, 0, st, 
          static_cast<const T*>(x), rows * L.cv, L.cv, mean, is_or_var, eps, gamma, beta,
          static_cast<const T*>(residual), act, alpha, static_cast<T*>(y), fs);
    } else {
      // odd channel counts: finalize as its own launch, then the scalar kernel
      if (kMode == 2) {
        ::mcn::launch(bn_finalize_kernel, (C + 127) / 128, 128, 0, st, fs.sums, 1.0 / fs.inv_count, C, fs.eps,
                                                            fs.momentum, fs.save_mean, fs.save_invstd,
                                                            fs.moving_mean, fs.moving_var);
        mean = fs.save_mean;
        is_or_var = fs.save_invstd;
      }
      long long n = rows * C;
      int grid = (int)std::min<long long>((n + 255) / 256, 8LL * num_sms());
      ::mcn::launch(bn_apply_scalar_kernel<T, kMode == 1>, grid, 256, 0, st, 
          static_cast<const T*>(x), n, C, mean, is_or_var, eps, gamma, beta,
          static_cast<const T*>(residual), act, alpha, static_cast<T*>(y));
    }
  });
  return after_launch("bn_apply");
}

extern "C" int mcn_bn_apply(int dtype, const void* x, long long rows, int C, const float* mean,
                            const float* invstd, const float* gamma, const float* beta,
                            const void* residual, int act, float act_alpha, void* y,
                            void* stream) {
  return bn_apply_impl<0>(dtype, x, rows, C, mean, invstd, 0.f, gamma, beta, residual, act,
                          act_alpha, y, BnSumsArgs{}, stream);
}
extern "C" int mcn_bn_infer(int dtype, const void* x, long long rows, int C, const float* mean,
                            const float* var, float eps, const float* gamma, const float* beta,
                            const void* residual, int act, float act_alpha, void* y,
                            void* stream) {
  return bn_apply_impl<1>(dtype, x, rows, C, mean, var, eps, gamma, beta, residual, act,
                          act_alpha, y, BnSumsArgs{}, stream);
}
extern "C" int mcn_bn_apply_stats(int dtype, const void* x, long long rows, int C,
                                  const double* sums, double count, float eps, float momentum,
                                  const float* gamma, const float* beta, const void* residual,
                                  int act, float act_alpha, void* y, float* save_mean,
                                  float* save_invstd, float* moving_mean, float* moving_var,
                                  void* stream) {
  MCN_REQUIRE(sums && save_mean && save_invstd && count > 0, "bn_apply_stats: bad argument");
  BnSumsArgs fs;
  fs.sums = sums;
  fs.inv_count = 1.0 / count;
  fs.bessel = count > 1.0 ? count / (count - 1.0) : 1.0;
  fs.eps = eps;
  fs.momentum = momentum;
  fs.save_mean = save_mean;
  fs.save_invstd = save_invstd;
  fs.moving_mean = moving_mean;
  fs.moving_var = moving_var;
  return bn_apply_impl<2>(dtype, x, rows, C, nullptr, nullptr, eps, gamma, beta, residual, act,
                          act_alpha, y, fs, stream);
}

// As mcn_bn_apply_stats, and also writes the ReLU bit mask of the stored output (bit e & 7 of byte
// e >> 3 <=> y[e] > 0; rows*C/8 bytes, rounded up to a multiple of 4) for mcn_bn_bwd_reduce_mask /
// mcn_bn_bwd_apply_mask: the layers with a fused residual need the sign of y in both backward passes,
// and 1 bit per element replaces two 2-byte reads.
extern "C" int mcn_bn_apply_stats_mask(int dtype, const void* x, long long rows, int C,
                                       const double* sums, double count, float eps, float momentum,
                                       const float* gamma, const float* beta, const void* residual,
                                       int act, float act_alpha, void* y, void* relu_mask,
                                       float* save_mean, float* save_invstd, float* moving_mean,
                                       float* moving_var, void* stream) {
  MCN_REQUIRE(sums && save_mean && save_invstd && count > 0 && relu_mask, "bn_apply_stats_mask: bad argument");
  MCN_REQUIRE(dtype == MCN_BF16 && act == MCN_ACT_RELU && bn_use_pipe(),
              "bn_apply_stats_mask: bf16 tensors with a ReLU only");
  BnSumsArgs fs;
  fs.sums = sums;
  fs.inv_count = 1.0 / count;
  fs.bessel = count > 1.0 ? count / (count - 1.0) : 1.0;
  fs.eps = eps;
  fs.momentum = momentum;
  fs.save_mean = save_mean;
  fs.save_invstd = save_invstd;
  fs.moving_mean = moving_mean;
  fs.moving_var = moving_var;
  return bn_apply_impl<2>(dtype, x, rows, C, nullptr, nullptr, eps, gamma, beta, residual, act,
                          act_alpha, y, fs, stream, static_cast<uint8_t*>(relu_mask));
}

template <typename T>
static void launch_bn_bwd_reduce(const SlabLaunch& L, size_t smem, cudaStream_t st, const void* dy, const void* x,
                                 const void* y, long long rows, int C, const float* mean, const float* invstd,
                                 const float* gamma, const float* beta, int act, float alpha, float* sum_dz,
                                 float* sum_dz_xhat, const XsScratch& xsc) {
  const T* pdy = static_cast<const T*>(dy);
  const T* px = static_cast<const T*>(x);
  const T* py = static_cast<const T*>(y);
  if (bn_reduce_pipe() && xsc.limbs != nullptr) {
    static bool ok_y = pipe_smem_ok(bn_reduce_pipe_kernel<T, false, true, float>, kPipeBytes + 16384);
    static bool ok_n = pipe_smem_ok(bn_reduce_pipe_kernel<T, false, false, float>, kPipeBytes + 16384);
    if (y != nullptr && ok_y) {
      ::mcn::launch(bn_reduce_pipe_kernel<T, false, true, float>, L.grid, 256, kPipeBytes + smem, st, pdy, px, py, rows,
                    C, L.slab_v, L.rowlanes, mean, invstd, gamma, beta, act, alpha, sum_dz, sum_dz_xhat, xsc);
      return;
    }
    if (y == nullptr && ok_n) {
      ::mcn::launch(bn_reduce_pipe_kernel<T, false, false, float>, L.grid, 256, kPipeBytes + smem, st, pdy, px, py, rows,
                    C, L.slab_v, L.rowlanes, mean, invstd, gamma, beta, act, alpha, sum_dz, sum_dz_xhat, xsc);
      return;
    }
  }
  const bool four = (y != nullptr) ? bn_bwd_unroll() == 4 : bn_noy_unroll() == 4;
  if (four && y != nullptr)
    ::mcn::launch(bn_bwd_reduce_kernel<T, 4, true>, L.grid, 256, smem, st, pdy, px, py, rows, C, L.slab_v, L.rowlanes, mean, invstd,
                                                               gamma, beta, act, alpha, sum_dz, sum_dz_xhat, xsc);
  else if (four && act == MCN_ACT_RELU)
    ::mcn::launch(bn_bwd_reduce_kernel<T, 4, false, false, MCN_ACT_RELU>, L.grid, 256, smem, st, pdy, px, py, rows, C,
                  L.slab_v, L.rowlanes, mean, invstd, gamma, beta, act, alpha, sum_dz, sum_dz_xhat, xsc);
  else if (four && act == MCN_ACT_NONE)
    ::mcn::launch(bn_bwd_reduce_kernel<T, 4, false, false, MCN_ACT_NONE>, L.grid, 256, smem, st, pdy, px, py, rows, C,
                  L.slab_v, L.rowlanes, mean, invstd, gamma, beta, act, alpha, sum_dz, sum_dz_xhat, xsc);
  else if (four)
    ::mcn::launch(bn_bwd_reduce_kernel<T, 4, false>, L.grid, 256, smem, st, pdy, px, py, rows, C, L.slab_v, L.rowlanes, mean, invstd,
                                                                gamma, beta, act, alpha, sum_dz, sum_dz_xhat, xsc);
  else if (y != nullptr)
    ::mcn::launch(bn_bwd_reduce_kernel<T, 2, true>, L.grid, 256, smem, st, pdy, px, py, rows, C, L.slab_v, L.rowlanes, mean, invstd,
                                                               gamma, beta, act, alpha, sum_dz, sum_dz_xhat, xsc);
  else
    ::mcn::launch(bn_bwd_reduce_kernel<T, 2, false>, L.grid, 256, smem, st, pdy, px, py, rows, C, L.slab_v, L.rowlanes, mean, invstd,
                                                                gamma, beta, act, alpha, sum_dz, sum_dz_xhat, xsc);
}

extern "C" int mcn_bn_bwd_reduce(int dtype, const void* dy, const void* x, const void* y,
                                 long long rows, int C, const float* mean, const float* invstd,
                                 const float* gamma, const float* beta, int act, float act_alpha,
                                 float* sum_dz, float* sum_dz_xhat, void* stream) {
  MCN_REQUIRE(dy && x && mean && invstd && sum_dz && sum_dz_xhat, "bn_bwd_reduce: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XsScratch xsc = xs_scratch(2 * C, "bn_bwd_reduce");
  if (xsc.limbs == nullptr) return MCN_EINVAL;
  MCN_DISPATCH_DTYPE(dtype, T, {
    SlabLaunch L;
    if (plan_slab<T>(rows, C, &L, 3 * num_sms())) {
      size_t smem = 2 * (size_t)L.rowlanes * L.slab_v * Vec16<T>::N * sizeof(float);
      launch_bn_bwd_reduce<T>(L, smem, st, dy, x, y, rows, C, mean, invstd, gamma, beta, act, act_alpha,
                              sum_dz, sum_dz_xhat, xsc);
    } else {
      MCN_REQUIRE((C + 127) / 128 <= kWsCounters, "bn_bwd_reduce: too many channels (%d)", C);
      dim3 grid((C + 127) / 128, (unsigned)std::min<long long>(rows, 4LL * num_sms()));
      ::mcn::launch(bn_bwd_reduce_scalar_kernel<T>, grid, 128, 0, st, 
          static_cast<const T*>(dy), static_cast<const T*>(x), static_cast<const T*>(y), rows, C,
          mean, invstd, gamma, beta, act, act_alpha, sum_dz, sum_dz_xhat, xsc);
    }
  });
  return after_launch("bn_bwd_reduce");
}

extern "C" int mcn_bn_bwd_reduce_mask(int dtype, const void* dy, const void* x, const void* relu_mask,
                                      long long rows, int C, const float* mean, const float* invstd,
                                      float* sum_dz, float* sum_dz_xhat, void* stream) {
  MCN_REQUIRE(dy && x && relu_mask && mean && invstd && sum_dz && sum_dz_xhat, "bn_bwd_reduce_mask: bad argument");
  MCN_REQUIRE(dtype == MCN_BF16 && C % 8 == 0, "bn_bwd_reduce_mask: bf16 tensors with C %% 8 == 0 only");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  XsScratch xsc = xs_scratch(2 * C, "bn_bwd_reduce_mask");
  if (xsc.limbs == nullptr) return MCN_EINVAL;
  typedef __nv_bfloat16 T;
  SlabLaunch L;
  MCN_REQUIRE(plan_slab<T>(rows, C, &L, 3 * num_sms()), "bn_bwd_reduce_mask: no slab plan for %lld x %d", rows, C);
  const size_t smem = 2 * (size_t)L.rowlanes * L.slab_v * Vec16<T>::N * sizeof(float);
  const T* py = static_cast<const T*>(relu_mask);
  const float* nul = nullptr;
  // the mask variant has two 16-byte streams like the y-less one: same unroll choice
  if (bn_noy_unroll() == 4)
    ::mcn::launch(bn_bwd_reduce_kernel<T, 4, true, true>, L.grid, 256, smem, st, static_cast<const T*>(dy),
                  static_cast<const T*>(x), py, rows, C, L.slab_v, L.rowlanes, mean, invstd, nul, nul,
                  (int)MCN_ACT_RELU, 0.f, sum_dz, sum_dz_xhat, xsc);
  else
    ::mcn::launch(bn_bwd_reduce_kernel<T, 2, true, true>, L.grid, 256, smem, st, static_cast<const T*>(dy),
                  static_cast<const T*>(x), py, rows, C, L.slab_v, L.rowlanes, mean, invstd, nul, nul,
                  (int)MCN_ACT_RELU, 0.f, sum_dz, sum_dz_xhat, xsc);
  return after_launch("bn_bwd_reduce_mask");
}

// Backward sums taken in a dgrad epilogue (mcn_conv2d_dgrad_tc_bnred) -> the two vectors the backward
// apply pass (and dbeta / dgamma) want: sum_dz += S1, sum_dz_xhat += invstd * (S2 - mean * S1), in fp64.
__global__ void bn_bwd_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ mean,
                                       const float* __restrict__ invstd, int C, float* __restrict__ sum_dz,
                                       float* __restrict__ sum_dz_xhat) {
  MCN_PDL_PROLOGUE();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double s1 = sums[c], s2 = sums[C + c];
  sum_dz[c] += static_cast<float>(s1);
  sum_dz_xhat[c] += static_cast<float>(static_cast<double>(invstd[c]) * (s2 - static_cast<double>(mean[c]) * s1));
}

extern "C" int mcn_bn_bwd_finalize(const double* sums, const float* mean, const float* invstd, int C,
                                   float* sum_dz, float* sum_dz_xhat, void* stream) {
  MCN_REQUIRE(sums && mean && invstd && sum_dz && sum_dz_xhat && C > 0, "bn_bwd_finalize: bad argument");
  ::mcn::launch(bn_bwd_finalize_kernel, (C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), sums, mean,
                invstd, C, sum_dz, sum_dz_xhat);
  return after_launch("bn_bwd_finalize");
}

template <typename T>
static void launch_bwd_apply_runs(unsigned grid, cudaStream_t st, const void* dy, const void* x,
                                  const void* y, long long nvec, int cv, const float* mean,
                                  const float* invstd, const float* gamma, const float* beta, int act,
                                  float alpha, const float* sum_dz, const float* sum_dz_xhat,
                                  float inv_count, void* dx, void* d_residual, int rpb) {
  const T* pdy = static_cast<const T*>(dy);
  const T* px = static_cast<const T*>(x);
  const T* py = static_cast<const T*>(y);
  T* pdx = static_cast<T*>(dx);
  T* pdr = static_cast<T*>(d_residual);
  if (y != nullptr && d_residual != nullptr)
    ::mcn::launch(bn_bwd_apply_runs_kernel<T, 4, true, true>, grid, 256, 0, st, 
        pdy, px, py, nvec, cv, mean, invstd, gamma, beta, act, alpha, sum_dz, sum_dz_xhat, inv_count, pdx, pdr, rpb);
  else if (y != nullptr)
    ::mcn::launch(bn_bwd_apply_runs_kernel<T, 4, true, false>, grid, 256, 0, st, 
        pdy, px, py, nvec, cv, mean, invstd, gamma, beta, act, alpha, sum_dz, sum_dz_xhat, inv_count, pdx, pdr, rpb);
  else if (d_residual != nullptr)
    ::mcn::launch(bn_bwd_apply_runs_kernel<T, 4, false, true>, grid, 256, 0, st, 
        pdy, px, py, nvec, cv, mean, invstd, gamma, beta, act, alpha, sum_dz, sum_dz_xhat, inv_count, pdx, pdr, rpb);
  else
    ::mcn::launch(bn_bwd_apply_runs_kernel<T, 4, false, false>, grid, 256, 0, st, 
        pdy, px, py, nvec, cv, mean, invstd, gamma, beta, act, alpha, sum_dz, sum_dz_xhat, inv_count, pdx, pdr, rpb);
}

#define MCN_BWD_APPLY_PIPE(Y_, R_)                                                                         \
  ::mcn::launch(bn_bwd_apply_pipe_kernel<T, (Y_) ? 1 : 0, R_>, grid, 256, kPipeBytes, st, static_cast<const T*>(dy), \
                static_cast<const T*>(x), static_cast<const T*>(y), nvec, cvr, mean, invstd, gamma, beta,  \
                act, act_alpha, sum_dz, sum_dz_xhat, inv_count, static_cast<T*>(dx),                       \
                static_cast<T*>(d_residual))

#define MCN_BWD_APPLY(U_, Y_)                                                                      \
  ::mcn::launch(bn_bwd_apply_kernel<T, U_, Y_>, L.grid, L.block, 0, st,                                        \
      static_cast<const T*>(dy), static_cast<const T*>(x), static_cast<const T*>(y), rows * L.cv,  \
      L.cv, mean, invstd, gamma, beta, act, act_alpha, sum_dz, sum_dz_xhat, inv_count,             \
      static_cast<T*>(dx), static_cast<T*>(d_residual))

extern "C" int mcn_bn_bwd_apply_mask(int dtype, const void* dy, const void* x, const void* relu_mask,
                                     long long rows, int C, const float* mean, const float* invstd,
                                     const float* gamma, const float* sum_dz, const float* sum_dz_xhat,
                                     double count, void* dx, void* d_residual, void* stream) {
  MCN_REQUIRE(dy && x && relu_mask && dx && mean && invstd && sum_dz && sum_dz_xhat && count > 0,
              "bn_bwd_apply_mask: bad argument");
  typedef __nv_bfloat16 T;
  const int cvr = (C % 8 == 0) ? C / 8 : 0;
  MCN_REQUIRE(dtype == MCN_BF16 && cvr > 0 && cvr <= 256 && bn_use_pipe(),
              "bn_bwd_apply_mask: bf16 tensors with C %% 8 == 0 and C <= 2048 only");
  MCN_REQUIRE(reinterpret_cast<uintptr_t>(relu_mask) % 4 == 0, "bn_bwd_apply_mask: the mask must be 4-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float inv_count = (float)(1.0 / count);
  const long long nvec = rows * cvr;
  const int nthr = (256 / cvr) * cvr;
  const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((nvec + nthr - 1) / nthr, 3LL * num_sms()));
  static bool ok_r = pipe_smem_ok(bn_bwd_apply_pipe_kernel<T, 2, true>, kPipeBytes + kMaskRingBytes);
  static bool ok_n = pipe_smem_ok(bn_bwd_apply_pipe_kernel<T, 2, false>, kPipeBytes + kMaskRingBytes);
  MCN_REQUIRE(ok_r && ok_n, "bn_bwd_apply_mask: cannot opt in to %d bytes of shared memory", kPipeBytes + kMaskRingBytes);
  const float* nul = nullptr;
  if (d_residual != nullptr)
    ::mcn::launch(bn_bwd_apply_pipe_kernel<T, 2, true>, grid, 256, kPipeBytes + kMaskRingBytes, st,
                  static_cast<const T*>(dy), static_cast<const T*>(x), static_cast<const T*>(relu_mask), nvec, cvr,
                  mean, invstd, gamma, nul, (int)MCN_ACT_RELU, 0.f, sum_dz, sum_dz_xhat, inv_count,
                  static_cast<T*>(dx), static_cast<T*>(d_residual));
  else
    ::mcn::launch(bn_bwd_apply_pipe_kernel<T, 2, false>, grid, 256, kPipeBytes + kMaskRingBytes, st,
                  static_cast<const T*>(dy), static_cast<const T*>(x), static_cast<const T*>(relu_mask), nvec, cvr,
                  mean, invstd, gamma, nul, (int)MCN_ACT_RELU, 0.f, sum_dz, sum_dz_xhat, inv_count,
                  static_cast<T*>(dx), static_cast<T*>(nullptr));
  return after_launch("bn_bwd_apply_mask");
}

extern "C" int mcn_bn_bwd_apply(int dtype, const void* dy, const void* x, const void* y,
                                long long rows, int C, const float* mean, const float* invstd,
                                const float* gamma, const float* beta, int act, float act_alpha,
                                const float* sum_dz, const float* sum_dz_xhat, double count,
                                void* dx, void* d_residual, void* stream) {
  MCN_REQUIRE(dy && x && dx && mean && invstd && sum_dz && sum_dz_xhat && count > 0,
              "bn_bwd_apply: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float inv_count = (float)(1.0 / count);
  MCN_DISPATCH_DTYPE(dtype, T, {
    ChanLaunch L;
    constexpr int V = Vec16<T>::N;
    const int cvr = (C % V == 0) ? C / V : 0;
    if (cvr > 0 && cvr <= 256 && bn_use_pipe()) {
      const long long nvec = rows * cvr;
      const int nthr = (256 / cvr) * cvr;
      const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>((nvec + nthr - 1) / nthr, 3LL * num_sms()));
      auto go = [&](auto kern) {
        ::mcn::launch(kern, grid, 256, kPipeBytes, st, static_cast<const T*>(dy), static_cast<const T*>(x),
                      static_cast<const T*>(y), nvec, cvr, mean, invstd, gamma, beta, act, act_alpha, sum_dz,
                      sum_dz_xhat, inv_count, static_cast<T*>(dx), static_cast<T*>(d_residual));
      };
      if (y != nullptr && d_residual != nullptr) MCN_BWD_APPLY_PIPE(true, true);
      else if (y != nullptr) MCN_BWD_APPLY_PIPE(true, false);
      else if (d_residual != nullptr) MCN_BWD_APPLY_PIPE(false, true);
      else if (act == MCN_ACT_RELU) go(bn_bwd_apply_pipe_kernel<T, 0, false, MCN_ACT_RELU>);
      else if (act == MCN_ACT_NONE) go(bn_bwd_apply_pipe_kernel<T, 0, false, MCN_ACT_NONE>);
      else MCN_BWD_APPLY_PIPE(false, false);
    } else if (cvr > 0 && cvr <= 256 && 256 % cvr == 0 && bn_bwd_use_runs()) {
      const long long nvec = rows * cvr;
      constexpr int U = 4;
      const long long runs = (nvec + 256LL * U - 1) / (256LL * U);
      int rpb = (int)std::max<long long>(1, std::min<long long>(8, runs / (4LL * 2 * num_sms())));
      const unsigned grid = (unsigned)((runs + rpb - 1) / rpb);
      launch_bwd_apply_runs<T>(grid, st, dy, x, y, nvec, cvr, mean, invstd, gamma, beta, act, act_alpha,
                               sum_dz, sum_dz_xhat, inv_count, dx, d_residual, rpb);
    } else if (plan<T>(rows, C, &L, 3, 256) && (sizeof(T) != 2 || L.block <= 256)) {
      // Bytes in flight decide these kernels (ncu: ~37 % of the warp slots, all stalled on loads): the
      // y-less variant has two read streams instead of three, so it keeps twice the vectors in flight
      // (768 threads x 4 x 2 x 16 B = 98 KB per SM, the same as the three-stream variant at U = 2).
      if (y != nullptr) {
        if (bn_bwd_unroll() == 4) MCN_BWD_APPLY(4, true);
        else MCN_BWD_APPLY(2, true);
      } else {
        if (bn_noy_unroll() == 2) MCN_BWD_APPLY(2, false);
        else MCN_BWD_APPLY(4, false);
      }
    } else {
      long long n = rows * C;
      int grid = (int)std::min<long long>((n + 255) / 256, 8LL * num_sms());
      ::mcn::launch(bn_bwd_apply_scalar_kernel<T>, grid, 256, 0, st, 
          static_cast<const T*>(dy), static_cast<const T*>(x), static_cast<const T*>(y), n, C,
          mean, invstd, gamma, beta, act, act_alpha, sum_dz, sum_dz_xhat, inv_count,
          static_cast<T*>(dx), static_cast<T*>(d_residual));
    }
  });
  return after_launch("bn_bwd_apply");
}
