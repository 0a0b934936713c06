// Group normalisation and weight standardisation.
// Replaces reference convnet.py:1928-2013 (group_norm: per-sample, per-group moments over
// H x W x C/G, then the per-channel affine) and convnet.py:1410-1419 (weight standardisation:
// w' = (w - mean_o) / (std_o + 1e-5) per output channel o over every other axis, applied inside the
// graph so the gradient flows back to the raw weights).  Neither is on the ResNet-50 hot path (the
// *_wsgn model files and the DCGAN 'gn' option use them), so the kernels are plain CUDA-core code
// built for exactness and run-to-run reproducibility: every reduction has a fixed order, no atomics.
#include "mcn_common.cuh"

namespace mcn {
namespace {

// Fixed-order block reduction of two doubles (blockDim.x == 256); result valid in every thread.
__device__ __forceinline__ void block_sum2(double& a, double& b) {
  __shared__ double sa[256], sb[256];
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sa[threadIdx.x] += sa[threadIdx.x + s];
      sb[threadIdx.x] += sb[threadIdx.x + s];
    }
    __syncthreads();
  }
  a = sa[0];
  b = sb[0];
  __syncthreads();
}

// ---------------------------------------------------------------- group norm
// grid (G, N): moments of group g of sample n.  save[(n*G+g)*2] = {mean, invstd}
template <typename T>
__global__ void __launch_bounds__(256)
gn_stats_kernel(const T* __restrict__ x, long long HW, int C, int G, float eps, float* __restrict__ save) {
  MCN_PDL_PROLOGUE();
  const int g = blockIdx.x, n = blockIdx.y, cg = C / G;
  const long long m = HW * cg;
  const T* base = x + (long long)n * HW * C + (long long)g * cg;
  double s1 = 0.0, s2 = 0.0;
  for (long long e = threadIdx.x; e < m; e += 256) {
    const long long p = e / cg;
    const int c = (int)(e - p * cg);
    const float v = to_f32(base[p * C + c]);
    s1 += v;
    s2 += (double)v * v;
  }
  block_sum2(s1, s2);
  if (threadIdx.x == 0) {
    const double mean = s1 / (double)m;
    double var = s2 / (double)m - mean * mean;
    if (var < 0.0) var = 0.0;
    save[((long long)n * G + g) * 2] = (float)mean;
    save[((long long)n * G + g) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
}

template <typename T>
__global__ void gn_apply_kernel(const T* __restrict__ x, long long total, long long HW, int C, int G,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const float* __restrict__ save, T* __restrict__ y) {
  MCN_PDL_PROLOGUE();
  const int cg = C / G;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long n = i / (HW * C);
    const float* sv = save + (n * G + c / cg) * 2;
    const float xh = (to_f32(x[i]) - sv[0]) * sv[1];
    y[i] = from_f32<T>(xh * (gamma ? gamma[c] : 1.f) + (beta ? beta[c] : 0.f));
  }
}

// grid (G, N): red[(n*G+g)*2] = {sum dy*gamma, sum dy*gamma*xhat} over the group
template <typename T>
__global__ void __launch_bounds__(256)
gn_bwd_group_kernel(const T* __restrict__ dy, const T* __restrict__ x, long long HW, int C, int G,
                    const float* __restrict__ gamma, const float* __restrict__ save,
                    float* __restrict__ red) {
  MCN_PDL_PROLOGUE();
  const int g = blockIdx.x, n = blockIdx.y, cg = C / G;
  const long long m = HW * cg;
  const long long off = (long long)n * HW * C + (long long)g * cg;
  const float mean = save[((long long)n * G + g) * 2], is = save[((long long)n * G + g) * 2 + 1];
  double s1 = 0.0, s2 = 0.0;
  for (long long e = threadIdx.x; e < m; e += 256) {
    const long long p = e / cg;
    const int c = (int)(e - p * cg);
    const long long i = off + p * C + c;
    const float dg = to_f32(dy[i]) * (gamma ? gamma[g * cg + c] : 1.f);
    const float xh = (to_f32(x[i]) - mean) * is;
    s1 += dg;
    s2 += (double)dg * xh;
  }
  block_sum2(s1, s2);
  if (threadIdx.x == 0) {
    red[((long long)n * G + g) * 2] = (float)s1;
    red[((long long)n * G + g) * 2 + 1] = (float)s2;
  }
}

// grid (ceil(C/32), N), block 256 = 32 channels x 8 pixel lanes:
// part[(n*C + c)*2] = {sum_hw dy*xhat, sum_hw dy} of sample n, summed in a fixed order
template <typename T>
__global__ void __launch_bounds__(256)
gn_bwd_channel_kernel(const T* __restrict__ dy, const T* __restrict__ x, long long HW, int C, int G,
                      const float* __restrict__ save, float* __restrict__ part) {
  MCN_PDL_PROLOGUE();
  __shared__ float sh1[8][33], sh2[8][33];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx, n = blockIdx.y, cg = C / G;
  float a1 = 0.f, a2 = 0.f;
  if (c < C) {
    const float mean = save[((long long)n * G + c / cg) * 2], is = save[((long long)n * G + c / cg) * 2 + 1];
    const long long off = (long long)n * HW * C + c;
    for (long long p = py; p < HW; p += 8) {
      const float d = to_f32(dy[off + p * C]);
      a1 = fmaf(d, (to_f32(x[off + p * C]) - mean) * is, a1);
      a2 += d;
    }
  }
  sh1[py][cx] = a1;
  sh2[py][cx] = a2;
  __syncthreads();
  if (py == 0 && c < C) {
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      t1 += sh1[k][cx];
      t2 += sh2[k][cx];
    }
    part[((long long)n * C + c) * 2] = t1;
    part[((long long)n * C + c) * 2 + 1] = t2;
  }
}

// dgamma[c] += sum_n part[n][c][0], dbeta[c] += sum_n part[n][c][1]   (sample order)
__global__ void gn_param_sum_kernel(const float* __restrict__ part, int N, int C, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta) {
  MCN_PDL_PROLOGUE();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float t1 = 0.f, t2 = 0.f;
  for (int n = 0; n < N; ++n) {
    t1 += part[((long long)n * C + c) * 2];
    t2 += part[((long long)n * C + c) * 2 + 1];
  }
  if (dgamma) dgamma[c] += t1;
  if (dbeta) dbeta[c] += t2;
}

// dx = invstd * (dy*gamma - s1/m - xhat * s2/m)
template <typename T>
__global__ void gn_bwd_apply_kernel(const T* __restrict__ dy, const T* __restrict__ x, long long total,
                                    long long HW, int C, int G, const float* __restrict__ gamma,
                                    const float* __restrict__ save, const float* __restrict__ red,
                                    T* __restrict__ dx) {
  MCN_PDL_PROLOGUE();
  const int cg = C / G;
  const float inv_m = 1.f / (float)(HW * cg);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const long long ng = (i / (HW * C)) * G + c / cg;
    const float mean = save[ng * 2], is = save[ng * 2 + 1];
    const float xh = (to_f32(x[i]) - mean) * is;
    const float dg = to_f32(dy[i]) * (gamma ? gamma[c] : 1.f);
    dx[i] = from_f32<T>(is * (dg - red[ng * 2] * inv_m - xh * red[ng * 2 + 1] * inv_m));
  }
}

// ---------------------------------------------------------------- weight standardisation
// w: [rows][cols] fp32, one statistic per COLUMN (output channel).  grid ceil(cols/32), block
// 256 = 32 columns x 8 row lanes; sums in fp64, fixed order.
__device__ __forceinline__ double col_sum8(double v, double (*sh)[33], int cx, int ry) {
  sh[ry][cx] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int k = 0; k < 8; ++k) t += sh[k][cx];
  __syncthreads();
  return t;
}

__global__ void __launch_bounds__(256)
ws_fwd_kernel(const float* __restrict__ w, int rows, int cols, float eps, float* __restrict__ w_std,
              float* __restrict__ stats) {
  MCN_PDL_PROLOGUE();
  __shared__ double sh[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const bool ok = c < cols;
  double s = 0.0;
  if (ok)
    for (int r = ry; r < rows; r += 8) s += w[(long long)r * cols + c];
  const double mean = col_sum8(s, sh, cx, ry) / rows;
  double q = 0.0;
  if (ok)
    for (int r = ry; r < rows; r += 8) {
      const double d = (double)w[(long long)r * cols + c] - mean;
      q += d * d;
    }
  const double sigma = sqrt(col_sum8(q, sh, cx, ry) / rows);
  if (!ok) return;
  const float fm = (float)mean, inv = (float)(1.0 / (sigma + (double)eps));
  for (int r = ry; r < rows; r += 8) w_std[(long long)r * cols + c] = (w[(long long)r * cols + c] - fm) * inv;
  if (ry == 0) {
    stats[c] = fm;
    stats[cols + c] = (float)sigma;
  }
}

// grad += (g - mean(g)) / s - c * sum(g*c) / (rows * sigma * s^2),  c = w - mean, s = sigma + eps
__global__ void __launch_bounds__(256)
ws_bwd_kernel(const float* __restrict__ g, const float* __restrict__ w, const float* __restrict__ stats,
              int rows, int cols, float eps, float* __restrict__ grad) {
  MCN_PDL_PROLOGUE();
  __shared__ double sh[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const bool ok = c < cols;
  const float mean = ok ? stats[c] : 0.f, sigma = ok ? stats[cols + c] : 1.f;
  double sg = 0.0, sgc = 0.0;
  if (ok)
    for (int r = ry; r < rows; r += 8) {
      const float gv = g[(long long)r * cols + c];
      sg += gv;
      sgc += (double)gv * (double)(w[(long long)r * cols + c] - mean);
    }
  const double tg = col_sum8(sg, sh, cx, ry), tgc = col_sum8(sgc, sh, cx, ry);
  if (!ok) return;
  const double s = (double)sigma + (double)eps;
  const float mg = (float)(tg / rows), inv_s = (float)(1.0 / s);
  const float k2 = sigma > 0.f ? (float)(tgc / ((double)rows * sigma * s * s)) : 0.f;
  for (int r = ry; r < rows; r += 8) {
    const long long i = (long long)r * cols + c;
    grad[i] += (g[i] - mg) * inv_s - (w[i] - mean) * k2;
  }
}

}  // namespace
}  // namespace mcn

using namespace mcn;

extern "C" int mcn_gn_fwd(int dtype, const void* x, int N, long long HW, int C, int G, float eps,
                          const float* gamma, const float* beta, void* y, float* save, void* stream) {
  MCN_REQUIRE(x && y && save && N > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0, "gn_fwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = (long long)N * HW * C;
  const int grid = (int)std::min<long long>((total + 255) / 256, 8LL * num_sms());
  MCN_DISPATCH_DTYPE(dtype, T, {
    ::mcn::launch(gn_stats_kernel<T>, dim3(G, N), 256, 0, st, static_cast<const T*>(x), HW, C, G, eps, save);
    int rc = after_launch("gn_stats");
    if (rc) return rc;
    ::mcn::launch(gn_apply_kernel<T>, grid, 256, 0, st, static_cast<const T*>(x), total, HW, C, G, gamma, beta,
                  static_cast<const float*>(save), static_cast<T*>(y));
  });
  return after_launch("gn_apply");
}

extern "C" int mcn_gn_bwd(int dtype, const void* dy, const void* x, int N, long long HW, int C, int G,
                          const float* gamma, const float* save, float* scratch, void* dx, float* dgamma,
                          float* dbeta, void* stream) {
  MCN_REQUIRE(dy && x && save && scratch && N > 0 && HW > 0 && C > 0 && G > 0 && C % G == 0,
              "gn_bwd: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long total = (long long)N * HW * C;
  const int grid = (int)std::min<long long>((total + 255) / 256, 8LL * num_sms());
  float* red = scratch;                              // [N*G*2]
  float* part = scratch + (long long)N * G * 2;      // [N*C*2]
  MCN_DISPATCH_DTYPE(dtype, T, {
    int rc;
    if (dx != nullptr) {
      ::mcn::launch(gn_bwd_group_kernel<T>, dim3(G, N), 256, 0, st, static_cast<const T*>(dy),
                    static_cast<const T*>(x), HW, C, G, gamma, save, red);
      if ((rc = after_launch("gn_bwd_group"))) return rc;
      ::mcn::launch(gn_bwd_apply_kernel<T>, grid, 256, 0, st, static_cast<const T*>(dy), static_cast<const T*>(x),
                    total, HW, C, G, gamma, save, static_cast<const float*>(red), static_cast<T*>(dx));
      if ((rc = after_launch("gn_bwd_apply"))) return rc;
    }
    if (dgamma != nullptr || dbeta != nullptr) {
      ::mcn::launch(gn_bwd_channel_kernel<T>, dim3((C + 31) / 32, N), 256, 0, st, static_cast<const T*>(dy),
                    static_cast<const T*>(x), HW, C, G, save, part);
      if ((rc = after_launch("gn_bwd_channel"))) return rc;
      ::mcn::launch(gn_param_sum_kernel, (C + 127) / 128, 128, 0, st, static_cast<const float*>(part), N, C, dgamma,
                    dbeta);
      if ((rc = after_launch("gn_param_sum"))) return rc;
    }
  });
  return MCN_OK;
}

extern "C" int mcn_ws_fwd(const float* w, int rows, int cols, float eps, float* w_std, float* stats,
                          void* stream) {
  MCN_REQUIRE(w && w_std && stats && rows > 0 && cols > 0, "ws_fwd: bad argument");
  ::mcn::launch(ws_fwd_kernel, (cols + 31) / 32, 256, 0, static_cast<cudaStream_t>(stream), w, rows, cols, eps,
                w_std, stats);
  return after_launch("ws_fwd");
}

extern "C" int mcn_ws_bwd(const float* g_std, const float* w, const float* stats, int rows, int cols,
                          float eps, float* grad, void* stream) {
  MCN_REQUIRE(g_std && w && stats && grad && rows > 0 && cols > 0, "ws_bwd: bad argument");
  ::mcn::launch(ws_bwd_kernel, (cols + 31) / 32, 256, 0, static_cast<cudaStream_t>(stream), g_std, w, stats, rows,
                cols, eps, grad);
  return after_launch("ws_bwd");
}
