// Loss kernels: fused forward + gradient.
// softmax cross-entropy replaces tf.nn.softmax_cross_entropy_with_logits_v2 + the weighting and
// masking of reference convnet.py:552-594 (mean over ALL rows, invalid rows contribute zero);
// sigmoid cross-entropy replaces tf.nn.sigmoid_cross_entropy_with_logits at gan.py:134-136.
#include <cfloat>

#include "mcn_common.cuh"
#include "xsum.cuh"

namespace mcn {
namespace {

// One warp per row.  labels[r] < 0 (or >= C) is the reference's all-zero one-hot row: the row is
// invalid, its loss and gradient are zero (convnet.py:448-449, 567-573).
//
// Targets t_c (convnet.py:574-577, 603-607; segmentation/segnet.py:116-121):
//   seg_h == 0 : t = onehot*(1-ls) + ls/C                          (classification)
//   seg_h  > 0 : t = onehot*(1-ls) + ls*avg5x5(onehot)             (rows are pixels of [N,seg_h,seg_w]
//                maps; tf.nn.avg_pool2d SAME divides by the in-bounds window size, ignored pixels
//                are zero rows of the one-hot map) — computed on the fly from the label map.
// CE = -sum_c t_c log p_c = lse*sum(t) - sum_c t_c z_c (targets need not sum to one).
// Focal variants (convnet.py:580-592), p_y = softmax probability of the TRUE class:
//   focal_gamma > 0 : CE *= (1 - p_y)^gamma                         (differentiated through)
//   sig_alpha   > 0 : CE *= stop_gradient(1 - sigmoid(alpha*(p_y - 0.5))) / (1 - sigmoid(-alpha/2))
struct XentOpts {
  float ls, focal_gamma, sig_alpha;
  int seg_h, seg_w;
};

__device__ __forceinline__ float xent_target(const XentOpts& o, const int32_t* __restrict__ labels,
                                             long long r, int y, int c, int C) {
  const float hot = (c == y) ? 1.f - o.ls : 0.f;
  if (o.ls <= 0.f) return hot;
  if (o.seg_h == 0) return hot + o.ls / static_cast<float>(C);
  const int w = static_cast<int>(r % o.seg_w);
  const long long q = r / o.seg_w;
  const int h = static_cast<int>(q % o.seg_h);
  const long long img = (q / o.seg_h) * o.seg_h * o.seg_w;
  int cnt = 0, nvalid = 0;
  for (int dy = -2; dy <= 2; ++dy) {
    const int hh = h + dy;
    if (hh < 0 || hh >= o.seg_h) continue;
    for (int dx = -2; dx <= 2; ++dx) {
      const int ww = w + dx;
      if (ww < 0 || ww >= o.seg_w) continue;
      ++nvalid;
      cnt += (labels[img + static_cast<long long>(hh) * o.seg_w + ww] == c) ? 1 : 0;
    }
  }
  return hot + o.ls * static_cast<float>(cnt) / static_cast<float>(nvalid);
}

// Segmentation targets for a whole row at once: lane j < 25 holds the label at position j of the
// row's 5x5 window (-2 outside the map), `nvalid` the in-bounds count; the smoothed target of class c
// is then 25 shuffles + compares.  (The per-class version above re-derives the pixel coordinates with
// 64-bit divisions and re-reads the 25 labels for every class and twice per row: the DeepLab loss
// launch took 2.5 ms for 4.2 M pixels x 21 classes.)
struct SegWindow {
  int label;     // this lane's window label
  int nvalid;
};
__device__ __forceinline__ SegWindow seg_window(const XentOpts& o, const int32_t* __restrict__ labels,
                                                long long r, int lane) {
  const int w = static_cast<int>(r % o.seg_w);
  const long long q = r / o.seg_w;
  const int h = static_cast<int>(q % o.seg_h);
  const long long img = (q / o.seg_h) * o.seg_h * o.seg_w;
  const int dy = lane / 5 - 2, dx = lane - (lane / 5) * 5 - 2;
  const int hh = h + dy, ww = w + dx;
  const bool inb = lane < 25 && hh >= 0 && hh < o.seg_h && ww >= 0 && ww < o.seg_w;
  SegWindow s;
  s.label = inb ? labels[img + static_cast<long long>(hh) * o.seg_w + ww] : -2;
  s.nvalid = __popc(__ballot_sync(0xffffffffu, inb));
  return s;
}
// all 32 lanes must call this together (c may differ per lane)
__device__ __forceinline__ float seg_target(const XentOpts& o, const SegWindow& sw, int y, int c) {
  int cnt = 0;
#pragma unroll
  for (int j = 0; j < 25; ++j) cnt += (__shfl_sync(0xffffffffu, sw.label, j) == c) ? 1 : 0;
  const float hot = (c == y) ? 1.f - o.ls : 0.f;
  return hot + o.ls * static_cast<float>(cnt) / static_cast<float>(sw.nvalid);
}

__global__ void softmax_xent_kernel(const float* __restrict__ logits,
                                    const int32_t* __restrict__ labels, long long rows, int C,
                                    const float* __restrict__ class_w, XentOpts o, float grad_scale,
                                    long long* __restrict__ loss_xs, float* __restrict__ dlogits,
                                    float* __restrict__ probs) {
  MCN_PDL_PROLOGUE();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  float block_loss = 0.f;
  for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows;
       r += (long long)gridDim.x * wpb) {
    const float* z = logits + r * C;
    float mx = -FLT_MAX;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, s));
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += __expf(z[c] - mx);
    se = warp_sum(se);
    const float lse = mx + __logf(se);
    const int y = labels ? labels[r] : -1;
    const bool valid = (y >= 0 && y < C);
    const float w = valid ? (class_w ? class_w[y] : 1.f) : 0.f;
    const float inv_se = 1.f / se;
    float ce = 0.f, st = 1.f, f_mul = 1.f, df = 0.f, s_mul = 1.f, py = 0.f;
    const bool seg = o.ls > 0.f && o.seg_h > 0;     // warp-uniform
    SegWindow sw{-2, 1};
    if (seg && valid) sw = seg_window(o, labels, r, lane);
    if (valid) {
      // sum_c t_c and sum_c t_c z_c
      float a = 0.f, b = 0.f;
      if (o.ls > 0.f) {
        for (int c0 = 0; c0 < C; c0 += 32) {       // whole-warp trips: seg_target shuffles
          const int c = c0 + lane;
          const float t = seg ? seg_target(o, sw, y, c) : xent_target(o, labels, r, y, c, C);
          if (c < C) {
            a += t;
            b = fmaf(t, z[c], b);
          }
        }
        a = warp_sum(a);
        b = warp_sum(b);
      } else {
        a = 1.f;
        b = z[y];
      }
      st = a;
      ce = lse * a - b;
      py = __expf(z[y] - mx) * inv_se;
      if (o.focal_gamma > 0.f) {
        const float om = fmaxf(1.f - py, 1e-12f);
        f_mul = __powf(om, o.focal_gamma);
        df = -o.focal_gamma * __powf(om, o.focal_gamma - 1.f) * py;      // dF/dz_c = df * ((c==y) - p_c)
      }
      if (o.sig_alpha > 0.f)
        s_mul = (1.f - 1.f / (1.f + __expf(-o.sig_alpha * (py - 0.5f)))) /
                (1.f - 1.f / (1.f + __expf(0.5f * o.sig_alpha)));
      if (lane == 0) block_loss += w * ce * f_mul * s_mul;
    }
    for (int c0 = 0; c0 < C; c0 += 32) {
      const int c = c0 + lane;
      float t = 0.f;
      if (dlogits && valid) t = seg ? seg_target(o, sw, y, c) : xent_target(o, labels, r, y, min(c, C - 1), C);
      if (c >= C) continue;
      const float p = __expf(z[c] - mx) * inv_se;
      if (probs) probs[r * C + c] = p;
      if (dlogits) {
        float g = 0.f;
        if (valid) g = f_mul * (p * st - t) + ce * df * ((c == y ? 1.f : 0.f) - p);
        dlogits[r * C + c] = grad_scale * w * s_mul * g;
      }
    }
  }
  if (loss_xs && lane == 0) xs::add(loss_xs, 1, 0, block_loss);
}

// Few classes (C <= 48: segmentation heads, small classifiers): one THREAD per row.  A block stages
// 256 consecutive rows in shared memory with coalesced loads (row pitch C | 1 words: conflict-free),
// each thread walks its own row there — no shuffles, no half-empty warps — overwrites it with the
// output (probabilities or gradient) and the block writes the chunk back coalesced.  The warp-per-row
// kernel above kept 21 of 32 lanes busy behind five dependent passes and two shuffle reductions per
// row: 2.2 ms (probabilities) + 2.9 ms (loss + gradient) for the 4.2 M x 21 DeepLab head.
// Smoothed segmentation targets t_c = hot_c + ls*cnt_c/nvalid enter through their sums:
//   sum_c t_c = (1-ls) + ls*n_lab/nvalid,  sum_c t_c z_c = (1-ls) z_y + (ls/nvalid) sum_j z[l_j],
// and the gradient's -t_c term is scattered per window label into the thread's own row.
__global__ void __launch_bounds__(256)
softmax_xent_rows_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels, long long rows,
                         int C, const float* __restrict__ class_w, XentOpts o, float grad_scale,
                         long long* __restrict__ loss_xs, float* __restrict__ out, int out_is_grad) {
  MCN_PDL_PROLOGUE();
  extern __shared__ float rows_sm[];
  const int pitch = C | 1;
  float* mine = rows_sm + threadIdx.x * pitch;
  float my_loss = 0.f;
  const long long chunks = (rows + 255) / 256;
  for (long long ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
    const long long r0 = ch * 256;
    const int nrows = static_cast<int>(min(256LL, rows - r0));
    const int nel = nrows * C;
    const float* src = logits + r0 * C;
    for (int i = threadIdx.x; i < nel; i += 256) {
      const int rr = i / C;
      rows_sm[rr * pitch + (i - rr * C)] = __ldcs(src + i);
    }
    __syncthreads();
    if (threadIdx.x < nrows) {
      const long long r = r0 + threadIdx.x;
      float mx = -FLT_MAX;
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, mine[c]);
      float se = 0.f;
      for (int c = 0; c < C; ++c) se += __expf(mine[c] - mx);
      const float lse = mx + __logf(se);
      const float inv_se = 1.f / se;
      const int y = labels ? labels[r] : -1;
      const bool valid = (y >= 0 && y < C);
      const float w = valid ? (class_w ? class_w[y] : 1.f) : 0.f;
      float ce = 0.f, st = 1.f, f_mul = 1.f, df = 0.f, s_mul = 1.f, py = 0.f;
      const bool seg = o.ls > 0.f && o.seg_h > 0;
      int wl[25];
      int nvalid = 1;
      if (valid) {
        float a = 1.f, b = mine[y];
        if (seg) {
          const int wq = static_cast<int>(r % o.seg_w);
          const long long q = r / o.seg_w;
          const int h = static_cast<int>(q % o.seg_h);
          const int32_t* lab = labels + (q / o.seg_h) * o.seg_h * o.seg_w;
          nvalid = 0;
          int nlab = 0;
          float zs = 0.f;
#pragma unroll
          for (int j = 0; j < 25; ++j) {
            const int hh = h + j / 5 - 2, ww = wq + j % 5 - 2;
            const bool inb = hh >= 0 && hh < o.seg_h && ww >= 0 && ww < o.seg_w;
            int l = -2;
            if (inb) {
              l = lab[static_cast<long long>(hh) * o.seg_w + ww];
              ++nvalid;
            }
            if (l < 0 || l >= C) l = -2;
            wl[j] = l;
            if (l >= 0) {
              ++nlab;
              zs += mine[l];
            }
          }
          const float k = o.ls / static_cast<float>(nvalid);
          a = (1.f - o.ls) + k * static_cast<float>(nlab);
          b = (1.f - o.ls) * mine[y] + k * zs;
        } else if (o.ls > 0.f) {
          float zs = 0.f;
          for (int c = 0; c < C; ++c) zs += mine[c];
          const float k = o.ls / static_cast<float>(C);
          a = (1.f - o.ls) + k * static_cast<float>(C);
          b = (1.f - o.ls) * mine[y] + k * zs;
        }
        st = a;
        ce = lse * a - b;
        py = __expf(mine[y] - mx) * inv_se;
        if (o.focal_gamma > 0.f) {
          const float om = fmaxf(1.f - py, 1e-12f);
          f_mul = __powf(om, o.focal_gamma);
          df = -o.focal_gamma * __powf(om, o.focal_gamma - 1.f) * py;
        }
        if (o.sig_alpha > 0.f)
          s_mul = (1.f - 1.f / (1.f + __expf(-o.sig_alpha * (py - 0.5f)))) /
                  (1.f - 1.f / (1.f + __expf(0.5f * o.sig_alpha)));
        my_loss += w * ce * f_mul * s_mul;
      }
      // the row becomes the output
      const float gs = grad_scale * w * s_mul;
      const float base_t = (o.ls > 0.f && !seg) ? o.ls / static_cast<float>(C) : 0.f;
      for (int c = 0; c < C; ++c) {
        const float p = __expf(mine[c] - mx) * inv_se;
        if (!out_is_grad) {
          mine[c] = p;
        } else {
          float g = 0.f;
          if (valid) {
            const float hot = (c == y) ? 1.f - o.ls : 0.f;
            g = f_mul * (p * st - (hot + base_t)) + ce * df * ((c == y ? 1.f : 0.f) - p);
          }
          mine[c] = gs * g;
        }
      }
      if (out_is_grad && valid && seg) {
        const float k = gs * f_mul * o.ls / static_cast<float>(nvalid);
#pragma unroll
        for (int j = 0; j < 25; ++j)
          if (wl[j] >= 0) mine[wl[j]] -= k;
      }
    }
    __syncthreads();
    if (out != nullptr) {
      float* dst = out + r0 * C;
      for (int i = threadIdx.x; i < nel; i += 256) {
        const int rr = i / C;
        dst[i] = rows_sm[rr * pitch + (i - rr * C)];
      }
    }
    __syncthreads();
  }
  my_loss = warp_sum(my_loss);
  if (loss_xs && (threadIdx.x & 31) == 0) xs::add(loss_xs, 1, 0, my_loss);
}

// loss = max(x,0) - x*z + log1p(exp(-|x|));  d/dx = sigmoid(x) - z
__global__ void sigmoid_xent_kernel(const float* __restrict__ logits, long long n, float label,
                                    float weight, float grad_scale, long long* __restrict__ loss_xs,
                                    float* __restrict__ dlogits, int accumulate) {
  MCN_PDL_PROLOGUE();
  float acc = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float x = logits[i];
    acc += weight * (fmaxf(x, 0.f) - x * label + log1pf(__expf(-fabsf(x))));
    if (dlogits) {
      float g = grad_scale * weight * (1.f / (1.f + __expf(-x)) - label);
      dlogits[i] = accumulate ? dlogits[i] + g : g;
    }
  }
  acc = warp_sum(acc);
  if (loss_xs && (threadIdx.x & 31) == 0) xs::add(loss_xs, 1, 0, acc);
}

}  // namespace
}  // namespace mcn

using namespace mcn;

extern "C" int mcn_softmax_xent(const float* logits, const int32_t* labels, long long rows, int C,
                                const float* class_w, float label_smoothing, float focal_gamma,
                                float sigmoid_focal_alpha, int seg_h, int seg_w, float grad_scale,
                                long long* loss_xs, float* dlogits, float* probs, void* stream) {
  MCN_REQUIRE(logits && rows > 0 && C > 0, "softmax_xent: bad argument");
  MCN_REQUIRE(seg_h >= 0 && seg_w >= 0 && (seg_h == 0 || (seg_w > 0 && rows % ((long long)seg_h * seg_w) == 0)),
              "softmax_xent: rows must be whole [seg_h, seg_w] label maps");
  XentOpts o{label_smoothing, focal_gamma, sigmoid_focal_alpha, seg_h, seg_w};
  const char* env_rows = getenv("MCN_XENT_ROWS");      // 0: always the warp-per-row kernel (A/B, tests)
  const bool rows_path = !(env_rows && env_rows[0] == '0');
  if (rows_path && 256 * (C | 1) * 4 <= 48 * 1024 && rows >= 4096 && !(dlogits != nullptr && probs != nullptr)) {
    const long long chunks = (rows + 255) / 256;
    const int grid = (int)std::max<long long>(1, std::min<long long>(chunks, 4LL * num_sms()));
    const size_t smem = static_cast<size_t>(256) * (C | 1) * sizeof(float);
    ::mcn::launch(softmax_xent_rows_kernel, grid, 256, smem, static_cast<cudaStream_t>(stream), logits, labels, rows,
                  C, class_w, o, grad_scale, loss_xs, dlogits != nullptr ? dlogits : probs,
                  dlogits != nullptr ? 1 : 0);
    return after_launch("softmax_xent");
  }
  const int wpb = 8;
  int grid = (int)std::max<long long>(1, std::min<long long>((rows + wpb - 1) / wpb, 8LL * num_sms()));
  ::mcn::launch(softmax_xent_kernel, grid, wpb * 32, 0, static_cast<cudaStream_t>(stream), 
      logits, labels, rows, C, class_w, o, grad_scale, loss_xs, dlogits, probs);
  return after_launch("softmax_xent");
}

extern "C" int mcn_sigmoid_xent(const float* logits, long long n, float label, float weight,
                                float grad_scale, long long* loss_xs, float* dlogits,
                                int accumulate_grad, void* stream) {
  MCN_REQUIRE(logits && n > 0, "sigmoid_xent: bad argument");
  int grid = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 4LL * num_sms()));
  ::mcn::launch(sigmoid_xent_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), 
      logits, n, label, weight, grad_scale, loss_xs, dlogits, accumulate_grad);
  return after_launch("sigmoid_xent");
}
