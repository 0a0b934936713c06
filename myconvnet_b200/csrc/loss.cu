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
    if (valid) {
      // sum_c t_c and sum_c t_c z_c
      float a = 0.f, b = 0.f;
      if (o.ls > 0.f) {
        for (int c = lane; c < C; c += 32) {
          const float t = xent_target(o, labels, r, y, c, C);
          a += t;
          b = fmaf(t, z[c], b);
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
    for (int c = lane; c < C; c += 32) {
      const float p = __expf(z[c] - mx) * inv_se;
      if (probs) probs[r * C + c] = p;
      if (dlogits) {
        float g = 0.f;
        if (valid) {
          const float t = xent_target(o, labels, r, y, c, C);
          g = f_mul * (p * st - t) + ce * df * ((c == y ? 1.f : 0.f) - p);
        }
        dlogits[r * C + c] = grad_scale * w * s_mul * g;
      }
    }
  }
  if (loss_xs && lane == 0) xs::add(loss_xs, 1, 0, block_loss);
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
