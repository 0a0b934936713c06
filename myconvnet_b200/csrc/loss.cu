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
__global__ void softmax_xent_kernel(const float* __restrict__ logits,
                                    const int32_t* __restrict__ labels, long long rows, int C,
                                    const float* __restrict__ class_w, float ls, float grad_scale,
                                    long long* __restrict__ loss_xs, float* __restrict__ dlogits,
                                    float* __restrict__ probs) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  float block_loss = 0.f;
  for (long long r = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); r < rows;
       r += (long long)gridDim.x * wpb) {
    const float* z = logits + r * C;
    float mx = -FLT_MAX;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, z[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += __expf(z[c] - mx);
    se = warp_sum(se);
    const float lse = mx + __logf(se);
    const int y = labels ? labels[r] : -1;
    const bool valid = (y >= 0 && y < C);
    const float w = valid ? (class_w ? class_w[y] : 1.f) : 0.f;
    // smoothed target t_c = onehot*(1-ls) + ls/C  (convnet.py:606); sum_c t_c = 1
    // CE = -sum_c t_c (z_c - lse) = lse - (1-ls) z_y - (ls/C) sum_c z_c
    float sz = 0.f;
    if (ls > 0.f) {
      for (int c = lane; c < C; c += 32) sz += z[c];
      sz = warp_sum(sz);
    }
    if (valid && lane == 0) {
      float ce = lse - (1.f - ls) * z[y] - (ls > 0.f ? ls / (float)C * sz : 0.f);
      block_loss += w * ce;
    }
    const float inv_se = 1.f / se;
    for (int c = lane; c < C; c += 32) {
      float p = __expf(z[c] - mx) * inv_se;
      if (probs) probs[r * C + c] = p;
      if (dlogits) {
        float t = (c == y ? 1.f - ls : 0.f) + ls / (float)C;
        dlogits[r * C + c] = grad_scale * w * (p - t);
      }
    }
  }
  if (loss_xs && lane == 0) xs::add(loss_xs, 1, 0, block_loss);
}

// loss = max(x,0) - x*z + log1p(exp(-|x|));  d/dx = sigmoid(x) - z
__global__ void sigmoid_xent_kernel(const float* __restrict__ logits, long long n, float label,
                                    float weight, float grad_scale, long long* __restrict__ loss_xs,
                                    float* __restrict__ dlogits, int accumulate) {
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
                                const float* class_w, float label_smoothing, float grad_scale,
                                long long* loss_xs, float* dlogits, float* probs, void* stream) {
  MCN_REQUIRE(logits && rows > 0 && C > 0, "softmax_xent: bad argument");
  const int wpb = 8;
  int grid = (int)std::max<long long>(1, std::min<long long>((rows + wpb - 1) / wpb, 8LL * num_sms()));
  softmax_xent_kernel<<<grid, wpb * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, labels, rows, C, class_w, label_smoothing, grad_scale, loss_xs, dlogits, probs);
  return after_launch("softmax_xent");
}

extern "C" int mcn_sigmoid_xent(const float* logits, long long n, float label, float weight,
                                float grad_scale, long long* loss_xs, float* dlogits,
                                int accumulate_grad, void* stream) {
  MCN_REQUIRE(logits && n > 0, "sigmoid_xent: bad argument");
  int grid = (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 4LL * num_sms()));
  sigmoid_xent_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, n, label, weight, grad_scale, loss_xs, dlogits, accumulate_grad);
  return after_launch("sigmoid_xent");
}
