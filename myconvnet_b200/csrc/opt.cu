// Fused multi-tensor optimiser step.
// Replaces, in ONE launch over every variable: the EMA shadow update (reference
// convnet.py:183-184,1401 — applied to the PRE-step value, optimizers.py:159,175), the L2 term of
// the loss gradient (convnet.py:563), apply_gradients for Nesterov momentum / RMSProp / Adam
// (optimizers.py:668-705), the decoupled weight decay on the post-step value
// (optimizers.py:163-172) and the per-use fp32->half cast of the weights (convnet.py:1421), here
// a bf16 copy in both tensor-core operand layouts.
#include "mcn_common.cuh"
#include "xsum.cuh"

namespace mcn {
namespace {

constexpr int kItems = 8;
constexpr int kBlock = 256;

// Squared global norm of the full gradient (data term averaged over ranks + L2 term), the
// quantity tf.clip_by_global_norm needs (reference optimizers.py:112-113).  Same grid as the step.
__global__ void __launch_bounds__(kBlock)
grad_sqnorm_kernel(const mcn_opt_tensor* __restrict__ table, const float* __restrict__ hp,
                   long long* __restrict__ out_xs) {
  MCN_PDL_PROLOGUE();
  const mcn_opt_tensor t = table[blockIdx.y];
  const long long base = (long long)blockIdx.x * (kBlock * kItems);
  if (base >= t.n || t.g == nullptr) return;
  const float gscale = hp[6];
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const long long i = base + k * kBlock + threadIdx.x;
    if (i < t.n) {
      const float wv = t.w[i];
      const float g = t.g[i] * gscale + t.l2 * wv + t.l1 * (wv > 0.f ? 1.f : (wv < 0.f ? -1.f : 0.f));
      acc = fmaf(g, g, acc);
    }
  }
  __shared__ float part[kBlock / 32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kBlock / 32; ++w) s += part[w];
    xs::add(out_xs, 1, 0, s);
  }
}

// hp: [0] lr  [1] momentum|beta1  [2] decay|beta2  [3] eps  [4] ema decay d_t
//     [5] adam lr_t  [6] gradient scale  [7] weight-decay multiplier  [8] clip threshold
//     [9] decoupled weight-decay form: 0 = w -= wd*w, 1 = w -= wd*sign(w) (l1_weight_decay),
//         2 = w -= wd*w/sqrt(1+(w/delta)^2) (pseudo-Huber, optimizers.py:165-171)  [10] Huber delta
__global__ void __launch_bounds__(kBlock)
opt_step_kernel(int kind, const mcn_opt_tensor* __restrict__ table, const float* __restrict__ hp,
                long long* __restrict__ l2_xs, const long long* __restrict__ grad_sqnorm_xs) {
  MCN_PDL_PROLOGUE();
  const mcn_opt_tensor t = table[blockIdx.y];
  const long long base = (long long)blockIdx.x * (kBlock * kItems);
  if (base >= t.n) return;
  const float lr = hp[0], mom = hp[1], b2 = hp[2], eps = hp[3], ema_d = hp[4], adam_lr = hp[5],
              gscale = hp[6], wd = t.wd * hp[7];
  const int wd_form = static_cast<int>(hp[9]);
  const float huber_delta = hp[10];
  // tf.clip_by_global_norm: g * clip / max(global_norm, clip)
  float clip = 1.f;
  if (grad_sqnorm_xs != nullptr) {
    const float thr = hp[8];
    clip = thr / fmaxf(static_cast<float>(sqrt(xs::read(grad_sqnorm_xs, 1, 0))), thr);
  }
  // regularisation loss over the PRE-step weights: l2 * sum(w^2)/2 (tf.nn.l2_loss, convnet.py:563)
  // + l1 * sum|w| (convnet.py:557)
  float l2_acc = 0.f;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    const long long i = base + k * kBlock + threadIdx.x;
    if (i >= t.n) continue;
    float w = t.w[i];
    l2_acc = fmaf(0.5f * t.l2 * w, w, l2_acc) + t.l1 * fabsf(w);
    if (t.ema) {
      float s = t.ema[i];
      t.ema[i] = s - (1.f - ema_d) * (s - w);
    }
    if (t.g != nullptr) {
      float g = (t.g[i] * gscale + t.l2 * w + t.l1 * (w > 0.f ? 1.f : (w < 0.f ? -1.f : 0.f))) * clip;
      if (kind == MCN_OPT_NESTEROV) {
        float a = mom * t.m[i] + g;
        t.m[i] = a;
        w -= lr * (g + mom * a);
      } else if (kind == MCN_OPT_RMSPROP) {
        float ms = b2 * t.v[i] + (1.f - b2) * g * g;
        t.v[i] = ms;
        float m = mom * t.m[i] + lr * g * rsqrtf(ms + eps);
        t.m[i] = m;
        w -= m;
      } else {
        float m = mom * t.m[i] + (1.f - mom) * g;
        float v = b2 * t.v[i] + (1.f - b2) * g * g;
        t.m[i] = m;
        t.v[i] = v;
        w -= adam_lr * m / (sqrtf(v) + eps);
      }
      if (wd != 0.f) {
        if (wd_form == 1) w -= wd * (w > 0.f ? 1.f : (w < 0.f ? -1.f : 0.f));
        else if (wd_form == 2) w -= wd * w / sqrtf(1.f + (w / huber_delta) * (w / huber_delta));
        else w -= wd * w;
      }
      t.w[i] = w;
    }
    if (t.w_bf16) reinterpret_cast<__nv_bfloat16*>(t.w_bf16)[i] = __float2bfloat16_rn(w);
    if (t.w_bf16_t) {
      // i = (tap*cin + ci)*cout + co  ->  (tap*cout + co)*cin + ci
      int co = (int)(i % t.cout);
      long long r = i / t.cout;
      int ci = (int)(r % t.cin);
      long long tap = r / t.cin;
      reinterpret_cast<__nv_bfloat16*>(t.w_bf16_t)[(tap * t.cout + co) * t.cin + ci] =
          __float2bfloat16_rn(w);
    }
  }
  if (l2_xs != nullptr && (t.l2 != 0.f || t.l1 != 0.f)) {
    l2_acc = warp_sum(l2_acc);
    if ((threadIdx.x & 31) == 0) xs::add(l2_xs, 1, 0, l2_acc);
  }
}

// out[t][c][r] += in[t][r][c]   (in: [taps][rows][cols])
__global__ void transpose_add_kernel(const float* __restrict__ in, int rows, int cols,
                                     float* __restrict__ out) {
  MCN_PDL_PROLOGUE();
  __shared__ float tile[32][33];
  const int t = blockIdx.z;
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  const float* src = in + (long long)t * rows * cols;
  float* dst = out + (long long)t * rows * cols;
  for (int r = threadIdx.y; r < 32; r += 8) {
    int rr = r0 + r, cc = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (rr < rows && cc < cols) ? src[(long long)rr * cols + cc] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    int cc = c0 + r, rr = r0 + threadIdx.x;
    if (rr < rows && cc < cols) dst[(long long)cc * rows + rr] += tile[threadIdx.x][r];
  }
}

// Tiled transpose of one weight tensor: fp32 [taps][cin][cout] -> bf16 same + bf16 [taps][cout][cin]
__global__ void weight_prep_kernel(const float* __restrict__ w, int cin, int cout,
                                   __nv_bfloat16* __restrict__ o_same,
                                   __nv_bfloat16* __restrict__ o_t) {
  MCN_PDL_PROLOGUE();
  __shared__ float tile[32][33];
  const int tap = blockIdx.z;
  const int ci0 = blockIdx.y * 32, co0 = blockIdx.x * 32;
  const float* src = w + (long long)tap * cin * cout;
  for (int r = threadIdx.y; r < 32; r += 8) {
    int ci = ci0 + r, co = co0 + threadIdx.x;
    float v = (ci < cin && co < cout) ? src[(long long)ci * cout + co] : 0.f;
    tile[r][threadIdx.x] = v;
    if (o_same && ci < cin && co < cout)
      o_same[(long long)tap * cin * cout + (long long)ci * cout + co] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  if (o_t)
    for (int r = threadIdx.y; r < 32; r += 8) {
      int co = co0 + r, ci = ci0 + threadIdx.x;
      if (ci < cin && co < cout)
        o_t[(long long)tap * cin * cout + (long long)co * cin + ci] =
            __float2bfloat16_rn(tile[threadIdx.x][r]);
    }
}

}  // namespace
}  // namespace mcn

using namespace mcn;

extern "C" int mcn_opt_step(int kind, const mcn_opt_tensor* table, int ntensors, long long max_n,
                            const float* hp, long long* l2_xs, const long long* grad_sqnorm_xs,
                            void* stream) {
  MCN_REQUIRE(table && hp && ntensors > 0 && max_n > 0, "opt_step: bad argument");
  MCN_REQUIRE(kind >= MCN_OPT_NESTEROV && kind <= MCN_OPT_ADAM, "opt_step: unknown optimiser %d", kind);
  MCN_REQUIRE(ntensors <= 65535, "opt_step: too many tensors");
  dim3 grid((unsigned)((max_n + kBlock * kItems - 1) / (kBlock * kItems)), (unsigned)ntensors);
  ::mcn::launch(opt_step_kernel, grid, kBlock, 0, static_cast<cudaStream_t>(stream), kind, table, hp, l2_xs,
                                                                         grad_sqnorm_xs);
  return after_launch("opt_step");
}

extern "C" int mcn_grad_sqnorm(const mcn_opt_tensor* table, int ntensors, long long max_n,
                               const float* hp, long long* out_xs, void* stream) {
  MCN_REQUIRE(table && hp && out_xs && ntensors > 0 && max_n > 0, "grad_sqnorm: bad argument");
  MCN_REQUIRE(ntensors <= 65535, "grad_sqnorm: too many tensors");
  dim3 grid((unsigned)((max_n + kBlock * kItems - 1) / (kBlock * kItems)), (unsigned)ntensors);
  ::mcn::launch(grad_sqnorm_kernel, grid, kBlock, 0, static_cast<cudaStream_t>(stream), table, hp, out_xs);
  return after_launch("grad_sqnorm");
}

extern "C" int mcn_weight_prep(const float* w_hwio_f32, int taps, int cin, int cout,
                               void* w_hwio_bf16, void* w_ohwi_bf16, void* stream) {
  MCN_REQUIRE(w_hwio_f32 && taps > 0 && cin > 0 && cout > 0, "weight_prep: bad argument");
  dim3 grid((cout + 31) / 32, (cin + 31) / 32, taps), block(32, 8);
  ::mcn::launch(weight_prep_kernel, grid, block, 0, static_cast<cudaStream_t>(stream), 
      w_hwio_f32, cin, cout, static_cast<__nv_bfloat16*>(w_hwio_bf16),
      static_cast<__nv_bfloat16*>(w_ohwi_bf16));
  return after_launch("weight_prep");
}

extern "C" int mcn_transpose_add_f32(const float* in, int taps, int rows, int cols, float* out,
                                     void* stream) {
  MCN_REQUIRE(in && out && taps > 0 && rows > 0 && cols > 0, "transpose_add: bad argument");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32, taps), block(32, 8);
  ::mcn::launch(transpose_add_kernel, grid, block, 0, static_cast<cudaStream_t>(stream), in, rows, cols, out);
  return after_launch("transpose_add");
}
