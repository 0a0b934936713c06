/*
 * libmcn — C ABI of the B200-native layer-op backend.
 *
 * The reference (dooyounggo/MyConvNet) has no native boundary: its layer ops are Python
 * methods of ConvNet (convnet.py:1382-2577) that call TensorFlow library ops, and its gradient
 * averaging is optimizers.py:89-177.  Each entry point below replaces one of those TF call sites;
 * the citation names the reference line whose op it replaces.  The Python host
 * (myconvnet_b200/) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless named host_*
 *   - the caller owns all memory; the library never allocates or frees caller buffers
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*)
 *   - return value: 0 on success, negative mcn_status otherwise; mcn_last_error() gives text
 *   - activations are NHWC, dense conv weights HWIO, exactly as in the reference
 */
#ifndef MCN_H_
#define MCN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum { MCN_OK = 0, MCN_EINVAL = -1, MCN_ECUDA = -2, MCN_EUNSUPPORTED = -3 } mcn_status;
typedef enum { MCN_F32 = 0, MCN_BF16 = 1, MCN_U8 = 2 } mcn_dtype;   /* U8: network input images only */
typedef enum {
  MCN_ACT_NONE = 0,
  MCN_ACT_RELU = 1,
  MCN_ACT_RELU6 = 2,
  MCN_ACT_LRELU = 3,
  MCN_ACT_TANH = 4,
  MCN_ACT_SIGMOID = 5,
  MCN_ACT_SWISH = 6
} mcn_act;

/* Geometry of one convolution.  pad_t/pad_l are the resolved TF SAME/VALID leading pads
 * (extra padding goes bottom/right); Ho/Wo the resulting output size. */
typedef struct {
  int N, H, W, Cin;
  int Cout;
  int kh, kw;
  int sh, sw;
  int dh, dw;
  int pad_t, pad_l;
  int Ho, Wo;
} mcn_conv_desc;

const char* mcn_last_error(void);
int mcn_version(void);
/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
long long mcn_launch_count(void);
/* Debug aid: cycles each warp role of the tensor-core conv kernels spent waiting on its barriers,
 * summed over all CTAs since the last reset.  Only the -DMCN_ROLE_TIMING build (MCN_ROLE_TIMING=1
 * python -m myconvnet_b200.build -> libmcn_timing.so, scripts/role_timing.py) counts; the production
 * library returns zeros. */
int mcn_debug_role_cycles(unsigned long long* out16, int reset);

/* ---- determinism: workspace and exact accumulators --------------------------------------
 * No entry point of this library uses floating-point atomics: every cross-block reduction is
 * either summed in a fixed order (split-K slices of the wgrad kernels) or accumulated in an
 * EXACT fixed-point accumulator ("xsum": three int64 limbs per sum, limb k of element i of an
 * n-element array at limbs[k*n + i]; value = l0*2^-80 + l1*2^-40 + l2), so two runs of the same
 * call give bit-identical results (the reference's TF ops are deterministic per tower).
 * The reductions need scratch memory: ONE device buffer per device, registered here, zero-filled
 * by the caller before the first use (every launch leaves it zeroed again).  Launches that share
 * it must be stream-ordered, as with a cuBLAS handle.  Calls that need it fail with MCN_EINVAL
 * when none is registered.  Size: mcn_workspace_min_bytes() plus, for the tensor-core / direct
 * wgrad kernels, one fp32 dw-sized slice per split (mcn_conv2d_wgrad_workspace_bytes gives the
 * preferred size for one convolution; less means fewer splits, never a different result class). */
int mcn_set_workspace(void* device_ptr, long long bytes);
long long mcn_workspace_min_bytes(void);
long long mcn_conv2d_wgrad_workspace_bytes(const mcn_conv_desc* d, int a_mode, int stem);
/* xsum limbs [3][n] -> out_f32[n] and/or out_f64[n] (either may be NULL); accumulate != 0 adds. */
int mcn_xsum_decode(const long long* limbs, int n, float* out_f32, double* out_f64,
                    int accumulate, void* stream);

/* ---- dense convolution, tensor-core path (tcgen05/TMEM + TMA), bf16 in / fp32 accumulate.
 * Replaces tf.nn.conv2d (convnet.py:1659), its autodiff Conv2DBackpropInput/Filter,
 * tf.nn.conv2d_transpose (convnet.py:2463, = dgrad) and tf.matmul (convnet.py:1743,1755; a
 * 1x1 conv on a [N,1,1,in] tensor).
 *   w_ohwi : bf16 [kh*kw][Cout][Cin]   (fprop operand; produced by mcn_weight_prep)
 *   w_hwio : bf16 [kh*kw][Cin][Cout]   (dgrad operand; the reference's own layout)
 *   dw     : fp32 [kh*kw][Cin][Cout]   (wgrad ACCUMULATES into it: zero it first)
 * a_mode: 0 = tiled TMA boxes (spatial tiles), 1 = im2col TMA (exact 128-pixel tiles).
 * accumulate != 0: the epilogue adds into the existing output (y += conv, dx += dgrad) — used for
 * tensors with several consumers, whose gradient contributions sum (no separate add pass). */
int mcn_conv2d_fprop_tc(const mcn_conv_desc* d, const void* x, const void* w_ohwi,
                        const float* bias, void* y, int y_dtype, int a_mode, int accumulate,
                        void* stream);
/* conv -> batch-norm fusion: same as mcn_conv2d_fprop_tc with a bf16 output, and the epilogue also
 * ACCUMULATES the per-channel statistics of the stored output into bn_sums (fp64 [2*Cout]:
 * sum y | sum y^2; zero first) — the statistics half of tf.nn.fused_batch_norm
 * (convnet.py:1883) without a second pass over y.  Needs Cout % 64 == 0. */
int mcn_conv2d_fprop_tc_stats(const mcn_conv_desc* d, const void* x, const void* w_ohwi,
                              const float* bias, void* y, int a_mode, double* bn_sums,
                              void* stream);
int mcn_conv2d_dgrad_tc(const mcn_conv_desc* d, const void* dy, const void* w_hwio, void* dx,
                        int dx_dtype, int a_mode, int accumulate, void* stream);
int mcn_conv2d_wgrad_tc(const mcn_conv_desc* d, const void* x, const void* dy, float* dw,
                        int a_mode, void* stream);

/* ---- dense / depthwise convolution, CUDA-core direct path (any dtype, any geometry).
 * Same call sites as above plus tf.nn.depthwise_conv2d (convnet.py:1645).  Weights HWIO
 * (depthwise: [kh][kw][C][mult]) in `wdtype`; wgrad ACCUMULATES into fp32 (zero first). */
int mcn_conv2d_fprop_direct(const mcn_conv_desc* d, int dtype, const void* x, int wdtype,
                            const void* w, const float* bias, void* y, void* stream);
int mcn_conv2d_dgrad_direct(const mcn_conv_desc* d, int dtype, const void* dy, int wdtype,
                            const void* w, void* dx, void* stream);
int mcn_conv2d_wgrad_direct(const mcn_conv_desc* d, int dtype, const void* x, const void* dy,
                            float* dw, void* stream);
int mcn_dwconv2d_fwd(const mcn_conv_desc* d, int mult, int dtype, const void* x, int wdtype,
                     const void* w, void* y, void* stream);
int mcn_dwconv2d_bwd_data(const mcn_conv_desc* d, int mult, int dtype, const void* dy, int wdtype,
                          const void* w, void* dx, void* stream);
int mcn_dwconv2d_bwd_filter(const mcn_conv_desc* d, int mult, int dtype, const void* x,
                            const void* dy, float* dw, void* stream);

/* ---- stem convolution: k x k (3 or 7), stride 2 along W, on an RGB image stored with FOUR
 * channels (8-byte pixels, 4th = 0; mcn_pad_rgb4).  Same call site as mcn_conv2d_fprop_tc
 * (tf.nn.conv2d, convnet.py:1659) for the first layer of resnet_v1_5.py:43 / efficientnet.py:62,
 * without an im2col matrix: the GEMM A tile is gathered from the image with 16-byte cp.async
 * copies.  d->Cin must be 4.  The weight is stored as [Kpad][Cout] with row
 * k = (r*(kw+1) + s)*4 + c  (tap s = kw and channel c = 3 are zero), Kpad = mcn_stem_conv_kpad(d):
 *   w_okp : bf16 [Cout][Kpad]   (fprop operand = mcn_weight_prep's transposed copy)
 *   dw    : fp32 [Kpad][Cout]   (wgrad ACCUMULATES; rows of the zero tap are left untouched)
 * bn_sums (may be NULL): fused batch-norm statistics as in mcn_conv2d_fprop_tc_stats. */
int mcn_stem_conv_kpad(const mcn_conv_desc* d);   /* 0 when the geometry is not supported */
int mcn_stem_conv_fprop(const mcn_conv_desc* d, const void* x4, const void* w_okp, const float* bias,
                        void* y, double* bn_sums, void* stream);
int mcn_stem_conv_wgrad(const mcn_conv_desc* d, const void* x4, const void* dy, float* dw,
                        void* stream);
/* bf16 [pixels][3] -> bf16 [pixels][4] (4th channel zero) */
int mcn_pad_rgb4(const void* x_bf16, long long pixels, void* y_bf16, void* stream);

/* fp32 master weight [taps][Cin][Cout] -> bf16 copies in both operand layouts. */
int mcn_weight_prep(const float* w_hwio_f32, int taps, int cin, int cout, void* w_hwio_bf16,
                    void* w_ohwi_bf16, void* stream);
/* Explicit im2col for channel counts the TMA path cannot address (Cin % 8 != 0: RGB stems).
 * col: bf16 [N*Ho*Wo][kpad], kpad >= kh*kw*Cin, zero padded. */
int mcn_im2col(const mcn_conv_desc* d, int dtype, const void* x, void* col, int kpad, void* stream);

/* ---- batch normalisation (tf.nn.fused_batch_norm, convnet.py:1883-1896,1916) ----
 * rows = N*H*W, C channels.
 * stats: sums[0..C) += sum x, sums[C..2C) += sum x^2 in fp64 (ACCUMULATES: zero first; for
 *        synchronised BN all-reduce this 2C vector across ranks before finalize).
 * finalize: mean, biased var -> invstd; also the reference's moving-stat EMA
 *        (convnet.py:1898-1901) with the Bessel-corrected variance; count = global rows.
 *        moving_mean/moving_var may be NULL (update_batch_norm off).
 * apply: y = act(gamma*(x-mean)*invstd + beta [+ residual]) in one pass; gamma/beta may be NULL
 *        (scale=False / shift=False). */
int mcn_bn_stats(int dtype, const void* x, long long rows, int C, double* sums, void* stream);
int mcn_bn_finalize(const double* sums, double count, int C, float eps, float momentum,
                    float* mean, float* invstd, float* moving_mean, float* moving_var,
                    void* stream);
/* frozen BN (update off, convnet.py:1916-1924): mean = moving_mean, invstd = 1/sqrt(moving_var+eps);
 * mcn_bn_apply then normalises with them, the backward pass is mcn_bn_bwd_reduce (dgamma, dbeta)
 * + mcn_bn_bwd_apply with ZERO sums (dx = dz*gamma*invstd: no batch-statistics terms). */
int mcn_bn_frozen_stats(const float* moving_mean, const float* moving_var, int C, float eps,
                        float* mean, float* invstd, void* stream);
int mcn_bn_apply(int dtype, const void* x, long long rows, int C, const float* mean,
                 const float* invstd, const float* gamma, const float* beta, const void* residual,
                 int act, float act_alpha, void* y, void* stream);
/* finalize + apply in one launch: every thread derives mean / invstd of its channels from the
 * (all-reduced) fp64 sums; block 0 also stores them for the backward pass (save_mean,
 * save_invstd) and performs the moving-statistics update (moving_* may be NULL). */
int mcn_bn_apply_stats(int dtype, const void* x, long long rows, int C, const double* sums,
                       double count, float eps, float momentum, const float* gamma,
                       const float* beta, const void* residual, int act, float act_alpha, void* y,
                       float* save_mean, float* save_invstd, float* moving_mean,
                       float* moving_var, void* stream);
/* ReLU bit mask for the layers with a fused residual (bf16, C % 8 == 0, C <= 2048, act = RELU).
 * Their backward passes need the sign of the OUTPUT y (the pre-activation cannot be rebuilt from x alone);
 * mcn_bn_apply_stats_mask is mcn_bn_apply_stats that also writes relu_mask: bit (e & 7) of byte (e >> 3)
 * is set iff the stored y[e] > 0 (rows*C/8 bytes; the buffer must be 4-byte aligned and readable up to
 * the next multiple of 4).  mcn_bn_bwd_reduce_mask / mcn_bn_bwd_apply_mask are mcn_bn_bwd_reduce /
 * mcn_bn_bwd_apply reading that mask instead of y: one bit per element instead of two bytes, twice. */
int mcn_bn_apply_stats_mask(int dtype, const void* x, long long rows, int C, const double* sums,
                            double count, float eps, float momentum, const float* gamma,
                            const float* beta, const void* residual, int act, float act_alpha, void* y,
                            void* relu_mask, float* save_mean, float* save_invstd, float* moving_mean,
                            float* moving_var, void* stream);
int mcn_bn_bwd_reduce_mask(int dtype, const void* dy, const void* x, const void* relu_mask,
                           long long rows, int C, const float* mean, const float* invstd,
                           float* sum_dz, float* sum_dz_xhat, void* stream);
int mcn_bn_bwd_apply_mask(int dtype, const void* dy, const void* x, const void* relu_mask,
                          long long rows, int C, const float* mean, const float* invstd,
                          const float* gamma, const float* sum_dz, const float* sum_dz_xhat,
                          double count, void* dx, void* d_residual, void* stream);
/* Backward reduction fused into the dgrad that PRODUCES the BN output's gradient (the conv consuming
 * the BN+ReLU output): stride-1 bf16 dgrad with dx = d(loss)/d(BN output), plus, from the same
 * epilogue tile, sums[c] += sum dz and sums[C + c] += sum dz*x over all pixels, where x = bn_x is
 * the BN INPUT (same shape as dx), dz = dx * act'(x*gamma*invstd + beta - mean*gamma*invstd), act =
 * MCN_ACT_NONE or MCN_ACT_RELU (the forward pass's own constants: same fmaf, same mask).  sums is an
 * fp64 [2*C] accumulator (zero it first; exact, order-independent like the forward statistics).
 * mcn_bn_bwd_finalize turns it into what mcn_bn_bwd_apply (and dbeta / dgamma) expect:
 * sum_dz += S1, sum_dz_xhat += invstd*(S2 - mean*S1).  Replaces mcn_bn_bwd_reduce for that layer.
 * With sum_dz / sum_dz_xhat given (both or neither), the block that completes a channel group applies
 * that very formula itself and no mcn_bn_bwd_finalize launch is needed (sums then stays untouched but
 * must still be a valid [2*C] buffer). */
int mcn_conv2d_dgrad_bnred_supported(const mcn_conv_desc* d, int a_mode);
int mcn_conv2d_dgrad_tc_bnred(const mcn_conv_desc* d, const void* dy, const void* w_hwio, void* dx,
                              int a_mode, const void* bn_x, const float* mean, const float* invstd,
                              const float* gamma, const float* beta, int act, double* sums,
                              float* sum_dz, float* sum_dz_xhat, void* stream);
int mcn_bn_bwd_finalize(const double* sums, const float* mean, const float* invstd, int C,
                        float* sum_dz, float* sum_dz_xhat, void* stream);
/* inference mode: y = act(gamma*(x-mean)/sqrt(var+eps)+beta [+ residual]) */
int mcn_bn_infer(int dtype, const void* x, long long rows, int C, const float* mean,
                 const float* var, float eps, const float* gamma, const float* beta,
                 const void* residual, int act, float act_alpha, void* y, void* stream);
/* backward, two passes.  dz = dy * act'(.): with `y` (the forward OUTPUT, post-activation and
 * post-residual) the derivative is rebuilt from the output (relu family); with y == NULL the
 * pre-activation is recomputed from x (swish etc.; only valid without a fused residual).
 * reduce: sum_dz[c] += sum dz (= dbeta), sum_dz_xhat[c] += sum dz*xhat (= dgamma): the LOCAL
 *         sums are the parameter gradients; for synchronised BN all-reduce a copy before apply.
 * apply : dx = gamma*invstd*(dz - sum_dz/count - xhat*sum_dz_xhat/count); d_residual = dz
 *         (may be NULL). */
int mcn_bn_bwd_reduce(int dtype, const void* dy, const void* x, const void* y, long long rows,
                      int C, const float* mean, const float* invstd, const float* gamma,
                      const float* beta, int act, float act_alpha, float* sum_dz,
                      float* sum_dz_xhat, void* stream);
int mcn_bn_bwd_apply(int dtype, const void* dy, const void* x, const void* y, long long rows,
                     int C, const float* mean, const float* invstd, const float* gamma,
                     const float* beta, int act, float act_alpha, const float* sum_dz,
                     const float* sum_dz_xhat, double count, void* dx, void* d_residual,
                     void* stream);

/* ---- pooling (tf.nn.max_pool convnet.py:1509, tf.nn.avg_pool :1548, tf.reduce_mean over H,W
 * resnet_v1_5.py:73).  argmax is the flattened input offset (h*W + w)*C + c within the image,
 * first maximum in row-major window order (TF CPU tie rule). */
int mcn_maxpool_fwd(int dtype, const void* x, int N, int H, int W, int C, int kh, int kw, int sh,
                    int sw, int pad_t, int pad_l, int Ho, int Wo, void* y, int32_t* argmax,
                    void* stream);
int mcn_maxpool_bwd(int dtype, const void* dy, const int32_t* argmax, int N, int H, int W, int C,
                    int kh, int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo, void* dx,
                    void* stream);
/* compact form used by the training step: the winner is stored as its TAP index a*kw + b inside the
 * window (one byte instead of four: the int32 argmax is twice the size of a bf16 tensor); same
 * tie rule.  mcn_maxpool_tap_to_argmax expands it to the TF index above, bit for bit.
 * Needs kh*kw <= 255 and C a multiple of the 16-byte vector (8 bf16 / 4 fp32). */
int mcn_maxpool_fwd_tap(int dtype, const void* x, int N, int H, int W, int C, int kh, int kw, int sh,
                        int sw, int pad_t, int pad_l, int Ho, int Wo, void* y, uint8_t* tap, void* stream);
int mcn_maxpool_bwd_tap(int dtype, const void* dy, const uint8_t* tap, int N, int H, int W, int C, int kh,
                        int kw, int sh, int sw, int pad_t, int pad_l, int Ho, int Wo, void* dx,
                        void* stream);
int mcn_maxpool_tap_to_argmax(const uint8_t* tap, int N, int H, int W, int C, int kh, int kw, int sh,
                              int sw, int pad_t, int pad_l, int Ho, int Wo, int32_t* argmax, void* stream);
int mcn_avgpool_fwd(int dtype, const void* x, int N, int H, int W, int C, int kh, int kw, int sh,
                    int sw, int pad_t, int pad_l, int Ho, int Wo, void* y, void* stream);
int mcn_avgpool_bwd(int dtype, const void* dy, int N, int H, int W, int C, int kh, int kw, int sh,
                    int sw, int pad_t, int pad_l, int Ho, int Wo, void* dx, void* stream);
int mcn_gap_fwd(int dtype, const void* x, int N, int HW, int C, void* y, int y_dtype, void* stream);
int mcn_gap_bwd(int dtype, const void* dy, int dy_dtype, int N, int HW, int C, void* dx,
                void* stream);

/* ---- element-wise (convnet.py:2500-2556, efficientnet.py:163) ---- */
int mcn_act_fwd(int dtype, const void* x, long long n, int act, float alpha, void* y, void* stream);
int mcn_act_bwd(int dtype, const void* dy, const void* x, long long n, int act, float alpha,
                void* dx, void* stream);
/* y = act(a + b)  (stochastic_depth with drop_rate 0 followed by relu) */
int mcn_add_act_fwd(int dtype, const void* a, const void* b, long long n, int act, float alpha,
                    void* y, void* stream);
/* dz = dy*act'(y)  (relu family, from the output) */
int mcn_add_act_bwd(int dtype, const void* dy, const void* y, long long n, int act, float alpha,
                    void* dz, void* stream);
/* a += b */
int mcn_accumulate(int dtype, void* a, const void* b, long long n, void* stream);
/* y[n,hw,c] = x[n,hw,c]*m[n,c]; backward gives dx and dm (SE excite) */
int mcn_scale_bcast_fwd(int dtype, const void* x, const void* m, int N, int HW, int C, void* y,
                        void* stream);
int mcn_scale_bcast_bwd(int dtype, const void* dy, const void* x, const void* m, int N, int HW,
                        int C, void* dx, float* dm, void* stream);
/* per-channel bias add / bias gradient (tf.nn.bias_add convnet.py:1694) */
int mcn_bias_add(int dtype, void* y, long long rows, int C, const float* bias, void* stream);
int mcn_bias_grad(int dtype, const void* dy, long long rows, int C, float* db, void* stream);
int mcn_cast(int src_dtype, const void* src, int dst_dtype, void* dst, long long n, void* stream);
/* network input prologue (convnet.py:449-471): x is [N,Hi,Wi,C] fp32 in [0,1] (MCN_F32) or raw uint8
 * (MCN_U8: divided by 255 first); centre crop to [N,H,W,C] with offsets (Hi-H)//2, (Wi-W)//2
 * (center_crop, convnet.py:1137-1149), y = (x - mean)*scale, cast to the compute dtype. */
int mcn_input_prep(const void* x, int src_dtype, int N, int Hi, int Wi, int H, int W, int C,
                   float mean, float scale, int dst_dtype, void* y, void* stream);
/* channel concat / split of NHWC tensors (tf.concat axis=-1, deeplabv3plus.py:100,110) */
int mcn_copy_channels(int dtype, const void* src, long long rows, int Csrc, int src_off, void* dst,
                      int Cdst, int dst_off, int Ccopy, int accumulate, void* stream);
/* bilinear resize (tf.image.resize_bilinear, convnet.py:2397). mode: 0 legacy, 1 align_corners,
 * 2 half_pixel_centers. */
int mcn_resize_bilinear_fwd(int dtype, const void* x, int N, int H, int W, int C, int Ho, int Wo,
                            int mode, void* y, void* stream);
int mcn_resize_bilinear_bwd(int dtype, const void* dy, int N, int H, int W, int C, int Ho, int Wo,
                            int mode, void* dx, void* stream);

/* nearest-neighbour resize (tf.image.resize_nearest_neighbor, convnet.py:2393-2395); modes as above:
 * 0 src = floor(dst*in/out), 1 src = round(dst*(in-1)/(out-1)), 2 src = floor((dst+0.5)*in/out). */
int mcn_resize_nearest_fwd(int dtype, const void* x, int N, int H, int W, int C, int Ho, int Wo,
                           int mode, void* y, void* stream);
int mcn_resize_nearest_bwd(int dtype, const void* dy, int N, int H, int W, int C, int Ho, int Wo,
                           int mode, void* dx, void* stream);

/* ---- random train-time ops.  Keep decisions are a pure function of (seed, step, layer, element):
 * Philox4x32-10 with counter (idx_lo, idx_hi, step, 0) and key (seed, layer); u = (bits>>8)*2^-24;
 * keep <=> u >= rate; kept values are scaled by 1/(1-rate).  seed and step are read on the device
 * from hp (the optimiser's hyper-parameter vector: uint32 at float slots 12 and 13), so a replayed
 * CUDA graph draws new masks every step and the backward pass recomputes them.
 * dropout: tf.nn.dropout(x, rate) (resnet_v1_5.py:75, efficientnet.py:117); element i uses output
 *          i%4 of counter i/4; the same call with dy in place of x is the backward pass.
 * sd_add : ConvNet.stochastic_depth (convnet.py:2500-2512): y = act(a*survived[n] + b) with
 *          survived[n] = (u_n >= rate)/(1-rate) per sample (counter = n, output 0);
 *          backward: dz = dy*act'(y), da = dz*survived[n], db = dz (either may be NULL). */
int mcn_dropout(int dtype, const void* x, long long n, float rate, const float* hp, int layer, void* y,
                void* stream);
int mcn_sd_add_fwd(int dtype, const void* a, const void* b, int N, long long per_sample, float rate,
                   const float* hp, int layer, int act, float alpha, void* y, void* stream);
int mcn_sd_add_bwd(int dtype, const void* dy, const void* y, int N, long long per_sample, float rate,
                   const float* hp, int layer, int act, float alpha, void* da, void* db, void* stream);

/* ---- losses (convnet.py:594-600, gan.py:134-138) ----
 * softmax cross-entropy on fp32 logits [rows][C] with int32 labels (-1 = all-zero one-hot row,
 * convnet.py:448-449).  loss_sum accumulates sum_rows w*valid*CE; dlogits = grad_scale *
 * w*valid*(softmax - smoothed_onehot).  probs may be NULL.  loss_xs: one xsum accumulator
 * (3 int64, zero first; decode on the host or with mcn_xsum_decode).
 * label_smoothing: targets onehot*(1-ls) + ls/C (convnet.py:603-607) or, with seg_h/seg_w > 0 (rows
 * are the pixels of [N,seg_h,seg_w] label maps), onehot*(1-ls) + ls*avg5x5(onehot)
 * (segmentation/segnet.py:116-121).  focal_gamma > 0: CE *= (1-p_true)^gamma; sigmoid_focal_alpha
 * > 0: CE *= stop_gradient(1-sigmoid(alpha*(p_true-0.5)))/(1-sigmoid(-alpha/2)) (convnet.py:580-592). */
int mcn_softmax_xent(const float* logits, const int32_t* labels, long long rows, int C,
                     const float* class_w, float label_smoothing, float focal_gamma,
                     float sigmoid_focal_alpha, int seg_h, int seg_w, float grad_scale,
                     long long* loss_xs, float* dlogits, float* probs, void* stream);
int mcn_sigmoid_xent(const float* logits, long long n, float label, float weight, float grad_scale,
                     long long* loss_xs, float* dlogits, int accumulate_grad, void* stream);

/* ---- optimiser: fused multi-tensor update (optimizers.py:149-176, 668-705; EMA
 * convnet.py:183-184,1401; L2 term convnet.py:563).  One launch updates every tensor in the
 * table.  Order per element: shadow EMA on the PRE-step value, g += l2*w,
 * optimiser rule, decoupled weight decay on the post-step value; then bf16 copies refreshed. */
typedef struct {
  float* w;            /* fp32 master */
  const float* g;      /* fp32 gradient (already averaged over ranks); NULL = not trainable:
                          only the EMA shadow and the bf16 copies are refreshed */
  float* m;            /* momentum / Adam m / RMSProp mom */
  float* v;            /* Adam v / RMSProp ms (NULL for SGD) */
  float* ema;          /* EMA shadow (NULL = none) */
  void* w_bf16;        /* bf16 copy, same layout (NULL = none) */
  void* w_bf16_t;      /* bf16 [taps][cout][cin] copy (NULL = none) */
  long long n;         /* elements */
  int taps, cin, cout; /* for the transposed copy */
  float l2;            /* L2 factor for this tensor (0 for biases/norm unless bias_norm_decay) */
  float wd;            /* decoupled weight decay factor (already scaled) */
  float l1;            /* L1 factor (convnet.py:555-558): g += l1*sign(w), loss += l1*sum|w| */
} mcn_opt_tensor;
typedef enum { MCN_OPT_NESTEROV = 0, MCN_OPT_RMSPROP = 1, MCN_OPT_ADAM = 2 } mcn_opt_kind;
/* table: device array of mcn_opt_tensor; hp (device, 11 floats): lr, momentum(beta1),
 * beta2/decay, eps, ema_decay_t, adam_lr_t, grad_scale, weight-decay multiplier, clip threshold,
 * weight-decay form (0 = wd*w, 1 = wd*sign(w), 2 = pseudo-Huber wd*w/sqrt(1+(w/delta)^2);
 * optimizers.py:163-172), Huber delta — read on device so a captured CUDA graph can be replayed
 * with new hyper-parameters. */
int mcn_opt_step(int kind, const mcn_opt_tensor* table, int ntensors, long long max_n,
                 const float* hp, long long* l2_loss_xs, const long long* grad_sqnorm_xs,
                 void* stream);
/* Gradient clipping (tf.clip_by_global_norm, optimizers.py:112-113): out_xs (one xsum accumulator,
 * zero first) += sum over all trainable tensors of (g*hp[6] + l2*w)^2.  Passing it to mcn_opt_step
 * as grad_sqnorm_xs scales every gradient by hp[8] / max(sqrt(sum), hp[8]); NULL = no clipping.
 * l2_loss_xs (may be NULL): xsum accumulator receiving sum_t l2_t * |w_t|^2 / 2 (pre-step). */
int mcn_grad_sqnorm(const mcn_opt_tensor* table, int ntensors, long long max_n, const float* hp,
                    long long* out_xs, void* stream);
/* out[t][c][r] += in[t][r][c]: gradient of a transposed-conv weight (stored [kh,kw,Cin,Cout],
 * reference convnet.py:2460-2462) from the wgrad of the underlying conv. */
int mcn_transpose_add_f32(const float* in, int taps, int rows, int cols, float* out, void* stream);

/* ---- multi-GPU: one-shot all-reduce (sum) of a small vector over peer-mapped memory ----
 * Replaces, for synchronised batch-norm, the reference's tower-after-tower statistics chain
 * (convnet.py:1898-1914) and the per-layer ncclAllReduce: every rank writes its vector into a
 * mailbox slot in every peer's memory over NVLink, raises a flag, waits for all flags and sums the
 * slots in rank order (bit-identical on all ranks).
 *   peers    : DEVICE array [world] of the peer-mapped base address of every rank's symmetric region
 *   mail_off : byte offset of this collective point's two mailboxes in a region: mailbox
 *              (sequence & 1) starts at mail_off + (sequence & 1)*parity_stride and holds
 *              [world][n0+n1] elements (double-buffered so that reuse can never overtake a slower
 *              peer's read, whatever the number of collective points per step)
 *              Default protocol ("LL"): every 8-byte word on the wire is {32 data bits | 32-bit sequence
 *              tag} (an fp64 value = two words), the receiver polls the words themselves: no system
 *              fence, no flag.  A mailbox then holds [world][n0+n1] values of 2*sizeof(T) bytes, i.e.
 *              parity_stride >= world*(n0+n1)*2*sizeof(T).  MCN_PEER_LL=0 selects the fence + flag protocol.
 *   flag_off : byte offset of its flag row ([world] uint64, zero-initialised; fence + flag protocol only)
 *   The wait for the peers is bounded by MCN_PEER_TIMEOUT_S seconds (default 1800).
 *   counter  : local device uint64[2] of this collective point, zero-initialised: the sequence number
 *              and the finishing ticket of the kernel's blocks (large vectors are shared by up to 8 blocks)
 *   src0/src1: local source segments (n0, n1 elements; src1 may be NULL when n1 == 0)
 *   dst      : local destination, n0+n1 elements (may alias src0) */
int mcn_peer_allreduce(const unsigned long long* peers, long long mail_off, long long parity_stride,
                       long long flag_off, unsigned long long* counter, int is_f64, const void* src0,
                       int n0, const void* src1, int n1, void* dst, int rank, int world, void* stream);

/* ---- group normalisation (convnet.py:1928-2013) and weight standardisation (convnet.py:1410-1419) ----
 * x, y, dy, dx: [N, HW, C] (NHWC with HW = H*W; HW = 1 for [N, C] tensors); G groups of C/G channels.
 * gn_fwd: save[(n*G+g)*2] = {mean, 1/sqrt(var+eps)} of group g of sample n (population variance),
 *         y = (x - mean) * invstd * gamma[c] + beta[c]   (gamma / beta may be NULL = 1 / 0).
 * gn_bwd: dx (may be NULL) = invstd*(dy*gamma - mean_g(dy*gamma) - xhat*mean_g(dy*gamma*xhat));
 *         dgamma[c] += sum dy*xhat, dbeta[c] += sum dy (either may be NULL).  scratch: fp32
 *         [2*N*G + 2*N*C].  Every sum has a fixed order: results are bit-reproducible.
 * ws_fwd: w [rows][cols] fp32 (cols = output channels, the LAST axis of the reference's weight):
 *         w_std = (w - mean_col) / (std_col + eps), stats = [mean | std] per column (population std).
 * ws_bwd: grad += d(loss)/dw given g_std = d(loss)/d(w_std) (the autodiff of the three lines above). */
int mcn_gn_fwd(int dtype, const void* x, int N, long long HW, int C, int G, float eps,
               const float* gamma, const float* beta, void* y, float* save, void* stream);
int mcn_gn_bwd(int dtype, const void* dy, const void* x, int N, long long HW, int C, int G,
               const float* gamma, const float* save, float* scratch, void* dx, float* dgamma,
               float* dbeta, void* stream);
int mcn_ws_fwd(const float* w, int rows, int cols, float eps, float* w_std, float* stats, void* stream);
int mcn_ws_bwd(const float* g_std, const float* w, const float* stats, int rows, int cols, float eps,
               float* grad, void* stream);

/* ---- generic helpers ---- */
int mcn_fill_f32(float* p, long long n, float v, void* stream);
int mcn_scale_f32(float* p, long long n, float s, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MCN_H_ */
