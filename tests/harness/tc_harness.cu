// Standalone parity harness for the tensor-core convolution entry points of libmcn.
// Each case runs in its own process (a device trap kills the context):
//   tc_harness list          -> number of cases
//   tc_harness <case-index>  -> runs one case, prints PASS/FAIL with max error
// The CPU reference here is a plain loop nest over bf16-rounded inputs (fp64 accumulate).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mcn.h"

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

static float bf16r(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }
static uint32_t rng_state = 12345;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

struct Case {
  const char* name;
  int op;  // 0 fprop, 1 dgrad, 2 wgrad
  int N, H, W, Cin, Cout, k, s, d;
  int a_mode;
  int same;  // 1 SAME, 0 VALID
  int bias;
  int out_f32;
  int timing;  // >0: also time it (iterations)
};

static void same_pad(int in, int k, int s, int d, int same, int* out, int* pad_lo) {
  if (same) {
    *out = (in + s - 1) / s;
    int tot = (*out - 1) * s + (k - 1) * d + 1 - in;
    if (tot < 0) tot = 0;
    *pad_lo = tot / 2;
  } else {
    *out = (in - (k - 1) * d + s - 1) / s;
    *pad_lo = 0;
  }
}

static std::vector<Case> cases() {
  std::vector<Case> c;
  // name, op, N,H,W,Cin,Cout,k,s,d, mode, same, bias, f32, timing
  c.push_back({"fprop 1x1 gemm 512x128x128", 0, 2, 16, 16, 128, 128, 1, 1, 1, 0, 1, 0, 0, 0});
  c.push_back({"fprop 1x1 ragged M200 K72 N40 bias f32", 0, 2, 10, 10, 72, 40, 1, 1, 1, 0, 1, 1, 1, 0});
  c.push_back({"fprop 1x1 N256 (block_n 256)", 0, 2, 16, 16, 64, 256, 1, 1, 1, 0, 1, 0, 0, 0});
  c.push_back({"fprop 3x3 s1 box", 0, 2, 14, 14, 64, 64, 3, 1, 1, 0, 1, 0, 0, 0});
  c.push_back({"fprop 3x3 s1 im2col", 0, 2, 14, 14, 64, 64, 3, 1, 1, 1, 1, 0, 0, 0});
  c.push_back({"fprop 3x3 s1 box 56x56 C128", 0, 1, 56, 56, 128, 128, 3, 1, 1, 0, 1, 0, 0, 0});
  c.push_back({"fprop 3x3 s1 im2col 56x56 C128", 0, 1, 56, 56, 128, 128, 3, 1, 1, 1, 1, 0, 0, 0});
  c.push_back({"fprop 3x3 s2 im2col", 0, 2, 28, 28, 64, 128, 3, 2, 1, 1, 1, 0, 0, 0});
  c.push_back({"fprop 3x3 d2 box", 0, 2, 16, 16, 64, 64, 3, 1, 2, 0, 1, 0, 0, 0});
  c.push_back({"fprop 3x3 d2 im2col", 0, 2, 16, 16, 64, 64, 3, 1, 2, 1, 1, 0, 0, 0});
  c.push_back({"fprop 1x1 s2 box(strided view)", 0, 2, 28, 28, 64, 128, 1, 2, 1, 0, 1, 0, 0, 0});
  c.push_back({"fprop 1x1 s2 im2col", 0, 2, 28, 28, 64, 128, 1, 2, 1, 1, 1, 0, 0, 0});
  c.push_back({"fprop 5x5 s2 im2col", 0, 2, 16, 16, 64, 64, 5, 2, 1, 1, 1, 1, 0, 0});
  c.push_back({"fprop 3x3 VALID im2col", 0, 2, 15, 15, 64, 64, 3, 1, 1, 1, 0, 0, 0, 0});
  c.push_back({"fprop 7x7 box batch-merged", 0, 5, 7, 7, 128, 64, 3, 1, 1, 0, 1, 0, 0, 0});
  c.push_back({"dgrad 1x1", 1, 2, 16, 16, 128, 64, 1, 1, 1, 0, 1, 0, 0, 0});
  c.push_back({"dgrad 3x3 s1 box", 1, 2, 14, 14, 64, 128, 3, 1, 1, 0, 1, 0, 0, 0});
  c.push_back({"dgrad 3x3 s1 im2col", 1, 2, 14, 14, 64, 128, 3, 1, 1, 1, 1, 0, 0, 0});
  c.push_back({"dgrad 3x3 d2 im2col", 1, 2, 16, 16, 64, 64, 3, 1, 2, 1, 1, 0, 0, 0});
  c.push_back({"wgrad 1x1", 2, 2, 16, 16, 128, 128, 1, 1, 1, 0, 1, 0, 0, 0});
  c.push_back({"wgrad 1x1 Cin64 Cout256", 2, 2, 16, 16, 64, 256, 1, 1, 1, 0, 1, 0, 0, 0});
  c.push_back({"wgrad 3x3 s1 box", 2, 2, 14, 14, 64, 64, 3, 1, 1, 0, 1, 0, 0, 0});
  c.push_back({"wgrad 3x3 s1 im2col", 2, 2, 14, 14, 64, 64, 3, 1, 1, 1, 1, 0, 0, 0});
  c.push_back({"wgrad 3x3 s2 im2col", 2, 2, 28, 28, 64, 128, 3, 2, 1, 1, 1, 0, 0, 0});
  c.push_back({"wgrad 1x1 s2 box", 2, 2, 28, 28, 64, 128, 1, 2, 1, 0, 1, 0, 0, 0});
  // halo mode (a_mode 2): one TMA box per input halo tile, taps = start offsets into it
  c.push_back({"fprop 3x3 s1 halo 56x56 64->64 (W resident)", 0, 2, 56, 56, 64, 64, 3, 1, 1, 2, 1, 0, 0, 0});
  c.push_back({"fprop 3x3 s1 halo 28x28 128->128", 0, 2, 28, 28, 128, 128, 3, 1, 1, 2, 1, 1, 0, 0});
  c.push_back({"fprop 3x3 s1 halo 30x20 64->256 ragged", 0, 2, 30, 20, 64, 256, 3, 1, 1, 2, 1, 0, 0, 0});
  c.push_back({"fprop 3x3 d2 halo 32x32 64->64", 0, 1, 32, 32, 64, 64, 3, 1, 2, 2, 1, 0, 0, 0});
  c.push_back({"fprop 5x5 s1 halo 32x32 64->128", 0, 1, 32, 32, 64, 128, 5, 1, 1, 2, 1, 0, 0, 0});
  c.push_back({"fprop 3x3 VALID halo 34x34 64->64", 0, 1, 34, 34, 64, 64, 3, 1, 1, 2, 0, 0, 0, 0});
  c.push_back({"dgrad 3x3 s1 halo 56x56 64<-64", 1, 2, 56, 56, 64, 64, 3, 1, 1, 2, 1, 0, 0, 0});
  c.push_back({"dgrad 3x3 s1 halo 28x28 128<-128", 1, 2, 28, 28, 128, 128, 3, 1, 1, 2, 1, 0, 0, 0});
  c.push_back({"dgrad 5x5 s1 halo 32x32 64<-128", 1, 1, 32, 32, 64, 128, 5, 1, 1, 2, 1, 0, 0, 0});
  c.push_back({"T fprop 3x3 56x56 64->64 b64 halo", 0, 64, 56, 56, 64, 64, 3, 1, 1, 2, 1, 0, 0, 20});
  c.push_back({"T fprop 3x3 28x28 128->128 b64 halo", 0, 64, 28, 28, 128, 128, 3, 1, 1, 2, 1, 0, 0, 20});
  c.push_back({"T fprop 3x3 28x28 128->128 b64 im2col", 0, 64, 28, 28, 128, 128, 3, 1, 1, 1, 1, 0, 0, 20});
  c.push_back({"T dgrad 3x3 56x56 64<-64 b64 halo", 1, 64, 56, 56, 64, 64, 3, 1, 1, 2, 1, 0, 0, 20});
  // timing cases (ResNet-50 shapes at batch 64)
  c.push_back({"T fprop 1x1 56x56 256->64 b64", 0, 64, 56, 56, 256, 64, 1, 1, 1, 0, 1, 0, 0, 20});
  c.push_back({"T fprop 1x1 14x14 1024->256 b64", 0, 64, 14, 14, 1024, 256, 1, 1, 1, 0, 1, 0, 0, 20});
  c.push_back({"T fprop 3x3 14x14 256->256 b64 box", 0, 64, 14, 14, 256, 256, 3, 1, 1, 0, 1, 0, 0, 20});
  c.push_back({"T fprop 3x3 14x14 256->256 b64 im2col", 0, 64, 14, 14, 256, 256, 3, 1, 1, 1, 1, 0, 0, 20});
  c.push_back({"T fprop 3x3 56x56 64->64 b64 im2col", 0, 64, 56, 56, 64, 64, 3, 1, 1, 1, 1, 0, 0, 20});
  c.push_back({"T dgrad 3x3 14x14 256->256 b64 im2col", 1, 64, 14, 14, 256, 256, 3, 1, 1, 1, 1, 0, 0, 20});
  c.push_back({"T wgrad 3x3 14x14 256->256 b64 im2col", 2, 64, 14, 14, 256, 256, 3, 1, 1, 1, 1, 0, 0, 20});
  c.push_back({"T wgrad 1x1 56x56 64->256 b64", 2, 64, 56, 56, 64, 256, 1, 1, 1, 0, 1, 0, 0, 20});
  // halo wgrad (a_mode 2): pair mode (Cin 64) and group mode (Cin >= 128)
  c.push_back({"wgrad 3x3 halo pair 56x56 64->64", 2, 2, 56, 56, 64, 64, 3, 1, 1, 2, 1, 0, 0, 0});
  c.push_back({"wgrad 3x3 halo pair 30x20 64->128 ragged", 2, 3, 30, 20, 64, 128, 3, 1, 1, 2, 1, 0, 0, 0});
  c.push_back({"wgrad 3x3 halo pair d2 32x32 64->64", 2, 1, 32, 32, 64, 64, 3, 1, 2, 2, 1, 0, 0, 0});
  c.push_back({"wgrad 3x3 halo group 28x28 128->128", 2, 2, 28, 28, 128, 128, 3, 1, 1, 2, 1, 0, 0, 0});
  c.push_back({"wgrad 3x3 halo group 14x14 256->256", 2, 3, 14, 14, 256, 256, 3, 1, 1, 2, 1, 0, 0, 0});
  c.push_back({"wgrad 3x3 halo group 30x20 192->64 ragged", 2, 2, 30, 20, 192, 64, 3, 1, 1, 2, 1, 0, 0, 0});
  c.push_back({"wgrad 3x3 halo group VALID 34x34 128->128", 2, 1, 34, 34, 128, 128, 3, 1, 1, 2, 0, 0, 0, 0});
  c.push_back({"T wgrad 3x3 56x56 64->64 b64 halo", 2, 64, 56, 56, 64, 64, 3, 1, 1, 2, 1, 0, 0, 20});
  c.push_back({"T wgrad 3x3 56x56 64->64 b64 im2col", 2, 64, 56, 56, 64, 64, 3, 1, 1, 1, 1, 0, 0, 20});
  c.push_back({"T wgrad 3x3 28x28 128->128 b64 halo", 2, 64, 28, 28, 128, 128, 3, 1, 1, 2, 1, 0, 0, 20});
  c.push_back({"T wgrad 3x3 28x28 128->128 b64 im2col", 2, 64, 28, 28, 128, 128, 3, 1, 1, 1, 1, 0, 0, 20});
  c.push_back({"T wgrad 3x3 14x14 256->256 b64 halo", 2, 64, 14, 14, 256, 256, 3, 1, 1, 2, 1, 0, 0, 20});
  // weight-stationary mode of the GEMM kernel needs >= 2 tiles per SM: larger batches
  c.push_back({"fprop 1x1 56x56 64->256 N16 (W stationary)", 0, 16, 56, 56, 64, 256, 1, 1, 1, 0, 1, 1, 0, 0});
  c.push_back({"dgrad 1x1 56x56 256<-64 N16 (W stationary)", 1, 16, 56, 56, 256, 64, 1, 1, 1, 0, 1, 0, 0, 0});
  c.push_back({"fprop 1x1 28x28 128->512 N64 (2 n-tiles)", 0, 64, 28, 28, 128, 512, 1, 1, 1, 0, 1, 0, 0, 0});
  c.push_back({"T fprop 1x1 56x56 64->256 b64", 0, 64, 56, 56, 64, 256, 1, 1, 1, 0, 1, 0, 0, 20});
  c.push_back({"T dgrad 1x1 56x56 256<-64 b64", 1, 64, 56, 56, 256, 64, 1, 1, 1, 0, 1, 0, 0, 20});
  c.push_back({"T fprop 1x1 28x28 128->512 b64", 0, 64, 28, 28, 128, 512, 1, 1, 1, 0, 1, 0, 0, 20});
  c.push_back({"T fprop 1x1 28x28 512->128 b64", 0, 64, 28, 28, 512, 128, 1, 1, 1, 0, 1, 0, 0, 20});
  return c;
}

int main(int argc, char** argv) {
  std::vector<Case> cs = cases();
  if (argc < 2 || !strcmp(argv[1], "list")) {
    printf("%zu\n", cs.size());
    return 0;
  }
  int idx = atoi(argv[1]);
  if (idx < 0 || idx >= (int)cs.size()) return 1;
  Case c = cs[idx];
  mcn_conv_desc d;
  d.N = c.N; d.H = c.H; d.W = c.W; d.Cin = c.Cin; d.Cout = c.Cout;
  d.kh = d.kw = c.k; d.sh = d.sw = c.s; d.dh = d.dw = c.d;
  same_pad(c.H, c.k, c.s, c.d, c.same, &d.Ho, &d.pad_t);
  same_pad(c.W, c.k, c.s, c.d, c.same, &d.Wo, &d.pad_l);
  const int taps = c.k * c.k;
  const size_t nx = (size_t)c.N * c.H * c.W * c.Cin;
  const size_t ny = (size_t)c.N * d.Ho * d.Wo * c.Cout;
  const size_t nw = (size_t)taps * c.Cin * c.Cout;
  std::vector<float> x(nx), w(nw), dy(ny), bias(c.Cout);
  for (auto& v : x) v = bf16r(frand());
  for (auto& v : w) v = bf16r(frand() * 0.25f);   // HWIO
  for (auto& v : dy) v = bf16r(frand());
  for (auto& v : bias) v = frand();
  std::vector<__nv_bfloat16> xb(nx), dyb(ny), w_hwio(nw), w_ohwi(nw);
  for (size_t i = 0; i < nx; ++i) xb[i] = __float2bfloat16_rn(x[i]);
  for (size_t i = 0; i < ny; ++i) dyb[i] = __float2bfloat16_rn(dy[i]);
  for (int t = 0; t < taps; ++t)
    for (int ci = 0; ci < c.Cin; ++ci)
      for (int co = 0; co < c.Cout; ++co) {
        float v = w[((size_t)t * c.Cin + ci) * c.Cout + co];
        w_hwio[((size_t)t * c.Cin + ci) * c.Cout + co] = __float2bfloat16_rn(v);
        w_ohwi[((size_t)t * c.Cout + co) * c.Cin + ci] = __float2bfloat16_rn(v);
      }
  void *dx_, *dw1, *dw2, *dy_, *dout;
  float* dbias;
  CK(cudaMalloc(&dx_, nx * 2));
  CK(cudaMalloc(&dy_, ny * 2));
  CK(cudaMalloc(&dw1, nw * 2));
  CK(cudaMalloc(&dw2, nw * 2));
  CK(cudaMalloc(&dbias, c.Cout * 4));
  size_t nout = c.op == 0 ? ny : (c.op == 1 ? nx : nw);
  CK(cudaMalloc(&dout, nout * 4));
  CK(cudaMemcpy(dx_, xb.data(), nx * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dy_, dyb.data(), ny * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw1, w_hwio.data(), nw * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dw2, w_ohwi.data(), nw * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, bias.data(), c.Cout * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dout, 0, nout * 4));

  auto run = [&]() -> int {
    if (c.op == 0)
      return mcn_conv2d_fprop_tc(&d, dx_, dw2, c.bias ? dbias : nullptr, dout,
                                 c.out_f32 ? MCN_F32 : MCN_BF16, c.a_mode, 0, 0);
    if (c.op == 1) return mcn_conv2d_dgrad_tc(&d, dy_, dw1, dout, MCN_BF16, c.a_mode, 0, 0);
    return mcn_conv2d_wgrad_tc(&d, dx_, dy_, (float*)dout, c.a_mode, 0);
  };
  int rc = run();
  if (rc != 0) {
    printf("[%2d] %-46s ERROR rc=%d: %s\n", idx, c.name, rc, mcn_last_error());
    return 3;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("[%2d] %-46s DEVICE FAULT: %s\n", idx, c.name, cudaGetErrorString(e));
    return 4;
  }
  if (c.timing > 0) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) { if (c.op == 2) CK(cudaMemsetAsync(dout, 0, nout * 4)); run(); }
    CK(cudaEventRecord(e0));
    for (int i = 0; i < c.timing; ++i) run();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= c.timing;
    double flops = 2.0 * c.N * d.Ho * d.Wo * (double)taps * c.Cin * c.Cout;
    double bytes = 2.0 * (nx + ny) + 2.0 * nw;
    printf("[%2d] %-46s %.3f ms  %.1f TFLOP/s  %.0f GB/s (algorithmic)\n", idx, c.name, ms,
           flops / ms * 1e-9, bytes / ms * 1e-6);
    return 0;
  }
  // ---- CPU reference
  std::vector<float> got(nout);
  if ((c.op == 0 && !c.out_f32) || c.op == 1) {
    std::vector<__nv_bfloat16> tmp(nout);
    CK(cudaMemcpy(tmp.data(), dout, nout * 2, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < nout; ++i) got[i] = __bfloat162float(tmp[i]);
  } else {
    CK(cudaMemcpy(got.data(), dout, nout * 4, cudaMemcpyDeviceToHost));
  }
  std::vector<double> ref(nout, 0.0);
  for (int n = 0; n < c.N; ++n)
    for (int p = 0; p < d.Ho; ++p)
      for (int q = 0; q < d.Wo; ++q)
        for (int r = 0; r < c.k; ++r)
          for (int s = 0; s < c.k; ++s) {
            int h = p * c.s + r * c.d - d.pad_t, ww = q * c.s + s * c.d - d.pad_l;
            if (h < 0 || h >= c.H || ww < 0 || ww >= c.W) continue;
            const float* xp = &x[(((size_t)n * c.H + h) * c.W + ww) * c.Cin];
            const float* dyp = &dy[(((size_t)n * d.Ho + p) * d.Wo + q) * c.Cout];
            const float* wp = &w[(size_t)(r * c.k + s) * c.Cin * c.Cout];
            if (c.op == 0) {
              double* o = &ref[(((size_t)n * d.Ho + p) * d.Wo + q) * c.Cout];
              for (int ci = 0; ci < c.Cin; ++ci)
                for (int co = 0; co < c.Cout; ++co) o[co] += (double)xp[ci] * wp[ci * c.Cout + co];
            } else if (c.op == 1) {
              double* o = &ref[(((size_t)n * c.H + h) * c.W + ww) * c.Cin];
              for (int ci = 0; ci < c.Cin; ++ci)
                for (int co = 0; co < c.Cout; ++co) o[ci] += (double)dyp[co] * wp[ci * c.Cout + co];
            } else {
              double* o = &ref[(size_t)(r * c.k + s) * c.Cin * c.Cout];
              for (int ci = 0; ci < c.Cin; ++ci)
                for (int co = 0; co < c.Cout; ++co) o[ci * c.Cout + co] += (double)xp[ci] * dyp[co];
            }
          }
  if (c.op == 0 && c.bias)
    for (size_t i = 0; i < nout; ++i) ref[i] += bias[i % c.Cout];
  double max_err = 0, max_ref = 0;
  size_t worst = 0;
  for (size_t i = 0; i < nout; ++i) {
    double e2 = fabs(got[i] - ref[i]);
    if (e2 > max_err) { max_err = e2; worst = i; }
    if (fabs(ref[i]) > max_ref) max_ref = fabs(ref[i]);
  }
  bool pass = max_err <= 1e-2 * max_ref + 1e-3 && max_ref > 0;
  printf("[%2d] %-46s %s  max_err %.4g  max_ref %.4g  (worst idx %zu got %.5g ref %.5g)\n", idx,
         c.name, pass ? "PASS" : "FAIL", max_err, max_ref, worst, got[worst], ref[worst]);
  if (!pass) {
    // first few mismatches help diagnose layout errors
    int shown = 0;
    for (size_t i = 0; i < nout && shown < 6; ++i)
      if (fabs(got[i] - ref[i]) > 1e-2 * max_ref + 1e-3) {
        printf("      idx %zu got %.5g ref %.5g\n", i, got[i], ref[i]);
        ++shown;
      }
  }
  return pass ? 0 : 5;
}
