// Bandwidth experiments for the streaming (batch-norm apply-like) kernels: y = relu(x*sc + sf) over a
// bf16 tensor with C channels.  Compares traversal orders, unroll depths, vector widths, cache
// hints and a TMA-bulk shared-memory ring, on tensors of 25 / 103 / 411 MB, both "cold" (L2
// flushed) and right after a producer kernel wrote x (what the training step does).
// Not a test: prints a table.  Usage: stream_harness [reps]
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e_ = (x);                                                              \
    if (e_ != cudaSuccess) {                                                           \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)

__device__ __forceinline__ uint4 ld_nc(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_cs(const void* p) {
  uint4 r;
  asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_cs(void* p, uint4 v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t fma_relu2(uint32_t w, float sc0, float sf0, float sc1, float sf1) {
  float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xFFFF0000u);
  lo = fmaxf(fmaf(lo, sc0, sf0), 0.f);
  hi = fmaxf(fmaf(hi, sc1, sf1), 0.f);
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 apply8(uint4 a, const float* sc, const float* sf) {
  uint4 o;
  o.x = fma_relu2(a.x, sc[0], sf[0], sc[1], sf[1]);
  o.y = fma_relu2(a.y, sc[2], sf[2], sc[3], sf[3]);
  o.z = fma_relu2(a.z, sc[4], sf[4], sc[5], sf[5]);
  o.w = fma_relu2(a.w, sc[6], sf[6], sc[7], sf[7]);
  return o;
}
__device__ __forceinline__ void load_coef(const float* scale, const float* shift, int c0, float* sc, float* sf) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = scale[c0 + i];
    sf[i] = shift[c0 + i];
  }
}

// ---- grid-stride persistent; HINT 0 = nc/no_allocate loads + default stores, 1 = .cs loads,
// 2 = .cs loads and .cs stores.  REV walks from the end of the tensor.
template <int U, int REV, int HINT>
__global__ void __launch_bounds__(512) apply_gs(const uint4* __restrict__ x, uint4* __restrict__ y, long long nvec,
                                                int cv, const float* scale, const float* shift) {
  float sc[8], sf[8];
  load_coef(scale, shift, (threadIdx.x % cv) * 8, sc, sf);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += U * stride) {
    uint4 a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      long long vv = v + u * stride;
      if (vv < nvec) {
        long long idx = REV ? (nvec - 1 - vv) : vv;
        a[u] = HINT == 0 ? ld_nc(x + idx) : ld_cs(x + idx);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      long long vv = v + u * stride;
      if (vv < nvec) {
        long long idx = REV ? (nvec - 1 - vv) : vv;
        // with REV the channel group of a vector is (idx % cv); keep coefficient registers valid by
        // requiring nvec % cv == 0 and blockDim % cv == 0 (then idx % cv == cv-1-(vv % cv))
        uint4 o = apply8(a[u], sc, sf);
        if (HINT == 2) st_cs(y + idx, o); else y[idx] = o;
      }
    }
  }
}

// ---- non-persistent: each block owns a contiguous run of block*U vectors
template <int U>
__global__ void __launch_bounds__(256) apply_blk(const uint4* __restrict__ x, uint4* __restrict__ y, long long nvec,
                                                 int cv, const float* scale, const float* shift) {
  float sc[8], sf[8];
  load_coef(scale, shift, (threadIdx.x % cv) * 8, sc, sf);
  const long long base = (long long)blockIdx.x * blockDim.x * U + threadIdx.x;
  uint4 a[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    long long vv = base + (long long)u * blockDim.x;
    if (vv < nvec) a[u] = ld_nc(x + vv);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    long long vv = base + (long long)u * blockDim.x;
    if (vv < nvec) y[vv] = apply8(a[u], sc, sf);
  }
}

// ---- persistent, blocked: every block owns ONE contiguous slice of the tensor and walks it
template <int U>
__global__ void __launch_bounds__(512) apply_slice(const uint4* __restrict__ x, uint4* __restrict__ y, long long nvec,
                                                   int cv, const float* scale, const float* shift) {
  float sc[8], sf[8];
  load_coef(scale, shift, (threadIdx.x % cv) * 8, sc, sf);
  // slice boundaries are multiples of blockDim so the channel group of a thread is fixed
  const long long per = ((nvec + gridDim.x - 1) / gridDim.x + blockDim.x - 1) / blockDim.x * blockDim.x;
  const long long v0 = per * blockIdx.x, v1 = min(nvec, v0 + per);
  for (long long v = v0 + threadIdx.x; v < v1; v += (long long)U * blockDim.x) {
    uint4 a[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      long long vv = v + (long long)u * blockDim.x;
      if (vv < v1) a[u] = ld_nc(x + vv);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      long long vv = v + (long long)u * blockDim.x;
      if (vv < v1) y[vv] = apply8(a[u], sc, sf);
    }
  }
}

// ---- TMA bulk ring: thread 0 streams 16 KB chunks global -> shared (cp.async.bulk + mbarrier),
// all threads consume from shared memory.  Bytes in flight are decoupled from registers.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int STAGES, int CHUNK>
__global__ void __launch_bounds__(256) apply_bulk(const uint4* __restrict__ x, uint4* __restrict__ y, long long nvec,
                                                  int cv, const float* scale, const float* shift) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(sm + (size_t)STAGES * CHUNK);
  float sc[8], sf[8];
  load_coef(scale, shift, (threadIdx.x % cv) * 8, sc, sf);
  constexpr int VPC = CHUNK / 16;           // vectors per chunk (multiple of blockDim)
  const long long nchunks = (nvec + VPC - 1) / VPC;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](long long chunk, int s) {
    long long v0 = chunk * VPC;
    uint32_t bytes = (uint32_t)(min((long long)VPC, nvec - v0) * 16);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[s])), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(sm + (size_t)s * CHUNK)),
                 "l"(x + v0), "r"(bytes), "r"(smem_u32(&bar[s]))
                 : "memory");
  };
  long long c = blockIdx.x;
  if (threadIdx.x == 0) {
    long long cc = c;
    for (int s = 0; s < STAGES && cc < nchunks; ++s, cc += gridDim.x) issue(cc, s);
  }
  int it = 0;
  for (; c < nchunks; c += gridDim.x, ++it) {
    const int s = it % STAGES;
    const uint32_t ph = (it / STAGES) & 1;
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(&bar[s])), "r"(ph) : "memory");
    }
    const uint4* src = reinterpret_cast<const uint4*>(sm + (size_t)s * CHUNK);
    const long long v0 = c * VPC;
#pragma unroll
    for (int k = 0; k < VPC / 256; ++k) {
      int i = k * 256 + threadIdx.x;
      if (v0 + i < nvec) y[v0 + i] = apply8(src[i], sc, sf);
    }
    __syncthreads();     // everyone has read stage s
    if (threadIdx.x == 0) {
      long long nxt = c + (long long)STAGES * gridDim.x;
      if (nxt < nchunks) issue(nxt, s);
    }
  }
}

// ---- multi-stream variants: NR read streams, NW write streams (bn_apply+residual = 2R1W,
// bn_bwd_reduce = 2R0W / 3R0W, bn_bwd_apply = 3R2W).  The arithmetic is a stand-in (sum of the inputs).
template <int NR, int NW, int U, int BLK>
__global__ void __launch_bounds__(BLK) multi_gs(const uint4* __restrict__ a0, const uint4* __restrict__ a1,
                                                const uint4* __restrict__ a2, uint4* __restrict__ o0,
                                                uint4* __restrict__ o1, long long nvec, float* sink) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  float acc = 0.f;
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += U * stride) {
    uint4 x[U], y[U], z[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      long long vv = v + u * stride;
      if (vv < nvec) {
        x[u] = ld_nc(a0 + vv);
        if (NR > 1) y[u] = ld_nc(a1 + vv);
        if (NR > 2) z[u] = ld_nc(a2 + vv);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      long long vv = v + u * stride;
      if (vv < nvec) {
        uint4 r = x[u];
        if (NR > 1) { r.x ^= y[u].x; r.y += y[u].y; r.z ^= y[u].z; r.w += y[u].w; }
        if (NR > 2) { r.x += z[u].x; r.y ^= z[u].y; r.z += z[u].z; r.w ^= z[u].w; }
        if (NW > 0) o0[vv] = r;
        if (NW > 1) o1[vv] = make_uint4(r.y, r.x, r.w, r.z);
        if (NW == 0) acc += __uint_as_float(r.x & 0x3FFFFFFFu);
      }
    }
  }
  if (NW == 0 && acc == 12345.678f) *sink = acc;
}
template <int NR, int NW, int U>
__global__ void __launch_bounds__(256) multi_runs(const uint4* __restrict__ a0, const uint4* __restrict__ a1,
                                                  const uint4* __restrict__ a2, uint4* __restrict__ o0,
                                                  uint4* __restrict__ o1, long long nvec, float* sink) {
  const long long base = (long long)blockIdx.x * 256 * U + threadIdx.x;
  uint4 x[U], y[U], z[U];
  float acc = 0.f;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    long long vv = base + u * 256;
    if (vv < nvec) {
      x[u] = ld_nc(a0 + vv);
      if (NR > 1) y[u] = ld_nc(a1 + vv);
      if (NR > 2) z[u] = ld_nc(a2 + vv);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    long long vv = base + u * 256;
    if (vv < nvec) {
      uint4 r = x[u];
      if (NR > 1) { r.x ^= y[u].x; r.y += y[u].y; r.z ^= y[u].z; r.w += y[u].w; }
      if (NR > 2) { r.x += z[u].x; r.y ^= z[u].y; r.z += z[u].z; r.w ^= z[u].w; }
      if (NW > 0) o0[vv] = r;
      if (NW > 1) o1[vv] = make_uint4(r.y, r.x, r.w, r.z);
      if (NW == 0) acc += __uint_as_float(r.x & 0x3FFFFFFFu);
    }
  }
  if (NW == 0 && acc == 12345.678f) *sink = acc;
}

// producer stand-in: writes x sequentially with 32-byte stores per thread (like the conv epilogue)
__global__ void writer(uint4* __restrict__ x, long long nvec, uint32_t seed) {
  for (long long v = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 2; v + 1 < nvec;
       v += (long long)gridDim.x * blockDim.x * 2) {
    uint32_t a = (uint32_t)v * 2654435761u + seed;
    uint32_t w = 0x3F803F80u ^ ((a & 0x007F007Fu));   // bf16 values near 1.0
    x[v] = make_uint4(w, w, w, w);
    x[v + 1] = make_uint4(w, w, w, w);
  }
}
__global__ void flush(uint4* p, long long n) {
  for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < n; v += (long long)gridDim.x * blockDim.x)
    p[v] = make_uint4(1, 2, 3, 4);
}

int main(int argc, char** argv) {
  int reps = argc > 1 ? atoi(argv[1]) : 5;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int C = 256, cv = C / 8;
  const long long sizes[3] = {25690112LL / 16, 102760448LL / 16, 411041792LL / 16};   // vectors
  uint4 *x, *y, *scratch;
  float *scale, *shift;
  const long long maxv = sizes[2];
  CK(cudaMalloc(&x, maxv * 16));
  CK(cudaMalloc(&y, maxv * 16));
  const long long nflush = (512LL << 20) / 16;
  CK(cudaMalloc(&scratch, nflush * 16));
  CK(cudaMalloc(&scale, C * 4));
  CK(cudaMalloc(&shift, C * 4));
  std::vector<float> h(C, 1.0f);
  CK(cudaMemcpy(scale, h.data(), C * 4, cudaMemcpyHostToDevice));
  for (auto& v : h) v = -0.5f;
  CK(cudaMemcpy(shift, h.data(), C * 4, cudaMemcpyHostToDevice));
  CK(cudaFuncSetAttribute(apply_bulk<6, 16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 16384 + 64));
  CK(cudaFuncSetAttribute(apply_bulk<4, 16384>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 16384 + 64));
  CK(cudaFuncSetAttribute(apply_bulk<3, 32768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 32768 + 64));

  struct Variant { const char* name; int id; };
  std::vector<Variant> vars = {
      {"gs U4 b512x2 nc (current)", 0}, {"gs U8 b512x2 nc", 1}, {"gs U2 b512x4 nc", 2}, {"gs U4 b512x2 REV", 3},
      {"gs U4 b512x2 .cs ld", 4}, {"gs U4 b512x2 .cs ld+st", 5}, {"gs U8 b512x2 .cs ld+st", 6},
      {"blk U4 (b256, non-persistent)", 7}, {"blk U8 (b256, non-persistent)", 8}, {"slice U4 b512x2", 9},
      {"slice U8 b512x2", 10}, {"bulk 6x16K b256x2", 11}, {"bulk 4x16K b256x3", 12}, {"bulk 3x32K b256x2", 13},
      {"gs U4 b512x1 nc", 14}, {"gs U8 b512x1 nc", 15},
  };
  auto launch = [&](int id, long long nvec) {
    switch (id) {
      case 0: apply_gs<4, 0, 0><<<sms * 2, 512>>>(x, y, nvec, cv, scale, shift); break;
      case 1: apply_gs<8, 0, 0><<<sms * 2, 512>>>(x, y, nvec, cv, scale, shift); break;
      case 2: apply_gs<2, 0, 0><<<sms * 4, 512>>>(x, y, nvec, cv, scale, shift); break;
      case 3: apply_gs<4, 1, 0><<<sms * 2, 512>>>(x, y, nvec, cv, scale, shift); break;
      case 4: apply_gs<4, 0, 1><<<sms * 2, 512>>>(x, y, nvec, cv, scale, shift); break;
      case 5: apply_gs<4, 0, 2><<<sms * 2, 512>>>(x, y, nvec, cv, scale, shift); break;
      case 6: apply_gs<8, 0, 2><<<sms * 2, 512>>>(x, y, nvec, cv, scale, shift); break;
      case 7: apply_blk<4><<<(unsigned)((nvec + 1023) / 1024), 256>>>(x, y, nvec, cv, scale, shift); break;
      case 8: apply_blk<8><<<(unsigned)((nvec + 2047) / 2048), 256>>>(x, y, nvec, cv, scale, shift); break;
      case 9: apply_slice<4><<<sms * 2, 512>>>(x, y, nvec, cv, scale, shift); break;
      case 10: apply_slice<8><<<sms * 2, 512>>>(x, y, nvec, cv, scale, shift); break;
      case 11: apply_bulk<6, 16384><<<sms * 2, 256, 6 * 16384 + 64>>>(x, y, nvec, cv, scale, shift); break;
      case 12: apply_bulk<4, 16384><<<sms * 3, 256, 4 * 16384 + 64>>>(x, y, nvec, cv, scale, shift); break;
      case 13: apply_bulk<3, 32768><<<sms * 2, 256, 3 * 32768 + 64>>>(x, y, nvec, cv, scale, shift); break;
      case 14: apply_gs<4, 0, 0><<<sms, 512>>>(x, y, nvec, cv, scale, shift); break;
      case 15: apply_gs<8, 0, 0><<<sms, 512>>>(x, y, nvec, cv, scale, shift); break;
    }
  };
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("%-34s | cold: 25MB 103MB 411MB | after writer: 25MB 103MB 411MB   (GB/s, read+write, best of %d)\n", "variant", reps);
  for (auto& v : vars) {
    double res[2][3];
    for (int cond = 0; cond < 2; ++cond)
      for (int si = 0; si < 3; ++si) {
        float best = 1e30f;
        for (int r = 0; r < reps + 1; ++r) {
          if (cond == 0) {
            writer<<<sms * 8, 256>>>(x, sizes[si], r);
            flush<<<sms * 8, 256>>>(scratch, nflush);
          } else {
            flush<<<sms * 8, 256>>>(scratch, nflush);
            writer<<<sms * 8, 256>>>(x, sizes[si], r);
          }
          CK(cudaEventRecord(e0));
          launch(v.id, sizes[si]);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaGetLastError());
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (r > 0 && ms < best) best = ms;
        }
        res[cond][si] = 2.0 * sizes[si] * 16 / (best * 1e-3) / 1e9;
      }
    printf("%-34s | %6.0f %6.0f %6.0f | %6.0f %6.0f %6.0f\n", v.name, res[0][0], res[0][1], res[0][2],
           res[1][0], res[1][1], res[1][2]);
    fflush(stdout);
  }
  // correctness spot check of the last variant against variant 0
  std::vector<uint32_t> a(4096), b(4096);
  launch(0, sizes[0]);
  CK(cudaMemcpy(a.data(), y, 4096 * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemset(y, 0, 4096 * 4));
  launch(11, sizes[0]);
  CK(cudaMemcpy(b.data(), y, 4096 * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < 4096; ++i) bad += a[i] != b[i];
  printf("bulk vs gs mismatches in first 16 KB: %d\n", bad);
  // ---- multi-stream table (103 MB and 411 MB tensors; all streams cold)
  {
    uint4 *b1, *b2, *o1;
    const long long nv[2] = {sizes[1], sizes[2]};
    CK(cudaMalloc(&b1, maxv * 16));
    CK(cudaMalloc(&b2, maxv * 16));
    CK(cudaMalloc(&o1, maxv * 16));
    CK(cudaMemset(b1, 1, maxv * 16));
    CK(cudaMemset(b2, 2, maxv * 16));
    struct MV { const char* name; int nr, nw, id; };
    std::vector<MV> mv = {
        {"2R1W gs U4 b512x2 (apply+res now)", 2, 1, 0}, {"2R1W gs U2 b512x2", 2, 1, 1}, {"2R1W runs U4", 2, 1, 2}, {"2R1W runs U8", 2, 1, 3},
        {"2R1W gs U4 b256x4", 2, 1, 4},
        {"2R0W gs U2 b256x3 (reduce now)", 2, 0, 5}, {"2R0W gs U4 b256x3", 2, 0, 6}, {"2R0W gs U4 b512x2", 2, 0, 7}, {"2R0W runs U8", 2, 0, 8},
        {"3R2W gs U2 b256x3 (bwd_apply now)", 3, 2, 9}, {"3R2W gs U4 b256x2", 3, 2, 10}, {"3R2W gs U2 b512x2", 3, 2, 11}, {"3R2W runs U4", 3, 2, 12},
        {"3R2W gs U1 b512x4", 3, 2, 13},
    };
    auto mlaunch = [&](int id, long long n) {
      switch (id) {
        case 0: multi_gs<2, 1, 4, 512><<<sms * 2, 512>>>(x, b1, b2, y, o1, n, shift); break;
        case 1: multi_gs<2, 1, 2, 512><<<sms * 2, 512>>>(x, b1, b2, y, o1, n, shift); break;
        case 2: multi_runs<2, 1, 4><<<(unsigned)((n + 1023) / 1024), 256>>>(x, b1, b2, y, o1, n, shift); break;
        case 3: multi_runs<2, 1, 8><<<(unsigned)((n + 2047) / 2048), 256>>>(x, b1, b2, y, o1, n, shift); break;
        case 4: multi_gs<2, 1, 4, 256><<<sms * 4, 256>>>(x, b1, b2, y, o1, n, shift); break;
        case 5: multi_gs<2, 0, 2, 256><<<sms * 3, 256>>>(x, b1, b2, y, o1, n, shift); break;
        case 6: multi_gs<2, 0, 4, 256><<<sms * 3, 256>>>(x, b1, b2, y, o1, n, shift); break;
        case 7: multi_gs<2, 0, 4, 512><<<sms * 2, 512>>>(x, b1, b2, y, o1, n, shift); break;
        case 8: multi_runs<2, 0, 8><<<(unsigned)((n + 2047) / 2048), 256>>>(x, b1, b2, y, o1, n, shift); break;
        case 9: multi_gs<3, 2, 2, 256><<<sms * 3, 256>>>(x, b1, b2, y, o1, n, shift); break;
        case 10: multi_gs<3, 2, 4, 256><<<sms * 2, 256>>>(x, b1, b2, y, o1, n, shift); break;
        case 11: multi_gs<3, 2, 2, 512><<<sms * 2, 512>>>(x, b1, b2, y, o1, n, shift); break;
        case 12: multi_runs<3, 2, 4><<<(unsigned)((n + 1023) / 1024), 256>>>(x, b1, b2, y, o1, n, shift); break;
        case 13: multi_gs<3, 2, 1, 512><<<sms * 4, 512>>>(x, b1, b2, y, o1, n, shift); break;
      }
    };
    printf("%-36s | 103MB 411MB per stream   (GB/s over all streams, cold, best of %d)\n", "multi-stream variant", reps);
    for (auto& m : mv) {
      double res[2];
      for (int si = 0; si < 2; ++si) {
        float best = 1e30f;
        for (int r = 0; r < reps + 1; ++r) {
          flush<<<sms * 8, 256>>>(scratch, nflush);
          CK(cudaEventRecord(e0));
          mlaunch(m.id, nv[si]);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaGetLastError());
          float ms;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          if (r > 0 && ms < best) best = ms;
        }
        res[si] = (double)(m.nr + m.nw) * nv[si] * 16 / (best * 1e-3) / 1e9;
      }
      printf("%-36s | %6.0f %6.0f\n", m.name, res[0], res[1]);
      fflush(stdout);
    }
  }
  return 0;
}
