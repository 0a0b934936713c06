// Dense convolution on the 5th-generation tensor cores (sm_100a).
//
// Replaces tf.nn.conv2d / Conv2DBackpropInput / Conv2DBackpropFilter / conv2d_transpose / matmul
// as called from reference convnet.py:1659, :2463, :1743 (see include/mcn.h).
//
// Formulation: implicit GEMM, one filter tap at a time.
//   fprop : Y[pixel, co]  = sum_tap sum_ci X[pixel shifted by tap, ci] * W[tap, co, ci]
//   dgrad : dX[pixel, ci] = sum_tap sum_co dY[pixel shifted by -tap, co] * W[tap, ci, co]
//   wgrad : dW[tap, ci, co] = sum_pixel X[pixel shifted by tap, ci] * dY[pixel, co]
// fprop and dgrad share one kernel (A = activations, K-major; B = weights, K-major);
// wgrad has its own (both operands MN-major, K = pixels, deterministic split-K through workspace slices).
//
// Data movement: every operand tile is brought in by TMA into 128-byte-swizzled shared memory.
// The shifted activation window of a tap is either a 4-D TMA *box* (a_mode 0: the tile is a
// TW x TH x TN block of output pixels, out-of-bounds rows/cols zero-filled = the padding) or
// a TMA *im2col* load (a_mode 1: the tile is 128 consecutive output pixels in n,h,w order; the
// hardware walks the window base with the conv stride and the tap offset carries the dilation).
// MMA: tcgen05.mma cta_group::1, M=128, N=block_n, K=16 per instruction, accumulators in TMEM.
// Roles per CTA (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
// warps 2..5 = epilogue (tcgen05.ld -> registers -> global).
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "mcn_common.cuh"
#include "sm100_ptx.cuh"
#include "xsum.cuh"

namespace mcn {
namespace {

// Per-role stall accounting (build with -DMCN_ROLE_TIMING, scripts/role_timing.py): cycles each
// role spends waiting on its barriers, summed over the CTAs of every launch since the last reset.
//   0 producer waits for a free stage   1 MMA waits for operands      2 MMA waits for a free accumulator
//   3 epilogue waits for an accumulator 4 epilogue total              5 CTA lifetime   6 CTAs   7 MMA total
#ifdef MCN_ROLE_TIMING
__device__ unsigned long long g_role_cycles[16];
#define RT_DECL unsigned long long rt_acc = 0, rt_acc2 = 0, rt_t0 = 0; (void)rt_acc2; (void)rt_t0
#define RT_BEGIN rt_t0 = clock64()
#define RT_END rt_acc += clock64() - rt_t0
#define RT_END2 rt_acc2 += clock64() - rt_t0
#define RT_FLUSH(slot) atomicAdd(&g_role_cycles[slot], rt_acc)
#define RT_FLUSH2(slot) atomicAdd(&g_role_cycles[slot], rt_acc2)
#define RT_ADD(slot, v) atomicAdd(&g_role_cycles[slot], static_cast<unsigned long long>(v))
#define RT_NOW() clock64()
#else
#define RT_DECL
#define RT_BEGIN
#define RT_END
#define RT_END2
#define RT_FLUSH(slot)
#define RT_FLUSH2(slot)
#define RT_ADD(slot, v)
#define RT_NOW() 0ll
#endif

constexpr int kMaxTaps = 52;
constexpr int kABytes = 128 * 128;  // one activation tile: 128 rows x 64 bf16

struct TapTab {
  int brow[kMaxTaps];   // row of this tap's weight slab in the B matrix
  short dh[kMaxTaps];   // tiled: spatial shift; im2col: tap offset (>= 0)
  short dw[kMaxTaps];
  signed char map[kMaxTaps];
};

struct TileGeom {
  int a_mode;
  int tiles_w, tiles_h;     // tiled: tile grid inside (W, H); tiles along N follow
  int TW, TH, TN, rows_box;  // tiled: box extents, rows_box = TW*TH*TN <= 128
  int Wo, Ho, Nb;            // logical pixel-space extents
  int str_w, str_h;          // im2col: traversal stride
  int low_w, low_h;          // im2col: lower corner (window base of pixel 0)
  long long m_total;         // Nb*Ho*Wo
};

struct PixelTile {
  int w0, h0, n0;   // tiled: box origin; im2col: window base (w,h) and image
  long long m0;     // im2col: first linear pixel
};

__device__ __forceinline__ PixelTile decode_tile(const TileGeom& g, int m_t) {
  PixelTile t;
  if (g.a_mode == 0) {
    int wt = m_t % g.tiles_w;
    int r = m_t / g.tiles_w;
    int ht = r % g.tiles_h;
    int nt = r / g.tiles_h;
    t.w0 = wt * g.TW;
    t.h0 = ht * g.TH;
    t.n0 = nt * g.TN;
    t.m0 = 0;
  } else {
    t.m0 = static_cast<long long>(m_t) * 128;
    int q0 = static_cast<int>(t.m0 % g.Wo);
    long long r = t.m0 / g.Wo;
    int p0 = static_cast<int>(r % g.Ho);
    t.n0 = static_cast<int>(r / g.Ho);
    t.w0 = q0 * g.str_w + g.low_w;
    t.h0 = p0 * g.str_h + g.low_h;
  }
  return t;
}

// Row `row` (0..127) of a pixel tile -> (n, p, q) and validity.
__device__ __forceinline__ bool row_coords(const TileGeom& g, const PixelTile& t, int row, int& n,
                                           int& p, int& q) {
  if (g.a_mode == 0) {
    int wi = row % g.TW;
    int r = row / g.TW;
    int hi = r % g.TH;
    int ni = r / g.TH;
    q = t.w0 + wi;
    p = t.h0 + hi;
    n = t.n0 + ni;
    return row < g.rows_box && q < g.Wo && p < g.Ho && n < g.Nb;
  }
  long long m = t.m0 + row;
  q = static_cast<int>(m % g.Wo);
  long long r = m / g.Wo;
  p = static_cast<int>(r % g.Ho);
  n = static_cast<int>(r / g.Ho);
  return m < g.m_total;
}

// ------------------------------------------------------------------ shared epilogue
struct EpiArgs {
  void* out;
  const float* bias;
  double* stats;   // fused batch-norm statistics: [2*n_total] fp64 (sum y, sum y^2), or nullptr
  long long* xs;   // exact accumulator limbs [3][2*n_total] in the workspace (xsum.cuh)
  unsigned int* xs_counter;   // one tile counter per n-tile
  int tiles_m;     // tiles per n-tile
  int out_f32, vec_ok, accumulate, block_n, n_total;
  int debug;   // bit0: skip stores, bit2: skip TMEM loads (timing experiments, direct epilogue only)
  int stg_bufs;    // TMA-store epilogue: 4 KB staging tiles per epilogue warp (1 or 2)
  int out_rank4;   // TMA-store epilogue: the output map is (C, W, H, N) instead of [M, C]
  // fused batch-norm BACKWARD reduction (dgrad whose output is the gradient of a BN+ReLU output):
  // stats then receives [sum dz | sum dz*x] with dz = dy * relu'(x*sc + sf), x = the BN input
  int red;                 // 0 off; otherwise 1 + activation code (1 = none, 2 = relu)
  const float* red_mean;   // saved mean / invstd of the forward pass, gamma / beta (may be NULL)
  const float* red_invstd;
  const float* red_gamma;
  const float* red_beta;
  // when set, the block that completes a channel group writes the FINAL backward sums itself:
  // red_out1[c] += sum dz, red_out2[c] += invstd*(sum dz*x - mean*sum dz)  (no finalize launch)
  float* red_out1;
  float* red_out2;
};

// Fused BN statistics (conv -> BN): the epilogue already holds every output element in
// registers, so the per-channel sum and sum of squares of the STORED (bf16-rounded) values are
// taken there and the separate statistics pass over y disappears.  A thread owns one pixel row,
// a channel's values are spread over the lanes: each epilogue warp parks its 32 rows x 64 packed
// bf16 channels in a private shared-memory tile (row pitch 36 words: conflict-free 16-byte row
// stores and conflict-free column reads), lane L then sums channels 2L and 2L+1 down the 32
// rows.  Partial sums go to a per-WARP shared accumulator (one n-tile wide, plain read-modify-
// write: no shared-memory float atomics, which are CAS loops) and are flushed into the exact
// fixed-point accumulator of xsum.cuh (integer atomics: the result does not depend on the order
// in which CTAs arrive) only when the persistent CTA moves to another n-tile or finishes; the
// flush that completes an n-tile decodes its limbs into the caller's fp64 sums (stats_flush).
constexpr int kStatRowWords = 36;
constexpr int kStatStageWords = 4 * 32 * kStatRowWords;          // four epilogue warps
constexpr int kStatAccWarp = 2 * 256;                            // [sum | sum sq] x block_n <= 256
constexpr int kStatAccFloats = 4 * kStatAccWarp;                 // one slice per epilogue warp
constexpr int kEpiStageBytes = kStatStageWords * 4;              // always: the store path stages through it
constexpr int kStatAccBytes = kStatAccFloats * 4;                // only with fused statistics
constexpr int kBarRegionBytes = 512;                             // mbarriers + tmem slot, halo tap table at +256

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// et: index of the thread among the 128 epilogue threads; `pending` = tiles of n-tile n_t whose
// sums this flush carries.  The ticket counter of an n-tile counts TILES: the flush that
// completes the n-tile's tiles_m tiles decodes just that n-tile's channels into the caller's fp64
// sums and clears limbs and counter — no grid-wide serial tail, and the workspace is zero again
// when the kernel ends.  `flag` is a shared-memory word visible to all epilogue threads.
// ticket == false: an intermediate flush (every 8 tiles, bounding the length of the fp32 partial sums):
// only the exact adds — no fence, no ticket; the tiles stay in the caller's `pending` count and the
// next ticketed flush (whose __threadfence orders ALL of this thread's earlier adds: thread et always
// owns the same columns) reports them.
__device__ __forceinline__ void stats_flush(const EpiArgs& e, float* acc, int n_t, int et, int pending,
                                            int* flag, bool ticket = true) {
  epi_bar_sync();   // every warp's shared-memory sums of the finished tiles have landed
  for (int c = et; c < e.block_n; c += 128) {
    const int col = n_t * e.block_n + c;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      s1 += acc[w * kStatAccWarp + c];
      s2 += acc[w * kStatAccWarp + 256 + c];
      acc[w * kStatAccWarp + c] = 0.f;
      acc[w * kStatAccWarp + 256 + c] = 0.f;
    }
    if (col < e.n_total) {
      xs::add(e.xs, 2 * e.n_total, col, s1);
      xs::add(e.xs, 2 * e.n_total, e.n_total + col, s2);
    }
  }
  if (!ticket) {
    epi_bar_sync();   // the partial sums are cleared before any warp accumulates again
    return;
  }
  __threadfence();   // the adds are ordered before the ticket
  epi_bar_sync();
  if (et == 0) {
    const unsigned int old = atomicAdd(e.xs_counter + n_t, static_cast<unsigned int>(pending));
    __threadfence();
    *flag = (old + static_cast<unsigned int>(pending) == static_cast<unsigned int>(e.tiles_m)) ? 1 : 0;
  }
  epi_bar_sync();
  if (*flag) {
    const int c0 = n_t * e.block_n;
    const int w = min(e.n_total - c0, e.block_n);
    if (e.red && e.red_out1 != nullptr) {
      // same arithmetic as bn_bwd_finalize_kernel (bn.cu), on the exact sums
      for (int i = et; i < w; i += 128) {
        const int c = c0 + i;
        const double s1 = xs::read_clear(e.xs, 2 * e.n_total, c);
        const double s2 = xs::read_clear(e.xs, 2 * e.n_total, e.n_total + c);
        e.red_out1[c] += static_cast<float>(s1);
        e.red_out2[c] += static_cast<float>(static_cast<double>(e.red_invstd[c]) *
                                            (s2 - static_cast<double>(e.red_mean[c]) * s1));
      }
    } else {
      for (int i = et; i < 2 * w; i += 128) {
        const int which = i >= w ? 1 : 0;
        const int idx = which * e.n_total + c0 + i - which * w;
        const double prev = e.stats[idx];
        e.stats[idx] = prev + xs::read_clear(e.xs, 2 * e.n_total, idx);
      }
    }
    if (et == 0) xs::release(e.xs_counter + n_t);
  }
  epi_bar_sync();
}

// (A shared-memory staged, fully coalesced write-out was tried in round 1 and measured SLOWER on
// the store-heavy 1x1 convolutions (24.8 vs 24.5 ms/step): ncu shows the epilogue warps bound by
// their own instruction stream (~1500 instructions per 128x256 tile at one warp per scheduler,
// 5.8 cycles per issued instruction), not by LSU wavefronts, so adding instructions loses.  The
// lever is more epilogue warps per TMEM quadrant / fewer instructions per element.)
// One output tile: wait for the accumulator, TMEM -> registers -> global.  Each thread owns one
// output pixel (row) and walks its channels 64 at a time: 64 bf16 = one full 128-byte line written
// with four 32-byte stores (fp32 output: eight).  The TMEM buffer is handed back to the MMA warp
// right after the last tcgen05.ld, before the stores drain.
__device__ __forceinline__ void epilogue_tile(const EpiArgs& e, uint32_t tmem_acc, int quad, int lane,
                                              bool valid, long long off, int n_t,
                                              uint64_t* tmem_full_bar, uint32_t tph,
                                              uint64_t* tmem_empty_bar, uint32_t* stat_stage,
                                              float* stat_acc, int chunk_first = 0, int chunk_step = 64) {
  // chunk_first / chunk_step: with two epilogue warps per TMEM lane quadrant (MCN_EPI_WARPS=8) the
  // pair splits the 64-channel chunks of a tile between them (first = 0 or 64, step = 128).
      // bf16 accumulate (dx += dgrad): the previous values of a 64-channel chunk are fetched
  // with four back-to-back 32-byte loads one chunk AHEAD (the first one before the
  // accumulator is even ready), so the global-load latency hides behind the MMAs.
  const bool acc_bf16 = e.accumulate && !e.out_f32 && e.vec_ok;
  uint32_t old[32];
  __nv_bfloat16* obase = reinterpret_cast<__nv_bfloat16*>(e.out) + off;
  auto prefetch_old = [&](int c0) {
    const int col0 = n_t * e.block_n + c0;
    if (valid && col0 + 64 <= e.n_total) {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        ptx::ld_global_v8(obase + col0 + j * 16, *reinterpret_cast<uint32_t(*)[8]>(&old[j * 8]));
    }
  };
  if (acc_bf16 && chunk_first < e.block_n) prefetch_old(chunk_first);
#ifdef MCN_ROLE_TIMING
  const long long rt_w0 = clock64();
#endif
  ptx::mbar_wait(tmem_full_bar, tph);
#ifdef MCN_ROLE_TIMING
  if (quad == 0 && lane == 0) RT_ADD(3, clock64() - rt_w0);
#endif
  ptx::tc_fence_after();
  if (chunk_first >= e.block_n) {
    // nothing to drain for this warp (narrow tile): it still takes part in the hand-back count
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(tmem_empty_bar);
    return;
  }
  for (int c0 = chunk_first; c0 < e.block_n; c0 += chunk_step) {
    uint32_t r[64];
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) +
                           static_cast<uint32_t>(c0);
    if (!(e.debug & 4)) {
      ptx::tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
      ptx::tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
      ptx::tmem_ld_wait();
    }
    if (c0 + chunk_step >= e.block_n) {
      // last read of this accumulator: hand the TMEM buffer back before the stores drain
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty_bar);
    }
    const int col0 = n_t * e.block_n + c0;
    const bool fullw = e.vec_ok && (col0 + 64 <= e.n_total);
    if (col0 >= e.n_total) continue;                               // warp-uniform
    const bool do_store = valid && !(e.debug & 1);
    if (e.stats == nullptr && !do_store) continue;                 // statistics need every lane
    if (e.bias != nullptr) {
#pragma unroll
      for (int j = 0; j < 64; ++j)
        if (fullw || col0 + j < e.n_total)
          r[j] = __float_as_uint(__uint_as_float(r[j]) + e.bias[col0 + j]);
    }
    if (e.out_f32) {
      if (!do_store) continue;
      float* o = reinterpret_cast<float*>(e.out) + off + col0;
      if (fullw) {
#pragma unroll
        for (int j = 0; j < 64; j += 8) {
          uint32_t v[8];
          if (e.accumulate) {
            ptx::ld_global_v8(o + j, v);
#pragma unroll
            for (int i = 0; i < 8; ++i)
              v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(r[j + i]));
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = r[j + i];
          }
          ptx::st_global_v8(o + j, v);
        }
      } else {
        for (int j = 0; j < 64; ++j)
          if (col0 + j < e.n_total)
            o[j] = __uint_as_float(r[j]) + (e.accumulate ? o[j] : 0.f);
      }
    } else {
      __nv_bfloat16* o = obase + col0;
      if (fullw) {
        uint32_t v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float lo = __uint_as_float(r[2 * i]), hi = __uint_as_float(r[2 * i + 1]);
          if (acc_bf16) {
            lo += __uint_as_float(old[i] << 16);
            hi += __uint_as_float(old[i] & 0xFFFF0000u);
          }
          __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
          v[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        if (acc_bf16 && c0 + chunk_step < e.block_n) prefetch_old(c0 + chunk_step);
        if (do_store) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            ptx::st_global_v8(o + j * 16, *reinterpret_cast<uint32_t(*)[8]>(&v[j * 8]));
        }
        if (e.stats != nullptr) {
          uint32_t* mine = stat_stage + lane * kStatRowWords;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<uint4*>(mine + 4 * j) =
                valid ? make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3])
                      : make_uint4(0u, 0u, 0u, 0u);
          __syncwarp();
          float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
          for (int r = 0; r < 32; ++r) {
            const uint32_t w = stat_stage[r * kStatRowWords + lane];
            const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xFFFF0000u);
            s1a += lo;
            s1b += hi;
            s2a = fmaf(lo, lo, s2a);
            s2b = fmaf(hi, hi, s2b);
          }
          __syncwarp();
          // this warp's own slice, this lane's own two channels: plain read-modify-write
          float2* a1 = reinterpret_cast<float2*>(stat_acc + c0 + 2 * lane);
          float2* a2 = reinterpret_cast<float2*>(stat_acc + 256 + c0 + 2 * lane);
          float2 t1 = *a1, t2 = *a2;
          t1.x += s1a;
          t1.y += s1b;
          t2.x += s2a;
          t2.y += s2b;
          *a1 = t1;
          *a2 = t2;
        }
      } else if (do_store) {
        for (int j = 0; j < 64; ++j)
          if (col0 + j < e.n_total) {
            float f = __uint_as_float(r[j]);
            if (e.accumulate) f += __bfloat162float(o[j]);
            o[j] = __float2bfloat16_rn(f);
          }
      }
    }
  }
}

// ------------------------------------------------------------------ TMA-store epilogue
// The direct epilogue above writes one 128-byte row segment per THREAD: every 32-byte store
// instruction of a warp touches 32 different lines, and the L1 store path (one line per cycle or
// so) — not HBM, not the tensor pipe — paced the store-heavy 1x1 convolutions (per-role counters,
// profiles/r02_role_timing.txt: the MMA warp spent 71 % of a 64->256 fprop waiting for a free
// accumulator).  Here a warp packs its 32 rows x 64 channels into a 4 KB shared-memory tile in the
// tensor map's 128-byte swizzle (conflict-free 16-byte stores) and one lane hands it to the TMA
// unit: cp.async.bulk.tensor store, or cp.reduce.async.bulk.tensor .add for the accumulating dgrad
// (the sum happens in the memory system: no read-modify-write through registers; the addend is
// rounded to bf16 first, i.e. dx = bf16(dx + bf16(acc))).  Rows and columns outside the tensor are
// clipped by the hardware, so partial tiles need no predicates.  Fused BN statistics read the same
// staged tile column-wise (conflict-free under the swizzle).
constexpr int kStgBytes = 4096;   // 32 rows x 128 B

struct OutTile {   // coordinates of a warp's 32-row slab: 2-D map (channel, c1); 4-D (channel, c1, c2, c3)
  int c1, c2, c3;
};

// kMode: 0 store only, 1 + forward BN statistics (sum y, sum y^2), 2 + backward BN reduction: the
// tile of the BN input x that matches this output tile arrives by TMA (xmap -> xbuf, xbar) and the
// column pass accumulates sum dz and sum dz*x.
template <int kMode>
__device__ __forceinline__ void epilogue_tile_tma(const EpiArgs& e, const CUtensorMap* omap,
                                                  const OutTile& o, uint32_t tmem_acc, int quad,
                                                  int lane, bool valid, int n_t,
                                                  uint64_t* tmem_full_bar, uint32_t tph,
                                                  uint64_t* tmem_empty_bar, uint8_t* stg, int& stg_i,
                                                  float* stat_acc, const CUtensorMap* xmap = nullptr,
                                                  uint8_t* xbuf = nullptr, uint64_t* xbar = nullptr,
                                                  uint32_t* xph = nullptr) {
  constexpr bool kStats = kMode != 0;
#ifdef MCN_ROLE_TIMING
  const long long rt_w0 = clock64();
#endif
  ptx::mbar_wait(tmem_full_bar, tph);
#ifdef MCN_ROLE_TIMING
  if (quad == 0 && lane == 0) RT_ADD(3, clock64() - rt_w0);
#endif
  ptx::tc_fence_after();
  const uint32_t swz = static_cast<uint32_t>(lane & 7) << 4;
  for (int c0 = 0; c0 < e.block_n; c0 += 64) {
    uint32_t r[64];
    const uint32_t taddr = tmem_acc + (static_cast<uint32_t>(quad * 32) << 16) +
                           static_cast<uint32_t>(c0);
    ptx::tmem_ld_32x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
    ptx::tmem_ld_32x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
    ptx::tmem_ld_wait();
    if (c0 + 64 >= e.block_n) {
      // last read of this accumulator: hand the TMEM buffer back before the stores drain
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(tmem_empty_bar);
    }
    const int col0 = n_t * e.block_n + c0;
    if (col0 >= e.n_total) continue;   // warp-uniform (n_total % 64 == 0: a chunk is whole or absent)
    if (kMode == 2 && lane == 0) {
      // the x tile of this chunk (the previous chunk's column pass ended with a __syncwarp)
      ptx::mbar_expect_tx(xbar, kStgBytes);
      if (e.out_rank4) ptx::tma_load_4d(xmap, xbar, xbuf, col0, o.c1, o.c2, o.c3);
      else ptx::tma_load_2d(xmap, xbar, xbuf, col0, o.c1);
    }
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
      v[i] = *reinterpret_cast<uint32_t*>(&h);
      if (kStats && !valid) v[i] = 0u;   // rows outside the tensor must not reach the sums
    }
    uint8_t* buf = stg + stg_i * kStgBytes;
    // the bulk store that last used this buffer has finished reading it
    if (lane == 0) {
      if (e.stg_bufs == 2) ptx::bulk_wait_read<1>();
      else ptx::bulk_wait_read<0>();
    }
    __syncwarp();
    const uint32_t rowaddr = ptx::smem_u32(buf) + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      ptx::st_shared_v4(rowaddr + ((static_cast<uint32_t>(j) << 4) ^ swz), v[4 * j], v[4 * j + 1],
                        v[4 * j + 2], v[4 * j + 3]);
    ptx::fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      if (e.out_rank4) {
        if (e.accumulate) ptx::tma_reduce_add_4d(omap, buf, col0, o.c1, o.c2, o.c3);
        else ptx::tma_store_4d(omap, buf, col0, o.c1, o.c2, o.c3);
      } else {
        if (e.accumulate) ptx::tma_reduce_add_2d(omap, buf, col0, o.c1);
        else ptx::tma_store_2d(omap, buf, col0, o.c1);
      }
      ptx::bulk_commit();
    }
    if (kStats) {
      // lane L sums channels 2L, 2L+1 down the warp's 32 rows; word L of row r sits in 16-byte
      // chunk (L/4) ^ (r & 7) of the swizzled row: 32 lanes, 32 banks
      const uint32_t* bw = reinterpret_cast<const uint32_t*>(buf);
      const int wq = lane >> 2, wr = lane & 3;
      float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
      if (kMode == 2) {
        // dz = dy * act'(x*sc + sf) with the forward pass's own constants (same fmaf, same sign)
        const int ch = col0 + 2 * lane;
        float sca = 1.f, scb = 1.f, sfa = 0.f, sfb = 0.f;
        const bool relu = e.red == 2;
        if (relu) {
          const float ia = e.red_invstd[ch], ib = e.red_invstd[ch + 1];
          sca = (e.red_gamma ? e.red_gamma[ch] : 1.f) * ia;
          scb = (e.red_gamma ? e.red_gamma[ch + 1] : 1.f) * ib;
          sfa = (e.red_beta ? e.red_beta[ch] : 0.f) - e.red_mean[ch] * sca;
          sfb = (e.red_beta ? e.red_beta[ch + 1] : 0.f) - e.red_mean[ch + 1] * scb;
        }
        ptx::mbar_wait(xbar, *xph);
        *xph ^= 1u;
        const uint32_t* xw = reinterpret_cast<const uint32_t*>(xbuf);
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          const int idx = rr * 32 + (((wq ^ (rr & 7)) << 2) | wr);
          const uint32_t w = bw[idx], xx = xw[idx];
          const float xlo = __uint_as_float(xx << 16), xhi = __uint_as_float(xx & 0xFFFF0000u);
          float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xFFFF0000u);
          if (relu) {
            lo = fmaf(xlo, sca, sfa) > 0.f ? lo : 0.f;
            hi = fmaf(xhi, scb, sfb) > 0.f ? hi : 0.f;
          }
          s1a += lo;
          s1b += hi;
          s2a = fmaf(lo, xlo, s2a);
          s2b = fmaf(hi, xhi, s2b);
        }
        __syncwarp();   // every lane is done with xbuf before lane 0 refills it
      } else {
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
          const uint32_t w = bw[rr * 32 + (((wq ^ (rr & 7)) << 2) | wr)];
          const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xFFFF0000u);
          s1a += lo;
          s1b += hi;
          s2a = fmaf(lo, lo, s2a);
          s2b = fmaf(hi, hi, s2b);
        }
      }
      float2* a1 = reinterpret_cast<float2*>(stat_acc + c0 + 2 * lane);
      float2* a2 = reinterpret_cast<float2*>(stat_acc + 256 + c0 + 2 * lane);
      float2 t1 = *a1, t2 = *a2;
      t1.x += s1a;
      t1.y += s1b;
      t2.x += s2a;
      t2.y += s2b;
      *a1 = t1;
      *a2 = t2;
    }
    stg_i = (stg_i + 1 == e.stg_bufs) ? 0 : stg_i + 1;
  }
}

// ------------------------------------------------------------------ fprop / dgrad kernel
struct GemmConvArgs {
  CUtensorMap mapA[4];
  CUtensorMap mapB;
  CUtensorMap mapOut;   // TMA-store epilogue: [M, Cout] bf16, box 64 x 32
  CUtensorMap mapRed;   // fused BN-backward reduction: the BN input x, same geometry as mapOut
  TileGeom g;
  int taps, k_chunks, ksteps_last, block_n, stages, tiles_n, tmem_cols, total_tiles;
  int b_stationary;   // the CTA's whole weight slab (all taps / k-chunks of its n-tile) stays in smem
  long long out_sn, out_sh, out_sw;  // element strides of the output pixel grid
  EpiArgs e;
  TapTab tab;
};

// Persistent: one CTA per SM walks tiles (tile = blockIdx.x + i*gridDim.x).  The shared-memory
// ring and its phases run across tile boundaries, so the producer prefetches the next tile while
// the MMA thread finishes the current one; the accumulator is double-buffered in TMEM so the
// epilogue of tile i overlaps the MMAs of tile i+1.
// The MMA role is ONE thread running a tight loop: descriptors are a constant upper half plus the
// shifted address, ring indices advance by increment, the four K=16 steps of a stage are unrolled.
// (Per-role counters showed the old per-iteration elect / descriptor rebuild / modulo chain cost
// ~160 cycles per tcgen05.mma — five times the 32 cycles an N=64 instruction occupies the pipe.)
// kTma: TMA-store epilogue (bf16 [M, Cout] output with Cout % 64 == 0) or the direct one.
template <bool kTma>
__global__ void __launch_bounds__(192, 1)
gemm_conv_kernel(const __grid_constant__ GemmConvArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = args.stages;
  const uint32_t b_bytes = static_cast<uint32_t>(args.block_n) * 128u;
  // Weight-stationary mode: small weight slabs (1x1 convs) are loaded ONCE per CTA behind the
  // activation ring instead of once per tile — for K = 64, N = 256 the weights were two thirds of
  // the shared-memory fill traffic.  Needs a constant n-tile per CTA (gridDim % tiles_n == 0).
  const bool bstat = args.b_stationary != 0;
  const uint32_t stage_bytes = bstat ? kABytes : kABytes + b_bytes;
  const int k_iters = args.taps * args.k_chunks;
  uint8_t* smemBs = smem + static_cast<size_t>(stages) * stage_bytes;
  uint8_t* stg_all = smemBs + (bstat ? static_cast<size_t>(k_iters) * b_bytes : 0);   // 1024-aligned
  uint8_t* xbuf_all = stg_all + (kTma ? static_cast<size_t>(4 * args.e.stg_bufs) * kStgBytes : 0);
  uint8_t* tail = xbuf_all + ((kTma && args.e.red) ? static_cast<size_t>(4) * kStgBytes : 0);
  uint64_t* xbar_all = reinterpret_cast<uint64_t*>(tail + 464);   // [4], one per epilogue warp
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + stages;
  uint64_t* tmem_full = empty + stages;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;     // [2]
  uint64_t* bstat_bar = tmem_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bstat_bar + 1);

  const int total_tiles = args.total_tiles;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&args.mapA[i]);
    ptx::prefetch_tmap(&args.mapB);
    if (kTma) ptx::prefetch_tmap(&args.mapOut);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tmem_full[b], 1);
      ptx::mbar_init(&tmem_empty[b], 4);   // one arrival per epilogue warp
    }
    ptx::mbar_init(bstat_bar, 1);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&xbar_all[i], 1);
    if (kTma && args.e.red) ptx::prefetch_tmap(&args.mapRed);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, static_cast<uint32_t>(args.tmem_cols));
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, tensor-map prefetch and the TMEM allocation above overlap the previous kernel's tail
  MCN_PDL_PROLOGUE();
  const long long rt_cta0 = RT_NOW();
  (void)rt_cta0;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (ptx::elect_one()) {
      RT_DECL;
      const uint32_t a_tx = (args.g.a_mode == 0) ? static_cast<uint32_t>(args.g.rows_box) * 128u
                                                 : static_cast<uint32_t>(kABytes);
      if (bstat) {
        const int n_t = blockIdx.x % args.tiles_n;
        ptx::mbar_expect_tx(bstat_bar, static_cast<uint32_t>(k_iters) * b_bytes);
        for (int t = 0; t < args.taps; ++t)
          for (int kc = 0; kc < args.k_chunks; ++kc)
            ptx::tma_load_2d(&args.mapB, bstat_bar,
                             smemBs + static_cast<size_t>(t * args.k_chunks + kc) * b_bytes, kc * 64,
                             args.tab.brow[t] + n_t * args.block_n);
      }
      int s = 0;
      uint32_t ph = 0;
      for (int tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x) {
        const int n_t = tile_id % args.tiles_n;
        const PixelTile tile = decode_tile(args.g, tile_id / args.tiles_n);
        for (int t = 0; t < args.taps; ++t) {
          for (int kc = 0; kc < args.k_chunks; ++kc) {
            RT_BEGIN;
            ptx::mbar_wait_quiet(&empty[s], ph ^ 1u);
            RT_END;
            ptx::mbar_expect_tx(&full[s], bstat ? a_tx : a_tx + b_bytes);
            uint8_t* sA = smem + static_cast<size_t>(s) * stage_bytes;
            uint8_t* sB = sA + kABytes;
            if (args.g.a_mode == 0) {
              ptx::tma_load_4d(&args.mapA[args.tab.map[t]], &full[s], sA, kc * 64,
                               tile.w0 + args.tab.dw[t], tile.h0 + args.tab.dh[t], tile.n0);
            } else {
              ptx::tma_load_im2col_4d(&args.mapA[0], &full[s], sA, kc * 64, tile.w0, tile.h0,
                                      tile.n0, static_cast<uint16_t>(args.tab.dw[t]),
                                      static_cast<uint16_t>(args.tab.dh[t]));
            }
            if (!bstat)
              ptx::tma_load_2d(&args.mapB, &full[s], sB, kc * 64,
                               args.tab.brow[t] + n_t * args.block_n);
            if (++s == stages) {
              s = 0;
              ph ^= 1u;
            }
          }
        }
      }
      RT_FLUSH(0);
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (one thread) ----------------
    if (ptx::elect_one()) {
      RT_DECL;
      const long long rt_m0 = RT_NOW();
      (void)rt_m0;
      const uint32_t idesc = ptx::make_idesc_bf16(128, args.block_n, 0, 0);
      const uint64_t dhi = ptx::smem_desc_hi(16, 1024);
      const uint32_t smem_a0 = ptx::smem_u32(smem);
      const uint32_t smem_b0 = ptx::smem_u32(smemBs);
      const int kch = args.k_chunks, last_steps = args.ksteps_last;
      if (bstat) ptx::mbar_wait_quiet(bstat_bar, 0);
      int s = 0, lt = 0;
      uint32_t ph = 0;
      for (int tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x, ++lt) {
        const int buf = lt & 1;
        const uint32_t tph = static_cast<uint32_t>(lt >> 1) & 1u;
        RT_BEGIN;
        ptx::mbar_wait_quiet(&tmem_empty[buf], tph ^ 1u);   // epilogue has drained this accumulator
        RT_END2;
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * args.block_n);
        uint32_t acc = 0u, b_stat_addr = smem_b0;
        int kc = 0;
        for (int kit = 0; kit < k_iters; ++kit) {
          RT_BEGIN;
          ptx::mbar_wait_quiet(&full[s], ph);
          RT_END;
          ptx::tc_fence_after();
          const uint32_t a_addr = smem_a0 + static_cast<uint32_t>(s) * stage_bytes;
          const uint64_t ad = ptx::smem_desc_at(dhi, a_addr);
          const uint64_t bd = ptx::smem_desc_at(dhi, bstat ? b_stat_addr : a_addr + kABytes);
          if (kc != kch - 1 || last_steps == 4) {
            ptx::umma_bf16(tmem_d, ad, bd, idesc, acc);
            ptx::umma_bf16(tmem_d, ad + 2, bd + 2, idesc, 1u);
            ptx::umma_bf16(tmem_d, ad + 4, bd + 4, idesc, 1u);
            ptx::umma_bf16(tmem_d, ad + 6, bd + 6, idesc, 1u);
          } else {
            for (int k = 0; k < last_steps; ++k)
              ptx::umma_bf16(tmem_d, ad + 2 * k, bd + 2 * k, idesc, (k == 0) ? acc : 1u);
          }
          acc = 1u;
          ptx::umma_commit(&empty[s]);
          b_stat_addr += b_bytes;
          if (++kc == kch) kc = 0;
          if (++s == stages) {
            s = 0;
            ph ^= 1u;
          }
        }
        ptx::umma_commit(&tmem_full[buf]);
      }
      RT_FLUSH(1);
      RT_FLUSH2(2);
      RT_ADD(7, RT_NOW() - rt_m0);
    }
  } else {
    // ---------------- epilogue ----------------
    const long long rt_e0 = RT_NOW();
    (void)rt_e0;
    const int quad = warp & 3;  // TMEM lane quadrant this warp may read
    const int row = quad * 32 + lane;
    const bool stats = args.e.stats != nullptr;
    uint32_t* stat_stage = reinterpret_cast<uint32_t*>(tail + kBarRegionBytes);
    // direct epilogue: [staging 18 KB][accumulators]; TMA epilogue: accumulators only
    float* stat_acc_all = kTma ? reinterpret_cast<float*>(stat_stage)
                               : reinterpret_cast<float*>(stat_stage + kStatStageWords);
    float* stat_acc = stat_acc_all + quad * kStatAccWarp;
    stat_stage += quad * 32 * kStatRowWords;
    if (stats) {
      for (int i = row; i < kStatAccFloats; i += 128) stat_acc_all[i] = 0.f;
      epi_bar_sync();
    }
    int* stat_flag = reinterpret_cast<int*>(tail + kBarRegionBytes - 16);
    uint8_t* stg = stg_all + static_cast<size_t>(quad * args.e.stg_bufs) * kStgBytes;
    int stg_i = 0;
    uint32_t xph = 0;
    int lt = 0, cur_nt = -1, pending = 0, since = 0;
    for (int tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const uint32_t tph = static_cast<uint32_t>(lt >> 1) & 1u;
      const int n_t = tile_id % args.tiles_n;
      if (stats) {
        if (n_t != cur_nt) {
          if (cur_nt >= 0) stats_flush(args.e, stat_acc_all, cur_nt, row, pending, stat_flag);
          cur_nt = n_t;
          pending = since = 0;
        } else if (since == 8) {
          // every 8 tiles: bounds the length of the fp32 partial sums (accuracy of sum x^2)
          stats_flush(args.e, stat_acc_all, cur_nt, row, pending, stat_flag, false);
          since = 0;
        }
      }
      ++pending;
      ++since;
      if (kTma) {
        // the tile is 128 consecutive rows of the [M, Cout] output matrix
        const long long m0 = static_cast<long long>(tile_id / args.tiles_n) * 128;
        const bool valid = m0 + row < args.g.m_total;
        const OutTile o{static_cast<int>(m0) + quad * 32, 0, 0};
        const uint32_t tacc = tmem_base + static_cast<uint32_t>(buf * args.e.block_n);
        if (args.e.red)
          epilogue_tile_tma<2>(args.e, &args.mapOut, o, tacc, quad, lane, valid, n_t, &tmem_full[buf], tph,
                               &tmem_empty[buf], stg, stg_i, stat_acc, &args.mapRed,
                               xbuf_all + static_cast<size_t>(quad) * kStgBytes, &xbar_all[quad], &xph);
        else if (stats)
          epilogue_tile_tma<1>(args.e, &args.mapOut, o, tacc, quad, lane, valid, n_t, &tmem_full[buf], tph,
                               &tmem_empty[buf], stg, stg_i, stat_acc);
        else
          epilogue_tile_tma<0>(args.e, &args.mapOut, o, tacc, quad, lane, valid, n_t, &tmem_full[buf], tph,
                               &tmem_empty[buf], stg, stg_i, stat_acc);
      } else {
        const PixelTile tile = decode_tile(args.g, tile_id / args.tiles_n);
        int n, p, q;
        const bool valid = row_coords(args.g, tile, row, n, p, q);
        const long long off = valid ? (n * args.out_sn + p * args.out_sh + q * args.out_sw) : 0;
        epilogue_tile(args.e, tmem_base + static_cast<uint32_t>(buf * args.e.block_n), quad, lane, valid,
                      off, n_t, &tmem_full[buf], tph, &tmem_empty[buf], stat_stage, stat_acc);
      }
    }
    if (stats && cur_nt >= 0) stats_flush(args.e, stat_acc_all, cur_nt, row, pending, stat_flag);
    if (kTma && lane == 0) ptx::bulk_wait_all();   // the staged tiles must outlive their stores
    if (row == 0) RT_ADD(4, RT_NOW() - rt_e0);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    RT_ADD(5, RT_NOW() - rt_cta0);
    RT_ADD(6, 1);
  }
  if (warp == 1) ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(args.tmem_cols));
}

// ------------------------------------------------------------------ halo kernel (k x k, stride 1)
// For stride-1 k x k convolutions the im2col feed re-reads every input pixel kh*kw times from L2.
// Here an output tile is 16 rows x 8 columns of one image and its INPUT halo
// ((16+(kh-1)dh) x (8+(kw-1)dw) pixels x 64 channels) is brought in by ONE TMA box per channel
// chunk; each filter tap is then just a different start address into the same shared-memory
// tile: 8 consecutive pixels of an output row are 8 consecutive 128-byte rows (one swizzle
// group) and successive output rows are SBO = halo_width*128 bytes apart.  Activation traffic
// drops from kh*kw x to ~1.4 x; weight tiles stream through their own ring (or stay resident for
// the whole persistent CTA when all taps fit: the 64->64 3x3 layers).
struct HaloArgs {
  CUtensorMap mapA;
  CUtensorMap mapB;
  CUtensorMap mapOut;    // TMA-store epilogue: (C, W, H, N) bf16, box 64 x 8 x 4 x 1
  CUtensorMap mapRed;    // fused BN-backward reduction: the BN input x, same geometry as mapOut
  EpiArgs e;
  int taps, k_chunks, tiles_n, tmem_cols, total_tiles;
  int tiles_w, tiles_h;
  int Wo, Ho, Nb;
  int hwb, hhb;          // halo box extent in pixels
  int org_w, org_h;      // input coordinate of the halo origin relative to the tile origin
  int a_stages, b_stages, halo_stride, b_stationary, use_base_offset;
  long long out_sn, out_sh, out_sw;
  int brow[kMaxTaps];
  short toff[kMaxTaps];  // halo row (pixel) offset of each tap
};

template <bool kTma>
__global__ void __launch_bounds__(224, 1)
halo_conv_kernel(const __grid_constant__ HaloArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t b_bytes = static_cast<uint32_t>(args.e.block_n) * 128u;
  const int nb_slots = args.b_stationary ? args.taps * args.k_chunks : args.b_stages;
  uint8_t* smemA = smem;
  uint8_t* smemB = smem + static_cast<size_t>(args.a_stages) * args.halo_stride;
  uint8_t* stg_all = smemB + static_cast<size_t>(nb_slots) * b_bytes;   // 1024-aligned
  uint8_t* xbuf_all = stg_all + (kTma ? static_cast<size_t>(4 * args.e.stg_bufs) * kStgBytes : 0);
  uint8_t* tail = xbuf_all + ((kTma && args.e.red) ? static_cast<size_t>(4) * kStgBytes : 0);
  uint64_t* xbar_all = reinterpret_cast<uint64_t*>(tail + 464);   // [4], one per epilogue warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* full_a = bars;
  uint64_t* empty_a = full_a + args.a_stages;
  uint64_t* full_b = empty_a + args.a_stages;      // [b_stages] (stationary: [0] only)
  uint64_t* empty_b = full_b + args.b_stages;
  uint64_t* tmem_full = empty_b + args.b_stages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint32_t* tap_desc = reinterpret_cast<uint32_t*>(tail + 256);   // [taps] descriptor steps (16-byte units)
  const int total_tiles = args.total_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&args.mapA);
    ptx::prefetch_tmap(&args.mapB);
    if (kTma) ptx::prefetch_tmap(&args.mapOut);
    for (int s = 0; s < args.a_stages; ++s) {
      ptx::mbar_init(&full_a[s], 1);
      ptx::mbar_init(&empty_a[s], 1);
    }
    for (int s = 0; s < args.b_stages; ++s) {
      ptx::mbar_init(&full_b[s], 1);
      ptx::mbar_init(&empty_b[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tmem_full[b], 1);
      ptx::mbar_init(&tmem_empty[b], 4);
    }
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&xbar_all[i], 1);
    if (kTma && args.e.red) ptx::prefetch_tmap(&args.mapRed);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    for (int t = lane; t < args.taps; t += 32) tap_desc[t] = static_cast<uint32_t>(args.toff[t]) * 8u;
    ptx::tmem_alloc(tmem_slot, static_cast<uint32_t>(args.tmem_cols));
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, tensor-map prefetch and the TMEM allocation above overlap the previous kernel's tail
  MCN_PDL_PROLOGUE();
  const int tiles_per_img = args.tiles_w * args.tiles_h;
  const long long rt_cta0 = RT_NOW();
  (void)rt_cta0;

  if (warp == 0) {
    // ---------------- activation (halo) producer ----------------
    if (ptx::elect_one()) {
      RT_DECL;
      const uint32_t a_tx = static_cast<uint32_t>(args.hwb * args.hhb) * 128u;
      int s = 0;
      uint32_t ph = 0;
      for (int tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x) {
        const int m = tile_id / args.tiles_n;
        const int n = m / tiles_per_img;
        const int r = m - n * tiles_per_img;
        const int h0 = (r / args.tiles_w) * 16, w0 = (r % args.tiles_w) * 8;
        for (int kc = 0; kc < args.k_chunks; ++kc) {
          RT_BEGIN;
          ptx::mbar_wait_quiet(&empty_a[s], ph ^ 1u);
          RT_END;
          ptx::mbar_expect_tx(&full_a[s], a_tx);
          ptx::tma_load_4d(&args.mapA, &full_a[s], smemA + static_cast<size_t>(s) * args.halo_stride,
                           kc * 64, w0 + args.org_w, h0 + args.org_h, n);
          if (++s == args.a_stages) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
      RT_FLUSH(0);
    }
  } else if (warp == 2) {
    // ---------------- weight producer ----------------
    if (ptx::elect_one()) {
      if (args.b_stationary) {
        // every (k-chunk, tap) weight tile fits: load once, keep for all tiles of this CTA
        const int n_t = blockIdx.x % args.tiles_n;   // tiles_n == 1 in this mode
        ptx::mbar_expect_tx(&full_b[0], static_cast<uint32_t>(nb_slots) * b_bytes);
        for (int kc = 0; kc < args.k_chunks; ++kc)
          for (int t = 0; t < args.taps; ++t)
            ptx::tma_load_2d(&args.mapB, &full_b[0],
                             smemB + static_cast<size_t>(kc * args.taps + t) * b_bytes, kc * 64,
                             args.brow[t] + n_t * args.e.block_n);
      } else {
        int s = 0;
        uint32_t ph = 0;
        for (int tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x) {
          const int n_t = tile_id % args.tiles_n;
          for (int kc = 0; kc < args.k_chunks; ++kc)
            for (int t = 0; t < args.taps; ++t) {
              ptx::mbar_wait_quiet(&empty_b[s], ph ^ 1u);
              ptx::mbar_expect_tx(&full_b[s], b_bytes);
              ptx::tma_load_2d(&args.mapB, &full_b[s], smemB + static_cast<size_t>(s) * b_bytes,
                               kc * 64, args.brow[t] + n_t * args.e.block_n);
              if (++s == args.b_stages) {
                s = 0;
                ph ^= 1u;
              }
            }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer (one thread, see gemm_conv_kernel) ----------------
    if (ptx::elect_one()) {
      RT_DECL;
      const long long rt_m0 = RT_NOW();
      (void)rt_m0;
      const uint32_t idesc = ptx::make_idesc_bf16(128, args.e.block_n, 0, 0);
      const uint64_t dhi_a = ptx::smem_desc_hi(16, static_cast<uint32_t>(args.hwb) * 128u);
      const uint64_t dhi_b = ptx::smem_desc_hi(16, 1024);
      const uint32_t smem_a0 = ptx::smem_u32(smemA), smem_b0 = ptx::smem_u32(smemB);
      const bool bstat = args.b_stationary != 0, use_bo = args.use_base_offset != 0;
      const int taps = args.taps;
      int sa = 0, sb = 0, lt = 0;
      uint32_t pha = 0, phb = 0;
      if (bstat) ptx::mbar_wait_quiet(&full_b[0], 0);
      for (int tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x, ++lt) {
        const int buf = lt & 1;
        const uint32_t tph = static_cast<uint32_t>(lt >> 1) & 1u;
        RT_BEGIN;
        ptx::mbar_wait_quiet(&tmem_empty[buf], tph ^ 1u);
        RT_END2;
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * args.e.block_n);
        uint32_t acc = 0u, b_stat_addr = smem_b0;
        for (int kc = 0; kc < args.k_chunks; ++kc) {
          RT_BEGIN;
          ptx::mbar_wait_quiet(&full_a[sa], pha);
          RT_END;
          ptx::tc_fence_after();
          const uint32_t a_base = smem_a0 + static_cast<uint32_t>(sa) * static_cast<uint32_t>(args.halo_stride);
          const uint64_t ad0 = ptx::smem_desc_at(dhi_a, a_base);
          for (int t = 0; t < taps; ++t) {
            uint32_t b_addr;
            if (bstat) {
              b_addr = b_stat_addr;
              b_stat_addr += b_bytes;
            } else {
              RT_BEGIN;
              ptx::mbar_wait_quiet(&full_b[sb], phb);
              RT_END;
              ptx::tc_fence_after();
              b_addr = smem_b0 + static_cast<uint32_t>(sb) * b_bytes;
            }
            const uint32_t step = tap_desc[t];
            uint64_t ad = ad0 + step;
            if (use_bo) ad |= static_cast<uint64_t>(((a_base >> 7) + (step >> 3)) & 7u) << 49;
            const uint64_t bd = ptx::smem_desc_at(dhi_b, b_addr);
            ptx::umma_bf16(tmem_d, ad, bd, idesc, acc);
            ptx::umma_bf16(tmem_d, ad + 2, bd + 2, idesc, 1u);
            ptx::umma_bf16(tmem_d, ad + 4, bd + 4, idesc, 1u);
            ptx::umma_bf16(tmem_d, ad + 6, bd + 6, idesc, 1u);
            acc = 1u;
            if (!bstat) {
              ptx::umma_commit(&empty_b[sb]);
              if (++sb == args.b_stages) {
                sb = 0;
                phb ^= 1u;
              }
            }
          }
          ptx::umma_commit(&empty_a[sa]);
          if (++sa == args.a_stages) {
            sa = 0;
            pha ^= 1u;
          }
        }
        ptx::umma_commit(&tmem_full[buf]);
      }
      RT_FLUSH(1);
      RT_FLUSH2(2);
      RT_ADD(7, RT_NOW() - rt_m0);
    }
  } else {
    // ---------------- epilogue (warps 3..6) ----------------
    const long long rt_e0 = RT_NOW();
    (void)rt_e0;
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const bool stats = args.e.stats != nullptr;
    uint32_t* stat_stage = reinterpret_cast<uint32_t*>(tail + kBarRegionBytes);
    float* stat_acc_all = kTma ? reinterpret_cast<float*>(stat_stage)
                               : reinterpret_cast<float*>(stat_stage + kStatStageWords);
    float* stat_acc = stat_acc_all + quad * kStatAccWarp;
    stat_stage += quad * 32 * kStatRowWords;
    if (stats) {
      for (int i = row; i < kStatAccFloats; i += 128) stat_acc_all[i] = 0.f;
      epi_bar_sync();
    }
    int* stat_flag = reinterpret_cast<int*>(tail + kBarRegionBytes - 16);
    uint8_t* stg = stg_all + static_cast<size_t>(quad * args.e.stg_bufs) * kStgBytes;
    int stg_i = 0;
    uint32_t xph = 0;
    int lt = 0, cur_nt = -1, pending = 0, since = 0;
    for (int tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x, ++lt) {
      const int buf = lt & 1;
      const uint32_t tph = static_cast<uint32_t>(lt >> 1) & 1u;
      const int n_t = tile_id % args.tiles_n;
      if (stats) {
        if (n_t != cur_nt) {
          if (cur_nt >= 0) stats_flush(args.e, stat_acc_all, cur_nt, row, pending, stat_flag);
          cur_nt = n_t;
          pending = since = 0;
        } else if (since == 8) {
          // every 8 tiles: bounds the length of the fp32 partial sums (accuracy of sum x^2)
          stats_flush(args.e, stat_acc_all, cur_nt, row, pending, stat_flag, false);
          since = 0;
        }
      }
      ++pending;
      ++since;
      const int m = tile_id / args.tiles_n;
      const int n = m / tiles_per_img;
      const int r = m - n * tiles_per_img;
      const int h0 = (r / args.tiles_w) * 16, w0 = (r % args.tiles_w) * 8;
      const int p = h0 + (row >> 3);
      const int q = w0 + (row & 7);
      const bool valid = p < args.Ho && q < args.Wo;
      if (kTma) {
        const OutTile o{w0, h0 + quad * 4, n};   // this warp's 4 output rows x 8 columns
        const uint32_t tacc = tmem_base + static_cast<uint32_t>(buf * args.e.block_n);
        if (args.e.red)
          epilogue_tile_tma<2>(args.e, &args.mapOut, o, tacc, quad, lane, valid, n_t, &tmem_full[buf], tph,
                               &tmem_empty[buf], stg, stg_i, stat_acc, &args.mapRed,
                               xbuf_all + static_cast<size_t>(quad) * kStgBytes, &xbar_all[quad], &xph);
        else if (stats)
          epilogue_tile_tma<1>(args.e, &args.mapOut, o, tacc, quad, lane, valid, n_t, &tmem_full[buf], tph,
                               &tmem_empty[buf], stg, stg_i, stat_acc);
        else
          epilogue_tile_tma<0>(args.e, &args.mapOut, o, tacc, quad, lane, valid, n_t, &tmem_full[buf], tph,
                               &tmem_empty[buf], stg, stg_i, stat_acc);
      } else {
        const long long off = valid ? (n * args.out_sn + p * args.out_sh + q * args.out_sw) : 0;
        epilogue_tile(args.e, tmem_base + static_cast<uint32_t>(buf * args.e.block_n), quad, lane, valid,
                      off, n_t, &tmem_full[buf], tph, &tmem_empty[buf], stat_stage, stat_acc);
      }
    }
    if (stats && cur_nt >= 0) stats_flush(args.e, stat_acc_all, cur_nt, row, pending, stat_flag);
    if (kTma && lane == 0) ptx::bulk_wait_all();
    if (row == 0) RT_ADD(4, RT_NOW() - rt_e0);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    RT_ADD(5, RT_NOW() - rt_cta0);
    RT_ADD(6, 1);
  }
  if (warp == 1) ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(args.tmem_cols));
}

// ------------------------------------------------------------------ wgrad kernel
struct WgradArgs {
  CUtensorMap mapX[4];
  CUtensorMap mapDy;
  TileGeom g;
  int taps, stages, block_n, nb_atoms, tmem_cols;
  int a_atoms;     // 64-channel atoms of x per CTA: 2 (M = 128) or 4 (two M = 128 accumulators: "wide" tile)
  int atom_bytes;  // shared-memory bytes of one atom: (pixels per stage) x 128
  int cin, cout;
  int tiles_mi, tiles_ni, splits, kblocks_total, ksteps;
  float* dw;          // split-K partial slices, or the gradient itself when K is not split
  long long ws_stride;  // elements between the slices of consecutive splits;
                        // dw; -1 = K is not split: every element has one owner, added in place
  // In-kernel ordered reduction (fuse_reduce): the CTAs of one output tile meet at a ticket counter
  // once their partial tiles are written (all CTAs of the grid are co-resident: one per SM), then
  // split s sums share s of the tile over the slices 0 .. splits-1 IN THAT ORDER and adds it into
  // dw_final: the separate splitk_reduce launch (and its kernel boundary) disappears, the order stays fixed.
  int fuse_reduce;
  float* dw_final;
  unsigned int* tickets;   // one per output tile, zero at launch, zero again at exit
  // Partial tiles through TMA stores (tma_out): each epilogue warp stages 32 rows x 32 fp32 columns
  // (128-byte rows, the map's swizzle) in the idle pipeline buffers and one lane issues a bulk store
  // into the 4-D view (Cout, Cin, tap, split) of the slices; rows >= Cin are clipped by the map.
  // (The direct write-out is one 128-byte row per thread: 256 KB per CTA took ~20k cycles.)
  int tma_out;
  CUtensorMap mapOut;
  TapTab tab;
};

// A thread's 32 consecutive output channels of one dw row: plain 16-byte stores into this split's
// private slice (summed later in split order by splitk_reduce), or — K not split, every element has
// one owner in the grid — a read-modify-write of the gradient itself.  No floating-point atomics.
__device__ __forceinline__ void wgrad_out32(float* o, const uint32_t (&r)[32], bool owner_rmw = false) {
  if (owner_rmw) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      float4 v = *reinterpret_cast<float4*>(o + j);
      v.x += __uint_as_float(r[j]);
      v.y += __uint_as_float(r[j + 1]);
      v.z += __uint_as_float(r[j + 2]);
      v.w += __uint_as_float(r[j + 3]);
      *reinterpret_cast<float4*>(o + j) = v;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<uint4*>(o + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
  }
}

__global__ void __launch_bounds__(192, 1)
wgrad_kernel(const __grid_constant__ WgradArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = args.stages;
  const uint32_t atom = static_cast<uint32_t>(args.atom_bytes);
  const uint32_t a_bytes = static_cast<uint32_t>(args.a_atoms) * atom;
  const uint32_t b_bytes = static_cast<uint32_t>(args.nb_atoms) * atom;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  const int m_halves = args.a_atoms >> 1;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(stages) * stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* tmem_full = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  int w = blockIdx.x;
  const int ni = w % args.tiles_ni;
  w /= args.tiles_ni;
  const int mi = w % args.tiles_mi;
  w /= args.tiles_mi;
  const int tap = w % args.taps;
  const int split = w / args.taps;
  const int kb0 = static_cast<int>(static_cast<long long>(split) * args.kblocks_total / args.splits);
  const int kb1 =
      static_cast<int>(static_cast<long long>(split + 1) * args.kblocks_total / args.splits);
  const long long rt_cta0 = RT_NOW();
  (void)rt_cta0;

  // Rows a tiled box does not cover must read as zero (K padding): clear the stage buffers once.
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = static_cast<int>(static_cast<size_t>(stages) * stage_bytes / 16);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
    ptx::fence_proxy_async();
  }
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 4; ++i) ptx::prefetch_tmap(&args.mapX[i]);
    ptx::prefetch_tmap(&args.mapDy);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, static_cast<uint32_t>(args.tmem_cols));
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, tensor-map prefetch and the TMEM allocation above overlap the previous kernel's tail
  MCN_PDL_PROLOGUE();

  if (warp == 0) {
    if (ptx::elect_one()) {
      const uint32_t rows_bytes =
          (args.g.a_mode == 0) ? static_cast<uint32_t>(args.g.rows_box) * 128u : kABytes;
      const uint32_t tx = rows_bytes * static_cast<uint32_t>(args.a_atoms + args.nb_atoms);
      RT_DECL;
      RT_ADD(2, RT_NOW() - rt_cta0);      // wgrad: slot 2 = prologue (smem clear, barrier init, TMEM alloc)
      for (int kb = kb0, it = 0; kb < kb1; ++kb, ++it) {
        const int s = it % stages;
        const uint32_t ph = static_cast<uint32_t>(it / stages) & 1u;
        RT_BEGIN;
        ptx::mbar_wait(&empty[s], ph ^ 1u);
        RT_END;
        ptx::mbar_expect_tx(&full[s], tx);
        uint8_t* sA = smem + static_cast<size_t>(s) * stage_bytes;
        uint8_t* sB = sA + a_bytes;
        const PixelTile t = decode_tile(args.g, kb);
        for (int a = 0; a < args.a_atoms; ++a) {
          const int c = mi * 64 * args.a_atoms + a * 64;
          if (args.g.a_mode == 0) {
            ptx::tma_load_4d(&args.mapX[args.tab.map[tap]], &full[s], sA + a * atom, c,
                             t.w0 + args.tab.dw[tap], t.h0 + args.tab.dh[tap], t.n0);
          } else {
            ptx::tma_load_im2col_4d(&args.mapX[0], &full[s], sA + a * atom, c, t.w0, t.h0, t.n0,
                                    static_cast<uint16_t>(args.tab.dw[tap]),
                                    static_cast<uint16_t>(args.tab.dh[tap]));
          }
        }
        for (int j = 0; j < args.nb_atoms; ++j) {
          const int c = ni * args.block_n + j * 64;
          if (args.g.a_mode == 0) {
            ptx::tma_load_4d(&args.mapDy, &full[s], sB + j * atom, c, t.w0, t.h0, t.n0);
          } else {
            // dy is addressed as a [m_total, Cout] matrix through a (C, M, 1, 1) map
            ptx::tma_load_4d(&args.mapDy, &full[s], sB + j * atom, c,
                             static_cast<int>(t.m0), 0, 0);
          }
        }
      }
      RT_FLUSH(0);
    }
  } else if (warp == 1) {
    // one thread, tight loop (see gemm_conv_kernel)
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, args.block_n, 1, 1);
      RT_DECL;
      const long long rt_m0 = RT_NOW();
      (void)rt_m0;
      // MN-major: 64-channel atoms `atom` bytes apart (LBO), 8-pixel groups 1024 B apart (SBO);
      // one instruction consumes 16 pixels = 2048 B of each atom (descriptor step 128).  A wide tile
      // (a_atoms == 4) runs two M = 128 accumulators off the same dy operand.
      const uint64_t dhi = ptx::smem_desc_hi(atom, 1024);
      const uint32_t smem0 = ptx::smem_u32(smem);
      const int ksteps = args.ksteps;
      int s = 0;
      uint32_t ph = 0, acc = 0u;
      for (int kb = kb0; kb < kb1; ++kb) {
        RT_BEGIN;
        ptx::mbar_wait_quiet(&full[s], ph);
        RT_END;
        ptx::tc_fence_after();
        const uint32_t a_addr = smem0 + static_cast<uint32_t>(s) * stage_bytes;
        const uint64_t bd = ptx::smem_desc_at(dhi, a_addr + a_bytes);
        for (int half = 0; half < m_halves; ++half) {
          const uint64_t ad = ptx::smem_desc_at(dhi, a_addr + static_cast<uint32_t>(half) * 2u * atom);
          const uint32_t td = tmem_base + static_cast<uint32_t>(half * args.block_n);
          if (ksteps == 8) {
#pragma unroll
            for (int k = 0; k < 8; ++k) ptx::umma_bf16(td, ad + 128 * k, bd + 128 * k, idesc, k == 0 ? acc : 1u);
          } else if (ksteps == 4) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ptx::umma_bf16(td, ad + 128 * k, bd + 128 * k, idesc, k == 0 ? acc : 1u);
          } else {
            for (int k = 0; k < ksteps; ++k)
              ptx::umma_bf16(td, ad + 128 * k, bd + 128 * k, idesc, k == 0 ? acc : 1u);
          }
        }
        acc = 1u;
        ptx::umma_commit(&empty[s]);
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
        }
      }
      ptx::umma_commit(tmem_full);
      RT_FLUSH(1);
      RT_ADD(7, RT_NOW() - rt_m0);
    }
  } else {
    const long long rt_e0 = RT_NOW();
    (void)rt_e0;
    ptx::mbar_wait(tmem_full, 0);
    if (threadIdx.x == 64) RT_ADD(3, RT_NOW() - rt_e0);
    ptx::tc_fence_after();
    const int quad = warp & 3;
    if (args.tma_out) {
      // every MMA has completed (tmem_full): the operand ring is free — two 4 KB staging tiles per warp
      uint8_t* stg = smem + static_cast<size_t>(quad) * 2 * 4096;
      const uint32_t swz = static_cast<uint32_t>(lane & 7) << 4;
      int bi = 0;
      for (int hc = 0; hc < m_halves * args.block_n; hc += 32) {
        const int half = hc / args.block_n, c0 = hc - half * args.block_n;
        const int row0 = mi * 64 * args.a_atoms + half * 128 + quad * 32;     // warp-uniform
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(hc), r);
        ptx::tmem_ld_wait();
        if (row0 >= args.cin || ni * args.block_n + c0 >= args.cout) continue;
        uint8_t* buf = stg + bi * 4096;
        if (lane == 0) ptx::bulk_wait_read<1>();      // the store that last used this buffer has read it
        __syncwarp();
        const uint32_t rowaddr = ptx::smem_u32(buf) + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          ptx::st_shared_v4(rowaddr + ((static_cast<uint32_t>(j) << 4) ^ swz), r[4 * j], r[4 * j + 1], r[4 * j + 2],
                            r[4 * j + 3]);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_4d(&args.mapOut, buf, ni * args.block_n + c0, row0, tap, split);
          ptx::bulk_commit();
        }
        bi ^= 1;
      }
      if (lane == 0) ptx::bulk_wait_all();
    } else
    for (int hc = 0; hc < m_halves * args.block_n; hc += 32) {
      const int half = hc / args.block_n, c0 = hc - half * args.block_n;
      const int ci = mi * 64 * args.a_atoms + half * 128 + quad * 32 + lane;
      uint32_t r[32];
      ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(hc), r);
      ptx::tmem_ld_wait();
      if (ci >= args.cin) continue;
      const int co0 = ni * args.block_n + c0;
      const bool rmw = args.ws_stride < 0;      // splits == 1: add straight into the gradient
      float* o = args.dw + (rmw ? 0 : static_cast<long long>(split) * args.ws_stride) +
                 (static_cast<long long>(tap) * args.cin + ci) * args.cout + co0;
      if ((args.cout & 3) == 0 && co0 + 32 <= args.cout) {
        wgrad_out32(o, r, rmw);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (co0 + j < args.cout) {
            if (rmw) o[j] += __uint_as_float(r[j]);
            else o[j] = __uint_as_float(r[j]);
          }
      }
    }
    if (args.fuse_reduce) {
      const int et = threadIdx.x - 64;
      const int S = args.splits;
      unsigned int* tk = args.tickets + (blockIdx.x - split * (gridDim.x / S));
      __threadfence();          // this thread's partial-tile stores are visible device-wide ...
      epi_bar_sync();           // ... for all 128 epilogue threads before the ticket is taken
      if (et == 0) {
        atomicAdd(tk, 1u);
        unsigned int seen;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(tk) : "memory");
          if (seen < static_cast<unsigned int>(S)) __nanosleep(100);
        } while (seen < static_cast<unsigned int>(S));
      }
      epi_bar_sync();
      // share `split` of the tile's rows; rows >= cin were never written (and are not summed)
      const int rows = m_halves * 128;
      const int r0 = static_cast<int>(static_cast<long long>(split) * rows / S);
      const int r1 = static_cast<int>(static_cast<long long>(split + 1) * rows / S);
      const int vec_per_row = args.block_n >> 2;
      const int items = (r1 - r0) * vec_per_row;
      const long long tile_off = (static_cast<long long>(tap) * args.cin + mi * 64 * args.a_atoms) * args.cout +
                                 ni * args.block_n;
      for (int i0 = et; i0 < items; i0 += 256) {
        // two items in flight per thread: 2*S independent 16-byte loads
        float4 acc[2];
        long long off[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int i = i0 + u * 128;
          const int r = r0 + i / vec_per_row, c4 = i - (i / vec_per_row) * vec_per_row;
          ok[u] = i < items && mi * 64 * args.a_atoms + r < args.cin;
          off[u] = tile_off + static_cast<long long>(r) * args.cout + c4 * 4;
          acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int sl = 0; sl < S; ++sl) {
          float4 p[2];
#pragma unroll
          for (int u = 0; u < 2; ++u)
            if (ok[u]) p[u] = __ldcg(reinterpret_cast<const float4*>(args.dw + sl * args.ws_stride + off[u]));
#pragma unroll
          for (int u = 0; u < 2; ++u)
            if (ok[u]) {
              acc[u].x += p[u].x;
              acc[u].y += p[u].y;
              acc[u].z += p[u].z;
              acc[u].w += p[u].w;
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (ok[u]) {
            float4* o = reinterpret_cast<float4*>(args.dw_final + off[u]);
            float4 v = *o;
            v.x += acc[u].x;
            v.y += acc[u].y;
            v.z += acc[u].z;
            v.w += acc[u].w;
            *o = v;
          }
      }
      // second pass through the ticket: the last CTA to finish reading leaves the counter at zero
      epi_bar_sync();
      if (et == 0) {
        const unsigned int old = atomicAdd(tk, 1u);
        if (old == 2u * static_cast<unsigned int>(S) - 1u) *tk = 0u;
      }
    }
    if (threadIdx.x == 64) RT_ADD(4, RT_NOW() - rt_e0);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) {
    RT_ADD(5, RT_NOW() - rt_cta0);
    RT_ADD(6, 1);
  }
  if (warp == 1) ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(args.tmem_cols));
}

// ------------------------------------------------------------------ wgrad, halo feed (k x k, stride 1)
// The tap-at-a-time wgrad above re-reads the dy tile for every tap and the x tile for every
// output-channel tile: with 128 x 128 output tiles it moves 64 KB of operands per 128^3 MACs
// and is bound by the L2 -> shared-memory rate, not by the tensor pipe.  Here a pixel tile is
// 16 rows x 8 columns of one image (as in halo_conv_kernel): its INPUT halo and its dy tile are
// loaded ONCE and every filter tap of the group is a different start address into the halo
// (MN-major operand: 8 consecutive pixels = 8 consecutive 128-byte rows = one swizzle group,
// the next output row is SBO = halo_width*128 bytes further).  Each tap owns a TMEM
// accumulator; a CTA keeps all of them for its whole range of pixel tiles (split-K over the
// pixels) and adds them into dw with fp32 red.global at the end.
//   pair mode  (Cin == 64): all taps in one CTA; two taps share one M=128 instruction, the
//               second tap being the "next 64-channel atom" at LBO = its byte distance in the halo.
//   group mode (Cin >= 128): one kernel row (kw taps) per CTA, M = 128 input channels held in two
//               halo boxes LBO apart.
constexpr int kWhMaxGroups = 8;
constexpr int kWhMaxAcc = 8;
struct WgradHaloArgs {
  CUtensorMap mapX;    // box (64 ch, hwb, box_h, 1)
  CUtensorMap mapDy;   // box (64 ch, 8, 16, 1)
  float* dw;           // split-K slices (ws_stride > 0)
  long long ws_stride;
  int n_acc, block_n, nb_atoms, a_atoms, a_box_bytes, a_stride, stages, tmem_cols;
  int hwb, tiles_w, tiles_h, tiles_total, units_ci, units_co, groups, splits;
  int cin, cout, ci_tile, org_w;
  int org_h[kWhMaxGroups];
  int aoff[kWhMaxGroups][kWhMaxAcc];   // byte offset of the accumulator's (first) tap in the halo
  int lbo[kWhMaxGroups][kWhMaxAcc];    // byte distance to the operand's second 64-row atom
  short tap_lo[kWhMaxGroups][kWhMaxAcc];  // tap written by accumulator rows 0..63 (-1: discard)
  short tap_hi[kWhMaxGroups][kWhMaxAcc];  // tap of rows 64..127 (-2: same tap, channels + 64)
  int tma_out;          // partial tiles through TMA stores (see WgradArgs::tma_out)
  CUtensorMap mapOut;   // fp32 4-D view (Cout, Cin, tap, split) of the slices
};

__global__ void __launch_bounds__(192, 1)
wgrad_halo_kernel(const __grid_constant__ WgradHaloArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = args.stages;
  const uint32_t a_bytes = static_cast<uint32_t>(args.a_atoms) * args.a_stride;
  const uint32_t b_bytes = static_cast<uint32_t>(args.nb_atoms) * kABytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(stages) * stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* tmem_full = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  int w = blockIdx.x;
  const int split = w % args.splits;
  w /= args.splits;
  const int g = w % args.groups;
  w /= args.groups;
  const int co_t = w % args.units_co;
  const int ci_t = w / args.units_co;
  const int t0 = static_cast<int>(static_cast<long long>(split) * args.tiles_total / args.splits);
  const int t1 = static_cast<int>(static_cast<long long>(split + 1) * args.tiles_total / args.splits);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&args.mapX);
    ptx::prefetch_tmap(&args.mapDy);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(tmem_slot, static_cast<uint32_t>(args.tmem_cols));
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, tensor-map prefetch and the TMEM allocation above overlap the previous kernel's tail
  MCN_PDL_PROLOGUE();
  const int tiles_per_img = args.tiles_w * args.tiles_h;

  if (warp == 0) {
    if (ptx::elect_one()) {
      const uint32_t tx = static_cast<uint32_t>(args.a_atoms) * args.a_box_bytes + b_bytes;
      for (int t = t0, it = 0; t < t1; ++t, ++it) {
        const int s = it % stages;
        const uint32_t ph = static_cast<uint32_t>(it / stages) & 1u;
        ptx::mbar_wait(&empty[s], ph ^ 1u);
        ptx::mbar_expect_tx(&full[s], tx);
        uint8_t* sA = smem + static_cast<size_t>(s) * stage_bytes;
        uint8_t* sB = sA + a_bytes;
        const int n = t / tiles_per_img;
        const int r = t - n * tiles_per_img;
        const int h0 = (r / args.tiles_w) * 16, w0 = (r % args.tiles_w) * 8;
        for (int a = 0; a < args.a_atoms; ++a)
          ptx::tma_load_4d(&args.mapX, &full[s], sA + static_cast<size_t>(a) * args.a_stride,
                           ci_t * args.ci_tile + a * 64, w0 + args.org_w, h0 + args.org_h[g], n);
        for (int j = 0; j < args.nb_atoms; ++j)
          ptx::tma_load_4d(&args.mapDy, &full[s], sB + static_cast<size_t>(j) * kABytes,
                           co_t * args.block_n + j * 64, w0, h0, n);
      }
    }
  } else if (warp == 1) {
    // one thread, tight loop (see gemm_conv_kernel)
    if (ptx::elect_one() && t1 > t0) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, args.block_n, 1, 1);
      const uint32_t row_pitch = static_cast<uint32_t>(args.hwb) * 128u;   // one output row further
      const uint32_t kstep_a = (2u * row_pitch) >> 4;   // K = 16 pixels = two output rows of the tile
      const uint64_t dhi_b = ptx::smem_desc_hi(kABytes, 1024);
      const uint32_t smem0 = ptx::smem_u32(smem);
      const int n_acc = args.n_acc;
      int s = 0;
      uint32_t ph = 0, accf = 0u;
      for (int t = t0; t < t1; ++t) {
        ptx::mbar_wait_quiet(&full[s], ph);
        ptx::tc_fence_after();
        const uint32_t a_addr = smem0 + static_cast<uint32_t>(s) * stage_bytes;
        const uint64_t bd = ptx::smem_desc_at(dhi_b, a_addr + a_bytes);
        for (int acc = 0; acc < n_acc; ++acc) {
          const uint64_t ad = ptx::make_smem_desc(a_addr + static_cast<uint32_t>(args.aoff[g][acc]),
                                                  static_cast<uint32_t>(args.lbo[g][acc]), row_pitch);
          const uint32_t td = tmem_base + static_cast<uint32_t>(acc * args.block_n);
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ptx::umma_bf16(td, ad + static_cast<uint64_t>(kstep_a * j), bd + 128 * j, idesc,
                           j == 0 ? accf : 1u);
        }
        accf = 1u;
        ptx::umma_commit(&empty[s]);
        if (++s == stages) {
          s = 0;
          ph ^= 1u;
        }
      }
      ptx::umma_commit(tmem_full);
    }
  } else if (t1 > t0) {
    ptx::mbar_wait(tmem_full, 0);
    ptx::tc_fence_after();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    if (args.tma_out) {
      // a warp's 32 accumulator rows are 32 consecutive input channels of ONE tap: one 32 x 32 box
      uint8_t* stg = smem + static_cast<size_t>(quad) * 2 * 4096;
      const uint32_t swz = static_cast<uint32_t>(lane & 7) << 4;
      int bi = 0;
      for (int acc = 0; acc < args.n_acc; ++acc) {
        const int tlo = args.tap_lo[g][acc], thi = args.tap_hi[g][acc];
        const int row_w = quad * 32;
        const int tap = (thi == -2) ? tlo : (row_w < 64 ? tlo : thi);
        const int ci0 = ci_t * args.ci_tile + ((thi == -2) ? row_w : (row_w & 63));
        for (int c0 = 0; c0 < args.block_n; c0 += 32) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                                 static_cast<uint32_t>(acc * args.block_n + c0), r);
          ptx::tmem_ld_wait();
          const int co0 = co_t * args.block_n + c0;
          if (tap < 0 || ci0 >= args.cin || co0 >= args.cout) continue;
          uint8_t* buf = stg + bi * 4096;
          if (lane == 0) ptx::bulk_wait_read<1>();
          __syncwarp();
          const uint32_t rowaddr = ptx::smem_u32(buf) + static_cast<uint32_t>(lane) * 128u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            ptx::st_shared_v4(rowaddr + ((static_cast<uint32_t>(j) << 4) ^ swz), r[4 * j], r[4 * j + 1],
                              r[4 * j + 2], r[4 * j + 3]);
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_4d(&args.mapOut, buf, co0, ci0, tap, split);
            ptx::bulk_commit();
          }
          bi ^= 1;
        }
      }
      if (lane == 0) ptx::bulk_wait_all();
    } else
    for (int acc = 0; acc < args.n_acc; ++acc) {
      const int tlo = args.tap_lo[g][acc], thi = args.tap_hi[g][acc];
      int tap, ci;
      if (thi == -2) {
        tap = tlo;
        ci = ci_t * args.ci_tile + row;
      } else {
        tap = row < 64 ? tlo : thi;
        ci = ci_t * args.ci_tile + (row & 63);
      }
      const bool ok = tap >= 0 && ci < args.cin;
      for (int c0 = 0; c0 < args.block_n; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                               static_cast<uint32_t>(acc * args.block_n + c0), r);
        ptx::tmem_ld_wait();
        if (!ok) continue;
        const int co0 = co_t * args.block_n + c0;
        if (co0 >= args.cout) continue;
        float* o = args.dw + static_cast<long long>(split) * args.ws_stride +
                   (static_cast<long long>(tap) * args.cin + ci) * args.cout + co0;
        wgrad_out32(o, r);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(args.tmem_cols));
}

// ------------------------------------------------------------------ stem convolution (RGB input)
// A k x k stride-2 convolution on a 3-channel image cannot be fed by TMA (6-byte pixels), and an
// explicit im2col matrix costs ~1 GB of HBM writes + two reads per step at batch 256.  Here the
// image is stored with 4 channels (8-byte pixels, 4th = 0) and the filter row is widened to an
// even tap count kwp = kw + 1 (the extra tap has a zero weight): with stride 2 and an even left
// pad the kwp*8 bytes of one filter row of one output pixel then start on a 16-byte boundary, so
// the GEMM A tile is gathered straight from the image with 16-byte cp.async copies into the
// 128-byte-swizzled K-major layout tcgen05 expects.  K index = (r*kwp + s)*4 + c.
//   fprop: M = 128 pixels, N = Cout, K = kh*kwp*4 (224 for 7x7), weights resident in shared memory.
//   wgrad: the same tile read as an MN-major operand (rows = pixels = K), dy tiles by TMA,
//          one TMEM accumulator per 128 k-rows, split-K over pixel tiles, fp32 red.global at the end.
struct StemArgs {
  CUtensorMap mapW;     // fprop: [Cout][Kpad] bf16, box 64 x Cout;  wgrad: dy [M][Cout], box 64 x 128
  CUtensorMap mapOut;   // fprop, TMA-store epilogue (use_tma): y as [M][Cout] bf16, box 64 x 32
  int use_tma;
  int async_full;       // producers signal a stage through cp.async.mbarrier.arrive (on completion)
  const __nv_bfloat16* x4;
  EpiArgs e;            // fprop output / statistics
  float* dw;            // wgrad output [Kpad][Cout]: split-K slices (ws_stride > 0) or the gradient
  long long ws_stride;
  int H, W, Ho, Wo, N, sh, pad_t, pad_l, kh, kw;
  int epr, cpr, rpc, kc, kc_alloc, ksteps_last, kp;   // elements / 16-byte chunks per filter row, rows per 128-byte line
  int stages, tmem_cols, total_tiles, cout, nb_atoms, splits;
  long long m_total;
};

// One producer thread per pixel row of the tile.  (Two threads per row, the filter rows split between
// them — eight gathering warps — measured SLOWER: stem fprop 377 -> 530 us, wgrad 238 -> 257 us; the
// gather is not bound by its instruction stream.)
constexpr int kStemProducers = 128;
constexpr int kStemProducerWarps = kStemProducers / 32;
constexpr int kStemMmaWarp = kStemProducerWarps;
constexpr int kStemThreads = kStemProducers + 32 + 128;

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, bool valid) {
  const uint32_t n = valid ? 16u : 0u;   // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
// the mbarrier receives one arrival from this thread when all its cp.async issued so far have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// One producer thread gathers the receptive field of pixel `m` (row `t` of the tile) into stage `sA`.
// The filter size is a template parameter: with one warp per scheduler the gather is bound by
// the latency of its own address arithmetic, so everything that can fold to a constant must
// (a generic version with run-time kh / kw spent ~8k cycles per tile, mostly in integer divisions).
template <int KH, int KW, int R0, int R1>
__device__ __forceinline__ void stem_gather(const StemArgs& a, uint8_t* sA, int t, uint32_t m) {
  constexpr int EPR = (KW + 1) * 4, CPR = EPR / 8, RPC = 64 / EPR;
  const bool live = m < static_cast<uint32_t>(a.m_total);
  uint32_t q = 0, p = 0, n = 0;
  if (live) {
    const uint32_t wo = static_cast<uint32_t>(a.Wo), ho = static_cast<uint32_t>(a.Ho);
    q = m % wo;
    const uint32_t r = m / wo;
    p = r % ho;
    n = r / ho;
  }
  const int w0 = static_cast<int>(q) * 2 - a.pad_l;                      // even: 16-byte aligned pixel pair
  const int h0 = static_cast<int>(p) * a.sh - a.pad_t;
  const __nv_bfloat16* img = a.x4 + static_cast<long long>(n) * a.H * a.W * 4;
  const uint32_t row_base = ptx::smem_u32(sA) + static_cast<uint32_t>(t) * 128u;
  const uint32_t sw = static_cast<uint32_t>(t & 7);
  bool wok[CPR];
#pragma unroll
  for (int j = 0; j < CPR; ++j) wok[j] = (w0 + 2 * j >= 0) && (w0 + 2 * j + 1 < a.W);
#pragma unroll
  for (int r = R0; r < R1; ++r) {
    const int h = h0 + r;
    const bool hin = live && h >= 0 && h < a.H;
    const __nv_bfloat16* src = img + ((hin ? h : 0) * a.W + w0) * 4;
    const uint32_t chunk_base = row_base + static_cast<uint32_t>(r / RPC) * kABytes;
#pragma unroll
    for (int j = 0; j < CPR; ++j) {
      const bool ok = hin && wok[j];
      cp_async_16(chunk_base + ((static_cast<uint32_t>((r % RPC) * CPR + j) ^ sw) << 4),
                  ok ? src + 8 * j : a.x4, ok);
    }
  }
}
// run-time dispatch on the two supported filter sizes (uniform across the block)
__device__ __forceinline__ void stem_gather_any(const StemArgs& a, uint8_t* sA, int t, uint32_t m) {
  if (a.kw == 7) stem_gather<7, 7, 0, 7>(a, sA, t, m);
  else stem_gather<3, 3, 0, 3>(a, sA, t, m);
}

// Roles (288 threads): warps 0-3 gather A tiles (cp.async, two stages in flight), warp 4 loads the
// weights once and issues the MMAs, warps 5-8 run the shared epilogue (incl. fused BN statistics).
template <bool kTma>
__global__ void __launch_bounds__(kStemThreads, 1)
stem_fprop_kernel(const __grid_constant__ StemArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = args.stages;
  const uint32_t a_stage = static_cast<uint32_t>(args.kc_alloc) * kABytes;
  const uint32_t b_chunk = static_cast<uint32_t>(args.e.block_n) * 128u;
  uint8_t* smemB = smem + static_cast<size_t>(stages) * a_stage;
  // TMA-store epilogue: per-warp staging tiles (1024-aligned: b_chunk is a multiple of 2 KB ... the
  // launcher only enables it when the offset is 1024-aligned)
  constexpr bool tma = kTma;
  const bool async_full = args.async_full != 0;
  uint8_t* stg_all = smemB + static_cast<size_t>(args.kc) * b_chunk;
  uint8_t* tail = stg_all + (tma ? static_cast<size_t>(4 * args.e.stg_bufs) * kStgBytes : 0);
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);
  uint64_t* empty = full + stages;
  uint64_t* wbar = empty + stages;
  uint64_t* tmem_full = wbar + 1;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int total_tiles = args.total_tiles;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&args.mapW);
    if (tma) ptx::prefetch_tmap(&args.mapOut);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full[s], kStemProducers);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(wbar, 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tmem_full[b], 1);
      ptx::mbar_init(&tmem_empty[b], 4);
    }
    ptx::fence_barrier_init();
  }
  if (warp == kStemMmaWarp) {
    ptx::tmem_alloc(tmem_slot, static_cast<uint32_t>(args.tmem_cols));
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, tensor-map prefetch and the TMEM allocation above overlap the previous kernel's tail
  MCN_PDL_PROLOGUE();

  if (warp < kStemProducerWarps) {
    // ---------------- gather producers ----------------
    const int t = threadIdx.x;
    int it = 0;
    RT_DECL;
    const long long rt_p0 = RT_NOW();
    (void)rt_p0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int s = it % stages;
      RT_BEGIN;
      ptx::mbar_wait(&empty[s], (static_cast<uint32_t>(it / stages) & 1u) ^ 1u);
      RT_END;
      stem_gather_any(args, smem + static_cast<size_t>(s) * a_stage, t, static_cast<uint32_t>(tile) * 128u + t);
      if (async_full) {
        // the stage's barrier gets this thread's arrival when its copies land: the producer moves on
        // to the next free stage at once.  (Waiting for tile i-1's copies only AFTER issuing tile i,
        // behind the wait for a free stage, chained gather -> MMA -> gather: role counters showed
        // producers, MMA thread and epilogue each idle 40-50 % of the kernel.)
        cp_async_arrive_noinc(&full[s]);
        continue;
      }
      cp_async_commit();
      if (it > 0) {
        RT_BEGIN;
        cp_async_wait<1>();               // the previous tile's copies have landed
        RT_END2;
        ptx::fence_proxy_async();         // generic-proxy writes -> visible to the tensor core
        ptx::mbar_arrive(&full[(it - 1) % stages]);
      }
    }
    if (t == 0) {      // slot 0: wait for a free stage, slot 8: wait for the copies, slot 9: producer lifetime
      RT_FLUSH(0);
      RT_FLUSH2(8);
      RT_ADD(9, RT_NOW() - rt_p0);
      RT_ADD(6, 1);
    }
    if (it > 0 && !async_full) {
      cp_async_wait<0>();
      ptx::fence_proxy_async();
      ptx::mbar_arrive(&full[(it - 1) % stages]);
    }
  } else if (warp == kStemMmaWarp) {
    // ---------------- weights + MMA issuer ----------------
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(wbar, static_cast<uint32_t>(args.kc) * b_chunk);
      for (int kc = 0; kc < args.kc; ++kc)
        ptx::tma_load_2d(&args.mapW, wbar, smemB + static_cast<size_t>(kc) * b_chunk, kc * 64, 0);
    }
    __syncwarp();
    ptx::mbar_wait(wbar, 0);
    const uint32_t idesc = ptx::make_idesc_bf16(128, args.e.block_n, 0, 0);
    int it = 0;
    RT_DECL;
    const long long rt_m0 = RT_NOW();
    (void)rt_m0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int s = it % stages;
      const int buf = it & 1;
      RT_BEGIN;
      ptx::mbar_wait(&tmem_empty[buf], (static_cast<uint32_t>(it >> 1) & 1u) ^ 1u);
      RT_END2;
      RT_BEGIN;
      ptx::mbar_wait(&full[s], static_cast<uint32_t>(it / stages) & 1u);
      RT_END;
      if (async_full) ptx::fence_proxy_async();   // the gathered tile (generic-proxy writes) -> tensor core
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t a_addr = ptx::smem_u32(smem + static_cast<size_t>(s) * a_stage);
        const uint32_t b_addr = ptx::smem_u32(smemB);
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(buf * args.e.block_n);
        for (int kc = 0; kc < args.kc; ++kc) {
          const int ksteps = (kc == args.kc - 1) ? args.ksteps_last : 4;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t ad = ptx::make_smem_desc(a_addr + kc * kABytes + k * 32, 16, 1024);
            const uint64_t bd = ptx::make_smem_desc(b_addr + kc * b_chunk + k * 32, 16, 1024);
            ptx::umma_bf16(tmem_d, ad, bd, idesc, (kc | k) != 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit(&empty[s]);
        ptx::umma_commit(&tmem_full[buf]);
      }
      __syncwarp();
    }
    if (lane == 0) {
      RT_FLUSH(1);
      RT_FLUSH2(2);
      RT_ADD(7, RT_NOW() - rt_m0);
      RT_ADD(5, RT_NOW() - rt_m0);
    }
  } else {
    // ---------------- epilogue (warps 5..8) ----------------
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    uint32_t* stat_stage = reinterpret_cast<uint32_t*>(tail + kBarRegionBytes);
    float* stat_acc_all = tma ? reinterpret_cast<float*>(stat_stage)
                              : reinterpret_cast<float*>(stat_stage + kStatStageWords);
    float* stat_acc = stat_acc_all + quad * kStatAccWarp;
    uint8_t* stg = stg_all + static_cast<size_t>(quad * args.e.stg_bufs) * kStgBytes;
    int stg_i = 0;
    stat_stage += quad * 32 * kStatRowWords;
    const bool stats = args.e.stats != nullptr;
    if (stats) {
      for (int i = row; i < kStatAccFloats; i += 128) stat_acc_all[i] = 0.f;
      epi_bar_sync();
    }
    int* stat_flag = reinterpret_cast<int*>(tail + kBarRegionBytes - 16);
    int it = 0, pending = 0, since = 0;
    const long long rt_e0 = RT_NOW();
    (void)rt_e0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      if (stats && since == 8) {
        stats_flush(args.e, stat_acc_all, 0, row, pending, stat_flag, false);
        since = 0;
      }
      ++pending;
      ++since;
      const long long m = static_cast<long long>(tile) * 128 + row;
      const bool valid = m < args.m_total;
      const uint32_t tacc = tmem_base + static_cast<uint32_t>(buf * args.e.block_n);
      const uint32_t tph = static_cast<uint32_t>(it >> 1) & 1u;
      if (tma) {
        // the tile is 128 consecutive rows of the [M, Cout] output matrix; the hardware clips the last one
        const OutTile o{tile * 128 + quad * 32, 0, 0};
        if (stats)
          epilogue_tile_tma<1>(args.e, &args.mapOut, o, tacc, quad, lane, valid, 0, &tmem_full[buf], tph,
                               &tmem_empty[buf], stg, stg_i, stat_acc);
        else
          epilogue_tile_tma<0>(args.e, &args.mapOut, o, tacc, quad, lane, valid, 0, &tmem_full[buf], tph,
                               &tmem_empty[buf], stg, stg_i, stat_acc);
      } else {
        epilogue_tile(args.e, tacc, quad, lane, valid, valid ? m * args.cout : 0, 0, &tmem_full[buf], tph,
                      &tmem_empty[buf], stat_stage, stat_acc);
      }
    }
    if (stats && pending > 0) stats_flush(args.e, stat_acc_all, 0, row, pending, stat_flag);
    if (tma && lane == 0) ptx::bulk_wait_all();   // the staged tiles must outlive their stores
    if (row == 0) RT_ADD(4, RT_NOW() - rt_e0);
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kStemMmaWarp) ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(args.tmem_cols));
}

// wgrad: warps 0-3 gather, warp 4 loads dy tiles by TMA and issues the MMAs, warps 5-8 write dw.
__global__ void __launch_bounds__(kStemThreads, 1)
stem_wgrad_kernel(const __grid_constant__ StemArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int stages = args.stages;
  const uint32_t a_stage = static_cast<uint32_t>(args.kc_alloc) * kABytes;
  const uint32_t b_stage = static_cast<uint32_t>(args.nb_atoms) * kABytes;
  const uint32_t stage_bytes = a_stage + b_stage;
  uint8_t* tail = smem + static_cast<size_t>(stages) * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(tail);      // producers' gather
  uint64_t* fullb = full + stages;                         // dy TMA
  uint64_t* empty = fullb + stages;
  uint64_t* tmem_full = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int split = blockIdx.x;
  const bool async_full = args.async_full != 0;
  const int t0 = static_cast<int>(static_cast<long long>(split) * args.total_tiles / args.splits);
  const int t1 = static_cast<int>(static_cast<long long>(split + 1) * args.total_tiles / args.splits);

  // k-chunks the gather never writes (padding up to a multiple of 128 k-rows) must read as zero
  for (int s = 0; s < stages; ++s) {
    uint4* z = reinterpret_cast<uint4*>(smem + static_cast<size_t>(s) * stage_bytes +
                                        static_cast<size_t>(args.kc) * kABytes);
    const int n16 = (args.kc_alloc - args.kc) * kABytes / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  ptx::fence_proxy_async();
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&args.mapW);
    for (int s = 0; s < stages; ++s) {
      ptx::mbar_init(&full[s], kStemProducers);
      ptx::mbar_init(&fullb[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(tmem_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == kStemMmaWarp) {
    ptx::tmem_alloc(tmem_slot, static_cast<uint32_t>(args.tmem_cols));
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // barriers, tensor-map prefetch and the TMEM allocation above overlap the previous kernel's tail
  MCN_PDL_PROLOGUE();
  const int m_tiles = args.kc_alloc / 2;     // accumulators of 128 k-rows

  if (warp < kStemProducerWarps) {
    const int t = threadIdx.x;
    int it = 0;
    for (int tile = t0; tile < t1; ++tile, ++it) {
      const int s = it % stages;
      ptx::mbar_wait(&empty[s], (static_cast<uint32_t>(it / stages) & 1u) ^ 1u);
      stem_gather_any(args, smem + static_cast<size_t>(s) * stage_bytes, t, static_cast<uint32_t>(tile) * 128u + t);
      if (async_full) {
        cp_async_arrive_noinc(&full[s]);
        continue;
      }
      cp_async_commit();
      if (it > 0) {
        cp_async_wait<1>();
        ptx::fence_proxy_async();
        ptx::mbar_arrive(&full[(it - 1) % stages]);
      }
    }
    if (it > 0 && !async_full) {
      cp_async_wait<0>();
      ptx::fence_proxy_async();
      ptx::mbar_arrive(&full[(it - 1) % stages]);
    }
  } else if (warp == kStemMmaWarp) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, args.e.block_n, 1, 1);
    // dy loads run one stage ahead of the MMAs: issue for tile i+1 before consuming tile i
    auto issue_dy = [&](int tile, int it) {
      const int s = it % stages;
      ptx::mbar_wait(&empty[s], (static_cast<uint32_t>(it / stages) & 1u) ^ 1u);
      ptx::mbar_expect_tx(&fullb[s], b_stage);
      uint8_t* sB = smem + static_cast<size_t>(s) * stage_bytes + a_stage;
      for (int j = 0; j < args.nb_atoms; ++j)
        ptx::tma_load_2d(&args.mapW, &fullb[s], sB + static_cast<size_t>(j) * kABytes, j * 64, tile * 128);
    };
    const bool leader = ptx::elect_one();
    if (leader && t0 < t1) issue_dy(t0, 0);
    __syncwarp();
    int it = 0;
    for (int tile = t0; tile < t1; ++tile, ++it) {
      const int s = it % stages;
      const uint32_t ph = static_cast<uint32_t>(it / stages) & 1u;
      if (leader && tile + 1 < t1 && stages > 1) issue_dy(tile + 1, it + 1);
      __syncwarp();
      ptx::mbar_wait(&full[s], ph);
      ptx::mbar_wait(&fullb[s], ph);
      if (async_full) ptx::fence_proxy_async();
      ptx::tc_fence_after();
      if (leader) {
        const uint32_t a_addr = ptx::smem_u32(smem + static_cast<size_t>(s) * stage_bytes);
        const uint32_t b_addr = a_addr + a_stage;
        for (int mt = 0; mt < m_tiles; ++mt)
          for (int k = 0; k < 8; ++k) {
            const uint64_t ad = ptx::make_smem_desc(a_addr + mt * 2 * kABytes + k * 2048, kABytes, 1024);
            const uint64_t bd = ptx::make_smem_desc(b_addr + k * 2048, kABytes, 1024);
            ptx::umma_bf16(tmem_base + static_cast<uint32_t>(mt * args.e.block_n), ad, bd, idesc,
                           (it | k) != 0 ? 1u : 0u);
          }
        ptx::umma_commit(&empty[s]);
        if (tile == t1 - 1) ptx::umma_commit(tmem_full);
      }
      __syncwarp();
    }
  } else if (t1 > t0) {
    ptx::mbar_wait(tmem_full, 0);
    ptx::tc_fence_after();
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    for (int mt = 0; mt < m_tiles; ++mt) {
      const int k = mt * 128 + row;
      const int s2 = (k % args.epr) >> 2;          // tap within the (widened) filter row
      const bool ok = k < args.kp && s2 < args.kw;    // the widening tap carries a zero weight: no gradient
      for (int c0 = 0; c0 < args.e.block_n; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) +
                               static_cast<uint32_t>(mt * args.e.block_n + c0), r);
        ptx::tmem_ld_wait();
        if (c0 >= args.cout) continue;
        if (!ok) {
          // rows of the widening tap / K padding carry no gradient: the slice still needs defined
          // values there because splitk_reduce sums whole slices
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        float* o = args.dw + static_cast<long long>(split) * args.ws_stride +
                   static_cast<long long>(k) * args.cout + c0;
        wgrad_out32(o, r);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == kStemMmaWarp) ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(args.tmem_cols));
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                   cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void* driver_symbol(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  return fn;
}

// bf16 tensor (C innermost) viewed as up to 4 dims; box[0] is always 64 channels (128 B).
int encode_tiled(CUtensorMap* m, const void* base, int rank, const uint64_t* dims,
                 const uint64_t* strides_bytes, const uint32_t* box) {
  static EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(driver_symbol("cuTensorMapEncodeTiled"));
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return MCN_ECUDA;
  }
  cuuint64_t gd[4];
  cuuint64_t gs[3];
  cuuint32_t bx[4];
  cuuint32_t es[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    if (i > 0) gs[i - 1] = strides_bytes[i];
  }
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), gd, gs, bx,
                  es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu box %u %u %u %u",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0),
              box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return MCN_ECUDA;
  }
  return MCN_OK;
}

// fp32 4-D view (Cout, Cin, tap, split) of the split-K slices, box 32 x 32 x 1 x 1 (128-byte rows)
int encode_wgrad_slices(CUtensorMap* m, const float* base, int cout, int cin, int taps, int splits,
                        long long slice_stride) {
  static EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(driver_symbol("cuTensorMapEncodeTiled"));
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable");
    return MCN_ECUDA;
  }
  cuuint64_t gd[4] = {(cuuint64_t)cout, (cuuint64_t)cin, (cuuint64_t)taps, (cuuint64_t)splits};
  cuuint64_t gs[3] = {(cuuint64_t)cout * 4, (cuuint64_t)cin * cout * 4, (cuuint64_t)slice_stride * 4};
  cuuint32_t bx[4] = {32, 32, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (wgrad slices) failed (%d): cout %d cin %d taps %d splits %d stride %lld",
              (int)r, cout, cin, taps, splits, slice_stride);
    return MCN_ECUDA;
  }
  return MCN_OK;
}

int encode_im2col(CUtensorMap* m, const void* base, int C, int W, int H, int N, int low_w,
                  int low_h, int up_w, int up_h, int str_w, int str_h) {
  static EncodeIm2colFn fn =
      reinterpret_cast<EncodeIm2colFn>(driver_symbol("cuTensorMapEncodeIm2col"));
  if (!fn) {
    set_error("cuTensorMapEncodeIm2col unavailable");
    return MCN_ECUDA;
  }
  cuuint64_t gd[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t gs[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  int lo[2] = {low_w, low_h};
  int up[2] = {up_w, up_h};
  cuuint32_t es[4] = {1, (cuuint32_t)str_w, (cuuint32_t)str_h, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gd, gs, lo, up,
                  64, 128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeIm2col failed (%d): C%d W%d H%d N%d lo(%d,%d) up(%d,%d) s(%d,%d)",
              (int)r, C, W, H, N, low_w, low_h, up_w, up_h, str_w, str_h);
    return MCN_ECUDA;
  }
  return MCN_OK;
}

// NHWC bf16 tensor -> (C, W, H, N) tiled map with the given box.
int encode_nhwc(CUtensorMap* m, const void* base, int C, int W, int H, int N, int bw, int bh,
                int bn) {
  uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)N};
  uint64_t st[4] = {2, (uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
  uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bh, (uint32_t)bn};
  return encode_tiled(m, base, 4, dims, st, box);
}
// [rows, K] bf16 matrix -> 2-D map, box 64 x box_rows.
int encode_matrix(CUtensorMap* m, const void* base, long long rows, int K, int box_rows) {
  uint64_t dims[2] = {(uint64_t)K, (uint64_t)rows};
  uint64_t st[2] = {2, (uint64_t)K * 2};
  uint32_t box[2] = {64, (uint32_t)box_rows};
  return encode_tiled(m, base, 2, dims, st, box);
}

int pick_block_n(int n) {
  if (n <= 64) return 64;
  if (n <= 128) return 128;
  if (n % 256 == 0) return 256;
  return 128;
}
int tmem_cols_for(int block_n) { return block_n <= 64 ? 64 : (block_n <= 128 ? 128 : 256); }

// Pixel-space tiling for the box mode: TW x TH x TN <= 128 output pixels per tile.
void pick_box(int Wo, int Ho, int Nb, int* TW, int* TH, int* TN) {
  int tw = std::min(Wo, 128);
  int th = 1, tn = 1;
  if (tw == Wo) {
    int max_th = std::min(Ho, 128 / tw);
    int best = 1, best_tiles = Ho;
    for (int c = 1; c <= max_th; ++c) {
      int tiles = (Ho + c - 1) / c;
      if (tiles < best_tiles) {
        best_tiles = tiles;
        best = c;
      }
    }
    th = best;
    if (th == Ho) tn = std::max(1, std::min(Nb, 128 / (tw * th)));
  }
  *TW = tw;
  *TH = th;
  *TN = tn;
}

struct PixelSpace {   // the pixel grid the GEMM M (fprop/dgrad) or K (wgrad) dimension walks
  int W, H, N;
};

void fill_geom_tiled(TileGeom* g, const PixelSpace& ps) {
  g->a_mode = 0;
  pick_box(ps.W, ps.H, ps.N, &g->TW, &g->TH, &g->TN);
  g->rows_box = g->TW * g->TH * g->TN;
  g->tiles_w = (ps.W + g->TW - 1) / g->TW;
  g->tiles_h = (ps.H + g->TH - 1) / g->TH;
  g->Wo = ps.W;
  g->Ho = ps.H;
  g->Nb = ps.N;
  g->str_w = g->str_h = 1;
  g->low_w = g->low_h = 0;
  g->m_total = static_cast<long long>(ps.W) * ps.H * ps.N;
}
int tiles_m_of(const TileGeom& g) {
  if (g.a_mode == 0) return g.tiles_w * g.tiles_h * ((g.Nb + g.TN - 1) / g.TN);
  return static_cast<int>((g.m_total + 127) / 128);
}

int smem_optin_limit() {
  static int lim = 0;
  if (!lim) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (lim <= 0) lim = 227 * 1024;
  }
  return lim;
}

// Fused statistics go through the workspace's exact accumulator (xsum.cuh).
int attach_xs(EpiArgs* e, int tiles_m, int tiles_n) {
  e->xs = nullptr;
  e->xs_counter = nullptr;
  e->tiles_m = tiles_m;
  if (e->stats == nullptr) return MCN_OK;
  if (tiles_n > kWsCounters) {
    set_error("conv fused statistics: too many n-tiles (%d)", tiles_n);
    return MCN_EINVAL;
  }
  const XsScratch sc = xs_scratch(2 * e->n_total, "conv fused statistics");
  if (sc.limbs == nullptr) return MCN_EINVAL;
  e->xs = sc.limbs;
  e->xs_counter = sc.counter;
  return MCN_OK;
}

// Deterministic split-K: every split writes its partial dw into a private slice of the workspace
// and splitk_reduce sums the slices in split order.
struct SplitPlan {
  float* base;         // first slice (nullptr in query / atomics mode)
  long long stride;    // elements per slice
  int splits;
};
// n: elements of dw; want: preferred split count.  query != nullptr: only report the bytes wanted.
int plan_splits(long long n, int want, SplitPlan* sp, long long* query, const char* who) {
  sp->stride = (n + 63) / 64 * 64;
  sp->splits = want;
  sp->base = nullptr;
  if (query != nullptr) {
    *query = kWsSplitOff + static_cast<long long>(want) * sp->stride * 4;
    return MCN_OK;
  }
  const Workspace w = current_workspace();
  if (w.base == nullptr) {
    set_error("%s: no workspace registered for this device (mcn_set_workspace)", who);
    return MCN_EINVAL;
  }
  const long long cap = (w.bytes - kWsSplitOff) / (sp->stride * 4);
  if (cap < 1) {
    set_error("%s: workspace too small: %lld bytes, one split needs %lld", who, w.bytes,
              kWsSplitOff + sp->stride * 4);
    return MCN_EINVAL;
  }
  sp->splits = static_cast<int>(std::min<long long>(want, cap));
  sp->base = reinterpret_cast<float*>(w.base + kWsSplitOff);
  return MCN_OK;
}

// Fused batch-norm backward reduction riding on a dgrad launch (mcn_conv2d_dgrad_tc_bnred).
struct BnRed {
  const void* x;   // the BN input (= the producing conv's output), same shape as dx
  const float *mean, *invstd, *gamma, *beta;
  int act;         // MCN_ACT_NONE or MCN_ACT_RELU
  double* sums;    // [2*C]: += sum dz | sum dz*x
  float* out1;     // optional: final sums written by the kernel itself (sum_dz += ...,
  float* out2;     //           sum_dz_xhat += ...): no mcn_bn_bwd_finalize launch
};
void attach_red(EpiArgs* e, const BnRed* red) {
  e->red = 0;
  if (red == nullptr) return;
  e->stats = red->sums;
  e->red = red->act == MCN_ACT_RELU ? 2 : 1;
  e->red_mean = red->mean;
  e->red_invstd = red->invstd;
  e->red_gamma = red->gamma;
  e->red_beta = red->beta;
  e->red_out1 = red->out1;
  e->red_out2 = red->out2;
}

// dense_out: the output is a plain [m_total, n_total] matrix whose row order is the tile order (1x1
// convolutions and the im2col feed) — the precondition of the TMA-store epilogue.
int launch_gemm_conv(GemmConvArgs& a, int tiles_m, bool dense_out, cudaStream_t st,
                     const BnRed* red = nullptr) {
  a.e.block_n = a.block_n;
  a.total_tiles = tiles_m * a.tiles_n;
  attach_red(&a.e, red);
  {
    const int rc = attach_xs(&a.e, tiles_m, a.tiles_n);
    if (rc) return rc;
  }
#ifdef MCN_ENABLE_DEBUG_SKIP
  {
    static int dbg = -1;      // timing experiments: MCN_DEBUG_SKIP bit0 no stores, bit2 no TMEM loads
    if (dbg < 0) {
      const char* e = getenv("MCN_DEBUG_SKIP");
      dbg = e ? atoi(e) : 0;
    }
    a.e.debug = dbg;
  }
#else
  a.e.debug = 0;
#endif
  static int tma_enabled = -1;
  if (tma_enabled < 0) {
    const char* e = getenv("MCN_TMA_STORE");      // MCN_TMA_STORE=0: direct epilogue everywhere (A/B)
    tma_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  // n_total % 8 == 0: the output row pitch is a multiple of 16 bytes; a partial last 64-channel chunk
  // (EfficientNet's 24 / 40 / 96 / 144 ... channels) is clipped by the hardware, its accumulator columns
  // are exact zeros (the weight rows beyond Cout are zero-filled on load) and stats_flush skips them.
  // MCN_TMA_STORE_MIN64=1 restores the n_total % 64 == 0 rule (A/B).
  static int min64 = -1;
  if (min64 < 0) {
    const char* e = getenv("MCN_TMA_STORE_MIN64");
    min64 = (e && e[0] == '1') ? 1 : 0;
  }
  const bool tma = tma_enabled && dense_out && !a.e.out_f32 && a.e.bias == nullptr &&
                   a.e.n_total % (min64 ? 64 : 8) == 0 && reinterpret_cast<uintptr_t>(a.e.out) % 16 == 0 &&
                   a.g.m_total < (1LL << 31);
  a.e.out_rank4 = 0;
  if (tma) {
    const int rc = encode_matrix(&a.mapOut, a.e.out, a.g.m_total, a.e.n_total, 32);
    if (rc) return rc;
  }
  if (red != nullptr) {
    if (!tma || a.e.accumulate) {
      set_error("dgrad_tc_bnred: this geometry has no TMA-store epilogue (mcn_conv2d_dgrad_bnred_supported)");
      return MCN_EINVAL;
    }
    const int rc = encode_matrix(&a.mapRed, red->x, a.g.m_total, a.e.n_total, 32);
    if (rc) return rc;
  }
  const int grid_n = std::min(a.total_tiles, num_sms());
  // weight-stationary when the CTA's weight slab is small and it sees one n-tile only
  static int ws_enabled = -1;
  if (ws_enabled < 0) {
    const char* e = getenv("MCN_WEIGHT_STATIONARY");
    ws_enabled = (e && e[0] == '0') ? 0 : ((e && e[0] == '2') ? 2 : 1);   // 2: also for tiny grids (tests)
  }
  const size_t b_slab = static_cast<size_t>(a.taps) * a.k_chunks * a.block_n * 128;
  a.b_stationary = (ws_enabled && b_slab <= 128 * 1024 && grid_n % a.tiles_n == 0 &&
                    (a.total_tiles >= 2 * grid_n || ws_enabled == 2)) ? 1 : 0;
  const uint32_t stage_bytes = a.b_stationary ? kABytes : kABytes + a.block_n * 128;
  // one persistent CTA per SM: the whole shared memory is the TMA ring, minus the epilogue's share
  auto fixed_for = [&](int bufs) {
    const size_t epi = tma ? static_cast<size_t>(4 * bufs + (red ? 4 : 0)) * kStgBytes + (a.e.stats ? kStatAccBytes : 0)
                           : (a.e.stats ? kEpiStageBytes + kStatAccBytes : 0);
    return kBarRegionBytes + epi + 1024 + (a.b_stationary ? b_slab : 0);
  };
  auto stages_for = [&](int bufs) {
    const size_t ring = std::min<size_t>(200 * 1024, static_cast<size_t>(smem_optin_limit()) - fixed_for(bufs));
    return std::max(2, std::min(8, static_cast<int>(ring / stage_bytes)));
  };
  // two staging tiles per warp unless the second one costs a ring stage the kernel is short of
  a.e.stg_bufs = (tma && stages_for(2) < stages_for(1) && stages_for(2) < 4) ? 1 : 2;
  const int stages = stages_for(a.e.stg_bufs);
  a.stages = stages;
  a.tmem_cols = 2 * tmem_cols_for(a.block_n);   // double-buffered accumulator
  const size_t smem = static_cast<size_t>(stages) * stage_bytes + fixed_for(a.e.stg_bufs);
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(gemm_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_optin_limit()) != cudaSuccess ||
        cudaFuncSetAttribute(gemm_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_optin_limit()) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(gemm_conv_kernel) failed");
      return MCN_ECUDA;
    }
    configured = true;
  }
  dim3 grid(static_cast<unsigned>(grid_n));
  if (tma)
    ::mcn::launch(gemm_conv_kernel<true>, grid, 192, smem, st, a);
  else
    ::mcn::launch(gemm_conv_kernel<false>, grid, 192, smem, st, a);
  return after_launch("gemm_conv_kernel");
}

// Halo mode (a_mode 2): stride-1 k x k convolution whose spatial size makes 16x8 tiles efficient.
// `flip` selects the dgrad tap order (correlation with mirrored taps).
constexpr double kHaloMinEff = 0.85;
constexpr int kHaloMinHw = 48 * 48;
bool halo_eligible(int cin, int kh, int kw, int sh, int sw, int dh, int dw, int Ho, int Wo) {
  if (sh != 1 || sw != 1 || (kh == 1 && kw == 1) || cin % 64 != 0 || kh * kw > kMaxTaps) return false;
  const int hwb = 8 + (kw - 1) * dw, hhb = 16 + (kh - 1) * dh;
  if (hwb > 256 || hhb > 256 || hwb * hhb * 128 > 40 * 1024) return false;
  // tile efficiency: fraction of the 16x8 tiles that is real output.  Measured on B200
  // (profiles/r01_halo_vs_im2col.txt): the halo feed wins on the large maps (56x56: 187 -> 144 us)
  // and loses below ~48x48, where the im2col feed's exact 128-pixel tiles matter more.
  const double eff = (double)(Ho * Wo) / ((double)((Ho + 15) / 16 * 16) * ((Wo + 7) / 8 * 8));
  // MCN_HALO_MIN_EFF / MCN_HALO_MIN_HW override the two thresholds (A/B; plan.py mirrors them)
  static double min_eff = -1.0;
  static int min_hw = -1;
  if (min_eff < 0.0) {
    const char* e = getenv("MCN_HALO_MIN_EFF");
    const char* h = getenv("MCN_HALO_MIN_HW");
    min_hw = h ? atoi(h) : kHaloMinHw;
    min_eff = e ? atof(e) : kHaloMinEff;
  }
  return eff >= min_eff && Ho * Wo >= min_hw;
}

int launch_halo(const void* in, int C_in, int W_in, int H_in, int N, const void* wmat, int taps_rows,
                int n_total, int kh, int kw, int dh, int dw, int org_h, int org_w, bool flip,
                int Ho, int Wo, void* out, int out_f32, const float* bias, int accumulate,
                double* stats, cudaStream_t st, const BnRed* red = nullptr) {
  HaloArgs a;
  std::memset(&a, 0, sizeof(a));
  a.e.stats = stats;
  attach_red(&a.e, red);
  stats = a.e.stats;
  int rc;
  a.hwb = 8 + (kw - 1) * dw;
  a.hhb = 16 + (kh - 1) * dh;
  if ((rc = encode_nhwc(&a.mapA, in, C_in, W_in, H_in, N, a.hwb, a.hhb, 1))) return rc;
  a.e.block_n = pick_block_n(n_total);
  a.tiles_n = (n_total + a.e.block_n - 1) / a.e.block_n;
  a.taps = kh * kw;
  a.k_chunks = C_in / 64;
  if ((rc = encode_matrix(&a.mapB, wmat, (long long)a.taps * taps_rows, C_in, a.e.block_n))) return rc;
  a.e.n_total = n_total;
  a.e.out = out;
  a.e.out_f32 = out_f32;
  a.e.bias = bias;
  a.e.accumulate = accumulate;
  a.e.vec_ok = (n_total % 16 == 0) && (reinterpret_cast<uintptr_t>(out) % 32 == 0);
  a.Wo = Wo;
  a.Ho = Ho;
  a.Nb = N;
  a.tiles_w = (Wo + 7) / 8;
  a.tiles_h = (Ho + 15) / 16;
  a.org_w = org_w;
  a.org_h = org_h;
  a.out_sw = n_total;
  a.out_sh = (long long)Wo * n_total;
  a.out_sn = (long long)Ho * Wo * n_total;
  for (int r = 0; r < kh; ++r)
    for (int s2 = 0; s2 < kw; ++s2) {
      const int t = r * kw + s2;
      a.brow[t] = t * taps_rows;
      const int rr = flip ? (kh - 1 - r) : r, ss = flip ? (kw - 1 - s2) : s2;
      a.toff[t] = (short)(rr * dh * a.hwb + ss * dw);
    }
  const uint32_t b_bytes = a.e.block_n * 128;
  a.halo_stride = ((a.hwb * a.hhb * 128) + 1023) / 1024 * 1024;
  const long long b_all = (long long)a.taps * a.k_chunks * b_bytes;
  a.b_stationary = (a.tiles_n == 1 && b_all <= 100 * 1024) ? 1 : 0;
  a.a_stages = 3;
  static int tma_enabled = -1;
  if (tma_enabled < 0) {
    const char* e = getenv("MCN_TMA_STORE");
    tma_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  const bool tma = tma_enabled && !out_f32 && bias == nullptr && n_total % 64 == 0 &&
                   reinterpret_cast<uintptr_t>(out) % 16 == 0;
  a.e.out_rank4 = 1;
  a.e.stg_bufs = 2;
  if (tma && (rc = encode_nhwc(&a.mapOut, out, n_total, Wo, Ho, N, 8, 4, 1))) return rc;
  if (red != nullptr) {
    if (!tma || accumulate) {
      set_error("dgrad_tc_bnred: this geometry has no TMA-store epilogue (mcn_conv2d_dgrad_bnred_supported)");
      return MCN_EINVAL;
    }
    if ((rc = encode_nhwc(&a.mapRed, red->x, n_total, Wo, Ho, N, 8, 4, 1))) return rc;
  }
  const size_t epi_bytes = tma ? static_cast<size_t>(4 * a.e.stg_bufs + (red ? 4 : 0)) * kStgBytes + (stats ? kStatAccBytes : 0)
                               : (stats ? kEpiStageBytes + kStatAccBytes : 0);
  const long long budget = static_cast<long long>(smem_optin_limit()) - 1024 - kBarRegionBytes -
                           static_cast<long long>(epi_bytes) - (long long)a.a_stages * a.halo_stride;
  if (a.b_stationary && b_all > budget) a.b_stationary = 0;   // resident weights do not fit: stream them
  if (a.b_stationary) {
    a.b_stages = 1;
  } else {
    a.b_stages = (int)std::max<long long>(2, std::min<long long>(8, budget / b_bytes));
  }
  static int use_bo = -1;
  if (use_bo < 0) {
    const char* e = getenv("MCN_HALO_BASE_OFFSET");
    use_bo = (e && e[0] == '1') ? 1 : 0;
  }
  a.use_base_offset = use_bo;
#ifdef MCN_ENABLE_DEBUG_SKIP
  {
    const char* e = getenv("MCN_DEBUG_SKIP");
    a.e.debug = e ? atoi(e) : 0;
  }
#else
  a.e.debug = 0;
#endif
  a.tmem_cols = 2 * tmem_cols_for(a.e.block_n);
  a.total_tiles = a.tiles_w * a.tiles_h * N * a.tiles_n;
  const int nb_slots = a.b_stationary ? a.taps * a.k_chunks : a.b_stages;
  size_t smem = (size_t)a.a_stages * a.halo_stride + (size_t)nb_slots * b_bytes + kBarRegionBytes +
                epi_bytes + 1024;
  if (smem > static_cast<size_t>(smem_optin_limit())) {
    set_error("halo_conv_kernel: %zu bytes of shared memory needed", smem);
    return MCN_EINVAL;
  }
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(halo_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_optin_limit()) != cudaSuccess ||
        cudaFuncSetAttribute(halo_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_optin_limit()) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(halo_conv_kernel) failed");
      return MCN_ECUDA;
    }
    configured = true;
  }
  if ((rc = attach_xs(&a.e, a.total_tiles / a.tiles_n, a.tiles_n))) return rc;
  dim3 grid((unsigned)std::min(a.total_tiles, num_sms()));
  if (tma)
    ::mcn::launch(halo_conv_kernel<true>, grid, 224, smem, st, a);
  else
    ::mcn::launch(halo_conv_kernel<false>, grid, 224, smem, st, a);
  return after_launch("halo_conv_kernel");
}

bool tc_shape_ok(const mcn_conv_desc* d) {
  return d->Cin % 8 == 0 && d->Cout % 8 == 0 && d->kh * d->kw <= kMaxTaps;
}

}  // namespace
}  // namespace mcn

using namespace mcn;

static int fprop_tc_impl(const mcn_conv_desc* d, const void* x, const void* w_ohwi,
                         const float* bias, void* y, int y_dtype, int a_mode, int accumulate,
                         double* bn_sums, void* stream) {
  MCN_REQUIRE(d && x && w_ohwi && y, "fprop_tc: null argument");
  if (bn_sums != nullptr) {
    MCN_REQUIRE(y_dtype == MCN_BF16 && !accumulate && d->Cout % 64 == 0 &&
                    reinterpret_cast<uintptr_t>(y) % 32 == 0,
                "fprop_tc: fused BN statistics need a bf16, 32-byte aligned, non-accumulating output "
                "with Cout %% 64 == 0 (Cout=%d)", d->Cout);
  }
  MCN_REQUIRE(d->Cin % 8 == 0, "fprop_tc: Cin=%d must be a multiple of 8 (16-byte TMA rows)", d->Cin);
  MCN_REQUIRE(d->kh * d->kw <= kMaxTaps, "fprop_tc: too many taps");
  const bool pointwise = d->kh == 1 && d->kw == 1 && d->sh == 1 && d->sw == 1;
  if (a_mode == 2) {
    if (halo_eligible(d->Cin, d->kh, d->kw, d->sh, d->sw, d->dh, d->dw, d->Ho, d->Wo))
      return launch_halo(x, d->Cin, d->W, d->H, d->N, w_ohwi, d->Cout, d->Cout, d->kh, d->kw, d->dh,
                         d->dw, -d->pad_t, -d->pad_l, false, d->Ho, d->Wo, y, y_dtype == MCN_F32, bias,
                         accumulate, bn_sums, static_cast<cudaStream_t>(stream));
    a_mode = 1;
  }
  if (a_mode == 1 && (d->Cin % 64 != 0 || pointwise)) a_mode = 0;
  MCN_REQUIRE(a_mode == 1 || (d->sh == 1 && d->sw == 1) || (d->kh == 1 && d->kw == 1),
              "fprop_tc: box mode supports stride 1 (or 1x1 kernels) only");
  GemmConvArgs a;
  std::memset(&a, 0, sizeof(a));
  int rc;
  const int taps = d->kh * d->kw;
  if (pointwise) {
    // pure GEMM: pixels form one long row
    PixelSpace ps{static_cast<int>(std::min<long long>((long long)d->N * d->H * d->W, 1LL << 30)), 1, 1};
    fill_geom_tiled(&a.g, ps);
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)ps.W, 1, 1};
    uint64_t stb[4] = {2, (uint64_t)d->Cin * 2, (uint64_t)ps.W * d->Cin * 2,
                       (uint64_t)ps.W * d->Cin * 2};
    uint32_t box[4] = {64, (uint32_t)a.g.TW, 1, 1};
    if ((rc = encode_tiled(&a.mapA[0], x, 4, dims, stb, box))) return rc;
    a.out_sw = d->Cout;
    a.out_sh = a.out_sn = 0;
  } else if (a_mode == 0) {
    PixelSpace ps{d->Wo, d->Ho, d->N};
    fill_geom_tiled(&a.g, ps);
    if (d->sh == 1 && d->sw == 1) {
      if ((rc = encode_nhwc(&a.mapA[0], x, d->Cin, d->W, d->H, d->N, a.g.TW, a.g.TH, a.g.TN)))
        return rc;
    } else {
      // 1x1 strided: the sampled pixels form a strided view of x
      uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->N};
      uint64_t stb[4] = {2, (uint64_t)d->sw * d->Cin * 2, (uint64_t)d->sh * d->W * d->Cin * 2,
                         (uint64_t)d->H * d->W * d->Cin * 2};
      uint32_t box[4] = {64, (uint32_t)a.g.TW, (uint32_t)a.g.TH, (uint32_t)a.g.TN};
      if ((rc = encode_tiled(&a.mapA[0], x, 4, dims, stb, box))) return rc;
    }
    a.out_sw = d->Cout;
    a.out_sh = (long long)d->Wo * d->Cout;
    a.out_sn = (long long)d->Ho * d->Wo * d->Cout;
  } else {
    a.g.a_mode = 1;
    a.g.Wo = d->Wo;
    a.g.Ho = d->Ho;
    a.g.Nb = d->N;
    a.g.m_total = (long long)d->N * d->Ho * d->Wo;
    a.g.str_w = d->sw;
    a.g.str_h = d->sh;
    a.g.low_w = -d->pad_l;
    a.g.low_h = -d->pad_t;
    // upper corner chosen so the window base takes exactly Wo (Ho) positions
    const int up_w = (d->Wo - 1) * d->sw - d->pad_l - (d->W - 1);
    const int up_h = (d->Ho - 1) * d->sh - d->pad_t - (d->H - 1);
    if ((rc = encode_im2col(&a.mapA[0], x, d->Cin, d->W, d->H, d->N, a.g.low_w, a.g.low_h, up_w,
                            up_h, d->sw, d->sh)))
      return rc;
    a.out_sw = d->Cout;
    a.out_sh = (long long)d->Wo * d->Cout;
    a.out_sn = (long long)d->Ho * d->Wo * d->Cout;
  }
  for (int i = 1; i < 4; ++i) a.mapA[i] = a.mapA[0];
  a.block_n = pick_block_n(d->Cout);
  a.tiles_n = (d->Cout + a.block_n - 1) / a.block_n;
  if ((rc = encode_matrix(&a.mapB, w_ohwi, (long long)taps * d->Cout, d->Cin, a.block_n))) return rc;
  a.taps = taps;
  a.k_chunks = (d->Cin + 63) / 64;
  {
    int rem = d->Cin - (a.k_chunks - 1) * 64;
    a.ksteps_last = (rem + 15) / 16;
  }
  a.e.n_total = d->Cout;
  a.e.out = y;
  a.e.stats = bn_sums;
  a.e.out_f32 = (y_dtype == MCN_F32);
  a.e.bias = bias;
  a.e.vec_ok = (d->Cout % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 32 == 0);
  a.e.accumulate = accumulate;
  for (int r = 0; r < d->kh; ++r)
    for (int s = 0; s < d->kw; ++s) {
      int t = r * d->kw + s;
      a.tab.brow[t] = t * d->Cout;
      a.tab.map[t] = 0;
      if (a.g.a_mode == 0 && !pointwise) {
        a.tab.dh[t] = (short)(r * d->dh - d->pad_t);
        a.tab.dw[t] = (short)(s * d->dw - d->pad_l);
      } else {
        a.tab.dh[t] = (short)(r * d->dh);
        a.tab.dw[t] = (short)(s * d->dw);
      }
    }
  return launch_gemm_conv(a, tiles_m_of(a.g), pointwise || a.g.a_mode == 1, static_cast<cudaStream_t>(stream));
}

extern "C" int mcn_conv2d_fprop_tc(const mcn_conv_desc* d, const void* x, const void* w_ohwi,
                                   const float* bias, void* y, int y_dtype, int a_mode,
                                   int accumulate, void* stream) {
  return fprop_tc_impl(d, x, w_ohwi, bias, y, y_dtype, a_mode, accumulate, nullptr, stream);
}
extern "C" int mcn_conv2d_fprop_tc_stats(const mcn_conv_desc* d, const void* x, const void* w_ohwi,
                                         const float* bias, void* y, int a_mode, double* bn_sums,
                                         void* stream) {
  MCN_REQUIRE(bn_sums != nullptr, "fprop_tc_stats: bn_sums is null");
  return fprop_tc_impl(d, x, w_ohwi, bias, y, MCN_BF16, a_mode, 0, bn_sums, stream);
}

// dgrad.
//   dx[n,h,w,ci] = sum_{r,s,co} dy[n, (h + pad_t - r*dh)/sh, (w + pad_l - s*dw)/sw, co] * W[r,s,ci,co]
// over the taps for which the divisions are exact.  For stride 1 this is a correlation of dy with
// the mirrored taps.  For stride > 1 the input pixels split into sh*sw phases (h % sh, w % sw);
// each phase is a stride-1 correlation of dy with the subset of taps of matching parity, written
// to a strided view of dx — one launch per phase, no zero-insertion, no wasted MACs.
static int dgrad_phase(const mcn_conv_desc* d, const void* dy, const void* w_hwio, void* dx,
                       int dx_dtype, int a_mode, int accumulate, int ph, int pw, bool* empty,
                       cudaStream_t st,
                       const BnRed* red = nullptr) {
  GemmConvArgs a;
  std::memset(&a, 0, sizeof(a));
  int rc;
  const size_t esz = (dx_dtype == MCN_F32) ? 4 : 2;
  // pixel grid of this phase
  const int Ha = (d->H - ph + d->sh - 1) / d->sh;
  const int Wa = (d->W - pw + d->sw - 1) / d->sw;
  // taps of matching parity and their dy offsets e (dy row = a + e)
  int nt = 0, e_h[kMaxTaps], e_w[kMaxTaps], tap_id[kMaxTaps];
  int min_eh = 1 << 30, min_ew = 1 << 30;
  for (int r = 0; r < d->kh; ++r) {
    int th = ph + d->pad_t - r * d->dh;
    if (((th % d->sh) + d->sh) % d->sh != 0) continue;
    for (int s = 0; s < d->kw; ++s) {
      int tw = pw + d->pad_l - s * d->dw;
      if (((tw % d->sw) + d->sw) % d->sw != 0) continue;
      // exact division (th, tw are multiples of the stride, possibly negative)
      e_h[nt] = th / d->sh;
      e_w[nt] = tw / d->sw;
      tap_id[nt] = r * d->kw + s;
      min_eh = std::min(min_eh, e_h[nt]);
      min_ew = std::min(min_ew, e_w[nt]);
      ++nt;
    }
  }
  *empty = (nt == 0 || Ha <= 0 || Wa <= 0);
  if (*empty) return MCN_OK;
  const bool unit = (d->sh == 1 && d->sw == 1);
  const bool pointwise = unit && d->kh == 1 && d->kw == 1;
  if (a_mode == 1 && (d->Cout % 64 != 0 || pointwise)) a_mode = 0;
  if (pointwise) {
    PixelSpace ps{static_cast<int>((long long)d->N * d->H * d->W), 1, 1};
    fill_geom_tiled(&a.g, ps);
    uint64_t dims[4] = {(uint64_t)d->Cout, (uint64_t)ps.W, 1, 1};
    uint64_t stb[4] = {2, (uint64_t)d->Cout * 2, (uint64_t)ps.W * d->Cout * 2,
                       (uint64_t)ps.W * d->Cout * 2};
    uint32_t box[4] = {64, (uint32_t)a.g.TW, 1, 1};
    if ((rc = encode_tiled(&a.mapA[0], dy, 4, dims, stb, box))) return rc;
    a.out_sw = d->Cin;
  } else if (a_mode == 0) {
    PixelSpace ps{Wa, Ha, d->N};
    fill_geom_tiled(&a.g, ps);
    if ((rc = encode_nhwc(&a.mapA[0], dy, d->Cout, d->Wo, d->Ho, d->N, a.g.TW, a.g.TH, a.g.TN)))
      return rc;
  } else {
    a.g.a_mode = 1;
    a.g.Wo = Wa;
    a.g.Ho = Ha;
    a.g.Nb = d->N;
    a.g.m_total = (long long)d->N * Ha * Wa;
    a.g.str_w = a.g.str_h = 1;
    a.g.low_w = min_ew;
    a.g.low_h = min_eh;
    const int up_w = (Wa - 1) + a.g.low_w - (d->Wo - 1);
    const int up_h = (Ha - 1) + a.g.low_h - (d->Ho - 1);
    if ((rc = encode_im2col(&a.mapA[0], dy, d->Cout, d->Wo, d->Ho, d->N, a.g.low_w, a.g.low_h,
                            up_w, up_h, 1, 1)))
      return rc;
  }
  if (!pointwise) {
    a.out_sw = (long long)d->sw * d->Cin;
    a.out_sh = (long long)d->sh * d->W * d->Cin;
    a.out_sn = (long long)d->H * d->W * d->Cin;
  }
  for (int i = 1; i < 4; ++i) a.mapA[i] = a.mapA[0];
  a.block_n = pick_block_n(d->Cin);
  a.tiles_n = (d->Cin + a.block_n - 1) / a.block_n;
  if ((rc = encode_matrix(&a.mapB, w_hwio, (long long)d->kh * d->kw * d->Cin, d->Cout, a.block_n)))
    return rc;
  a.taps = nt;
  a.k_chunks = (d->Cout + 63) / 64;
  {
    int rem = d->Cout - (a.k_chunks - 1) * 64;
    a.ksteps_last = (rem + 15) / 16;
  }
  a.e.n_total = d->Cin;
  a.e.out = static_cast<uint8_t*>(dx) + ((size_t)ph * d->W + pw) * d->Cin * esz;
  a.e.out_f32 = (dx_dtype == MCN_F32);
  a.e.bias = nullptr;
  a.e.vec_ok = (d->Cin % 16 == 0) && (reinterpret_cast<uintptr_t>(a.e.out) % 32 == 0);
  a.e.accumulate = accumulate;
  for (int t = 0; t < nt; ++t) {
    a.tab.brow[t] = tap_id[t] * d->Cin;
    a.tab.map[t] = 0;
    if (a.g.a_mode == 0 && !pointwise) {
      a.tab.dh[t] = (short)e_h[t];
      a.tab.dw[t] = (short)e_w[t];
    } else {
      a.tab.dh[t] = (short)(e_h[t] - min_eh);
      a.tab.dw[t] = (short)(e_w[t] - min_ew);
    }
  }
  return launch_gemm_conv(a, tiles_m_of(a.g), pointwise || (a.g.a_mode == 1 && unit), st, red);
}

extern "C" int mcn_conv2d_dgrad_tc(const mcn_conv_desc* d, const void* dy, const void* w_hwio,
                                   void* dx, int dx_dtype, int a_mode, int accumulate,
                                   void* stream) {
  MCN_REQUIRE(d && dy && w_hwio && dx, "dgrad_tc: null argument");
  MCN_REQUIRE(d->Cout % 8 == 0, "dgrad_tc: Cout=%d must be a multiple of 8", d->Cout);
  MCN_REQUIRE(d->kh * d->kw <= kMaxTaps, "dgrad_tc: too many taps");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t esz = (dx_dtype == MCN_F32) ? 4 : 2;
  if (a_mode == 2) {
    // stride-1 dgrad = correlation of dy with the mirrored taps; halo origin pad - (k-1)*d
    if (halo_eligible(d->Cout, d->kh, d->kw, d->sh, d->sw, d->dh, d->dw, d->H, d->W))
      return launch_halo(dy, d->Cout, d->Wo, d->Ho, d->N, w_hwio, d->Cin, d->Cin, d->kh, d->kw, d->dh,
                         d->dw, d->pad_t - (d->kh - 1) * d->dh, d->pad_l - (d->kw - 1) * d->dw, true,
                         d->H, d->W, dx, dx_dtype == MCN_F32, nullptr, accumulate, nullptr, st);
    a_mode = 1;
  }
  // phases without a contributing tap (e.g. 1x1 stride 2) stay zero
  bool any_empty = false;
  for (int ph = 0; ph < d->sh && !any_empty; ++ph)
    for (int pw = 0; pw < d->sw && !any_empty; ++pw) {
      bool hit = false;
      for (int r = 0; r < d->kh && !hit; ++r)
        for (int s = 0; s < d->kw && !hit; ++s)
          hit = (((ph + d->pad_t - r * d->dh) % d->sh) == 0) &&
                (((pw + d->pad_l - s * d->dw) % d->sw) == 0);
      any_empty = !hit;
    }
  if (any_empty && !accumulate) {
    if (cudaMemsetAsync(dx, 0, (size_t)d->N * d->H * d->W * d->Cin * esz, st) != cudaSuccess) {
      set_error("dgrad_tc: memset failed");
      return MCN_ECUDA;
    }
  }
  for (int ph = 0; ph < d->sh; ++ph)
    for (int pw = 0; pw < d->sw; ++pw) {
      bool empty = false;
      int rc = dgrad_phase(d, dy, w_hwio, dx, dx_dtype, a_mode, accumulate, ph, pw, &empty, st);
      if (rc) return rc;
    }
  return MCN_OK;
}

// dgrad (stride 1, bf16) whose epilogue also takes the batch-norm BACKWARD sums of the layer that
// produced the conv's input: dx is d(loss)/d(BN output); with x = the BN input,
// sums[c] += sum dz, sums[C + c] += sum dz*x, dz = dx * act'(x*gamma*invstd + beta - mean*gamma*invstd).
extern "C" int mcn_conv2d_dgrad_bnred_supported(const mcn_conv_desc* d, int a_mode) {
  if (d == nullptr || d->sh != 1 || d->sw != 1 || d->Cin % 64 != 0 || d->Cout % 8 != 0) return 0;
  if (d->kh * d->kw > kMaxTaps || (long long)d->N * d->H * d->W >= (1LL << 31)) return 0;
  const char* e = getenv("MCN_TMA_STORE");
  if (e && e[0] == '0') return 0;
  const bool pointwise = d->kh == 1 && d->kw == 1;
  if (a_mode == 2 && halo_eligible(d->Cout, d->kh, d->kw, d->sh, d->sw, d->dh, d->dw, d->H, d->W)) return 1;
  if (pointwise) return 1;
  return (a_mode >= 1 && d->Cout % 64 == 0) ? 1 : 0;
}

extern "C" int mcn_conv2d_dgrad_tc_bnred(const mcn_conv_desc* d, const void* dy, const void* w_hwio,
                                         void* dx, int a_mode, const void* bn_x, const float* mean,
                                         const float* invstd, const float* gamma, const float* beta,
                                         int act, double* sums, float* sum_dz, float* sum_dz_xhat,
                                         void* stream) {
  MCN_REQUIRE(d && dy && w_hwio && dx && bn_x && mean && invstd && sums, "dgrad_tc_bnred: null argument");
  MCN_REQUIRE((sum_dz == nullptr) == (sum_dz_xhat == nullptr), "dgrad_tc_bnred: sum_dz and sum_dz_xhat go together");
  MCN_REQUIRE(act == MCN_ACT_NONE || act == MCN_ACT_RELU, "dgrad_tc_bnred: activation %d not supported", act);
  MCN_REQUIRE(mcn_conv2d_dgrad_bnred_supported(d, a_mode), "dgrad_tc_bnred: geometry not supported");
  MCN_REQUIRE(reinterpret_cast<uintptr_t>(dx) % 16 == 0 && reinterpret_cast<uintptr_t>(bn_x) % 16 == 0,
              "dgrad_tc_bnred: dx and bn_x must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const BnRed red{bn_x, mean, invstd, gamma, beta, act, sums, sum_dz, sum_dz_xhat};
  if (a_mode == 2) {
    if (halo_eligible(d->Cout, d->kh, d->kw, d->sh, d->sw, d->dh, d->dw, d->H, d->W))
      return launch_halo(dy, d->Cout, d->Wo, d->Ho, d->N, w_hwio, d->Cin, d->Cin, d->kh, d->kw, d->dh,
                         d->dw, d->pad_t - (d->kh - 1) * d->dh, d->pad_l - (d->kw - 1) * d->dw, true,
                         d->H, d->W, dx, 0, nullptr, 0, nullptr, st, &red);
    a_mode = 1;
  }
  bool empty = false;
  return dgrad_phase(d, dy, w_hwio, dx, MCN_BF16, a_mode, 0, 0, 0, &empty, st, &red);
}

// Halo wgrad: returns 1 when the geometry is eligible and the launch was issued, 0 when the
// caller should use the tap-at-a-time kernel, negative on error.
static int try_wgrad_halo(const mcn_conv_desc* d, const void* x, const void* dy, float* dw,
                          cudaStream_t st, long long* query) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("MCN_WGRAD_HALO");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled) return 0;
  if (d->sh != 1 || d->sw != 1 || (d->kh == 1 && d->kw == 1)) return 0;
  if (d->Cin % 64 != 0 || d->Cout % 64 != 0) return 0;
  if (d->kh > kWhMaxGroups || d->kw > kWhMaxAcc) return 0;
  const int taps = d->kh * d->kw;
  WgradHaloArgs a;
  std::memset(&a, 0, sizeof(a));
  a.hwb = 8 + (d->kw - 1) * d->dw;
  const bool pair = d->Cin == 64 && taps >= 2 && ((taps + 1) / 2) <= kWhMaxAcc &&
                    ((taps + 1) / 2) * 64 <= 512;
  int box_h;
  if (pair) {
    a.groups = 1;
    box_h = 16 + (d->kh - 1) * d->dh;
    a.block_n = 64;
    a.n_acc = (taps + 1) / 2;
    a.a_atoms = 1;
    a.ci_tile = 64;
  } else {
    if (d->Cin < 128) return 0;
    a.groups = d->kh;
    box_h = 16;
    a.block_n = (d->kw * 128 <= 512 && d->Cout % 128 == 0) ? 128 : 64;
    if (d->kw * a.block_n > 512) return 0;
    a.n_acc = d->kw;
    a.a_atoms = 2;
    a.ci_tile = 128;
  }
  if (a.hwb > 256 || box_h > 256) return 0;
  a.a_box_bytes = a.hwb * box_h * 128;
  if (a.a_box_bytes > 48 * 1024) return 0;
  a.a_stride = (a.a_box_bytes + 1023) / 1024 * 1024;
  a.tiles_w = (d->Wo + 7) / 8;
  a.tiles_h = (d->Ho + 15) / 16;
  const double eff = (double)(d->Ho * d->Wo) / ((double)a.tiles_h * 16 * a.tiles_w * 8);
  if (eff < 0.7) return 0;
  a.nb_atoms = a.block_n / 64;
  const uint32_t stage_bytes = a.a_atoms * a.a_stride + a.nb_atoms * kABytes;
  a.stages = std::min(4, (int)((216 * 1024) / stage_bytes));
  if (a.stages < 2) return 0;
  int cols = a.n_acc * a.block_n;
  a.tmem_cols = 32;
  while (a.tmem_cols < cols) a.tmem_cols *= 2;
  a.tiles_total = d->N * a.tiles_w * a.tiles_h;
  a.units_ci = (d->Cin + a.ci_tile - 1) / a.ci_tile;
  a.units_co = (d->Cout + a.block_n - 1) / a.block_n;
  const int units = a.units_ci * a.units_co * a.groups;
  a.splits = std::max(1, std::min(a.tiles_total, num_sms() / units));
  a.cin = d->Cin;
  a.cout = d->Cout;
  const long long dw_elems = (long long)taps * d->Cin * d->Cout;
  SplitPlan sp;
  {
    const int prc = plan_splits(dw_elems, a.splits, &sp, query, "wgrad_tc (halo)");
    if (prc) return prc;
    if (query != nullptr) return 1;
  }
  a.splits = sp.splits;
  a.ws_stride = sp.stride;
  a.dw = sp.stride ? sp.base : dw;
  a.org_w = -d->pad_l;
  auto off = [&](int r, int s2) { return (r * d->dh * a.hwb + s2 * d->dw) * 128; };
  if (pair) {
    a.org_h[0] = -d->pad_t;
    for (int i = 0; i < a.n_acc; ++i) {
      int lo = 2 * i, hi = 2 * i + 1;
      bool discard_lo = false;
      if (hi >= taps) {   // odd tap count: the last tap rides in the upper half next to its predecessor
        hi = taps - 1;
        lo = taps - 2;
        discard_lo = true;
      }
      a.aoff[0][i] = off(lo / d->kw, lo % d->kw);
      a.lbo[0][i] = off(hi / d->kw, hi % d->kw) - a.aoff[0][i];
      a.tap_lo[0][i] = (short)(discard_lo ? -1 : lo);
      a.tap_hi[0][i] = (short)hi;
      if (a.lbo[0][i] <= 0) return 0;
    }
  } else {
    for (int r = 0; r < d->kh; ++r) {
      a.org_h[r] = -d->pad_t + r * d->dh;
      for (int s2 = 0; s2 < d->kw; ++s2) {
        a.aoff[r][s2] = s2 * d->dw * 128;
        a.lbo[r][s2] = a.a_stride;
        a.tap_lo[r][s2] = (short)(r * d->kw + s2);
        a.tap_hi[r][s2] = -2;
      }
    }
  }
  int rc;
  if ((rc = encode_nhwc(&a.mapX, x, d->Cin, d->W, d->H, d->N, a.hwb, box_h, 1))) return rc;
  if ((rc = encode_nhwc(&a.mapDy, dy, d->Cout, d->Wo, d->Ho, d->N, 8, 16, 1))) return rc;
  size_t smem = (size_t)a.stages * stage_bytes + kBarRegionBytes + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_optin_limit()) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(wgrad_halo_kernel) failed");
      return MCN_ECUDA;
    }
    configured = true;
  }
  dim3 grid((unsigned)(units * a.splits));
  {
    static int tma_out_enabled = -1;
    if (tma_out_enabled < 0) {
      const char* e = getenv("MCN_WGRAD_TMA_OUT");
      tma_out_enabled = (e && e[0] == '0') ? 0 : 1;
    }
    a.tma_out = 0;
    if (tma_out_enabled && sp.stride && d->Cout % 32 == 0 && (size_t)a.stages * stage_bytes >= 8 * 4096) {
      if ((rc = encode_wgrad_slices(&a.mapOut, sp.base, d->Cout, d->Cin, taps, a.splits, sp.stride))) return rc;
      a.tma_out = 1;
    }
  }
  ::mcn::launch(wgrad_halo_kernel, grid, 192, smem, st, a);
  rc = after_launch("wgrad_halo_kernel");
  if (rc) return rc;
  if (sp.stride && (rc = launch_splitk_reduce(sp.base, sp.stride, a.splits, dw_elems, dw, st))) return rc;
  return 1;
}

static int wgrad_tc_impl(const mcn_conv_desc* d, const void* x, const void* dy, float* dw,
                         int a_mode, void* stream, long long* query) {
  MCN_REQUIRE(d && (query || (x && dy && dw)), "wgrad_tc: null argument");
  MCN_REQUIRE(d->Cin % 8 == 0 && d->Cout % 8 == 0, "wgrad_tc: channels must be multiples of 8");
  if (a_mode == 2) {
    int rc = try_wgrad_halo(d, x, dy, dw, static_cast<cudaStream_t>(stream), query);
    if (rc != 0) return rc < 0 ? rc : MCN_OK;
    a_mode = 1;   // not eligible: tap-at-a-time kernel with the im2col feed
  }
  MCN_REQUIRE(d->kh * d->kw <= kMaxTaps, "wgrad_tc: too many taps");
  const bool pointwise = d->kh == 1 && d->kw == 1 && d->sh == 1 && d->sw == 1;
  if (a_mode == 1 && (d->Cin % 64 != 0 || pointwise)) a_mode = 0;
  MCN_REQUIRE(a_mode == 1 || (d->sh == 1 && d->sw == 1) || (d->kh == 1 && d->kw == 1),
              "wgrad_tc: box mode supports stride 1 (or 1x1 kernels) only");
  WgradArgs a;
  std::memset(&a, 0, sizeof(a));
  int rc;
  const int taps = d->kh * d->kw;
  a.block_n = pick_block_n(d->Cout);
  if (a.block_n > 128) a.block_n = 128;  // keep a stage at 64 KB
  a.a_atoms = 2;
  a.atom_bytes = kABytes;
  // Wide tiles for the 1x1 convolutions.  Per-role counters put the MMA thread of these launches in
  // "waiting for operands" with the L2 -> shared-memory traffic at the chip's ~6300 B/cycle cap: a
  // 128 x 128 output tile moves 512 operand bytes per 128 x 128 x 1 MACs.  Two M = 128 accumulators
  // (256 input channels) against a 256-wide dy tile halve that; the pixel box shrinks to 64 rows so
  // three stages of (4 + 4) x 8 KB still fit.  MCN_WGRAD_WIDE=0 restores the narrow tiles (A/B).
  static int wide_enabled = -1;
  if (wide_enabled < 0) {
    const char* e = getenv("MCN_WGRAD_WIDE");
    wide_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  const bool wide = wide_enabled && pointwise && (d->Cin % 256 == 0 || d->Cout % 256 == 0) &&
                    (long long)d->N * d->H * d->W >= 4096;
  if (wide) {
    if (d->Cin % 256 == 0) a.a_atoms = 4;
    if (d->Cout % 256 == 0) a.block_n = 256;
    a.atom_bytes = 64 * 128;
  }
  a.nb_atoms = a.block_n / 64;
  if (pointwise) {
    PixelSpace ps{static_cast<int>((long long)d->N * d->H * d->W), 1, 1};
    fill_geom_tiled(&a.g, ps);
    if (wide) {   // 64-pixel boxes
      a.g.TW = std::min(ps.W, 64);
      a.g.rows_box = a.g.TW;
      a.g.tiles_w = (ps.W + a.g.TW - 1) / a.g.TW;
    }
    uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)ps.W, 1, 1};
    uint64_t stb[4] = {2, (uint64_t)d->Cin * 2, (uint64_t)ps.W * d->Cin * 2,
                       (uint64_t)ps.W * d->Cin * 2};
    uint32_t box[4] = {64, (uint32_t)a.g.TW, 1, 1};
    if (!query && (rc = encode_tiled(&a.mapX[0], x, 4, dims, stb, box))) return rc;
    uint64_t dimsy[4] = {(uint64_t)d->Cout, (uint64_t)ps.W, 1, 1};
    uint64_t sty[4] = {2, (uint64_t)d->Cout * 2, (uint64_t)ps.W * d->Cout * 2,
                       (uint64_t)ps.W * d->Cout * 2};
    if (!query && (rc = encode_tiled(&a.mapDy, dy, 4, dimsy, sty, box))) return rc;
  } else if (a_mode == 0) {
    PixelSpace ps{d->Wo, d->Ho, d->N};
    fill_geom_tiled(&a.g, ps);
    if (d->sh == 1 && d->sw == 1) {
      if (!query && (rc = encode_nhwc(&a.mapX[0], x, d->Cin, d->W, d->H, d->N, a.g.TW, a.g.TH, a.g.TN)))
        return rc;
    } else {
      uint64_t dims[4] = {(uint64_t)d->Cin, (uint64_t)d->Wo, (uint64_t)d->Ho, (uint64_t)d->N};
      uint64_t stb[4] = {2, (uint64_t)d->sw * d->Cin * 2, (uint64_t)d->sh * d->W * d->Cin * 2,
                         (uint64_t)d->H * d->W * d->Cin * 2};
      uint32_t box[4] = {64, (uint32_t)a.g.TW, (uint32_t)a.g.TH, (uint32_t)a.g.TN};
      if (!query && (rc = encode_tiled(&a.mapX[0], x, 4, dims, stb, box))) return rc;
    }
    if (!query && (rc = encode_nhwc(&a.mapDy, dy, d->Cout, d->Wo, d->Ho, d->N, a.g.TW, a.g.TH, a.g.TN)))
      return rc;
  } else {
    a.g.a_mode = 1;
    a.g.Wo = d->Wo;
    a.g.Ho = d->Ho;
    a.g.Nb = d->N;
    a.g.m_total = (long long)d->N * d->Ho * d->Wo;
    a.g.str_w = d->sw;
    a.g.str_h = d->sh;
    a.g.low_w = -d->pad_l;
    a.g.low_h = -d->pad_t;
    const int up_w = (d->Wo - 1) * d->sw - d->pad_l - (d->W - 1);
    const int up_h = (d->Ho - 1) * d->sh - d->pad_t - (d->H - 1);
    if (!query && (rc = encode_im2col(&a.mapX[0], x, d->Cin, d->W, d->H, d->N, a.g.low_w, a.g.low_h, up_w,
                            up_h, d->sw, d->sh)))
      return rc;
    uint64_t dimsy[4] = {(uint64_t)d->Cout, (uint64_t)a.g.m_total, 1, 1};
    uint64_t sty[4] = {2, (uint64_t)d->Cout * 2, (uint64_t)a.g.m_total * d->Cout * 2,
                       (uint64_t)a.g.m_total * d->Cout * 2};
    uint32_t box[4] = {64, 128, 1, 1};
    if (!query && (rc = encode_tiled(&a.mapDy, dy, 4, dimsy, sty, box))) return rc;
  }
  for (int i = 1; i < 4; ++i) a.mapX[i] = a.mapX[0];
  a.taps = taps;
  a.cin = d->Cin;
  a.cout = d->Cout;
  a.tiles_mi = (d->Cin + 64 * a.a_atoms - 1) / (64 * a.a_atoms);
  a.tiles_ni = (d->Cout + a.block_n - 1) / a.block_n;
  a.kblocks_total = tiles_m_of(a.g);
  a.ksteps = (a.g.a_mode == 0) ? (a.g.rows_box + 15) / 16 : 8;
  {
    const int base = taps * a.tiles_mi * a.tiles_ni;
    // two full waves of one-CTA-per-SM at most: rounding the split count UP (304 CTAs for a base of
    // 16) left a third, nearly empty wave behind.  Wide tiles: one wave (every extra split is another
    // 256 KB partial tile to write and sum)
    static int wave_pct = -1;      // MCN_WGRAD_WAVE_PCT: fraction of the SMs a wave of splits may fill (A/B)
    if (wave_pct < 0) {
      const char* e = getenv("MCN_WGRAD_WAVE_PCT");
      wave_pct = e ? std::max(10, std::min(100, atoi(e))) : 100;
    }
    int want = std::max(1, ((wide ? 1 : 2) * num_sms() * wave_pct / 100) / base);
    // one CTA per SM is resident at a time, so a grid that already fills 3/4 of the SMs gains
    // nothing from splitting K — and an un-split K needs no slices and no second pass
    if (4 * base >= 3 * num_sms()) want = 1;
    a.splits = std::max(1, std::min(want, a.kblocks_total));
  }
  const long long dw_elems = (long long)taps * d->Cin * d->Cout;
  SplitPlan sp;
  sp.base = nullptr;
  sp.stride = 0;
  sp.splits = 1;
  if (a.splits == 1) {
    if (query != nullptr) {
      *query = kWsMinBytes;
      return MCN_OK;
    }
    a.ws_stride = -1;
    a.dw = dw;
  } else {
    if ((rc = plan_splits(dw_elems, a.splits, &sp, query, "wgrad_tc"))) return rc;
    if (query != nullptr) return MCN_OK;
    a.splits = sp.splits;
    a.ws_stride = sp.stride;
    a.dw = sp.stride ? sp.base : dw;
  }
  for (int r = 0; r < d->kh; ++r)
    for (int s = 0; s < d->kw; ++s) {
      int t = r * d->kw + s;
      a.tab.brow[t] = 0;
      a.tab.map[t] = 0;
      if (a.g.a_mode == 0 && !pointwise) {
        a.tab.dh[t] = (short)(r * d->dh - d->pad_t);
        a.tab.dw[t] = (short)(s * d->dw - d->pad_l);
      } else {
        a.tab.dh[t] = (short)(r * d->dh);
        a.tab.dw[t] = (short)(s * d->dw);
      }
    }
  const uint32_t stage_bytes = (a.a_atoms + a.nb_atoms) * a.atom_bytes;
  a.stages = std::max(2, std::min(4, (int)((200 * 1024) / stage_bytes)));
  a.tmem_cols = tmem_cols_for(a.block_n) * (a.a_atoms / 2);   // 64 .. 512, a power of two
  size_t smem = (size_t)a.stages * stage_bytes + (2 * a.stages + 1) * 8 + 16 + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_optin_limit()) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(wgrad_kernel) failed");
      return MCN_ECUDA;
    }
    configured = true;
  }
  dim3 grid((unsigned)(taps * a.tiles_mi * a.tiles_ni * a.splits));
  // in-kernel reduction: needs every CTA resident at once (the tiles' CTAs wait for each other), whole
  // 16-byte vectors and full tiles along Cout
  static int fuse_enabled = -1;
  if (fuse_enabled < 0) {
    // Off by default — measured SLOWER (wgrad class 4.10 -> 4.84 ms per step, +20 us on every 1x1
    // layer): 128 epilogue threads per SM keep ~4 KB of loads in flight where the stand-alone
    // splitk_reduce grid (16 x 148 blocks) keeps the whole memory system busy.  MCN_WGRAD_FUSE_REDUCE=1
    // selects it (tests / A-B).
    const char* e = getenv("MCN_WGRAD_FUSE_REDUCE");
    fuse_enabled = (e && e[0] == '1') ? 1 : 0;
  }
  a.fuse_reduce = 0;
  if (fuse_enabled && sp.stride && (int)grid.x <= num_sms() && d->Cout % a.block_n == 0 &&
      taps * a.tiles_mi * a.tiles_ni <= kWsCounters) {
    const Workspace w = current_workspace();
    a.fuse_reduce = 1;
    a.dw_final = dw;
    a.tickets = reinterpret_cast<unsigned int*>(w.base + kWsCounterOff);
  }
  // partial tiles through TMA stores: split-K slices only, 16-byte aligned rows, staging fits the ring
  static int tma_out_enabled = -1;
  if (tma_out_enabled < 0) {
    const char* e = getenv("MCN_WGRAD_TMA_OUT");      // 0: one 128-byte row per thread (A/B)
    tma_out_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  a.tma_out = 0;
  if (tma_out_enabled && sp.stride && !a.fuse_reduce && d->Cout % 4 == 0 && d->Cout % 32 == 0 &&
      (size_t)a.stages * stage_bytes >= 8 * 4096) {
    if ((rc = encode_wgrad_slices(&a.mapOut, sp.base, d->Cout, d->Cin, taps, a.splits, sp.stride))) return rc;
    a.tma_out = 1;
  }
  ::mcn::launch(wgrad_kernel, grid, 192, smem, static_cast<cudaStream_t>(stream), a);
  if ((rc = after_launch("wgrad_kernel"))) return rc;
  if (sp.stride && !a.fuse_reduce)
    return launch_splitk_reduce(sp.base, sp.stride, a.splits, dw_elems, dw, static_cast<cudaStream_t>(stream));
  return MCN_OK;
}

extern "C" int mcn_conv2d_wgrad_tc(const mcn_conv_desc* d, const void* x, const void* dy,
                                   float* dw, int a_mode, void* stream) {
  return wgrad_tc_impl(d, x, dy, dw, a_mode, stream, nullptr);
}


// ------------------------------------------------------------------ stem convolution: host side
namespace {
// Fills the geometry fields of StemArgs; returns false when the convolution is not a "stem" the
// gather kernels handle (see the comment above StemArgs).
bool stem_geometry(const mcn_conv_desc* d, StemArgs* a) {
  if (d->Cin != 4 || d->sw != 2 || d->dw != 1 || d->dh != 1 || (d->pad_l & 1) || (d->W & 1)) return false;
  if (!((d->kw == 3 && d->kh == 3) || (d->kw == 7 && d->kh == 7))) return false;   // instantiated filter sizes
  if (static_cast<long long>(d->N) * d->Ho * d->Wo >= (1LL << 31) - 256) return false;
  if (static_cast<long long>(d->H) * d->W * 4 >= (1LL << 31)) return false;
  if (d->Cout % 16 != 0 || d->Cout > 256) return false;
  a->H = d->H; a->W = d->W; a->Ho = d->Ho; a->Wo = d->Wo; a->N = d->N;
  a->sh = d->sh; a->pad_t = d->pad_t; a->pad_l = d->pad_l; a->kh = d->kh; a->kw = d->kw;
  a->epr = (d->kw + 1) * 4;
  a->cpr = a->epr / 8;
  a->rpc = 64 / a->epr;
  a->kc = (d->kh + a->rpc - 1) / a->rpc;
  a->kc_alloc = (a->kc + 1) / 2 * 2;
  a->kp = d->kh * a->epr;
  a->ksteps_last = (a->kp - (a->kc - 1) * 64 + 15) / 16;
  a->cout = d->Cout;
  a->m_total = static_cast<long long>(d->N) * d->Ho * d->Wo;
  a->total_tiles = static_cast<int>((a->m_total + 127) / 128);
  static int async_full = -1;
  if (async_full < 0) {
    const char* e = getenv("MCN_STEM_ASYNC_FULL");      // 0: commit / wait_group hand-off (A/B)
    async_full = (e && e[0] == '0') ? 0 : 1;
  }
  a->async_full = async_full;
  return true;
}
}  // namespace

extern "C" int mcn_stem_conv_kpad(const mcn_conv_desc* d) {
  StemArgs a;
  std::memset(&a, 0, sizeof(a));
  if (!d || !stem_geometry(d, &a)) return 0;
  return a.kc_alloc * 64;
}

extern "C" int mcn_stem_conv_fprop(const mcn_conv_desc* d, const void* x4, const void* w_okp,
                                   const float* bias, void* y, double* bn_sums, void* stream) {
  MCN_REQUIRE(d && x4 && w_okp && y, "stem_conv_fprop: null argument");
  StemArgs a;
  std::memset(&a, 0, sizeof(a));
  MCN_REQUIRE(stem_geometry(d, &a), "stem_conv_fprop: geometry not supported (needs Cin=4 padded input, "
              "stride-2 3x3/7x7, even W and left pad, Cout %% 16 == 0)");
  MCN_REQUIRE(bn_sums == nullptr || d->Cout % 64 == 0, "stem_conv_fprop: fused statistics need Cout %% 64 == 0");
  MCN_REQUIRE(reinterpret_cast<uintptr_t>(x4) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 32 == 0,
              "stem_conv_fprop: misaligned tensor");
  a.x4 = static_cast<const __nv_bfloat16*>(x4);
  const int kpad = a.kc_alloc * 64;
  int rc;
  if ((rc = encode_matrix(&a.mapW, w_okp, d->Cout, kpad, d->Cout))) return rc;
  a.e.block_n = d->Cout;
  a.e.n_total = d->Cout;
  a.e.out = y;
  a.e.bias = bias;
  a.e.stats = bn_sums;
  a.e.out_f32 = 0;
  a.e.accumulate = 0;
  a.e.vec_ok = (d->Cout % 16 == 0);
  a.nb_atoms = 0;
  const size_t a_stage = static_cast<size_t>(a.kc_alloc) * kABytes;
  // TMA-store epilogue (per-role counters: with the direct epilogue, one 128-byte row per thread, the
  // epilogue warps were busy 94 % of the kernel and the producers idle half of it)
  static int tma_enabled = -1;
  if (tma_enabled < 0) {
    const char* e = getenv("MCN_TMA_STORE");
    tma_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  a.use_tma = (tma_enabled && bias == nullptr && d->Cout % 64 == 0 && a.m_total < (1LL << 31) &&
               (static_cast<size_t>(a.kc) * d->Cout * 128) % 1024 == 0) ? 1 : 0;
  a.e.stg_bufs = 2;
  a.e.out_rank4 = 0;
  if (a.use_tma && (rc = encode_matrix(&a.mapOut, y, a.m_total, d->Cout, 32))) return rc;
  const size_t fixed = static_cast<size_t>(a.kc) * d->Cout * 128 + kBarRegionBytes +
                       (a.use_tma ? static_cast<size_t>(4 * a.e.stg_bufs) * kStgBytes + (bn_sums ? kStatAccBytes : 0)
                                  : (bn_sums ? kEpiStageBytes + kStatAccBytes : 0)) + 1024;
  a.stages = static_cast<int>(std::min<size_t>(3, (static_cast<size_t>(smem_optin_limit()) - fixed) / a_stage));
  MCN_REQUIRE(a.stages >= 2, "stem_conv_fprop: shared memory budget too small");
  a.tmem_cols = 2 * tmem_cols_for(d->Cout);
  const size_t smem = a.stages * a_stage + fixed;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(stem_fprop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_optin_limit()) != cudaSuccess ||
        cudaFuncSetAttribute(stem_fprop_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_optin_limit()) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(stem_fprop_kernel) failed");
      return MCN_ECUDA;
    }
    configured = true;
  }
  if ((rc = attach_xs(&a.e, a.total_tiles, 1))) return rc;
  dim3 grid(static_cast<unsigned>(std::min(a.total_tiles, num_sms())));
  if (a.use_tma) ::mcn::launch(stem_fprop_kernel<true>, grid, kStemThreads, smem, static_cast<cudaStream_t>(stream), a);
  else ::mcn::launch(stem_fprop_kernel<false>, grid, kStemThreads, smem, static_cast<cudaStream_t>(stream), a);
  return after_launch("stem_fprop_kernel");
}

static int stem_wgrad_impl(const mcn_conv_desc* d, const void* x4, const void* dy, float* dw,
                           void* stream, long long* query) {
  MCN_REQUIRE(d && (query || (x4 && dy && dw)), "stem_conv_wgrad: null argument");
  StemArgs a;
  std::memset(&a, 0, sizeof(a));
  MCN_REQUIRE(stem_geometry(d, &a) && d->Cout % 64 == 0,
              "stem_conv_wgrad: geometry not supported (see mcn_stem_conv_fprop; Cout %% 64 == 0)");
  a.x4 = static_cast<const __nv_bfloat16*>(x4);
  int rc;
  a.splits = std::max(1, std::min(a.total_tiles, num_sms()));
  const long long dw_elems = (long long)a.kc_alloc * 64 * d->Cout;
  SplitPlan sp;
  if ((rc = plan_splits(dw_elems, a.splits, &sp, query, "stem_conv_wgrad"))) return rc;
  if (query != nullptr) return MCN_OK;
  a.splits = sp.splits;
  a.ws_stride = sp.stride;
  a.dw = sp.stride ? sp.base : dw;
  if ((rc = encode_matrix(&a.mapW, dy, a.m_total, d->Cout, 128))) return rc;
  a.e.block_n = d->Cout;
  a.nb_atoms = d->Cout / 64;
  const size_t stage_bytes = static_cast<size_t>(a.kc_alloc + a.nb_atoms) * kABytes;
  a.stages = static_cast<int>(std::min<size_t>(3, (static_cast<size_t>(smem_optin_limit()) - kBarRegionBytes - 1024) / stage_bytes));
  MCN_REQUIRE(a.stages >= 2, "stem_conv_wgrad: shared memory budget too small");
  int cols = (a.kc_alloc / 2) * d->Cout;
  MCN_REQUIRE(cols <= 512, "stem_conv_wgrad: too many accumulator columns");
  a.tmem_cols = 32;
  while (a.tmem_cols < cols) a.tmem_cols *= 2;
  const size_t smem = a.stages * stage_bytes + kBarRegionBytes + 1024;
  static bool configured = false;
  if (!configured) {
    if (cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             smem_optin_limit()) != cudaSuccess) {
      set_error("cudaFuncSetAttribute(stem_wgrad_kernel) failed");
      return MCN_ECUDA;
    }
    configured = true;
  }
  ::mcn::launch(stem_wgrad_kernel, a.splits, kStemThreads, smem, static_cast<cudaStream_t>(stream), a);
  if ((rc = after_launch("stem_wgrad_kernel"))) return rc;
  if (sp.stride)
    return launch_splitk_reduce(sp.base, sp.stride, a.splits, dw_elems, dw, static_cast<cudaStream_t>(stream));
  return MCN_OK;
}

extern "C" int mcn_stem_conv_wgrad(const mcn_conv_desc* d, const void* x4, const void* dy, float* dw,
                                   void* stream) {
  return stem_wgrad_impl(d, x4, dy, dw, stream, nullptr);
}

// Bytes of workspace (mcn_set_workspace) the wgrad of this convolution would like: the fixed
// reduction area plus one dw-sized fp32 slice per split.  stem != 0: mcn_stem_conv_wgrad geometry.
// A smaller workspace still works (fewer splits) down to one slice.
extern "C" long long mcn_conv2d_wgrad_workspace_bytes(const mcn_conv_desc* d, int a_mode, int stem) {
  long long q = 0;
  const int rc = stem ? stem_wgrad_impl(d, nullptr, nullptr, nullptr, nullptr, &q)
                      : wgrad_tc_impl(d, nullptr, nullptr, nullptr, a_mode, nullptr, &q);
  return rc ? static_cast<long long>(rc) : q;
}

// Debug: per-role stall cycles accumulated by the -DMCN_ROLE_TIMING build (zeros otherwise).
extern "C" int mcn_debug_role_cycles(unsigned long long* out16, int reset) {
  MCN_REQUIRE(out16 != nullptr, "debug_role_cycles: null argument");
#ifdef MCN_ROLE_TIMING
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out16, g_role_cycles, sizeof(unsigned long long) * 16);
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_role_cycles, z, sizeof(z));
  }
  return MCN_OK;
#else
  for (int i = 0; i < 16; ++i) out16[i] = 0;
  (void)reset;
  return MCN_OK;
#endif
}
