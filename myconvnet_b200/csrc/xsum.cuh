// Order-independent (bit-reproducible) reductions.
//
// Every cross-block reduction of the training step used to end in floating-point atomics, whose
// summation order changes from run to run: two runs of the same step gave different low bits and,
// through ReLU masks, visibly different losses a few steps later.  Two tools replace them:
//
//  1. xs:: — an EXACT fixed-point accumulator.  A sum is held in three signed 64-bit limbs, each a
//     40-bit window of one long fixed-point number (bit weights 2^-80 .. 2^39, the top limb keeps
//     growing to 2^62).  An fp32 addend (24-bit mantissa) is split over at most two adjacent
//     limbs and added with integer atomics — integer addition is associative, so the result does
//     not depend on the order in which blocks arrive.  Limb k of element i of an n-element
//     accumulator array lives at limbs[k*n + i].  Addends below 2^-80 lose their low bits (by
//     truncation of the addend itself, still order-independent); |addend| >= 2^62, inf and NaN
//     poison the sum (it decodes as NaN).
//  2. A per-device workspace (mcn_set_workspace): the limbs of kernels whose consumers want
//     plain float/double sums live there; the LAST block of such a kernel (a ticket counter in the
//     workspace) decodes them, adds them to the caller's output and clears limbs and counter, so
//     the workspace is all-zero again when the kernel ends.  Launches that share a workspace must
//     be stream-ordered (as with a cuBLAS handle).  The split-K partial tiles of the wgrad
//     kernels use the rest of the workspace and are summed in split order by splitk_reduce.
#pragma once
#include <cstdint>

#include "mcn_common.cuh"

namespace mcn {

struct Workspace {
  unsigned char* base;
  long long bytes;
};
// Workspace registered for the current device (runtime.cu); base == nullptr when none is set.
Workspace current_workspace();

constexpr long long kWsCounterOff = 0;            // unsigned int ticket counters (one per channel group)
constexpr int kWsCounters = 1024;
constexpr long long kWsXsOff = 4096;              // limbs of the in-flight reduction
constexpr int kWsXsMax = 32768;                   // sums per launch
constexpr long long kWsSplitOff = 1 << 20;        // split-K partial tiles start here
constexpr long long kWsMinBytes = kWsSplitOff;

struct XsScratch {
  long long* limbs;
  unsigned int* counter;
};
// Host: the workspace's limb area for a launch that reduces n sums; limbs == nullptr (and the
// error text set) when no workspace is registered or n is too large.
XsScratch xs_scratch(int n, const char* who);

namespace xs {

constexpr int kLimbs = 3;
constexpr int kBits = 40;
constexpr int kE0 = -80;   // weight of bit 0 of limb 0

__device__ __forceinline__ void limb_add(long long* p, long long v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v));
}

__device__ __forceinline__ void add(long long* limbs, int n, int i, float v) {
  const uint32_t u = __float_as_uint(v);
  if ((u & 0x7fffffffu) == 0u) return;
  int ex = static_cast<int>((u >> 23) & 0xffu);
  long long* top = limbs + static_cast<size_t>(2) * n + i;
  if (ex == 0xff) {   // inf / NaN: poison
    limb_add(top, 1LL << 62);
    return;
  }
  unsigned long long m = u & 0x7fffffu;
  if (ex != 0) m |= 0x800000ull; else ex = 1;
  int pos = ex - 150 - kE0;   // v = m * 2^(ex-150): bit position of the mantissa's LSB
  if (pos < 0) {
    if (pos <= -24) return;
    m >>= -pos;
    pos = 0;
    if (m == 0) return;
  }
  const bool neg = (u >> 31) != 0u;
  if (pos >= 2 * kBits) {
    const int sh = pos - 2 * kBits;
    if (sh > 38) {
      limb_add(top, 1LL << 62);
      return;
    }
    const long long t = static_cast<long long>(m << sh);
    limb_add(top, neg ? -t : t);
    return;
  }
  const int k = pos >= kBits ? 1 : 0;
  const int r = pos - k * kBits;
  const unsigned long long wide = m << r;                       // < 2^(24+39)
  const long long lo = static_cast<long long>(wide & ((1ull << kBits) - 1ull));
  const long long hi = static_cast<long long>(wide >> kBits);
  long long* p = limbs + static_cast<size_t>(k) * n + i;
  if (lo) limb_add(p, neg ? -lo : lo);
  if (hi) limb_add(p + n, neg ? -hi : hi);
}

__device__ __forceinline__ double decode(long long l0, long long l1, long long l2) {
  const long long a2 = l2 < 0 ? -l2 : l2;
  if (a2 >= (1LL << 61)) return __longlong_as_double(0x7ff8000000000000LL);
  double v = static_cast<double>(l0) * 0x1p-80 + static_cast<double>(l1) * 0x1p-40;
  return v + static_cast<double>(l2);
}
// The limbs were written by other blocks' atomics (performed at L2): read around L1.
__device__ __forceinline__ double read(const long long* limbs, int n, int i) {
  return decode(__ldcg(limbs + i), __ldcg(limbs + static_cast<size_t>(n) + i),
                __ldcg(limbs + static_cast<size_t>(2) * n + i));
}
__device__ __forceinline__ double read_clear(long long* limbs, int n, int i) {
  const double v = read(limbs, n, i);
  __stcg(limbs + i, 0LL);
  __stcg(limbs + static_cast<size_t>(n) + i, 0LL);
  __stcg(limbs + static_cast<size_t>(2) * n + i, 0LL);
  return v;
}

// Ticket taken by one thread per block once ALL the block's adds have been issued (the caller
// synchronises the block before and broadcasts the result after).  True for the last of
// `blocks` arrivals; that caller decodes, clears the limbs and finally calls release().
__device__ __forceinline__ bool take_ticket(unsigned int* counter, unsigned int blocks) {
  __threadfence();
  const unsigned int t = atomicAdd(counter, 1u);
  __threadfence();
  return t == blocks - 1u;
}
__device__ __forceinline__ void release(unsigned int* counter) { __stcg(counter, 0u); }

// Whole-block helper: every thread has issued its adds; returns true (for all threads) in the
// last block of the grid.
__device__ __forceinline__ bool block_is_last(unsigned int* counter, unsigned int blocks) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0 && threadIdx.z == 0)
    s_last = take_ticket(counter, blocks) ? 1 : 0;
  __syncthreads();
  return s_last != 0;
}

}  // namespace xs

// dw[i] += sum_{s < splits} ws[s*stride + i], summed in split order (deterministic split-K).
int launch_splitk_reduce(const float* ws, long long stride, int splits, long long n, float* dw,
                         cudaStream_t st);

}  // namespace mcn
