// Shared pieces of the selection kernels (topk.cu, select_compact.cu): order-preserving keys, the block scan and the
// generic single-CTA radix select + bitonic sort.  Anonymous namespace: each translation unit gets its own copy.
#pragma once
#include "common.cuh"

namespace fkv {
namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxBins = 4096;
constexpr int kSmemSortMax = 16384;              // composites kept in shared memory (128 KB)
constexpr int kMaxCluster = 8;
constexpr int kCandMax = 4096;

__device__ inline uint32_t ordered_key(float x) {
  x += 0.0f;                                      // -0.0 -> +0.0 so both zeros compare equal
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// Inclusive block scan of one uint32 per thread (kThreads threads); returns inclusive value and
// the block total through `total`.  `wsum` is kWarps words of shared scratch.
__device__ inline uint32_t block_scan_incl(uint32_t v, uint32_t* wsum, uint32_t& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  __syncthreads();                                // scratch free from the previous use
  if (lane == 31) wsum[warp] = v;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    wsum[lane] = w;                               // inclusive warp totals
  }
  __syncthreads();
  total = wsum[kWarps - 1];
  return v + (warp ? wsum[warp - 1] : 0u);
}

// ---- generic single-CTA selection over a row in global memory ---------------------------------------
__device__ __noinline__ void select_row_generic(const float* __restrict__ x, int n, int k, int order, int tail,
                                   int64_t* __restrict__ out, uint32_t* hist, uint32_t* wsum, uint32_t* bc,
                                   unsigned long long* comp, int sort_cap) {
  const int tid = threadIdx.x;
  // ---- radix select: exact threshold key tau + number of tau-valued keys still needed -----------
  uint32_t prefix = 0, prefix_mask = 0;
  uint32_t need = (uint32_t)k;
  if (k < n) {
    const int shifts[3] = {20, 8, 0};
    const int bits[3] = {12, 12, 8};
    for (int pass = 0; pass < 3; ++pass) {
      const int shift = shifts[pass];
      const uint32_t nb = 1u << bits[pass];
      for (int i = tid; i < kMaxBins; i += kThreads) hist[i] = 0;
      __syncthreads();
      for (int j0 = 0; j0 < n; j0 += kThreads) {
        const int j = j0 + tid;
        const bool valid = j < n;
        const uint32_t key = valid ? ordered_key(x[j]) : 0u;
        const bool hit = valid && ((key & prefix_mask) == prefix);
        const uint32_t bin = (key >> shift) & (nb - 1);
        // warp-uniform fast path: identical bins collapse into one atomic (all-equal rows)
        const uint32_t hits = __ballot_sync(0xffffffffu, hit);
        if (hits) {
          const int leader = __ffs(hits) - 1;
          const uint32_t lbin = __shfl_sync(0xffffffffu, bin, leader);
          const uint32_t same = __ballot_sync(0xffffffffu, hit && bin == lbin);
          if (same == hits) {
            if ((tid & 31) == leader) atomicAdd(&hist[lbin], __popc(hits));
          } else if (hit) {
            atomicAdd(&hist[bin], 1u);
          }
        }
      }
      __syncthreads();
      // suffix scan from the top bin: thread t owns bins [top - 4t - 3, top - 4t] (top = nb-1)
      const int per = kMaxBins / kThreads;          // 4
      uint32_t local[4];
      uint32_t lsum = 0;
#pragma unroll
      for (int i = 0; i < per; ++i) {
        const int bin = (int)nb - 1 - (tid * per + i);
        local[i] = (bin >= 0) ? hist[bin] : 0u;
        lsum += local[i];
      }
      uint32_t total;
      const uint32_t incl = block_scan_incl(lsum, wsum, total);
      const uint32_t excl = incl - lsum;             // elements in bins above this thread's range
      if (excl < need && need <= incl) {
        uint32_t above = excl;
#pragma unroll
        for (int i = 0; i < per; ++i) {
          if (above < need && need <= above + local[i]) {
            bc[0] = (uint32_t)((int)nb - 1 - (tid * per + i));
            bc[1] = need - above;
          }
          above += local[i];
        }
      }
      __syncthreads();
      prefix |= bc[0] << shift;
      prefix_mask |= (nb - 1) << shift;
      need = bc[1];
      __syncthreads();
    }
  } else {
    prefix = 0; prefix_mask = 0xffffffffu; need = 0;
  }
  const uint32_t tau = prefix;
  const bool take_all = (k >= n);

  // ---- index-ordered compaction ---------------------------------------------------------------------
  const bool by_value = (order == FKV_ORDER_VALUE_DESC);
  uint32_t base_gt = 0, base_eq = 0;
  for (int j0 = 0; j0 < n; j0 += kThreads * 4) {
    const int j = j0 + tid * 4;
    uint32_t key[4];
    uint32_t gt = 0, eq = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool valid = (j + i) < n;
      key[i] = valid ? ordered_key(x[j + i]) : 0u;
      const bool g = valid && (take_all || key[i] > tau);
      const bool e = valid && !take_all && key[i] == tau;
      gt += g; eq += e;
    }
    uint32_t total;
    const uint32_t incl = block_scan_incl((eq << 16) | gt, wsum, total);
    uint32_t pre_gt = base_gt + ((incl & 0xffffu) - gt);
    uint32_t pre_eq = base_eq + ((incl >> 16) - eq);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const bool valid = (j + i) < n;
      const bool g = valid && (take_all || key[i] > tau);
      const bool e = valid && !take_all && key[i] == tau;
      if (g || (e && pre_eq < need)) {
        const uint32_t pos = pre_gt + min(pre_eq, need);
        if (by_value) comp[pos] = ((unsigned long long)key[i] << 32) | (uint32_t)(0xffffffffu - (uint32_t)(j + i));
        else out[pos] = (int64_t)(j + i);
      }
      pre_gt += g; pre_eq += e;
    }
    base_gt += total & 0xffffu;
    base_eq += total >> 16;
  }

  if (!by_value) {
    for (int i = tid; i < tail; i += kThreads) out[k + i] = (int64_t)(n + i);
    return;
  }

  // ---- bitonic sort, descending on (key, ~index) ------------------------------------------------------
  const int cap = sort_cap;
  for (int i = k + tid; i < cap; i += kThreads) comp[i] = 0ull;      // padding sorts last
  __syncthreads();
  for (int size = 2; size <= cap; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < cap / 2; t += kThreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long va = comp[lo], vb = comp[hi];
        if ((va < vb) == desc) { comp[lo] = vb; comp[hi] = va; }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < k; i += kThreads) out[i] = (int64_t)(0xffffffffu - (uint32_t)(comp[i] & 0xffffffffull));
  for (int i = tid; i < tail; i += kThreads) out[k + i] = (int64_t)(n + i);
}


}  // namespace
}  // namespace fkv
