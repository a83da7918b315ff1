// K3: deterministic top-k (value descending, ties -> lowest index) for the per-head KV selection
// and the global TSP selection.  Replaces Tensor.topk / sum+topk+cat+sort at
// baselines/fastkv/utils.py:113 and :127-130.
//
// One 1024-thread CTA per row:
//   phase 0  (nsum > 1) fixed-order sum of the source rows -> dense fp32 row (TSP importance)
//   phase 1  MSD radix select on order-preserving uint32 keys, digits 12 + 12 + 8 bits, each pass
//            filtering on the prefix found so far -> exact threshold key tau and the number of
//            tau-valued elements still needed
//   phase 2  index-ordered compaction (block scan): every key > tau plus the first `need` keys
//            == tau in index order  -> the selected set, ascending by index
//   phase 3  VALUE_DESC only: bitonic sort of (key, ~index) composites in shared memory
#include "common.cuh"

namespace fkv {
namespace {

constexpr int kThreads = 1024;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxBins = 4096;
constexpr int kSmemSortMax = 16384;              // composites kept in shared memory (128 KB)

__device__ inline uint32_t ordered_key(float x) {
  x += 0.0f;                                      // -0.0 -> +0.0 so both zeros compare equal
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct TopkArgs {
  const float* src; int64_t src_stride_row, src_stride_sum;
  int n, k, order, tail, nsum;
  int64_t* out; int64_t out_stride_row;
  float* dense;            // [rows][n] summed rows (workspace or imp_out) when nsum > 1
  unsigned long long* sort_ws;   // [rows][pow2(k)] global sort scratch when k > kSmemSortMax
  int sort_cap;            // pow2 >= k
};

// Inclusive block scan of one uint32 per thread (kThreads threads); returns inclusive value and
// the block total through `total`.  `wsum` is kWarps+1 words of shared scratch.
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

__global__ void __launch_bounds__(kThreads, 1)
topk_row_kernel(TopkArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw);                    // [4096]
  uint32_t* wsum = hist + kMaxBins;                                          // [32]
  uint32_t* bc = wsum + kWarps;                                              // [4] broadcast slots
  unsigned long long* sort_smem = reinterpret_cast<unsigned long long*>(bc + 4);

  const int row = blockIdx.x, tid = threadIdx.x;
  const int n = a.n, k = a.k;
  int64_t* out = a.out + (int64_t)row * a.out_stride_row;

  // ---- phase 0: fixed-order head sum (baselines/fastkv/utils.py:127) -----------------------------
  const float* x;
  if (a.nsum > 1) {
    float* dense = a.dense + (int64_t)row * n;
    const float* base = a.src + (int64_t)row * a.src_stride_row;
    for (int j = tid; j < n; j += kThreads) {
      float acc = base[j];
      for (int s = 1; s < a.nsum; ++s) acc += base[(int64_t)s * a.src_stride_sum + j];
      dense[j] = acc;
    }
    __syncthreads();
    x = dense;
  } else {
    x = a.src + (int64_t)row * a.src_stride_row;
  }

  // ---- phase 1: radix select ----------------------------------------------------------------------
  uint32_t prefix = 0, prefix_mask = 0;
  uint32_t need = (uint32_t)k;                     // elements still to take inside the current prefix
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
    prefix = 0; prefix_mask = 0xffffffffu; need = 0;   // k == n: everything is "greater" than tau = 0 key
  }
  const uint32_t tau = prefix;
  const bool take_all = (k >= n);

  // ---- phase 2: index-ordered compaction ----------------------------------------------------------
  const bool by_value = (a.order == FKV_ORDER_VALUE_DESC);
  unsigned long long* comp = nullptr;
  if (by_value) comp = (a.sort_cap <= kSmemSortMax) ? sort_smem : a.sort_ws + (int64_t)row * a.sort_cap;

  uint32_t base_gt = 0, base_eq = 0;                 // selected-so-far counts
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
    for (int i = tid; i < a.tail; i += kThreads) out[k + i] = (int64_t)(n + i);
    return;
  }

  // ---- phase 3: bitonic sort, descending on (key, ~index) -------------------------------------------
  const int cap = a.sort_cap;
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
  for (int i = tid; i < a.tail; i += kThreads) out[k + i] = (int64_t)(n + i);
}

inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace

size_t topk_workspace_bytes(int rows, int n, int k) {
  size_t b = align_up((size_t)rows * n * sizeof(float), 256);                 // dense summed rows
  const int cap = next_pow2(k);
  if (cap > kSmemSortMax) b += align_up((size_t)rows * cap * sizeof(unsigned long long), 256);
  return b + 256;
}

int launch_topk(const fkv_topk_params& p, cudaStream_t s) {
  TopkArgs a{};
  a.src = p.src; a.src_stride_row = p.src_stride_row; a.src_stride_sum = p.src_stride_sum;
  a.n = p.n; a.k = p.k; a.order = p.order; a.tail = p.tail; a.nsum = p.nsum < 1 ? 1 : p.nsum;
  a.out = p.out; a.out_stride_row = p.out_stride_row;
  a.sort_cap = next_pow2(p.k);
  char* ws = reinterpret_cast<char*>(p.workspace);
  size_t off = 0;
  if (a.nsum > 1) {
    if (p.imp_out) a.dense = p.imp_out;
    else {
      if (!ws) return FKV_EWORKSPACE;
      a.dense = reinterpret_cast<float*>(ws);
      off = align_up((size_t)p.rows * p.n * sizeof(float), 256);
    }
  }
  const bool by_value = p.order == FKV_ORDER_VALUE_DESC;
  size_t smem = (kMaxBins + kWarps + 4) * sizeof(uint32_t);
  if (by_value) {
    if (a.sort_cap <= kSmemSortMax) smem += (size_t)a.sort_cap * sizeof(unsigned long long);
    else {
      const size_t need = off + (size_t)p.rows * a.sort_cap * sizeof(unsigned long long);
      if (!ws || p.workspace_bytes < need) return FKV_EWORKSPACE;
      a.sort_ws = reinterpret_cast<unsigned long long*>(ws + off);
    }
  }
  static bool attr_set = false;
  if (!attr_set) {
    const int smem_max = (int)((kMaxBins + kWarps + 4) * sizeof(uint32_t) + (size_t)kSmemSortMax * sizeof(unsigned long long));
    FKV_CUDA_TRY(cudaFuncSetAttribute(topk_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    attr_set = true;
  }
  topk_row_kernel<<<p.rows, kThreads, smem, s>>>(a);
  FKV_LAUNCH_CHECK();
  return FKV_OK;
}

}  // namespace fkv
