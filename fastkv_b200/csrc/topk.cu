// K3: deterministic top-k (value descending, ties -> lowest index) for the per-head KV selection
// and the global TSP selection.  Replaces Tensor.topk / sum+topk+cat+sort at
// baselines/fastkv/utils.py:113 and :127-130.
//
// Two kernels:
//  * topk_cluster_kernel — the fast path (n <= 131072, k <= 4096).  One thread-block CLUSTER of
//    C = ceil(n / 16384) CTAs per row; each CTA keeps its slice of the row in registers (16 values
//    per thread), so the row is read from L2/HBM exactly once.
//      1. local 4096-bin histogram of the top 12 key bits (shared-memory atomics)
//      2. histograms are reduced across the cluster through distributed shared memory (each CTA
//         owns 4096/C bins), the owner of the bin holding the k-th largest value finds it
//      3. every key in or above that bin (k <= count <= 4096, else fall back) is sent to the
//         leader CTA's shared memory (DSMEM stores)
//      4. the leader ranks the candidates without a sort: counting-scatter by bin, then each
//         candidate counts the larger keys inside its own (small) bin -> final position
//      5. INDEX_ASC (TSP): the k-th composite key is broadcast and every CTA compacts its own
//         selected indices in index order (block scan + cluster prefix)
//  * topk_row_kernel — generic fallback (any n, any k, heavy ties): one CTA per row, MSD radix
//    select (12 + 12 + 8 bits) over global memory, index-ordered compaction, bitonic sort.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "topk_common.cuh"

namespace cg = cooperative_groups;

namespace fkv {
namespace {

// optional KV compaction fused behind the per-head selection (K3 + K4 in one launch)
struct FusedGather {
  const char* k; const char* v; char* ko; char* vo;
  int64_t k_sb, k_sh, k_sl, v_sb, v_sh, v_sl, ko_sb, ko_sh, ko_sl, vo_sb, vo_sh, vo_sl;   // bytes
  int L, W, h0, head_div, Hsel, row_chunks, enabled;
};

struct TopkArgs {
  const float* src; int64_t src_stride_row, src_stride_sum;
  int n, k, order, tail, nsum;
  int64_t* out; int64_t out_stride_row;
  float* dense;            // [rows][n] summed rows (workspace or imp_out) when nsum > 1
  unsigned long long* sort_ws;   // [rows][pow2(k)] global sort scratch when k > kSmemSortMax
  int sort_cap;            // pow2 >= k
  int cluster;             // CTAs per row (cluster path)
  int write_dense;         // cluster path: also store the summed row (importance output)
  unsigned long long* trace;   // FKV_TOPK_TRACE debugging aid
  FusedGather g;
};

__global__ void __launch_bounds__(kThreads, 1)
topk_row_kernel(TopkArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw);                    // [4096]
  uint32_t* wsum = hist + kMaxBins;                                          // [32]
  uint32_t* bc = wsum + kWarps;                                              // [4] broadcast slots
  unsigned long long* sort_smem = reinterpret_cast<unsigned long long*>(bc + 4);

  const int row = blockIdx.x, tid = threadIdx.x;
  const int n = a.n;
  int64_t* out = a.out + (int64_t)row * a.out_stride_row;

  // fixed-order head sum (baselines/fastkv/utils.py:127)
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
  unsigned long long* comp = nullptr;
  if (a.order == FKV_ORDER_VALUE_DESC)
    comp = (a.sort_cap <= kSmemSortMax) ? sort_smem : a.sort_ws + (int64_t)row * a.sort_cap;
  select_row_generic(x, n, a.k, a.order, a.tail, out, hist, wsum, bc, comp, a.sort_cap);
}

// ---- cluster fast path -------------------------------------------------------------------------------
constexpr int kSliceKeys = 16384;                 // keys per CTA, held in shared memory
constexpr int kKeysPadded = kSliceKeys + kSliceKeys / 32;   // one pad word per 32 keys: conflict-free blocked reads
constexpr int kTargetKeys = 4096;                 // preferred keys per CTA (more CTAs per row = shorter phases)
constexpr int kKmax = 2048;                       // largest k served by the cluster path

struct ClusterSmem {
  uint32_t keys[kKeysPadded];       // order-preserving keys of this CTA's slice
  uint32_t hist[kMaxBins];          // local 12-bit histogram; leader: adaptive-bin counts / cursors
  uint32_t ghist[kMaxBins];         // leader: cluster-wide 12-bit histogram (remote atomics from every CTA)
  unsigned long long cand[kCandMax];     // leader: gathered candidates (unordered)
  unsigned long long sortbuf[kCandMax];  // leader: candidates grouped by adaptive bin (also the generic path's sort buffer)
  uint16_t bbase[kMaxBins];         // leader: candidates in higher adaptive bins (<= 4096)
  uint32_t cnt_cta[kMaxCluster];    // INDEX_ASC: selected count per CTA
  uint32_t wsum[kWarps];
  uint32_t bc[16];                  // [0] bin*, [1] count above bin*, [2] count in bin*, [3] gather cursor,
                                    // [4..5] k-th composite (lo, hi), [6] top non-empty bin, [11] fallback flag
};

__device__ inline int key_slot(int i) { return i + (i >> 5); }

// bitonic sort of `cap` (power of two) composites, descending; all kThreads threads participate
__device__ inline void bitonic_desc(unsigned long long* buf, int cap) {
  const int tid = threadIdx.x;
  for (int size = 2; size <= cap; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < cap / 2; t += kThreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long va = buf[lo], vb = buf[hi];
        if ((va < vb) == desc) { buf[lo] = vb; buf[hi] = va; }
      }
      __syncthreads();
    }
  }
}

#define TK_TRACE(i) do { if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.trace[i] = t_; } } while (0)

// kByValue: per-head selection in score order (+ optional fused KV copy); false: the TSP row (index order, head sum).
// Two instances, so that neither carries the other's code and register pressure (64-register cap at 1024 threads).
template <bool kByValue>
__global__ void __launch_bounds__(kThreads, 1)
topk_cluster_kernel(TopkArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ClusterSmem& sm = *reinterpret_cast<ClusterSmem*>(smem_raw);
  TK_TRACE(0);
  cg::cluster_group cluster = cg::this_cluster();
  const int C = a.cluster;
  const bool multi = C > 1;                         // C == 1: plain CTA, no cluster traffic at all
  const int crank = multi ? (int)cluster.block_rank() : 0;
  const int row = blockIdx.x / C;
  const int tid = threadIdx.x, lane = tid & 31;
  const int n = a.n, k = a.k;
  int64_t* out = a.out + (int64_t)row * a.out_stride_row;
  auto csync = [&]() { if (multi) cluster.sync(); else __syncthreads(); };
  ClusterSmem& lead = multi ? *cluster.map_shared_rank(&sm, 0) : sm;   // the leader CTA's shared memory

  for (int i = tid; i < kMaxBins; i += kThreads) { sm.hist[i] = 0; sm.ghist[i] = 0; }
  if (tid < 16) sm.bc[tid] = 0;
  csync();                                          // the leader's ghist is zero before anybody adds to it
  asm volatile("griddepcontrol.wait;" ::: "memory");   // programmatic dependent launch: the producer of `src` has completed

  // ---- 1. load the slice (fixed-order head sum for TSP), keys -> smem ---------------------------------
  const int per_cta = ((n + C - 1) / C + 3) & ~3;
  const int base = crank * per_cta;
  const int cnt = max(0, min(n, base + per_cta) - base);       // keys of this CTA
  const float* src = a.src + (int64_t)row * a.src_stride_row;
  {
    const bool vec_ok = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)(a.src_stride_sum * 4)) & 15) == 0;
    float* dense = a.write_dense ? a.dense + (int64_t)row * n : nullptr;
    auto emit = [&](int i, float x0, float x1, float x2, float x3) {   // 4 consecutive keys -> smem (+ dense row)
      const float x[4] = {x0, x1, x2, x3};
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (i + e < cnt) {
          if (dense) dense[base + i + e] = x[e];
          sm.keys[key_slot(i + e)] = ordered_key(x[e]);
        }
    };
    auto load4 = [&](const float* p, int i, float4& r) {                // p: row base, keys [i, i+4) of the slice
      if (vec_ok && i + 3 < cnt) r = *reinterpret_cast<const float4*>(p + base + i);
      else {
        r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < cnt) r.x = p[base + i];
        if (i + 1 < cnt) r.y = p[base + i + 1];
        if (i + 2 < cnt) r.z = p[base + i + 2];
        if (i + 3 < cnt) r.w = p[base + i + 3];
      }
    };
    if (a.nsum == 1) {
      constexpr int U = 4;                          // float4 loads in flight per thread
      for (int i0 = tid * 4; i0 < cnt; i0 += kThreads * 4 * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) load4(src, i0 + u * kThreads * 4, v[u]);
#pragma unroll
        for (int u = 0; u < U; ++u) emit(i0 + u * kThreads * 4, v[u].x, v[u].y, v[u].z, v[u].w);
      }
    } else {
      for (int i = tid * 4; i < cnt; i += kThreads * 4) {
        float4 acc;
        load4(src, i, acc);
        for (int h0 = 1; h0 < a.nsum; h0 += 4) {    // fixed head order (utils.py:127), 4 loads in flight
          float4 t4[4];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (h0 + q < a.nsum) load4(src + (int64_t)(h0 + q) * a.src_stride_sum, i, t4[q]);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (h0 + q < a.nsum) { acc.x += t4[q].x; acc.y += t4[q].y; acc.z += t4[q].z; acc.w += t4[q].w; }
        }
        emit(i, acc.x, acc.y, acc.z, acc.w);
      }
    }
  }
  __syncthreads();
  // 12-bit histogram with run-length aggregation: neighbouring keys often share a bin, so a warp
  // issues one shared-memory atomic per run instead of one per key
  for (int i0 = 0; i0 < cnt; i0 += kThreads) {
    const int i = i0 + tid;
    const uint32_t bin = (i < cnt) ? (sm.keys[key_slot(i)] >> 20) : 0xffffffffu;
    const uint32_t prev = __shfl_up_sync(0xffffffffu, bin, 1);
    const bool head = (lane == 0) || (bin != prev);
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    if (head && i < cnt) {
      const uint32_t rest = (lane == 31) ? 0u : (heads >> (lane + 1));
      const int len = rest ? (__ffs(rest)) : (32 - lane);
      atomicAdd(&sm.hist[bin], (uint32_t)len);
    }
  }
  __syncthreads();
  TK_TRACE(1);

  // ---- 2. cluster histogram in the leader (only the few non-empty bins travel), leader finds bin* ------
  if (multi) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t c = sm.hist[tid * 4 + i];
      if (c) atomicAdd(&lead.ghist[tid * 4 + i], c);
    }
    cluster.sync();
  }
  TK_TRACE(2);
  if (crank == 0) {
    const uint32_t* gh = multi ? sm.ghist : sm.hist;
    uint32_t local[4], ls = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) { local[i] = gh[kMaxBins - 1 - (tid * 4 + i)]; ls += local[i]; }
    uint32_t total;
    const uint32_t incl = block_scan_incl(ls, sm.wsum, total);
    uint32_t above = incl - ls;
    uint32_t b_star = 0, n_ab = 0, n_in = 0;
    bool found = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (above < (uint32_t)k && (uint32_t)k <= above + local[i]) {
        found = true; b_star = (uint32_t)(kMaxBins - 1 - (tid * 4 + i)); n_ab = above; n_in = local[i];
      }
      if (above == 0 && local[i] != 0)                                // first non-empty bin from the top
        sm.bc[6] = (uint32_t)(kMaxBins - 1 - (tid * 4 + i));
      above += local[i];
    }
    if (found) {
      for (int c = 0; c < C; ++c) {
        uint32_t* rbc = multi ? cluster.map_shared_rank(sm.bc, c) : sm.bc;
        rbc[0] = b_star; rbc[1] = n_ab; rbc[2] = n_in;
      }
    }
  }
  csync();
  TK_TRACE(3);
  const uint32_t bin_star = sm.bc[0];
  uint32_t thr = bin_star << 20;                    // candidates are the keys >= thr
  uint32_t n_above = sm.bc[1];
  uint32_t n_cand = n_above + sm.bc[2];
  constexpr bool by_value = kByValue;

  if (n_cand > (uint32_t)kCandMax) {
    // ---- 2b. crowded threshold bin (narrow score distributions put thousands of keys within one
    // 1/8-octave bin): split bin* by the next 12 key bits, same cluster histogram once more --------
    for (int i = tid; i < kMaxBins; i += kThreads) { sm.hist[i] = 0; sm.ghist[i] = 0; }
    csync();                                        // the leader's ghist is zero before anybody adds to it
    for (int i = tid; i < cnt; i += kThreads) {
      const uint32_t key = sm.keys[key_slot(i)];
      if ((key >> 20) == bin_star) atomicAdd(&sm.hist[(key >> 8) & 0xfffu], 1u);
    }
    __syncthreads();
    if (multi) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t c = sm.hist[tid * 4 + i];
        if (c) atomicAdd(&lead.ghist[tid * 4 + i], c);
      }
      cluster.sync();
    }
    if (crank == 0) {
      const uint32_t* gh = multi ? sm.ghist : sm.hist;
      const uint32_t kk = (uint32_t)k - n_above;    // still needed from inside bin*
      uint32_t local[4], ls = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) { local[i] = gh[kMaxBins - 1 - (tid * 4 + i)]; ls += local[i]; }
      uint32_t total;
      const uint32_t incl = block_scan_incl(ls, sm.wsum, total);
      uint32_t above = incl - ls;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (above < kk && kk <= above + local[i]) {
          const uint32_t sub = (uint32_t)(kMaxBins - 1 - (tid * 4 + i));
          for (int c = 0; c < C; ++c) {
            uint32_t* rbc = multi ? cluster.map_shared_rank(sm.bc, c) : sm.bc;
            rbc[7] = (bin_star << 20) | (sub << 8); rbc[1] = n_above + above; rbc[2] = local[i];
          }
        }
        above += local[i];
      }
    }
    csync();
    thr = sm.bc[7];
    n_above = sm.bc[1];
    n_cand = n_above + sm.bc[2];
  }

  if (n_cand > (uint32_t)kCandMax) {
    // heavy ties / very flat rows: the leader re-selects from global memory with the generic path
    if (a.write_dense) __threadfence();
    csync();                                        // dense row (if any) fully written, remote smem no longer needed
    if (crank == 0) {
      const float* x = (a.nsum > 1) ? a.dense + (int64_t)row * n : src;
      select_row_generic(x, n, k, a.order, a.tail, out, sm.hist, sm.wsum, sm.bc, sm.sortbuf, a.sort_cap);
      if (a.g.enabled && by_value) {                // fused compaction on the fallback path: leader only
        __syncthreads();
        const FusedGather& g = a.g;
        const int b = row / g.Hsel, hs = row - b * g.Hsel;
        const int src_head = (g.h0 + hs) / g.head_div;
        const int lane_chunk = tid % g.row_chunks, rows_par = kThreads / g.row_chunks;
        for (int r = tid / g.row_chunks; r < k + g.W; r += rows_par) {
          const int64_t sr = (r < k) ? out[r] : (int64_t)(g.L - g.W + (r - k));
          st_stream16(g.ko + b * g.ko_sb + hs * g.ko_sh + (int64_t)r * g.ko_sl + lane_chunk * 16,
                      ld_stream16(g.k + b * g.k_sb + src_head * g.k_sh + sr * g.k_sl + lane_chunk * 16));
          st_stream16(g.vo + b * g.vo_sb + hs * g.vo_sh + (int64_t)r * g.vo_sl + lane_chunk * 16,
                      ld_stream16(g.v + b * g.v_sb + src_head * g.v_sh + sr * g.v_sl + lane_chunk * 16));
        }
      }
    }
    return;
  }

  // the leader's adaptive-bin counters (sm.hist is free from here on) are zeroed now, behind the candidate traffic
  const uint32_t klo = thr;
  const uint32_t khi = (sm.bc[6] << 20) | 0xfffffu;                     // upper bound of the top non-empty 12-bit bin
  // bin = (key - klo) >> bin_shift < 4096: a power-of-two map of the candidates' key range (2048..4096 bins in use)
  const int range_bits = 32 - __clz(khi - klo);
  const int bin_shift = range_bits > 12 ? range_bits - 12 : 0;
  if (crank == 0)
    for (int i = tid; i < kMaxBins; i += kThreads) sm.hist[i] = 0;

  // ---- 3. send candidates (bin >= bin*) to the leader: count, one cursor atomic per warp, write ------
  {
    uint32_t mine = 0;
    for (int i = tid; i < cnt; i += kThreads) mine += (sm.keys[key_slot(i)] >= thr);
    uint32_t pre = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += t;
    }
    const uint32_t warp_total = __shfl_sync(0xffffffffu, pre, 31);
    uint32_t wbase = 0;
    if (lane == 31 && warp_total) wbase = atomicAdd(&lead.bc[3], warp_total);
    wbase = __shfl_sync(0xffffffffu, wbase, 31);
    uint32_t pos = wbase + pre - mine;
    for (int i = tid; i < cnt; i += kThreads) {
      const uint32_t key = sm.keys[key_slot(i)];
      if (key >= thr)
        lead.cand[pos++] = ((unsigned long long)key << 32) | (uint32_t)(0xffffffffu - (uint32_t)(base + i));
    }
  }
  csync();
  TK_TRACE(4);

  // ---- 4. leader: rank the candidates without a general sort ------------------------------------------
  // Counting scatter on 4096 ADAPTIVE bins (the candidates' key range mapped linearly onto the bins,
  // so a bin holds ~1 key), then every candidate counts the larger keys in its own bin:
  // rank = keys in higher bins + larger keys in the bin.  A skewed row (a bin with more than 64
  // keys, e.g. exact ties) takes the bitonic sort instead.
  if (crank == 0) {
    uint32_t* cursor = sm.hist;
    uint16_t* bbase = sm.bbase;
    unsigned long long* grouped = sm.sortbuf;
    auto fine_bin = [&](uint32_t key) { return (key - klo) >> bin_shift; };
    TK_TRACE(8);
    for (int i = tid; i < (int)n_cand; i += kThreads) atomicAdd(&cursor[fine_bin((uint32_t)(sm.cand[i] >> 32))], 1u);
    __syncthreads();
    TK_TRACE(9);
    bool skew;
    {
      uint32_t local[4], ls = 0, mx = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) { local[i] = cursor[kMaxBins - 1 - (tid * 4 + i)]; ls += local[i]; mx = max(mx, local[i]); }
      skew = __syncthreads_or(mx > 64u);
      uint32_t total;
      const uint32_t inc = block_scan_incl(ls, sm.wsum, total);
      uint32_t above = inc - ls;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int b = kMaxBins - 1 - (tid * 4 + i);
        bbase[b] = (uint16_t)above;
        above += local[i];
        cursor[b] = 0;                              // becomes the scatter cursor (each thread owns its 4 bins here)
      }
    }
    __syncthreads();
    TK_TRACE(10);
    unsigned long long kth = 0ull;
    if (!skew) {
      for (int i = tid; i < (int)n_cand; i += kThreads) {
        const unsigned long long c = sm.cand[i];
        const uint32_t b = fine_bin((uint32_t)(c >> 32));
        grouped[bbase[b] + atomicAdd(&cursor[b], 1u)] = c;
      }
      __syncthreads();
      TK_TRACE(11);
      for (int i = tid; i < (int)n_cand; i += kThreads) {
        const unsigned long long c = grouped[i];
        const uint32_t b = fine_bin((uint32_t)(c >> 32));
        const uint32_t s0 = bbase[b], s1 = s0 + cursor[b];
        uint32_t rank = s0;
        for (uint32_t q = s0; q < s1; ++q) rank += (grouped[q] > c);
        if (rank < (uint32_t)k) {
          if (by_value) out[rank] = (int64_t)(0xffffffffu - (uint32_t)(c & 0xffffffffull));
          if (rank == (uint32_t)k - 1) { sm.bc[4] = (uint32_t)(c & 0xffffffffull); sm.bc[5] = (uint32_t)(c >> 32); }
        }
      }
      __syncthreads();
      kth = ((unsigned long long)sm.bc[5] << 32) | sm.bc[4];
    } else {
      int cap = 1;
      while (cap < (int)n_cand) cap <<= 1;
      for (int i = (int)n_cand + tid; i < cap; i += kThreads) sm.cand[i] = 0ull;
      __syncthreads();
      bitonic_desc(sm.cand, cap);
      if (by_value)
        for (int i = tid; i < k; i += kThreads) out[i] = (int64_t)(0xffffffffu - (uint32_t)(sm.cand[i] & 0xffffffffull));
      kth = sm.cand[k - 1];
    }
    if (by_value) {
      for (int i = tid; i < a.tail; i += kThreads) out[k + i] = (int64_t)(n + i);
    } else {
      __syncthreads();                              // every thread has read bc[4..5] before they are rewritten
      if (tid == 0) {
        for (int cc = 0; cc < C; ++cc) {
          uint32_t* rbc = multi ? cluster.map_shared_rank(sm.bc, cc) : sm.bc;
          rbc[4] = (uint32_t)(kth & 0xffffffffull);
          rbc[5] = (uint32_t)(kth >> 32);
        }
      }
    }
  }
  TK_TRACE(5);
  csync();
  TK_TRACE(7);
  if constexpr (by_value) {
    if (a.g.enabled) {
      // ---- 4b. fused KV compaction (baselines/fastkv/utils.py:114-121): the whole cluster copies the
      // selected rows of this (batch, head) + the W window rows; indices come from `out` (global,
      // visible cluster-wide after the barrier above)
      const FusedGather& g = a.g;
      const int b = row / g.Hsel, hs = row - b * g.Hsel;
      const int src_head = (g.h0 + hs) / g.head_div;
      const char* kb = g.k + b * g.k_sb + src_head * g.k_sh;
      const char* vb = g.v + b * g.v_sb + src_head * g.v_sh;
      char* kob = g.ko + b * g.ko_sb + hs * g.ko_sh;
      char* vob = g.vo + b * g.vo_sb + hs * g.vo_sh;
      const int rows_total = k + g.W;
      const int lane_chunk = tid % g.row_chunks, row_in_cta = tid / g.row_chunks;
      const int rows_par = kThreads / g.row_chunks;
      for (int rb = crank * rows_par + row_in_cta; rb < rows_total; rb += 4 * rows_par * C) {
        int64_t sr[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = rb + u * rows_par * C;
          sr[u] = (r < rows_total) ? ((r < k) ? __ldcg(out + r) : (int64_t)(g.L - g.W + (r - k))) : -1;
        }
        uint4 kv[4], vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (sr[u] >= 0) {
            kv[u] = ld_stream16(kb + sr[u] * g.k_sl + lane_chunk * 16);
            vv[u] = ld_stream16(vb + sr[u] * g.v_sl + lane_chunk * 16);
          }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (sr[u] >= 0) {
            const int r = rb + u * rows_par * C;
            st_stream16(kob + (int64_t)r * g.ko_sl + lane_chunk * 16, kv[u]);
            st_stream16(vob + (int64_t)r * g.vo_sl + lane_chunk * 16, vv[u]);
          }
      }
    }
    TK_TRACE(6);
    return;
  }

  // ---- 5. INDEX_ASC: every CTA compacts its own selected keys in index order ---------------------------
  const unsigned long long kth_c = ((unsigned long long)sm.bc[5] << 32) | sm.bc[4];
  constexpr int kPerThread = kSliceKeys / kThreads;               // 16 consecutive keys per thread
  uint32_t mine = 0;
  uint32_t selmask = 0;
#pragma unroll
  for (int e = 0; e < kPerThread; ++e) {
    const int i = tid * kPerThread + e;
    if (i < cnt) {
      const unsigned long long c = ((unsigned long long)sm.keys[key_slot(i)] << 32) | (uint32_t)(0xffffffffu - (uint32_t)(base + i));
      if (c >= kth_c) { ++mine; selmask |= 1u << e; }
    }
  }
  uint32_t cta_total;
  const uint32_t inc = block_scan_incl(mine, sm.wsum, cta_total);
  uint32_t pos = inc - mine;
  if (multi) {
    if (tid < C) cluster.map_shared_rank(sm.cnt_cta, tid)[crank] = cta_total;
    cluster.sync();
    for (int c = 0; c < crank; ++c) pos += sm.cnt_cta[c];
  }
#pragma unroll
  for (int e = 0; e < kPerThread; ++e)
    if (selmask & (1u << e)) out[pos++] = (int64_t)(base + tid * kPerThread + e);
  if (crank == 0)
    for (int i = tid; i < a.tail; i += kThreads) out[k + i] = (int64_t)(n + i);
  TK_TRACE(6);
  if (multi) cluster.sync();                        // keep this CTA's smem alive until all remote accesses finished
}

inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace

size_t topk_workspace_bytes(int rows, int n, int k) {
  size_t b = align_up((size_t)rows * n * sizeof(float), 256);                 // dense summed rows
  const int cap = next_pow2(k);
  if (cap > kSmemSortMax) b += align_up((size_t)rows * cap * sizeof(unsigned long long), 256);
  return b + 256;
}

static int launch_topk_impl(const fkv_topk_params& p, const fkv_gather_kv_params* fg, bool* fused, cudaStream_t s);

int launch_topk(const fkv_topk_params& p, cudaStream_t s) { return launch_topk_impl(p, nullptr, nullptr, s); }

int launch_topk_fused(const fkv_topk_params& p, const fkv_gather_kv_params& g, bool* fused, cudaStream_t s) {
  return launch_topk_impl(p, &g, fused, s);
}

static int launch_topk_impl(const fkv_topk_params& p, const fkv_gather_kv_params* fg, bool* fused, cudaStream_t s) {
  if (fused) *fused = false;
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
      off = align_up((size_t)p.rows * p.n * sizeof(float), 256);
      if (!ws || p.workspace_bytes < off) return FKV_EWORKSPACE;      // the summed rows live in the caller's scratch
      a.dense = reinterpret_cast<float*>(ws);
    }
  }
  const bool by_value = p.order == FKV_ORDER_VALUE_DESC;

  static const bool no_cluster = getenv("FKV_TOPK_NO_CLUSTER") != nullptr;
  // CTAs per row: ~4096 keys per CTA (short phases), at least enough for the slice to fit in smem
  long long want = ((long long)p.n * a.nsum + kTargetKeys - 1) / kTargetKeys;   // summed rows read nsum x the bytes
  int C = (int)(want > kMaxCluster ? kMaxCluster : (want < 1 ? 1 : want));
  const bool fits = ceil_div(p.n, kMaxCluster) <= kSliceKeys;
  if (!no_cluster && fits && p.k <= kKmax && (p.src_stride_row % 4) == 0) {
    int Cp = 1;
    while (Cp < C) Cp <<= 1;                        // cluster sizes 1, 2, 4, 8 (divides the 4096 bins)
    a.cluster = Cp;
    if (fg && by_value && Cp >= 4) {                  // enough CTAs per row for the copy to be worth fusing
      const int es = dtype_size(fg->dtype);
      FusedGather& g = a.g;
      g.k = (const char*)fg->k.ptr; g.v = (const char*)fg->v.ptr;
      g.ko = (char*)const_cast<void*>(fg->k_out.ptr); g.vo = (char*)const_cast<void*>(fg->v_out.ptr);
      g.k_sb = fg->k.stride_b * es; g.k_sh = fg->k.stride_h * es; g.k_sl = fg->k.stride_l * es;
      g.v_sb = fg->v.stride_b * es; g.v_sh = fg->v.stride_h * es; g.v_sl = fg->v.stride_l * es;
      g.ko_sb = fg->k_out.stride_b * es; g.ko_sh = fg->k_out.stride_h * es; g.ko_sl = fg->k_out.stride_l * es;
      g.vo_sb = fg->v_out.stride_b * es; g.vo_sh = fg->v_out.stride_h * es; g.vo_sl = fg->v_out.stride_l * es;
      g.L = fg->L; g.W = fg->window_size; g.h0 = fg->h0; g.head_div = fg->head_div; g.Hsel = fg->h1 - fg->h0;
      g.row_chunks = fg->D * es / 16;
      g.enabled = 1;
      if (fused) *fused = true;
    }
    a.write_dense = (a.nsum > 1);                   // the fallback inside the kernel needs the dense row
    {
      int rc_attr;
      if ((rc_attr = ensure_dyn_smem(topk_cluster_kernel<true>, (int)sizeof(ClusterSmem))) != FKV_OK) return rc_attr;
      if ((rc_attr = ensure_dyn_smem(topk_cluster_kernel<false>, (int)sizeof(ClusterSmem))) != FKV_OK) return rc_attr;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.rows * Cp);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = sizeof(ClusterSmem);
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = Cp; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    static const bool trace_on = getenv("FKV_TOPK_TRACE") != nullptr;
    static unsigned long long* trace_buf = nullptr;
    if (trace_on) {
      if (!trace_buf) FKV_CUDA_TRY(cudaMalloc(&trace_buf, 16 * sizeof(unsigned long long)));
      a.trace = trace_buf;
    }
    if (by_value) FKV_CUDA_TRY(cudaLaunchKernelEx(&cfg, topk_cluster_kernel<true>, a));
    else FKV_CUDA_TRY(cudaLaunchKernelEx(&cfg, topk_cluster_kernel<false>, a));
    count_launch();
    if (trace_on) {
      cudaStreamSynchronize(s);
      unsigned long long h[16];
      cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
      fprintf(stderr, "[topk trace] n=%d rows=%d C=%d order=%d:", p.n, p.rows, Cp, p.order);
      for (int i = 1; i <= 11; ++i) fprintf(stderr, " %d:%.2f", i, (double)(h[i] - h[0]) / 1e3);
      fprintf(stderr, "\n");
    }
    return FKV_OK;
  }

  size_t smem = (kMaxBins + kWarps + 4) * sizeof(uint32_t);
  if (by_value) {
    if (a.sort_cap <= kSmemSortMax) smem += (size_t)a.sort_cap * sizeof(unsigned long long);
    else {
      const size_t need = off + (size_t)p.rows * a.sort_cap * sizeof(unsigned long long);
      if (!ws || p.workspace_bytes < need) return FKV_EWORKSPACE;
      a.sort_ws = reinterpret_cast<unsigned long long*>(ws + off);
    }
  }
  {
    const int smem_max = (int)((kMaxBins + kWarps + 4) * sizeof(uint32_t) + (size_t)kSmemSortMax * sizeof(unsigned long long));
    const int rc_attr = ensure_dyn_smem(topk_row_kernel, smem_max);
    if (rc_attr != FKV_OK) return rc_attr;
  }
  topk_row_kernel<<<p.rows, kThreads, smem, s>>>(a);
  FKV_LAUNCH_CHECK();
  return FKV_OK;
}

}  // namespace fkv
