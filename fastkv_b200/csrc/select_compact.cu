// K3 + K4 in one launch: per-head top-k (score descending, ties -> lowest index) with the KV compaction fused behind
// it.  Replaces Tensor.topk + expand + 2 x gather + 2 x cat at baselines/fastkv/utils.py:113-121 (SnapKV twin
// snapkv/utils.py:88-101) for the per-head selection of fkv_compress.
//
// Design (no clusters; every SM busy; every score is read from L2 exactly once):
//   * the 12-bit histogram of a score row (sign + exponent + 3 mantissa bits of the order-preserving key) is built by
//     the PRODUCER of the scores (score_tc_kernel's pass-2 epilogue adds every score it stores to a [rows][4096]
//     table in the workspace), so this kernel starts with the threshold bin one block scan away;
//   * a row is served by R CTAs (rows * R ~ number of SMs).  CTA r scans slice r of the row and appends its
//     candidates (keys in or above the threshold bin, k .. 4096 per row) to the row's list in global memory (one
//     cursor atomic per warp); a per-row arrive/spin barrier (the launch is cooperative, so the R CTAs are
//     co-resident) publishes the list.  Redundant full-row scans by all R CTAs were measured first: 16 x the L2
//     traffic made the scan 6 us at 32k and 2 x 17 us at 128k;
//   * every CTA then loads the whole list (~21 KB), counts it on 4096 ADAPTIVE bins (the candidates' key range mapped
//     linearly onto the bins, so a bin holds ~1 key), suffix-scans the counts into rank offsets and groups the
//     candidates by bin in shared memory: no sort;
//   * the CTAs of a row split the candidates by token index (idx mod R).  A thread ranks an owned candidate exactly:
//     rank = keys in higher bins (suffix sum) + larger keys inside its own (tiny) bin; the CTA writes idx_out[rank]
//     and copies the K and V rows straight to row `rank` of the compact cache (16-byte chunks, 4 in flight per
//     thread): the copy is spread over all rows * R CTAs;
//   * crowded threshold bin (narrow score distributions): the bin is split by the next 12 key bits through a second
//     per-row table (slice histograms flushed with REDs, one more barrier); still more than 4096 candidates (heavy
//     exact ties): CTA 0 of the row falls back to the generic radix select;
//   * no producer-built table (fp32 scoring path): the CTAs build it from their slices first (one more barrier).
// The last CTA to leave a row zeroes its tables and counters, so the workspace is clean for the next launch.
#include <cstdio>

#include "topk_common.cuh"

namespace fkv {
namespace {

constexpr int kSelKmax = 2048;                    // largest k served (candidates must fit the 4096-entry buffers)
constexpr int kSelCand = 4096;
constexpr int kSpinLimit = 1 << 22;               // bounded spin: a broken barrier flags the workspace instead of hanging

// per-row control words in the persistent workspace region (zero between launches)
// (word 2-3 is ONE 64-bit word: low half = list cursor, high half = arrivals at the list barrier, so the CTA that sees
// the barrier open has the list length from the same load)
enum { kCtlBarH = 0, kCtlBarS = 1, kCtlList = 2, kCtlDepart = 4, kCtlWords = 8 };
static_assert(kCtlWords * 4 == 32, "control block size is part of the workspace layout (common.cuh)");

struct SelArgs {
  const float* score; int64_t stride_row;         // [rows][n], rows = B * Hsel
  int n, k, R, rshift;                            // R = CTAs per row (power of two), rshift = log2(R)
  int have_hist;                                  // the producer of the scores filled hist[]
  int64_t* idx_out; int64_t idx_stride_row;
  uint32_t* ctl;                                  // [rows][kCtlWords]
  uint32_t* hist;                                 // [rows][4096] 12-bit histogram of the row
  uint32_t* subhist;                              // [rows][4096] second-level histogram of the threshold bin
  unsigned long long* list;                       // [rows][kSelCand] candidate lists (scratch)
  uint32_t* err;                                  // workspace error word
  const char* kp; const char* vp; char* ko; char* vo;
  int64_t k_sb, k_sh, k_sl, v_sb, v_sh, v_sl, ko_sb, ko_sh, ko_sl, vo_sb, vo_sh, vo_sl;   // bytes
  int L, W, h0, head_div, Hsel, row_chunks;
  int sort_cap;                                   // pow2 >= k (generic fallback)
  unsigned long long* trace;                      // FKV_SELECT_TRACE=1: %globaltimer stamps per CTA
};

#define SEL_TRACE(i) do { if (a.trace && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); a.trace[16 + blockIdx.x * 8 + (i)] = t_; } } while (0)

struct SelSmem {
  unsigned long long list[kSelCand];              // the row's candidates (key << 32 | ~index) as published
  unsigned long long grouped[kSelCand];           // the same, grouped by adaptive bin, bins descending
  uint32_t cursor[kMaxBins];                      // slice histograms, then adaptive-bin counts
  uint32_t fill[kMaxBins];                        // adaptive-bin scatter cursors
  uint32_t fine_off[kMaxBins];                    // candidates in strictly higher adaptive bins == segment offset
  uint2 picked[kSelKmax];                         // (token, rank) of the selected candidates owned by this CTA
  uint32_t wsum[kWarps];
  uint32_t bc[16];                                // [0] bin*, [1] keys above bin*, [2] keys in bin*, [6] top non-empty bin,
                                                  // [9] entries in picked[], [11] total of the last scan,
                                                  // [13] / [14] min / max key of a short row
};

// inverse of ordered_key
__device__ __forceinline__ float key_to_float(uint32_t key) {
  return __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
}

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Per-row barrier between the R CTAs that serve a row: everything this CTA wrote before is visible to the others
// after.  One thread arrives (release) and spins (acquire); the rest of the CTA waits at the closing __syncthreads.
// (the arrival is a gpu-scope RELEASE by one thread after the CTA barrier: cumulativity covers the whole CTA's writes,
// which is cheaper than a membar by every thread)
__device__ __forceinline__ void row_barrier(uint32_t* counter, uint32_t R, uint32_t* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    int spins = 0;
    while (ld_acquire_u32(counter) < R) {
      if (++spins > kSpinLimit) { *err = 2u; break; }        // never hang the device: flag the workspace and go on
    }
  }
  __syncthreads();
}

// The list barrier: arrivals live in the high half of the 64-bit list word; returns the list length (low half).
__device__ __forceinline__ uint32_t list_barrier(unsigned long long* word, uint32_t R, uint32_t* err, uint32_t* bcast) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(word), "l"(1ull << 32) : "memory");
    unsigned long long v;
    int spins = 0;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(word) : "memory");
      if ((uint32_t)(v >> 32) >= R) break;
      if (++spins > kSpinLimit) { *err = 2u; break; }
    }
    *bcast = (uint32_t)(v & 0xffffffffull);
  }
  __syncthreads();
  return *bcast;
}

// The slice of the row this CTA scans: f(index, value) for each of its keys (float4 granules, remainder to the last CTA).
template <typename F>
__device__ __forceinline__ void slice_pass(const float* __restrict__ src, int n, bool vec_ok, uint32_t r, int R, F&& f) {
  const int tid = threadIdx.x;
  if (vec_ok) {
    const int n4 = n >> 2;
    const int per = (n4 + R - 1) / R;
    const int i_begin = min(n4, (int)r * per), i_end = min(n4, i_begin + per);
    const float4* src4 = reinterpret_cast<const float4*>(src);
    for (int i0 = i_begin; i0 < i_end; i0 += kThreads * 4) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kThreads + tid;
        if (i < i_end) v[u] = __ldg(src4 + i);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * kThreads + tid;
        if (i < i_end) { f(4 * i, v[u].x); f(4 * i + 1, v[u].y); f(4 * i + 2, v[u].z); f(4 * i + 3, v[u].w); }
      }
    }
    if ((int)r == R - 1)
      for (int j = (n4 << 2) + tid; j < n; j += kThreads) f(j, __ldg(src + j));
  } else {
    const int per = (n + R - 1) / R;
    const int j_begin = min(n, (int)r * per), j_end = min(n, j_begin + per);
    for (int j = j_begin + tid; j < j_end; j += kThreads) f(j, __ldg(src + j));
  }
}

// Suffix scan of a 4096-bin histogram held 4 bins per thread (thread t: bins 4095-4t .. 4092-4t, top first).
// Reports the bin holding the `need`-th largest key through bc[0..2], the top non-empty bin through bc[6] and the
// total count through bc[11]; off_out (optional) receives the number of keys in strictly higher bins for every bin.
__device__ __forceinline__ void suffix_scan_bins(const uint32_t (&local)[4], uint32_t need, uint32_t* off_out,
                                                 uint32_t* wsum, uint32_t* bc) {
  const int tid = threadIdx.x;
  const uint32_t ls = local[0] + local[1] + local[2] + local[3];
  uint32_t total;
  const uint32_t incl = block_scan_incl(ls, wsum, total);
  uint32_t above = incl - ls;
  if (tid == 0) bc[11] = total;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int bin = kMaxBins - 1 - (tid * 4 + i);
    if (off_out) off_out[bin] = above;
    if (need && above < need && need <= above + local[i]) { bc[0] = (uint32_t)bin; bc[1] = above; bc[2] = local[i]; }
    if (above == 0 && local[i] != 0) bc[6] = (uint32_t)bin;
    above += local[i];
  }
  __syncthreads();
}

// this thread's 4 bins (top first) of a [4096] table in global memory
__device__ __forceinline__ void load_bins(const uint32_t* table, uint32_t (&local)[4]) {
  const uint4 h = __ldcg(reinterpret_cast<const uint4*>(table + (kMaxBins - 4 - threadIdx.x * 4)));
  local[0] = h.w; local[1] = h.z; local[2] = h.y; local[3] = h.x;
}

// add this CTA's shared-memory slice histogram to the row's table (non-empty bins only) and clear it
__device__ __forceinline__ void flush_bins(uint32_t* smem_bins, uint32_t* table) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int bin = threadIdx.x * 4 + i;
    const uint32_t c = smem_bins[bin];
    if (c) { atomicAdd(table + bin, c); smem_bins[bin] = 0; }
  }
}

__global__ void __launch_bounds__(kThreads, 1)
select_compact_kernel(SelArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SelSmem& sm = *reinterpret_cast<SelSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int row = blockIdx.x >> a.rshift;
  const uint32_t r = blockIdx.x & (a.R - 1);
  const uint32_t R = (uint32_t)a.R;
  const int n = a.n, k = a.k;

  // prologue: nothing here depends on the producer of the scores
  SEL_TRACE(0);
  for (int i = tid; i < kMaxBins; i += kThreads) { sm.cursor[i] = 0; sm.fill[i] = 0; }
  if (tid < 16) sm.bc[tid] = (tid == 13) ? 0xffffffffu : 0u;
  const int b = row / a.Hsel, hs = row - b * a.Hsel;
  const int src_head = (a.h0 + hs) / a.head_div;
  const char* kb = a.kp + b * a.k_sb + src_head * a.k_sh;
  const char* vb = a.vp + b * a.v_sb + src_head * a.v_sh;
  char* kob = a.ko + b * a.ko_sb + hs * a.ko_sh;
  char* vob = a.vo + b * a.vo_sb + hs * a.vo_sh;
  int64_t* out = a.idx_out + (int64_t)row * a.idx_stride_row;
  uint32_t* ctl = a.ctl + (size_t)row * kCtlWords;
  uint32_t* hist = a.hist + (size_t)row * kMaxBins;
  uint32_t* subhist = a.subhist + (size_t)row * kMaxBins;
  unsigned long long* glist = a.list + (size_t)row * kSelCand;
  const int chunks2 = 2 * a.row_chunks;             // 16-byte chunks of one K row + one V row
  __syncthreads();
  pdl_enter();                                      // the scores (and their histogram) are complete and visible
  SEL_TRACE(1);

  const float* src = a.score + (int64_t)row * a.stride_row;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(src) & 15) == 0;

  // the W window rows are kept unconditionally (utils.py:116-121): one CTA of the row copies them right away
  if (r == R - 1) {
    for (int c = tid; c < a.W * chunks2; c += kThreads) {
      const int wr = c / chunks2, ch = c - wr * chunks2;
      const bool is_v = ch >= a.row_chunks;
      const int cc = is_v ? ch - a.row_chunks : ch;
      const int64_t sr = a.L - a.W + wr;
      const char* s = (is_v ? vb + sr * a.v_sl : kb + sr * a.k_sl) + cc * 16;
      char* d = (is_v ? vob + (int64_t)(k + wr) * a.vo_sl : kob + (int64_t)(k + wr) * a.ko_sl) + cc * 16;
      st_stream16(d, ld_stream16(s));
    }
  }

  // Short rows (n <= 4096, e.g. every layer behind the TSP layer): all keys fit the candidate buffers, so every CTA
  // loads the whole row itself and ranks its share: no histogram, no list exchange, no barrier, no control words.
  const bool small = n <= kSelCand;
  uint32_t n_cand = (uint32_t)n, thr = 0u, bin_star = 0u, top_bin = 0u, n_above = 0u;
  if (small) {
    uint32_t kmin = 0xffffffffu, kmax = 0u;
    for (int j = tid; j < n; j += kThreads) {
      const uint32_t key = ordered_key(__ldg(src + j));
      sm.list[j] = ((unsigned long long)key << 32) | (uint32_t)(0xffffffffu - (uint32_t)j);
      kmin = min(kmin, key); kmax = max(kmax, key);
    }
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    if (lane == 0) { atomicMin(&sm.bc[13], kmin); atomicMax(&sm.bc[14], kmax); }
    __syncthreads();
  } else {
  // ---- A. threshold bin from the row's 12-bit histogram ----------------------------------------------------
  if (!a.have_hist) {
    // no producer-built table (fp32 scoring path): histogram of the slice in shared memory, flushed to the row's table
    slice_pass(src, n, vec_ok, r, a.R, [&](int, float x) { atomicAdd(&sm.cursor[ordered_key(x) >> 20], 1u); });
    __syncthreads();
    flush_bins(sm.cursor, hist);
    row_barrier(ctl + kCtlBarH, R, a.err);
  }
  uint32_t local[4];
  load_bins(hist, local);
  suffix_scan_bins(local, (uint32_t)k, nullptr, sm.wsum, sm.bc);
  SEL_TRACE(2);
  bin_star = sm.bc[0];
  n_above = sm.bc[1];
  top_bin = sm.bc[6];
  n_cand = n_above + sm.bc[2];
  thr = bin_star << 20;                             // candidates are the keys >= thr

  // ---- B. crowded threshold bin: split it by the next 12 key bits through the row's second-level table --------
  if (n_cand > (uint32_t)kSelCand) {
    const float lo_f = key_to_float(bin_star << 20), hi_f = key_to_float((bin_star << 20) | 0xfffffu);
    const bool by_float = (lo_f == lo_f) && (hi_f == hi_f);   // the key map is monotonic: one float range test per key
    slice_pass(src, n, vec_ok, r, a.R, [&](int, float x) {
      if (by_float) {
        if (x >= lo_f && x <= hi_f) atomicAdd(&sm.cursor[(ordered_key(x) >> 8) & 0xfffu], 1u);
      } else {
        const uint32_t key = ordered_key(x);
        if ((key >> 20) == bin_star) atomicAdd(&sm.cursor[(key >> 8) & 0xfffu], 1u);
      }
    });
    __syncthreads();
    flush_bins(sm.cursor, subhist);
    row_barrier(ctl + kCtlBarS, R, a.err);
    uint32_t l2[4];
    load_bins(subhist, l2);
    suffix_scan_bins(l2, (uint32_t)k - n_above, nullptr, sm.wsum, sm.bc);
    thr = (bin_star << 20) | (sm.bc[0] << 8);
    n_cand = n_above + sm.bc[1] + sm.bc[2];
  }
  }                                                 // !small
  SEL_TRACE(3);

  if (n_cand > (uint32_t)kSelCand) {
    // ---- heavy exact ties / very flat rows: CTA 0 of the row re-selects with the generic radix path and copies ----
    if (r == 0) {
      __syncthreads();
      select_row_generic(src, n, k, FKV_ORDER_VALUE_DESC, 0, out, sm.cursor, sm.wsum, sm.bc, sm.grouped, a.sort_cap);
      __syncthreads();
      for (int c = tid; c < k * chunks2; c += kThreads) {
        const int rr = c / chunks2, ch = c - rr * chunks2;
        const bool is_v = ch >= a.row_chunks;
        const int cc = is_v ? ch - a.row_chunks : ch;
        const int64_t sr = out[rr];
        const char* s = (is_v ? vb + sr * a.v_sl : kb + sr * a.k_sl) + cc * 16;
        char* d = (is_v ? vob + (int64_t)rr * a.vo_sl : kob + (int64_t)rr * a.ko_sl) + cc * 16;
        st_stream16(d, ld_stream16(s));
      }
      __syncthreads();
      for (int i = tid; i < kMaxBins; i += kThreads) sm.cursor[i] = 0;   // (the generic path used it as its histogram)
    }
  } else {
    uint32_t n_list, klo, khi;
    const uint32_t rmask = R - 1u;
    // (L2 prefetches of the owned candidates' K / V rows at this point were measured: no gain, 20.5 vs 20.6 us; so were
    // prefetches at kernel start on the k == n short rows, whose owned tokens are known up front: 19.4 vs 19.2 us)
    if (small) {
      n_list = (uint32_t)n; klo = sm.bc[13]; khi = sm.bc[14];
    } else {
    // ---- C. publish this slice's candidates in the row's list (one cursor atomic per warp and load round) --------
    const float thr_f = key_to_float(thr);
    const bool by_float = (thr_f == thr_f);          // monotonic key map: the filter is one float compare per key
    {
      // two steps per round so that the warp needs ONE global atomic: count, reserve, write
      const int n4 = n >> 2;
      const int per = vec_ok ? (n4 + a.R - 1) / a.R : (n + a.R - 1) / a.R;
      const int lim = vec_ok ? n4 : n;
      const int i_begin = min(lim, (int)r * per), i_end = min(lim, i_begin + per);
      for (int i0 = i_begin; i0 < i_end; i0 += kThreads * 4) {       // uniform over the CTA
        float x[16];
        uint32_t cm = 0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * kThreads + tid;
          if (vec_ok) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i < i_end) v = __ldg(reinterpret_cast<const float4*>(src) + i);
            x[4 * u] = v.x; x[4 * u + 1] = v.y; x[4 * u + 2] = v.z; x[4 * u + 3] = v.w;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              if (i < i_end && (by_float ? (x[4 * u + c] >= thr_f) : (ordered_key(x[4 * u + c]) >= thr))) cm |= 1u << (4 * u + c);
          } else {
            x[4 * u] = (i < i_end) ? __ldg(src + i) : 0.f;
            if (i < i_end && (by_float ? (x[4 * u] >= thr_f) : (ordered_key(x[4 * u]) >= thr))) cm |= 1u << (4 * u);
          }
        }
        // warp-level reservation
        const uint32_t mine = (uint32_t)__popc(cm);
        uint32_t pre = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xffffffffu, pre, o);
          if (lane >= o) pre += t;
        }
        const uint32_t wtotal = __shfl_sync(0xffffffffu, pre, 31);
        uint32_t wbase = 0;
        if (lane == 31 && wtotal)
          wbase = (uint32_t)atomicAdd(reinterpret_cast<unsigned long long*>(ctl + kCtlList), (unsigned long long)wtotal);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        uint32_t pos = wbase + pre - mine;
#pragma unroll
        for (int e = 0; e < 16; ++e)
          if (cm & (1u << e)) {
            const int ie = i0 + (e >> 2) * kThreads + tid;
            const int j = vec_ok ? 4 * ie + (e & 3) : ie;
            if (pos < (uint32_t)kSelCand)
              glist[pos] = ((unsigned long long)ordered_key(x[e]) << 32) | (uint32_t)(0xffffffffu - (uint32_t)j);
            ++pos;
          }
      }
      if (vec_ok && r == R - 1) {                   // <= 3 tail keys past the last float4
        const int j = (n4 << 2) + tid;
        if (j < n) {
          const float xv = __ldg(src + j);
          if (by_float ? (xv >= thr_f) : (ordered_key(xv) >= thr)) {
            const uint32_t pos = (uint32_t)atomicAdd(reinterpret_cast<unsigned long long*>(ctl + kCtlList), 1ull);
            if (pos < (uint32_t)kSelCand)
              glist[pos] = ((unsigned long long)ordered_key(xv) << 32) | (uint32_t)(0xffffffffu - (uint32_t)j);
          }
        }
      }
    }
    n_list = min(list_barrier(reinterpret_cast<unsigned long long*>(ctl + kCtlList), R, a.err, &sm.bc[15]), (uint32_t)kSelCand);
    klo = thr; khi = (top_bin << 20) | 0xfffffu;
    for (uint32_t i = tid; i < n_list; i += kThreads) {
      sm.list[i] = __ldcg(glist + i);
    }
    __syncthreads();
    }                                               // !small
    SEL_TRACE(4);

    // ---- D0. count the list on 4096 adaptive bins over the candidates' key range ---------------------------------
    const int range_bits = 32 - __clz(khi - klo);
    const int fshift = range_bits > 12 ? range_bits - 12 : 0;     // fine bin = (key - klo) >> fshift < 4096
    for (uint32_t i = tid; i < n_list; i += kThreads)
      atomicAdd(&sm.cursor[((uint32_t)(sm.list[i] >> 32) - klo) >> fshift], 1u);
    __syncthreads();
    uint32_t fl[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) fl[i] = sm.cursor[kMaxBins - 1 - (tid * 4 + i)];
    suffix_scan_bins(fl, 0u, sm.fine_off, sm.wsum, sm.bc);
    for (uint32_t i = tid; i < n_list; i += kThreads) {           // group by bin
      const unsigned long long c = sm.list[i];
      const uint32_t fb = ((uint32_t)(c >> 32) - klo) >> fshift;
      sm.grouped[sm.fine_off[fb] + atomicAdd(&sm.fill[fb], 1u)] = c;
    }
    __syncthreads();

    SEL_TRACE(6);
    // ---- D1. exact ranks of the owned candidates (idx mod R); the selected ones go to a (token, rank) list ------
    for (uint32_t p = tid; p < n_list; p += kThreads) {
      const unsigned long long c = sm.grouped[p];
      const uint32_t tok = 0xffffffffu - (uint32_t)(c & 0xffffffffull);
      if ((tok & rmask) != r) continue;
      const uint32_t fb = ((uint32_t)(c >> 32) - klo) >> fshift;
      const uint32_t s0 = sm.fine_off[fb], cnt = sm.cursor[fb];
      if (s0 >= (uint32_t)k) continue;               // k or more keys in higher bins: cannot be selected
      uint32_t rank = s0;
      for (uint32_t q = 0; q < cnt; ++q) rank += (sm.grouped[s0 + q] > c) ? 1u : 0u;
      if (rank < (uint32_t)k) {
        out[rank] = (int64_t)tok;
        sm.picked[atomicAdd(&sm.bc[9], 1u)] = make_uint2(tok, rank);
      }
    }
    __syncthreads();
    SEL_TRACE(7);

    // ---- D2. the whole CTA copies its selected K / V rows: every thread keeps 4 x 16 bytes in flight ------------
    const int total = (int)sm.bc[9] * chunks2;
    for (int c0 = tid; c0 < total; c0 += kThreads * 4) {
      uint4 data[4];
      int64_t dst_off[4];
      bool dv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * kThreads;
        if (c < total) {
          const int e = c / chunks2, ch = c - e * chunks2;
          const uint2 pr = sm.picked[e];
          const bool is_v = ch >= a.row_chunks;
          const int cc2 = (is_v ? ch - a.row_chunks : ch) * 16;
          data[u] = ld_stream16((is_v ? vb + (int64_t)pr.x * a.v_sl : kb + (int64_t)pr.x * a.k_sl) + cc2);
          dst_off[u] = (int64_t)pr.y * (is_v ? a.vo_sl : a.ko_sl) + cc2;
          dv[u] = is_v;
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int c = c0 + u * kThreads;
        if (c < total) st_stream16((dv[u] ? vob : kob) + dst_off[u], data[u]);
      }
    }
  }

  SEL_TRACE(5);
  if (a.trace && blockIdx.x == 0 && tid == 0) { a.trace[8] = n_cand; a.trace[9] = bin_star; a.trace[10] = thr; }
  // the last CTA to leave the row hands clean tables and counters to the next launch (warp 0 only: the other
  // warps are done; short rows never touched the tables)
  if (!small && tid < 32) {
    uint32_t last = 0;
    if (lane == 0) last = (atomicAdd(ctl + kCtlDepart, 1u) == R - 1u) ? 1u : 0u;
    if (__shfl_sync(0xffffffffu, last, 0)) {
      for (int i = lane; i < kMaxBins / 4; i += 32) {
        reinterpret_cast<uint4*>(hist)[i] = make_uint4(0u, 0u, 0u, 0u);
        reinterpret_cast<uint4*>(subhist)[i] = make_uint4(0u, 0u, 0u, 0u);
      }
      if (lane < kCtlWords) ctl[lane] = 0u;
    }
  }
}

inline int next_pow2_sel(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace

bool select_compact_supported(const fkv_topk_params& p, const fkv_gather_kv_params& g) {
  static const bool off = getenv("FKV_NO_SELECT_COMPACT") != nullptr;
  if (off) return false;
  if (p.order != FKV_ORDER_VALUE_DESC || p.nsum > 1 || p.tail != 0) return false;
  if (p.k < 1 || p.k > kSelKmax || p.k > p.n) return false;
  if ((long long)p.rows * p.n >= (1ll << 31) || p.rows > kHistMaxRows) return false;
  const int es = dtype_size(g.dtype);
  const int row_chunks = g.D * es / 16;
  return row_chunks >= 8 && row_chunks <= 32;
}

// `workspace`: base of the call's workspace (persistent region in front); `slot`: 1 or 2, the table set to use;
// `have_hist`: the producer of the scores filled the slot's 12-bit histograms for these rows.
int launch_select_compact(const fkv_topk_params& p, const fkv_gather_kv_params& g, void* workspace, int slot,
                          bool have_hist, cudaStream_t s, int max_ctas) {
  if (!workspace || p.rows > kHistMaxRows || slot < 1 || slot > kHistSlots) return FKV_EWORKSPACE;
  char* wsb = static_cast<char*>(workspace);
  const size_t row0 = (size_t)(slot - 1) * kHistMaxRows;
  const int es = dtype_size(g.dtype);
  SelArgs a{};
  a.score = p.src; a.stride_row = p.src_stride_row; a.n = p.n; a.k = p.k;
  a.idx_out = p.out; a.idx_stride_row = p.out_stride_row;
  a.have_hist = have_hist ? 1 : 0;
  a.ctl = reinterpret_cast<uint32_t*>(wsb + kWsCtlOffset) + row0 * kCtlWords;
  a.hist = reinterpret_cast<uint32_t*>(wsb + kWsHistOffset) + row0 * kMaxBins;
  a.subhist = reinterpret_cast<uint32_t*>(wsb + kWsSubHistOffset) + row0 * kMaxBins;
  a.list = reinterpret_cast<unsigned long long*>(wsb + kWsListOffset) + row0 * kSelCand;
  a.err = reinterpret_cast<uint32_t*>(wsb) + 5;
  a.kp = (const char*)g.k.ptr; a.vp = (const char*)g.v.ptr;
  a.ko = (char*)const_cast<void*>(g.k_out.ptr); a.vo = (char*)const_cast<void*>(g.v_out.ptr);
  a.k_sb = g.k.stride_b * es; a.k_sh = g.k.stride_h * es; a.k_sl = g.k.stride_l * es;
  a.v_sb = g.v.stride_b * es; a.v_sh = g.v.stride_h * es; a.v_sl = g.v.stride_l * es;
  a.ko_sb = g.k_out.stride_b * es; a.ko_sh = g.k_out.stride_h * es; a.ko_sl = g.k_out.stride_l * es;
  a.vo_sb = g.v_out.stride_b * es; a.vo_sh = g.v_out.stride_h * es; a.vo_sl = g.v_out.stride_l * es;
  a.L = g.L; a.W = g.window_size; a.h0 = g.h0; a.head_div = g.head_div; a.Hsel = g.h1 - g.h0;
  a.row_chunks = g.D * es / 16;
  a.sort_cap = next_pow2_sel(p.k);
  // CTAs per row: fill the chip once (the launch is cooperative: all CTAs must be co-resident), but leave every CTA at
  // least ~32 selected rows to rank and copy
  int R = 1;
  int sms = num_sms();
  if (max_ctas > 0 && max_ctas < sms) sms = max_ctas;
  static const int env_r = getenv("FKV_SELECT_R") ? atoi(getenv("FKV_SELECT_R")) : 0;
  while (R < 32 && (long long)p.rows * (R * 2) <= sms && p.k / (R * 2) >= 32) R *= 2;
  if (env_r > 0) { R = 1; while (R * 2 <= env_r && R < 32 && (long long)p.rows * (R * 2) <= sms) R *= 2; }
  a.R = R;
  a.rshift = 0;
  while ((1 << a.rshift) < R) ++a.rshift;
  static const bool trace_on = getenv("FKV_SELECT_TRACE") != nullptr;
  static unsigned long long* trace_buf = nullptr;
  if (trace_on) {
    if (!trace_buf) FKV_CUDA_TRY(cudaMalloc(&trace_buf, (16 + 8 * 4096) * sizeof(unsigned long long)));
    a.trace = trace_buf;
  }
  int rc;
  const size_t smem = sizeof(SelSmem);
  if ((rc = ensure_dyn_smem(select_compact_kernel, (int)smem)) != FKV_OK) return rc;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.rows * R);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  static const bool no_coop = getenv("FKV_NO_COOP") != nullptr;   // experiments only
  if (R > 1 && p.n > kSelCand && !no_coop) {        // the R CTAs of a row wait for each other: they must be co-resident
    attr[na].id = cudaLaunchAttributeCooperative;
    attr[na].val.cooperative = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  FKV_CUDA_TRY(cudaLaunchKernelEx(&cfg, select_compact_kernel, a));
  count_launch();
  if (trace_on) {                                   // debugging aid only: synchronises
    cudaStreamSynchronize(s);
    static unsigned long long h[16 + 8 * 4096];
    const int nct = p.rows * R > 4096 ? 4096 : p.rows * R;
    cudaMemcpy(h, trace_buf, (16 + 8 * (size_t)nct) * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < nct; ++c) if (h[16 + c * 8] < t0) t0 = h[16 + c * 8];
    const char* names[8] = {"start", "enter", "thr-bin", "refined", "list published", "copied (end)", "grouped", "ranked"};
    fprintf(stderr, "[select trace] n=%d k=%d rows=%d R=%d hist=%d n_cand(cta0) %llu bin* %llu thr %llx\n", p.n, p.k, p.rows, R, have_hist ? 1 : 0, h[8], h[9], h[10]);
    for (int st : {0, 1, 2, 3, 4, 6, 7, 5}) {
      unsigned long long mn = ~0ull, mx = 0; double avg = 0; int amx = 0;
      for (int c = 0; c < nct; ++c) { const unsigned long long v = h[16 + c * 8 + st] - t0; if (v < mn) mn = v; if (v > mx) { mx = v; amx = c; } avg += v; }
      fprintf(stderr, "[select trace]   %-14s min %7.2f avg %7.2f max %7.2f us (cta %d)\n", names[st], mn / 1e3, avg / nct / 1e3, mx / 1e3, amx);
    }
  }
  return FKV_OK;
}

}  // namespace fkv
