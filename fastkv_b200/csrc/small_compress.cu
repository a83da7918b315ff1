// Opt-in one-kernel path for short contexts (FKV_SMALL_FUSE=1): see small_compress_kernel below.  Shares the TMA /
// tcgen05 plumbing of score_tc.cu through tc_common.cuh.
#include "tc_common.cuh"

namespace fkv {
namespace {

// ---- short contexts: the whole update_kv chain in ONE cluster kernel --------------------------------------------
// Post-TSP layers see only tsp_length (2048) tokens: three launches (score, top-k, compaction) of a few
// microseconds each are then dominated by launch gaps and cross-CTA round trips through L2.  Here one
// thread-block cluster of 8 CTAs owns a (b, kv head) pair end to end:
//   TMA -> tcgen05.mma -> logits of <= 4 tiles per CTA in TMEM           (baselines/fastkv/utils.py:93-101)
//   exact column max + sum over the CTA's keys, exchanged through DSMEM   (:103)
//   votes -> pool -> group-sum from the logits still in TMEM              (:104-112)
//   all scores of the pair sent to the leader as (key, ~index) composites, bitonic sort in its smem  (:113)
//   the sorted indices fan back out through DSMEM and all 8 CTAs copy the selected K/V rows + window (:114-121)
namespace cg = cooperative_groups;
constexpr int kSfCluster = 8;
constexpr int kSfTiles = 4;                        // tiles per CTA (TMEM columns 4 x 32, smem 4 x 32 KB)
constexpr int kSfWorkers = 512;                    // worker threads (warps 0-3 also own the TMEM lanes)
constexpr int kSfThreads = kSfWorkers + 32;        // + the TMA/MMA warp
constexpr int kSfIssueWarp = kSfWorkers / 32;
constexpr int kSfMaxKeys = 4096;

struct SfArgs {
  int L, P, D, nkv, kv0, n_tiles, S, half, ks, pooling, is_f16, q_heads_per_kv, n_sel, W, sort_cap;
  int kpos_l, kpos_h, kpos_b, qpos_l, qpos_h, qpos_b;
  float c;
  float* score; int64_t score_sb, score_sh; int out_h0;
  int64_t* idx_out; int64_t idx_sb, idx_sh;
  const char* k; const char* v; char* ko; char* vo;
  int64_t k_sb, k_sh, k_sl, v_sb, v_sh, v_sl, ko_sb, ko_sh, ko_sl, vo_sb, vo_sh, vo_sl;   // bytes
  int row_chunks;
  unsigned long long* trace;                          // FKV_SF_TRACE=1: %globaltimer stamps of CTA 0
};
#define SF_TRACE(i) do { if (a.trace && blockIdx.x == 0 && threadIdx.x == 0) a.trace[i] = gtimer(); } while (0)

__device__ __forceinline__ uint32_t sf_ordered_key(float x) {   // same order-preserving map as the top-k kernels
  x += 0.0f;
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <int W, int G>
__global__ void __launch_bounds__(kSfThreads, 1)
small_compress_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_q, const SfArgs a) {
  constexpr int N = W * G;
  constexpr int NMMA = N < 16 ? 16 : N;
  static_assert(N <= 32 && (N & 7) == 0, "unsupported window x group size");
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (base - smem_u32(smem_dyn));
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  SF_TRACE(0);
  const int pair = blockIdx.x / kSfCluster;
  const int b = pair / a.nkv, kvl = pair - b * a.nkv, kvh = a.kv0 + kvl;

  const int tile_bytes = kTile * a.D * 2;
  const int q_bytes = (NMMA * a.D * 2 + 1023) & ~1023;
  const uint32_t k_off = 0;
  const uint32_t q_off = kSfTiles * tile_bytes;
  unsigned long long* cand = reinterpret_cast<unsigned long long*>(sm + q_off + q_bytes);   // leader: [sort_cap] composites
  float2* part = reinterpret_cast<float2*>(cand + kSfMaxKeys);          // [cluster][N] per-CTA (max, sum)
  float2* colstat = part + kSfCluster * N;                              // [N] final (max, 1/sum)
  float* wm = reinterpret_cast<float*>(colstat + N);                    // [4 warps][N] column max / sum staging
  float* cmax = wm + 4 * N;                                             // [N] CTA column max
  float* vbuf = cmax + N;                                               // [2][G][128]
  int* sel = reinterpret_cast<int*>(vbuf + 2 * G * kTile);              // [rows of this CTA] source rows to copy
  uint32_t* cursor = reinterpret_cast<uint32_t*>(sel + kSfMaxKeys / kSfCluster + 64);   // [4096] leader: adaptive-bin counts / cursors
  uint64_t* bars = reinterpret_cast<uint64_t*>(cursor + 4096);        // kfull[4] tfull[4] qfull
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
  auto bar_kfull = [&](int i) { return smem_u32(bars + i); };
  auto bar_tfull = [&](int i) { return smem_u32(bars + 4 + i); };
  const uint32_t bar_qfull = smem_u32(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, tid = threadIdx.x;
  int T = 0;                                         // my tiles: crank, crank + 8, ...
  for (int t = crank; t < a.n_tiles; t += kSfCluster) ++T;

  if (warp == kSfIssueWarp) {
    if (lane < 9) mbar_init(smem_u32(bars + lane), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(128));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_k) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
    }
  }
  if (crank == 0) {                                  // sort padding (composites 0 order after every real key), zero bin counts
    for (int i = a.P + tid; i < a.sort_cap; i += kSfThreads) cand[i] = 0ull;
    for (int i = tid; i < 4096; i += kSfThreads) cursor[i] = 0u;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_enter();                                       // nothing global is touched before this point
  SF_TRACE(1);

  if (warp == kSfIssueWarp && lane == 0) {                      // all loads of this CTA are issued up front
    const int kl1 = a.kpos_l == 1, kl2 = a.kpos_l == 2, kl3 = a.kpos_l == 3;
    const int hb1 = kvh * (a.kpos_h == 1) + b * (a.kpos_b == 1), hb2 = kvh * (a.kpos_h == 2) + b * (a.kpos_b == 2),
              hb3 = kvh * (a.kpos_h == 3) + b * (a.kpos_b == 3);
    const int slabs = a.D >> 6;
    if (T > 0) {
      mbar_expect_tx(bar_qfull, N * a.D * 2);
      for (int slab = 0; slab < slabs; ++slab)
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const int hq = kvh * a.q_heads_per_kv + g;
          tma_load_4d(base + q_off + slab * (NMMA * 128) + g * (W * 128), &tm_q, bar_qfull, slab * 64,
                      a.P * (a.qpos_l == 1) + hq * (a.qpos_h == 1) + b * (a.qpos_b == 1),
                      a.P * (a.qpos_l == 2) + hq * (a.qpos_h == 2) + b * (a.qpos_b == 2),
                      a.P * (a.qpos_l == 3) + hq * (a.qpos_h == 3) + b * (a.qpos_b == 3));
        }
    }
    for (int i = 0; i < T; ++i) {
      const int row0 = (crank + i * kSfCluster) * a.S - a.half;
      mbar_expect_tx(bar_kfull(i), tile_bytes);
      const uint32_t kdst = base + k_off + i * tile_bytes;
      const int c1 = row0 * kl1 + hb1, c2 = row0 * kl2 + hb2, c3 = row0 * kl3 + hb3;
      tma_load_4d(kdst, &tm_k, bar_kfull(i), 0, c1, c2, c3);
      if (slabs == 2) tma_load_4d(kdst + kTile * 128, &tm_k, bar_kfull(i), 64, c1, c2, c3);
    }
  }
  cluster.sync();                                    // #1: every CTA of the cluster is running (DSMEM is safe from here)
  SF_TRACE(2);

  if (warp == kSfIssueWarp && lane == 0 && T > 0) {             // MMA issue
    const uint32_t idesc = (1u << 4) | ((a.is_f16 ? 0u : 1u) << 7) | ((a.is_f16 ? 0u : 1u) << 10) |
                           ((uint32_t)(NMMA >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
    const uint64_t desc_hi = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
    const uint32_t qa = base + q_off;
    const uint64_t bd0 = desc_hi | (uint64_t)((qa & 0x3ffffu) >> 4);
    mbar_wait(bar_qfull, 0);
    for (int i = 0; i < T; ++i) {
      mbar_wait(bar_kfull(i), 0);
      tc_fence_after();
      const uint32_t ka = base + k_off + i * tile_bytes;
      const uint64_t ad0 = desc_hi | (uint64_t)((ka & 0x3ffffu) >> 4);
      const uint32_t d_tmem = tmem_base + i * NMMA;
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(d_tmem, ad0 + 2 * k, bd0 + 2 * k, idesc, k > 0 ? 1u : 0u);
      if (a.D == 128) {
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_f16(d_tmem, ad0 + ((kTile * 128) >> 4) + 2 * k, bd0 + ((NMMA * 128) >> 4) + 2 * k, idesc, 1u);
      }
      umma_commit(bar_tfull(i));
    }
  }

  // ---------------- pass 1 (warps 0-3, thread = tile row): exact column max, then column sum ----------------
  const int row = (warp & 3) * 32 + lane;
  const uint32_t my_tmem = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  auto load_tile = [&](int i, float* s, bool& own, int& key) {   // masked, unscaled logits of tile i
    tmem_load_row<N>(my_tmem + i * NMMA, s);
    tmem_wait_ld();
    const int t = crank + i * kSfCluster;
    key = t * a.S - a.half + row;
    own = (row >= a.half) && (row < a.half + a.S) && (key < a.L);
    const int kp = key - a.P;                        // window key index: window row r sees window keys <= r
    if (t * a.S - a.half + kTile > a.P) {
#pragma unroll
      for (int n = 0; n < N; ++n) s[n] = (kp > (n % W)) ? -INFINITY : s[n];
    }
  };
  if (warp < 4) {
    float m[N];
#pragma unroll
    for (int n = 0; n < N; ++n) m[n] = kEmptyMax;
    for (int i = 0; i < T; ++i) {
      mbar_wait(bar_tfull(i), 0);
      tc_fence_after();
      if (i == 0) SF_TRACE(3);
      float s[N]; bool own; int key;
      load_tile(i, s, own, key);
#pragma unroll
      for (int n = 0; n < N; ++n) m[n] = fmaxf(m[n], warp_max_f32(own ? s[n] * a.c : -INFINITY));
    }
    if (lane == 0) {
#pragma unroll
      for (int n = 0; n < N; ++n) wm[warp * N + n] = m[n];
    }
    named_bar(1, 128);
    if (tid < N) cmax[tid] = fmaxf(fmaxf(wm[tid], wm[N + tid]), fmaxf(wm[2 * N + tid], wm[3 * N + tid]));
    named_bar(1, 128);
#pragma unroll
    for (int n = 0; n < N; ++n) m[n] = cmax[n];
    float z[N];
#pragma unroll
    for (int n = 0; n < N; ++n) z[n] = 0.f;
    for (int i = 0; i < T; ++i) {
      float s[N]; bool own; int key;
      load_tile(i, s, own, key);
      if (own) {
#pragma unroll
        for (int n = 0; n < N; ++n) z[n] += ex2(fmaf(s[n], a.c, -m[n]));
      }
    }
    // column sums across the warp (halving butterfly: lane l ends with column l), then across the 4 warps
    {
      float tz[N];
#pragma unroll
      for (int n = 0; n < N; ++n) tz[n] = z[n];
#pragma unroll
      for (int h = N / 2; h >= 1; h >>= 1) {
        const bool up = (lane & h) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
          const float keep = up ? tz[i + h] : tz[i];
          const float send = up ? tz[i] : tz[i + h];
          tz[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
        }
      }
#pragma unroll
      for (int o = 16; o >= N; o >>= 1) tz[0] += __shfl_xor_sync(0xffffffffu, tz[0], o);
      named_bar(1, 128);                             // wm is free again
      if (lane < N) wm[warp * N + lane] = tz[0];
    }
    named_bar(1, 128);
    if (tid < N) part[crank * N + tid] = make_float2(cmax[tid], (wm[tid] + wm[N + tid]) + (wm[2 * N + tid] + wm[3 * N + tid]));
  }
  __syncthreads();
  // my (max, sum) per column -> every CTA of the cluster
  SF_TRACE(4);
  for (int i = tid; i < kSfCluster * N; i += kSfThreads) {
    const int dst = i / N, n = i - dst * N;
    if (dst != crank) cluster.map_shared_rank(part, dst)[crank * N + n] = part[crank * N + n];
  }
  cluster.sync();                                    // #2: all partials everywhere
  SF_TRACE(5);
  if (tid < N) {
    float M = kEmptyMax;
#pragma unroll
    for (int c = 0; c < kSfCluster; ++c) M = fmaxf(M, part[c * N + tid].x);
    float Z = 0.f;
#pragma unroll
    for (int c = 0; c < kSfCluster; ++c) Z += part[c * N + tid].y * ex2(part[c * N + tid].x - M);
    colstat[tid] = make_float2(M, 1.0f / Z);
  }
  __syncthreads();

  // ---------------- pass 2: votes -> pool -> group sum -> score; composites to the leader ------------------
  unsigned long long* lead_cand = cluster.map_shared_rank(cand, 0);
  if (warp < 4) {
    float mn[N], zi[N];
#pragma unroll
    for (int n = 0; n < N; ++n) { const float2 v = colstat[n]; mn[n] = v.x; zi[n] = v.y; }
    const float pad = (a.pooling == FKV_POOL_MAX) ? -INFINITY : 0.f;
    const float inv_ks = 1.0f / (float)a.ks;
    for (int i = 0; i < T; ++i) {
      float s[N]; bool own1; int key;
      load_tile(i, s, own1, key);
      const bool in_prefix = (key >= 0) && (key < a.P);
      float* vb = vbuf + ((i & 1) * G) * kTile;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float vote = 0.f;
#pragma unroll
        for (int r = 0; r < W; ++r)                  // rows summed in order (utils.py:104)
          vote = fmaf(ex2(fmaf(s[g * W + r], a.c, -mn[g * W + r])), zi[g * W + r], vote);
        vb[g * kTile + row] = in_prefix ? vote : pad;
      }
      named_bar(1, 128);
      if ((row >= a.half) && (row < a.half + a.S) && in_prefix) {
        float acc = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float* vr = vb + g * kTile + row;
          float pooled;
          if (a.pooling == FKV_POOL_MAX) {
            pooled = -INFINITY;
            for (int d = -a.half; d <= a.half; ++d) pooled = fmaxf(pooled, vr[d]);
          } else {
            float sum = 0.f;
            if (a.half == 3) {
#pragma unroll
              for (int d = -3; d <= 3; ++d) sum += vr[d];
            } else {
              for (int d = -a.half; d <= a.half; ++d) sum += vr[d];
            }
            pooled = sum * inv_ks;                    // count_include_pad: always / kernel_size
          }
          acc += pooled;                              // GQA group-sum, g in order (utils.py:112)
        }
        a.score[(int64_t)b * a.score_sb + (int64_t)(kvh - a.out_h0) * a.score_sh + key] = acc;
        lead_cand[key] = ((unsigned long long)sf_ordered_key(acc) << 32) | (uint32_t)(0xffffffffu - (uint32_t)key);
      }
    }
  }
  SF_TRACE(6);
  cluster.sync();                                    // #3: the leader holds every score of the pair
  SF_TRACE(7);

  // ---------------- selection in the leader (value desc, index asc) ------------------------------------------
  // Counting scatter on 4096 ADAPTIVE bins (the pair's key range mapped linearly onto the bins, ~0.5 keys per bin),
  // then every key counts the larger keys of its own bin: rank = keys in higher bins + larger keys in the bin.
  // The K tiles are dead by now (all MMAs retired before pass 2 ended): their shared memory holds the scratch.
  // Rows with a crowded bin (exact ties) take a plain bitonic sort instead.
  int* sorted_idx = reinterpret_cast<int*>(sm);                              // [n] source row per rank (leader, read by all)
  if (crank == 0) {
    unsigned long long* grouped = reinterpret_cast<unsigned long long*>(sm + 16384);   // [n] keys grouped by bin
    uint32_t* bbase = reinterpret_cast<uint32_t*>(sm + 16384 + 32768);                  // [4096] keys in larger bins
    uint32_t* wtot = bbase + 4096;                                                      // [16] scan / reduction scratch
    const int n = a.P;
    // key range
    uint32_t lo = 0xffffffffu, hi = 0u;
    constexpr int UB = 4;                              // independent elements per thread and round (latency, not work, bounds this)
    for (int i0 = tid; i0 < n; i0 += kSfThreads * UB) {
      uint32_t key[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) { const int i = i0 + u * kSfThreads; key[u] = (i < n) ? (uint32_t)(cand[i] >> 32) : 0u; }
#pragma unroll
      for (int u = 0; u < UB; ++u)
        if (i0 + u * kSfThreads < n) { lo = min(lo, key[u]); hi = max(hi, key[u]); }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if (lane == 0) { wtot[warp] = lo; wtot[32 + warp] = hi; }
    __syncthreads();
    lo = wtot[0]; hi = wtot[32];
#pragma unroll
    for (int w8 = 1; w8 < kSfThreads / 32; ++w8) { lo = min(lo, wtot[w8]); hi = max(hi, wtot[32 + w8]); }
    SF_TRACE(12);
    const int range_bits = 32 - __clz(hi - lo);       // bin = (key - lo) >> shift < 4096 (2048..4096 bins in use)
    const int bin_shift = range_bits > 12 ? range_bits - 12 : 0;
    auto rbin = [&](uint32_t key) { return 4095u - ((key - lo) >> bin_shift); };   // 0 = largest keys
    for (int i0 = tid; i0 < n; i0 += kSfThreads * UB) {
      uint32_t bn[UB];
#pragma unroll
      for (int u = 0; u < UB; ++u) { const int i = i0 + u * kSfThreads; bn[u] = (i < n) ? rbin((uint32_t)(cand[i] >> 32)) : 0xffffffffu; }
#pragma unroll
      for (int u = 0; u < UB; ++u)
        if (bn[u] != 0xffffffffu) atomicAdd(&cursor[bn[u]], 1u);
    }
    __syncthreads();
    SF_TRACE(13);
    // exclusive scan over the 4096 bins: kSfWorkers threads x BPT bins
    constexpr int BPT = 4096 / kSfWorkers;
    uint32_t loc[BPT], ls = 0, mx = 0;
    if (tid < kSfWorkers) {
#pragma unroll
      for (int i = 0; i < BPT; ++i) { loc[i] = cursor[tid * BPT + i]; ls += loc[i]; mx = max(mx, loc[i]); }
    }
    const bool skew = __syncthreads_or(mx > 64u);
    if (!skew) {
      uint32_t inc = ls;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
      if (tid < kSfWorkers && lane == 31) wtot[64 + warp] = inc;
      __syncthreads();
      if (tid < kSfWorkers) {
        uint32_t above = inc - ls;
        for (int w8 = 0; w8 < warp; ++w8) above += wtot[64 + w8];
#pragma unroll
        for (int i = 0; i < BPT; ++i) { bbase[tid * BPT + i] = above; above += loc[i]; cursor[tid * BPT + i] = 0; }
      }
      __syncthreads();
      SF_TRACE(14);
      for (int i0 = tid; i0 < n; i0 += kSfThreads * UB) {
        unsigned long long c[UB];
        uint32_t bn[UB], pos[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) { const int i = i0 + u * kSfThreads; c[u] = (i < n) ? cand[i] : 0ull; bn[u] = (i < n) ? rbin((uint32_t)(c[u] >> 32)) : 0xffffffffu; }
#pragma unroll
        for (int u = 0; u < UB; ++u)
          if (bn[u] != 0xffffffffu) pos[u] = bbase[bn[u]] + atomicAdd(&cursor[bn[u]], 1u);
#pragma unroll
        for (int u = 0; u < UB; ++u)
          if (bn[u] != 0xffffffffu) grouped[pos[u]] = c[u];
      }
      __syncthreads();
      SF_TRACE(15);
      for (int i0 = tid; i0 < n; i0 += kSfThreads * UB) {
        unsigned long long c[UB];
        uint32_t s0[UB], s1[UB];
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          const int i = i0 + u * kSfThreads;
          c[u] = (i < n) ? grouped[i] : 0ull;
          const uint32_t bn = rbin((uint32_t)(c[u] >> 32));
          s0[u] = (i < n) ? bbase[bn] : 0u;
          s1[u] = (i < n) ? s0[u] + cursor[bn] : 0u;
        }
#pragma unroll
        for (int u = 0; u < UB; ++u) {
          uint32_t rank = s0[u];
          for (uint32_t q = s0[u]; q < s1[u]; ++q) rank += (grouped[q] > c[u]);
          if (i0 + u * kSfThreads < n) sorted_idx[rank] = (int)(0xffffffffu - (uint32_t)(c[u] & 0xffffffffull));
        }
      }
    } else {
      for (int size = 2; size <= a.sort_cap; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
          for (int t = tid; t < a.sort_cap / 2; t += kSfThreads) {
            const int l2 = 2 * t - (t & (stride - 1));
            const int h2 = l2 + stride;
            const bool desc = ((l2 & size) == 0);
            const unsigned long long va = cand[l2], vb2 = cand[h2];
            if ((va < vb2) == desc) { cand[l2] = vb2; cand[h2] = va; }
          }
          __syncthreads();
        }
      }
      for (int i = tid; i < n; i += kSfThreads) sorted_idx[i] = (int)(0xffffffffu - (uint32_t)(cand[i] & 0xffffffffull));
    }
    __syncthreads();
    int64_t* out = a.idx_out + (int64_t)b * a.idx_sb + (int64_t)(kvh - a.out_h0) * a.idx_sh;
    for (int r = tid; r < a.n_sel; r += kSfThreads) out[r] = (int64_t)sorted_idx[r];
  }
  SF_TRACE(8);
  cluster.sync();                                    // #4: sorted list ready
  SF_TRACE(9);

  // ---------------- compaction: every CTA copies a contiguous share of the C rows (utils.py:114-121) -------
  const int C = a.n_sel + a.W;
  const int per = (C + kSfCluster - 1) / kSfCluster;
  const int r0 = crank * per, r1 = min(C, r0 + per);
  const int* lead_sorted = cluster.map_shared_rank(sorted_idx, 0);
  for (int r = r0 + tid; r < r1; r += kSfThreads)
    sel[r - r0] = (r < a.n_sel) ? lead_sorted[r] : (a.L - a.W + (r - a.n_sel));
  cluster.sync();                                    // #5: nobody reads the leader's smem after this point
  SF_TRACE(10);
  if (tid < kSfWorkers) {
    const int lane_chunk = tid % a.row_chunks, rows_par = kSfWorkers / a.row_chunks;
    const char* kb = a.k + b * a.k_sb + kvh * a.k_sh;
    const char* vbp = a.v + b * a.v_sb + kvh * a.v_sh;
    char* kob = a.ko + b * a.ko_sb + (kvh - a.out_h0) * a.ko_sh;
    char* vob = a.vo + b * a.vo_sb + (kvh - a.out_h0) * a.vo_sh;
    for (int rb = r0 + tid / a.row_chunks; rb < r1; rb += rows_par * 4) {
      int64_t src_row[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int r = rb + u * rows_par; src_row[u] = (r < r1) ? (int64_t)sel[r - r0] : -1; }
      uint4 kv[4], vv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (src_row[u] >= 0) {
          kv[u] = ld_stream16(kb + src_row[u] * a.k_sl + lane_chunk * 16);
          vv[u] = ld_stream16(vbp + src_row[u] * a.v_sl + lane_chunk * 16);
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (src_row[u] >= 0) {
          const int r = rb + u * rows_par;
          st_stream16(kob + (int64_t)r * a.ko_sl + lane_chunk * 16, kv[u]);
          st_stream16(vob + (int64_t)r * a.vo_sl + lane_chunk * 16, vv[u]);
        }
    }
  }
  SF_TRACE(11);
  tc_fence_before();
  __syncthreads();
  if (warp == kSfIssueWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(128));
  }
}

struct SfPlan { int G, W, N, nkv, kv0, n_tiles, S, half, sort_cap; size_t smem; bool ok; };

SfPlan make_sf_plan(const fkv_compress_params& cp) {
  const fkv_score_params& p = cp.score;
  SfPlan pl{};
  pl.ok = false;
  // Opt-in (FKV_SMALL_FUSE=1, read per call so tests can toggle it): measured on B200 it only ties the three-kernel
  // path (18.8 us vs ~21 us per 2048-token layer including launch gaps, step 1.092 vs 1.098 ms) because the
  // single-CTA ranking of the 2040 scores (5.5 us) and 5 cluster barriers eat what the saved launches give.
  const char* on = getenv("FKV_SMALL_FUSE");
  if (!on || on[0] != '1' || p.mode != FKV_MODE_FASTKV || cp.tsp_k > 0 || p.impl == FKV_IMPL_SIMPLE) return pl;
  if (p.dtype != FKV_BF16 && p.dtype != FKV_F16) return pl;
  if (!(p.D == 64 || p.D == 128)) return pl;
  pl.G = p.Hq / p.Hkv; pl.W = p.window_size; pl.N = pl.G * pl.W;
  if (pl.W != 8 || !(pl.G == 1 || pl.G == 2 || pl.G == 4)) return pl;
  pl.half = p.kernel_size / 2;
  pl.S = kTile - 2 * pl.half;
  if (pl.S < 32) return pl;
  const int P = p.L - pl.W;
  if (P < 1 || P > kSfMaxKeys) return pl;
  pl.n_tiles = ceil_div(p.L, pl.S);
  if (pl.n_tiles > kSfCluster * kSfTiles) return pl;
  pl.kv0 = p.h0; pl.nkv = p.h1 - p.h0;
  if (pl.nkv < 1) return pl;
  int cap = 2;
  while (cap < P) cap <<= 1;
  pl.sort_cap = cap;
  const int NMMA = pl.N < 16 ? 16 : pl.N;
  const size_t tile_bytes = (size_t)kTile * p.D * 2;
  const size_t q_bytes = ((size_t)NMMA * p.D * 2 + 1023) & ~(size_t)1023;
  pl.smem = 1024 + kSfTiles * tile_bytes + q_bytes + (size_t)kSfMaxKeys * 8 + (size_t)(kSfCluster + 1) * pl.N * 8 +
            (size_t)5 * pl.N * 4 + (size_t)2 * pl.G * kTile * 4 + (kSfMaxKeys / kSfCluster + 64) * 4 + 4096 * 4 + 9 * 8 + 16 + 256;
  pl.ok = encode_fn() != nullptr;
  return pl;
}

template <int W, int G>
int launch_sf_inst(const CUtensorMap& mk, const CUtensorMap& mq, const SfArgs& a, int n_ctas, size_t smem, cudaStream_t s) {
  { const int rc_attr = ensure_dyn_smem(small_compress_kernel<W, G>, 227 * 1024); if (rc_attr != FKV_OK) return rc_attr; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_ctas);
  cfg.blockDim = dim3(kSfThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kSfCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  FKV_CUDA_TRY(cudaLaunchKernelEx(&cfg, small_compress_kernel<W, G>, mk, mq, a));
  count_launch();
  return FKV_OK;
}

}  // namespace

bool small_compress_supported(const fkv_compress_params& cp) { return make_sf_plan(cp).ok; }

// score -> top-k -> compaction of a short-context call in one cluster kernel (see small_compress_kernel)
int launch_small_compress(const fkv_compress_params& cp, cudaStream_t s) {
  const fkv_score_params& p = cp.score;
  const SfPlan pl = make_sf_plan(cp);
  if (!pl.ok) return FKV_EMODE;
  const int es = dtype_size(p.dtype);
  SfArgs a{};
  a.L = p.L; a.P = p.L - pl.W; a.D = p.D; a.nkv = pl.nkv; a.kv0 = pl.kv0; a.n_tiles = pl.n_tiles; a.S = pl.S; a.half = pl.half;
  a.ks = p.kernel_size; a.pooling = p.pooling; a.is_f16 = (p.dtype == FKV_F16); a.q_heads_per_kv = pl.G;
  a.n_sel = cp.capacity - pl.W; a.W = pl.W; a.sort_cap = pl.sort_cap;
  a.c = (1.0f / sqrtf((float)p.D)) * 1.4426950408889634f;
  a.score = p.score; a.score_sb = p.score_stride_b; a.score_sh = p.score_stride_h; a.out_h0 = p.h0;
  a.idx_out = cp.idx_out; a.idx_sb = (int64_t)(p.h1 - p.h0) * a.n_sel; a.idx_sh = a.n_sel;
  a.k = (const char*)p.k.ptr; a.v = (const char*)cp.v.ptr;
  a.ko = (char*)const_cast<void*>(cp.k_out.ptr); a.vo = (char*)const_cast<void*>(cp.v_out.ptr);
  a.k_sb = p.k.stride_b * es; a.k_sh = p.k.stride_h * es; a.k_sl = p.k.stride_l * es;
  a.v_sb = cp.v.stride_b * es; a.v_sh = cp.v.stride_h * es; a.v_sl = cp.v.stride_l * es;
  a.ko_sb = cp.k_out.stride_b * es; a.ko_sh = cp.k_out.stride_h * es; a.ko_sl = cp.k_out.stride_l * es;
  a.vo_sb = cp.v_out.stride_b * es; a.vo_sh = cp.v_out.stride_h * es; a.vo_sl = cp.v_out.stride_l * es;
  a.row_chunks = p.D * es / 16;
  CUtensorMap mk, mq;
  int rc;
  if ((rc = make_map(&mk, p.k, p.dtype, p.B, p.Hkv, p.L, p.D, kTile, &a.kpos_l, &a.kpos_h, &a.kpos_b)) != FKV_OK) return rc;
  if ((rc = make_map(&mq, p.q, p.dtype, p.B, p.Hq, p.L, p.D, pl.W, &a.qpos_l, &a.qpos_h, &a.qpos_b)) != FKV_OK) return rc;
  const int n_ctas = p.B * pl.nkv * kSfCluster;
  static const bool trace_on = getenv("FKV_SF_TRACE") != nullptr;
  static unsigned long long* trace_buf = nullptr;
  if (trace_on) {
    if (!trace_buf) FKV_CUDA_TRY(cudaMalloc(&trace_buf, 16 * sizeof(unsigned long long)));
    a.trace = trace_buf;
  }
  if (pl.G == 1) rc = launch_sf_inst<8, 1>(mk, mq, a, n_ctas, pl.smem, s);
  else if (pl.G == 2) rc = launch_sf_inst<8, 2>(mk, mq, a, n_ctas, pl.smem, s);
  else rc = launch_sf_inst<8, 4>(mk, mq, a, n_ctas, pl.smem, s);
  if (trace_on && rc == FKV_OK) {                       // debugging aid: synchronises
    cudaStreamSynchronize(s);
    unsigned long long h[16];
    cudaMemcpy(h, trace_buf, sizeof(h), cudaMemcpyDeviceToHost);
    const char* nm[12] = {"start", "setup", "csync1", "tile0", "p1_done", "csync2", "p2_done", "csync3", "sorted", "csync4", "csync5", "copied"};
    fprintf(stderr, "[sf trace] L=%d:", p.L);
    for (int i = 1; i < 12; ++i) fprintf(stderr, " %s %.2f", nm[i], (double)(h[i] - h[0]) / 1e3);
    fprintf(stderr, " | minmax %.2f hist %.2f scan %.2f scatter %.2f", (double)(h[12] - h[0]) / 1e3, (double)(h[13] - h[0]) / 1e3, (double)(h[14] - h[0]) / 1e3, (double)(h[15] - h[0]) / 1e3);
    fprintf(stderr, "\n");
  }
  return rc;
}

}  // namespace fkv
