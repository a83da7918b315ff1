// K1/K2 tensor-core path, hand-written for sm_100a (TMA -> shared memory -> tcgen05.mma -> TMEM -> registers):
// observation-window scoring + pooling + GQA group-sum (baselines/fastkv/utils.py:93-112 and the SnapKV twin) as one
// persistent cooperative kernel that reads K exactly once.  Sibling kernels on the same plumbing (tc_common.cuh):
// gemfilter_tc.cu (GemFilter scores, single pass) and small_compress.cu (opt-in one-kernel short-context path).
//
// score_tc_kernel
// Work unit  = (batch b, kv head h, key tile t).  A tile is 128 consecutive key rows (UMMA M) that
//              start `half = kernel_size/2` rows before the S = 128 - 2*half keys it OWNS, so the
//              pooling halo is recomputed instead of exchanged between CTAs.
// MMA        = logits[128 keys x N] = K_tile[128 x D] . Qwin[N x D]^T, N = G*W <= 32 window-rows of the
//              query heads of that kv head (wider groups: one launch per head sub-group), bf16/f16
//              operands, fp32 accumulators in TMEM.  All 512 TMEM columns are a cache of NB = 512/N logit tiles.
// Phases     = the (b, h) pairs are processed in groups whose tiles fit that cache on every CTA, so the
//              logits of a tile stay on chip between its two passes:
//   pass 1   per row (max, sum exp2) over the owned keys (lazy per-warp rescaling), flushed per (b, h)
//            pair as epoch-tagged records; a CTA then assembles the final statistics of its pairs from the
//            records of the CTAs that share them (no grid barrier: records validate themselves)
//   pass 2   p = exp2(x - max) / sum from the logits in TMEM -> votes per query head -> pool1d over the
//            halo'd tile -> group-sum -> score[b, h, j]; reading a tile frees its TMEM block for the next phase
// Warp roles (512 threads = 4 aligned warpgroups): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator (both stream
// through all phases without stopping), warps 2-3 idle, warps 4-15 = three epilogue warpgroups taking tiles round-robin;
// setmaxnreg moves registers from the control warpgroup (80 per thread) to the epilogue (144 per thread).
#include "tc_common.cuh"

namespace fkv {
namespace {

struct TcArgs {
  int B, L, P, D, nkv, kv0, n_tiles, S, half, ks, pooling, group_sum, is_f16;
  int U, maxlp, nst, maxc;
  int npairs, ppp, nph;                                 // (b, h) pairs, pairs per phase, phases
  int kpos_l, kpos_h, kpos_b, qpos_l, qpos_h, qpos_b;   // coordinate slots of the stride-sorted TMA dims
  int q_heads_per_kv;                                   // model G (query head = kvh * G + g0 + g)
  int g0, accumulate;                                   // head sub-group of this launch; add to the score already there
  float c;                                              // softmax scale * log2(e)
  unsigned long long* part;                             // [pair][contributor rank][N][2] epoch-tagged (max | sum) words
  unsigned int* sync_counter;                           // workspace header: [1] departures (left zero), u64 at [2] = launch epoch
  unsigned long long* trace;                            // optional [grid][16] globaltimer stamps (FKV_TC_TRACE=1)
  float* score; int64_t score_sb, score_sh; int out_h0;
  unsigned int* hist;                                   // optional [B * hist_hpb][4096] 12-bit histogram of the stored scores
  int hist_hpb;                                         // selection heads per batch element in that table
};

// order-preserving key of a score (same map as topk_common.cuh) -> its 12-bit histogram bin
__device__ __forceinline__ uint32_t score_bin(float x) {
  x += 0.0f;
  const uint32_t u = __float_as_uint(x);
  return ((u & 0x80000000u) ? ~u : (u | 0x80000000u)) >> 20;
}
// one RED per distinct bin of the warp: pooled scores of neighbouring keys mostly share a 1/8-octave bin
__device__ __forceinline__ void hist_red(unsigned int* hrow, float score, bool valid) {
  const uint32_t bin = valid ? score_bin(score) : 0xffffffffu;
  const uint32_t peers = __match_any_sync(0xffffffffu, bin);
  if (valid && (int)(threadIdx.x & 31) == __ffs(peers) - 1)
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(hrow + bin), "r"((uint32_t)__popc(peers)) : "memory");
}

// kDef: the reference's default epilogue (avgpool, kernel_size 7, FastKV group-sum, single head group) compiled
// without the other variants, so the pass-2 loop stays small and contiguous in the instruction cache
template <int W, int G, bool kDef>
__global__ void __launch_bounds__(kTcThreads, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_q, const TcArgs a) {
  constexpr int N = W * G;                          // logit rows per (b, kv head)
  constexpr int NMMA = N < 16 ? 16 : N;             // UMMA N (multiple of 16 for M = 128)
  constexpr int NB = 512 / NMMA;                    // accumulator blocks: ALL of TMEM is a logits cache
  static_assert(N <= 32 && (N & 7) == 0, "unsupported window x group size");

  extern __shared__ unsigned char smem_dyn[];
  if (threadIdx.x == kCtrlThreads) FKV_TRACE(0);
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (base - smem_u32(smem_dyn));

  const int tile_bytes = kTile * a.D * 2;
  const int q_bytes = NMMA * a.D * 2;
  // carve-up (all operand regions 1024-byte aligned for the 128B swizzle)
  const uint32_t k_off = 0;
  const uint32_t q_off = k_off + a.nst * tile_bytes;
  const uint32_t misc_off = q_off + 2 * ((q_bytes + 1023) & ~1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + misc_off);   // full[6] empty[6] qfull[2] qempty[2] tfull[32] tempty[32]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 80);
  float2* stage = reinterpret_cast<float2*>(tmem_slot + 4);             // [epilogue warps][N]
  float2* colstat = stage + kEpiWarps * N;                              // [maxlp][N]
  float* vbuf = reinterpret_cast<float*>(colstat + kMaxLocalPairs * N); // [wg][2 buf][G][128]
  int* ncontrib = reinterpret_cast<int*>(vbuf + kNumWG * 2 * G * kTile);     // [maxlp] CTAs contributing to each local pair
  int* ph_u0 = ncontrib + kMaxLocalPairs;                               // [nph] first unit of this CTA's slice of the phase
  int* ph_T = ph_u0 + kMaxPhases;                                       // [nph] tiles of this CTA in the phase
  // pass 1 leaves exp2(x*c - ref) in TMEM instead of the logits; ref = the owning warp's running reference at that time
  float* tile_ref = reinterpret_cast<float*>(ph_T + kMaxPhases);        // [NB][4 warps][N] references a tile's values are relative to
  unsigned int* tile_ver = reinterpret_cast<unsigned int*>(tile_ref + NB * 4 * N);   // [NB][4] version of those references
  float* cstage = reinterpret_cast<float*>(tile_ver + NB * 4);          // [epilogue warps][N] coefficient staging (pass 2)

  auto bar_full = [&](int s) { return smem_u32(bars + s); };
  auto bar_empty = [&](int s) { return smem_u32(bars + 6 + s); };
  auto bar_qfull = [&](int s) { return smem_u32(bars + 12 + s); };
  auto bar_qempty = [&](int s) { return smem_u32(bars + 14 + s); };
  auto bar_tfull = [&](int s) { return smem_u32(bars + 16 + s); };
  auto bar_tempty = [&](int s) { return smem_u32(bars + 48 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = gridDim.x, cta = blockIdx.x;
  const int nph = a.nph;
  // Phase ph covers the (b, h) pairs [ph * ppp, ...); its tiles are split evenly over the grid, at most
  // NB per CTA, so a tile's logits stay in TMEM between its two passes and K is read exactly once.
  for (int ph = threadIdx.x; ph < nph; ph += kTcThreads) {
    const int p0 = ph * a.ppp;
    const int Uph = min(a.ppp, a.npairs - p0) * a.n_tiles;
    const int r0 = unit_begin(cta, Uph, grid), r1 = unit_begin(cta + 1, Uph, grid);
    ph_u0[ph] = p0 * a.n_tiles + r0;
    ph_T[ph] = r1 - r0;
  }

  if (warp == 0) {                                  // 80 barriers, a few per lane
    for (int s = lane; s < kMaxStages; s += 32) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    if (lane < 2) { mbar_init(bar_qfull(lane), 1); mbar_init(bar_qempty(lane), 1); }
    for (int s = lane; s < NB; s += 32) { mbar_init(bar_tfull(s), 1); mbar_init(bar_tempty(s), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tm_q) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_enter();                                      // nothing global is touched before this point
  if (threadIdx.x == kCtrlThreads) FKV_TRACE(1);
  // register re-allocation between the aligned warpgroups: the control warpgroup keeps 80 registers per thread, the
  // three epilogue warpgroups get 144 (the 128-register cap of a 512-thread block left the epilogue spilling)
  // (each setmaxnreg sits at the top of the region it governs, like CUTLASS' warpgroup_reg_dealloc / _alloc)

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kCtrlRegs));
  if (warp == 0) {
    // ===================================== TMA producer =====================================
    // One thread; everything in the loop is incremental (no divisions, no local arrays): at
    // ~5 cycles per dependent instruction a fat loop body would pace the whole pipeline.
    if (lane == 0) {
      const int kl1 = a.kpos_l == 1, kl2 = a.kpos_l == 2, kl3 = a.kpos_l == 3;
      const int kh1 = a.kpos_h == 1, kh2 = a.kpos_h == 2, kh3 = a.kpos_h == 3;
      const int kb1 = a.kpos_b == 1, kb2 = a.kpos_b == 2, kb3 = a.kpos_b == 3;
      const int ql1 = a.qpos_l == 1, ql2 = a.qpos_l == 2, ql3 = a.qpos_l == 3;
      const int qh1 = a.qpos_h == 1, qh2 = a.qpos_h == 2, qh3 = a.qpos_h == 3;
      const int qb1 = a.qpos_b == 1, qb2 = a.qpos_b == 2, qb3 = a.qpos_b == 3;
      const int slabs = a.D >> 6;
      int cur_pair = -1, qu = -1;
      int hb1 = 0, hb2 = 0, hb3 = 0;                 // head/batch part of the K coordinates
      int slot = 0, ephase = 1;                      // ring slot + parity for the empty-barrier wait
      int m = 0;                                     // tiles issued so far
      for (int ph = 0; ph < nph; ++ph) {
        const int T = ph_T[ph];
        if (T == 0) continue;
        int pair = ph_u0[ph] / a.n_tiles, t = ph_u0[ph] - pair * a.n_tiles;
        for (int i = 0; i < T; ++i, ++m) {
          const bool new_pair = (pair != cur_pair);
          int pb = 0, pkvh = 0;
          if (new_pair) {
            cur_pair = pair; ++qu;
            pb = pair / a.nkv; pkvh = a.kv0 + pair - pb * a.nkv;
            hb1 = pkvh * kh1 + pb * kb1; hb2 = pkvh * kh2 + pb * kb2; hb3 = pkvh * kh3 + pb * kb3;
          }
          if (m >= a.nst) mbar_wait(bar_empty(slot), ephase);
          FKV_TRACE2(0, m);
          mbar_expect_tx(bar_full(slot), tile_bytes);
          const uint32_t kdst = base + k_off + slot * tile_bytes;
          const int row0 = t * a.S - a.half;         // may be negative / run past L: TMA zero-fills
          const int c1 = row0 * kl1 + hb1, c2 = row0 * kl2 + hb2, c3 = row0 * kl3 + hb3;
          tma_load_4d(kdst, &tm_k, bar_full(slot), 0, c1, c2, c3);
          if (slabs == 2) tma_load_4d(kdst + kTile * 128, &tm_k, bar_full(slot), 64, c1, c2, c3);
          if (new_pair) {                            // new (b, h): stage its N query rows (B operand) after the K tile
            const int qb = qu & 1;
            if (qu >= 2) mbar_wait(bar_qempty(qb), ((qu >> 1) - 1) & 1);
            mbar_expect_tx(bar_qfull(qb), N * a.D * 2);
            const uint32_t qdst = base + q_off + qb * ((q_bytes + 1023) & ~1023);
            for (int slab = 0; slab < slabs; ++slab)
#pragma unroll
              for (int g = 0; g < G; ++g) {
                const int hq = pkvh * a.q_heads_per_kv + a.g0 + g;
                tma_load_4d(qdst + slab * (NMMA * 128) + g * (W * 128), &tm_q, bar_qfull(qb), slab * 64,
                            a.P * ql1 + hq * qh1 + pb * qb1, a.P * ql2 + hq * qh2 + pb * qb2, a.P * ql3 + hq * qh3 + pb * qb3);
              }
          }
          if (++slot == a.nst) { slot = 0; ephase ^= 1; }
          if (++t == a.n_tiles) { t = 0; ++pair; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((a.is_f16 ? 0u : 1u) << 7) | ((a.is_f16 ? 0u : 1u) << 10) |
                             ((uint32_t)(NMMA >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
      const uint64_t desc_hi = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      const bool two_slabs = (a.D == 128);
      int cur_pair = -1, qu = -1;
      int slot = 0, fphase = 0;
      uint32_t qa = 0;
      int m = 0;
      for (int ph = 0; ph < nph; ++ph) {
        const int T = ph_T[ph];
        if (T == 0) continue;
        int pair = ph_u0[ph] / a.n_tiles, t = ph_u0[ph] - pair * a.n_tiles;
        for (int i = 0; i < T; ++i, ++m) {
          if (pair != cur_pair) {
            if (qu >= 0) umma_commit(bar_qempty(qu & 1));        // the previous pair's Q buffer is free once its MMAs retire
            cur_pair = pair; ++qu;
            mbar_wait(bar_qfull(qu & 1), (qu >> 1) & 1);
            qa = base + q_off + (qu & 1) * ((q_bytes + 1023) & ~1023);
          }
          const int blk = m & (NB - 1);
          mbar_wait(bar_full(slot), fphase);
          FKV_TRACE2(1, m);
          // the block is free once the tile that used it before (m - NB) went through its SECOND pass;
          // tempty[blk] completes one phase per such read, so the waiter is never more than one phase behind
          if (m >= NB) mbar_wait(bar_tempty(blk), ((m / NB) - 1) & 1);
          tc_fence_after();
          const uint32_t ka = base + k_off + slot * tile_bytes;
          const uint32_t d_tmem = tmem_base + blk * NMMA;
          const uint64_t ad0 = desc_hi | (uint64_t)((ka & 0x3ffffu) >> 4);
          const uint64_t bd0 = desc_hi | (uint64_t)((qa & 0x3ffffu) >> 4);
          // K advances 16 elements = 32 B (encoded +2) inside the 128B swizzle atom; the second
          // 64-wide slab of D starts one full operand slab later
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d_tmem, ad0 + 2 * k, bd0 + 2 * k, idesc, k > 0 ? 1u : 0u);
          if (two_slabs) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(d_tmem, ad0 + ((kTile * 128) >> 4) + 2 * k, bd0 + ((NMMA * 128) >> 4) + 2 * k, idesc, 1u);
          }
          umma_commit(bar_empty(slot));              // smem slot reusable once these MMAs retire
          umma_commit(bar_tfull(blk));               // accumulator ready for the epilogue
          FKV_TRACE2(2, m);
          if (++slot == a.nst) { slot = 0; fphase ^= 1; }
          if (++t == a.n_tiles) { t = 0; ++pair; }
        }
      }
    }
  }
  } else {
    // ===================================== epilogue =========================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kEpiRegs));
    const int et = threadIdx.x - kCtrlThreads;       // 0..383
    const int ew = et >> 5;                          // epilogue warp 0..7
    const int wg = ew >> 2;                          // warpgroup: tiles are dealt round-robin
    const int quarter = warp & 3;                    // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;             // tile row == TMEM lane
    const uint32_t my_tmem = tmem_base + ((uint32_t)(quarter * 32) << 16);
    // Per-warp running reference max (warp-uniform, replicated in every lane) and per-lane partial
    // sums, both in registers.  In pass 2 the same registers hold the final (max, 1/sum) per row.
    float mn[N], z[N];
#pragma unroll
    for (int n = 0; n < N; ++n) { mn[n] = kEmptyMax; z[n] = 0.f; }
    unsigned int ref_ver = 1u;                       // bumped whenever this warp's references change (pass 1)
    // The host side refuses workspaces that never went through fkv_workspace_init (FKV_EWORKSPACE).  Should the
    // header be overwritten afterwards, never abort the context across the ABI: flag it in the header's error word
    // (read back by fkv_workspace_status) and carry on; the launch stays self-consistent (all CTAs read one epoch).
    if (et == 0 && blockIdx.x == 0 && __ldcg(a.sync_counter + 4) != kWorkspaceMagic) a.sync_counter[5] = 1u;
    // launch epoch (bumped by the last CTA of every launch) -> the two record tags of this launch
    const unsigned long long epoch = __ldcg(reinterpret_cast<const unsigned long long*>(a.sync_counter + 2)) + 1ull;
    const unsigned int tag_m = 0x7fc00000u | (unsigned)(epoch & 0x3fffffu);
    const unsigned int tag_z = 0x7fc00000u | (unsigned)((epoch >> 22) & 0x3fffffu);

    const float pad = (a.pooling == FKV_POOL_MAX) ? -INFINITY : 0.f;
    const float inv_ks = 1.0f / (float)a.ks;
    int jbase = 0;                                   // CTA-wide index of the phase's first tile (TMEM block = index % NB)
    int vcount = 0;                                  // pass-2 tiles done by this warpgroup (vote double buffer)
    for (int ph = 0; ph < nph; ++ph) {
    const int T = ph_T[ph];                          // tiles of this CTA in the phase (<= NB)
    if (T == 0) continue;                            // uniform over the CTA
    const int u0 = ph_u0[ph];
    const int pair0 = u0 / a.n_tiles;
    const int nlp = (u0 + T - 1) / a.n_tiles - pair0 + 1;   // local (b, h) pairs touched
    const int p0 = ph * a.ppp;                       // first pair of the phase
    const int Uph = min(a.ppp, a.npairs - p0) * a.n_tiles;
#pragma unroll
    for (int n = 0; n < N; ++n) { mn[n] = kEmptyMax; z[n] = 0.f; }
    ++ref_ver;
    int flushed = 0;                                 // local pairs already flushed by this warpgroup
    auto flush_upto = [&](int lp_end) {              // flush local pairs [flushed, lp_end)
      for (; flushed < lp_end; ++flushed) {
        // Column sums across the warp by a halving butterfly: after the exchange with lane^16 a lane
        // keeps the half of the columns selected by its bit 4, then bit 3, ... so lane l ends up with
        // the total (and the reference max) of column l (N/32 groups of 32 columns): 31 shuffles
        // instead of 160, and little straight-line code (cold instruction fetch dominates here).
#pragma unroll
        for (int g0 = 0; g0 < N; g0 += 32) {
          constexpr int CW = N < 32 ? N : 32;         // columns in this group
          float tz[CW], tm[CW];
#pragma unroll
          for (int i = 0; i < CW; ++i) { tz[i] = z[g0 + i]; tm[i] = mn[g0 + i]; }
#pragma unroll
          for (int h = CW / 2; h >= 1; h >>= 1) {
            const bool up = (lane & h) != 0;
#pragma unroll
            for (int i = 0; i < h; ++i) {
              const float keep = up ? tz[i + h] : tz[i];
              const float send = up ? tz[i] : tz[i + h];
              tz[i] = keep + __shfl_xor_sync(0xffffffffu, send, h);
              tm[i] = up ? tm[i + h] : tm[i];
            }
          }
          // lanes >= CW (N < 32) hold partial duplicates: finish their reduction over the unused lane bits
#pragma unroll
          for (int o = 16; o >= CW; o >>= 1) tz[0] += __shfl_xor_sync(0xffffffffu, tz[0], o);
          if (lane < CW) stage[ew * N + g0 + lane] = make_float2(tm[0], tz[0]);
        }
#pragma unroll
        for (int n = 0; n < N; ++n) { mn[n] = kEmptyMax; z[n] = 0.f; }
        ++ref_ver;
        named_bar(2, kEpiThreads);
        if (et < N) {
          float M = kEmptyMax;
#pragma unroll
          for (int w8 = 0; w8 < kEpiWarps; ++w8) M = fmaxf(M, stage[w8 * N + et].x);
          float Z = 0.f;
#pragma unroll
          for (int w8 = 0; w8 < kEpiWarps; ++w8) {
            const float2 v = stage[w8 * N + et];
            Z += v.y * ex2(v.x - M);
          }
          const int pr = pair0 + flushed;
          const int krank = cta - cta_of_unit((pr - p0) * a.n_tiles, Uph, grid);
          st_record(a.part + (((size_t)pr * a.maxc + krank) * N + et) * 2, tag_word(M, tag_m), tag_word(Z, tag_z));
        }
        named_bar(2, kEpiThreads);
      }
    };

    // ---------------- pass 1: per-row softmax statistics over owned keys ----------------------
    int pair, t;
    { const int u = u0 + wg; pair = u / a.n_tiles; t = u - pair * a.n_tiles; }
    // (one extra trip past the last tile performs the phase's final flush: the flush code, 10 % of the kernel's
    // instructions, then exists once in the instruction stream)
    for (int j = wg; ; j += kNumWG) {
      const int blk = (jbase + j) & (NB - 1);
      flush_upto(j < T ? pair - pair0 : nlp);
      if (j >= T) break;
      mbar_wait(bar_tfull(blk), ((jbase + j) / NB) & 1);
      if (et == 0 && jbase + j == 0) FKV_TRACE(2);
      tc_fence_after();
      float s[N];
      tmem_load_row<N>(my_tmem + blk * NMMA, s);
      tmem_wait_ld();
      const int key = t * a.S - a.half + row;
      const bool valid = (key >= 0) && (key < a.L);  // a real key (own or pooling halo)
      const bool own = (row >= a.half) && (row < a.half + a.S) && (key < a.L);
      const bool has_window = (t * a.S - a.half + kTile > a.P);   // warp-uniform
      const int kp = key - a.P;                      // window key index (>= 0 inside the window)
      // window row r sees window keys <= r (utils.py:95-101): mask the others to -inf
      if (has_window) {
#pragma unroll
        for (int n = 0; n < N; ++n) s[n] = (kp > (n % W)) ? -INFINITY : s[n];
      }
      // cheap check first: does any key of the tile exceed its row's reference by more than tau?
      float dmax = -INFINITY;
#pragma unroll
      for (int n = 0; n < N; ++n) dmax = fmaxf(dmax, fmaf(s[n], a.c, -mn[n]));
      if (__any_sync(0xffffffffu, valid && dmax > kLazyTau)) {
        // rare (always on the first tile of a pair): raise the references to the warp-wide column max
#pragma unroll
        for (int n = 0; n < N; ++n) {
          const float nm = fmaxf(mn[n], warp_max_f32(valid ? s[n] * a.c : -INFINITY));
          z[n] *= ex2(mn[n] - nm);
          mn[n] = nm;
        }
        ++ref_ver;
      }
      // the tile's values go back to TMEM as e = exp2(x*c - ref): pass 2 is then one FMA per value, no MUFU
#pragma unroll
      for (int n = 0; n < N; ++n) s[n] = valid ? ex2(fmaf(s[n], a.c, -mn[n])) : 0.f;
      tmem_store_row<N>(my_tmem + blk * NMMA, s);
      if (own) {
#pragma unroll
        for (int n = 0; n < N; ++n) z[n] += s[n];
      }
      if (lane == 0) {                               // the references are warp-uniform: one lane records them for pass 2
        float4* dst = reinterpret_cast<float4*>(tile_ref + (blk * 4 + quarter) * N);
#pragma unroll
        for (int n = 0; n < N; n += 4) dst[n >> 2] = make_float4(mn[n], mn[n + 1], mn[n + 2], mn[n + 3]);
        tile_ver[blk * 4 + quarter] = ref_ver;
      }
      t += kNumWG;
      while (t >= a.n_tiles) { t -= a.n_tiles; ++pair; }
    }
    tmem_wait_st();                                  // this thread's stores to TMEM are complete before pass 2 reads them
    if (et == 0 && ph == nph - 1) FKV_TRACE(4);
    if (et == 0 && ph == 0 && nph > 1) FKV_TRACE(9);

    // ---------------- final (max, 1/sum) per row for every local pair ---------------------------
    // No grid barrier: a CTA only needs the partials of the CTAs that share its (b, h) pairs, and every
    // record says by itself whether it belongs to this launch.
    if (et < nlp) {
      const int ub = (pair0 + et - p0) * a.n_tiles;
      ncontrib[et] = cta_of_unit(ub + a.n_tiles - 1, Uph, grid) - cta_of_unit(ub, Uph, grid) + 1;
    }
    named_bar(2, kEpiThreads);
    // 8 lanes per row; every lane first issues all of its (up to 4) record loads, then merges.
    for (int idx = et; idx < nlp * N * 8; idx += kEpiThreads) {
      const int sub = idx & 7, n = (idx >> 3) % N, lp = (idx >> 3) / N;
      const int cnt = ncontrib[lp];
      const unsigned long long* src = a.part + (((size_t)(pair0 + lp) * a.maxc) * N + n) * 2;
      float M = kEmptyMax, Z = 0.f;
      for (int c0 = sub; c0 < cnt; c0 += 32) {
        unsigned long long w0[4], w1[4];
        bool ok;
        // (this wait is unbounded: a spin counter here was measured at +0.5 us at 32k and +11 us for B = 8 x 32k — the
        // kernel is that sensitive to its instruction footprint.  Its partners are always resident: the launch is
        // cooperative and callers never co-schedule another persistent grid unless both fit the chip, ops.SelectOverlap)
        do {
          ok = true;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = c0 + 8 * i;
            if (c < cnt) ld_record(src + (size_t)c * N * 2, w0[i], w1[i]);
            else { w0[i] = tag_word(kEmptyMax, tag_m); w1[i] = tag_word(0.f, tag_z); }
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) ok = ok && ((unsigned)(w0[i] >> 32) == tag_m) && ((unsigned)(w1[i] >> 32) == tag_z);
        } while (!ok);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float vm = __uint_as_float((unsigned)w0[i]), vz = __uint_as_float((unsigned)w1[i]);
          const float nm = fmaxf(M, vm);
          Z = Z * ex2(M - nm) + vz * ex2(vm - nm);
          M = nm;
        }
      }
#pragma unroll
      for (int o = 1; o < 8; o <<= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, M, o), oz = __shfl_xor_sync(0xffffffffu, Z, o);
        const float nm = fmaxf(M, om);
        Z = Z * ex2(M - nm) + oz * ex2(om - nm);
        M = nm;
      }
      if (sub == 0) colstat[lp * N + n] = make_float2(M, 1.0f / Z);
    }
    named_bar(2, kEpiThreads);
    if (et == 0 && ph == nph - 1) FKV_TRACE(6);
    if (et == 0 && ph == 0 && nph > 1) FKV_TRACE(10);

    // ---------------- pass 2: votes -> pool -> group sum -> score ------------------------------
    // every tile of the phase is still resident in TMEM; reading it frees the block for the next phase
    int cur_lp = -1, b = 0, kvl = 0;
    unsigned int cur_ver = 0u;                       // version of the references the coefficients in mn[] belong to
    { const int u = u0 + wg; pair = u / a.n_tiles; t = u - pair * a.n_tiles; }
    for (int j2 = wg; j2 < T; j2 += kNumWG) {
      const int lp = pair - pair0;
      const int blk = (jbase + j2) & (NB - 1);
      if (lp != cur_lp) { b = pair / a.nkv; kvl = pair - b * a.nkv; }
      // coefficient of column n for this tile: exp2(ref_n - max_n) / sum_n; it changes only when the warp's references
      // did (about once per pair), so it normally stays in registers from tile to tile
      const unsigned int ver = tile_ver[blk * 4 + quarter];
      if (ver != cur_ver || lp != cur_lp) {           // warp-uniform
        if (lane < N) {
          const float2 st = colstat[lp * N + lane];
          cstage[ew * N + lane] = ex2(tile_ref[(blk * 4 + quarter) * N + lane] - st.x) * st.y;
        }
        __syncwarp();
        const float4* cs = reinterpret_cast<const float4*>(cstage + ew * N);
#pragma unroll
        for (int n = 0; n < N; n += 4) { const float4 c4 = cs[n >> 2]; mn[n] = c4.x; mn[n + 1] = c4.y; mn[n + 2] = c4.z; mn[n + 3] = c4.w; }
        __syncwarp();
        cur_ver = ver; cur_lp = lp;
      }
      float s[N];
      tmem_load_row<N>(my_tmem + blk * NMMA, s);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(bar_tempty(blk));
      const int key = t * a.S - a.half + row;
      const bool in_prefix = (key >= 0) && (key < a.P);
      float* vb = vbuf + ((wg * 2 + (vcount++ & 1)) * G) * kTile;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float vote = 0.f;
#pragma unroll
        for (int r = 0; r < W; ++r)                  // rows summed in order (utils.py:104)
          vote = fmaf(s[g * W + r], mn[g * W + r], vote);
        vb[g * kTile + row] = in_prefix ? vote : (kDef ? 0.f : pad);
      }
      named_bar(3 + wg, 128);
      const bool own = (row >= a.half) && (row < a.half + a.S) && in_prefix;
      if constexpr (kDef) {
        float acc = 0.f;
        if (own) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const float* vr = vb + g * kTile + row;
            float sum = 0.f;
#pragma unroll
            for (int d = -3; d <= 3; ++d) sum += vr[d];
            acc += sum * inv_ks;                      // avg_pool1d (count_include_pad), GQA group-sum with g in order
          }
          a.score[(int64_t)b * a.score_sb + (int64_t)(a.kv0 + kvl - a.out_h0) * a.score_sh + key] = acc;
        }
        if (a.hist) hist_red(a.hist + (size_t)(b * a.hist_hpb + a.kv0 + kvl - a.out_h0) * 4096, acc, own);
      } else {
        float pooled_g[G];
        float acc = 0.f;
        if (own) {
          // a later head sub-group continues the sum of the earlier launches: ((p0 + p1) + p2) + p3 as in utils.py:112
          acc = a.accumulate ? a.score[(int64_t)b * a.score_sb + (int64_t)(a.kv0 + kvl - a.out_h0) * a.score_sh + key] : 0.f;
  #pragma unroll
          for (int g = 0; g < G; ++g) {
            float pooled;
            if (a.pooling == FKV_POOL_MAX) {
              pooled = -INFINITY;
              for (int d = -a.half; d <= a.half; ++d) pooled = fmaxf(pooled, vb[g * kTile + row + d]);
            } else {
              float sum = 0.f;
              const float* vr = vb + g * kTile + row;
              if (a.half == 3) {                        // kernel_size 7 (the reference default), unrolled
  #pragma unroll
                for (int d = -3; d <= 3; ++d) sum += vr[d];
              } else if (a.half == 2) {
  #pragma unroll
                for (int d = -2; d <= 2; ++d) sum += vr[d];
              } else {
                for (int d = -a.half; d <= a.half; ++d) sum += vr[d];
              }
              pooled = sum * inv_ks;                    // count_include_pad: always / kernel_size
            }
            pooled_g[g] = pooled;
            if (a.group_sum) acc += pooled;             // GQA group-sum, g in order (utils.py:112)
            else a.score[(int64_t)b * a.score_sb + (int64_t)((a.kv0 + kvl) * a.q_heads_per_kv + a.g0 + g - a.out_h0) * a.score_sh + key] = pooled;
          }
          if (a.group_sum) a.score[(int64_t)b * a.score_sb + (int64_t)(a.kv0 + kvl - a.out_h0) * a.score_sh + key] = acc;
        }
        if (a.hist) {                                  // warp-uniform: the match below needs all 32 lanes
          if (a.group_sum) hist_red(a.hist + (size_t)(b * a.hist_hpb + a.kv0 + kvl - a.out_h0) * 4096, acc, own);
          else {
#pragma unroll
            for (int g = 0; g < G; ++g)
              hist_red(a.hist + (size_t)(b * a.hist_hpb + (a.kv0 + kvl) * a.q_heads_per_kv + a.g0 + g - a.out_h0) * 4096,
                       own ? pooled_g[g] : 0.f, own);
          }
        }
      }
      t += kNumWG;
      while (t >= a.n_tiles) { t -= a.n_tiles; ++pair; }
    }
    if (et == 0 && ph == 0 && nph > 1) FKV_TRACE(11);
    jbase += T;
    }                                                // phases
  }

  if (threadIdx.x == kCtrlThreads) {
    FKV_TRACE(7);
    // departure count: the last CTA out resets it and advances the epoch, so the next launch on this
    // stream needs no memset node in front of it and can tell its own records from older ones
    const unsigned int prev = atomicAdd(a.sync_counter + 1, 1u);
    if (prev == (unsigned)grid - 1) {
      a.sync_counter[1] = 0u;
      *reinterpret_cast<unsigned long long*>(a.sync_counter + 2) += 1ull;
      __threadfence();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == kCtrlThreads) FKV_TRACE(8);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

struct Plan {
  int G, Gm, W, N, nkv, kv0, n_tiles, S, half, U, grid, maxlp, nst, maxc, npairs, ppp, nph;   // G = heads per launch, Gm = model group size
  size_t smem, part_bytes;
  bool ok;
};

Plan make_plan(const fkv_score_params& p) {
  Plan pl{};
  pl.ok = false;
  if (p.dtype != FKV_BF16 && p.dtype != FKV_F16) return pl;
  if (p.mode != FKV_MODE_FASTKV && p.mode != FKV_MODE_SNAPKV) return pl;
  const int Gm = p.Hq / p.Hkv;
  pl.W = p.window_size;
  if (!(pl.W == 8 || pl.W == 16 || pl.W == 32)) return pl;
  // At most 32 logit rows per MMA tile (running max + sum + logits = 96 live registers per epilogue thread):
  // wider groups (W=16/32 at G=4, G=8 models) run as several launches over head sub-groups, each reading K once.
  int Gk = Gm;
  while (Gk > 1 && (Gk * pl.W > 32 || Gm % Gk || (Gk != 2 && Gk != 4))) --Gk;   // instances exist for 1, 2 and 4 heads
  if (Gk * pl.W > 32) return pl;
  pl.G = Gk; pl.Gm = Gm; pl.N = pl.G * pl.W;
  if (p.mode == FKV_MODE_SNAPKV) {
    if (p.h0 % Gm || p.h1 % Gm) return pl;           // whole kv-head groups only
    pl.kv0 = p.h0 / Gm; pl.nkv = (p.h1 - p.h0) / Gm;
  } else {
    pl.kv0 = p.h0; pl.nkv = p.h1 - p.h0;
  }
  pl.half = p.kernel_size / 2;
  pl.S = kTile - 2 * pl.half;
  if (pl.S < 32) return pl;
  pl.n_tiles = ceil_div(p.L, pl.S);
  const long long U = (long long)p.B * pl.nkv * pl.n_tiles;
  if (U * num_sms() >= (1ll << 31)) return pl;        // the kernel splits units with 32-bit arithmetic
  pl.U = (int)U;
  static const int env_grid = getenv("FKV_TC_GRID") ? atoi(getenv("FKV_TC_GRID")) : 0;   // experiments: leave SMs to a concurrent kernel
  int sms = (env_grid > 0 && env_grid < num_sms()) ? env_grid : num_sms();
  if (p.max_ctas > 0 && p.max_ctas < sms) sms = p.max_ctas;
  pl.grid = (int)(U < sms ? U : sms);
  // phases: groups of `ppp` (b, h) pairs whose tiles fit in the TMEM blocks of the CTAs (<= NB tiles per
  // CTA and phase), so every tile's logits stay on chip between the two passes
  pl.npairs = p.B * pl.nkv;
  {
    const int NBk = 512 / (pl.N < 16 ? 16 : pl.N);
    long long cap_tiles = NBk;                                           // per CTA and phase
    if ((long long)(kMaxLocalPairs - 1) * pl.n_tiles < cap_tiles) cap_tiles = (long long)(kMaxLocalPairs - 1) * pl.n_tiles;
    static const int env_ppp = getenv("FKV_TC_PPP") ? atoi(getenv("FKV_TC_PPP")) : 0;
    int ppp = 1;
    while (ppp < pl.npairs && ceil_div((long long)(ppp + 1) * pl.n_tiles, pl.grid) <= cap_tiles) ++ppp;
    if (env_ppp > 0 && env_ppp < ppp) ppp = env_ppp;                     // experiments: smaller phases
    if (ceil_div((long long)ppp * pl.n_tiles, pl.grid) > NBk) return pl; // one pair alone overflows TMEM (~290k tokens at N = 32)
    pl.nph = ceil_div(pl.npairs, ppp);
    if (pl.nph > kMaxPhases) return pl;
    pl.ppp = ceil_div(pl.npairs, pl.nph);                                // same phase count, balanced sizes
    const int t_ph = (int)ceil_div((long long)pl.ppp * pl.n_tiles, pl.grid);
    pl.maxlp = ceil_div(t_ph, pl.n_tiles) + 1;
    if (pl.maxlp > kMaxLocalPairs) return pl;
  }
  const int NMMA = pl.N < 16 ? 16 : pl.N;
  const size_t tile_bytes = (size_t)kTile * p.D * 2;
  const size_t q_bytes = ((size_t)NMMA * p.D * 2 + 1023) & ~(size_t)1023;
  const size_t NBp = 512 / NMMA;
  const size_t misc = 80 * 8 + 16 + (size_t)kEpiWarps * pl.N * 8 + (size_t)kMaxLocalPairs * pl.N * 8 + (size_t)kNumWG * 2 * pl.G * kTile * 4 +
                      kMaxLocalPairs * 4 + 2 * kMaxPhases * 4 +
                      NBp * 4 * pl.N * 4 + NBp * 4 * 4 + (size_t)kEpiWarps * pl.N * 4;   // tile references + versions + coefficient staging
  const size_t budget = 227 * 1024 - 1024 - 2 * q_bytes - misc - 256;
  int nst = (int)(budget / tile_bytes);
  if (nst > kMaxStages) nst = kMaxStages;
  if (nst < 2) return pl;
  pl.nst = nst;
  pl.smem = 1024 + nst * tile_bytes + 2 * q_bytes + misc + 256;
  pl.maxc = pl.n_tiles < pl.grid ? pl.n_tiles : pl.grid;   // CTAs that can contribute to one pair
  pl.part_bytes = align_up((size_t)pl.npairs * pl.maxc * pl.N * 2 * sizeof(unsigned long long), 256);
  pl.ok = true;
  return pl;
}

template <int W, int G, bool kDef>
int launch_inst(const CUtensorMap& mk, const CUtensorMap& mq, const TcArgs& a, int grid, size_t smem, cudaStream_t s) {
  { const int rc_attr = ensure_dyn_smem(score_tc_kernel<W, G, kDef>, 227 * 1024); if (rc_attr != FKV_OK) return rc_attr; }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  static const bool no_coop = getenv("FKV_NO_COOP") != nullptr;   // experiments only
  if (!no_coop) {
    attr[na].id = cudaLaunchAttributeCooperative;    // all CTAs co-resident: CTAs wait on each other's partials
    attr[na].val.cooperative = 1;
    ++na;
  }
  if (pdl_enabled()) {                               // programmatic dependent launch: the prologue overlaps the previous kernel's tail
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  FKV_CUDA_TRY(cudaLaunchKernelEx(&cfg, score_tc_kernel<W, G, kDef>, mk, mq, a));
  count_launch();
  return FKV_OK;
}


}  // namespace

bool score_tensor_supported(const fkv_score_params& p) {
  if (encode_fn() == nullptr) return false;
  return p.mode == FKV_MODE_GEMFILTER ? gemfilter_tensor_supported(p) : make_plan(p).ok;
}

int launch_score_tensor(const fkv_score_params& p, cudaStream_t s) {
  const bool with_hist = p.hist_slot >= 1 && p.hist_slot <= kHistSlots && p.mode != FKV_MODE_GEMFILTER &&
                         p.B * (p.h1 - p.h0) <= kHistMaxRows;
  if (p.mode == FKV_MODE_GEMFILTER) return launch_score_gemfilter(p, s);
  const Plan pl = make_plan(p);
  if (!pl.ok) return FKV_EMODE;
  if (p.workspace_bytes < pl.part_bytes + FKV_WORKSPACE_HEADER_BYTES) return FKV_EWORKSPACE;

  TcArgs a{};
  a.B = p.B; a.L = p.L; a.P = p.L - p.window_size; a.D = p.D;
  a.nkv = pl.nkv; a.kv0 = pl.kv0; a.n_tiles = pl.n_tiles; a.S = pl.S; a.half = pl.half; a.ks = p.kernel_size;
  a.pooling = p.pooling; a.group_sum = (p.mode == FKV_MODE_FASTKV); a.is_f16 = (p.dtype == FKV_F16);
  a.U = pl.U; a.maxlp = pl.maxlp; a.nst = pl.nst; a.maxc = pl.maxc; a.q_heads_per_kv = pl.Gm;
  a.npairs = pl.npairs; a.ppp = pl.ppp; a.nph = pl.nph;
  a.c = (1.0f / sqrtf((float)p.D)) * 1.4426950408889634f;
  char* ws = reinterpret_cast<char*>(p.workspace);
  a.sync_counter = reinterpret_cast<unsigned int*>(ws);
  a.part = reinterpret_cast<unsigned long long*>(ws + FKV_WORKSPACE_HEADER_BYTES);
  void* hist = with_hist ? ws + kWsHistOffset + (size_t)(p.hist_slot - 1) * kHistMaxRows * 4096 * 4 : nullptr;
  a.score = p.score; a.score_sb = p.score_stride_b; a.score_sh = p.score_stride_h; a.out_h0 = p.h0;

  CUtensorMap mk, mq;
  int rc;
  if ((rc = make_map(&mk, p.k, p.dtype, p.B, p.Hkv, p.L, p.D, kTile, &a.kpos_l, &a.kpos_h, &a.kpos_b)) != FKV_OK) return rc;
  if ((rc = make_map(&mq, p.q, p.dtype, p.B, p.Hq, p.L, p.D, pl.W, &a.qpos_l, &a.qpos_h, &a.qpos_b)) != FKV_OK) return rc;
  static const bool trace_on = getenv("FKV_TC_TRACE") != nullptr;
  static unsigned long long* trace_buf = nullptr;
  if (trace_on) {
    if (!trace_buf) FKV_CUDA_TRY(cudaMalloc(&trace_buf, (256 * 16 + 1024) * sizeof(unsigned long long)));
    a.trace = trace_buf;
  }

  int rc_launch = FKV_EMODE;
#define FKV_TC_CASE(w, g) if (pl.W == w && pl.G == g) rc_launch = def_epi ? launch_inst<w, g, true>(mk, mq, a, pl.grid, pl.smem, s) \
                                                                          : launch_inst<w, g, false>(mk, mq, a, pl.grid, pl.smem, s)
  for (int g0 = 0; g0 < pl.Gm; g0 += pl.G) {         // one launch per head sub-group (a single one for W*G <= 32)
    a.g0 = g0;
    a.accumulate = (a.group_sum && g0 > 0) ? 1 : 0;
    // the histogram counts FINAL scores: with a group sum spread over several launches only the last one adds to it
    a.hist = (hist && (!a.group_sum || g0 + pl.G >= pl.Gm)) ? reinterpret_cast<unsigned int*>(hist) : nullptr;
    a.hist_hpb = p.h1 - p.h0;
    const bool def_epi = a.group_sum && !a.accumulate && a.pooling == FKV_POOL_AVG && a.half == 3;
    FKV_TC_CASE(8, 1); FKV_TC_CASE(8, 2); FKV_TC_CASE(8, 4);
    FKV_TC_CASE(16, 1); FKV_TC_CASE(16, 2);
    FKV_TC_CASE(32, 1);
    if (rc_launch != FKV_OK) return rc_launch;
  }
#undef FKV_TC_CASE
  if (trace_on && rc_launch == FKV_OK) {                // debugging aid only: synchronises and prints phase times
    cudaStreamSynchronize(s);
    static unsigned long long host[256 * 16 + 1024];
    cudaMemcpy(host, trace_buf, sizeof(host), cudaMemcpyDeviceToHost);
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < pl.grid; ++c) if (host[c * 16] < t0) t0 = host[c * 16];
    const char* names[12] = {"start", "setup", "first_acc", "-", "last_flushed", "-", "last_colstat", "pass2_end", "exit",
                             "ph0_flushed", "ph0_colstat", "ph0_pass2_end"};
    fprintf(stderr, "[fkv trace] L=%d grid=%d pairs=%d ppp=%d phases=%d tiles/pair=%d stages=%d\n", p.L, pl.grid, pl.npairs, pl.ppp, pl.nph, pl.n_tiles, pl.nst);
    for (int sidx = 0; sidx < (pl.nph > 1 ? 12 : 9); ++sidx) {
      if (sidx == 5 || sidx == 3) continue;
      unsigned long long mn = ~0ull, mx = 0; double avg = 0;
      for (int c = 0; c < pl.grid; ++c) { unsigned long long v = host[c * 16 + sidx] - t0; if (v < mn) mn = v; if (v > mx) mx = v; avg += v; }
      fprintf(stderr, "[fkv trace] %-12s min %8.2f us  avg %8.2f us  max %8.2f us\n", names[sidx], mn / 1e3, avg / pl.grid / 1e3, mx / 1e3);
    }
    if (atoi(getenv("FKV_TC_TRACE")) >= 2) {            // CTA 0, tile by tile: TMA issue, operands landed (MMA issue), MMA committed
      const unsigned long long* r = host + 256 * 16;
      for (int m = 0; m < 64 && r[m]; ++m)
        fprintf(stderr, "[fkv trace] cta0 tile %2d  tma %7.2f  mma %7.2f  committed %7.2f us\n", m, (r[m] - t0) / 1e3, (r[64 + m] - t0) / 1e3,
                (r[128 + m] - t0) / 1e3);
    }
  }
  return rc_launch;
}

}  // namespace fkv
