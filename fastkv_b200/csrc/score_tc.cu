// K1/K2 tensor-core paths, hand-written for sm_100a (TMA -> shared memory -> tcgen05.mma -> TMEM -> registers):
//   score_tc_kernel        observation-window scoring + pooling + GQA group-sum (baselines/fastkv/utils.py:93-112
//                          and the SnapKV twin), one persistent cooperative kernel that reads K exactly once
//   gemfilter_tc_kernel    GemFilter selection scores (baselines/gemfilter/utils.py:25-33), single pass
//   small_compress_kernel  opt-in: the whole update_kv chain of a short-context layer in one cluster kernel
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
// Warp roles (448 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM allocator (both stream
// through all phases without stopping), warps 2-13 = three epilogue warpgroups taking tiles round-robin.
#include <cuda.h>
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace fkv {
namespace {

constexpr int kTile = 128;
constexpr int kNumWG = 3;                  // epilogue warpgroups (tiles are dealt round-robin to them)
constexpr int kEpiThreads = 128 * kNumWG;
constexpr int kEpiWarps = 4 * kNumWG;
constexpr int kTcThreads = 64 + kEpiThreads;
constexpr int kMaxLocalPairs = 8;
constexpr int kMaxStages = 6;
constexpr int kMaxPhases = 512;            // (b, h) pairs are processed in groups whose logits fit in TMEM
constexpr float kLazyTau = 6.0f;          // log2 units of headroom before a running max is raised
constexpr float kEmptyMax = -1.0e30f;     // "no key seen yet" (finite, so exp2(x - m) never sees inf - inf)

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
};

// ---- raw PTX wrappers ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tWAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra WAIT_DONE;\n\tbra WAIT_LOOP;\n\tWAIT_DONE:\n\t}"
      ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
template <int N>
__device__ __forceinline__ void tmem_load_row(uint32_t taddr, float* v) {
  if constexpr (N == 8) tmem_ld8(taddr, v);
  else if constexpr (N == 16) tmem_ld16(taddr, v);
  else {
#pragma unroll
    for (int c = 0; c < N; c += 32) tmem_ld32(taddr + c, v + c);
  }
}

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Phase timelines (%globaltimer stamps printed by the host side) are a build-time option: python fastkv_b200/_build.py
// with FKV_BUILD_TRACE=1 compiles them in, FKV_TC_TRACE=1 then prints them.  The default build carries no
// tracing code: the single-thread TMA / MMA loops pace the pipeline and every instruction in them counts.
#ifdef FKV_TC_TRACING
#define FKV_TRACE2(role, m) do { if (a.trace && blockIdx.x == 0 && (m) < 64) a.trace[256 * 16 + (role) * 64 + (m)] = gtimer(); } while (0)
#define FKV_TRACE(slot) do { if (a.trace) a.trace[(size_t)blockIdx.x * 16 + (slot)] = gtimer(); } while (0)
#else
#define FKV_TRACE2(role, m) do { } while (0)
#define FKV_TRACE(slot) do { } while (0)
#endif

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: LBO(enc)=1, SBO = 1024 B, version 1.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// Partial softmax statistics travel between CTAs as two self-validating 64-bit words per row:
// low half = the fp32 value, high half = a quiet-NaN pattern carrying 22 bits of the launch epoch
// (44 bits over the two words).  A reader simply re-loads a record until both tags match this launch,
// so the exchange needs no flag, no fence and no grid-wide barrier: data and "ready" are one word.
__device__ __forceinline__ unsigned long long tag_word(float v, unsigned int tag) {
  return (unsigned long long)__float_as_uint(v) | ((unsigned long long)tag << 32);
}
__device__ __forceinline__ void st_record(unsigned long long* p, unsigned long long w0, unsigned long long w1) {
  asm volatile("st.relaxed.gpu.global.v2.b64 [%0], {%1, %2};" ::"l"(p), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ void ld_record(const unsigned long long* p, unsigned long long& w0, unsigned long long& w1) {
  asm volatile("ld.relaxed.gpu.global.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(p) : "memory");
}

// balanced split of U units over the grid: u0(c) = floor(c * U / grid); with fewer units than CTAs the
// first U CTAs take one unit each, so the CTAs that share a (b, h) pair are always consecutive and non-empty
// (32-bit arithmetic: the plan guarantees U * grid < 2^31)
__device__ __forceinline__ int unit_begin(int c, int U, int grid) {
  return U >= grid ? (int)((unsigned)c * (unsigned)U / (unsigned)grid) : min(c, U);
}
__device__ __forceinline__ int cta_of_unit(int u, int U, int grid) {
  return U >= grid ? (int)((((unsigned)u + 1u) * (unsigned)grid - 1u) / (unsigned)U) : u;
}

// warp-wide float max in one instruction (CREDUX, sm_100a)
__device__ __forceinline__ float warp_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
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
  if (threadIdx.x == 64) FKV_TRACE(0);
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
  if (threadIdx.x == 64) FKV_TRACE(1);

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
  } else {
    // ===================================== epilogue =========================================
    const int et = threadIdx.x - 64;                 // 0..255
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
    // a workspace that never went through fkv_workspace_init has a garbage departure counter / epoch and would
    // make stale records look valid: fail loudly instead of computing on them
    if (et == 0 && blockIdx.x == 0 && __ldcg(a.sync_counter + 4) != kWorkspaceMagic) {
      printf("fastkv_b200: fkv_score got a workspace that was not passed through fkv_workspace_init\n");
      __trap();
    }
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
      const bool own = (row >= a.half) && (row < a.half + a.S) && (key < a.L);
      const bool has_window = (t * a.S - a.half + kTile > a.P);   // warp-uniform
      const int kp = key - a.P;                      // window key index (>= 0 inside the window)
      // window row r sees window keys <= r (utils.py:95-101): mask the others to -inf
      if (has_window) {
#pragma unroll
        for (int n = 0; n < N; ++n) s[n] = (kp > (n % W)) ? -INFINITY : s[n];
      }
      // cheap check first: does any owned key exceed its row's reference by more than tau?
      float dmax = -INFINITY;
#pragma unroll
      for (int n = 0; n < N; ++n) dmax = fmaxf(dmax, fmaf(s[n], a.c, -mn[n]));
      if (__any_sync(0xffffffffu, own && dmax > kLazyTau)) {
        // rare (always on the first tile of a pair): raise the references to the warp-wide column max
#pragma unroll
        for (int n = 0; n < N; ++n) {
          const float nm = fmaxf(mn[n], warp_max_f32(own ? s[n] * a.c : -INFINITY));
          z[n] *= ex2(mn[n] - nm);
          mn[n] = nm;
        }
      }
      if (own) {
#pragma unroll
        for (int n = 0; n < N; ++n) z[n] += ex2(fmaf(s[n], a.c, -mn[n]));
      }
      t += kNumWG;
      while (t >= a.n_tiles) { t -= a.n_tiles; ++pair; }
    }
    if (et == 0 && ph == nph - 1) FKV_TRACE(4);

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

    // ---------------- pass 2: votes -> pool -> group sum -> score ------------------------------
    // every tile of the phase is still resident in TMEM; reading it frees the block for the next phase
    int cur_lp = -1, b = 0, kvl = 0;
    { const int u = u0 + wg; pair = u / a.n_tiles; t = u - pair * a.n_tiles; }
    for (int j2 = wg; j2 < T; j2 += kNumWG) {
      const int lp = pair - pair0;
      const int blk = (jbase + j2) & (NB - 1);
      if (lp != cur_lp) {                            // (max, 1/sum) of this pair's rows -> registers
        cur_lp = lp;
        b = pair / a.nkv; kvl = pair - b * a.nkv;
#pragma unroll
        for (int n = 0; n < N; ++n) { const float2 v = colstat[lp * N + n]; mn[n] = v.x; z[n] = v.y; }
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
          vote = fmaf(ex2(fmaf(s[g * W + r], a.c, -mn[g * W + r])), z[g * W + r], vote);
        vb[g * kTile + row] = in_prefix ? vote : (kDef ? 0.f : pad);
      }
      named_bar(3 + wg, 128);
      const bool own = (row >= a.half) && (row < a.half + a.S) && in_prefix;
      if (own) {
        if constexpr (kDef) {
          float acc = 0.f;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const float* vr = vb + g * kTile + row;
            float sum = 0.f;
#pragma unroll
            for (int d = -3; d <= 3; ++d) sum += vr[d];
            acc += sum * inv_ks;                      // avg_pool1d (count_include_pad), GQA group-sum with g in order
          }
          a.score[(int64_t)b * a.score_sb + (int64_t)(a.kv0 + kvl - a.out_h0) * a.score_sh + key] = acc;
        } else {
          // a later head sub-group continues the sum of the earlier launches: ((p0 + p1) + p2) + p3 as in utils.py:112
          float acc = a.accumulate ? a.score[(int64_t)b * a.score_sb + (int64_t)(a.kv0 + kvl - a.out_h0) * a.score_sh + key] : 0.f;
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
            if (a.group_sum) acc += pooled;             // GQA group-sum, g in order (utils.py:112)
            else a.score[(int64_t)b * a.score_sb + (int64_t)((a.kv0 + kvl) * a.q_heads_per_kv + a.g0 + g - a.out_h0) * a.score_sh + key] = pooled;
          }
          if (a.group_sum) a.score[(int64_t)b * a.score_sb + (int64_t)(a.kv0 + kvl - a.out_h0) * a.score_sh + key] = acc;
        }
      }
      t += kNumWG;
      while (t >= a.n_tiles) { t -= a.n_tiles; ++pair; }
    }
    jbase += T;
    }                                                // phases
  }

  if (threadIdx.x == 64) {
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
  if (threadIdx.x == 64) FKV_TRACE(8);
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// ---- GemFilter selection scores on tensor cores (baselines/gemfilter/utils.py:25-56) ------------------------
// ip[b, j] = sum over all query heads of q[b, hq, L-1, :] . K[b, hq / G, j, :]  (no scale, no softmax, no
// window exclusion), then avg_pool1d(kernel, pad = kernel/2) over j.  One unit = (b, 128-key tile): the
// K tiles of ALL kv heads for those keys go through one TMEM accumulator (the MMA's accumulate flag does
// the sum over kv heads, column g holds query head g of every kv head), the epilogue adds the G columns,
// pools through shared memory and writes the score.  One pass, K read once, no cross-CTA exchange.
struct GfArgs {
  int B, L, D, Hkv, G, n_tiles, S, half, ks, is_f16, U, nst;
  int kpos_l, kpos_h, kpos_b, qpos_l, qpos_h, qpos_b;
  float* score; int64_t score_sb;
};

constexpr int kGfThreads = 64 + 128;
constexpr int kGfCols = 16;                       // UMMA N (query heads per kv head, padded)

__global__ void __launch_bounds__(kGfThreads, 1)
gemfilter_tc_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_q, const GfArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  unsigned char* sm = smem_dyn + (base - smem_u32(smem_dyn));
  const int tile_bytes = kTile * a.D * 2;
  const int qtile_bytes = kGfCols * a.D * 2;        // one kv head's query rows (B operand), 2 or 4 KB
  const uint32_t k_off = 0;
  const uint32_t q_off = k_off + a.nst * tile_bytes;
  const uint32_t misc_off = q_off + ((a.Hkv * qtile_bytes + 1023) & ~1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sm + misc_off);      // full[6] empty[6] qfull qempty tfull[2] tempty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  float* vbuf = reinterpret_cast<float*>(tmem_slot + 4);            // [2][128] head-summed products of a tile
  auto bar_full = [&](int s) { return smem_u32(bars + s); };
  auto bar_empty = [&](int s) { return smem_u32(bars + 6 + s); };
  const uint32_t bar_qfull = smem_u32(bars + 12), bar_qempty = smem_u32(bars + 13);
  auto bar_tfull = [&](int s) { return smem_u32(bars + 14 + s); };
  auto bar_tempty = [&](int s) { return smem_u32(bars + 16 + s); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grid = gridDim.x, cta = blockIdx.x;
  const int u0 = unit_begin(cta, a.U, grid), u1 = unit_begin(cta + 1, a.U, grid);
  if (warp == 0) {
    for (int s = lane; s < kMaxStages; s += 32) { mbar_init(bar_full(s), 1); mbar_init(bar_empty(s), 1); }
    if (lane == 0) { mbar_init(bar_qfull, 1); mbar_init(bar_qempty, 1); }
    if (lane < 2) { mbar_init(bar_tfull(lane), 1); mbar_init(bar_tempty(lane), 128); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(32));
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
  pdl_enter();
  const int slabs = a.D >> 6;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (lane == 0) {
      const int kl1 = a.kpos_l == 1, kl2 = a.kpos_l == 2, kl3 = a.kpos_l == 3;
      const int kh1 = a.kpos_h == 1, kh2 = a.kpos_h == 2, kh3 = a.kpos_h == 3;
      const int kb1 = a.kpos_b == 1, kb2 = a.kpos_b == 2, kb3 = a.kpos_b == 3;
      const int ql1 = a.qpos_l == 1, ql2 = a.qpos_l == 2, ql3 = a.qpos_l == 3;
      const int qh1 = a.qpos_h == 1, qh2 = a.qpos_h == 2, qh3 = a.qpos_h == 3;
      const int qb1 = a.qpos_b == 1, qb2 = a.qpos_b == 2, qb3 = a.qpos_b == 3;
      int cur_b = -1, nq = 0;
      int slot = 0, ephase = 1, m = 0;
      int b = u0 / a.n_tiles, t = u0 - b * a.n_tiles;
      for (int u = u0; u < u1; ++u) {
        if (b != cur_b) {                            // the last query row of every head of this batch entry
          cur_b = b;
          if (nq > 0) mbar_wait(bar_qempty, (nq - 1) & 1);
          ++nq;
          mbar_expect_tx(bar_qfull, a.Hkv * a.G * a.D * 2);
          const int lq = a.L - 1;
          for (int h = 0; h < a.Hkv; ++h)
            for (int g = 0; g < a.G; ++g) {
              const int hq = h * a.G + g;
              for (int slab = 0; slab < slabs; ++slab)
                tma_load_4d(base + q_off + h * qtile_bytes + slab * (kGfCols * 128) + g * 128, &tm_q, bar_qfull, slab * 64,
                            lq * ql1 + hq * qh1 + b * qb1, lq * ql2 + hq * qh2 + b * qb2, lq * ql3 + hq * qh3 + b * qb3);
            }
        }
        const int row0 = t * a.S - a.half;
        for (int h = 0; h < a.Hkv; ++h, ++m) {
          if (m >= a.nst) mbar_wait(bar_empty(slot), ephase);
          mbar_expect_tx(bar_full(slot), tile_bytes);
          const uint32_t kdst = base + k_off + slot * tile_bytes;
          const int c1 = row0 * kl1 + h * kh1 + b * kb1, c2 = row0 * kl2 + h * kh2 + b * kb2, c3 = row0 * kl3 + h * kh3 + b * kb3;
          tma_load_4d(kdst, &tm_k, bar_full(slot), 0, c1, c2, c3);
          if (slabs == 2) tma_load_4d(kdst + kTile * 128, &tm_k, bar_full(slot), 64, c1, c2, c3);
          if (++slot == a.nst) { slot = 0; ephase ^= 1; }
        }
        if (++t == a.n_tiles) { t = 0; ++b; }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =======================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((a.is_f16 ? 0u : 1u) << 7) | ((a.is_f16 ? 0u : 1u) << 10) |
                             ((uint32_t)(kGfCols >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
      const uint64_t desc_hi = (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
      int cur_b = -1, nq = 0, slot = 0, fphase = 0;
      int b = u0 / a.n_tiles, t = u0 - b * a.n_tiles;
      for (int u = u0, ui = 0; u < u1; ++u, ++ui) {
        if (b != cur_b) {
          if (nq > 0) umma_commit(bar_qempty);        // the previous batch entry's query tiles are free once its MMAs retire
          cur_b = b;
          mbar_wait(bar_qfull, nq & 1);
          ++nq;
        }
        const int blk = ui & 1;
        if (ui >= 2) mbar_wait(bar_tempty(blk), ((ui >> 1) - 1) & 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + blk * kGfCols;
        for (int h = 0; h < a.Hkv; ++h) {
          mbar_wait(bar_full(slot), fphase);
          tc_fence_after();
          const uint32_t ka = base + k_off + slot * tile_bytes;
          const uint32_t qa = base + q_off + h * qtile_bytes;
          const uint64_t ad0 = desc_hi | (uint64_t)((ka & 0x3ffffu) >> 4);
          const uint64_t bd0 = desc_hi | (uint64_t)((qa & 0x3ffffu) >> 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_f16(d_tmem, ad0 + 2 * k, bd0 + 2 * k, idesc, (h > 0 || k > 0) ? 1u : 0u);
          if (slabs == 2) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16(d_tmem, ad0 + ((kTile * 128) >> 4) + 2 * k, bd0 + ((kGfCols * 128) >> 4) + 2 * k, idesc, 1u);
          }
          umma_commit(bar_empty(slot));
          if (++slot == a.nst) { slot = 0; fphase ^= 1; }
        }
        umma_commit(bar_tfull(blk));
        if (++t == a.n_tiles) { t = 0; ++b; }
      }
    }
  } else {
    // ===================================== epilogue =========================================
    const int et = threadIdx.x - 64;
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t my_tmem = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int b = u0 / a.n_tiles, t = u0 - b * a.n_tiles;
    for (int u = u0, ui = 0; u < u1; ++u, ++ui) {
      const int blk = ui & 1;
      mbar_wait(bar_tfull(blk), (ui >> 1) & 1);
      tc_fence_after();
      float s[kGfCols];
      tmem_load_row<kGfCols>(my_tmem + blk * kGfCols, s);
      tmem_wait_ld();
      tc_fence_before();
      mbar_arrive(bar_tempty(blk));
      float ip = 0.f;
#pragma unroll
      for (int g = 0; g < kGfCols; ++g) ip += (g < a.G) ? s[g] : 0.f;
      const int key = t * a.S - a.half + row;
      const bool in_range = (key >= 0) && (key < a.L);
      float* vb = vbuf + (ui & 1) * kTile;
      vb[row] = in_range ? ip : 0.f;                 // avg_pool1d zero padding
      named_bar(2, 128);
      if (row >= a.half && row < a.half + a.S && in_range) {
        float sum = 0.f;
        for (int d = -a.half; d <= a.half; ++d) sum += vb[row + d];
        a.score[(int64_t)b * a.score_sb + key] = sum / (float)a.ks;
      }
      if (++t == a.n_tiles) { t = 0; ++b; }
      (void)et;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(32));
  }
}

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

// ---- host side -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      (void)cudaGetLastError();
  });
  return fn;
}

// [B,H,L,D] view -> 4-D tensor map (dim0 = D, the other three sorted by stride) with a
// (64 x box_rows x 1 x 1) box on (D, L); reports the coordinate slot of l / h / b.
int make_map(CUtensorMap* map, const fkv_view4& v, int dtype, int B, int H, int L, int D, int box_rows, int* pos_l,
             int* pos_h, int* pos_b) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return (int)cudaErrorNotSupported;
  struct Dim { uint64_t size, stride; int box, id; } d[3] = {
      {(uint64_t)L, (uint64_t)v.stride_l * 2, box_rows, 0},
      {(uint64_t)H, (uint64_t)v.stride_h * 2, 1, 1},
      {(uint64_t)B, (uint64_t)v.stride_b * 2, 1, 2}};
  for (int i = 0; i < 3; ++i)
    if (d[i].size == 1) d[i].stride = 0;             // degenerate dims get a harmless stride below
  // sort by stride (degenerate dims last)
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j) {
      const uint64_t si = d[i].size == 1 ? ~0ull : d[i].stride, sj = d[j].size == 1 ? ~0ull : d[j].stride;
      if (sj < si) { Dim t = d[i]; d[i] = d[j]; d[j] = t; }
    }
  cuuint64_t gdim[4] = {(cuuint64_t)D, d[0].size, d[1].size, d[2].size};
  cuuint64_t gstr[3];
  uint64_t prev = (uint64_t)D * 2;
  for (int i = 0; i < 3; ++i) {
    uint64_t s = d[i].stride;
    if (d[i].size == 1 || s == 0) s = prev * (i == 0 ? 1 : gdim[i]);   // any multiple of 16 works for a size-1 dim
    gstr[i] = s;
    prev = s;
  }
  cuuint32_t box[4] = {64, (cuuint32_t)d[0].box, (cuuint32_t)d[1].box, (cuuint32_t)d[2].box};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  for (int i = 0; i < 3; ++i) {
    if (d[i].id == 0) *pos_l = i + 1;
    if (d[i].id == 1) *pos_h = i + 1;
    if (d[i].id == 2) *pos_b = i + 1;
  }
  const CUtensorMapDataType dt = dtype == FKV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUresult r = fn(map, dt, 4, const_cast<void*>(v.ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? FKV_OK : (int)cudaErrorInvalidValue;
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
  while (Gk > 1 && (Gk * pl.W > 32 || Gm % Gk)) --Gk;
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
  pl.grid = (int)(U < num_sms() ? U : num_sms());
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
  const size_t misc = 80 * 8 + 16 + (size_t)kEpiWarps * pl.N * 8 + (size_t)kMaxLocalPairs * pl.N * 8 + (size_t)kNumWG * 2 * pl.G * kTile * 4 +
                      kMaxLocalPairs * 4 + 2 * kMaxPhases * 4;
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
  static bool attr_set = false;
  if (!attr_set) {
    FKV_CUDA_TRY(cudaFuncSetAttribute(score_tc_kernel<W, G, kDef>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeCooperative;      // all CTAs co-resident: CTAs wait on each other's partials
  attr[0].val.cooperative = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;             // + programmatic dependent launch: the prologue overlaps the previous kernel's tail
  FKV_CUDA_TRY(cudaLaunchKernelEx(&cfg, score_tc_kernel<W, G, kDef>, mk, mq, a));
  count_launch();
  return FKV_OK;
}


struct GfPlan { int G, n_tiles, S, half, U, grid, nst; size_t smem; bool ok; };

GfPlan make_gf_plan(const fkv_score_params& p) {
  GfPlan pl{};
  pl.ok = false;
  if (p.mode != FKV_MODE_GEMFILTER || (p.dtype != FKV_BF16 && p.dtype != FKV_F16)) return pl;
  if (!(p.D == 64 || p.D == 128) || p.Hkv < 1 || p.Hq % p.Hkv) return pl;
  if (p.qh1 > p.qh0 && !(p.qh0 == 0 && p.qh1 == p.Hq)) return pl;   // head-window partial sums: simple path
  pl.G = p.Hq / p.Hkv;
  if (pl.G > kGfCols) return pl;
  pl.half = p.kernel_size / 2;
  pl.S = kTile - 2 * pl.half;
  if (pl.S < 32 || p.L < 1) return pl;
  pl.n_tiles = ceil_div(p.L, pl.S);
  const long long U = (long long)p.B * pl.n_tiles;
  if (U * num_sms() >= (1ll << 31)) return pl;
  pl.U = (int)U;
  pl.grid = (int)(U < num_sms() ? U : num_sms());
  const size_t tile_bytes = (size_t)kTile * p.D * 2;
  const size_t q_bytes = ((size_t)p.Hkv * kGfCols * p.D * 2 + 1023) & ~(size_t)1023;
  const size_t misc = 18 * 8 + 16 + 2 * kTile * 4 + 64;
  if (q_bytes + misc + 2 * tile_bytes + 2048 > 227 * 1024) return pl;
  int nst = (int)((227 * 1024 - 1024 - q_bytes - misc - 256) / tile_bytes);
  if (nst > kMaxStages) nst = kMaxStages;
  if (nst < 2) return pl;
  pl.nst = nst;
  pl.smem = 1024 + nst * tile_bytes + q_bytes + misc + 256;
  pl.ok = true;
  return pl;
}

int launch_gemfilter_tensor(const fkv_score_params& p, cudaStream_t s) {
  const GfPlan pl = make_gf_plan(p);
  if (!pl.ok) return FKV_EMODE;
  GfArgs a{};
  a.B = p.B; a.L = p.L; a.D = p.D; a.Hkv = p.Hkv; a.G = pl.G; a.n_tiles = pl.n_tiles; a.S = pl.S; a.half = pl.half;
  a.ks = p.kernel_size; a.is_f16 = (p.dtype == FKV_F16); a.U = pl.U; a.nst = pl.nst;
  a.score = p.score; a.score_sb = p.score_stride_b;
  CUtensorMap mk, mq;
  int rc;
  if ((rc = make_map(&mk, p.k, p.dtype, p.B, p.Hkv, p.L, p.D, kTile, &a.kpos_l, &a.kpos_h, &a.kpos_b)) != FKV_OK) return rc;
  if ((rc = make_map(&mq, p.q, p.dtype, p.B, p.Hq, p.L, p.D, 1, &a.qpos_l, &a.qpos_h, &a.qpos_b)) != FKV_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    FKV_CUDA_TRY(cudaFuncSetAttribute(gemfilter_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  FKV_CUDA_TRY(launch_pdl(gemfilter_tc_kernel, dim3(pl.grid), dim3(kGfThreads), pl.smem, s, mk, mq, a));
  count_launch();
  return FKV_OK;
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
  static bool attr_set = false;
  if (!attr_set) {
    FKV_CUDA_TRY(cudaFuncSetAttribute(small_compress_kernel<W, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
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

bool score_tensor_supported(const fkv_score_params& p) {
  if (encode_fn() == nullptr) return false;
  return p.mode == FKV_MODE_GEMFILTER ? make_gf_plan(p).ok : make_plan(p).ok;
}

int launch_score_tensor(const fkv_score_params& p, cudaStream_t s) {
  if (p.mode == FKV_MODE_GEMFILTER) return launch_gemfilter_tensor(p, s);
  const Plan pl = make_plan(p);
  if (!pl.ok) return FKV_EMODE;
  if (p.workspace_bytes < pl.part_bytes + 256) return FKV_EWORKSPACE;
  TcArgs a{};
  a.B = p.B; a.L = p.L; a.P = p.L - p.window_size; a.D = p.D;
  a.nkv = pl.nkv; a.kv0 = pl.kv0; a.n_tiles = pl.n_tiles; a.S = pl.S; a.half = pl.half; a.ks = p.kernel_size;
  a.pooling = p.pooling; a.group_sum = (p.mode == FKV_MODE_FASTKV); a.is_f16 = (p.dtype == FKV_F16);
  a.U = pl.U; a.maxlp = pl.maxlp; a.nst = pl.nst; a.maxc = pl.maxc; a.q_heads_per_kv = pl.Gm;
  a.npairs = pl.npairs; a.ppp = pl.ppp; a.nph = pl.nph;
  a.c = (1.0f / sqrtf((float)p.D)) * 1.4426950408889634f;
  char* ws = reinterpret_cast<char*>(p.workspace);
  a.sync_counter = reinterpret_cast<unsigned int*>(ws);
  a.part = reinterpret_cast<unsigned long long*>(ws + 256);
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
    const char* names[9] = {"start", "setup", "first_acc", "last_p1_end", "last_flushed", "-", "last_colstat", "pass2_end", "exit"};
    fprintf(stderr, "[fkv trace] L=%d grid=%d pairs=%d ppp=%d phases=%d tiles/pair=%d stages=%d\n", p.L, pl.grid, pl.npairs, pl.ppp, pl.nph, pl.n_tiles, pl.nst);
    for (int sidx = 0; sidx < 9; ++sidx) {
      if (sidx == 5) continue;
      unsigned long long mn = ~0ull, mx = 0; double avg = 0;
      for (int c = 0; c < pl.grid; ++c) { unsigned long long v = host[c * 16 + sidx] - t0; if (v < mn) mn = v; if (v > mx) mx = v; avg += v; }
      fprintf(stderr, "[fkv trace] %-12s min %8.2f us  avg %8.2f us  max %8.2f us\n", names[sidx], mn / 1e3, avg / pl.grid / 1e3, mx / 1e3);
    }
  }
  return rc_launch;
}

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
