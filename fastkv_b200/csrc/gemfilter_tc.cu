// GemFilter selection scores on tensor cores: see gemfilter_tc_kernel below.  Shares the TMA / tcgen05 plumbing of
// score_tc.cu through tc_common.cuh.
#include "tc_common.cuh"

namespace fkv {
namespace {

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
  { const int rc_attr = ensure_dyn_smem(gemfilter_tc_kernel, 227 * 1024); if (rc_attr != FKV_OK) return rc_attr; }
  FKV_CUDA_TRY(launch_pdl(gemfilter_tc_kernel, dim3(pl.grid), dim3(kGfThreads), pl.smem, s, mk, mq, a));
  count_launch();
  return FKV_OK;
}


}  // namespace

bool gemfilter_tensor_supported(const fkv_score_params& p) { return encode_fn() != nullptr && make_gf_plan(p).ok; }

int launch_score_gemfilter(const fkv_score_params& p, cudaStream_t s) { return launch_gemfilter_tensor(p, s); }

}  // namespace fkv
