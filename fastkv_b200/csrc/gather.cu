// K4: index-gather kernels.  Bit-exact 128-bit copies, one launch for K and V.
//   gather_kv_kernel    compacts the KV cache per selection head + appends the W window rows
//                       (baselines/fastkv/utils.py:114-121)
//   gather_rows_kernel  propagates hidden states (+ position ids) of the TSP-selected tokens
//                       (baselines/fastkv/llama_model.py:255-257)
#include "common.cuh"

namespace fkv {
namespace {

struct GatherKvArgs {
  const char* k; const char* v; char* ko; char* vo;
  int64_t k_sb, k_sh, k_sl, v_sb, v_sh, v_sl;          // bytes
  int64_t ko_sb, ko_sh, ko_sl, vo_sb, vo_sh, vo_sl;    // bytes
  const int64_t* idx; int64_t idx_sb, idx_sh;
  int L, W, n_sel, h0, head_div, Hsel;
  int row_chunks;                                       // 16-byte chunks per row
  int rows_per_cta;
};

// grid (ceil(C / rows_per_cta), Hsel, B); a row is copied by `row_chunks` adjacent threads so one
// warp moves 32 x 16 B = two 256-byte bf16 rows per instruction, fully coalesced on both sides.
__global__ void __launch_bounds__(256)
gather_kv_kernel(GatherKvArgs a) {
  pdl_enter();
  const int hs = blockIdx.y, b = blockIdx.z;
  const int C = a.n_sel + a.W;
  const int lane_chunk = threadIdx.x % a.row_chunks;
  const int row_in_cta = threadIdx.x / a.row_chunks;
  const int rows_par = blockDim.x / a.row_chunks;
  const int src_head = (a.h0 + hs) / a.head_div;
  const int64_t* idx = a.idx + (int64_t)b * a.idx_sb + (int64_t)hs * a.idx_sh;
  const char* kb = a.k + b * a.k_sb + src_head * a.k_sh;
  const char* vb = a.v + b * a.v_sb + src_head * a.v_sh;
  char* kob = a.ko + b * a.ko_sb + hs * a.ko_sh;
  char* vob = a.vo + b * a.vo_sb + hs * a.vo_sh;
  const int r0 = blockIdx.x * a.rows_per_cta;
  const int r1 = min(r0 + a.rows_per_cta, C);
  // 4 rows per thread per round: all index loads, then all 8 payload loads, then the 8 stores, so
  // every thread keeps 128 bytes in flight (the copy is pure latency otherwise)
  for (int rb = r0 + row_in_cta; rb < r1; rb += rows_par * 4) {
    int64_t src_row[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = rb + u * rows_par;
      src_row[u] = (r < r1) ? ((r < a.n_sel) ? idx[r] : (int64_t)(a.L - a.W + (r - a.n_sel))) : -1;
    }
    uint4 kv[4], vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (src_row[u] >= 0) {
        kv[u] = ld_stream16(kb + src_row[u] * a.k_sl + lane_chunk * 16);
        vv[u] = ld_stream16(vb + src_row[u] * a.v_sl + lane_chunk * 16);
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

struct GatherRowsArgs {
  const char* src; char* out;
  int64_t src_sb, src_sl, out_sb, out_sl, row_bytes;
  const int64_t* idx; int64_t idx_sb;
  const int64_t* pos_in; int64_t pos_sb; int64_t* pos_out; int64_t pos_out_sb;
  int T;
};

// grid (T, B): one CTA per destination row; 8 KB bf16 hidden rows -> 512 x 16 B, two per thread.
__global__ void __launch_bounds__(256)
gather_rows_kernel(GatherRowsArgs a) {
  pdl_enter();
  const int i = blockIdx.x, b = blockIdx.y;
  const int64_t src_row = a.idx[(int64_t)b * a.idx_sb + i];
  const char* s = a.src + b * a.src_sb + src_row * a.src_sl;
  char* d = a.out + b * a.out_sb + (int64_t)i * a.out_sl;
  const int chunks = (int)(a.row_bytes / 16);
  for (int c = threadIdx.x; c < chunks; c += blockDim.x) st_stream16(d + c * 16, ld_stream16(s + c * 16));
  if (a.pos_in != nullptr && threadIdx.x == 0)
    a.pos_out[(int64_t)b * a.pos_out_sb + i] = a.pos_in[(int64_t)b * a.pos_sb + src_row];
}

}  // namespace

int launch_gather_kv(const fkv_gather_kv_params& p, cudaStream_t s) {
  const int es = dtype_size(p.dtype);
  GatherKvArgs a{};
  a.k = (const char*)p.k.ptr; a.v = (const char*)p.v.ptr;
  a.ko = (char*)const_cast<void*>(p.k_out.ptr); a.vo = (char*)const_cast<void*>(p.v_out.ptr);
  a.k_sb = p.k.stride_b * es; a.k_sh = p.k.stride_h * es; a.k_sl = p.k.stride_l * es;
  a.v_sb = p.v.stride_b * es; a.v_sh = p.v.stride_h * es; a.v_sl = p.v.stride_l * es;
  a.ko_sb = p.k_out.stride_b * es; a.ko_sh = p.k_out.stride_h * es; a.ko_sl = p.k_out.stride_l * es;
  a.vo_sb = p.v_out.stride_b * es; a.vo_sh = p.v_out.stride_h * es; a.vo_sl = p.v_out.stride_l * es;
  a.idx = p.idx; a.idx_sb = p.idx_stride_b; a.idx_sh = p.idx_stride_h;
  a.L = p.L; a.W = p.window_size; a.n_sel = p.n_sel; a.h0 = p.h0; a.head_div = p.head_div;
  a.Hsel = p.h1 - p.h0;
  a.row_chunks = p.D * es / 16;                    // 16 (bf16, D=128), 8 (bf16 D=64), 32 (f32 D=128)
  const int C = p.n_sel + p.window_size;
  // enough CTAs to cover the chip a few times over without starving each of work
  const int rows_par = 256 / a.row_chunks;
  int rows_per_cta = rows_par * 4;
  const int64_t heads = (int64_t)a.Hsel * p.B;
  while (rows_per_cta > rows_par && heads * ceil_div(C, rows_per_cta) < 2 * num_sms()) rows_per_cta -= rows_par;
  a.rows_per_cta = rows_per_cta;
  FKV_CUDA_TRY(launch_pdl(gather_kv_kernel, dim3(ceil_div(C, rows_per_cta), a.Hsel, p.B), dim3(256), 0, s, a));
  count_launch();
  return FKV_OK;
}

int launch_gather_rows(const fkv_gather_rows_params& p, cudaStream_t s) {
  GatherRowsArgs a{};
  a.src = (const char*)p.src; a.out = (char*)p.out;
  a.src_sb = p.src_stride_b_bytes; a.src_sl = p.src_stride_l_bytes;
  a.out_sb = p.out_stride_b_bytes; a.out_sl = p.out_stride_l_bytes;
  a.row_bytes = p.row_bytes;
  a.idx = p.idx; a.idx_sb = p.idx_stride_b;
  a.pos_in = p.pos_in; a.pos_sb = p.pos_stride_b; a.pos_out = p.pos_out; a.pos_out_sb = p.pos_out_stride_b;
  a.T = p.T;
  FKV_CUDA_TRY(launch_pdl(gather_rows_kernel, dim3(p.T, p.B), dim3(256), 0, s, a));
  count_launch();
  return FKV_OK;
}

}  // namespace fkv
