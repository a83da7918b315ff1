// K1/K2 "simple" path: fp32 CUDA-core scoring for every dtype (bf16 / f16 / f32 inputs).
//
// This is the numerical checker for the tensor-core path and the kernel used for fp32 inputs
// (BASELINE config 1).  It follows the reference's arithmetic order as closely as a parallel
// kernel can: logits = (q . k) / sqrt(D) in fp32, row softmax over all L keys with the window's
// causal mask, window rows summed in row order, pooling per query head, then the GQA group-sum
// (baselines/fastkv/utils.py:94-112).
//
//   pass 1  score_tile_kernel<T, kStats>   per 128-key tile: per-row (max, sum exp) partials
//   reduce  stats_reduce_kernel            partials -> global (max, sum exp) per row
//   pass 2  score_tile_kernel<T, kVotes>   votes[b, hq, j] = sum_r softmax(row r)[j]
//   pool    pool_kernel                    score[b, h, j]  = sum_g pool1d(votes[b, hG+g, :])[j]
#include "common.cuh"

namespace fkv {
namespace {

constexpr int kTileKeys = 128;    // keys per CTA tile == threads per CTA (one key per thread)
constexpr int kRowChunk = 32;     // window x group rows handled per register pass

enum { kStats = 0, kVotes = 1, kRaw = 2 };

struct SimpleArgs {
  const void* q; int64_t q_sb, q_sh, q_sl;
  const void* k; int64_t k_sb, k_sh, k_sl;
  int B, L, D, W, P;
  int mode;
  int h0;            // first selection head
  int G;             // query heads per selection head handled by one CTA (fastkv: Hq/Hkv, snapkv: 1)
  int kv_div;        // kv head = selection head / kv_div           (snapkv: Hq/Hkv, else 1)
  int q_mul;         // first query head = selection head * q_mul    (fastkv: G, snapkv: 1)
  int qh0, qh1;      // gemfilter: query-head window
  int n_rows;        // rows per selection head: G*W (gemfilter: query heads of this kv head in window)
  int n_tiles;
  float2* partial;   // [B][Hsel][n_tiles][n_rows]
  float2* stats;     // [B][Hsel][n_rows]
  float* votes;      // [B][Hsel*G][P]
  int Hsel;
};

template <typename T, int PASS>
__global__ void __launch_bounds__(kTileKeys)
score_tile_kernel(SimpleArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int D = a.D;
  const int kstride = D + 4;                       // floats; 16-byte rows, conflict-free LDS.128
  float* ks = smem;                                // [128][D+4]
  float* qs = ks + kTileKeys * kstride;            // [32][D]
  float* lg = qs + kRowChunk * D;                  // [32][129]   (stats pass only)

  const int tile = blockIdx.x, hs = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x;
  const int sel_head = a.h0 + hs;
  const int kv_head = sel_head / a.kv_div;
  const int key0 = tile * kTileKeys;

  // ---- stage the K tile as fp32 -------------------------------------------------------------
  {
    const T* kbase = reinterpret_cast<const T*>(a.k) + (int64_t)b * a.k_sb + (int64_t)kv_head * a.k_sh;
    const int chunks_per_row = D / 8;
    for (int c = tid; c < kTileKeys * chunks_per_row; c += kTileKeys) {
      const int row = c / chunks_per_row, col = (c % chunks_per_row) * 8;
      float v[8];
      if (key0 + row < a.L) {
        Elem<T>::load8(kbase + (int64_t)(key0 + row) * a.k_sl + col, v);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
      }
      float4* dst = reinterpret_cast<float4*>(ks + row * kstride + col);
      dst[0] = make_float4(v[0], v[1], v[2], v[3]);
      dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }

  const int key = key0 + tid;
  const float inv_scale = sqrtf((float)D);
  float carry = 0.f;                                // running window-row sum for the current query head

  for (int rc = 0; rc < a.n_rows; rc += kRowChunk) {
    __syncthreads();                                // K tile visible / previous chunk's qs+lg consumed
    // ---- stage up to 32 query rows as fp32 ----------------------------------------------------
    {
      const int chunks_per_row = D / 8;
      for (int c = tid; c < kRowChunk * chunks_per_row; c += kTileKeys) {
        const int i = c / chunks_per_row, col = (c % chunks_per_row) * 8;
        const int n = rc + i;
        float v[8];
#pragma unroll
        for (int x = 0; x < 8; ++x) v[x] = 0.f;
        if (n < a.n_rows) {
          int hq, qrow;
          if (a.mode == FKV_MODE_GEMFILTER) {
            const int lo = max(a.qh0, kv_head * a.G);
            hq = lo + n; qrow = a.L - 1;
          } else {
            hq = sel_head * a.q_mul + n / a.W; qrow = a.P + n % a.W;
          }
          const T* qp = reinterpret_cast<const T*>(a.q) + (int64_t)b * a.q_sb + (int64_t)hq * a.q_sh +
                        (int64_t)qrow * a.q_sl + col;
          Elem<T>::load8(qp, v);
        }
        float4* dst = reinterpret_cast<float4*>(qs + i * D + col);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    __syncthreads();

    // ---- 32 dot products per key, d ascending ----------------------------------------------------
    float acc[kRowChunk];
#pragma unroll
    for (int i = 0; i < kRowChunk; ++i) acc[i] = 0.f;
    const float4* krow = reinterpret_cast<const float4*>(ks + tid * kstride);
    for (int d4 = 0; d4 < D / 4; ++d4) {
      const float4 kv = krow[d4];
#pragma unroll
      for (int i = 0; i < kRowChunk; ++i) {
        const float4 qv = reinterpret_cast<const float4*>(qs + i * D)[d4];
        acc[i] = fmaf(kv.x, qv.x, acc[i]);
        acc[i] = fmaf(kv.y, qv.y, acc[i]);
        acc[i] = fmaf(kv.z, qv.z, acc[i]);
        acc[i] = fmaf(kv.w, qv.w, acc[i]);
      }
    }

    if (PASS == kRaw) {
      // GemFilter: raw inner products, no scale / softmax / mask (gemfilter/utils.py:26-27)
      if (key < a.L) {
        const int lo = max(a.qh0, kv_head * a.G);
#pragma unroll
        for (int i = 0; i < kRowChunk; ++i) {
          const int n = rc + i;
          if (n < a.n_rows) {
            const int ql = lo + n - a.qh0;          // local query-head slot
            a.votes[((int64_t)b * (a.qh1 - a.qh0) + ql) * a.P + key] = acc[i];
          }
        }
      }
    } else if (PASS == kStats) {
#pragma unroll
      for (int i = 0; i < kRowChunk; ++i) {
        const int n = rc + i;
        const int r = n % a.W;
        float s = acc[i] / inv_scale;
        const bool visible = (key < a.L) && (key < a.P || key - a.P <= r);
        lg[i * (kTileKeys + 1) + tid] = visible ? s : -INFINITY;
      }
      __syncthreads();
      const int warp = tid >> 5, lane = tid & 31;
      for (int i = warp; i < kRowChunk; i += kTileKeys / 32) {
        const int n = rc + i;
        if (n >= a.n_rows) break;
        float v[4];
        float m = -INFINITY;
#pragma unroll
        for (int x = 0; x < 4; ++x) { v[x] = lg[i * (kTileKeys + 1) + lane + 32 * x]; m = fmaxf(m, v[x]); }
        m = warp_max(m);
        float z = 0.f;
        if (m > -INFINITY) {
#pragma unroll
          for (int x = 0; x < 4; ++x) z += expf(v[x] - m);
        }
        z = warp_sum(z);
        if (lane == 0)
          a.partial[(((int64_t)b * a.Hsel + hs) * a.n_tiles + tile) * a.n_rows + n] = make_float2(m, z);
      }
    } else {  // kVotes
      const float2* st = a.stats + ((int64_t)b * a.Hsel + hs) * a.n_rows;
#pragma unroll
      for (int i = 0; i < kRowChunk; ++i) {
        const int n = rc + i;
        if (n < a.n_rows) {
          const int g = n / a.W, r = n % a.W;
          const float2 mz = st[n];
          const float s = acc[i] / inv_scale;
          const float p = expf(s - mz.x) / mz.y;
          carry = (r == 0) ? p : carry + p;        // sum(dim=-2): rows in order (utils.py:104)
          if (r == a.W - 1 && key < a.P)
            a.votes[((int64_t)b * a.Hsel * a.G + (int64_t)hs * a.G + g) * a.P + key] = carry;
        }
      }
    }
  }
}

// partial (max, sumexp) per tile -> global per row.  grid (Hsel, B), one thread per row.
__global__ void stats_reduce_kernel(const float2* __restrict__ partial, float2* __restrict__ stats,
                                    int n_tiles, int n_rows, int Hsel) {
  const int hs = blockIdx.x, b = blockIdx.y;
  const float2* base = partial + ((int64_t)b * Hsel + hs) * n_tiles * n_rows;
  for (int n = threadIdx.x; n < n_rows; n += blockDim.x) {
    float m = -INFINITY;
    for (int t = 0; t < n_tiles; ++t) m = fmaxf(m, base[(int64_t)t * n_rows + n].x);
    float z = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const float2 v = base[(int64_t)t * n_rows + n];
      if (v.x > -INFINITY) z += v.y * expf(v.x - m);
    }
    stats[((int64_t)b * Hsel + hs) * n_rows + n] = make_float2(m, z);
  }
}

// score[b,h,j] = sum_g pool1d(votes[b, h*G+g, :])[j]          (fastkv / snapkv)
// score[b,0,j] = avgpool1d( sum_q votes[b, q, :] )[j]         (gemfilter: head-sum before the pool)
__global__ void pool_kernel(const float* __restrict__ votes, float* __restrict__ score, int64_t score_sb,
                            int64_t score_sh, int P, int G, int Hsel, int ks, int pooling, int gem) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int hs = blockIdx.y, b = blockIdx.z;
  if (j >= P) return;
  const int half = ks / 2;
  float out;
  if (gem) {
    const float* v = votes + (int64_t)b * G * P;
    float sum = 0.f;
    for (int d = -half; d <= half; ++d) {
      const int jj = j + d;
      if (jj < 0 || jj >= P) continue;
      float hsum = 0.f;
      for (int q = 0; q < G; ++q) hsum += v[(int64_t)q * P + jj];      // torch.sum(dim=1), head order
      sum += hsum;
    }
    out = sum / (float)ks;
  } else {
    out = 0.f;
    for (int g = 0; g < G; ++g) {
      const float* v = votes + ((int64_t)b * Hsel * G + (int64_t)hs * G + g) * P;
      float pooled;
      if (pooling == FKV_POOL_AVG) {
        float sum = 0.f;
        for (int d = -half; d <= half; ++d) {
          const int jj = j + d;
          if (jj >= 0 && jj < P) sum += v[jj];
        }
        pooled = sum / (float)ks;                  // count_include_pad: always / kernel_size
      } else {
        pooled = -INFINITY;
        for (int d = -half; d <= half; ++d) {
          const int jj = j + d;
          if (jj >= 0 && jj < P) pooled = fmaxf(pooled, v[jj]);
        }
      }
      out += pooled;                               // view(.., G, P).sum(dim=-2): g in order (utils.py:112)
    }
  }
  score[(int64_t)b * score_sb + (int64_t)hs * score_sh + j] = out;
}

template <typename T>
int run_simple(const fkv_score_params& p, cudaStream_t s) {
  SimpleArgs a{};
  a.q = p.q.ptr; a.q_sb = p.q.stride_b; a.q_sh = p.q.stride_h; a.q_sl = p.q.stride_l;
  a.k = p.k.ptr; a.k_sb = p.k.stride_b; a.k_sh = p.k.stride_h; a.k_sl = p.k.stride_l;
  a.B = p.B; a.L = p.L; a.D = p.D; a.mode = p.mode;
  const int Gm = p.Hq / p.Hkv;
  const int n_tiles = ceil_div(p.L, kTileKeys);
  a.n_tiles = n_tiles;
  const size_t smem = (size_t)(kTileKeys * (p.D + 4) + kRowChunk * p.D + kRowChunk * (kTileKeys + 1)) * sizeof(float);
  {                                                  // once per (device, kernel)
    const int smem_max = (int)((kTileKeys * (128 + 4) + kRowChunk * 128 + kRowChunk * (kTileKeys + 1)) * sizeof(float));
    int rc_attr;
    if ((rc_attr = ensure_dyn_smem(score_tile_kernel<T, kStats>, smem_max)) != FKV_OK) return rc_attr;
    if ((rc_attr = ensure_dyn_smem(score_tile_kernel<T, kVotes>, smem_max)) != FKV_OK) return rc_attr;
    if ((rc_attr = ensure_dyn_smem(score_tile_kernel<T, kRaw>, smem_max)) != FKV_OK) return rc_attr;
  }
  char* ws = reinterpret_cast<char*>(p.workspace) + FKV_WORKSPACE_HEADER_BYTES;   // past the library's persistent region

  if (p.mode == FKV_MODE_GEMFILTER) {
    const int qh0 = p.qh0, qh1 = (p.qh1 > p.qh0) ? p.qh1 : p.Hq;
    a.W = 1; a.P = p.L; a.G = Gm; a.kv_div = 1; a.q_mul = Gm; a.qh0 = qh0; a.qh1 = qh1;
    a.votes = reinterpret_cast<float*>(ws);
    a.Hsel = 1;
    // one launch per kv head touched by [qh0,qh1): n_rows differs at the window edges
    for (int kvh = qh0 / Gm; kvh * Gm < qh1; ++kvh) {
      const int lo = max(qh0, kvh * Gm), hi = min(qh1, (kvh + 1) * Gm);
      a.h0 = kvh; a.n_rows = hi - lo;
      score_tile_kernel<T, kRaw><<<dim3(n_tiles, 1, p.B), kTileKeys, smem, s>>>(a);
      FKV_LAUNCH_CHECK();
    }
    pool_kernel<<<dim3(ceil_div(p.L, 256), 1, p.B), 256, 0, s>>>(a.votes, p.score, p.score_stride_b, p.score_stride_h,
                                                                p.L, qh1 - qh0, 1, p.kernel_size, FKV_POOL_AVG, 1);
    FKV_LAUNCH_CHECK();
    return FKV_OK;
  }

  const int Hsel = p.h1 - p.h0;
  a.W = p.window_size; a.P = p.L - p.window_size; a.h0 = p.h0; a.Hsel = Hsel;
  if (p.mode == FKV_MODE_FASTKV) { a.G = Gm; a.kv_div = 1; a.q_mul = Gm; }
  else { a.G = 1; a.kv_div = Gm; a.q_mul = 1; }
  a.n_rows = a.G * a.W;
  size_t off = 0;
  a.partial = reinterpret_cast<float2*>(ws + off);
  off = align_up(off + (size_t)p.B * Hsel * n_tiles * a.n_rows * sizeof(float2), 256);
  a.stats = reinterpret_cast<float2*>(ws + off);
  off = align_up(off + (size_t)p.B * Hsel * a.n_rows * sizeof(float2), 256);
  a.votes = reinterpret_cast<float*>(ws + off);

  const dim3 grid(n_tiles, Hsel, p.B);
  score_tile_kernel<T, kStats><<<grid, kTileKeys, smem, s>>>(a);
  FKV_LAUNCH_CHECK();
  stats_reduce_kernel<<<dim3(Hsel, p.B), 128, 0, s>>>(a.partial, a.stats, n_tiles, a.n_rows, Hsel);
  FKV_LAUNCH_CHECK();
  score_tile_kernel<T, kVotes><<<grid, kTileKeys, smem, s>>>(a);
  FKV_LAUNCH_CHECK();
  pool_kernel<<<dim3(ceil_div(a.P, 256), Hsel, p.B), 256, 0, s>>>(a.votes, p.score, p.score_stride_b, p.score_stride_h,
                                                                  a.P, a.G, Hsel, p.kernel_size, p.pooling, 0);
  FKV_LAUNCH_CHECK();
  return FKV_OK;
}

}  // namespace

// Scratch for the simple path (also an upper bound for the tensor path, which needs less).
size_t score_workspace_bytes(int mode, int B, int Hq, int Hkv, int L, int W) {
  const int n_tiles = ceil_div(L, kTileKeys);
  if (mode == FKV_MODE_GEMFILTER) return FKV_WORKSPACE_HEADER_BYTES + align_up((size_t)B * Hq * L * sizeof(float), 256);
  const int Hsel = (mode == FKV_MODE_FASTKV) ? Hkv : Hq;
  const int G = (mode == FKV_MODE_FASTKV) ? Hq / Hkv : 1;
  size_t b = 0;
  b = align_up(b + (size_t)B * Hsel * n_tiles * G * W * sizeof(float2), 256);
  b = align_up(b + (size_t)B * Hsel * G * W * sizeof(float2), 256);
  b = align_up(b + (size_t)B * Hsel * G * (size_t)L * sizeof(float), 256);
  return b + 1024 + FKV_WORKSPACE_HEADER_BYTES;
}

int launch_score_simple(const fkv_score_params& p, cudaStream_t s) {
  switch (p.dtype) {
    case FKV_BF16: return run_simple<__nv_bfloat16>(p, s);
    case FKV_F16: return run_simple<__half>(p, s);
    case FKV_F32: return run_simple<float>(p, s);
  }
  return FKV_EDTYPE;
}

}  // namespace fkv
