// K5: split-K decode attention (q_len == 1) over the compact KV cache, GQA-aware, plus the
// in-place cache append that replaces DynamicCache.update's torch.cat.
// Replaces flash_attn_func at baselines/fastkv/llama_model.py:79-81 (decode branch :143-145).
//
//   decode_attn_kernel   grid (splits, Hq / Gt, B), ONE launch.  A CTA owns R = 64 / 128 / 256 consecutive cache rows of
//                        one kv head and serves Gt query heads of its group (K/V bytes are read once per group).
//                        All its K and V rows travel at once as 16-byte asynchronous copies (cp.async, L1 bypassed) into
//                        PADDED shared-memory rows — one DRAM round trip per CTA — and are then consumed 64 rows at a
//                        time with an online softmax: scores key-per-thread (no shuffles in the dot product), the
//                        softmax per head by one warp, P.V dimension-per-thread.
//                        Partials (max, sum, unnormalised output) go to the workspace; the last CTA of a (batch, head
//                        group) to arrive — an acq_rel ticket in the workspace's persistent region — merges them
//                        (float4 loads, splits spread over thread groups: one L2 round trip) and writes the output,
//                        so there is no second "combine" launch.  R is chosen so that the grid is a single wave and
//                        the merge short (9 splits at S = 2304, B = 1; 5 at S = 640, B = 8).
//   Measured dead ends: row-sized cp.async.bulk copies (128 per CTA serialise in the TMA unit: 20 us per call); 36 splits
//   of 64 rows merged by a scalar loop (each dependent L2 load stalls the in-order warp: 17 us per call).
#include <cstdio>

#include "common.cuh"

namespace fkv {
namespace {

constexpr int kDecThreads = 256;
constexpr int kDecChunk = 64;                         // cache rows consumed per online-softmax step
constexpr int kDecParts = kDecThreads / 32;           // score stage: one warp per slice of the head dim
constexpr int kDecMaxSplitsPerGroup = 4;              // merge: partial rows a thread keeps in flight per head

struct DecodeArgs {
  const void* q; int64_t q_sb, q_sh;
  const char* kc; int64_t k_sb, k_sh, k_sl;           // bytes
  const char* vc; int64_t v_sb, v_sh, v_sl;
  void* out; int64_t o_sb, o_sh;
  float* part_o;      // [B][Hq][splits][D]
  float2* part_ml;    // [B][Hq][splits]
  unsigned int* tickets;   // [B][Hq / Gt] arrival counters (persistent region of the workspace, zero between launches)
  const int32_t* seq_len_dev;
  int seq_len, max_seq_len, B, Hq, Hkv, G, D, splits, rows_per_cta;
  float scale_log2;   // scale * log2(e)
};

template <typename T> __device__ inline float elem_to_float(T v);
template <> __device__ inline float elem_to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ inline float elem_to_float<__half>(__half v) { return __half2float(v); }
template <> __device__ inline float elem_to_float<float>(float v) { return v; }
template <typename T> __device__ inline T float_to_elem(float v);
template <> __device__ inline __nv_bfloat16 float_to_elem<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ inline __half float_to_elem<__half>(float v) { return __float2half_rn(v); }
template <> __device__ inline float float_to_elem<float>(float v) { return v; }

__device__ __forceinline__ uint32_t dsmem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// FKV_DEC_TRACE=1: %globaltimer stamps of every CTA (debugging aid; the launcher synchronises and prints them)
__device__ unsigned long long g_dec_trace[8 * 2048];
__device__ int g_dec_trace_on = 0;
#define DEC_TRACE(i) do { if (g_dec_trace_on && threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); \
  const int c_ = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z); if (c_ < 2048) g_dec_trace[c_ * 8 + (i)] = t_; } } while (0)

// Gt = query heads served by one CTA (divides the model's group size G); D = head dim
template <typename T, int D, int Gt>
__global__ void __launch_bounds__(kDecThreads)
decode_attn_kernel(DecodeArgs a) {
  constexpr int ROWB = D * (int)sizeof(T);            // bytes per cache row
  constexpr int PADB = ROWB + 16;                     // padded shared-memory row: conflict-free 16-byte reads, key per thread
  constexpr int HALF = D / 2;                         // dims per thread in the score stage (2 threads per key)
  constexpr int CPR = ROWB / 16;                      // 16-byte chunks per row
  extern __shared__ __align__(128) unsigned char smem[];
  const int R = a.rows_per_cta;
  unsigned char* ks = smem;                           // [R][PADB]
  unsigned char* vs = ks + (size_t)R * PADB;          // [R][PADB]
  float* qs = reinterpret_cast<float*>(vs + (size_t)R * PADB);  // [Gt][D] scaled query, fp32 (later: merge scratch)
  float* sc = qs + Gt * D;                            // [8 dim parts][Gt][64] partial dot products (later: merge scratch)
  float* pT = sc + kDecParts * Gt * kDecChunk;        // [64][Gt] probabilities of the current step, heads contiguous
  float* red = pT + kDecChunk * Gt;                   // [Gt][4]: running max, running sum, rescale factor of this step, -
  __shared__ unsigned int s_last;

  const int split = blockIdx.x, hg = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  DEC_TRACE(0);
  int S = a.seq_len_dev ? *a.seq_len_dev : a.seq_len;
  S = min(S, a.max_seq_len);                          // never read past the rows the cache views really have
  const int nblocks = (S + R - 1) / R;                // R-row blocks of this (batch, head group)
  const int active = min(nblocks, a.splits);          // CTAs that have rows (block j goes to CTA j mod splits)
  if (split >= active) return;
  const int hq0 = hg * Gt, kvh = hq0 / a.G;
  constexpr int KPARTS = kDecThreads / D;             // P.V: 2 (D = 128) or 4 (D = 64) key parts per output dimension
  const int d = tid % D, kpart = tid / D;
  float o[Gt];
#pragma unroll
  for (int g = 0; g < Gt; ++g) o[g] = 0.f;
  if (tid < Gt) { red[4 * tid] = -INFINITY; red[4 * tid + 1] = 0.f; }

  for (int blk = split; blk < nblocks; blk += a.splits) {       // one trip unless the cache is longer than splits * R rows
  const int s0 = blk * R, rows = min(R, S - s0);
  {
    // a thread always copies the same 16-byte column, kDecThreads / CPR rows apart: two adds per copy
    constexpr int RSTEP = kDecThreads / CPR;          // rows covered by one pass of the CTA
    const int col = tid % CPR, r_first = tid / CPR;
    const char* ksrc = a.kc + b * a.k_sb + kvh * a.k_sh + (int64_t)(s0 + r_first) * a.k_sl + col * 16;
    const char* vsrc = a.vc + b * a.v_sb + kvh * a.v_sh + (int64_t)(s0 + r_first) * a.v_sl + col * 16;
    uint32_t kdst = dsmem_u32(ks + (size_t)r_first * PADB + col * 16);
    uint32_t vdst = dsmem_u32(vs + (size_t)r_first * PADB + col * 16);
    const int64_t kstep = (int64_t)RSTEP * a.k_sl, vstep = (int64_t)RSTEP * a.v_sl;
    for (int r = r_first; r < rows; r += RSTEP) {
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(kdst), "l"(ksrc) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(vdst), "l"(vsrc) : "memory");
      ksrc += kstep; vsrc += vstep; kdst += RSTEP * PADB; vdst += RSTEP * PADB;
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  DEC_TRACE(1);
  // meanwhile: the Gt query rows, scaled, as fp32
  if (blk == split) {
    for (int i = tid; i < Gt * D; i += kDecThreads) {
      const int g = i / D, dd = i - g * D;
      const T* qp = reinterpret_cast<const T*>(a.q) + (int64_t)b * a.q_sb + (int64_t)(hq0 + g) * a.q_sh + dd;
      qs[i] = elem_to_float<T>(*qp) * a.scale_log2;
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  DEC_TRACE(2);

  for (int c0 = 0; c0 < rows; c0 += kDecChunk) {
    const int crow = min(kDecChunk, rows - c0);
    // ---- scores: warp = slice of the head dim, lane = key pair (k, k + 32): q is read once for two keys, K rows as
    // conflict-free 16-byte chunks; no shuffles ------------------------------------------------------------------
    {
      constexpr int PD = D / kDecParts;               // dims per warp: 16 (D = 128) or 8 (D = 64)
      constexpr int E = 16 / (int)sizeof(T);          // elements per 16 bytes
      constexpr int NCH = PD / E > 0 ? PD / E : 1;    // 16-byte chunks per key and slice
      float acc0[Gt], acc1[Gt];
#pragma unroll
      for (int g = 0; g < Gt; ++g) { acc0[g] = 0.f; acc1[g] = 0.f; }
      const bool ok0 = lane < crow, ok1 = lane + 32 < crow;
      const unsigned char* kr0 = ks + (size_t)(c0 + lane) * PADB + warp * PD * (int)sizeof(T);
      const unsigned char* kr1 = kr0 + (size_t)32 * PADB;
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        uint4 raw0 = make_uint4(0u, 0u, 0u, 0u), raw1 = make_uint4(0u, 0u, 0u, 0u);
        if (ok0) raw0 = *reinterpret_cast<const uint4*>(kr0 + c * 16);
        if (ok1) raw1 = *reinterpret_cast<const uint4*>(kr1 + c * 16);
        const T* k0t = reinterpret_cast<const T*>(&raw0);
        const T* k1t = reinterpret_cast<const T*>(&raw1);
        float kf0[E], kf1[E];
#pragma unroll
        for (int e = 0; e < E; ++e) { kf0[e] = elem_to_float<T>(k0t[e]); kf1[e] = elem_to_float<T>(k1t[e]); }
#pragma unroll
        for (int g = 0; g < Gt; ++g) {
          const float* qg = qs + g * D + warp * PD + c * E;
#pragma unroll
          for (int e = 0; e < E; e += 4) {
            const float4 q4 = *reinterpret_cast<const float4*>(qg + e);
            acc0[g] = fmaf(kf0[e], q4.x, acc0[g]); acc0[g] = fmaf(kf0[e + 1], q4.y, acc0[g]);
            acc0[g] = fmaf(kf0[e + 2], q4.z, acc0[g]); acc0[g] = fmaf(kf0[e + 3], q4.w, acc0[g]);
            acc1[g] = fmaf(kf1[e], q4.x, acc1[g]); acc1[g] = fmaf(kf1[e + 1], q4.y, acc1[g]);
            acc1[g] = fmaf(kf1[e + 2], q4.z, acc1[g]); acc1[g] = fmaf(kf1[e + 3], q4.w, acc1[g]);
          }
        }
      }
#pragma unroll
      for (int g = 0; g < Gt; ++g) {
        sc[(warp * Gt + g) * kDecChunk + lane] = acc0[g];
        sc[(warp * Gt + g) * kDecChunk + lane + 32] = acc1[g];
      }
    }
    __syncthreads();
    // ---- online softmax over this step's keys, one warp per head (round-robin) -----------------------------------
    for (int g = warp; g < Gt; g += kDecParts) {
      float x0 = 0.f, x1 = 0.f;
#pragma unroll
      for (int pz = 0; pz < kDecParts; ++pz) {
        x0 += sc[(pz * Gt + g) * kDecChunk + lane];
        x1 += sc[(pz * Gt + g) * kDecChunk + lane + 32];
      }
      if (lane >= crow) x0 = -INFINITY;
      if (lane + 32 >= crow) x1 = -INFINITY;
      float m_old = (lane == 0) ? red[4 * g] : 0.f;   // one lane reads and later rewrites the running max: no intra-warp hazard
      m_old = __shfl_sync(0xffffffffu, m_old, 0);
      const float m = fmaxf(m_old, warp_max(fmaxf(x0, x1)));
      const float p0 = exp2f(x0 - m), p1 = exp2f(x1 - m);        // exp2(-inf) = 0 for the missing rows
      const float alpha = exp2f(m_old - m);                       // first step: exp2(-inf) = 0
      const float l = warp_sum(p0 + p1);
      pT[lane * Gt + g] = p0;
      pT[(lane + 32) * Gt + g] = p1;
      if (lane == 0) { red[4 * g] = m; red[4 * g + 1] = red[4 * g + 1] * alpha + l; red[4 * g + 2] = alpha; }
    }
    __syncthreads();
    // ---- P.V: thread = (output dimension, key part), all Gt heads; probabilities of a key are one broadcast read ----
#pragma unroll
    for (int g = 0; g < Gt; ++g) o[g] *= red[4 * g + 2];
    for (int k = kpart; k < crow; k += KPARTS) {
      const float v = elem_to_float<T>(*reinterpret_cast<const T*>(vs + (size_t)(c0 + k) * PADB + d * (int)sizeof(T)));
      if constexpr (Gt % 4 == 0) {
#pragma unroll
        for (int g = 0; g < Gt; g += 4) {
          const float4 p4 = *reinterpret_cast<const float4*>(pT + k * Gt + g);
          o[g] = fmaf(p4.x, v, o[g]); o[g + 1] = fmaf(p4.y, v, o[g + 1]);
          o[g + 2] = fmaf(p4.z, v, o[g + 2]); o[g + 3] = fmaf(p4.w, v, o[g + 3]);
        }
      } else {
#pragma unroll
        for (int g = 0; g < Gt; ++g) o[g] = fmaf(pT[k * Gt + g], v, o[g]);
      }
    }
    __syncthreads();                                  // sc / red (and, on the next trip, the K / V rows) are rewritten next
  }
  }                                                   // blocks
  DEC_TRACE(3);
  if constexpr (KPARTS > 1) {                         // fold the key parts onto part 0 (sc is free: [KPARTS][Gt][D] fits it)
    if (kpart > 0) {
#pragma unroll
      for (int g = 0; g < Gt; ++g) sc[((kpart - 1) * Gt + g) * D + d] = o[g];
    }
    __syncthreads();
    if (kpart == 0) {
#pragma unroll
      for (int pz = 0; pz < KPARTS - 1; ++pz)
#pragma unroll
        for (int g = 0; g < Gt; ++g) o[g] += sc[(pz * Gt + g) * D + d];
    }
    __syncthreads();
  }
  const bool writer = (kpart == 0);

  if (active == 1) {                                  // a single CTA has all the rows: no partials, no merge
    if (writer) {
#pragma unroll
      for (int g = 0; g < Gt; ++g) {
        T* op = reinterpret_cast<T*>(a.out) + (int64_t)b * a.o_sb + (int64_t)(hq0 + g) * a.o_sh + d;
        *op = float_to_elem<T>(o[g] / red[4 * g + 1]);
      }
    }
    return;
  }
  // ---- partials out; the last CTA of this (batch, head group) to arrive merges them -----------------------------
  if (writer) {
#pragma unroll
    for (int g = 0; g < Gt; ++g) {
      const int64_t slot = ((int64_t)b * a.Hq + hq0 + g) * a.splits + split;
      a.part_o[slot * D + d] = o[g];
      if (d == 0) a.part_ml[slot] = make_float2(red[4 * g], red[4 * g + 1]);
    }
  }
  __syncthreads();
  if (tid == 0) {
    // one acq_rel ticket by one thread publishes the whole CTA's partials (ordered before it by the barrier) and,
    // for the last arrival, makes every other CTA's partials visible to the whole CTA after the next barrier
    unsigned int* ticket = a.tickets + (size_t)b * gridDim.y + hg;
    unsigned int prev;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(ticket) : "memory");
    s_last = (prev == (unsigned)active - 1u) ? 1u : 0u;
    if (s_last) *ticket = 0u;                         // clean for the next launch (stream-ordered)
  }
  __syncthreads();
  DEC_TRACE(4);
  if (!s_last) return;

  // ---- merge (one CTA per batch x head group).  Every load is issued before its first use: a dependent FMA would
  // stall the in-order warp on each L2 round trip.  Thread = (float4 column of D, split group): D/4 columns x SG groups.
  constexpr int COLS = D / 4, SG = kDecThreads / COLS;           // 32 x 8 (D = 128), 16 x 16 (D = 64)
  constexpr int MAXS = kDecMaxSplitsPerGroup;                     // splits per thread group held in registers (host: splits <= SG * MAXS)
  float2* mlbuf = reinterpret_cast<float2*>(sc);                  // [Gt][active] (max, sum) of every split
  for (int i = tid; i < Gt * active; i += kDecThreads) {
    const int g = i / active, sp = i - g * active;
    mlbuf[i] = __ldcg(&a.part_ml[((int64_t)b * a.Hq + hq0 + g) * a.splits + sp]);
  }
  const int col = tid % COLS, sg = tid / COLS;
  float4 pv[Gt][MAXS];
#pragma unroll
  for (int g = 0; g < Gt; ++g)
#pragma unroll
    for (int u = 0; u < MAXS; ++u) {
      const int sp = sg + u * SG;
      pv[g][u] = (sp < active) ? __ldcg(reinterpret_cast<const float4*>(a.part_o + (((int64_t)b * a.Hq + hq0 + g) * a.splits + sp) * D) + col)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  __syncthreads();
  float4* accbuf = reinterpret_cast<float4*>(ks);                 // [SG][Gt][COLS] partial sums (the K rows are dead now)
#pragma unroll
  for (int g = 0; g < Gt; ++g) {
    float mm = -INFINITY;
    for (int sp = 0; sp < active; ++sp) mm = fmaxf(mm, mlbuf[g * active + sp].x);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < MAXS; ++u) {
      const int sp = sg + u * SG;
      if (sp < active) {
        const float f = exp2f(mlbuf[g * active + sp].x - mm);
        acc.x = fmaf(pv[g][u].x, f, acc.x); acc.y = fmaf(pv[g][u].y, f, acc.y);
        acc.z = fmaf(pv[g][u].z, f, acc.z); acc.w = fmaf(pv[g][u].w, f, acc.w);
      }
    }
    accbuf[(sg * Gt + g) * COLS + col] = acc;
  }
  __syncthreads();
  if (tid < D) {                                      // thread = output dimension again
#pragma unroll
    for (int g = 0; g < Gt; ++g) {
      float mm = -INFINITY, ll = 0.f;
      for (int sp = 0; sp < active; ++sp) mm = fmaxf(mm, mlbuf[g * active + sp].x);
      for (int sp = 0; sp < active; ++sp) ll = fmaf(mlbuf[g * active + sp].y, exp2f(mlbuf[g * active + sp].x - mm), ll);
      float acc = 0.f;
#pragma unroll
      for (int q2 = 0; q2 < SG; ++q2) acc += reinterpret_cast<const float*>(accbuf + (q2 * Gt + g) * COLS)[tid];
      T* op = reinterpret_cast<T*>(a.out) + (int64_t)b * a.o_sb + (int64_t)(hq0 + g) * a.o_sh + tid;
      *op = float_to_elem<T>(acc / ll);
    }
  }
  DEC_TRACE(5);
}

template <typename T, int D, int Gt>
int run_decode(const DecodeArgs& a, cudaStream_t s) {
  constexpr int PADB = D * (int)sizeof(T) + 16;
  constexpr int SG = kDecThreads / (D / 4);
  if (a.splits > SG * kDecMaxSplitsPerGroup) return FKV_EINVAL;
  const size_t smem = (size_t)2 * a.rows_per_cta * PADB + (size_t)Gt * D * 4 + (size_t)kDecParts * Gt * kDecChunk * 4 +
                      (size_t)kDecChunk * Gt * 4 + (size_t)Gt * 4 * 4 + 64;
  int rc;
  if ((rc = ensure_dyn_smem(decode_attn_kernel<T, D, Gt>, 200 * 1024)) != FKV_OK) return rc;
  static const bool trace_on = getenv("FKV_DEC_TRACE") != nullptr;
  if (trace_on) {
    const int one = 1;
    cudaMemcpyToSymbol(g_dec_trace_on, &one, sizeof(one));
    static unsigned long long zero[8 * 2048];
    cudaMemcpyToSymbol(g_dec_trace, zero, sizeof(zero));
  }
  FKV_CUDA_TRY(launch_pdl(decode_attn_kernel<T, D, Gt>, dim3(a.splits, a.Hq / Gt, a.B), dim3(kDecThreads), smem, s, a));
  count_launch();
  if (trace_on) {
    cudaStreamSynchronize(s);
    static unsigned long long h[8 * 2048];
    cudaMemcpyFromSymbol(h, g_dec_trace, sizeof(h));
    const int n = min(2048, a.splits * (a.Hq / Gt) * a.B);
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < n; ++c) if (h[c * 8] && h[c * 8] < t0) t0 = h[c * 8];
    const char* nm[6] = {"start", "copies issued", "rows landed", "computed", "ticket", "merged (last CTA)"};
    fprintf(stderr, "[decode trace] grid %d x %d x %d, rows/CTA %d, smem %zu\n", a.splits, a.Hq / Gt, a.B, a.rows_per_cta, smem);
    for (int st = 0; st < 6; ++st) {
      unsigned long long mn = ~0ull, mx = 0; double avg = 0; int cnt = 0;
      for (int c = 0; c < n; ++c) { if (!h[c * 8 + st]) continue; const unsigned long long v = h[c * 8 + st] - t0; mn = v < mn ? v : mn; mx = v > mx ? v : mx; avg += v; ++cnt; }
      if (cnt) fprintf(stderr, "[decode trace]   %-18s min %6.2f avg %6.2f max %6.2f us (%d CTAs)\n", nm[st], mn / 1e3, avg / cnt / 1e3, mx / 1e3, cnt);
    }
  }
  return FKV_OK;
}

template <typename T, int D>
int dispatch_group(const DecodeArgs& a, cudaStream_t s) {
  // query heads per CTA: the whole group when it is 1, 2, 4 or 8 wide, otherwise its largest such divisor
  int gt = 8;
  while (gt > 1 && (a.G % gt)) gt >>= 1;
  switch (gt) {
    case 1: return run_decode<T, D, 1>(a, s);
    case 2: return run_decode<T, D, 2>(a, s);
    case 4: return run_decode<T, D, 4>(a, s);
    case 8: return run_decode<T, D, 8>(a, s);
  }
  return FKV_EINVAL;
}

template <typename T>
int dispatch_dim(const DecodeArgs& a, cudaStream_t s) {
  if (a.D == 128) return dispatch_group<T, 128>(a, s);
  if (a.D == 64) return dispatch_group<T, 64>(a, s);
  return FKV_EHEADDIM;
}

// rows per CTA: the larger of 128 / 64 whose grid still fits one wave, so the merge stays short; at most
// `max_splits` CTAs per (batch, head group): longer caches make a CTA loop over several blocks
int pick_rows_per_cta(int groups, int bound, int row_bytes, int max_splits, int* splits_out) {
  const int sms = num_sms();
  int best = 64;
  for (int r = 128; r >= 64; r >>= 1) {              // (256 rows per CTA was measured slower: half the CTAs, same latency chain)
    const size_t smem = (size_t)2 * r * (row_bytes + 16) + 8192;
    const int per_sm = (int)((size_t)220 * 1024 / smem);
    if (per_sm < 1) continue;
    const long long ctas = (long long)groups * min(ceil_div(bound, r), max_splits);
    if (ctas <= (long long)sms * per_sm || r == 64) { best = r; break; }
  }
  int splits = ceil_div(bound, best);
  if (splits > max_splits) splits = max_splits;
  *splits_out = splits;
  return best;
}

struct AppendArgs {
  const char* kn; const char* vn; char* kc; char* vc;
  int64_t kn_sb, kn_sh, vn_sb, vn_sh, kc_sb, kc_sh, kc_sl, vc_sb, vc_sh, vc_sl;
  int32_t* seq_len_dev; int seq_len, max_seq_len, row_bytes;
};

__global__ void cache_append_kernel(AppendArgs a) {
  const int h = blockIdx.x, b = blockIdx.y;
  const int S = a.seq_len_dev ? *a.seq_len_dev : a.seq_len;
  if (S >= a.max_seq_len) return;                    // full cache: drop (host checks capacity)
  for (int c = threadIdx.x; c < a.row_bytes / 16; c += blockDim.x) {
    const uint4 kv = *reinterpret_cast<const uint4*>(a.kn + b * a.kn_sb + h * a.kn_sh + c * 16);
    const uint4 vv = *reinterpret_cast<const uint4*>(a.vn + b * a.vn_sb + h * a.vn_sh + c * 16);
    *reinterpret_cast<uint4*>(a.kc + b * a.kc_sb + h * a.kc_sh + (int64_t)S * a.kc_sl + c * 16) = kv;
    *reinterpret_cast<uint4*>(a.vc + b * a.vc_sb + h * a.vc_sh + (int64_t)S * a.vc_sl + c * 16) = vv;
  }
}
// a full cache dropped the row: the length must not move past the rows that exist
__global__ void bump_kernel(int32_t* p, int max_seq_len) { if (*p < max_seq_len) *p += 1; }

}  // namespace

size_t decode_workspace_bytes(int B, int Hq, int D, int max_seq_len) {
  int splits = ceil_div(max_seq_len > 0 ? max_seq_len : 1, kDecChunk);             // upper bound (64 rows per CTA)
  if (splits > 64) splits = 64;
  return align_up((size_t)B * Hq * splits * D * sizeof(float), 256) + align_up((size_t)B * Hq * splits * sizeof(float2), 256);
}

// `p.workspace` is the caller's workspace BASE: the arrival tickets live in its persistent region, the split
// partials in the scratch past it.
int launch_decode_attn(const fkv_decode_attn_params& p, cudaStream_t s) {
  const int es = dtype_size(p.dtype);
  DecodeArgs a{};
  a.q = p.q; a.q_sb = p.q_stride_b; a.q_sh = p.q_stride_h;
  a.kc = (const char*)p.k_cache.ptr; a.k_sb = p.k_cache.stride_b * es; a.k_sh = p.k_cache.stride_h * es; a.k_sl = p.k_cache.stride_l * es;
  a.vc = (const char*)p.v_cache.ptr; a.v_sb = p.v_cache.stride_b * es; a.v_sh = p.v_cache.stride_h * es; a.v_sl = p.v_cache.stride_l * es;
  a.out = p.out; a.o_sb = p.out_stride_b; a.o_sh = p.out_stride_h;
  a.seq_len_dev = p.seq_len_dev; a.seq_len = p.seq_len;
  a.B = p.B; a.Hq = p.Hq; a.Hkv = p.Hkv; a.D = p.D; a.G = p.Hq / p.Hkv;
  // the grid covers the rows the cache can hold (a device-resident length may grow under graph replay); CTAs past the
  // current length exit at once
  const int bound = p.seq_len_dev ? p.max_seq_len : (p.max_seq_len > 0 ? min(p.seq_len, p.max_seq_len) : p.seq_len);
  a.max_seq_len = p.max_seq_len > 0 ? p.max_seq_len : p.seq_len;
  int gt0 = 8;
  while (gt0 > 1 && (a.G % gt0)) gt0 >>= 1;
  const int max_splits = (kDecThreads / (p.D / 4)) * kDecMaxSplitsPerGroup;      // what the merge holds in registers
  a.rows_per_cta = pick_rows_per_cta(p.B * (p.Hq / gt0), bound < 1 ? 1 : bound, p.D * es, max_splits, &a.splits);
  const float scale = p.scale != 0.f ? p.scale : 1.0f / sqrtf((float)p.D);
  a.scale_log2 = scale * 1.4426950408889634f;
  const size_t need_o = align_up((size_t)p.B * p.Hq * a.splits * p.D * sizeof(float), 256);
  const size_t need = FKV_WORKSPACE_HEADER_BYTES + need_o + (size_t)p.B * p.Hq * a.splits * sizeof(float2);
  if (!p.workspace || p.workspace_bytes < need) return FKV_EWORKSPACE;
  int gt = 8;
  while (gt > 1 && (a.G % gt)) gt >>= 1;
  if ((size_t)p.B * (p.Hq / gt) > kWsTickets) return FKV_EINVAL;
  char* wsb = reinterpret_cast<char*>(p.workspace);
  a.tickets = reinterpret_cast<unsigned int*>(wsb + kWsTicketOffset);
  a.part_o = reinterpret_cast<float*>(wsb + FKV_WORKSPACE_HEADER_BYTES);
  a.part_ml = reinterpret_cast<float2*>(wsb + FKV_WORKSPACE_HEADER_BYTES + need_o);
  switch (p.dtype) {
    case FKV_BF16: return dispatch_dim<__nv_bfloat16>(a, s);
    case FKV_F16: return dispatch_dim<__half>(a, s);
    case FKV_F32: return dispatch_dim<float>(a, s);
  }
  return FKV_EDTYPE;
}

int launch_cache_append(const fkv_cache_append_params& p, cudaStream_t s) {
  const int es = dtype_size(p.dtype);
  AppendArgs a{};
  a.kn = (const char*)p.k_new.ptr; a.vn = (const char*)p.v_new.ptr;
  a.kc = (char*)const_cast<void*>(p.k_cache.ptr); a.vc = (char*)const_cast<void*>(p.v_cache.ptr);
  a.kn_sb = p.k_new.stride_b * es; a.kn_sh = p.k_new.stride_h * es;
  a.vn_sb = p.v_new.stride_b * es; a.vn_sh = p.v_new.stride_h * es;
  a.kc_sb = p.k_cache.stride_b * es; a.kc_sh = p.k_cache.stride_h * es; a.kc_sl = p.k_cache.stride_l * es;
  a.vc_sb = p.v_cache.stride_b * es; a.vc_sh = p.v_cache.stride_h * es; a.vc_sl = p.v_cache.stride_l * es;
  a.seq_len_dev = p.seq_len_dev; a.seq_len = p.seq_len; a.max_seq_len = p.max_seq_len;
  a.row_bytes = p.D * es;
  cache_append_kernel<<<dim3(p.Hkv, p.B), 32, 0, s>>>(a);
  FKV_LAUNCH_CHECK();
  if (p.bump && p.seq_len_dev) {
    bump_kernel<<<1, 1, 0, s>>>(p.seq_len_dev, p.max_seq_len);
    FKV_LAUNCH_CHECK();
  }
  return FKV_OK;
}

}  // namespace fkv
