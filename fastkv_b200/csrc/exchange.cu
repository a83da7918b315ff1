// The TSP layer's one exchange, written over peer memory instead of a library collective:
// KV-head-sharded ranks store their per-head score rows straight into each other's HBM (NVLink / NVSwitch P2P
// stores) so that every rank can sum ALL heads before its global top-k (baselines/fastkv/utils.py:127).
//
//   exchange_push_kernel   handshake ("my region may be overwritten") -> 16-byte stores of the local rows into
//                          every rank's region -> one release of a per-source DATA flag per peer
//   exchange_wait_kernel   acquires the peers' DATA flags; the consumer (fkv_topk, nsum = all heads) follows on
//                          the same stream
//   allreduce_kernel /     the small in-place sum of tensor-parallel decode steps over the same kind of region
//   allreduce_ll_kernel    (fkv_allreduce_sum; second half of this file)
// Everything is stream-ordered device work: capturable in a CUDA graph, no host round trip, no NCCL.
//
// Region header of the score exchange (u32 words; FKV_EXCHANGE_HEADER_BYTES = 256):
//   [0]        number of exchanges completed on the owner (bumped by its wait kernel)             (owner only)
//   [2]        error word: 1 = a handshake / data wait timed out                                  (owner only)
//   [8 + s]    READY word written by rank s: "s has entered push number e" (=> s consumed exchange e - 1)
//   [16 + s]   DATA counter, bumped once by every push CTA of rank s after its rows are in this region
//   [24 + s]   DATA count the owner has already consumed from rank s                               (owner only)
// All words only grow (compared as signed differences), so nothing is ever reset and a replayed graph needs no
// host-side state: the epoch and the expected counts live in the region.
//
// Measured between two B200s (tools/exchange_trace.py, profiles/r2_exchange.md): the first version (per-thread
// system fences, a last-CTA-out ticket, release stores for every flag) took 17.5 us for 512 KiB per rank, of which
// 5 us were CTAs queueing behind each other's remote stores, and 5.5 us three serialised system fences of the last
// CTA.  Now every CTA publishes its own rows with ONE fence and one remote atomic, and the receiver counts CTAs.
#include "common.cuh"

namespace fkv {
namespace {

constexpr int kXEpoch = 0, kXErr = 2, kXReady = 8, kXData = 16, kXSeen = 24;
constexpr int kXTrace = 32;           // u64[16] %globaltimer stamps of the last exchange (FKV_X_TRACING builds only)
#ifdef FKV_X_TRACING
#define FKV_XT(i) reinterpret_cast<unsigned long long*>(mine + kXTrace)[i] = now_ns()
#else
#define FKV_XT(i)
#endif
constexpr int kXThreads = 256;
static_assert(kXSeen + FKV_MAX_PEERS <= 32 && FKV_MAX_PEERS == 8, "exchange header layout");

struct XArgs {
  const float* score;
  int64_t sb, sh;                     // element strides of the local rows
  int B, hl, rank, world;
  int64_t P;
  unsigned int* region[FKV_MAX_PEERS];
  unsigned long long timeout_ns;
  unsigned int push_ctas;             // grid of the push kernel: what every receiver adds to its expected DATA count
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_add_relaxed_sys(unsigned int* p, unsigned int v) {
  asm volatile("red.relaxed.sys.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_sys() { asm volatile("fence.acq_rel.sys;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_peer16(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// Spin until *word has reached `e` (signed difference: the words only grow); false on timeout.  Relaxed polls: the
// caller fences once after the wait where it needs acquire semantics.
__device__ bool wait_word(const unsigned int* word, unsigned int e, unsigned long long timeout_ns) {
  if ((int)(ld_relaxed_sys(word) - e) >= 0) return true;
  const unsigned long long t0 = now_ns();
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
      if ((int)(ld_relaxed_sys(word) - e) >= 0) return true;
      __nanosleep(20);
    }
    if (now_ns() - t0 > timeout_ns) return false;
  }
}

__global__ void __launch_bounds__(kXThreads)
exchange_push_kernel(const __grid_constant__ XArgs a) {
  __shared__ unsigned int* region[FKV_MAX_PEERS];               // (indexing the kernel parameter dynamically would spill it to local memory)
#pragma unroll
  for (int r = 0; r < FKV_MAX_PEERS; ++r)
    if (threadIdx.x == r) region[r] = a.region[r];
  __syncthreads();
  unsigned int* mine = region[a.rank];
  const unsigned int e = ld_relaxed_gpu(mine + kXEpoch) + 1u;    // bumped by this rank's wait kernel, i.e. after every CTA here
  const int tid = threadIdx.x;
  if (blockIdx.x == 0 && tid == 0) FKV_XT(0);
  // (i) handshake.  Entering push number e proves that this rank's consumer of exchange e - 1 has finished (stream
  // order), so peers may overwrite this region; they learn it from the READY word.  One CTA tells, every CTA listens.
  // Nothing is published by READY (the reads it reports are complete by the kernel boundary): relaxed accesses.
  if (blockIdx.x == 0 && tid < a.world && tid != a.rank) st_relaxed_sys(region[tid] + kXReady + a.rank, e);
  if (tid < a.world && tid != a.rank) {
    if (!wait_word(mine + kXReady + tid, e, a.timeout_ns)) mine[kXErr] = 1u;
  }
  __syncthreads();
  if (blockIdx.x == 0 && tid == 0) FKV_XT(1);

  // (ii) the rows.  Each 16-byte chunk is loaded once (L2-resident: the scoring kernel has just written it) and stored
  // to every rank's region, peers first so that the NVLink stores are in flight while the local one is issued.
  const int H = a.hl * a.world;
  const int64_t row_elems = a.P;
  const bool vec = ((row_elems & 3) == 0) && ((a.sb & 3) == 0) && ((a.sh & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(a.score) & 15) == 0);
  const int rows = a.B * a.hl;
  if (vec) {
    const int64_t chunks_per_row = row_elems >> 2;
    const int64_t total = (int64_t)rows * chunks_per_row;
    for (int64_t c0 = (int64_t)blockIdx.x * kXThreads * 2 + tid; c0 < total; c0 += (int64_t)gridDim.x * kXThreads * 2) {
      float4 v[2];
      int64_t dst_off[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t c = c0 + (int64_t)u * kXThreads;
        dst_off[u] = -1;
        if (c < total) {
          const int r = (int)(c / chunks_per_row);
          const int64_t i = (c - (int64_t)r * chunks_per_row) << 2;
          const int b = r / a.hl, h = r - b * a.hl;
          v[u] = *reinterpret_cast<const float4*>(a.score + b * a.sb + h * a.sh + i);
          dst_off[u] = ((int64_t)(b * H + a.rank * a.hl + h)) * row_elems + i;
        }
      }
      for (int pp = 1; pp <= a.world; ++pp) {
        const int peer = (a.rank + pp) % a.world;              // own region last
        float* dst = reinterpret_cast<float*>(region[peer] + FKV_EXCHANGE_HEADER_BYTES / 4);
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (dst_off[u] >= 0) st_peer16(dst + dst_off[u], v[u]);
      }
    }
  } else {                                                       // odd row lengths / unaligned views: 4-byte stores
    const int64_t total = (int64_t)rows * row_elems;
    for (int64_t c = (int64_t)blockIdx.x * kXThreads + tid; c < total; c += (int64_t)gridDim.x * kXThreads) {
      const int r = (int)(c / row_elems);
      const int64_t i = c - (int64_t)r * row_elems;
      const int b = r / a.hl, h = r - b * a.hl;
      const float val = a.score[b * a.sb + h * a.sh + i];
      const int64_t off = ((int64_t)(b * H + a.rank * a.hl + h)) * row_elems + i;
      for (int pp = 1; pp <= a.world; ++pp) {
        const int peer = (a.rank + pp) % a.world;
        reinterpret_cast<float*>(region[peer] + FKV_EXCHANGE_HEADER_BYTES / 4)[off] = val;
      }
    }
  }
  if (blockIdx.x == 0 && tid == 0) FKV_XT(2);

  // (iii) publish, per CTA: the CTA barrier orders every thread's stores before the ONE system-scope fence of the
  // thread that then bumps the DATA counter in each peer's region (a remote atomic: the receiver counts CTAs).
  __syncthreads();
  if (tid < a.world && tid != a.rank) {
    fence_sys();
    red_add_relaxed_sys(region[tid] + kXData + a.rank, 1u);
  }
  if (blockIdx.x == 0 && tid == 0) FKV_XT(3);
}

__global__ void __launch_bounds__(32)
exchange_wait_kernel(const __grid_constant__ XArgs a) {
  unsigned int* mine = nullptr;
#pragma unroll
  for (int r = 0; r < FKV_MAX_PEERS; ++r)
    if (a.rank == r) mine = a.region[r];
  const int tid = threadIdx.x;
  if (tid == 0) FKV_XT(6);
  if (tid < a.world && tid != a.rank) {
    const unsigned int want = mine[kXSeen + tid] + a.push_ctas;  // every rank pushes the same shape with the same grid
    if (!wait_word(mine + kXData + tid, want, a.timeout_ns)) mine[kXErr] = 1u;
    mine[kXSeen + tid] = want;
  }
  fence_sys();                                                   // acquire: the rows behind the counters (the kernel boundary that follows orders the consumer)
  __syncwarp();
  if (tid == 0) {
    mine[kXEpoch] = mine[kXEpoch] + 1u;                          // the push in front of this kernel has completed: stream order
    FKV_XT(7);
  }
}

int push_grid(const fkv_exchange_params& p) {
  const int64_t chunks = ((int64_t)p.B * p.heads_local * p.P + 3) / 4;
  const int64_t want = (chunks + kXThreads - 1) / kXThreads;     // one 16-byte chunk per thread until the cap
  static const int env_cap = getenv("FKV_X_CTAS") ? atoi(getenv("FKV_X_CTAS")) : 0;   // tuning aid; must agree on all ranks
  int cap = p.max_ctas > 0 ? p.max_ctas : (env_cap > 0 ? env_cap : 128);   // remote stores drain slowly per SM: spread them over most of the chip
  if (cap > num_sms()) cap = num_sms();
  return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

// ---- small all-reduce (sum) over peer memory: see fkv_allreduce_sum in the public header -----------------------------
constexpr int kArEpoch = 0, kArTicket = 1, kArErr = 2;
constexpr int kArFlags = 64;          // u32 word index of flags[source][cta] (8 x 32 words)
static_assert(kArFlags * 4 + FKV_MAX_PEERS * FKV_ALLREDUCE_MAX_CTAS * 4 <= FKV_ALLREDUCE_HEADER_BYTES, "all-reduce header layout");

struct ArArgs {
  char* data;
  int64_t chunks;                     // 16-byte chunks
  int dtype, rank, world;
  size_t slot_bytes;
  unsigned int* region[FKV_MAX_PEERS];
  unsigned long long timeout_ns;
};

__device__ __forceinline__ uint4 ld_cg16(const void* p) {       // L2 only: the line may hold a two-exchanges-old value in L1
  uint4 r;
  asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}

template <int DT>
__device__ __forceinline__ void ar_accumulate(float* acc, const uint4& v) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  if constexpr (DT == FKV_F32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] += __uint_as_float(w[i]);
  } else if constexpr (DT == FKV_BF16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { acc[2 * i] += __uint_as_float(w[i] << 16); acc[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u); }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      acc[2 * i] += f.x; acc[2 * i + 1] += f.y;
    }
  }
}

template <int DT>
__device__ __forceinline__ uint4 ar_pack(const float* acc) {
  uint4 o;
  uint32_t* w = reinterpret_cast<uint32_t*>(&o);
  if constexpr (DT == FKV_F32) {
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = __float_as_uint(acc[i]);
  } else if constexpr (DT == FKV_BF16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { const __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * i], acc[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&h); }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) { const __half2 h = __floats2half2_rn(acc[2 * i], acc[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&h); }
  }
  return o;
}

template <int DT>
__global__ void __launch_bounds__(kXThreads)
allreduce_kernel(const __grid_constant__ ArArgs a) {
  __shared__ unsigned int* region[FKV_MAX_PEERS];
#pragma unroll
  for (int r = 0; r < FKV_MAX_PEERS; ++r)
    if (threadIdx.x == r) region[r] = a.region[r];
  __syncthreads();
  unsigned int* mine = region[a.rank];
  const int tid = threadIdx.x;
  const unsigned int e = ld_relaxed_gpu(mine + kArEpoch) + 1u;   // the last CTA out publishes e after every CTA has read this
  const size_t slot_off = FKV_ALLREDUCE_HEADER_BYTES + (size_t)(e & 1u) * a.world * a.slot_bytes;
  const int64_t per = (a.chunks + gridDim.x - 1) / gridDim.x;
  const int64_t c0 = (int64_t)blockIdx.x * per, c1 = min(a.chunks, c0 + per);
  // 1. this CTA's range of the local partial goes into slot [e & 1][rank] of every peer
  for (int64_t c = c0 + tid; c < c1; c += kXThreads) {
    const uint4 v = *reinterpret_cast<const uint4*>(a.data + c * 16);
    for (int pp = 1; pp < a.world; ++pp) {
      const int peer = (a.rank + pp) % a.world;
      char* dst = reinterpret_cast<char*>(region[peer]) + slot_off + (size_t)a.rank * a.slot_bytes + (a.slot_bytes >> 1) + c * 16;
      asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
    }
  }
  __syncthreads();
  // 2. publish (one fence + one flag per peer), then wait for the same range from every peer
  if (tid < a.world && tid != a.rank) {
    fence_sys();
    st_relaxed_sys(region[tid] + kArFlags + a.rank * FKV_ALLREDUCE_MAX_CTAS + blockIdx.x, e);
    if (!wait_word(mine + kArFlags + tid * FKV_ALLREDUCE_MAX_CTAS + blockIdx.x, e, a.timeout_ns)) mine[kArErr] = 1u;
    fence_sys();
  }
  __syncthreads();
  // 3. sum in rank order, fp32, in place
  for (int64_t c = c0 + tid; c < c1; c += kXThreads) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < a.world; ++r) {
      const uint4 v = (r == a.rank) ? *reinterpret_cast<const uint4*>(a.data + c * 16)
                                    : ld_cg16(reinterpret_cast<const char*>(mine) + slot_off + (size_t)r * a.slot_bytes + (a.slot_bytes >> 1) + c * 16);
      ar_accumulate<DT>(acc, v);
    }
    *reinterpret_cast<uint4*>(a.data + c * 16) = ar_pack<DT>(acc);
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned int prev = atomicAdd(mine + kArTicket, 1u);
    if (prev == gridDim.x - 1) {
      mine[kArTicket] = 0u;
      __threadfence();
      mine[kArEpoch] = e;
    }
  }
}

// Low-latency variant for the sizes decode steps actually sum (<= 64 KiB): no fence and no separate flag.  Every 16-byte
// store carries 8 bytes of payload and the exchange number twice ({d0, e, d1, e}: a 16-byte store lands as a whole), the
// receiver spins on the line itself until both tags read e — one NVLink one-way latency instead of a fence round trip
// plus a flag.  Tagged lines take 2 x the tensor's bytes and live in the FIRST half of a slot, the plain rows of the
// fence + flag kernel in the second half, so the two never overlay (a stale payload word can therefore never be taken for a
// tag); tags of older exchanges (e - 2 in the same slot) never match.
__device__ __forceinline__ void st_ll16(void* p, uint32_t d0, uint32_t d1, uint32_t e) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(d0), "r"(e), "r"(d1), "r"(e) : "memory");
}
__device__ __forceinline__ uint4 ld_ll16(const void* p) {
  uint4 r;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
  return r;
}

template <int DT>
__global__ void __launch_bounds__(kXThreads)
allreduce_ll_kernel(const __grid_constant__ ArArgs a) {
  __shared__ unsigned int* region[FKV_MAX_PEERS];
#pragma unroll
  for (int r = 0; r < FKV_MAX_PEERS; ++r)
    if (threadIdx.x == r) region[r] = a.region[r];
  __syncthreads();
  unsigned int* mine = region[a.rank];
  const int tid = threadIdx.x;
  const unsigned int e = ld_relaxed_gpu(mine + kArEpoch) + 1u;
  const size_t slot_off = FKV_ALLREDUCE_HEADER_BYTES + (size_t)(e & 1u) * a.world * a.slot_bytes;
  const int64_t c = (int64_t)blockIdx.x * kXThreads + tid;       // one 16-byte chunk per thread (the launcher sizes the grid)
  const bool active = c < a.chunks;
  uint4 own = make_uint4(0u, 0u, 0u, 0u);
  if (active) {
    own = *reinterpret_cast<const uint4*>(a.data + c * 16);
    for (int pp = 1; pp < a.world; ++pp) {
      const int peer = (a.rank + pp) % a.world;
      char* dst = reinterpret_cast<char*>(region[peer]) + slot_off + (size_t)a.rank * a.slot_bytes + c * 32;
      st_ll16(dst, own.x, own.y, e);
      st_ll16(dst + 16, own.z, own.w, e);
    }
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    bool timed_out = false;
    for (int r = 0; r < a.world; ++r) {                          // rank order
      uint4 v = own;
      if (r != a.rank) {
        const char* src = reinterpret_cast<const char*>(mine) + slot_off + (size_t)r * a.slot_bytes + c * 32;
        uint4 lo = ld_ll16(src), hi = ld_ll16(src + 16);
        if (lo.y != e || lo.w != e || hi.y != e || hi.w != e) {
          const unsigned long long t0 = now_ns();
          int spins = 0;
          do {
            lo = ld_ll16(src); hi = ld_ll16(src + 16);
            if ((++spins & 63) == 0 && now_ns() - t0 > a.timeout_ns) { timed_out = true; break; }
          } while (lo.y != e || lo.w != e || hi.y != e || hi.w != e);
        }
        v = make_uint4(lo.x, lo.z, hi.x, hi.z);
      }
      ar_accumulate<DT>(acc, v);
    }
    if (timed_out) mine[kArErr] = 1u;
    *reinterpret_cast<uint4*>(a.data + c * 16) = ar_pack<DT>(acc);
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned int prev = atomicAdd(mine + kArTicket, 1u);
    if (prev == gridDim.x - 1) {
      mine[kArTicket] = 0u;
      __threadfence();
      mine[kArEpoch] = e;
    }
  }
}

XArgs make_args(const fkv_exchange_params& p) {
  XArgs a{};
  a.score = p.score; a.sb = p.score_stride_b; a.sh = p.score_stride_h;
  a.B = p.B; a.hl = p.heads_local; a.rank = p.rank; a.world = p.world; a.P = p.P;
  for (int r = 0; r < p.world; ++r) a.region[r] = static_cast<unsigned int*>(p.region[r]);
  a.timeout_ns = (unsigned long long)(p.timeout_ms > 0 ? p.timeout_ms : 20000) * 1000000ull;
  a.push_ctas = (unsigned)push_grid(p);
  return a;
}

}  // namespace

int launch_exchange_push(const fkv_exchange_params& p, cudaStream_t s) {
  const XArgs a = make_args(p);
  exchange_push_kernel<<<a.push_ctas, kXThreads, 0, s>>>(a);
  FKV_LAUNCH_CHECK();
  return FKV_OK;
}

int launch_exchange_wait(const fkv_exchange_params& p, cudaStream_t s) {
  const XArgs a = make_args(p);
  exchange_wait_kernel<<<1, 32, 0, s>>>(a);
  FKV_LAUNCH_CHECK();
  return FKV_OK;
}

int launch_allreduce_sum(const fkv_allreduce_params& p, cudaStream_t s) {
  ArArgs a{};
  const int es = dtype_size(p.dtype);
  a.data = static_cast<char*>(p.data);
  a.chunks = p.n * es / 16;
  a.dtype = p.dtype; a.rank = p.rank; a.world = p.world;
  a.slot_bytes = ((p.region_bytes - FKV_ALLREDUCE_HEADER_BYTES) / (2 * (size_t)p.world)) & ~(size_t)255;
  for (int r = 0; r < p.world; ++r) a.region[r] = static_cast<unsigned int*>(p.region[r]);
  a.timeout_ns = (unsigned long long)(p.timeout_ms > 0 ? p.timeout_ms : 20000) * 1000000ull;
  int64_t want = (a.chunks + kXThreads - 1) / kXThreads;       // one chunk per thread up to 32 CTAs (same on every rank)
  static const bool no_ll = getenv("FKV_AR_NO_LL") != nullptr;  // experiments: always take the fence + flag kernel
  const size_t bytes = (size_t)a.chunks * 16;
  if (!no_ll && bytes <= (64u << 10) && 2 * bytes <= (a.slot_bytes >> 1)) {   // decode-step sizes: tagged 16-byte lines, no fence
    const int g = (int)want;                                    // <= 16 CTAs for 64 KiB
    if (p.dtype == FKV_F32) allreduce_ll_kernel<FKV_F32><<<g, kXThreads, 0, s>>>(a);
    else if (p.dtype == FKV_BF16) allreduce_ll_kernel<FKV_BF16><<<g, kXThreads, 0, s>>>(a);
    else allreduce_ll_kernel<FKV_F16><<<g, kXThreads, 0, s>>>(a);
    FKV_LAUNCH_CHECK();
    return FKV_OK;
  }
  const int grid = (int)(want < 1 ? 1 : (want > FKV_ALLREDUCE_MAX_CTAS ? FKV_ALLREDUCE_MAX_CTAS : want));
  if (p.dtype == FKV_F32) allreduce_kernel<FKV_F32><<<grid, kXThreads, 0, s>>>(a);
  else if (p.dtype == FKV_BF16) allreduce_kernel<FKV_BF16><<<grid, kXThreads, 0, s>>>(a);
  else allreduce_kernel<FKV_F16><<<grid, kXThreads, 0, s>>>(a);
  FKV_LAUNCH_CHECK();
  return FKV_OK;
}

int exchange_status(const void* region, cudaStream_t s) {
  unsigned int words[4] = {0, 0, 0, 0};
  FKV_CUDA_TRY(cudaMemcpyAsync(words, region, sizeof(words), cudaMemcpyDeviceToHost, s));
  FKV_CUDA_TRY(cudaStreamSynchronize(s));
  return words[kXErr] ? FKV_ETIMEOUT : FKV_OK;
}

}  // namespace fkv
