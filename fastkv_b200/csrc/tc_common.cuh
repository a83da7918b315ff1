// Shared pieces of the tcgen05 kernels (score_tc.cu, gemfilter_tc.cu, small_compress.cu): tile constants, raw PTX
// wrappers (mbarrier, TMA, tcgen05.mma / commit / ld, fences), the epoch-tagged record helpers, the balanced unit split,
// and the host-side tensor-map builder.  Everything lives in an anonymous namespace: each translation unit gets its own copy.
#pragma once
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
constexpr int kCtrlThreads = 128;          // warpgroup 0: warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-3 idle
constexpr int kTcThreads = kCtrlThreads + kEpiThreads;   // four ALIGNED warpgroups, so that setmaxnreg can move registers
                                                         // from the two single-thread control warps to the epilogue
constexpr int kCtrlRegs = 56, kEpiRegs = 152;            // 128 * 56 + 384 * 152 = 65536 = 512 * 128
constexpr int kMaxLocalPairs = 8;
constexpr int kMaxStages = 6;
constexpr int kMaxPhases = 512;            // (b, h) pairs are processed in groups whose logits fit in TMEM
constexpr float kLazyTau = 6.0f;          // log2 units of headroom before a running max is raised
constexpr float kEmptyMax = -1.0e30f;     // "no key seen yet" (finite, so exp2(x - m) never sees inf - inf)

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
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
        "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])) : "memory");
}
template <int N>
__device__ __forceinline__ void tmem_store_row(uint32_t taddr, const float* v) {
  if constexpr (N == 8) tmem_st8(taddr, v);
  else {
#pragma unroll
    for (int c = 0; c < N; c += 16) tmem_st16(taddr + c, v + c);
  }
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
// Encoded maps are cached per translation unit, keyed by everything that enters the encoding (pointer, dtype, shape,
// strides, box): a prefill calls the same layer buffers again and again, and cuTensorMapEncodeTiled costs ~1 us of host
// time per map.  Mutex-guarded (the C ABI is re-entrant), 64 entries, round-robin replacement.
struct MapKey {
  const void* ptr; int64_t sb, sh, sl; int dtype, B, H, L, D, box_rows;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && sb == o.sb && sh == o.sh && sl == o.sl && dtype == o.dtype && B == o.B && H == o.H && L == o.L &&
           D == o.D && box_rows == o.box_rows;
  }
};
struct MapEntry { MapKey key; CUtensorMap map; int pos_l, pos_h, pos_b; bool valid; };

int make_map_uncached(CUtensorMap* map, const fkv_view4& v, int dtype, int B, int H, int L, int D, int box_rows, int* pos_l,
                      int* pos_h, int* pos_b);

int make_map(CUtensorMap* map, const fkv_view4& v, int dtype, int B, int H, int L, int D, int box_rows, int* pos_l,
             int* pos_h, int* pos_b) {
  constexpr int kEntries = 64;
  static std::mutex mu;
  static MapEntry cache[kEntries];
  static int next = 0;
  const MapKey key{v.ptr, v.stride_b, v.stride_h, v.stride_l, dtype, B, H, L, D, box_rows};
  {
    std::lock_guard<std::mutex> lock(mu);
    for (int i = 0; i < kEntries; ++i)
      if (cache[i].valid && cache[i].key == key) {
        *map = cache[i].map; *pos_l = cache[i].pos_l; *pos_h = cache[i].pos_h; *pos_b = cache[i].pos_b;
        return FKV_OK;
      }
  }
  const int rc = make_map_uncached(map, v, dtype, B, H, L, D, box_rows, pos_l, pos_h, pos_b);
  if (rc != FKV_OK) return rc;
  std::lock_guard<std::mutex> lock(mu);
  MapEntry& e = cache[next];
  next = (next + 1) % kEntries;
  e.key = key; e.map = *map; e.pos_l = *pos_l; e.pos_h = *pos_h; e.pos_b = *pos_b; e.valid = true;
  return FKV_OK;
}

int make_map_uncached(CUtensorMap* map, const fkv_view4& v, int dtype, int B, int H, int L, int D, int box_rows, int* pos_l,
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

}  // namespace
}  // namespace fkv
