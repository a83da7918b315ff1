// Shared device/host helpers for the fastkv_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdlib>
#include <type_traits>
#include <stdint.h>

#include "../../include/fastkv_b200.h"

#define FKV_NUM_SMS_FALLBACK 148

#define FKV_CUDA_TRY(expr)                      \
  do {                                          \
    cudaError_t _e = (expr);                    \
    if (_e != cudaSuccess) return (int)_e;      \
  } while (0)

// Launch check that never synchronises: picks up configuration errors only.
#define FKV_LAUNCH_CHECK()                      \
  do {                                          \
    cudaError_t _e = cudaGetLastError();        \
    if (_e != cudaSuccess) return (int)_e;      \
    ::fkv::count_launch();                      \
  } while (0)

namespace fkv {

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

inline int dtype_size(int dtype) { return dtype == FKV_F32 ? 4 : 2; }

// layout of the persistent region at the front of every workspace (FKV_WORKSPACE_HEADER_BYTES in the public header)
constexpr int kHistMaxRows = FKV_SELECT_MAX_ROWS;                  // score rows (B * selection heads) per slot
constexpr int kHistSlots = 2;
constexpr int kHistTotalRows = kHistSlots * kHistMaxRows;          // a slot's rows are [(slot-1) * kHistMaxRows, slot * kHistMaxRows)
constexpr size_t kWsCtlOffset = 256;                               // u32[rows][8]: per-row barrier counters / list cursor
constexpr size_t kWsHistOffset = kWsCtlOffset + kHistTotalRows * 32;                      // u32[rows][4096]: 12-bit histogram of a score row
constexpr size_t kWsSubHistOffset = kWsHistOffset + (size_t)kHistTotalRows * 4096 * 4;    // same, second level (threshold bin)
constexpr size_t kWsListOffset = kWsSubHistOffset + (size_t)kHistTotalRows * 4096 * 4;    // u64[rows][4096]: candidate lists
constexpr size_t kWsTicketOffset = kWsListOffset + (size_t)kHistTotalRows * 4096 * 8;     // u32[1024]: decode-attention arrival tickets
constexpr size_t kWsTickets = 1024;
static_assert(kWsTicketOffset + kWsTickets * 4 == FKV_WORKSPACE_HEADER_BYTES, "workspace header layout");

constexpr uint32_t kWorkspaceMagic = 0xF457CB20u;   // written at byte 16 of a workspace by fkv_workspace_init

void count_launch();   // bumps the process-wide kernel-launch counter (abi.cu)

// ---- programmatic dependent launch (PDL) ------------------------------------------------------
// The kernels of the update_kv chain are launched with programmatic stream serialization and call
// pdl_enter() before touching global memory, so their prologue (barrier init, TMEM allocation, smem
// zeroing) overlaps the drain of the previous kernel.  Measured on B200: -0.5 % on the 82-launch
// bench step with the implicit trigger (dependents released when the primary's CTAs exit); an
// explicit early griddepcontrol.launch_dependents made the step 4 % SLOWER (early-resident
// dependents compete with the primary's CTAs for SM resources), so none is issued.
// FKV_NO_PDL=1 turns the attribute off.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
inline bool pdl_enabled() {
  static const bool on = getenv("FKV_NO_PDL") == nullptr;
  return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif
int num_sms();   // cudaDevAttrMultiProcessorCount of the CURRENT device, cached per device (abi.cu)
// Opt a kernel in to `bytes` of dynamic shared memory once per (device, kernel): function attributes are per
// device, and several host threads may launch concurrently (abi.cu; mutex-guarded, a no-op after the first call).
int ensure_dyn_smem(const void* kernel, int bytes);
template <typename K>
inline int ensure_dyn_smem(K* kernel, int bytes) { return ensure_dyn_smem(reinterpret_cast<const void*>(kernel), bytes); }
// true iff `ws` lies inside a buffer that went through fkv_workspace_init in this process (abi.cu)
bool workspace_registered(const void* ws, size_t bytes);

// ---- element loads -> fp32 -----------------------------------------------------------------
template <typename T> struct Elem;
template <> struct Elem<__nv_bfloat16> {
  static constexpr int kPer16B = 8;
  __device__ static inline void load8(const __nv_bfloat16* p, float* out) {   // 16-byte aligned
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {                    // bf16 -> fp32 is a 16-bit shift
      out[2 * i] = __uint_as_float(w[i] << 16);
      out[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <> struct Elem<__half> {
  static constexpr int kPer16B = 8;
  __device__ static inline void load8(const __half* p, float* out) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __half2* h = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __half22float2(h[i]);
      out[2 * i] = f.x;
      out[2 * i + 1] = f.y;
    }
  }
};
template <> struct Elem<float> {
  static constexpr int kPer16B = 4;
  __device__ static inline void load8(const float* p, float* out) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
    out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
  }
};

// ---- warp / block reductions ---------------------------------------------------------------
__device__ inline float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ inline float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// 128-bit streaming global accesses (read-once data: keep it out of L1).
__device__ inline uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ inline void st_stream16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

}  // namespace fkv

// ---- per-file launchers (each .cu implements its own; abi.cu validates and dispatches) -----
namespace fkv {
int launch_score_simple(const fkv_score_params& p, cudaStream_t s);
// p.hist_slot 1 / 2: also add the 12-bit histogram of every stored score to that slot's tables in the workspace
// (B * (h1 - h0) <= kHistMaxRows rows, consumed and re-zeroed by launch_select_compact)
int launch_score_tensor(const fkv_score_params& p, cudaStream_t s);
bool score_tensor_supported(const fkv_score_params& p);
bool gemfilter_tensor_supported(const fkv_score_params& p);     // gemfilter_tc.cu
int launch_score_gemfilter(const fkv_score_params& p, cudaStream_t s);
bool small_compress_supported(const fkv_compress_params& p);   // short contexts: one cluster kernel for the whole chain
int launch_small_compress(const fkv_compress_params& p, cudaStream_t s);
size_t score_workspace_bytes(int mode, int B, int Hq, int Hkv, int L, int W);
int launch_topk(const fkv_topk_params& p, cudaStream_t s);
// per-head selection with the KV compaction fused behind it when the cluster kernel is used (*fused)
int launch_topk_fused(const fkv_topk_params& p, const fkv_gather_kv_params& g, bool* fused, cudaStream_t s);
size_t topk_workspace_bytes(int rows, int n, int k);
// select_compact.cu: per-head top-k + KV compaction in one non-cluster launch (consumes the producer-built histogram)
bool select_compact_supported(const fkv_topk_params& p, const fkv_gather_kv_params& g);
int launch_select_compact(const fkv_topk_params& p, const fkv_gather_kv_params& g, void* workspace, int slot,
                          bool have_hist, cudaStream_t s, int max_ctas = 0);
int launch_gather_kv(const fkv_gather_kv_params& p, cudaStream_t s);
int launch_gather_rows(const fkv_gather_rows_params& p, cudaStream_t s);
int launch_decode_attn(const fkv_decode_attn_params& p, cudaStream_t s);
int launch_cache_append(const fkv_cache_append_params& p, cudaStream_t s);
size_t decode_workspace_bytes(int B, int Hq, int D, int max_seq_len);
// exchange.cu: the TSP layer's score exchange over peer memory
int launch_exchange_push(const fkv_exchange_params& p, cudaStream_t s);
int launch_exchange_wait(const fkv_exchange_params& p, cudaStream_t s);
int exchange_status(const void* region, cudaStream_t s);
int launch_allreduce_sum(const fkv_allreduce_params& p, cudaStream_t s);
}  // namespace fkv
