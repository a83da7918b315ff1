// extern "C" surface of libfastkv_b200.so: argument validation, workspace carving and dispatch.
// No torch headers, no allocation, no synchronisation (see include/fastkv_b200.h).
#include <atomic>
#include <cstring>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace fkv {

static std::atomic<unsigned long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

constexpr int kMaxDevices = 64;

static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
  return dev;
}

int num_sms() {
  static std::atomic<int> cached[kMaxDevices];          // zero-initialised; one slot per device ordinal
  const int dev = current_device();
  if (dev >= 0 && dev < kMaxDevices) {
    const int v = cached[dev].load(std::memory_order_relaxed);
    if (v > 0) return v;
  }
  int sms = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    (void)cudaGetLastError();
    return FKV_NUM_SMS_FALLBACK;
  }
  if (dev >= 0 && dev < kMaxDevices) cached[dev].store(sms, std::memory_order_relaxed);
  return sms;
}

int ensure_dyn_smem(const void* kernel, int bytes) {
  static std::mutex mu;
  static std::vector<std::pair<int, const void*>> done;   // (device, kernel) pairs already opted in
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& e : done)
    if (e.first == dev && e.second == kernel) return FKV_OK;
  FKV_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.emplace_back(dev, kernel);
  return FKV_OK;
}

// workspaces that went through fkv_workspace_init in this process: (device, base, bytes)
struct WsRange { int dev; uintptr_t base; size_t bytes; };
static std::mutex g_ws_mu;
static std::vector<WsRange> g_ws;

static void register_workspace(void* ws, size_t bytes) {
  const int dev = current_device();
  const uintptr_t b = reinterpret_cast<uintptr_t>(ws);
  std::lock_guard<std::mutex> lock(g_ws_mu);
  // a new buffer replaces every older registration it overlaps (the allocator handed the range out again)
  for (size_t i = 0; i < g_ws.size();) {
    const WsRange& r = g_ws[i];
    if (r.dev == dev && b < r.base + r.bytes && r.base < b + bytes) { g_ws[i] = g_ws.back(); g_ws.pop_back(); }
    else ++i;
  }
  g_ws.push_back({dev, b, bytes});
}

bool workspace_registered(const void* ws, size_t bytes) {
  const int dev = current_device();
  const uintptr_t b = reinterpret_cast<uintptr_t>(ws);
  std::lock_guard<std::mutex> lock(g_ws_mu);
  for (const WsRange& r : g_ws)
    if (r.dev == dev && b >= r.base && b + bytes <= r.base + r.bytes) return true;
  return false;
}

namespace {

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int check_view(const fkv_view4& v, int es) {
  if (v.ptr == nullptr) return FKV_EINVAL;
  if (!aligned16(v.ptr)) return FKV_EALIGN;
  if ((v.stride_b * es) % 16 || (v.stride_h * es) % 16 || (v.stride_l * es) % 16) return FKV_EALIGN;
  return FKV_OK;
}

int selection_heads(const fkv_score_params& p) {
  return p.mode == FKV_MODE_FASTKV ? p.Hkv : (p.mode == FKV_MODE_SNAPKV ? p.Hq : 1);
}

int validate_score(const fkv_score_params& p) {
  if (p.struct_size != sizeof(fkv_score_params)) return FKV_EINVAL;
  if (p.mode < FKV_MODE_FASTKV || p.mode > FKV_MODE_GEMFILTER) return FKV_EMODE;
  if (p.impl < FKV_IMPL_AUTO || p.impl > FKV_IMPL_TENSOR) return FKV_EMODE;
  if (p.dtype < FKV_BF16 || p.dtype > FKV_F32) return FKV_EDTYPE;
  if (p.D != 64 && p.D != 128) return FKV_EHEADDIM;
  if (p.B < 1 || p.Hq < 1 || p.Hkv < 1 || p.L < 1 || p.Hq % p.Hkv) return FKV_EINVAL;
  const int G = p.Hq / p.Hkv;
  if (G > 64) return FKV_EINVAL;                       // any group size: the tensor path splits it into sub-groups of 1/2/4 heads
  if (p.kernel_size < 1 || p.kernel_size > 63 || (p.kernel_size & 1) == 0) return FKV_EPOOL;
  if (p.mode != FKV_MODE_GEMFILTER) {
    if (p.pooling != FKV_POOL_AVG && p.pooling != FKV_POOL_MAX) return FKV_EPOOL;
    if (p.window_size < 1 || p.window_size > 64 || p.window_size >= p.L) return FKV_EWINDOW;
    if (p.h0 < 0 || p.h1 <= p.h0 || p.h1 > selection_heads(p)) return FKV_ERANGE;
  } else {
    if (p.qh0 < 0 || p.qh1 > p.Hq || (p.qh1 != 0 && p.qh1 <= p.qh0) || (p.qh1 == 0 && p.qh0 != 0)) return FKV_ERANGE;
  }
  const int es = dtype_size(p.dtype);
  int rc;
  if ((rc = check_view(p.q, es)) != FKV_OK) return rc;
  if ((rc = check_view(p.k, es)) != FKV_OK) return rc;
  if (p.score == nullptr) return FKV_EINVAL;
  if (p.workspace == nullptr ||
      p.workspace_bytes < score_workspace_bytes(p.mode, p.B, p.Hq, p.Hkv, p.L, p.window_size))
    return FKV_EWORKSPACE;
  // the header (launch epoch, departure counter) is persistent state: refuse a buffer that never went through
  // fkv_workspace_init instead of computing on a garbage epoch (no device-side trap, nothing is launched)
  if (!workspace_registered(p.workspace, FKV_WORKSPACE_HEADER_BYTES)) return FKV_EWORKSPACE;
  if (p.hist_slot < 0 || p.hist_slot > kHistSlots || p.max_ctas < 0) return FKV_EINVAL;
  return FKV_OK;
}

// *hist_used: whether the scoring kernel filled the optional histogram table (tensor path only)
int run_score(const fkv_score_params& p, cudaStream_t s, bool* hist_used = nullptr) {
  int impl = p.impl;
  if (hist_used) *hist_used = false;
  if (impl == FKV_IMPL_AUTO) impl = score_tensor_supported(p) ? FKV_IMPL_TENSOR : FKV_IMPL_SIMPLE;
  if (impl == FKV_IMPL_TENSOR) {
    if (!score_tensor_supported(p)) return FKV_EMODE;
    if (hist_used)
      *hist_used = p.hist_slot >= 1 && p.hist_slot <= kHistSlots && p.mode != FKV_MODE_GEMFILTER &&
                   p.B * (p.h1 - p.h0) <= kHistMaxRows;
    return launch_score_tensor(p, s);
  }
  return launch_score_simple(p, s);
}

}  // namespace
}  // namespace fkv

using namespace fkv;

extern "C" {

int fkv_abi_version(void) { return FKV_ABI_VERSION; }

uint64_t fkv_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* fkv_strerror(int code) {
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  switch (code) {
    case FKV_OK: return "ok";
    case FKV_EINVAL: return "invalid argument (null pointer, bad size, struct_size mismatch or head counts)";
    case FKV_EDTYPE: return "unsupported dtype (bf16, f16, f32 only)";
    case FKV_EHEADDIM: return "unsupported head_dim (64 or 128 only)";
    case FKV_EPOOL: return "Pooling method not supported (avgpool/maxpool with an odd kernel_size < 64)";
    case FKV_ETOPK: return "top-k size out of range (need 1 <= k <= n)";
    case FKV_EALIGN: return "pointer or stride not 16-byte aligned";
    case FKV_EWORKSPACE: return "workspace missing or too small (see fkv_workspace_bytes)";
    case FKV_EMODE: return "unknown or unsupported mode / impl";
    case FKV_EWINDOW: return "window_size out of range (1..64 and < L)";
    case FKV_ERANGE: return "head range outside the selection heads";
    case FKV_ETIMEOUT: return "a peer never arrived at the score exchange";
  }
  return "unknown fastkv_b200 error";
}

size_t fkv_workspace_bytes(int32_t mode, int32_t B, int32_t Hq, int32_t Hkv, int32_t L, int32_t window_size,
                           int32_t capacity, int32_t tsp_k) {
  const int Hsel = mode == FKV_MODE_FASTKV ? Hkv : (mode == FKV_MODE_SNAPKV ? Hq : 1);
  size_t b = score_workspace_bytes(mode, B, Hq, Hkv, L, window_size);
  const int k_head = capacity > window_size ? capacity - window_size : 1;
  b += topk_workspace_bytes(B * Hsel, L, k_head);
  if (tsp_k > 0) b += topk_workspace_bytes(B, L, tsp_k);
  return b + 4096;                                     // score_workspace_bytes already includes the header
}

int fkv_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
  if (!workspace || workspace_bytes < FKV_WORKSPACE_HEADER_BYTES) return FKV_EWORKSPACE;
  FKV_CUDA_TRY(cudaMemsetAsync(workspace, 0, workspace_bytes, (cudaStream_t)stream));
  register_workspace(workspace, workspace_bytes);
  // header word 4 = "initialised" mark; the scoring kernel traps on a workspace that never went through here
  static const uint32_t magic = kWorkspaceMagic;
  FKV_CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(workspace) + 16, &magic, sizeof(magic), cudaMemcpyHostToDevice, (cudaStream_t)stream));
  return FKV_OK;
}

size_t fkv_topk_workspace_bytes(int32_t rows, int32_t n, int32_t k) {
  return FKV_WORKSPACE_HEADER_BYTES + topk_workspace_bytes(rows, n, k);
}

size_t fkv_decode_workspace_bytes(int32_t B, int32_t Hq, int32_t D, int32_t max_seq_len) {
  return FKV_WORKSPACE_HEADER_BYTES + decode_workspace_bytes(B, Hq, D, max_seq_len);
}

int fkv_workspace_status(const void* workspace, void* stream) {
  if (!workspace) return FKV_EWORKSPACE;
  uint32_t hdr[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  FKV_CUDA_TRY(cudaMemcpyAsync(hdr, workspace, sizeof(hdr), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  FKV_CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
  return (hdr[4] == kWorkspaceMagic && hdr[5] == 0u) ? FKV_OK : FKV_EWORKSPACE;
}

const char* fkv_score_impl_name(const fkv_score_params* p) {
  if (!p) return "invalid";
  if (p->impl == FKV_IMPL_SIMPLE) return "simple";
  return score_tensor_supported(*p) ? "tensor" : "simple";
}

int fkv_score(const fkv_score_params* p, void* stream) {
  if (!p) return FKV_EINVAL;
  const int rc = validate_score(*p);
  if (rc != FKV_OK) return rc;
  return run_score(*p, (cudaStream_t)stream);
}

int fkv_topk(const fkv_topk_params* p, void* stream) {
  if (!p || p->struct_size != sizeof(fkv_topk_params)) return FKV_EINVAL;
  if (p->rows < 1 || p->n < 1 || !p->src || !p->out || p->tail < 0) return FKV_EINVAL;
  if (p->k < 1 || p->k > p->n) return FKV_ETOPK;
  if (p->order != FKV_ORDER_VALUE_DESC && p->order != FKV_ORDER_INDEX_ASC) return FKV_EMODE;
  fkv_topk_params q = *p;                              // scratch starts past the scoring header
  if (q.workspace) {
    if (q.workspace_bytes < FKV_WORKSPACE_HEADER_BYTES) return FKV_EWORKSPACE;
    q.workspace = static_cast<char*>(q.workspace) + FKV_WORKSPACE_HEADER_BYTES;
    q.workspace_bytes -= FKV_WORKSPACE_HEADER_BYTES;
  }
  return launch_topk(q, (cudaStream_t)stream);
}

int fkv_gather_kv(const fkv_gather_kv_params* p, void* stream) {
  if (!p || p->struct_size != sizeof(fkv_gather_kv_params)) return FKV_EINVAL;
  if (p->dtype < FKV_BF16 || p->dtype > FKV_F32) return FKV_EDTYPE;
  if (p->D != 64 && p->D != 128) return FKV_EHEADDIM;
  if (p->B < 1 || p->L < 1 || p->n_sel < 0 || p->window_size < 0 || p->h1 <= p->h0 || p->head_div < 1 || !p->idx)
    return FKV_EINVAL;
  if (p->window_size > p->L) return FKV_EWINDOW;
  const int es = dtype_size(p->dtype);
  int rc;
  if ((rc = check_view(p->k, es)) || (rc = check_view(p->v, es)) || (rc = check_view(p->k_out, es)) ||
      (rc = check_view(p->v_out, es)))
    return rc;
  return launch_gather_kv(*p, (cudaStream_t)stream);
}

int fkv_select_compact(const fkv_select_compact_params* p, void* stream) {
  if (!p || p->struct_size != sizeof(fkv_select_compact_params)) return FKV_EINVAL;
  const fkv_gather_kv_params& g = p->gather;
  if (g.struct_size != sizeof(fkv_gather_kv_params)) return FKV_EINVAL;
  if (g.dtype < FKV_BF16 || g.dtype > FKV_F32) return FKV_EDTYPE;
  if (g.D != 64 && g.D != 128) return FKV_EHEADDIM;
  if (g.B < 1 || g.L < 1 || g.n_sel < 1 || g.window_size < 0 || g.h1 <= g.h0 || g.head_div < 1 || !p->score || !p->idx_out)
    return FKV_EINVAL;
  if (g.window_size >= g.L) return FKV_EWINDOW;
  const int P = g.L - g.window_size, Hsel = g.h1 - g.h0;
  if (g.n_sel > P) return FKV_ETOPK;
  if (p->hist_slot < 1 || p->hist_slot > kHistSlots) return FKV_EINVAL;
  const int es = dtype_size(g.dtype);
  int rc;
  if ((rc = check_view(g.k, es)) || (rc = check_view(g.v, es)) || (rc = check_view(g.k_out, es)) || (rc = check_view(g.v_out, es)))
    return rc;
  if (!p->workspace || p->workspace_bytes < FKV_WORKSPACE_HEADER_BYTES ||
      !workspace_registered(p->workspace, FKV_WORKSPACE_HEADER_BYTES))
    return FKV_EWORKSPACE;
  fkv_topk_params tk{};
  tk.struct_size = sizeof(tk);
  tk.rows = g.B * Hsel; tk.n = P; tk.k = g.n_sel; tk.order = FKV_ORDER_VALUE_DESC; tk.tail = 0; tk.nsum = 1;
  tk.src = p->score; tk.src_stride_row = P; tk.src_stride_sum = 0;
  tk.out = p->idx_out; tk.out_stride_row = g.n_sel;
  if (!select_compact_supported(tk, g)) return FKV_EMODE;
  return launch_select_compact(tk, g, p->workspace, p->hist_slot, p->have_hist != 0 && P > 4096, (cudaStream_t)stream,
                               p->max_ctas);
}

int fkv_gather_rows(const fkv_gather_rows_params* p, void* stream) {
  if (!p || p->struct_size != sizeof(fkv_gather_rows_params)) return FKV_EINVAL;
  if (p->B < 1 || p->L < 1 || p->T < 1 || !p->src || !p->out || !p->idx) return FKV_EINVAL;
  if (p->row_bytes < 16 || p->row_bytes % 16 || !aligned16(p->src) || !aligned16(p->out) ||
      p->src_stride_b_bytes % 16 || p->src_stride_l_bytes % 16 || p->out_stride_b_bytes % 16 || p->out_stride_l_bytes % 16)
    return FKV_EALIGN;
  if (p->pos_in && !p->pos_out) return FKV_EINVAL;
  return launch_gather_rows(*p, (cudaStream_t)stream);
}

static int check_exchange(const fkv_exchange_params* p) {
  if (!p || p->struct_size != sizeof(fkv_exchange_params)) return FKV_EINVAL;
  if (p->world < 1 || p->world > FKV_MAX_PEERS || p->rank < 0 || p->rank >= p->world) return FKV_EINVAL;
  if (p->B < 1 || p->heads_local < 1 || p->P < 1 || !p->score) return FKV_EINVAL;
  for (int r = 0; r < p->world; ++r) {
    if (!p->region[r]) return FKV_EINVAL;
    if (!aligned16(p->region[r])) return FKV_EALIGN;
  }
  const size_t need = (size_t)FKV_EXCHANGE_HEADER_BYTES + (size_t)p->B * p->world * p->heads_local * (size_t)p->P * 4;
  if (p->region_bytes < need) return FKV_EWORKSPACE;
  return FKV_OK;
}

int fkv_exchange_push(const fkv_exchange_params* p, void* stream) {
  if (int rc = check_exchange(p)) return rc;
  return launch_exchange_push(*p, (cudaStream_t)stream);
}

int fkv_exchange_wait(const fkv_exchange_params* p, void* stream) {
  if (int rc = check_exchange(p)) return rc;
  return launch_exchange_wait(*p, (cudaStream_t)stream);
}

int fkv_allreduce_sum(const fkv_allreduce_params* p, void* stream) {
  if (!p || p->struct_size != sizeof(fkv_allreduce_params)) return FKV_EINVAL;
  if (p->world < 1 || p->world > FKV_MAX_PEERS || p->rank < 0 || p->rank >= p->world || p->n < 1 || !p->data) return FKV_EINVAL;
  if (p->dtype < FKV_BF16 || p->dtype > FKV_F32) return FKV_EDTYPE;
  const size_t bytes = (size_t)p->n * dtype_size(p->dtype);
  if (bytes % 16 || !aligned16(p->data)) return FKV_EALIGN;
  for (int r = 0; r < p->world; ++r) {
    if (!p->region[r]) return FKV_EINVAL;
    if (!aligned16(p->region[r])) return FKV_EALIGN;
  }
  if (p->region_bytes < FKV_ALLREDUCE_HEADER_BYTES + 2 * (size_t)p->world * 256) return FKV_EWORKSPACE;
  const size_t slot = ((p->region_bytes - FKV_ALLREDUCE_HEADER_BYTES) / (2 * (size_t)p->world)) & ~(size_t)255;
  if (bytes > slot / 2) return FKV_EWORKSPACE;         // (half a slot: the other half holds the tagged lines of small sums)
  if (p->world == 1) return FKV_OK;                    // nothing to add
  return launch_allreduce_sum(*p, (cudaStream_t)stream);
}

int fkv_exchange_status(const void* local_region, void* stream) {
  if (!local_region) return FKV_EINVAL;
  return exchange_status(local_region, (cudaStream_t)stream);
}

int fkv_peer_alloc(size_t bytes, void** ptr_out, unsigned char handle_out[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t");
  if (!bytes || !ptr_out || !handle_out) return FKV_EINVAL;
  void* ptr = nullptr;
  FKV_CUDA_TRY(cudaMalloc(&ptr, bytes));
  cudaError_t e = cudaMemset(ptr, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) { cudaFree(ptr); return (int)e; }
  memcpy(handle_out, &h, 64);
  *ptr_out = ptr;
  return FKV_OK;
}

int fkv_peer_open(const unsigned char handle[64], void** ptr_out) {
  if (!handle || !ptr_out) return FKV_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, 64);
  FKV_CUDA_TRY(cudaIpcOpenMemHandle(ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
  return FKV_OK;
}

int fkv_peer_close(void* mapped_ptr) {
  if (!mapped_ptr) return FKV_EINVAL;
  FKV_CUDA_TRY(cudaIpcCloseMemHandle(mapped_ptr));
  return FKV_OK;
}

int fkv_peer_free(void* ptr) {
  if (!ptr) return FKV_EINVAL;
  FKV_CUDA_TRY(cudaFree(ptr));
  return FKV_OK;
}

int fkv_decode_attn(const fkv_decode_attn_params* p, void* stream) {
  if (!p || p->struct_size != sizeof(fkv_decode_attn_params)) return FKV_EINVAL;
  if (p->dtype < FKV_BF16 || p->dtype > FKV_F32) return FKV_EDTYPE;
  if (p->D != 64 && p->D != 128) return FKV_EHEADDIM;
  if (p->B < 1 || p->Hq < 1 || p->Hkv < 1 || p->Hq % p->Hkv || !p->q || !p->out) return FKV_EINVAL;
  if (!p->seq_len_dev && p->seq_len < 1) return FKV_EINVAL;
  const int es = dtype_size(p->dtype);
  int rc;
  if ((rc = check_view(p->k_cache, es)) || (rc = check_view(p->v_cache, es))) return rc;
  if ((p->q_stride_b * es) % 16 || (p->q_stride_h * es) % 16 || (p->out_stride_b * es) % 16 || (p->out_stride_h * es) % 16)
    return FKV_EALIGN;
  if (!aligned16(p->q) || !aligned16(p->out)) return FKV_EALIGN;
  if (p->seq_len_dev && p->max_seq_len < 1) return FKV_EINVAL;      // the device-side length needs its bound
  // the kernel keeps its arrival tickets in the persistent region: the workspace must be an initialised one
  if (!p->workspace || p->workspace_bytes < FKV_WORKSPACE_HEADER_BYTES ||
      !workspace_registered(p->workspace, FKV_WORKSPACE_HEADER_BYTES))
    return FKV_EWORKSPACE;
  return launch_decode_attn(*p, (cudaStream_t)stream);
}

int fkv_cache_append(const fkv_cache_append_params* p, void* stream) {
  if (!p || p->struct_size != sizeof(fkv_cache_append_params)) return FKV_EINVAL;
  if (p->dtype < FKV_BF16 || p->dtype > FKV_F32) return FKV_EDTYPE;
  if (p->D != 64 && p->D != 128) return FKV_EHEADDIM;
  if (p->B < 1 || p->Hkv < 1 || p->max_seq_len < 1) return FKV_EINVAL;
  const int es = dtype_size(p->dtype);
  int rc;
  if ((rc = check_view(p->k_new, es)) || (rc = check_view(p->v_new, es)) || (rc = check_view(p->k_cache, es)) ||
      (rc = check_view(p->v_cache, es)))
    return rc;
  return launch_cache_append(*p, (cudaStream_t)stream);
}

int fkv_compress(const fkv_compress_params* p, void* stream) {
  if (!p || p->struct_size != sizeof(fkv_compress_params)) return FKV_EINVAL;
  const fkv_score_params& sp = p->score;
  if (sp.mode == FKV_MODE_GEMFILTER) return FKV_EMODE;
  int rc = validate_score(sp);
  if (rc != FKV_OK) return rc;
  const int W = sp.window_size, P = sp.L - W;
  const int n_sel = p->capacity - W;
  if (n_sel < 1 || n_sel > P) return FKV_ETOPK;
  if (p->tsp_k < 0 || p->tsp_k > P) return FKV_ETOPK;
  if (p->tsp_k > 0 && (!p->tsp_out || sp.mode != FKV_MODE_FASTKV)) return FKV_EINVAL;
  if (!p->idx_out) return FKV_EINVAL;
  const int Hsel = sp.h1 - sp.h0;
  const int es = dtype_size(sp.dtype);
  if ((rc = check_view(p->v, es)) || (rc = check_view(p->k_out, es)) || (rc = check_view(p->v_out, es))) return rc;
  // the score call owns the front of the workspace, the top-k calls the rest
  const size_t score_ws = score_workspace_bytes(sp.mode, sp.B, sp.Hq, sp.Hkv, sp.L, W);
  const size_t head_ws = topk_workspace_bytes(sp.B * Hsel, P, n_sel);
  const size_t tsp_ws = p->tsp_k > 0 ? topk_workspace_bytes(sp.B, P, p->tsp_k) : 0;
  if (sp.workspace_bytes < score_ws + head_ws + tsp_ws) return FKV_EWORKSPACE;
  char* ws = reinterpret_cast<char*>(sp.workspace);
  cudaStream_t s = (cudaStream_t)stream;

  // short contexts (post-TSP layers): score + top-k + compaction in one cluster kernel
  if (small_compress_supported(*p)) return launch_small_compress(*p, s);

  // per-head top-k needs rows addressed as one flat list: require the dense [B,Hsel,P] layout
  if (sp.score_stride_h != P || sp.score_stride_b != (int64_t)Hsel * P) return FKV_EINVAL;
  fkv_topk_params tk{};
  tk.struct_size = sizeof(tk);
  tk.rows = sp.B * Hsel; tk.n = P; tk.k = n_sel; tk.order = FKV_ORDER_VALUE_DESC; tk.tail = 0; tk.nsum = 1;
  tk.src = sp.score; tk.src_stride_row = P; tk.src_stride_sum = 0;
  tk.out = p->idx_out; tk.out_stride_row = n_sel;
  tk.workspace = ws + score_ws; tk.workspace_bytes = head_ws;
  fkv_gather_kv_params g{};
  g.struct_size = sizeof(g);
  g.dtype = sp.dtype; g.B = sp.B; g.L = sp.L; g.D = sp.D; g.window_size = W; g.n_sel = n_sel;
  g.h0 = sp.h0; g.h1 = sp.h1; g.head_div = sp.mode == FKV_MODE_SNAPKV ? sp.Hq / sp.Hkv : 1;
  g.k = sp.k; g.v = p->v; g.idx = p->idx_out; g.idx_stride_b = (int64_t)Hsel * n_sel; g.idx_stride_h = n_sel;
  g.k_out = p->k_out; g.v_out = p->v_out;
  // K1+K2, building the per-row 12-bit histogram on the way when the fused select + compact kernel will consume it
  const bool use_select = select_compact_supported(tk, g);
  bool hist_used = false;
  fkv_score_params sp2 = sp;
  const int slot = (sp.hist_slot >= 1 && sp.hist_slot <= kHistSlots) ? sp.hist_slot : 1;
  sp2.hist_slot = (use_select && P > 4096) ? slot : 0;                      // short rows need no histogram
  if ((rc = run_score(sp2, s, &hist_used)) != FKV_OK) return rc;
  bool fused = false;
  if (use_select) {                                                         // K3 + K4 in one launch
    if ((rc = launch_select_compact(tk, g, ws, slot, hist_used, s)) != FKV_OK) return rc;
    fused = true;
  } else if ((rc = launch_topk_fused(tk, g, &fused, s)) != FKV_OK) return rc;   // K3 (+ K4 when the cluster kernel runs)

  if (p->tsp_k > 0) {
    fkv_topk_params tt{};
    tt.struct_size = sizeof(tt);
    tt.rows = sp.B; tt.n = P; tt.k = p->tsp_k; tt.order = FKV_ORDER_INDEX_ASC; tt.tail = W; tt.nsum = Hsel;
    tt.src = sp.score; tt.src_stride_row = (int64_t)Hsel * P; tt.src_stride_sum = P;
    tt.out = p->tsp_out; tt.out_stride_row = p->tsp_k + W;
    tt.imp_out = p->imp_out;
    tt.workspace = ws + score_ws + head_ws; tt.workspace_bytes = tsp_ws;
    if ((rc = launch_topk(tt, s)) != FKV_OK) return rc;
    if (p->imp_out && Hsel == 1)                     // a single selection head: the importance IS its score row
      FKV_CUDA_TRY(cudaMemcpy2DAsync(p->imp_out, (size_t)P * sizeof(float), sp.score, (size_t)sp.score_stride_b * sizeof(float),
                                     (size_t)P * sizeof(float), sp.B, cudaMemcpyDeviceToDevice, s));
  }
  return fused ? FKV_OK : launch_gather_kv(g, s);
}

}  // extern "C"
