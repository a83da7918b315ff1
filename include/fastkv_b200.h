/*
 * fastkv_b200 — C ABI of the B200-native TSP + KV-cache-compression hot path.
 *
 * This is the drop-in boundary: every entry point below replaces a stretch of ATen calls in the
 * reference (dongwonjo/FastKV), cited as file:line relative to the reference tree.  All functions
 *   - take raw DEVICE pointers, explicit element strides and sizes (no torch types),
 *   - never allocate, never synchronise, never throw: they only enqueue kernels on `stream`
 *     (a cudaStream_t passed as void*; NULL = legacy default stream),
 *   - return 0 (FKV_OK), a negative FKV_E* argument error, or a positive cudaError_t.
 * Tensors are "[B,H,L,D] views": element (b,h,l,d) lives at ptr + b*stride_b + h*stride_h +
 * l*stride_l + d (strides in ELEMENTS, D contiguous), which covers both [B,H,L,D] and the
 * [B,L,H,D] memory that HF attention layers actually produce.
 */
#ifndef FASTKV_B200_H_
#define FASTKV_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FKV_ABI_VERSION 6

/* Every `workspace` argument below is one kind of buffer: device memory, 256-byte aligned, passed ONCE through
 * fkv_workspace_init.  Its first FKV_WORKSPACE_HEADER_BYTES are persistent state owned by the library: 256 bytes
 * for the scoring kernel (launch epoch, departure counter, error word) and, for each of the two histogram SLOTS,
 * FKV_SELECT_MAX_ROWS score rows' worth of barrier counters (32 bytes), first- and second-level 4096-bin histograms
 * (filled by the scoring kernel / the select + compact kernel, which consumes both and leaves everything zeroed
 * again) and candidate lists (4096 x 8 bytes); finally 1024 arrival tickets of the decode-attention kernel.  Two slots let a caller run fkv_select_compact of call i on a side
 * stream while fkv_score of call i+1 already fills the other slot.
 * Every entry point starts its own scratch PAST that region by itself, so one buffer can be shared by all calls
 * issued on a stream and no caller ever has to offset a pointer.  The *_workspace_bytes queries include it. */
#define FKV_SELECT_MAX_ROWS 64
#define FKV_WORKSPACE_HEADER_BYTES (256 + 128 * 32 + 2 * 128 * 4096 * 4 + 128 * 4096 * 8 + 4096)

/* ---- enums ------------------------------------------------------------------------------- */
enum { FKV_BF16 = 0, FKV_F16 = 1, FKV_F32 = 2 };                 /* element type of Q/K/V/hidden   */
enum { FKV_POOL_AVG = 0, FKV_POOL_MAX = 1 };                     /* 'avgpool' / 'maxpool'          */
enum { FKV_MODE_FASTKV = 0, FKV_MODE_SNAPKV = 1, FKV_MODE_GEMFILTER = 2 };
enum { FKV_ORDER_VALUE_DESC = 0, FKV_ORDER_INDEX_ASC = 1 };     /* output order of fkv_topk       */
enum { FKV_IMPL_AUTO = 0, FKV_IMPL_SIMPLE = 1, FKV_IMPL_TENSOR = 2 };   /* scoring kernel choice  */

enum {
  FKV_OK = 0,
  FKV_EINVAL = -1,      /* null pointer / non-positive size / inconsistent head counts            */
  FKV_EDTYPE = -2,      /* dtype not in {bf16, f16, f32}                                           */
  FKV_EHEADDIM = -3,    /* head_dim not in {64, 128}                                               */
  FKV_EPOOL = -4,       /* pooling not avg/max, or kernel_size even / out of range                 */
  FKV_ETOPK = -5,       /* k < 1 or k > n                                                          */
  FKV_EALIGN = -6,      /* pointer or stride breaks the 16-byte row alignment the kernels need     */
  FKV_EWORKSPACE = -7,  /* workspace missing or smaller than fkv_workspace_bytes()                 */
  FKV_EMODE = -8,       /* unknown mode / impl, or combination not supported                       */
  FKV_EWINDOW = -9,     /* window_size < 1, > 64, or >= L                                          */
  FKV_ERANGE = -10,     /* head range [h0,h1) outside the selection heads                          */
  FKV_ETIMEOUT = -11    /* a peer never showed up in fkv_exchange_push / _wait (fkv_exchange_status) */
};

/* ---- tensor view ------------------------------------------------------------------------- */
typedef struct fkv_view4 {          /* [B,H,L,D] view, D contiguous; strides in elements          */
  const void* ptr;
  int64_t stride_b, stride_h, stride_l;
} fkv_view4;

/* ---- K1 + K2: observation-window scoring, pooling, GQA group-sum ------------------------ */
/*
 * Replaces baselines/fastkv/utils.py:93-112 (repeat_kv, matmul, mask, softmax, window row-sum,
 * avg/max_pool1d, group-sum), baselines/snapkv/utils.py:72-87 (same without the group-sum, one
 * selection head per query head) and baselines/gemfilter/utils.py:25-33 (last-row raw dot
 * products, sum over query heads, avg_pool1d).
 *
 * "Selection heads": FASTKV -> the Hkv kv heads; SNAPKV -> the Hq query heads (K of query head
 * hq is kv head hq / (Hq/Hkv)); GEMFILTER -> one virtual head 0 that sums the query heads in
 * [qh0,qh1) (used for head sharding; default all).  The call computes selection heads [h0,h1).
 *
 * Output: score[b, h-h0, j] fp32 for j in [0, P), P = L - W (GEMFILTER: P = L), laid out with
 * score_stride_b / score_stride_h (elements).
 */
typedef struct fkv_score_params {
  uint32_t struct_size;             /* sizeof(fkv_score_params), ABI guard                        */
  int32_t mode, dtype, impl;
  int32_t B, Hq, Hkv, L, D;
  int32_t window_size, kernel_size, pooling;
  int32_t h0, h1;                   /* selection-head range computed by this call                 */
  int32_t qh0, qh1;                 /* GEMFILTER only: query heads summed; 0,0 = all              */
  fkv_view4 q;                      /* [B,Hq,L,D]                                                 */
  fkv_view4 k;                      /* [B,Hkv,L,D]                                                */
  float* score;                     /* [B,h1-h0,P]                                                */
  int64_t score_stride_b, score_stride_h;
  void* workspace;
  size_t workspace_bytes;
  int32_t hist_slot;                /* 0: none.  1 or 2: additionally add the 12-bit histogram of every score written
                                       to that slot's tables in the workspace, for a following fkv_select_compact with
                                       the same slot (tensor-core path, B*(h1-h0) <= FKV_SELECT_MAX_ROWS; otherwise
                                       the field is ignored and fkv_select_compact must be told hist_slot's tables
                                       are empty through have_hist = 0) */
  int32_t max_ctas;                 /* 0: use every SM.  > 0: cap the persistent grid of the tensor-core kernel, leaving
                                       the other SMs to a kernel running concurrently on another stream (e.g. the
                                       fkv_select_compact of the previous call)                                       */
} fkv_score_params;

int fkv_score(const fkv_score_params* p, void* stream);

/* ---- K3: deterministic top-k ------------------------------------------------------------- */
/*
 * Replaces Tensor.topk at baselines/fastkv/utils.py:113 (per head, order VALUE_DESC) and the
 * sum(dim=-2) + topk + arange/cat + sort chain at :127-130 (TSP, order INDEX_ASC with
 * `tail` = W trailing window indices appended), also gemfilter/utils.py:34 + :60.
 *
 * Row r, element j is   sum_{s < nsum} src[r*src_stride_row + s*src_stride_sum + j]   (summed in
 * increasing s: the fixed-order TSP head-sum of :127; nsum = 1 for a plain row).
 * Total order: value descending, then index ascending (the reference's tie order is unspecified).
 * Output out[r*out_stride_row + i], i < k, then `tail` extra entries n, n+1, .. n+tail-1.
 * If `imp_out` is non-NULL and nsum > 1 the summed row is also written there ([rows,n] dense).
 */
typedef struct fkv_topk_params {
  uint32_t struct_size;
  int32_t rows, n, k, order, tail, nsum;
  const float* src;
  int64_t src_stride_row, src_stride_sum;
  int64_t* out;
  int64_t out_stride_row;
  float* imp_out;
  void* workspace;
  size_t workspace_bytes;
} fkv_topk_params;

int fkv_topk(const fkv_topk_params* p, void* stream);

/* ---- K4: KV compaction ------------------------------------------------------------------- */
/*
 * Replaces the expand + 2x gather + 2x cat at baselines/fastkv/utils.py:114-121 (and the SnapKV
 * twin snapkv/utils.py:88-101).  For selection head h in [h0,h1):
 *   out[b,h-h0,i,:]        = src[b, h / head_div, idx[b,h-h0,i], :]      i < n_sel
 *   out[b,h-h0,n_sel+r,:]  = src[b, h / head_div, L-W+r, :]              r < W
 * for src in {k, v}.  Outputs are [B,h1-h0,*,D] views (stride_l = D normally) so they can be a
 * slice of a pre-allocated cache.  head_div = 1 (FASTKV) or Hq/Hkv (SNAPKV).  Bit-exact copy.
 */
typedef struct fkv_gather_kv_params {
  uint32_t struct_size;
  int32_t dtype, B, L, D, window_size, n_sel, h0, h1, head_div;
  fkv_view4 k, v;
  const int64_t* idx;               /* [B,h1-h0,n_sel]                                            */
  int64_t idx_stride_b, idx_stride_h;
  fkv_view4 k_out, v_out;           /* writable; [B,h1-h0,n_sel+W,D]                              */
} fkv_gather_kv_params;

int fkv_gather_kv(const fkv_gather_kv_params* p, void* stream);

/* ---- K3 + K4 fused: per-head top-k with the KV compaction behind it, one launch ----------- */
/*
 * Replaces Tensor.topk + expand + 2x gather + 2x cat at baselines/fastkv/utils.py:113-121 (SnapKV twin
 * snapkv/utils.py:88-101) in ONE kernel: for every selection head row of `score` ([B, h1-h0, P] dense fp32, P = L - W)
 * the k = gather.n_sel best keys in (score desc, index asc) order go to idx_out [B, h1-h0, k] and their K / V rows +
 * the W window rows to gather.k_out / v_out (same semantics as fkv_gather_kv; gather.idx is ignored).
 * hist_slot (1 or 2) names the table set of the workspace to use; have_hist != 0 says fkv_score already filled that
 * slot's histogram for exactly these rows (fkv_score_params.hist_slot), 0 makes the kernel build it itself.
 * Limits: B*(h1-h0) <= FKV_SELECT_MAX_ROWS, k <= 2048 (FKV_EMODE otherwise: use fkv_topk + fkv_gather_kv).
 * Calls that share a workspace AND a slot must be stream-ordered; different slots may overlap.
 */
typedef struct fkv_select_compact_params {
  uint32_t struct_size;
  int32_t hist_slot, have_hist;
  const float* score;               /* [B, h1-h0, P] dense                                        */
  int64_t* idx_out;                 /* [B, h1-h0, n_sel] dense                                    */
  fkv_gather_kv_params gather;
  void* workspace;
  size_t workspace_bytes;
  int32_t max_ctas;                 /* 0: use every SM.  > 0: cap the grid (see fkv_score_params.max_ctas)  */
} fkv_select_compact_params;

int fkv_select_compact(const fkv_select_compact_params* p, void* stream);

/* ---- K4b: hidden-state / position-id propagation ---------------------------------------- */
/*
 * Replaces the two torch.gather calls at baselines/fastkv/llama_model.py:255-257 (Mistral
 * mistral_model.py:220-222):  out[b,i,:] = src[b, idx[b,i], :]  (row_bytes per row, multiple of
 * 16) and, when pos_in != NULL, pos_out[b,i] = pos_in[b*pos_stride_b + idx[b,i]] (int64;
 * pos_stride_b = 0 broadcasts a [1,L] position_ids).  Bit-exact copy.
 */
typedef struct fkv_gather_rows_params {
  uint32_t struct_size;
  int32_t B, L, T;
  int64_t row_bytes;
  const void* src;   int64_t src_stride_b_bytes, src_stride_l_bytes;
  void* out;         int64_t out_stride_b_bytes, out_stride_l_bytes;
  const int64_t* idx; int64_t idx_stride_b;
  const int64_t* pos_in; int64_t pos_stride_b;
  int64_t* pos_out;  int64_t pos_out_stride_b;
} fkv_gather_rows_params;

int fkv_gather_rows(const fkv_gather_rows_params* p, void* stream);

/* ---- K5: split-K decode attention over the compact cache --------------------------------- */
/*
 * Replaces flash_attn_func at q_len == 1 (baselines/fastkv/llama_model.py:79-81 via :181-183;
 * mistral_model.py:143-153): out[b,hq,:] = softmax(q[b,hq,:] . K[b,hq/G,0:S,:]^T * scale) V.
 * q/out are [B,Hq,D] with stride_b / stride_h (elements); caches are [B,Hkv,S_max,D] views.
 * S is `seq_len`, or *seq_len_dev (int32 on device, read by the kernel) when that is non-NULL so
 * a captured CUDA graph can be replayed while the cache grows; either way S is clamped to
 * max_seq_len (> 0), the number of rows the cache views really have.
 */
typedef struct fkv_decode_attn_params {
  uint32_t struct_size;
  int32_t dtype, B, Hq, Hkv, D, seq_len, max_seq_len, num_splits;   /* num_splits 0 = auto      */
  float scale;                                                       /* 0 = 1/sqrt(D)            */
  const void* q;  int64_t q_stride_b, q_stride_h;
  fkv_view4 k_cache, v_cache;
  void* out;      int64_t out_stride_b, out_stride_h;
  const int32_t* seq_len_dev;
  void* workspace;
  size_t workspace_bytes;
} fkv_decode_attn_params;

int fkv_decode_attn(const fkv_decode_attn_params* p, void* stream);

/*
 * Appends one new K/V row per (b, kv head) at row *seq_len_dev (or seq_len) of the cache and,
 * if bump != 0, increments *seq_len_dev afterwards (replaces DynamicCache.update's torch.cat at
 * baselines/fastkv/llama_model.py:144).  k_new/v_new are [B,Hkv,1,D] views.  A full cache
 * (row index >= max_seq_len) drops the row AND leaves *seq_len_dev unchanged.
 */
typedef struct fkv_cache_append_params {
  uint32_t struct_size;
  int32_t dtype, B, Hkv, D, seq_len, max_seq_len, bump;
  fkv_view4 k_new, v_new;
  fkv_view4 k_cache, v_cache;       /* writable                                                   */
  int32_t* seq_len_dev;
} fkv_cache_append_params;

int fkv_cache_append(const fkv_cache_append_params* p, void* stream);

/* ---- fused chain: one host call for update_kv -------------------------------------------- */
/*
 * score -> per-head top-k -> KV compaction [-> TSP top-k], i.e. the whole of
 * FastKVCluster.update_kv (baselines/fastkv/utils.py:93-130) or SnapKVCluster.update_kv
 * (baselines/snapkv/utils.py:72-101) behind one call, for selection heads [h0,h1).
 * tsp_k > 0 additionally emits tsp_out[b, 0:tsp_k+W] (ascending) from the sum over the heads
 * computed by THIS call (single-GPU case; multi-GPU callers exchange `score` first and call
 * fkv_topk themselves).  score.hist_slot (0 = library's choice) selects the table set used between
 * the scoring and the select + compact kernel.
 */
typedef struct fkv_compress_params {
  uint32_t struct_size;
  fkv_score_params score;           /* score.score / workspace as for fkv_score                   */
  int32_t capacity;                 /* C: rows kept per head including the W window rows          */
  int32_t tsp_k;                    /* T - W, or 0                                                */
  fkv_view4 v;                      /* [B,Hkv,L,D]                                                */
  int64_t* idx_out;                 /* [B,h1-h0,C-W]                                              */
  fkv_view4 k_out, v_out;           /* [B,h1-h0,C,D]                                              */
  int64_t* tsp_out;                 /* [B,T] or NULL                                              */
  float* imp_out;                   /* [B,P] or NULL                                              */
} fkv_compress_params;

int fkv_compress(const fkv_compress_params* p, void* stream);

/* ---- the TSP layer's score exchange over peer memory (multi-GPU, KV heads sharded over the ranks) -------------- */
/* baselines/fastkv/utils.py:127 sums the scores of ALL kv heads before the TSP top-k, so KV-head-sharded ranks must
 * see each other's [B, heads_local, P] fp32 score rows once per prefill.  Instead of a library all-gather, every
 * rank owns one PEER REGION (device memory mapped into every rank of the box: CUDA IPC or a torch symmetric-memory
 * allocation, NVLink / NVSwitch underneath) and the ranks store their rows straight into each other's regions:
 *
 *   region = [FKV_EXCHANGE_HEADER_BYTES of flags and counters, zero-initialised once]
 *            [gathered scores: fp32 [B][world * heads_local][P], head-major, the layout fkv_topk sums with nsum]
 *
 * fkv_exchange_push  (i) tells every peer "my previous TSP top-k is done, my region may be overwritten" and waits for
 *                    the same word from every peer, (ii) stores the local rows into every rank's region (its own
 *                    included) with 16-byte stores, (iii) releases a per-source DATA flag in every region.
 * fkv_exchange_wait  one small CTA that acquires the DATA flags of all peers; kernels enqueued behind it on the same
 *                    stream (fkv_topk with src = region + FKV_EXCHANGE_HEADER_BYTES, nsum = world * heads_local)
 *                    read the complete gathered rows.
 * Both only enqueue kernels (CUDA-graph capturable, no host synchronisation, no NCCL); work that does not depend on
 * the exchange (the rank-local per-head selection and compaction) can be enqueued between the two.  One exchange may
 * be in flight per region; push / wait / consumer must be enqueued on ONE stream per rank, in that order, and every
 * rank must issue the same sequence of exchanges.  A wait that sees no peer for `timeout_ms` (default 20 000) gives
 * up, records it in the region (fkv_exchange_status) and lets the stream continue rather than hang the GPU.        */
#define FKV_MAX_PEERS 8
#define FKV_EXCHANGE_HEADER_BYTES 256
typedef struct fkv_exchange_params {
  uint32_t struct_size;
  int32_t rank, world;                /* world <= FKV_MAX_PEERS                                              */
  int32_t B, heads_local;             /* local rows: [B, heads_local, P]                                     */
  int64_t P;
  const float* score;                 /* local scores (device); element strides below, P contiguous          */
  int64_t score_stride_b, score_stride_h;
  void* region[FKV_MAX_PEERS];        /* region[r] = rank r's peer region AS MAPPED IN THIS PROCESS          */
  size_t region_bytes;                /* >= FKV_EXCHANGE_HEADER_BYTES + B * world * heads_local * P * 4      */
  int32_t max_ctas;                   /* 0 = library default; MUST be the same on every rank (receivers count push CTAs) */
  int32_t timeout_ms;                 /* 0 = 20 000                                                           */
} fkv_exchange_params;
int fkv_exchange_push(const fkv_exchange_params* p, void* stream);
int fkv_exchange_wait(const fkv_exchange_params* p, void* stream);
/* SYNCHRONISES `stream`; FKV_OK, or FKV_ETIMEOUT when a push / wait on this rank's region ever timed out. */
int fkv_exchange_status(const void* local_region, void* stream);
/* Set-up helpers for callers without their own peer-memory allocator (these four DO allocate / map and are not
 * stream-ordered): cudaMalloc + zero + cudaIpcGetMemHandle, cudaIpcOpenMemHandle (enables peer access lazily),
 * cudaIpcCloseMemHandle, cudaFree.  `handle` is the 64-byte cudaIpcMemHandle_t, to be sent to the other ranks. */
int fkv_peer_alloc(size_t bytes, void** ptr_out, unsigned char handle_out[64]);
int fkv_peer_open(const unsigned char handle[64], void** ptr_out);
int fkv_peer_close(void* mapped_ptr);
int fkv_peer_free(void* ptr);

/* ---- small all-reduce over peer memory (tensor-parallel DECODE steps of a head-sharded host) ---------------------- */
/* Outside the §8 hot path proper, but what makes a KV-head-sharded decode step CUDA-graph capturable: the two hidden-state
 * sums per layer of a tensor-parallel model are 8 KiB at batch 1 — pure latency — and NCCL cannot be captured on every
 * stack.  One kernel: every rank stores its partial into a slot of every peer's region, fences, raises a per-(source,
 * CTA) flag to the exchange number, waits for its peers' flags and adds the world partials in RANK ORDER in fp32 (the
 * same bits on every rank), in place.  Two slots alternate by the exchange number kept in the region, so no handshake is
 * needed and a captured graph replays without host state.  Same region ownership rules as fkv_exchange_*; this entry
 * point wants its OWN region: [FKV_ALLREDUCE_HEADER_BYTES of counters and flags, zero-initialised once]
 * [2 slots x world sources x slot bytes], slot bytes = (region_bytes - header) / (2 * world) rounded down to 256; a
 * tensor may take up to HALF a slot (sums of <= 64 KiB that fit a quarter travel as tagged 16-byte lines — payload and
 * exchange number in one store, no fence — in the other half).                                                      */
#define FKV_ALLREDUCE_HEADER_BYTES 2048
#define FKV_ALLREDUCE_MAX_CTAS 32
typedef struct fkv_allreduce_params {
  uint32_t struct_size;
  int32_t rank, world;                /* world <= FKV_MAX_PEERS                                              */
  int32_t dtype;                      /* FKV_BF16 / FKV_F16 / FKV_F32                                        */
  int64_t n;                          /* elements; n * element size: a multiple of 16, at most half a slot   */
  void* data;                         /* in / out (device, 16-byte aligned): partial in, sum over ranks out  */
  void* region[FKV_MAX_PEERS];        /* region[r] = rank r's all-reduce region AS MAPPED IN THIS PROCESS    */
  size_t region_bytes;
  int32_t timeout_ms;                 /* 0 = 20 000                                                           */
} fkv_allreduce_params;
int fkv_allreduce_sum(const fkv_allreduce_params* p, void* stream);

/* ---- queries ----------------------------------------------------------------------------- */
/* Bytes of scratch fkv_score / fkv_topk / fkv_compress need for these sizes (upper bound).     */
/* A workspace must be passed through fkv_workspace_init ONCE after allocation: it zeroes the buffer,
 * i.e. the header (departure counter + launch epoch of the scoring kernel) and the region where CTAs
 * exchange epoch-tagged partial softmax statistics, and registers the address range with the library.
 * fkv_score / fkv_compress return FKV_EWORKSPACE (nothing is launched, nothing traps) for a buffer
 * that never went through it.  Calls keep the header consistent. */
int fkv_workspace_init(void* workspace, size_t workspace_bytes, void* stream);
/* Debugging aid, SYNCHRONISES `stream`: reads the header back and returns FKV_OK, or FKV_EWORKSPACE
 * when the header is not an initialised one or a kernel flagged it as inconsistent (a kernel never
 * aborts the context: it flags the header and carries on). */
int fkv_workspace_status(const void* workspace, void* stream);
size_t fkv_workspace_bytes(int32_t mode, int32_t B, int32_t Hq, int32_t Hkv, int32_t L, int32_t window_size,
                           int32_t capacity, int32_t tsp_k);
size_t fkv_topk_workspace_bytes(int32_t rows, int32_t n, int32_t k);
size_t fkv_decode_workspace_bytes(int32_t B, int32_t Hq, int32_t D, int32_t max_seq_len);
const char* fkv_strerror(int code);
int fkv_abi_version(void);
/* Number of kernels this library has launched in this process (bench.py's gpu_launches). */
uint64_t fkv_launch_count(void);
/* Name of the scoring kernel that FKV_IMPL_AUTO picks for these parameters ("simple"/"tensor"). */
const char* fkv_score_impl_name(const fkv_score_params* p);

#ifdef __cplusplus
}
#endif
#endif /* FASTKV_B200_H_ */
