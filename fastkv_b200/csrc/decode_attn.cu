// K5: split-K decode attention (q_len == 1) over the compact KV cache, GQA-aware, plus the
// in-place cache append that replaces DynamicCache.update's torch.cat.
// Replaces flash_attn_func at baselines/fastkv/llama_model.py:79-81 (decode branch :143-145).
//
//   decode_split_kernel    grid (splits, Hkv, B): each CTA streams its slice of K/V rows once and
//                          serves all G = Hq/Hkv query heads of that kv head (K/V bytes read once)
//   decode_combine_kernel  grid (Hq, B): merges the split partials (log-sum-exp weighted)
#include "common.cuh"

namespace fkv {
namespace {

constexpr int kDecThreads = 128;
constexpr int kDecWarps = kDecThreads / 32;

struct DecodeArgs {
  const void* q; int64_t q_sb, q_sh;
  const void* kc; int64_t k_sb, k_sh, k_sl;
  const void* vc; int64_t v_sb, v_sh, v_sl;
  void* out; int64_t o_sb, o_sh;
  float* part_o;      // [B][Hq][splits][D]
  float2* part_ml;    // [B][Hq][splits]
  const int32_t* seq_len_dev;
  int seq_len, B, Hq, Hkv, D, G, splits;
  float scale_log2;   // scale * log2(e)
};

template <typename T, int VEC> struct LaneVec;          // VEC elements per lane as fp32
template <int VEC> struct LaneVec<__nv_bfloat16, VEC> {
  __device__ static inline void load(const __nv_bfloat16* p, float* o) {
    if (VEC == 4) {
      const uint2 r = *reinterpret_cast<const uint2*>(p);
      o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
      o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
    } else {
      const uint32_t r = *reinterpret_cast<const uint32_t*>(p);
      o[0] = __uint_as_float(r << 16); o[1] = __uint_as_float(r & 0xffff0000u);
    }
  }
  __device__ static inline void store(__nv_bfloat16* p, const float* o) {
#pragma unroll
    for (int i = 0; i < VEC; i += 2)
      *reinterpret_cast<__nv_bfloat162*>(p + i) = __floats2bfloat162_rn(o[i], o[i + 1]);
  }
};
template <int VEC> struct LaneVec<__half, VEC> {
  __device__ static inline void load(const __half* p, float* o) {
#pragma unroll
    for (int i = 0; i < VEC; i += 2) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(p + i));
      o[i] = f.x; o[i + 1] = f.y;
    }
  }
  __device__ static inline void store(__half* p, const float* o) {
#pragma unroll
    for (int i = 0; i < VEC; i += 2) *reinterpret_cast<__half2*>(p + i) = __floats2half2_rn(o[i], o[i + 1]);
  }
};
template <int VEC> struct LaneVec<float, VEC> {
  __device__ static inline void load(const float* p, float* o) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) o[i] = p[i];
  }
  __device__ static inline void store(float* p, const float* o) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) p[i] = o[i];
  }
};

template <typename T, int D, int G>
__global__ void __launch_bounds__(kDecThreads)
decode_split_kernel(DecodeArgs a) {
  constexpr int VEC = D / 32;
  __shared__ float sm_o[kDecWarps][G][D];
  __shared__ float2 sm_ml[kDecWarps][G];

  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.seq_len_dev ? *a.seq_len_dev : a.seq_len;
  const int per = ceil_div(S, a.splits);
  const int s0 = min(split * per, S), s1 = min(s0 + per, S);

  float qv[G][VEC], o[G][VEC], m[G], l[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const T* qp = reinterpret_cast<const T*>(a.q) + (int64_t)b * a.q_sb + (int64_t)(kvh * G + g) * a.q_sh + lane * VEC;
    LaneVec<T, VEC>::load(qp, qv[g]);
    m[g] = -INFINITY; l[g] = 0.f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) o[g][i] = 0.f;
  }
  const T* kb = reinterpret_cast<const T*>(a.kc) + (int64_t)b * a.k_sb + (int64_t)kvh * a.k_sh + lane * VEC;
  const T* vb = reinterpret_cast<const T*>(a.vc) + (int64_t)b * a.v_sb + (int64_t)kvh * a.v_sh + lane * VEC;

  constexpr int U = 4;                               // rows in flight per warp
  for (int base = s0 + warp * U; base < s1; base += kDecWarps * U) {
    float kk[U][VEC], vv[U][VEC];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int row = min(base + u, s1 - 1);         // clamp: tail rows are loaded but not used
      LaneVec<T, VEC>::load(kb + (int64_t)row * a.k_sl, kk[u]);
      LaneVec<T, VEC>::load(vb + (int64_t)row * a.v_sl, vv[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (base + u < s1) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float d = 0.f;
#pragma unroll
          for (int i = 0; i < VEC; ++i) d = fmaf(qv[g][i], kk[u][i], d);
          d = warp_sum(d) * a.scale_log2;
          const float mn = fmaxf(m[g], d);
          const float alpha = exp2f(m[g] - mn);      // first row: exp2(-inf) = 0
          const float p = exp2f(d - mn);
          l[g] = l[g] * alpha + p;
#pragma unroll
          for (int i = 0; i < VEC; ++i) o[g][i] = fmaf(o[g][i], alpha, p * vv[u][i]);
          m[g] = mn;
        }
      }
    }
  }

#pragma unroll
  for (int g = 0; g < G; ++g) {
#pragma unroll
    for (int i = 0; i < VEC; ++i) sm_o[warp][g][lane * VEC + i] = o[g][i];
    if (lane == 0) sm_ml[warp][g] = make_float2(m[g], l[g]);
  }
  __syncthreads();
  // merge the warps: thread t owns output column t (D <= 128 == kDecThreads)
  for (int g = 0; g < G; ++g) {
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) mm = fmaxf(mm, sm_ml[w][g].x);
    float ll = 0.f, acc = 0.f;
#pragma unroll
    for (int w = 0; w < kDecWarps; ++w) {
      const float2 ml = sm_ml[w][g];
      const float f = (ml.x > -INFINITY) ? exp2f(ml.x - mm) : 0.f;
      ll += ml.y * f;
      if (threadIdx.x < D) acc += sm_o[w][g][threadIdx.x] * f;
    }
    const int64_t slot = ((int64_t)b * a.Hq + kvh * G + g) * a.splits + split;
    if (threadIdx.x < D) a.part_o[slot * D + threadIdx.x] = acc;
    if (threadIdx.x == 0) a.part_ml[slot] = make_float2(mm, ll);
  }
}

template <typename T, int D>
__global__ void __launch_bounds__(D)
decode_combine_kernel(DecodeArgs a) {
  const int hq = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const int64_t base = ((int64_t)b * a.Hq + hq) * a.splits;
  float mm = -INFINITY;
  for (int s = 0; s < a.splits; ++s) mm = fmaxf(mm, a.part_ml[base + s].x);
  float ll = 0.f, acc = 0.f;
  for (int s = 0; s < a.splits; ++s) {
    const float2 ml = a.part_ml[base + s];
    const float f = (ml.x > -INFINITY) ? exp2f(ml.x - mm) : 0.f;
    ll += ml.y * f;
    acc += a.part_o[(base + s) * D + t] * f;
  }
  const float r = acc / ll;
  T* op = reinterpret_cast<T*>(a.out) + (int64_t)b * a.o_sb + (int64_t)hq * a.o_sh + t;
  if constexpr (sizeof(T) == 4) *op = r;
  else if constexpr (std::is_same<T, __half>::value) *op = __float2half_rn(r);
  else *op = __float2bfloat16_rn(r);
}

template <typename T, int D, int G>
int run_decode(const DecodeArgs& a, cudaStream_t s) {
  decode_split_kernel<T, D, G><<<dim3(a.splits, a.Hkv, a.B), kDecThreads, 0, s>>>(a);
  FKV_LAUNCH_CHECK();
  decode_combine_kernel<T, D><<<dim3(a.Hq, a.B), D, 0, s>>>(a);
  FKV_LAUNCH_CHECK();
  return FKV_OK;
}

template <typename T, int D>
int dispatch_group(const DecodeArgs& a, cudaStream_t s) {
  switch (a.G) {
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

int pick_splits(int B, int Hkv, int max_seq_len) {
  // enough CTAs to fill the chip once, at least 64 rows per split
  int splits = ceil_div(num_sms(), B * Hkv);
  const int cap = ceil_div(max_seq_len, 64);
  if (splits > cap) splits = cap;
  if (splits < 1) splits = 1;
  if (splits > 64) splits = 64;
  return splits;
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
__global__ void bump_kernel(int32_t* p) { *p += 1; }

}  // namespace

size_t decode_workspace_bytes(int B, int Hq, int D, int max_seq_len) {
  const int splits = 64;
  (void)max_seq_len;
  return align_up((size_t)B * Hq * splits * D * sizeof(float), 256) + align_up((size_t)B * Hq * splits * sizeof(float2), 256);
}

int launch_decode_attn(const fkv_decode_attn_params& p, cudaStream_t s) {
  DecodeArgs a{};
  a.q = p.q; a.q_sb = p.q_stride_b; a.q_sh = p.q_stride_h;
  a.kc = p.k_cache.ptr; a.k_sb = p.k_cache.stride_b; a.k_sh = p.k_cache.stride_h; a.k_sl = p.k_cache.stride_l;
  a.vc = p.v_cache.ptr; a.v_sb = p.v_cache.stride_b; a.v_sh = p.v_cache.stride_h; a.v_sl = p.v_cache.stride_l;
  a.out = p.out; a.o_sb = p.out_stride_b; a.o_sh = p.out_stride_h;
  a.seq_len_dev = p.seq_len_dev; a.seq_len = p.seq_len;
  a.B = p.B; a.Hq = p.Hq; a.Hkv = p.Hkv; a.D = p.D; a.G = p.Hq / p.Hkv;
  const int bound = p.max_seq_len > 0 ? p.max_seq_len : p.seq_len;
  a.splits = p.num_splits > 0 ? (p.num_splits > 64 ? 64 : p.num_splits) : pick_splits(p.B, p.Hkv, bound);
  const float scale = p.scale != 0.f ? p.scale : 1.0f / sqrtf((float)p.D);
  a.scale_log2 = scale * 1.4426950408889634f;
  const size_t need_o = align_up((size_t)p.B * p.Hq * a.splits * p.D * sizeof(float), 256);
  const size_t need = need_o + (size_t)p.B * p.Hq * a.splits * sizeof(float2);
  if (!p.workspace || p.workspace_bytes < need) return FKV_EWORKSPACE;
  a.part_o = reinterpret_cast<float*>(p.workspace);
  a.part_ml = reinterpret_cast<float2*>(reinterpret_cast<char*>(p.workspace) + need_o);
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
    bump_kernel<<<1, 1, 0, s>>>(p.seq_len_dev);
    FKV_LAUNCH_CHECK();
  }
  return FKV_OK;
}

}  // namespace fkv
