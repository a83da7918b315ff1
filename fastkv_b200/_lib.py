"""ctypes binding of libfastkv_b200.so (the C ABI declared in include/fastkv_b200.h).

There is deliberately no fallback: if the shared library is missing and cannot be built with
nvcc, importing a kernel entry point raises.  Nothing here touches the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

from . import _build

# ---- enums (mirror include/fastkv_b200.h) -----------------------------------------------------
ABI_VERSION = 6
BF16, F16, F32 = 0, 1, 2
POOL_AVG, POOL_MAX = 0, 1
MODE_FASTKV, MODE_SNAPKV, MODE_GEMFILTER = 0, 1, 2
ORDER_VALUE_DESC, ORDER_INDEX_ASC = 0, 1
IMPL_AUTO, IMPL_SIMPLE, IMPL_TENSOR = 0, 1, 2
EPOOL = -4
ETIMEOUT = -11
MAX_PEERS = 8
EXCHANGE_HEADER = 256                      # FKV_EXCHANGE_HEADER_BYTES
ALLREDUCE_HEADER = 2048                    # FKV_ALLREDUCE_HEADER_BYTES

_DTYPES = {torch.bfloat16: BF16, torch.float16: F16, torch.float32: F32}
_POOLS = {"avgpool": POOL_AVG, "maxpool": POOL_MAX}


class View4(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("stride_b", C.c_int64), ("stride_h", C.c_int64), ("stride_l", C.c_int64)]


class ScoreParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("mode", C.c_int32), ("dtype", C.c_int32), ("impl", C.c_int32),
        ("B", C.c_int32), ("Hq", C.c_int32), ("Hkv", C.c_int32), ("L", C.c_int32), ("D", C.c_int32),
        ("window_size", C.c_int32), ("kernel_size", C.c_int32), ("pooling", C.c_int32),
        ("h0", C.c_int32), ("h1", C.c_int32), ("qh0", C.c_int32), ("qh1", C.c_int32),
        ("q", View4), ("k", View4),
        ("score", C.c_void_p), ("score_stride_b", C.c_int64), ("score_stride_h", C.c_int64),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("hist_slot", C.c_int32), ("max_ctas", C.c_int32),
    ]


class TopkParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("rows", C.c_int32), ("n", C.c_int32), ("k", C.c_int32),
        ("order", C.c_int32), ("tail", C.c_int32), ("nsum", C.c_int32),
        ("src", C.c_void_p), ("src_stride_row", C.c_int64), ("src_stride_sum", C.c_int64),
        ("out", C.c_void_p), ("out_stride_row", C.c_int64),
        ("imp_out", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class GatherKvParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32), ("B", C.c_int32), ("L", C.c_int32), ("D", C.c_int32),
        ("window_size", C.c_int32), ("n_sel", C.c_int32), ("h0", C.c_int32), ("h1", C.c_int32), ("head_div", C.c_int32),
        ("k", View4), ("v", View4),
        ("idx", C.c_void_p), ("idx_stride_b", C.c_int64), ("idx_stride_h", C.c_int64),
        ("k_out", View4), ("v_out", View4),
    ]


class SelectCompactParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("hist_slot", C.c_int32), ("have_hist", C.c_int32),
        ("score", C.c_void_p), ("idx_out", C.c_void_p),
        ("gather", GatherKvParams),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("max_ctas", C.c_int32),
    ]


class GatherRowsParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("B", C.c_int32), ("L", C.c_int32), ("T", C.c_int32),
        ("row_bytes", C.c_int64),
        ("src", C.c_void_p), ("src_stride_b_bytes", C.c_int64), ("src_stride_l_bytes", C.c_int64),
        ("out", C.c_void_p), ("out_stride_b_bytes", C.c_int64), ("out_stride_l_bytes", C.c_int64),
        ("idx", C.c_void_p), ("idx_stride_b", C.c_int64),
        ("pos_in", C.c_void_p), ("pos_stride_b", C.c_int64),
        ("pos_out", C.c_void_p), ("pos_out_stride_b", C.c_int64),
    ]


class DecodeAttnParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32), ("B", C.c_int32), ("Hq", C.c_int32), ("Hkv", C.c_int32),
        ("D", C.c_int32), ("seq_len", C.c_int32), ("max_seq_len", C.c_int32), ("num_splits", C.c_int32),
        ("scale", C.c_float),
        ("q", C.c_void_p), ("q_stride_b", C.c_int64), ("q_stride_h", C.c_int64),
        ("k_cache", View4), ("v_cache", View4),
        ("out", C.c_void_p), ("out_stride_b", C.c_int64), ("out_stride_h", C.c_int64),
        ("seq_len_dev", C.c_void_p),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
    ]


class CacheAppendParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("dtype", C.c_int32), ("B", C.c_int32), ("Hkv", C.c_int32), ("D", C.c_int32),
        ("seq_len", C.c_int32), ("max_seq_len", C.c_int32), ("bump", C.c_int32),
        ("k_new", View4), ("v_new", View4), ("k_cache", View4), ("v_cache", View4),
        ("seq_len_dev", C.c_void_p),
    ]


class CompressParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("score", ScoreParams),
        ("capacity", C.c_int32), ("tsp_k", C.c_int32),
        ("v", View4),
        ("idx_out", C.c_void_p),
        ("k_out", View4), ("v_out", View4),
        ("tsp_out", C.c_void_p), ("imp_out", C.c_void_p),
    ]


class ExchangeParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("rank", C.c_int32), ("world", C.c_int32),
        ("B", C.c_int32), ("heads_local", C.c_int32), ("P", C.c_int64),
        ("score", C.c_void_p), ("score_stride_b", C.c_int64), ("score_stride_h", C.c_int64),
        ("region", C.c_void_p * 8), ("region_bytes", C.c_size_t),
        ("max_ctas", C.c_int32), ("timeout_ms", C.c_int32),
    ]


class AllreduceParams(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32), ("rank", C.c_int32), ("world", C.c_int32), ("dtype", C.c_int32),
        ("n", C.c_int64), ("data", C.c_void_p),
        ("region", C.c_void_p * 8), ("region_bytes", C.c_size_t),
        ("timeout_ms", C.c_int32),
    ]


EXPORTS = [
    "fkv_abi_version", "fkv_strerror", "fkv_workspace_bytes", "fkv_workspace_init", "fkv_workspace_status",
    "fkv_topk_workspace_bytes", "fkv_decode_workspace_bytes",
    "fkv_score_impl_name", "fkv_launch_count",
    "fkv_score", "fkv_topk", "fkv_gather_kv", "fkv_select_compact", "fkv_gather_rows", "fkv_decode_attn", "fkv_cache_append",
    "fkv_compress",
    "fkv_exchange_push", "fkv_exchange_wait", "fkv_exchange_status", "fkv_allreduce_sum",
    "fkv_peer_alloc", "fkv_peer_open", "fkv_peer_close", "fkv_peer_free",
]

_lib = None
_lock = threading.Lock()


def load():
    """Load (building first if the .so is absent or stale and nvcc is available) and return the CDLL."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB
        if not os.path.exists(path) or (_build.is_stale() and _have_nvcc()):
            _build.build_library()
        lib = C.CDLL(path)
        lib.fkv_abi_version.restype = C.c_int
        if lib.fkv_abi_version() != ABI_VERSION:
            raise RuntimeError(f"libfastkv_b200.so ABI {lib.fkv_abi_version()} != binding {ABI_VERSION}; rebuild")
        lib.fkv_strerror.restype = C.c_char_p
        lib.fkv_strerror.argtypes = [C.c_int]
        lib.fkv_workspace_bytes.restype = C.c_size_t
        lib.fkv_workspace_bytes.argtypes = [C.c_int32] * 8
        lib.fkv_workspace_init.restype = C.c_int
        lib.fkv_workspace_init.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        lib.fkv_workspace_status.restype = C.c_int
        lib.fkv_workspace_status.argtypes = [C.c_void_p, C.c_void_p]
        lib.fkv_topk_workspace_bytes.restype = C.c_size_t
        lib.fkv_topk_workspace_bytes.argtypes = [C.c_int32] * 3
        lib.fkv_decode_workspace_bytes.restype = C.c_size_t
        lib.fkv_decode_workspace_bytes.argtypes = [C.c_int32] * 4
        lib.fkv_launch_count.restype = C.c_uint64
        lib.fkv_score_impl_name.restype = C.c_char_p
        lib.fkv_score_impl_name.argtypes = [C.POINTER(ScoreParams)]
        for name, st in [("fkv_score", ScoreParams), ("fkv_topk", TopkParams), ("fkv_gather_kv", GatherKvParams),
                         ("fkv_select_compact", SelectCompactParams),
                         ("fkv_gather_rows", GatherRowsParams), ("fkv_decode_attn", DecodeAttnParams),
                         ("fkv_cache_append", CacheAppendParams), ("fkv_compress", CompressParams),
                         ("fkv_exchange_push", ExchangeParams), ("fkv_exchange_wait", ExchangeParams),
                         ("fkv_allreduce_sum", AllreduceParams)]:
            fn = getattr(lib, name)
            fn.restype = C.c_int
            fn.argtypes = [C.POINTER(st), C.c_void_p]
        lib.fkv_exchange_status.restype = C.c_int
        lib.fkv_exchange_status.argtypes = [C.c_void_p, C.c_void_p]
        lib.fkv_peer_alloc.restype = C.c_int
        lib.fkv_peer_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
        lib.fkv_peer_open.restype = C.c_int
        lib.fkv_peer_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        lib.fkv_peer_close.restype = C.c_int
        lib.fkv_peer_close.argtypes = [C.c_void_p]
        lib.fkv_peer_free.restype = C.c_int
        lib.fkv_peer_free.argtypes = [C.c_void_p]
        _lib = lib
    return _lib


def _have_nvcc():
    from shutil import which
    return which(os.environ.get("NVCC", "nvcc")) is not None


def check(rc: int):
    """Map a C return code to the exception the reference would raise."""
    if rc == 0:
        return
    msg = load().fkv_strerror(rc).decode()
    if rc < 0:
        raise ValueError(f"fastkv_b200: {msg} (code {rc})")
    raise RuntimeError(f"fastkv_b200: CUDA error {rc}: {msg}")


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise ValueError(f"fastkv_b200: unsupported dtype {dt}") from None


def pool_code(name: str) -> int:
    try:
        return _POOLS[name]
    except KeyError:
        raise ValueError("Pooling method not supported") from None      # baselines/fastkv/utils.py:110


def view4(t: torch.Tensor) -> View4:
    """[B,H,L,D] tensor (any b/h/l strides, unit stride on D) -> fkv_view4."""
    if t.dim() != 4 or t.stride(3) != 1:
        raise ValueError("fastkv_b200: expected a [B,H,L,D] tensor with contiguous last dim")
    return View4(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("fastkv_b200 has no CPU path: tensors must live on a CUDA device")


def require_cuda_or_pinned(*tensors):
    """Gather SOURCES may also be page-locked host memory: the kernels then read only the selected
    rows over PCIe/NVLink-C2C (zero-copy) instead of the caller staging the whole tensor."""
    for t in tensors:
        if t is not None and not t.is_cuda and not (hasattr(t, "is_pinned") and t.is_pinned()):
            raise RuntimeError("fastkv_b200 has no CPU path: tensors must live on a CUDA device (or pinned host memory for gather sources)")


_workspaces = {}


def workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """Grow-only per-(device, stream) scratch buffer; kernels on one stream run in order."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 16 << 20), dtype=torch.uint8, device=device)
        check(load().fkv_workspace_init(buf.data_ptr(), buf.numel(), torch.cuda.current_stream(device).cuda_stream))
        _workspaces[key] = buf
    return buf


SELECT_MAX_ROWS = 64
WS_HEADER = 256 + 128 * 32 + 2 * 128 * 4096 * 4 + 128 * 4096 * 8 + 4096   # FKV_WORKSPACE_HEADER_BYTES: the library's persistent region


def scratch(device: torch.device, nbytes: int):
    """(ptr, bytes) of the shared per-stream buffer for the entry points that use their workspace as plain scratch
    (fkv_topk, fkv_decode_attn).  The library itself starts past the scoring header (ABI 4), so the same pointer
    serves every entry point; ``nbytes`` comes from the *_workspace_bytes queries, which include the header."""
    buf = workspace(device, nbytes)
    return buf.data_ptr(), buf.numel()


def workspace_status(device: torch.device) -> int:
    """fkv_workspace_status of the current stream's buffer (synchronises; debugging / tests)."""
    buf = workspace(device, WS_HEADER)
    return load().fkv_workspace_status(buf.data_ptr(), torch.cuda.current_stream(device).cuda_stream)
