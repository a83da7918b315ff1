"""Functional wrappers over the C ABI: one Python function per ``fkv_*`` entry point.

Each wrapper only marshals pointers / strides / sizes, allocates the outputs with torch and
enqueues the kernels on torch's current stream.  No arithmetic happens in Python.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L


class SelectOverlap:
    """Runs the select + compact stage (K3 + K4) of ``compress`` call i on a side stream, so that it overlaps whatever
    the caller's stream does next — in a prefill the layer's own attention and MLP, in bench.py the scoring kernel of
    the next layer.  The compact K/V of a prefill are not read before the first decode step, so nothing on the
    caller's stream has to wait for them until ``join()``.

    The two histogram slots of the workspace (include/fastkv_b200.h) are used alternately: call i + 1 may score while
    call i still selects; call i + 2 re-uses call i's slot and therefore waits for its select stage first.
    Works eagerly and under CUDA-graph capture (fork / join through events)."""

    SELECT_CTAS = 64                                   # measured on B200 (148 SMs): 84 scoring + 64 select CTAs is the best split

    def __init__(self, device=None, partition: bool = False):
        """``partition``: split the SMs between the scoring kernel on the caller's stream and the select + compact
        kernel on the side stream (back-to-back update_kv calls, as in bench.py: both kernels are persistent and
        latency-bound, so each loses little with fewer SMs and they truly run side by side).
        Without it the two kernels are never allowed to run at the same time (the next scoring kernel waits for the
        previous select stage): both are persistent kernels whose CTAs wait for each other, and two such grids that do
        not FIT the chip together can starve each other of SMs.  The stage then overlaps only what the caller does
        between two calls (attention, MLP), which is the case in a model forward."""
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        sms = torch.cuda.get_device_properties(self.device).multi_processor_count
        self.select_ctas = min(self.SELECT_CTAS, sms // 2) if partition else 0
        self.score_ctas = sms - self.select_ctas if partition else 0
        self.side = torch.cuda.Stream(self.device)
        self.done = [None, None]
        self.calls = 0
        self._keep = []                                # buffers the side stream may still be using

    def next_slot(self) -> int:
        slot = (self.calls % 2) + 1
        self.calls += 1
        return slot

    def join(self, stream=None):
        """Make ``stream`` (default: current) wait for every select + compact stage issued so far."""
        stream = torch.cuda.current_stream(self.device) if stream is None else stream
        for i, ev in enumerate(self.done):
            if ev is not None:
                stream.wait_event(ev)
                self.done[i] = None
        self._keep.clear()


class HostPipeline:
    """Round-robin CUDA streams for ``update_kv`` calls whose inputs live in pinned HOST memory.

    A blocking host call is H2D(K) -> kernels -> D2H(compact K/V) -> synchronize: the bus idles in one direction at
    any time and entirely while the host waits.  With a pipeline attached (``FastKVCluster.host_pipeline``) call i
    runs on stream i mod depth and returns WITHOUT synchronising, torch's ``non_blocking=True`` contract: the returned
    pinned tensors are complete after ``join()``.  The H2D copy of call i + 1 then overlaps the kernels and the D2H
    copy of call i (PCIe is full duplex, the GPU has separate copy engines per direction)."""

    def __init__(self, device=None, depth: int = 2):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(max(1, depth))]
        self.calls = 0

    def next_stream(self) -> torch.cuda.Stream:
        s = self.streams[self.calls % len(self.streams)]
        self.calls += 1
        return s

    def last_stream(self) -> torch.cuda.Stream:
        """The stream of the most recent call: work enqueued on it is ordered behind that call's results."""
        return self.streams[(self.calls - 1) % len(self.streams)]

    def join(self) -> None:
        """Block the host until every call issued so far has delivered its results."""
        for s in self.streams:
            s.synchronize()


def _score_params(mode, q, k, window_size, kernel_size, pooling, score, h0, h1, qh0, qh1, impl, ws):
    B, Hq, Lq, D = q.shape
    _, Hkv, Lk, Dk = k.shape
    if Lq != Lk or D != Dk:
        raise ValueError("fastkv_b200: query/key shape mismatch")
    if q.dtype != k.dtype:
        raise ValueError("fastkv_b200: query/key dtype mismatch")
    p = L.ScoreParams()
    p.struct_size = C.sizeof(L.ScoreParams)
    p.mode, p.dtype, p.impl = mode, L.dtype_code(q.dtype), impl
    p.B, p.Hq, p.Hkv, p.L, p.D = B, Hq, Hkv, Lq, D
    p.window_size, p.kernel_size, p.pooling = window_size, kernel_size, pooling
    p.h0, p.h1, p.qh0, p.qh1 = h0, h1, qh0, qh1
    p.q, p.k = L.view4(q), L.view4(k)
    p.score = score.data_ptr()
    p.score_stride_b, p.score_stride_h = score.stride(0), score.stride(1)
    p.workspace, p.workspace_bytes = ws.data_ptr(), ws.numel()
    return p


def _check_kv_pair(k: torch.Tensor, v: torch.Tensor):
    """The kernels copy V with K's dtype and geometry: refuse a mismatched pair instead of mis-reading V."""
    if v.dtype != k.dtype:
        raise ValueError(f"fastkv_b200: key/value dtype mismatch ({k.dtype} vs {v.dtype})")
    if tuple(v.shape) != tuple(k.shape):
        raise ValueError(f"fastkv_b200: key/value shape mismatch ({tuple(k.shape)} vs {tuple(v.shape)})")


def selection_heads(mode: int, Hq: int, Hkv: int) -> int:
    return Hkv if mode == L.MODE_FASTKV else (Hq if mode == L.MODE_SNAPKV else 1)


def score(q: torch.Tensor, k: torch.Tensor, *, window_size: int, kernel_size: int, pooling: str = "avgpool",
          mode: int = L.MODE_FASTKV, head_range: Optional[Tuple[int, int]] = None,
          q_head_range: Optional[Tuple[int, int]] = None, impl: int = L.IMPL_AUTO, hist_slot: int = 0,
          max_ctas: int = 0) -> torch.Tensor:
    """Pooled, group-summed selection scores [B, h1-h0, P] fp32 (P = L - W; GemFilter: P = L).

    ``hist_slot`` 1 / 2 additionally fills that slot's score-row histograms in the workspace (for a following
    ``fkv_select_compact``; the caller must then run it, it is what re-zeroes them); ``max_ctas`` caps the grid."""
    L.require_cuda(q, k)
    lib = L.load()
    B, Hq, Lq, _ = q.shape
    Hkv = k.shape[1]
    hs = selection_heads(mode, Hq, Hkv)
    h0, h1 = head_range if head_range is not None else (0, hs)
    qh0, qh1 = q_head_range if q_head_range is not None else (0, 0)
    if mode == L.MODE_GEMFILTER:
        h0, h1, window_size = 0, 1, 1
    P = Lq if mode == L.MODE_GEMFILTER else Lq - window_size
    out = torch.empty((B, h1 - h0, max(P, 0)), dtype=torch.float32, device=q.device)
    ws = L.workspace(q.device, lib.fkv_workspace_bytes(mode, B, Hq, Hkv, Lq, window_size, window_size + 1, 0))
    pool = L.pool_code(pooling) if mode != L.MODE_GEMFILTER else L.POOL_AVG
    p = _score_params(mode, q, k, window_size, kernel_size, pool, out, h0, h1, qh0, qh1, impl, ws)
    p.hist_slot, p.max_ctas = hist_slot, max_ctas
    L.check(lib.fkv_score(C.byref(p), L.stream_ptr(q.device)))
    return out


def topk(src: torch.Tensor, k: int, *, order: int = L.ORDER_VALUE_DESC, tail: int = 0, nsum: int = 1,
         return_importance: bool = False):
    """Deterministic top-k over the last dim.

    ``src`` is [R, n] (nsum == 1) or [R, nsum, n] (rows are summed over dim 1 in index order first).
    Returns int64 [R, k + tail] (and the summed rows [R, n] when ``return_importance``)."""
    L.require_cuda(src)
    lib = L.load()
    if src.dtype != torch.float32 or src.stride(-1) != 1:
        raise ValueError("fastkv_b200: top-k source must be fp32 with contiguous last dim")
    if nsum == 1:
        if src.dim() != 2:
            raise ValueError("fastkv_b200: expected [rows, n]")
        rows, n = src.shape
        stride_row, stride_sum = src.stride(0), 0
    else:
        if src.dim() != 3 or src.shape[1] != nsum:
            raise ValueError("fastkv_b200: expected [rows, nsum, n]")
        rows, _, n = src.shape
        stride_row, stride_sum = src.stride(0), src.stride(1)
    out = torch.empty((rows, k + tail), dtype=torch.int64, device=src.device)
    imp = torch.empty((rows, n), dtype=torch.float32, device=src.device) if (return_importance and nsum > 1) else None
    ws_ptr, ws_bytes = L.scratch(src.device, lib.fkv_topk_workspace_bytes(rows, n, k))
    p = L.TopkParams()
    p.struct_size = C.sizeof(L.TopkParams)
    p.rows, p.n, p.k, p.order, p.tail, p.nsum = rows, n, k, order, tail, nsum
    p.src, p.src_stride_row, p.src_stride_sum = src.data_ptr(), stride_row, stride_sum
    p.out, p.out_stride_row = out.data_ptr(), out.stride(0)
    p.imp_out = imp.data_ptr() if imp is not None else None
    p.workspace, p.workspace_bytes = ws_ptr, ws_bytes
    L.check(lib.fkv_topk(C.byref(p), L.stream_ptr(src.device)))
    return (out, imp) if return_importance else out


def gather_kv(k: torch.Tensor, v: torch.Tensor, idx: torch.Tensor, *, window_size: int,
              head_range: Optional[Tuple[int, int]] = None, head_div: int = 1,
              k_out: Optional[torch.Tensor] = None, v_out: Optional[torch.Tensor] = None):
    """Compact K/V rows selected by ``idx`` [B,Hsel,n_sel] and append the last W rows."""
    L.require_cuda(idx)
    L.require_cuda_or_pinned(k, v)
    lib = L.load()
    B, _, Lk, D = k.shape
    _check_kv_pair(k, v)
    Hsel, n_sel = idx.shape[1], idx.shape[2]
    h0, h1 = head_range if head_range is not None else (0, Hsel)
    if idx.dtype != torch.int64 or idx.stride(2) != 1:
        raise ValueError("fastkv_b200: idx must be int64 with contiguous last dim")
    Cc = n_sel + window_size
    # sources may be page-locked host tensors (zero-copy reads); the outputs always live on the GPU that runs the kernel
    if k_out is None:
        k_out = torch.empty((B, Hsel, Cc, D), dtype=k.dtype, device=idx.device)
    if v_out is None:
        v_out = torch.empty((B, Hsel, Cc, D), dtype=v.dtype, device=idx.device)
    p = L.GatherKvParams()
    p.struct_size = C.sizeof(L.GatherKvParams)
    p.dtype, p.B, p.L, p.D = L.dtype_code(k.dtype), B, Lk, D
    p.window_size, p.n_sel, p.h0, p.h1, p.head_div = window_size, n_sel, h0, h1, head_div
    p.k, p.v = L.view4(k), L.view4(v)
    p.idx, p.idx_stride_b, p.idx_stride_h = idx.data_ptr(), idx.stride(0), idx.stride(1)
    p.k_out, p.v_out = L.view4(k_out), L.view4(v_out)
    L.check(lib.fkv_gather_kv(C.byref(p), L.stream_ptr(idx.device)))     # idx is always on the GPU; K/V may be pinned host
    return k_out, v_out


def gather_rows(hidden: torch.Tensor, idx: torch.Tensor, position_ids: Optional[torch.Tensor] = None):
    """hidden [B,L,H] -> [B,T,H] rows picked by idx [B,T]; optionally position_ids [B|1, L] -> [B,T]."""
    L.require_cuda(idx, position_ids)
    L.require_cuda_or_pinned(hidden)
    lib = L.load()
    B, Ln, H = hidden.shape
    T = idx.shape[1]
    if hidden.stride(2) != 1 or idx.dtype != torch.int64 or idx.stride(1) != 1:
        raise ValueError("fastkv_b200: hidden needs a contiguous last dim and idx must be int64")
    es = hidden.element_size()
    out = torch.empty((B, T, H), dtype=hidden.dtype, device=idx.device)
    p = L.GatherRowsParams()
    p.struct_size = C.sizeof(L.GatherRowsParams)
    p.B, p.L, p.T, p.row_bytes = B, Ln, T, H * es
    p.src, p.src_stride_b_bytes, p.src_stride_l_bytes = hidden.data_ptr(), hidden.stride(0) * es, hidden.stride(1) * es
    p.out, p.out_stride_b_bytes, p.out_stride_l_bytes = out.data_ptr(), out.stride(0) * es, out.stride(1) * es
    p.idx, p.idx_stride_b = idx.data_ptr(), idx.stride(0)
    pos_out = None
    if position_ids is not None:
        if position_ids.dtype != torch.int64 or position_ids.dim() != 2 or position_ids.stride(1) != 1:
            raise ValueError("fastkv_b200: position_ids must be int64 [B or 1, L]")
        pos_out = torch.empty((B, T), dtype=torch.int64, device=idx.device)
        p.pos_in = position_ids.data_ptr()
        p.pos_stride_b = position_ids.stride(0) if position_ids.shape[0] == B and B > 1 else 0
        p.pos_out, p.pos_out_stride_b = pos_out.data_ptr(), pos_out.stride(0)
    L.check(lib.fkv_gather_rows(C.byref(p), L.stream_ptr(idx.device)))
    return (out, pos_out) if position_ids is not None else out


def _compress_overlapped(lib, ov, mode, q, k, v, W, kernel_size, pooling, capacity, h0, h1, impl, ws, score_buf, idx,
                         k_out, v_out, tsp, imp, tsp_k, after_score=None):
    """score on the current stream, select + compact on the side stream, TSP top-k on the current stream."""
    dev = q.device
    B, Hq, Lq, D = q.shape
    Hkv = k.shape[1]
    Hsel, P, n_sel = h1 - h0, Lq - W, capacity - W
    main = torch.cuda.current_stream(dev)
    slot = ov.next_slot()
    if ov.select_ctas == 0:                            # no SM partition: never co-schedule the two persistent kernels
        for ev in ov.done:
            if ev is not None:
                main.wait_event(ev)
    elif ov.done[slot - 1] is not None:                # the call that used this slot's tables two calls ago
        main.wait_event(ov.done[slot - 1])
    sp = _score_params(mode, q, k, W, kernel_size, L.pool_code(pooling), score_buf, h0, h1, 0, 0, impl, ws)
    sp.hist_slot = slot if P > 4096 else 0
    sp.max_ctas = ov.score_ctas
    have_hist = int(sp.hist_slot != 0 and lib.fkv_score_impl_name(C.byref(sp)) == b"tensor")
    L.check(lib.fkv_score(C.byref(sp), main.cuda_stream))
    if after_score is not None:
        after_score(score_buf)
    fork = torch.cuda.Event()
    fork.record(main)
    ov.side.wait_event(fork)
    sc = L.SelectCompactParams()
    sc.struct_size = C.sizeof(L.SelectCompactParams)
    sc.hist_slot, sc.have_hist = slot, have_hist
    sc.score, sc.idx_out = score_buf.data_ptr(), idx.data_ptr()
    g = sc.gather
    g.struct_size = C.sizeof(L.GatherKvParams)
    g.dtype, g.B, g.L, g.D = L.dtype_code(k.dtype), B, Lq, D
    g.window_size, g.n_sel, g.h0, g.h1 = W, n_sel, h0, h1
    g.head_div = Hq // Hkv if mode == L.MODE_SNAPKV else 1
    g.k, g.v = L.view4(k), L.view4(v)
    g.idx, g.idx_stride_b, g.idx_stride_h = idx.data_ptr(), idx.stride(0), idx.stride(1)
    g.k_out, g.v_out = L.view4(k_out), L.view4(v_out)
    sc.workspace, sc.workspace_bytes = ws.data_ptr(), ws.numel()
    sc.max_ctas = ov.select_ctas
    with torch.cuda.stream(ov.side):
        L.check(lib.fkv_select_compact(C.byref(sc), ov.side.cuda_stream))
        done = torch.cuda.Event()
        done.record(ov.side)
    ov.done[slot - 1] = done
    if not torch.cuda.is_current_stream_capturing():
        for t in (score_buf, idx, k_out, v_out, k, v):  # allocated on the caller's stream, used on the side stream
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(ov.side)
    ov._keep.append(ws)
    if tsp_k:
        ws_ptr, ws_bytes = L.scratch(dev, lib.fkv_topk_workspace_bytes(B, P, tsp_k))
        tp = L.TopkParams()
        tp.struct_size = C.sizeof(L.TopkParams)
        tp.rows, tp.n, tp.k, tp.order, tp.tail, tp.nsum = B, P, tsp_k, L.ORDER_INDEX_ASC, W, Hsel
        tp.src, tp.src_stride_row, tp.src_stride_sum = score_buf.data_ptr(), score_buf.stride(0), score_buf.stride(1)
        tp.out, tp.out_stride_row = tsp.data_ptr(), tsp.stride(0)
        tp.imp_out = imp.data_ptr() if imp is not None else None
        tp.workspace, tp.workspace_bytes = ws_ptr, ws_bytes
        L.check(lib.fkv_topk(C.byref(tp), main.cuda_stream))
        if imp is not None and Hsel == 1:
            imp.copy_(score_buf[:, 0])


def decode_attn(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, seq_len: int, *,
                seq_len_dev: Optional[torch.Tensor] = None, scale: float = 0.0, num_splits: int = 0,
                out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """q [B,Hq,D], caches [B,Hkv,S_max,D] -> out [B,Hq,D]; attends over rows [0, seq_len)."""
    L.require_cuda(q, k_cache, v_cache, seq_len_dev)
    lib = L.load()
    B, Hq, D = q.shape
    Hkv, Smax = k_cache.shape[1], k_cache.shape[2]
    if q.stride(2) != 1:
        raise ValueError("fastkv_b200: q needs a contiguous last dim")
    if out is None:
        out = torch.empty((B, Hq, D), dtype=q.dtype, device=q.device)
    ws_ptr, ws_bytes = L.scratch(q.device, lib.fkv_decode_workspace_bytes(B, Hq, D, Smax))
    p = L.DecodeAttnParams()
    p.struct_size = C.sizeof(L.DecodeAttnParams)
    p.dtype, p.B, p.Hq, p.Hkv, p.D = L.dtype_code(q.dtype), B, Hq, Hkv, D
    p.seq_len, p.max_seq_len, p.num_splits, p.scale = seq_len, Smax, num_splits, scale
    p.q, p.q_stride_b, p.q_stride_h = q.data_ptr(), q.stride(0), q.stride(1)
    p.k_cache, p.v_cache = L.view4(k_cache), L.view4(v_cache)
    p.out, p.out_stride_b, p.out_stride_h = out.data_ptr(), out.stride(0), out.stride(1)
    p.seq_len_dev = seq_len_dev.data_ptr() if seq_len_dev is not None else None
    p.workspace, p.workspace_bytes = ws_ptr, ws_bytes
    L.check(lib.fkv_decode_attn(C.byref(p), L.stream_ptr(q.device)))
    return out


def cache_append(k_new: torch.Tensor, v_new: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor,
                 seq_len: int, *, seq_len_dev: Optional[torch.Tensor] = None, bump: bool = False) -> None:
    """Write k_new/v_new [B,Hkv,1,D] into row ``seq_len`` (or *seq_len_dev) of the caches."""
    L.require_cuda(k_new, v_new, k_cache, v_cache, seq_len_dev)
    lib = L.load()
    B, Hkv, _, D = k_new.shape
    p = L.CacheAppendParams()
    p.struct_size = C.sizeof(L.CacheAppendParams)
    p.dtype, p.B, p.Hkv, p.D = L.dtype_code(k_new.dtype), B, Hkv, D
    p.seq_len, p.max_seq_len, p.bump = seq_len, k_cache.shape[2], int(bump)
    p.k_new, p.v_new, p.k_cache, p.v_cache = L.view4(k_new), L.view4(v_new), L.view4(k_cache), L.view4(v_cache)
    p.seq_len_dev = seq_len_dev.data_ptr() if seq_len_dev is not None else None
    L.check(lib.fkv_cache_append(C.byref(p), L.stream_ptr(k_new.device)))


def compress(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, window_size: int, kernel_size: int,
             pooling: str, capacity: int, tsp_len: int = 0, mode: int = L.MODE_FASTKV,
             head_range: Optional[Tuple[int, int]] = None, impl: int = L.IMPL_AUTO,
             k_out: Optional[torch.Tensor] = None, v_out: Optional[torch.Tensor] = None,
             return_aux: bool = False, overlap: Optional[SelectOverlap] = None, after_score=None):
    """The fused update_kv chain: score -> per-head top-k -> KV compaction [-> TSP top-k].

    Returns (k_out [B,Hsel,C,D], v_out, tsp_idx [B,T] | None) and, with ``return_aux``, a dict with
    the scores [B,Hsel,P], per-head indices [B,Hsel,C-W] and (TSP) importance [B,P].

    ``overlap``: run the select + compact stage on the overlap object's side stream (k_out / v_out / idx are then
    complete only after ``overlap.join()``); the scores and the TSP indices stay on the current stream.
    ``after_score(score)``: called once the scoring kernel is enqueued (on the overlapped path: before the select
    stage is; otherwise after the whole chain) — the hook the multi-GPU score exchange starts from."""
    L.require_cuda(q, k)
    L.require_cuda_or_pinned(v)
    lib = L.load()
    B, Hq, Lq, D = q.shape
    Hkv = k.shape[1]
    _check_kv_pair(k, v)
    hs = selection_heads(mode, Hq, Hkv)
    h0, h1 = head_range if head_range is not None else (0, hs)
    Hsel = h1 - h0
    W = window_size
    P = Lq - W
    tsp_k = (tsp_len - W) if tsp_len else 0
    if tsp_len and tsp_k < 1:
        raise ValueError("fastkv_b200: tsp_length must exceed window_size")
    dev = q.device
    score_buf = torch.empty((B, Hsel, max(P, 1)), dtype=torch.float32, device=dev)
    idx = torch.empty((B, Hsel, max(capacity - W, 1)), dtype=torch.int64, device=dev)
    if k_out is None:
        k_out = torch.empty((B, Hsel, capacity, D), dtype=k.dtype, device=dev)
    if v_out is None:
        v_out = torch.empty((B, Hsel, capacity, D), dtype=v.dtype, device=dev)
    tsp = torch.empty((B, tsp_len), dtype=torch.int64, device=dev) if tsp_k else None
    imp = torch.empty((B, P), dtype=torch.float32, device=dev) if (tsp_k and return_aux) else None
    ws = L.workspace(dev, lib.fkv_workspace_bytes(mode, B, Hq, Hkv, Lq, W, capacity, tsp_k))
    if overlap is not None and v.is_cuda and B * Hsel <= L.SELECT_MAX_ROWS and 1 <= capacity - W <= 2048 and capacity - W <= P:
        _compress_overlapped(lib, overlap, mode, q, k, v, W, kernel_size, pooling, capacity, h0, h1, impl, ws, score_buf, idx,
                             k_out, v_out, tsp, imp, tsp_k, after_score)
        if return_aux:
            return k_out, v_out, tsp, dict(score=score_buf, idx=idx, importance=imp)
        return k_out, v_out, tsp
    p = L.CompressParams()
    p.struct_size = C.sizeof(L.CompressParams)
    p.score = _score_params(mode, q, k, W, kernel_size, L.pool_code(pooling), score_buf, h0, h1, 0, 0, impl, ws)
    p.capacity, p.tsp_k = capacity, tsp_k
    p.v = L.view4(v)
    p.idx_out = idx.data_ptr()
    p.k_out, p.v_out = L.view4(k_out), L.view4(v_out)
    p.tsp_out = tsp.data_ptr() if tsp is not None else None
    p.imp_out = imp.data_ptr() if imp is not None else None
    L.check(lib.fkv_compress(C.byref(p), L.stream_ptr(dev)))
    if after_score is not None:
        after_score(score_buf)
    if return_aux:
        return k_out, v_out, tsp, dict(score=score_buf, idx=idx, importance=imp)
    return k_out, v_out, tsp
