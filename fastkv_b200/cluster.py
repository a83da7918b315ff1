"""Per-layer compressor objects with the reference's plugin surface.

``FastKVCluster`` mirrors ``baselines/fastkv/utils.py:49-134`` (constructor, ``reset``,
``update_kv`` signature, mutable knob attributes, early-return and ``None`` semantics);
``SnapKVCluster`` mirrors ``baselines/snapkv/utils.py:26-102``; ``standard_dis_index`` /
``find_context`` mirror ``baselines/gemfilter/utils.py:25-56``.  All arithmetic happens in the
CUDA kernels behind ``fastkv_b200.ops``; there is no torch or CPU fallback.
"""
from __future__ import annotations


import torch

from . import _lib as L
from . import ops


def _to_pinned_host(t: torch.Tensor) -> torch.Tensor:
    """Device result -> page-locked host tensor (asynchronous; the caller synchronises once)."""
    out = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
    out.copy_(t, non_blocking=True)
    return out


def _stage_to_device(t: torch.Tensor, device: torch.device) -> torch.Tensor:
    """Host buffer -> device (the e2e path of bench.py): pinned memory copies asynchronously."""
    return t.to(device, non_blocking=True)


class FastKVCluster:
    """Drop-in for the reference's ``FastKVCluster`` (baselines/fastkv/utils.py:49)."""

    def __init__(self, window_size=8, max_capacity_prompt=512, kernel_size=7, pooling='avgpool',
                 tsp_layer=False, tsp_length=2048, tsp_rate=0.25, retain_rate=0.25, eviction_mode="constant"):
        self.reset(window_size, max_capacity_prompt, kernel_size, pooling, tsp_layer, tsp_length, tsp_rate,
                   retain_rate, eviction_mode)
        # B200-side extras (not part of the reference surface)
        self.score_impl = L.IMPL_AUTO           # scoring kernel choice (auto / simple / tensor)
        self.shard = None                       # fastkv_b200.dist.HeadShard for KV-head-sharded runs
        self.keep_aux = False                   # keep scores / indices of the last call in self.last_aux
        self.last_aux = None
        # optional callable (batch, kv_heads, rows, head_dim, dtype, device) -> (k_buf, v_buf) [B,Hkv,rows,D] views:
        # the compaction kernel then writes straight into the caller's cache slot.  Calls on HOST
        # tensors ask with device = cpu: the compact K/V are then copied into the returned (pinned) host buffers instead of
        # freshly allocated ones (None = allocate)
        self.out_provider = None
        # optional ops.SelectOverlap: the select + compact stage runs on its side stream and the returned K/V are
        # complete only after overlap.join() (the model glue joins at the end of the prefill, bench.py at the end of a step)
        self.overlap = None
        # optional ops.HostPipeline: calls on pinned HOST tensors run on its streams round-robin and return without
        # synchronising (results complete after host_pipeline.join()); None = every host call blocks until its results landed
        self.host_pipeline = None

    def reset(self, window_size=8, max_capacity_prompt=512, kernel_size=7, pooling='avgpool', tsp_layer=False,
              tsp_length=2048, tsp_rate=0.25, retain_rate=0.25, eviction_mode="constant"):
        self.window_size = window_size
        self.max_capacity_prompt = max_capacity_prompt
        assert self.max_capacity_prompt - self.window_size > 0      # baselines/fastkv/utils.py:54
        self.kernel_size = kernel_size
        self.pooling = pooling
        self.tsp_layer = tsp_layer
        # the reference's reset() stores a 1-tuple here (utils.py:73, trailing comma); that is a
        # defect and is not reproduced
        self.tsp_length = tsp_length
        self.retain_rate = retain_rate
        self.eviction_mode = eviction_mode
        self.tsp_rate = tsp_rate

    # ------------------------------------------------------------------------------------------
    def update_kv(self, key_states, query_states, value_states, attention_mask, num_key_value_groups, layer_idx=None):
        """(key_states[B,Hkv,L,D], query_states[B,Hq,L,D], value_states) -> (k[B,Hkv,C,D], v, tsp_idx | None).

        Contract of baselines/fastkv/utils.py:80-134: prefill only (k_len == q_len);
        ``attention_mask`` and ``layer_idx`` are accepted and ignored; when ``q_len <
        max_capacity_prompt`` the inputs are returned by reference with ``None``."""
        assert key_states.shape[-2] == query_states.shape[-2]
        bsz, num_heads, q_len, head_dim = query_states.shape

        if self.eviction_mode == "proportional":
            self.max_capacity_prompt = int(q_len * self.retain_rate)

        if q_len < self.max_capacity_prompt:
            return key_states, value_states, None

        if self.tsp_layer and self.eviction_mode == "proportional":
            self.tsp_length = int(q_len * self.tsp_rate)
        want_tsp = bool(self.tsp_layer and q_len > self.tsp_length)

        host_io = not key_states.is_cuda
        if host_io and self.host_pipeline is not None and key_states.is_pinned():
            with torch.cuda.stream(self.host_pipeline.next_stream()):
                return self._update_kv_impl(key_states, query_states, value_states, num_key_value_groups, want_tsp, blocking=False)
        return self._update_kv_impl(key_states, query_states, value_states, num_key_value_groups, want_tsp, blocking=True)

    def _update_kv_impl(self, key_states, query_states, value_states, num_key_value_groups, want_tsp, blocking):
        bsz, num_heads, q_len, head_dim = query_states.shape
        host_io = not key_states.is_cuda
        if host_io:
            if not torch.cuda.is_available():
                raise RuntimeError("fastkv_b200 has no CPU path: a CUDA device is required")
            dev = torch.device("cuda", torch.cuda.current_device())
            w = self.window_size
            # only the window rows of Q are read (utils.py:94): ship those, not the whole tensor
            q_win = _stage_to_device(query_states[:, :, q_len - w:, :], dev)
            query_states = _RowOffsetView(q_win.contiguous(), q_len)
            key_states = _stage_to_device(key_states, dev)
            # V is only ever gathered (C rows per head): page-locked V stays on the host and the
            # compaction kernel reads just those rows; pageable V has to be staged
            if not value_states.is_pinned():
                value_states = _stage_to_device(value_states, dev)
        else:
            L.require_cuda(key_states, query_states, value_states)

        if key_states.shape[1] * num_key_value_groups != num_heads:
            raise ValueError("fastkv_b200: num_key_value_groups does not match the head counts")

        k_dst = v_dst = None
        if self.out_provider is not None and not host_io:
            k_dst, v_dst = self.out_provider(bsz, key_states.shape[1], self.max_capacity_prompt, head_dim,
                                             key_states.dtype, key_states.device)
        if self.shard is not None and self.shard.world_size > 1 and want_tsp:
            k, v, tsp = self._update_kv_sharded(key_states, query_states, value_states, k_dst, v_dst)
        else:
            res = ops.compress(query_states, key_states, value_states, window_size=self.window_size,
                               kernel_size=self.kernel_size, pooling=self.pooling,
                               capacity=self.max_capacity_prompt, tsp_len=self.tsp_length if want_tsp else 0,
                               mode=L.MODE_FASTKV, impl=self.score_impl, return_aux=self.keep_aux,
                               k_out=k_dst, v_out=v_dst, overlap=None if host_io else self.overlap)
            k, v, tsp = res[0], res[1], res[2]
            if self.keep_aux:
                self.last_aux = res[3]
        if host_io:
            dst = None
            if self.out_provider is not None:          # the caller's own (pinned) host cache slot for this layer
                dst = self.out_provider(bsz, k.shape[1], k.shape[2], head_dim, k.dtype, torch.device("cpu"))
            if dst is not None:
                dst[0].copy_(k, non_blocking=True)
                dst[1].copy_(v, non_blocking=True)
                k, v = dst
            else:
                k, v = _to_pinned_host(k), _to_pinned_host(v)
            tsp = _to_pinned_host(tsp) if tsp is not None else None
            if blocking:
                torch.cuda.current_stream().synchronize()  # results are on the host when we return
        return k, v, tsp

    def _update_kv_sharded(self, key_states, query_states, value_states, k_dst=None, v_dst=None):
        """KV-head-sharded TSP layer: local compress, exchange per-head scores, identical global top-k."""
        w = self.window_size
        # (p2p exchange: the stores to the peers are enqueued right behind the scoring kernel, before the rank-local
        # selection and compaction, which do not depend on them — SURVEY.md §8e)
        k, v, _, aux = ops.compress(query_states, key_states, value_states, window_size=w,
                                    kernel_size=self.kernel_size, pooling=self.pooling,
                                    capacity=self.max_capacity_prompt, tsp_len=0, mode=L.MODE_FASTKV,
                                    impl=self.score_impl, return_aux=True, overlap=self.overlap,
                                    k_out=k_dst, v_out=v_dst, after_score=self.shard.begin_exchange)
        tsp, imp = self.shard.global_tsp(aux["score"], self.tsp_length - w, w)
        if self.keep_aux:
            aux["importance"] = imp
            self.last_aux = aux
        return k, v, tsp


class _RowOffsetView:
    """Duck-typed [B,Hq,L,D] tensor backed by a [B,Hq,W,D] window buffer (host-staging path only)."""

    def __init__(self, win: torch.Tensor, q_len: int):
        self._win = win
        b, h, w, d = win.shape
        self.shape = torch.Size((b, h, q_len, d))
        self.dtype = win.dtype
        self.device = win.device
        self.is_cuda = True
        self._offset_rows = q_len - w

    def dim(self):
        return 4

    def stride(self, i=None):
        s = self._win.stride()
        return s if i is None else s[i]

    def data_ptr(self):
        return self._win.data_ptr() - self._offset_rows * self._win.stride(2) * self._win.element_size()


class SnapKVCluster:
    """Drop-in for the reference's ``SnapKVCluster`` (baselines/snapkv/utils.py:26).

    ``update_kv`` accepts what the reference's call site passes — K/V already expanded to one head
    per query head (baselines/snapkv/llama_model.py:385-386) — and also the un-expanded
    [B,Hkv,L,D] tensors, in which case the kernels read each kv head once for its whole group and
    the 4x ``repeat_kv`` copy never happens."""

    def __init__(self, window_size=64, max_capacity_prompt=256 + 64, kernel_size=5, pooling='avgpool', merge=None,
                 recent_size=32, ratio=0.4, retain_rate=0.25, eviction_mode="constant"):
        self.window_size = window_size
        self.max_capacity_prompt = max_capacity_prompt
        assert self.max_capacity_prompt - self.window_size > 0
        self.kernel_size = kernel_size
        self.pooling = pooling
        self.merge = merge
        self.recent_size = recent_size
        self.ratio = ratio
        self.retain_rate = retain_rate
        self.eviction_mode = eviction_mode
        self.score_impl = L.IMPL_AUTO
        self.keep_aux = False
        self.last_aux = None
        self.overlap = None                     # see FastKVCluster.overlap

    def reset(self, window_size=64, max_capacity_prompt=256 + 64, kernel_size=5, pooling='avgpool', merge=None,
              retain_rate=0.25, eviction_mode="constant"):
        # the reference's reset() reads undefined names `ratio` / `recent_size` (snapkv/utils.py:51-52);
        # here those two attributes are simply left untouched
        self.window_size = window_size
        self.max_capacity_prompt = max_capacity_prompt
        assert self.max_capacity_prompt - self.window_size > 0
        self.kernel_size = kernel_size
        self.pooling = pooling
        self.merge = merge
        self.retain_rate = retain_rate
        self.eviction_mode = eviction_mode

    def update_kv(self, key_states, query_states, value_states, attention_mask, num_key_value_groups):
        assert key_states.shape[-2] == query_states.shape[-2]
        bsz, num_heads, q_len, head_dim = query_states.shape
        if self.eviction_mode == "proportional":
            self.max_capacity_prompt = int(q_len * self.retain_rate)
        if q_len < self.max_capacity_prompt:
            return key_states, value_states
        if self.merge is not None:
            # merge_kv is undefined in the reference (snapkv/utils.py:93): calling it there raises NameError
            raise NotImplementedError("merge is not supported (merge_kv is undefined in the reference)")
        L.require_cuda(key_states, query_states, value_states)
        expanded = key_states.shape[1] == num_heads
        # expanded K (Hkv == Hq): every head is its own group, i.e. FASTKV arithmetic with G = 1
        mode = L.MODE_FASTKV if expanded else L.MODE_SNAPKV
        res = ops.compress(query_states, key_states, value_states, window_size=self.window_size,
                           kernel_size=self.kernel_size, pooling=self.pooling, capacity=self.max_capacity_prompt,
                           tsp_len=0, mode=mode, impl=self.score_impl, return_aux=self.keep_aux, overlap=self.overlap)
        if self.keep_aux:
            self.last_aux = res[3]
        return res[0], res[1]


# ---- GemFilter selection (baselines/gemfilter/utils.py:25-56) ---------------------------------
def standard_dis_index(data, queries, k, norm=1, pool=False, kernel_size=5, sum_over_heads=False):
    """Signature of gemfilter/utils.py:25.  Only the combination the reference actually uses
    (``pool=True, sum_over_heads=True``, un-expanded or expanded K) is implemented in CUDA."""
    if not (pool and sum_over_heads):
        raise NotImplementedError("fastkv_b200 implements the pool=True, sum_over_heads=True selection only")
    L.require_cuda(data, queries)
    q_last = queries[:, :, -1:, :]
    n = data.shape[2]
    # the kernel reads query row L-1 of a [B,Hq,L,D] view: present the single row as such a view
    qv = _RowOffsetView(q_last.contiguous(), n)
    scores = ops.score(qv, data, window_size=1, kernel_size=kernel_size, mode=L.MODE_GEMFILTER)   # [B,1,n]
    idx = ops.topk(scores.view(-1, n), k, order=L.ORDER_VALUE_DESC).view(scores.shape[0], 1, k)
    distances = torch.gather(scores, -1, idx)
    if norm != 1:
        distances = distances / norm
    return distances, idx


def find_context(self, query_states, key_states, print_idx_dis=False):
    """gemfilter/utils.py:41-56: records ``self.indecies`` on the selection layer (batch 1 only).

    ``key_states`` may be the un-expanded [1,Hkv,n,D] tensor (no repeat_kv needed)."""
    b, h, n, d = key_states.shape
    if self.eviction_mode == "proportional":
        self.topk = int(n * self.retain_rate)
    if self.indecies is None and self.layer_idx == self.select_layer_idx:
        assert b == 1
        _, indices = standard_dis_index(key_states, query_states, min(self.topk, n), pool=True, sum_over_heads=True)
        self.indecies = indices
    return
