"""Pre-allocated compact KV cache (replaces HF ``DynamicCache`` on the FastKV path).

The reference stores the compacted K/V in a ``DynamicCache`` whose ``update`` re-allocates and
``torch.cat``s the whole layer cache on every decode step (baselines/fastkv/llama_model.py:142-145).
Here every layer owns one ``[B, Hkv, capacity, D]`` buffer sized ``compact rows + max_new_tokens``:
the prefill writes the compacted rows once, decode steps append one row in place
(``fkv_cache_append``) and the decode-attention kernel reads rows ``[0, seq_len)``.

``update()`` keeps HF's ``Cache.update`` calling convention so the patched attention forwards read
like the reference's; ``get_seq_length()`` returns the layer-0 cache length exactly like
``DynamicCache`` did for the reference's benchmark loop (benchmark/e2e.py:174 passes no position
ids, so decode positions continue from the COMPACT length there).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import ops


class CompactKVCache:
    """HF-Cache-compatible (duck-typed) container of per-layer compact K/V buffers."""

    is_compileable = False

    def __init__(self, config=None, max_new_tokens: int = 256):
        self.max_new_tokens = int(max_new_tokens)
        self.k: List[Optional[torch.Tensor]] = []
        self.v: List[Optional[torch.Tensor]] = []
        self.seq_lens: List[int] = []                 # host mirror of the per-layer length
        self.seq_lens_dev: List[Optional[torch.Tensor]] = []   # int32 [1] on device (CUDA-graph friendly)
        self.true_length = 0                           # tokens the model has seen (prompt + generated), NOT the compact
                                                       # row count: the generate() path continues positions from here
                                                       # (baselines/monkeypatch.py:280-288)
        self.device_lengths = False                    # True: decode steps advance the device-side lengths only

    # -- HF Cache surface -----------------------------------------------------------------------
    def __len__(self):
        return len(self.k)

    def get_seq_length(self, layer_idx: int = 0) -> int:
        if layer_idx >= len(self.seq_lens):
            return 0
        return self.seq_lens[layer_idx]

    def get_mask_sizes(self, cache_position=None, layer_idx: int = 0):   # transformers >= 4.5x mask helpers
        q = int(cache_position.shape[0]) if cache_position is not None else 1
        return self.get_seq_length(layer_idx) + q, 0

    def note_tokens(self, n: int):
        """Called once per model forward with the number of new tokens (prompt length, then 1 per decode step)."""
        self.true_length += int(n)

    def _grow_lists(self, layer_idx: int):
        while len(self.k) <= layer_idx:
            self.k.append(None); self.v.append(None); self.seq_lens.append(0); self.seq_lens_dev.append(None)

    def reserve(self, layer_idx: int, batch: int, kv_heads: int, rows: int, head_dim: int, dtype, device):
        """Allocate (once) the layer buffer with room for ``rows`` compact rows + max_new_tokens."""
        self._grow_lists(layer_idx)
        cap = rows + self.max_new_tokens
        cur = self.k[layer_idx]
        if cur is None or cur.shape[0] != batch or cur.shape[1] != kv_heads or cur.shape[2] < cap or cur.dtype != dtype:
            self.k[layer_idx] = torch.empty((batch, kv_heads, cap, head_dim), dtype=dtype, device=device)
            self.v[layer_idx] = torch.empty((batch, kv_heads, cap, head_dim), dtype=dtype, device=device)
            self.seq_lens_dev[layer_idx] = torch.zeros(1, dtype=torch.int32, device=device)
        return self.k[layer_idx], self.v[layer_idx]

    def set_length(self, layer_idx: int, n: int):
        self.seq_lens[layer_idx] = int(n)
        self.seq_lens_dev[layer_idx].fill_(int(n))

    def update(self, key_states: torch.Tensor, value_states: torch.Tensor, layer_idx: int, cache_kwargs=None
               ) -> Tuple[torch.Tensor, torch.Tensor]:
        """Prefill (rows > 1): store the given (already compacted) rows.  Decode (1 row): append in place.

        Returns views of the valid rows, like ``DynamicCache.update``."""
        b, h, rows, d = key_states.shape
        self._grow_lists(layer_idx)
        if rows > 1 or self.k[layer_idx] is None:
            kbuf, vbuf = self.reserve(layer_idx, b, h, rows, d, key_states.dtype, key_states.device)
            if key_states.data_ptr() != kbuf.data_ptr():      # the compress kernel may already have written in place
                kbuf[:, :, :rows].copy_(key_states)
                vbuf[:, :, :rows].copy_(value_states)
            self.set_length(layer_idx, rows)
        else:
            kbuf, vbuf = self.k[layer_idx], self.v[layer_idx]
            if h != kbuf.shape[1] or b != kbuf.shape[0] or d != kbuf.shape[3]:
                raise ValueError(f"CompactKVCache.update: new row has shape {tuple(key_states.shape)} but layer {layer_idx} caches "
                                 f"{tuple(kbuf.shape)} (SnapKV caches hold one head per QUERY head: expand K/V first)")
            n = self.seq_lens[layer_idx]
            if n >= kbuf.shape[2]:
                raise RuntimeError("CompactKVCache is full: raise max_new_tokens")
            if key_states.is_cuda:
                ops.cache_append(key_states, value_states, kbuf, vbuf, n)
            else:                                             # CPU plumbing tests only (no kernels involved)
                kbuf[:, :, n] = key_states[:, :, 0]
                vbuf[:, :, n] = value_states[:, :, 0]
            self.seq_lens[layer_idx] = n + 1
            self.seq_lens_dev[layer_idx].add_(1)
        n = self.seq_lens[layer_idx]
        return self.k[layer_idx][:, :, :n], self.v[layer_idx][:, :, :n]

    def sync_lengths(self):
        """Refresh the host mirrors from the device-side lengths (after graph-replayed decode steps)."""
        for i, d in enumerate(self.seq_lens_dev):
            if d is not None:
                self.seq_lens[i] = int(d.item())

    def buffers(self, layer_idx: int):
        """(k_buffer, v_buffer, seq_len, seq_len_dev) for the decode-attention kernel."""
        return self.k[layer_idx], self.v[layer_idx], self.seq_lens[layer_idx], self.seq_lens_dev[layer_idx]
