"""CPU oracle for FastKV's TSP + KV-compression hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fastkv_b200/`` may import this module: it is the
checker that ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use, never the thing that is shipped or measured as the product.

It restates, step by step and exposing every intermediate, what the reference computes in

* ``baselines/fastkv/utils.py:80-134``     (``FastKVCluster.update_kv``)
* ``baselines/snapkv/utils.py:57-102``     (``SnapKVCluster.update_kv``)
* ``baselines/gemfilter/utils.py:25-56``   (``standard_dis_index`` / ``find_context``)
* ``baselines/fastkv/llama_model.py:252-259`` (TSP hidden-state / position-id gather)
* flash-attn 2.6.3 ``flash_attn_func`` at q_len == 1 (third-party, not vendored in the reference:
  restated here as fp32 ``softmax(q K^T / sqrt(D)) V``; parity for that row is *unpinned* by the
  reference because it has no test or golden vector touching it).

Parity status: **pinned**.  ``tests/golden/make_golden.py`` imports the unmodified reference
classes from ``/root/reference`` and stores their outputs; ``tests/test_oracle_golden.py`` checks
this restatement against those stored outputs bit-for-bit (the restatement issues the same ATen
ops in the same order, so on the same torch build the results are identical, ties included).

All functions take torch CPU tensors (the same code also runs on CUDA tensors: bench.py times it there as the
"reference torch path on this GPU" baseline).  ``compute_dtype`` mirrors the reference's behaviour of
working in the dtype of the incoming tensors (fp32 under the parity protocol of SURVEY.md §8c).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def expand_kv_heads(x: torch.Tensor, groups: int) -> torch.Tensor:
    """[B,Hkv,L,D] -> [B,Hkv*groups,L,D]; query head hq reads kv head hq // groups.

    Follows baselines/fastkv/utils.py:13-22 (``repeat_kv``)."""
    if groups == 1:
        return x
    b, h, l, d = x.shape
    return x.unsqueeze(2).expand(b, h, groups, l, d).reshape(b, h * groups, l, d)


def window_causal_bias(window: int, dtype: torch.dtype, device=None) -> torch.Tensor:
    """[W,W] additive bias: 0 where col <= row, finfo(dtype).min elsewhere.

    Follows baselines/fastkv/utils.py:95-99."""
    bias = torch.full((window, window), torch.finfo(dtype).min, dtype=dtype, device=device)
    cols = torch.arange(window, device=device)
    bias.masked_fill_(cols[None, :] <= cols[:, None], 0)
    return bias


def pool1d(x: torch.Tensor, kernel_size: int, pooling: str) -> torch.Tensor:
    """Same-length pooling along the last dim.  baselines/fastkv/utils.py:105-110.

    avg: zero padding, always divides by kernel_size (count_include_pad); max: -inf padding."""
    if pooling == "avgpool":
        return F.avg_pool1d(x, kernel_size=kernel_size, padding=kernel_size // 2, stride=1)
    if pooling == "maxpool":
        return F.max_pool1d(x, kernel_size=kernel_size, padding=kernel_size // 2, stride=1)
    raise ValueError("Pooling method not supported")


def stable_topk_desc(x: torch.Tensor, k: int) -> torch.Tensor:
    """Top-k indices along the last dim, ordered by (value descending, index ascending).

    This is the deterministic total order the CUDA kernels implement; the reference's
    ``Tensor.topk`` (baselines/fastkv/utils.py:113) leaves the order of equal values unspecified."""
    order = torch.sort(x, dim=-1, descending=True, stable=True).indices
    return order[..., :k].contiguous()


# --------------------------------------------------------------------------------------------
# FastKV / SnapKV
# --------------------------------------------------------------------------------------------
@dataclass
class CompressTrace:
    """Every intermediate of one ``update_kv`` call (None where the reference skips the step)."""
    skipped: bool = False                       # q_len < max_capacity_prompt early return
    capacity: int = 0                           # C actually used (after proportional rewrite)
    tsp_length: int = 0                         # T actually used
    votes: Optional[torch.Tensor] = None        # [B,Hq,P]   window-summed softmax mass per key
    pooled: Optional[torch.Tensor] = None       # [B,Hq,P]
    score: Optional[torch.Tensor] = None        # [B,Hsel,P] (Hsel = Hkv for FastKV, Hq for SnapKV)
    idx: Optional[torch.Tensor] = None          # [B,Hsel,C-W] int64, score-descending
    key: Optional[torch.Tensor] = None          # [B,Hsel,C,D]
    value: Optional[torch.Tensor] = None
    importance: Optional[torch.Tensor] = None   # [B,P]  sum over heads (TSP layer only)
    tsp_idx: Optional[torch.Tensor] = None      # [B,T] int64 ascending, or None
    extra: dict = field(default_factory=dict)


def window_votes(q: torch.Tensor, k_expanded: torch.Tensor, window: int) -> torch.Tensor:
    """Observation-window scoring.  baselines/fastkv/utils.py:94-104.

    q [B,Hq,L,D], k_expanded [B,Hq,L,D]  ->  votes [B,Hq,L-W] in q.dtype.
    The softmax normaliser runs over all L keys (window keys included, causally masked inside the
    window); the returned votes drop the last W key columns."""
    d = q.shape[-1]
    logits = torch.matmul(q[..., -window:, :], k_expanded.transpose(2, 3)) / math.sqrt(d)
    logits[:, :, -window:, -window:] += window_causal_bias(window, logits.dtype, logits.device)[None, None]
    probs = F.softmax(logits, dim=-1, dtype=torch.float32).to(q.dtype)
    return probs[:, :, -window:, :-window].sum(dim=-2)


def fastkv_update_kv(
    key_states: torch.Tensor,
    query_states: torch.Tensor,
    value_states: torch.Tensor,
    num_key_value_groups: int,
    *,
    window_size: int = 8,
    max_capacity_prompt: int = 512,
    kernel_size: int = 7,
    pooling: str = "avgpool",
    tsp_layer: bool = False,
    tsp_length: int = 2048,
    tsp_rate: float = 0.25,
    retain_rate: float = 0.25,
    eviction_mode: str = "constant",
    deterministic_ties: bool = False,
) -> CompressTrace:
    """Restatement of ``FastKVCluster.update_kv`` (baselines/fastkv/utils.py:80-134).

    ``deterministic_ties=False`` uses ``Tensor.topk`` exactly like the reference;
    ``True`` uses the (score desc, index asc) total order of the CUDA kernels."""
    assert key_states.shape[-2] == query_states.shape[-2]                    # :82
    bsz, _, q_len, head_dim = query_states.shape
    tr = CompressTrace()
    cap = max_capacity_prompt
    if eviction_mode == "proportional":                                      # :86-87
        cap = int(q_len * retain_rate)
    tr.capacity = cap
    tr.tsp_length = tsp_length
    if q_len < cap:                                                          # :89-91 (strict <)
        tr.skipped = True
        tr.key, tr.value = key_states, value_states
        return tr

    w = window_size
    k_exp = expand_kv_heads(key_states, num_key_value_groups)                # :93
    tr.votes = window_votes(query_states, k_exp, w)                          # :94-104
    tr.pooled = pool1d(tr.votes, kernel_size, pooling)                       # :105-110
    tr.score = tr.pooled.view(bsz, -1, num_key_value_groups, q_len - w).sum(dim=-2)   # :112
    select = stable_topk_desc if deterministic_ties else (lambda x, k: x.topk(k, dim=-1).indices)
    tr.idx = select(tr.score, cap - w)                                       # :113
    gidx = tr.idx.unsqueeze(-1).expand(-1, -1, -1, head_dim)                 # :114
    tr.key = torch.cat([key_states[:, :, :-w, :].gather(2, gidx), key_states[:, :, -w:, :]], dim=2)      # :116-120
    tr.value = torch.cat([value_states[:, :, :-w, :].gather(2, gidx), value_states[:, :, -w:, :]], dim=2)  # :117-121

    t_len = tsp_length
    if tsp_layer and eviction_mode == "proportional":                        # :123-124
        t_len = int(q_len * tsp_rate)
    tr.tsp_length = t_len
    if tsp_layer and q_len > t_len:                                          # :126 (strict >)
        tr.importance = tr.score.sum(dim=-2)                                 # :127
        picked = select(tr.importance, t_len - w)
        tail = torch.arange(q_len - w, q_len, device=key_states.device).unsqueeze(0).repeat(bsz, 1)    # :128
        tr.tsp_idx = torch.sort(torch.cat([picked, tail], dim=-1), dim=1).values   # :129-130
    return tr


def snapkv_update_kv(
    key_states: torch.Tensor,
    query_states: torch.Tensor,
    value_states: torch.Tensor,
    *,
    window_size: int = 64,
    max_capacity_prompt: int = 320,
    kernel_size: int = 5,
    pooling: str = "avgpool",
    retain_rate: float = 0.25,
    eviction_mode: str = "constant",
    deterministic_ties: bool = False,
) -> CompressTrace:
    """Restatement of ``SnapKVCluster.update_kv`` (baselines/snapkv/utils.py:57-102).

    The caller has already expanded K/V to one head per query head
    (baselines/snapkv/llama_model.py:385-386), so all tensors are [B,Hq,L,D]; there is no group
    sum and no TSP output.  ``merge`` must stay None (``merge_kv`` is undefined upstream)."""
    assert key_states.shape[-2] == query_states.shape[-2]
    _, _, q_len, head_dim = query_states.shape
    tr = CompressTrace()
    cap = max_capacity_prompt
    if eviction_mode == "proportional":
        cap = int(q_len * retain_rate)
    tr.capacity = cap
    if q_len < cap:
        tr.skipped = True
        tr.key, tr.value = key_states, value_states
        return tr
    w = window_size
    tr.votes = window_votes(query_states, key_states, w)
    tr.pooled = pool1d(tr.votes, kernel_size, pooling)
    tr.score = tr.pooled
    select = stable_topk_desc if deterministic_ties else (lambda x, k: x.topk(k, dim=-1).indices)
    tr.idx = select(tr.score, cap - w)
    gidx = tr.idx.unsqueeze(-1).expand(-1, -1, -1, head_dim)
    tr.key = torch.cat([key_states[:, :, :-w, :].gather(2, gidx), key_states[:, :, -w:, :]], dim=2)
    tr.value = torch.cat([value_states[:, :, :-w, :].gather(2, gidx), value_states[:, :, -w:, :]], dim=2)
    return tr


# --------------------------------------------------------------------------------------------
# GemFilter selection
# --------------------------------------------------------------------------------------------
def gemfilter_select(
    query_states: torch.Tensor,
    key_states: torch.Tensor,
    num_key_value_groups: int,
    topk: int,
    *,
    kernel_size: int = 5,
    retain_rate: Optional[float] = None,
    eviction_mode: str = "constant",
    deterministic_ties: bool = False,
):
    """Restatement of ``find_context`` -> ``standard_dis_index``
    (baselines/gemfilter/utils.py:41-56 and :25-39).

    Last query row only, raw dot products (no 1/sqrt(D), no softmax, no mask), sum over all query
    heads, avg_pool1d(k=5), top-min(topk, n) over all n keys.  Returns (scores [1,1,n],
    indices [1,1,k] in descending-score order, sorted_indices [k] ascending as used by
    ``get_layer_context`` :58-61)."""
    b, _, n, _ = key_states.shape
    assert b == 1                                                            # :47
    if eviction_mode == "proportional":                                      # :44-45
        topk = int(n * retain_rate)
    k_exp = expand_kv_heads(key_states, num_key_value_groups)                # :48-49
    ip = torch.matmul(query_states[:, :, -1:, :], k_exp.transpose(-1, -2))[:, :, 0, :]   # :26-27
    ip = torch.sum(ip, dim=1, keepdim=True)                                  # :29
    ip = F.avg_pool1d(ip, kernel_size=kernel_size, padding=kernel_size // 2, stride=1)    # :31-32
    k = min(topk, n)
    idx = stable_topk_desc(ip, k) if deterministic_ties else torch.topk(ip, k, dim=-1).indices
    return ip, idx, torch.sort(idx[0, 0, :]).values


# --------------------------------------------------------------------------------------------
# model-glue pieces on the path
# --------------------------------------------------------------------------------------------
def tsp_propagate(hidden_states: torch.Tensor, position_ids: torch.Tensor, tsp_idx: torch.Tensor):
    """Hidden-state and position-id gather at the TSP layer.

    baselines/fastkv/llama_model.py:252-259 (Mistral twin mistral_model.py:217-224)."""
    new_pos = torch.gather(position_ids, 1, tsp_idx)
    new_hidden = torch.gather(hidden_states, 1, tsp_idx.unsqueeze(-1).expand(-1, -1, hidden_states.size(2)))
    return new_hidden, new_pos


def decode_attention(q: torch.Tensor, k_cache: torch.Tensor, v_cache: torch.Tensor, seq_len: int) -> torch.Tensor:
    """q_len == 1 attention over the compact cache, fp32 math.

    Call sites: baselines/fastkv/llama_model.py:79-81 via :181-183 (flash_attn_func, causal —
    trivially all-visible at q_len 1).  q [B,Hq,D]; caches [B,Hkv,S_max,D]; uses rows [0, seq_len)."""
    b, hq, d = q.shape
    hkv = k_cache.shape[1]
    g = hq // hkv
    k = expand_kv_heads(k_cache[:, :, :seq_len].float(), g)
    v = expand_kv_heads(v_cache[:, :, :seq_len].float(), g)
    s = torch.einsum("bhd,bhsd->bhs", q.float(), k) / math.sqrt(d)
    p = torch.softmax(s, dim=-1)
    return torch.einsum("bhs,bhsd->bhd", p, v)


# --------------------------------------------------------------------------------------------
# comparison rules (SURVEY.md §8c parity protocol)
# --------------------------------------------------------------------------------------------
def near_tie_set_mismatch(ref_scores: torch.Tensor, ref_idx: torch.Tensor, got_idx: torch.Tensor,
                          rtol: float = 1e-5) -> int:
    """Count selections that differ from the reference for a reason other than a near-tie.

    ref_scores [n] fp32 (the reference's own scores), ref_idx / got_idx 1-D index sets of equal
    size.  An element present in one set but not the other is tolerated iff its reference score
    is within ``rtol`` (relative) of the reference's k-th (boundary) score — i.e. the two
    implementations only disagree on elements the reference itself cannot separate by more than
    rtol.  Returns the number of non-tolerated differing elements (0 == parity)."""
    ref_set = set(ref_idx.tolist())
    got_set = set(got_idx.tolist())
    assert len(ref_set) == ref_idx.numel(), "reference indices not unique"
    if len(got_set) != got_idx.numel():
        return got_idx.numel() - len(got_set) + len(ref_set ^ got_set)
    if ref_set == got_set:
        return 0
    boundary = ref_scores[ref_idx].min().item()
    tol = rtol * max(abs(boundary), torch.finfo(torch.float32).tiny)
    bad = 0
    for j in ref_set ^ got_set:
        if abs(ref_scores[j].item() - boundary) > tol:
            bad += 1
    return bad


def order_is_descending(scores: torch.Tensor, idx: torch.Tensor, rtol: float = 1e-5) -> bool:
    """True iff scores[idx] is non-increasing up to rtol (cache layout is score-descending)."""
    v = scores[idx]
    if v.numel() < 2:
        return True
    drop = v[1:] - v[:-1]
    return bool((drop <= rtol * v[:-1].abs().clamp_min(torch.finfo(torch.float32).tiny)).all())
