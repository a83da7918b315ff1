"""CPU restatement of the reference's FastKV / SnapKV model glue for a HF Llama/Mistral-style model.

TEST INFRASTRUCTURE ONLY (see oracle/fastkv_oracle.py).  The reference's own glue
(baselines/fastkv/llama_model.py, mistral_model.py, snapkv/llama_model.py, monkeypatch.py) cannot be imported
under the installed transformers (it needs 4.45 symbols), so this file restates it on top of the HF sub-modules:

* attention, prefill: compress with update_kv, cache the COMPACT K/V, attend over the FULL K/V
  (llama_model.py:136-142, :181-183; Mistral mistral_model.py:100-107 with flash-attn's index-based
  ``sliding_window`` :143-153)
* decoder layer: after the MLP residual of the TSP layer gather hidden states and position ids
  (llama_model.py:252-259; mistral_model.py:217-224)
* model loop: swap position ids, recompute rotary embeddings, cut to the last token
  (llama_model.py:368-371, :392; mistral_model.py:328-330, :351)
* decode: append the new K/V row to the compact cache and attend over it (llama_model.py:143-145); positions of the
  generated tokens continue from the TRUE prompt length when driven by ``generate`` (monkeypatch.py:280-288) or from the
  layer-0 cache length in the benchmark loop (benchmark/e2e.py:174)
* SnapKV: ``repeat_kv`` BEFORE compression and before every cache update, so the cache holds one head per query head
  (snapkv/llama_model.py:385-404)
Parity for these rows is unpinned by reference outputs (the reference glue cannot run here); the
compressor inside is the pinned oracle.
"""
import torch
import torch.nn.functional as F

from . import fastkv_oracle as O


def _causal_band_mask(n, sliding_window):
    """Additive-free boolean mask [n,n]: key j visible to query i iff j <= i and (no window or i - j < window).
    flash-attn's ``window_size=(w, 0)`` semantics as the reference's call uses them (index-based)."""
    i = torch.arange(n)
    keep = i[None, :] <= i[:, None]
    if sliding_window is not None and n > sliding_window:
        keep &= (i[:, None] - i[None, :]) < sliding_window
    return keep


def _attend(attn, q, k, v, sliding_window=None):
    g = attn.num_key_value_groups
    n = q.shape[2]
    ke, ve = O.expand_kv_heads(k, q.shape[1] // k.shape[1]), O.expand_kv_heads(v, q.shape[1] // v.shape[1])
    if n > 1 and sliding_window is not None and n > sliding_window:
        return F.scaled_dot_product_attention(q, ke, ve, attn_mask=_causal_band_mask(n, sliding_window)[None, None], scale=attn.scaling)
    return F.scaled_dot_product_attention(q, ke, ve, is_causal=(n > 1), scale=attn.scaling)


def _project(layer, hidden, pos_emb, apply_rotary_pos_emb):
    attn = layer.self_attn
    x = layer.input_layernorm(hidden)
    shape = (*x.shape[:-1], -1, attn.head_dim)
    q = attn.q_proj(x).view(shape).transpose(1, 2)
    k = attn.k_proj(x).view(shape).transpose(1, 2)
    v = attn.v_proj(x).view(shape).transpose(1, 2)
    q, k = apply_rotary_pos_emb(q, k, *pos_emb)
    return q, k, v


def _finish_layer(layer, residual, o, b):
    attn = layer.self_attn
    o = o.transpose(1, 2).reshape(b, o.shape[2], -1)
    hidden = residual + attn.o_proj(o)
    return hidden + layer.mlp(layer.post_attention_layernorm(hidden))


@torch.no_grad()
def fastkv_prefill(model, input_ids, knobs_per_layer, apply_rotary_pos_emb, sliding_window=None):
    """Returns (last_hidden [B,1,H], per-layer list of (k, v, tsp_idx | None), per-layer traces)."""
    m = model.model
    hidden = m.embed_tokens(input_ids)
    b, l, _ = hidden.shape
    position_ids = torch.arange(l).unsqueeze(0)
    pos_emb = m.rotary_emb(hidden, position_ids=position_ids)
    caches, traces = [], []
    for i, layer in enumerate(m.layers):
        attn = layer.self_attn
        q, k, v = _project(layer, hidden, pos_emb, apply_rotary_pos_emb)
        tr = O.fastkv_update_kv(k, q, v, attn.num_key_value_groups, deterministic_ties=True, **knobs_per_layer[i])
        caches.append((tr.key, tr.value, tr.tsp_idx))
        traces.append(tr)
        hidden = _finish_layer(layer, hidden, _attend(attn, q, k, v, sliding_window), b)
        if knobs_per_layer[i].get("tsp_layer") and tr.tsp_idx is not None:
            hidden, position_ids = O.tsp_propagate(hidden, position_ids.expand(b, -1), tr.tsp_idx)
            pos_emb = m.rotary_emb(hidden, position_ids=position_ids)
    hidden = m.norm(hidden)
    return hidden[:, -1:, :], caches, traces


@torch.no_grad()
def snapkv_prefill(model, input_ids, knobs_per_layer, apply_rotary_pos_emb):
    """SnapKV twin (snapkv/llama_model.py:385-404): K/V expanded to one head per query head, compressed per query
    head, no TSP.  Returns (last_hidden, per-layer (k [B,Hq,C,D], v), traces)."""
    m = model.model
    hidden = m.embed_tokens(input_ids)
    b, l, _ = hidden.shape
    pos_emb = m.rotary_emb(hidden, position_ids=torch.arange(l).unsqueeze(0))
    caches, traces = [], []
    for i, layer in enumerate(m.layers):
        attn = layer.self_attn
        q, k, v = _project(layer, hidden, pos_emb, apply_rotary_pos_emb)
        ke, ve = O.expand_kv_heads(k, attn.num_key_value_groups), O.expand_kv_heads(v, attn.num_key_value_groups)   # :385-386
        tr = O.snapkv_update_kv(ke, q, ve, deterministic_ties=True, **knobs_per_layer[i])                       # :401-402
        caches.append((tr.key, tr.value))
        traces.append(tr)
        hidden = _finish_layer(layer, hidden, _attend(attn, q, ke, ve), b)
    hidden = m.norm(hidden)
    return hidden[:, -1:, :], caches, traces


@torch.no_grad()
def decode_steps(model, caches, first_token, start_position, n_steps, apply_rotary_pos_emb, expand_new_kv=False):
    """Greedy decode over the compact caches: returns (tokens [B, n_steps], positions used [n_steps]).

    ``caches``: per layer [k, v] (compact, possibly of different lengths per layer).  Every step appends the new
    K/V row (``expand_new_kv``: repeat_kv first, the SnapKV cache layout) and attends over all cached rows
    (llama_model.py:143-145).  ``start_position`` is the position id of ``first_token``: the true prompt length on the
    ``generate`` path (monkeypatch.py:280-288), the layer-0 cache length in the benchmark loop (e2e.py:174)."""
    m = model.model
    caches = [[c[0].clone(), c[1].clone()] for c in caches]
    tok, toks, used = first_token, [], []
    for t in range(n_steps):
        pos = torch.full((tok.shape[0], 1), start_position + t, dtype=torch.long)
        used.append(start_position + t)
        hidden = m.embed_tokens(tok)
        pos_emb = m.rotary_emb(hidden, position_ids=pos)
        b = hidden.shape[0]
        for i, layer in enumerate(m.layers):
            attn = layer.self_attn
            q, k, v = _project(layer, hidden, pos_emb, apply_rotary_pos_emb)
            if expand_new_kv:
                k, v = O.expand_kv_heads(k, attn.num_key_value_groups), O.expand_kv_heads(v, attn.num_key_value_groups)
            caches[i][0] = torch.cat([caches[i][0], k], dim=2)
            caches[i][1] = torch.cat([caches[i][1], v], dim=2)
            hidden = _finish_layer(layer, hidden, _attend(attn, q, caches[i][0], caches[i][1]), b)
        logits = model.lm_head(m.norm(hidden)[:, -1:, :])
        tok = logits[:, -1, :].argmax(dim=-1, keepdim=True)
        toks.append(tok)
    return torch.cat(toks, dim=1), used
