"""CPU restatement of the reference's FastKV model glue for a HF Llama/Mistral-style model.

TEST INFRASTRUCTURE ONLY (see oracle/fastkv_oracle.py).  The reference's own glue
(baselines/fastkv/llama_model.py, mistral_model.py) cannot be imported under the installed
transformers (it needs 4.45 symbols), so this file restates it on top of the HF sub-modules:

* attention, prefill: compress with update_kv, cache the COMPACT K/V, attend over the FULL K/V
  (llama_model.py:136-142, :181-183)
* decoder layer: after the MLP residual of the TSP layer gather hidden states and position ids
  (llama_model.py:252-259)
* model loop: swap position ids, recompute rotary embeddings, cut to the last token
  (llama_model.py:368-371, :392)
Parity for these rows is unpinned by reference outputs (the reference glue cannot run here); the
compressor inside is the pinned oracle.
"""
import torch
import torch.nn.functional as F

from . import fastkv_oracle as O


@torch.no_grad()
def fastkv_prefill(model, input_ids, knobs_per_layer, apply_rotary_pos_emb):
    """Returns (last_hidden [B,1,H], per-layer list of (k, v, tsp_idx | None), per-layer traces)."""
    m = model.model
    hidden = m.embed_tokens(input_ids)
    b, l, _ = hidden.shape
    position_ids = torch.arange(l).unsqueeze(0)
    pos_emb = m.rotary_emb(hidden, position_ids=position_ids)
    caches, traces = [], []
    for i, layer in enumerate(m.layers):
        attn = layer.self_attn
        residual = hidden
        x = layer.input_layernorm(hidden)
        shape = (*x.shape[:-1], -1, attn.head_dim)
        q = attn.q_proj(x).view(shape).transpose(1, 2)
        k = attn.k_proj(x).view(shape).transpose(1, 2)
        v = attn.v_proj(x).view(shape).transpose(1, 2)
        q, k = apply_rotary_pos_emb(q, k, *pos_emb)
        tr = O.fastkv_update_kv(k, q, v, attn.num_key_value_groups, deterministic_ties=True, **knobs_per_layer[i])
        caches.append((tr.key, tr.value, tr.tsp_idx))
        traces.append(tr)
        g = attn.num_key_value_groups
        o = F.scaled_dot_product_attention(q, O.expand_kv_heads(k, g), O.expand_kv_heads(v, g), is_causal=True,
                                           scale=attn.scaling)
        o = o.transpose(1, 2).reshape(b, q.shape[2], -1)
        hidden = residual + attn.o_proj(o)
        residual = hidden
        hidden = residual + layer.mlp(layer.post_attention_layernorm(hidden))
        if knobs_per_layer[i].get("tsp_layer") and tr.tsp_idx is not None:
            hidden, position_ids = O.tsp_propagate(hidden, position_ids.expand(b, -1), tr.tsp_idx)
            pos_emb = m.rotary_emb(hidden, position_ids=position_ids)
    hidden = m.norm(hidden)
    return hidden[:, -1:, :], caches, traces
