"""Monkeypatch entry points with the reference's names, re-hosted on the INSTALLED transformers.

Mirrors ``baselines/monkeypatch.py``: ``replace_llama(method)`` :12, ``replace_mistral(method)`` :59,
``set_model(model, args)`` :104, plus ``compress_fastkv`` / ``init_fastkv``
(``baselines/fastkv/utils.py:25,137``).  The reference patches transformers 4.45 symbols
(``LlamaFlashAttention2``, ``LLAMA_ATTENTION_CLASSES``) that no longer exist; here the three
``forward`` methods of the current ``*Attention`` / ``*DecoderLayer`` / ``*Model`` classes are
replaced and every attention module gets its ``kv_cluster``:

* attention forward  — prefill: ``kv_cluster.update_kv`` compacts K/V into the cache while the layer
  still attends over the FULL K/V; decode: append + attention over the compact cache
  (baselines/fastkv/llama_model.py:136-145, :181-183; mistral_model.py:100-107)
* decoder layer      — after the MLP residual of the TSP layer, propagate only the selected tokens'
  hidden states and position ids (llama_model.py:252-259; mistral_model.py:217-224)
* model forward      — swap in the new position ids / rotary embeddings after the TSP layer and
  return only the last token's hidden state (llama_model.py:368-371, :392; mistral_model.py:328-351)

Supported methods: ``fastkv``, ``snapkv`` (selection through the same kernels), ``fullkv``.
The other baselines of the reference (h2o, streamingllm, pyramidinfer, gemfilter's two-pass driver)
are outside the hot path and raise ``NotImplementedError``.
"""
from __future__ import annotations

import torch

from . import ops
from .cache import CompactKVCache
from .cluster import FastKVCluster, SnapKVCluster

_ORIGINALS = {}


def _repeat_kv(x, n_rep):
    """[B,Hkv,S,D] -> [B,Hkv*n_rep,S,D] (query head hq reads kv head hq // n_rep), baselines/snapkv/llama_model.py:385."""
    if n_rep == 1:
        return x
    b, h, s, d = x.shape
    return x[:, :, None, :, :].expand(b, h, n_rep, s, d).reshape(b, h * n_rep, s, d)


# ---------------------------------------------------------------------------------------------
# knob plumbing (same attribute / argument names as the reference)
# ---------------------------------------------------------------------------------------------
def init_fastkv(self):
    """baselines/fastkv/utils.py:137-138."""
    self.kv_cluster = FastKVCluster()


def init_snapkv(self):
    """baselines/snapkv/utils.py:105-126 (defaults written onto self.config when absent)."""
    cfg = self.config
    for name, default in (("window_size", 32), ("max_capacity_prompt", 4096), ("kernel_size", 5),
                          ("pooling", "avgpool"), ("merge", None), ("retain_rate", 0.25), ("eviction_mode", "constant")):
        if not hasattr(cfg, name):
            setattr(cfg, name, default)
    self.kv_cluster = SnapKVCluster(window_size=cfg.window_size, max_capacity_prompt=cfg.max_capacity_prompt,
                                    kernel_size=cfg.kernel_size, pooling=cfg.pooling, merge=cfg.merge,
                                    retain_rate=cfg.retain_rate, eviction_mode=cfg.eviction_mode)


def compress_fastkv(model, args):
    """baselines/fastkv/utils.py:25-46: write the knobs onto every layer's cluster."""
    layers = len(model.model.layers)
    for i in range(layers):
        cl = model.model.layers[i].self_attn.kv_cluster
        cl.window_size = args.window_size[i]
        cl.kernel_size = args.kernel_size[i]
        cl.pooling = args.pooling
        cl.max_capacity_prompt = args.max_capacity_prompts
        cl.tsp_length = args.tsp_len
        cl.tsp_rate = args.tsp_rate
        cl.eviction_mode = args.eviction_mode
        cl.tsp_layer = (i == args.tsp_idx)
        if i <= args.tsp_idx:
            cl.retain_rate = args.retain_rate
        else:
            cl.retain_rate = args.retain_rate / args.tsp_rate


def set_model(model, args):
    """baselines/monkeypatch.py:104-147 (fastkv / snapkv / fullkv branches)."""
    if args.max_capacity_prompts != -1:
        max_capacity_prompts = args.max_capacity_prompts
    else:
        # the reference leaves this unbound (monkeypatch.py:105-106); fail loudly instead
        raise ValueError("max_capacity_prompts == -1 is not supported")
    if args.method == "fullkv":
        return
    if args.method == "gemfilter":                             # monkeypatch.py:149-151
        set_topk(model, args, mode="gemfilter")
        return
    if args.method not in ("fastkv", "snapkv"):
        raise NotImplementedError(f"method {args.method!r} is outside the FastKV hot path")
    window_size, kernel_size, pooling, retain_rate = args.window_size, args.kernel_size, args.pooling, args.retain_rate
    layers = len(model.model.layers)
    if not isinstance(window_size, list):
        window_size = [window_size] * layers
    if not isinstance(max_capacity_prompts, list):
        max_capacity_prompts = [max_capacity_prompts] * layers
    if not isinstance(kernel_size, list):
        kernel_size = [kernel_size] * layers
    if not isinstance(retain_rate, list):
        retain_rate = [retain_rate] * layers
    for i in range(layers):
        cfg = model.model.layers[i].self_attn.config
        cfg.window_size = window_size[i]
        cfg.max_capacity_prompt = max_capacity_prompts[i]
        cfg.kernel_size = kernel_size[i]
        cfg.pooling = pooling
        cfg.merge = getattr(args, "merge", None)
        cfg.retain_rate = retain_rate[i]
        cfg.eviction_mode = args.eviction_mode
    if args.method == "fastkv":
        args.window_size = window_size
        args.kernel_size = kernel_size
        for layer in model.model.layers:
            if not isinstance(getattr(layer.self_attn, "kv_cluster", None), FastKVCluster):
                init_fastkv(layer.self_attn)
        compress_fastkv(model, args)
    else:  # snapkv: per-layer clusters read their knobs from the layer config (snapkv/utils.py:105-126)
        for i, layer in enumerate(model.model.layers):
            init_snapkv(layer.self_attn)
            layer.self_attn.kv_cluster.max_capacity_prompt = max_capacity_prompts[i]
            layer.self_attn.kv_cluster.window_size = window_size[i]
            layer.self_attn.kv_cluster.kernel_size = kernel_size[i]


def _prefill_mask(config, q_len, like, use_sliding_window):
    """(attention_mask, extra kwargs) for the attention interface of ``config._attn_implementation``.

    The reference always runs flash-attn with ``causal=True`` (+ ``sliding_window`` on Mistral,
    baselines/fastkv/mistral_model.py:143-153).  ``sdpa`` / ``flash_attention_*`` take the causal flag; ``eager``
    (and any interface that ignores ``is_causal``) gets an explicit additive mask, otherwise its prefill would be
    bidirectional.  A sliding window shorter than the prompt needs the band mask as well (sdpa) or the flash
    kernels' window argument."""
    impl = getattr(config, "_attn_implementation", None) or "eager"
    window = getattr(config, "sliding_window", None) if use_sliding_window else None
    if q_len == 1:
        return None, {"is_causal": False}
    banded = window is not None and q_len > window
    if impl.startswith("flash_attention") or (impl == "sdpa" and not banded):
        return None, {}                                # causal flag (+ sliding_window kwarg for the flash kernels)
    if impl not in ("sdpa", "eager"):
        raise NotImplementedError(f"fastkv_b200: attention implementation {impl!r} is not supported on the FastKV path "
                                  "(use sdpa, flash_attention_2 or eager)")
    i = torch.arange(q_len, device=like.device)
    keep = i[None, :] <= i[:, None]
    if banded:
        keep &= (i[:, None] - i[None, :]) < window
    mask = torch.zeros((q_len, q_len), dtype=like.dtype, device=like.device).masked_fill_(~keep, torch.finfo(like.dtype).min)
    return mask[None, None], {"is_causal": False}


# ---------------------------------------------------------------------------------------------
# patched forwards (shared by Llama and Mistral; `mod` is the transformers modeling module)
# ---------------------------------------------------------------------------------------------
def _make_forwards(mod, method, sliding_window_attr=False):
    apply_rotary_pos_emb = mod.apply_rotary_pos_emb
    eager_attention_forward = mod.eager_attention_forward
    ALL_ATTENTION_FUNCTIONS = mod.ALL_ATTENTION_FUNCTIONS
    BaseModelOutputWithPast = mod.BaseModelOutputWithPast

    def attn_forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None, **kwargs):
        input_shape = hidden_states.shape[:-1]
        q_len = hidden_states.shape[1]
        hidden_shape = (*input_shape, -1, self.head_dim)
        query_states = self.q_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        key_states = self.k_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        value_states = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        cos, sin = position_embeddings
        query_states, key_states = apply_rotary_pos_emb(query_states, key_states, cos, sin)

        self.tsp_idx = None                            # never let a previous call's selection leak into this one
        self.rows_kept = None                          # set when this call returns attention rows for the TSP-kept tokens only
        decode_native = False
        if past_key_values is not None:
            if q_len > 1:
                # prefill: compact K/V go to the cache, the layer itself still attends over everything
                if not hasattr(self, "kv_cluster"):
                    init_fastkv(self) if method == "fastkv" else init_snapkv(self)
                if method == "fastkv":
                    if isinstance(past_key_values, CompactKVCache):
                        # the compaction kernel writes straight into this layer's slot of the pre-allocated cache
                        cache, li = past_key_values, self.layer_idx
                        self.kv_cluster.out_provider = lambda b, h, rows, d, dt, dev: tuple(
                            buf[:, :, :rows] for buf in cache.reserve(li, b, h, rows, d, dt, dev))
                    k_c, v_c, self.tsp_idx = self.kv_cluster.update_kv(
                        key_states, query_states, value_states, attention_mask, self.num_key_value_groups, self.layer_idx)
                    self.kv_cluster.out_provider = None
                else:
                    k_c, v_c = self.kv_cluster.update_kv(
                        key_states, query_states, value_states, attention_mask, self.num_key_value_groups)
                    self.tsp_idx = None
                past_key_values.update(k_c, v_c, self.layer_idx)
            else:
                if method == "snapkv" and key_states.shape[1] != query_states.shape[1]:
                    # the SnapKV cache holds one head per QUERY head (snapkv/llama_model.py:385-386 repeat_kv's before
                    # the cache update in both phases): expand the new row the same way before appending it
                    key_states = _repeat_kv(key_states, self.num_key_value_groups)
                    value_states = _repeat_kv(value_states, self.num_key_value_groups)
                native_ok = (isinstance(past_key_values, CompactKVCache) and key_states.is_cuda
                             and query_states.shape[1] % key_states.shape[1] == 0 and self.head_dim in (64, 128))
                if native_ok and past_key_values.device_lengths:
                    # CUDA-graph friendly step: the row index and the attended length live on the device
                    kbuf, vbuf, _, ndev = past_key_values.buffers(self.layer_idx)
                    ops.cache_append(key_states, value_states, kbuf, vbuf, 0, seq_len_dev=ndev, bump=True)
                    out = ops.decode_attn(query_states[:, :, 0, :], kbuf, vbuf, 0, seq_len_dev=ndev, scale=self.scaling)
                    attn_output = self.o_proj(out.reshape(input_shape[0], 1, -1))
                    return attn_output, None
                key_states, value_states = past_key_values.update(key_states, value_states, self.layer_idx)
                decode_native = native_ok
        if decode_native:
            # K5: split-K decode attention straight over the pre-allocated compact cache
            kbuf, vbuf, n, _ = past_key_values.buffers(self.layer_idx)
            out = ops.decode_attn(query_states[:, :, 0, :], kbuf, vbuf, n, scale=self.scaling)     # [B,Hq,D]
            attn_output = out.reshape(input_shape[0], 1, -1)
            attn_weights = None
        else:
            attention_interface = ALL_ATTENTION_FUNCTIONS.get_interface(self.config._attn_implementation, eager_attention_forward)
            extra = {}
            if sliding_window_attr:
                extra["sliding_window"] = getattr(self.config, "sliding_window", None)
            impl = getattr(self.config, "_attn_implementation", None) or "eager"
            if (self.tsp_idx is not None and q_len > 1 and impl in ("sdpa", "eager")
                    and getattr(self.config, "fastkv_sparse_tsp_layer", True)):
                # only the T kept tokens leave the TSP layer (llama_model.py:252-259 gathers them after
                # the MLP), and attention / o_proj / MLP rows are independent of each other, so the layer computes just those
                # rows: T queries over all L keys under the same causal (and sliding-window) rule, in original positions
                rows = self.tsp_idx                                                # [B,T] ascending original positions
                b, hq, _, d = query_states.shape
                t_keep = rows.shape[1]
                q_sel = query_states.gather(2, rows[:, None, :, None].expand(b, hq, t_keep, d))
                j = torch.arange(q_len, device=rows.device)[None, None, :]
                keep = j <= rows[:, :, None]
                window = extra.get("sliding_window")
                if window is not None and q_len > window:
                    keep &= (rows[:, :, None] - j) < window
                mask = torch.zeros((b, 1, t_keep, q_len), dtype=q_sel.dtype, device=q_sel.device)
                mask.masked_fill_(~keep[:, None], torch.finfo(q_sel.dtype).min)
                kw = dict(kwargs)
                kw["is_causal"] = False
                attn_output, attn_weights = attention_interface(
                    self, q_sel, key_states, value_states, mask,
                    dropout=0.0 if not self.training else self.attention_dropout, scaling=self.scaling, **extra, **kw)
                attn_output = attn_output.reshape(b, t_keep, -1).contiguous()
                self.rows_kept = rows
            else:
                attn_output, attn_weights = attention_interface(
                    self, query_states, key_states, value_states, attention_mask,
                    dropout=0.0 if not self.training else self.attention_dropout, scaling=self.scaling, **extra, **kwargs)
                attn_output = attn_output.reshape(*input_shape, -1).contiguous()
        attn_output = self.o_proj(attn_output)
        return attn_output, attn_weights

    def decoder_forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_values=None,
                        use_cache=False, position_embeddings=None, **kwargs):
        residual = hidden_states
        hidden_states = self.input_layernorm(hidden_states)
        hidden_states, _ = self.self_attn(
            hidden_states=hidden_states, attention_mask=attention_mask, position_ids=position_ids,
            past_key_values=past_key_values, use_cache=use_cache, position_embeddings=position_embeddings, **kwargs)
        rows_kept = getattr(self.self_attn, "rows_kept", None)
        kept_positions = None
        if rows_kept is not None:
            # the attention already returned the TSP-kept rows only: gather the residual stream (and the position ids)
            # here instead of after the MLP; everything below is per token, so the result is the reference's
            residual, kept_positions = ops.gather_rows(residual, rows_kept, position_ids)
        hidden_states = residual + hidden_states
        residual = hidden_states
        hidden_states = self.post_attention_layernorm(hidden_states)
        hidden_states = self.mlp(hidden_states)
        hidden_states = residual + hidden_states

        # [FastKV] select important tokens and adjust position_ids at the TSP layer
        tsp_idx = getattr(self.self_attn, "tsp_idx", None)
        cluster = getattr(self.self_attn, "kv_cluster", None)
        if rows_kept is not None:
            self.new_position_ids = kept_positions
        elif cluster is not None and getattr(cluster, "tsp_layer", False) and tsp_idx is not None:
            hidden_states, self.new_position_ids = ops.gather_rows(hidden_states, tsp_idx, position_ids)
        else:
            self.new_position_ids = None
        return hidden_states

    def model_forward(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None,
                      inputs_embeds=None, use_cache=None, **kwargs):
        if (input_ids is None) ^ (inputs_embeds is not None):
            raise ValueError("You must specify exactly one of input_ids or inputs_embeds")
        for drop in ("output_attentions", "output_hidden_states", "return_dict", "cache_position"):
            kwargs.pop(drop, None)
        if inputs_embeds is None:
            inputs_embeds = self.embed_tokens(input_ids)
        if use_cache is None:
            use_cache = self.config.use_cache
        if use_cache and past_key_values is None:
            past_key_values = CompactKVCache(self.config, max_new_tokens=getattr(self.config, "fastkv_max_new_tokens", 256))
        if position_ids is None:
            past_seen = past_key_values.get_seq_length() if past_key_values is not None else 0
            position_ids = (torch.arange(inputs_embeds.shape[1], device=inputs_embeds.device) + past_seen).unsqueeze(0)
        q_len = inputs_embeds.shape[1]
        if isinstance(past_key_values, CompactKVCache):
            past_key_values.note_tokens(q_len)         # true (un-compacted) token count, read by the generate() path
        if attention_mask is not None and attention_mask.dim() == 2 and not bool(attention_mask.all()):
            # the reference's padded-batch path is broken (llama_model.py:57,77)
            raise NotImplementedError("padded batches are not supported on the FastKV path")
        # un-padded prompts: causal attention (prefill, optionally inside Mistral's sliding window) / everything
        # visible (decode over a compact cache that is always shorter than the window)
        hidden_states = inputs_embeds
        position_embeddings = self.rotary_emb(hidden_states, position_ids=position_ids)
        causal_mask, is_causal_kw = _prefill_mask(self.config, q_len, hidden_states, sliding_window_attr)
        # prefill into the pre-allocated compact cache: the select + compact stage of every layer runs on a side stream
        # under the layer's own attention / MLP (nothing reads the compact K/V before the first decode step)
        overlap = None
        if (q_len > 1 and method == "fastkv" and hidden_states.is_cuda and isinstance(past_key_values, CompactKVCache)
                and getattr(self.config, "fastkv_overlap_select", True)):
            overlap = getattr(self, "_fkv_overlap", None)
            if overlap is None or overlap.device != hidden_states.device:
                overlap = self._fkv_overlap = ops.SelectOverlap(hidden_states.device)
        for decoder_layer in self.layers[: self.config.num_hidden_layers]:
            cluster = getattr(decoder_layer.self_attn, "kv_cluster", None)
            if cluster is not None and hasattr(cluster, "overlap"):
                cluster.overlap = overlap
            hidden_states = decoder_layer(
                hidden_states, attention_mask=causal_mask, position_embeddings=position_embeddings,
                position_ids=position_ids, past_key_values=past_key_values, use_cache=use_cache, **is_causal_kw, **kwargs)
            new_position_ids = getattr(decoder_layer, "new_position_ids", None)
            if new_position_ids is not None:          # [FastKV] later layers run on the TSP-selected tokens only
                position_ids = new_position_ids
                position_embeddings = self.rotary_emb(hidden_states, position_ids=position_ids)
                # the kept tokens stay in ascending order: still causal (and banded by INDEX, as the reference's
                # flash-attn call is) over the shorter sequence
                causal_mask, is_causal_kw = _prefill_mask(self.config, hidden_states.shape[1], hidden_states, sliding_window_attr)
        if overlap is not None:
            overlap.join()                            # the compact cache is complete when the prefill returns
            for decoder_layer in self.layers[: self.config.num_hidden_layers]:
                cluster = getattr(decoder_layer.self_attn, "kv_cluster", None)
                if cluster is not None and hasattr(cluster, "overlap"):
                    cluster.overlap = None
        hidden_states = self.norm(hidden_states)
        hidden_states = hidden_states[:, -1:, :]      # [FastKV] only the last token feeds lm_head
        return BaseModelOutputWithPast(last_hidden_state=hidden_states, past_key_values=past_key_values)

    return attn_forward, decoder_forward, model_forward


def _make_gemfilter_attn_forward(mod, sliding_window_attr=False):
    """Attention forward of the GemFilter baseline (baselines/gemfilter/llama_model.py:104-160, mistral twin):
    in select mode the selection layer records the token indices (``find_context``, our ``fkv_score`` GemFilter
    mode + ``fkv_topk``) and nothing is cached; otherwise stock attention with a stock cache update."""
    apply_rotary_pos_emb = mod.apply_rotary_pos_emb
    eager_attention_forward = mod.eager_attention_forward
    ALL_ATTENTION_FUNCTIONS = mod.ALL_ATTENTION_FUNCTIONS
    from .cluster import find_context

    def attn_forward(self, hidden_states, position_embeddings=None, attention_mask=None, past_key_values=None, **kwargs):
        input_shape = hidden_states.shape[:-1]
        hidden_shape = (*input_shape, -1, self.head_dim)
        query_states = self.q_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        key_states = self.k_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        value_states = self.v_proj(hidden_states).view(hidden_shape).transpose(1, 2)
        cos, sin = position_embeddings
        query_states, key_states = apply_rotary_pos_emb(query_states, key_states, cos, sin)
        if self.select_mode:                                   # llama_model.py:141-143
            self.indecies = None
            find_context(self, query_states, key_states)
        if not self.select_mode and past_key_values is not None:   # :145-150
            key_states, value_states = past_key_values.update(key_states, value_states, self.layer_idx)
        attention_interface = ALL_ATTENTION_FUNCTIONS.get_interface(self.config._attn_implementation, eager_attention_forward)
        extra = {"sliding_window": getattr(self.config, "sliding_window", None)} if sliding_window_attr else {}
        attn_output, attn_weights = attention_interface(
            self, query_states, key_states, value_states, attention_mask,
            dropout=0.0 if not self.training else self.attention_dropout, scaling=self.scaling, **extra, **kwargs)
        attn_output = self.o_proj(attn_output.reshape(*input_shape, -1).contiguous())
        return attn_output, attn_weights

    return attn_forward


# ---------------------------------------------------------------------------------------------
# generate() path: decode positions continue from the TRUE prompt length (monkeypatch.py:249-389)
# ---------------------------------------------------------------------------------------------
def _prepare_inputs_for_generation(self, input_ids, past_key_values=None, attention_mask=None, inputs_embeds=None,
                                   cache_position=None, position_ids=None, use_cache=True, **kwargs):
    """``prepare_inputs_for_generation_llama`` / ``_mistral`` of the reference (baselines/monkeypatch.py:249-330,
    :333-389), re-hosted on the installed transformers.

    The point of the override (:280-288): with a COMPACTED cache the number of cached rows says nothing about
    where the next token sits, so position ids come from the attention mask (``cumsum - 1``, last
    ``input_ids.shape[1]`` entries) — i.e. they continue from the true prompt length — or, when no mask is
    given, from the cache's own count of tokens seen (``CompactKVCache.true_length``)."""
    has_past = past_key_values is not None and past_key_values.get_seq_length() > 0
    true_len = getattr(past_key_values, "true_length", None) if has_past else None
    if has_past:
        if inputs_embeds is not None and cache_position is not None:                      # Exception 1 (:273-274)
            input_ids = input_ids[:, -cache_position.shape[0]:]
        elif cache_position is not None and input_ids.shape[1] != cache_position.shape[0]:  # default case (:275-276)
            input_ids = input_ids[:, cache_position]
        elif cache_position is None and input_ids.shape[1] > 1:
            # transformers >= 5 passes no cache_position but the number of unprocessed tokens
            nxt = kwargs.get("next_sequence_length")
            if nxt is not None:
                input_ids = input_ids[:, -int(nxt):]
            else:
                seen = true_len if true_len else (attention_mask.shape[1] - 1 if attention_mask is not None else input_ids.shape[1] - 1)
                input_ids = input_ids[:, seen:]
    if attention_mask is not None and position_ids is None:                               # :278-288
        position_ids = attention_mask.long().cumsum(-1) - 1
        position_ids.masked_fill_(attention_mask == 0, 1)
        if has_past:
            position_ids = position_ids[:, -input_ids.shape[1]:]
        position_ids = position_ids.clone(memory_format=torch.contiguous_format)
    elif position_ids is None and has_past and true_len:
        n = input_ids.shape[1]
        position_ids = torch.arange(true_len, true_len + n, device=input_ids.device).unsqueeze(0).expand(input_ids.shape[0], -1).contiguous()
    elif position_ids is not None and has_past and position_ids.shape[1] != input_ids.shape[1]:
        # full-length position ids built by generate() from the mask: keep the unprocessed tokens' (true) positions
        position_ids = position_ids[:, -input_ids.shape[1]:].clone(memory_format=torch.contiguous_format)
    if inputs_embeds is not None and not has_past:                                        # :291-295
        model_inputs = {"inputs_embeds": inputs_embeds, "input_ids": None}
    else:
        model_inputs = {"input_ids": input_ids.clone(memory_format=torch.contiguous_format), "inputs_embeds": None}
    model_inputs.update({"position_ids": position_ids, "cache_position": cache_position, "past_key_values": past_key_values,
                         "use_cache": use_cache, "attention_mask": attention_mask})
    for k, v in kwargs.items():                        # e.g. logits_to_keep: pass through untouched
        if k not in ("next_sequence_length", "is_first_iteration"):
            model_inputs.setdefault(k, v)
    return model_inputs


def prepare_inputs_for_generation_llama(self, input_ids, past_key_values=None, attention_mask=None, inputs_embeds=None,
                                        cache_position=None, position_ids=None, use_cache=True, **kwargs):
    """baselines/monkeypatch.py:249-330."""
    return _prepare_inputs_for_generation(self, input_ids, past_key_values, attention_mask, inputs_embeds, cache_position,
                                          position_ids, use_cache, **kwargs)


def prepare_inputs_for_generation_mistral(self, input_ids, past_key_values=None, attention_mask=None, inputs_embeds=None,
                                          cache_position=None, position_ids=None, use_cache=True, **kwargs):
    """baselines/monkeypatch.py:333-389 (the sliding-window crop of the attention mask there, :352-360, only acts on
    caches longer than the window, which a compact cache never is)."""
    return _prepare_inputs_for_generation(self, input_ids, past_key_values, attention_mask, inputs_embeds, cache_position,
                                          position_ids, use_cache, **kwargs)


def _patch(mod, prefix, method, sliding):
    attn_cls = getattr(mod, prefix + "Attention")
    layer_cls = getattr(mod, prefix + "DecoderLayer")
    model_cls = getattr(mod, prefix + "Model")
    lm_cls = getattr(mod, prefix + "ForCausalLM")
    key = prefix
    if key not in _ORIGINALS:
        _ORIGINALS[key] = (attn_cls.forward, layer_cls.forward, model_cls.forward, attn_cls.__init__,
                           lm_cls.prepare_inputs_for_generation)
    o_attn, o_layer, o_model, o_init, o_prep = _ORIGINALS[key]
    if method == "fullkv":
        attn_cls.forward, layer_cls.forward, model_cls.forward, attn_cls.__init__ = o_attn, o_layer, o_model, o_init
        lm_cls.prepare_inputs_for_generation = o_prep
        return
    if method not in ("fastkv", "snapkv", "gemfilter"):
        raise NotImplementedError(f"No method found for {method}: only fastkv / snapkv / gemfilter / fullkv are on the hot path")
    # every method but fullkv decodes from the true prompt length (monkeypatch.py:55-56, :101-102)
    lm_cls.prepare_inputs_for_generation = (prepare_inputs_for_generation_mistral if prefix == "Mistral"
                                            else prepare_inputs_for_generation_llama)
    if method == "gemfilter":
        # selection-mode baseline: only the attention forward differs from stock (gemfilter/llama_model.py)
        gf_forward = _make_gemfilter_attn_forward(mod, sliding_window_attr=sliding)

        def gemfilter_init(self, *args, **kwargs):             # LlamaGemFilterAttention.__init__ (:88-98)
            o_init(self, *args, **kwargs)
            self.indecies = None
            self.topk = 1024
            self.rate = 0.25
            self.retain_rate = 0.25
            self.eviction_mode = "constant"
            self.select_layer_idx = 13
            self.select_mode = False

        attn_cls.__init__, attn_cls.forward = gemfilter_init, gf_forward
        layer_cls.forward, model_cls.forward = o_layer, o_model
        return
    attn_f, layer_f, model_f = _make_forwards(mod, method, sliding_window_attr=sliding)

    def patched_init(self, *args, **kwargs):
        o_init(self, *args, **kwargs)
        if method == "fastkv":
            init_fastkv(self)                          # LlamaFastKVAttention.__init__ (llama_model.py:89-91)
        self.tsp_idx = None

    attn_cls.__init__ = patched_init
    attn_cls.forward = attn_f
    layer_cls.forward = layer_f
    model_cls.forward = model_f


def replace_llama(method):
    """baselines/monkeypatch.py:12-56."""
    from transformers.models.llama import modeling_llama as mod
    _patch(mod, "Llama", method, sliding=False)


def replace_mistral(method):
    """baselines/monkeypatch.py:59-102."""
    from transformers.models.mistral import modeling_mistral as mod
    _patch(mod, "Mistral", method, sliding=True)


# ---------------------------------------------------------------------------------------------
# CUDA-graphed decode step over the compact cache
# ---------------------------------------------------------------------------------------------
class GraphedDecode:
    """One decode step (all layers: projections, in-place cache append, K5 attention, MLP, lm_head,
    greedy argmax) captured once as a CUDA graph and replayed per token.  The cache lengths and the
    position id live on the device, so replays need no host work at all."""

    def __init__(self, model, cache, first_token, warmup=2):
        self.model, self.cache = model, cache
        dev = first_token.device
        self.tok = first_token.clone()
        # the reference's benchmark loop passes no position ids: positions continue from the
        # (compact) layer-0 cache length (benchmark/e2e.py:174)
        self.pos = torch.full((first_token.shape[0], 1), cache.get_seq_length(), dtype=torch.int64, device=dev)
        cache.device_lengths = True
        self.eager_tokens = []
        for _ in range(warmup):                       # real decode steps, run eagerly (allocations, attributes)
            self._step()
            self.eager_tokens.append(self.tok.clone())
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step()
        # the capture itself did not execute: nothing to undo on the device; the host-side token count did advance
        cache.true_length -= 1
        # the graph bakes in the scratch pointer of the decode kernels: keep that buffer alive even if a
        # later, larger call replaces it in the per-stream table
        from . import _lib as _L
        self._scratch_refs = list(_L._workspaces.values())

    @torch.no_grad()
    def _step(self):
        out = self.model(input_ids=self.tok, past_key_values=self.cache, position_ids=self.pos, use_cache=True)
        self.tok.copy_(out.logits[:, -1, :].argmax(dim=-1, keepdim=True))
        self.pos.add_(1)

    @torch.no_grad()
    def generate(self, n):
        """n more tokens (after the warm-up ones); returns [B, warmup + n] token ids."""
        self.cache.sync_lengths()
        room = min(k.shape[2] - s for k, s in zip(self.cache.k, self.cache.seq_lens) if k is not None)
        if n > room:
            # the replayed step appends through device-side lengths: the host check of CompactKVCache.update is not on
            # that path, so refuse here instead of dropping rows silently
            raise RuntimeError(f"CompactKVCache has room for {room} more tokens, {n} requested: raise max_new_tokens")
        toks = list(self.eager_tokens)
        for _ in range(n):
            self.graph.replay()
            toks.append(self.tok.clone())
        self.cache.sync_lengths()
        self.cache.true_length += n
        return torch.cat(toks, dim=1)


# ---------------------------------------------------------------------------------------------
# the reference's hand-rolled greedy loop (benchmark/e2e.py:162-183), used by bench / tests
# ---------------------------------------------------------------------------------------------
@torch.no_grad()
def greedy_generate(model, input_ids, genlen, attention_mask=None, max_new_tokens=None):
    """Prefill + (genlen - 1) decode steps; returns (token ids [B, genlen], cache)."""
    model.config.fastkv_max_new_tokens = max_new_tokens or genlen
    if hasattr(model, "model"):
        model.model.config.fastkv_max_new_tokens = model.config.fastkv_max_new_tokens
    out = model(input_ids, attention_mask=attention_mask, use_cache=True)
    cache = out.past_key_values
    tok = out.logits[:, -1, :].argmax(dim=-1, keepdim=True)
    toks = [tok]
    for _ in range(genlen - 1):
        out = model(input_ids=tok, past_key_values=cache, use_cache=True)
        tok = out.logits[:, -1, :].argmax(dim=-1, keepdim=True)
        toks.append(tok)
    return torch.cat(toks, dim=1), cache


# ---------------------------------------------------------------------------------------------
# GemFilter two-pass driver (baselines/gemfilter/utils.py:58-158)
# ---------------------------------------------------------------------------------------------
def set_topk(model, args, mode="gemfilter"):
    """gemfilter/utils.py:78-89."""
    if mode != "gemfilter":
        raise NotImplementedError
    for layer in model.model.layers:
        layer.self_attn.topk = args.max_capacity_prompts
        layer.self_attn.retain_rate = args.retain_rate
        layer.self_attn.eviction_mode = args.eviction_mode


def set_select_mode(model, mode):
    """gemfilter/utils.py:91-95."""
    for layer in model.model.layers:
        layer.self_attn.select_mode = mode


def set_select_layer(model, select_layer_idx):
    """gemfilter/utils.py:97-104."""
    if select_layer_idx is None:
        return model.model.layers[0].self_attn.select_layer_idx
    for layer in model.model.layers:
        layer.self_attn.select_layer_idx = select_layer_idx
    return select_layer_idx


def reduce_layer(model, layer_idx):
    """gemfilter/utils.py:68-71: run only layers [0, layer_idx] in the selection pass."""
    original_layers = model.model.layers
    model.model.layers = model.model.layers[:layer_idx + 1]
    return model, original_layers


def recover_layer(model, layers):
    """gemfilter/utils.py:73-75."""
    model.model.layers = layers
    return model


def get_layer_context(model, tokenizer, input_ids, layer_idx, print_context=False):
    """gemfilter/utils.py:58-66: the selected token ids in their original order, shape [1, k]."""
    idx = model.model.layers[layer_idx].self_attn.indecies[0, 0, :]
    values, _ = torch.sort(idx)
    new_input_ids = input_ids.gather(0, values.to(input_ids.device))
    if print_context and tokenizer is not None:
        print(tokenizer.decode(new_input_ids))
    return new_input_ids.unsqueeze(0)


@torch.no_grad()
def gemfilter_generate_selection_prefill(input_ids, attn_mask, model, tokenizer=None, select_layer_idx=None, print_context=False):
    """gemfilter/utils.py:140-158: pass 1 over layers [0, select layer] picks the tokens, pass 2 prefills the
    whole model on them.  Returns (first predicted token [1,1], past_key_values)."""
    set_select_mode(model, True)
    select_layer_idx = set_select_layer(model, select_layer_idx)
    model, original_layers = reduce_layer(model, select_layer_idx)
    try:
        model(input_ids, attention_mask=attn_mask, use_cache=False, logits_to_keep=1)
        new_input_ids = get_layer_context(model, tokenizer, input_ids[0], select_layer_idx, print_context=print_context)
    finally:
        model = recover_layer(model, original_layers)
        set_select_mode(model, False)
    outputs = model(new_input_ids, use_cache=True, logits_to_keep=1)
    pred_token_idx = outputs.logits[:, -1, :].argmax(dim=-1).unsqueeze(1)
    return pred_token_idx, outputs.past_key_values


@torch.no_grad()
def gemfilter_generate(model, tokenizer, pred_token_idx, past_key_values, max_gen_len):
    """gemfilter/utils.py:107-121 (greedy loop; stops at the tokenizer's EOS when a tokenizer is given)."""
    eos = getattr(tokenizer, "eos_token_id", None)
    generated_ids = [pred_token_idx.item()]
    for _ in range(max_gen_len):
        outputs = model(input_ids=pred_token_idx, past_key_values=past_key_values, use_cache=True)
        past_key_values = outputs.past_key_values
        pred_token_idx = outputs.logits[:, -1, :].argmax(dim=-1).unsqueeze(1)
        if eos is not None and pred_token_idx.item() == eos:
            break
        generated_ids.append(pred_token_idx.item())
    return generated_ids


@torch.no_grad()
def gemfilter_generate_selection(input_ids, attn_mask, model, tokenizer=None, max_gen_len=50, select_layer_idx=None,
                                 print_context=False):
    """gemfilter/utils.py:123-138.  Returns the decoded string with a tokenizer, the token ids without one."""
    pred, cache = gemfilter_generate_selection_prefill(input_ids, attn_mask, model, tokenizer, select_layer_idx, print_context)
    output_ids = gemfilter_generate(model, tokenizer, pred, cache, max_gen_len=max_gen_len)
    if tokenizer is None:
        return output_ids
    return tokenizer.decode(output_ids, skip_special_tokens=True).strip()
