"""CPU tests of the monkeypatch plumbing (no kernels run: prompts shorter than the KV budget take
the reference's early-return path, baselines/fastkv/utils.py:89-91)."""
import types

import pytest
import torch


def _tiny_llama(layers=2):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=layers, num_attention_heads=2,
                      num_key_value_heads=1, head_dim=64, vocab_size=97, max_position_embeddings=512,
                      attn_implementation="sdpa")
    torch.manual_seed(0)
    return LlamaForCausalLM(cfg).eval()


def _args(**kw):
    base = dict(method="fastkv", max_capacity_prompts=64, window_size=8, kernel_size=7, pooling="avgpool",
                retain_rate=0.1, eviction_mode="constant", tsp_len=32, tsp_rate=0.2, tsp_idx=1, merge=None)
    base.update(kw)
    return types.SimpleNamespace(**base)


@pytest.fixture
def patched():
    from fastkv_b200 import patch
    patch.replace_llama("fastkv")
    yield patch
    patch.replace_llama("fullkv")


def test_set_model_writes_the_reference_knobs(patched):
    model = _tiny_llama(3)
    args = _args(tsp_idx=1, eviction_mode="proportional")
    patched.set_model(model, args)
    cl = [l.self_attn.kv_cluster for l in model.model.layers]
    assert [c.tsp_layer for c in cl] == [False, True, False]                 # fastkv/utils.py:38-41
    assert cl[0].retain_rate == 0.1 and cl[2].retain_rate == pytest.approx(0.1 / 0.2)   # :43-46
    assert all(c.window_size == 8 and c.kernel_size == 7 and c.max_capacity_prompt == 64 for c in cl)
    assert all(c.tsp_length == 32 and c.pooling == "avgpool" and c.eviction_mode == "proportional" for c in cl)
    assert model.model.layers[0].self_attn.config.max_capacity_prompt == 64  # monkeypatch.py:132-139
    assert isinstance(args.window_size, list)                                # :145-146


def test_short_prompt_matches_stock_model_and_decodes(patched):
    from fastkv_b200.cache import CompactKVCache
    model = _tiny_llama()
    ids = torch.randint(0, 97, (1, 20), generator=torch.Generator().manual_seed(1))
    patched.replace_llama("fullkv")
    with torch.no_grad():
        ref = model(ids, use_cache=True)
        ref_next = ref.logits[:, -1].argmax(-1, keepdim=True)
        ref2 = model(ref_next, past_key_values=ref.past_key_values, use_cache=True).logits[:, -1]
    patched.replace_llama("fastkv")
    patched.set_model(model, _args())
    with torch.no_grad():
        out = model(ids, use_cache=True)
    assert out.logits.shape[1] == 1                                          # last-token cut (llama_model.py:392)
    assert torch.allclose(out.logits[:, -1], ref.logits[:, -1], atol=1e-5)
    cache = out.past_key_values
    assert isinstance(cache, CompactKVCache) and cache.get_seq_length() == 20
    with torch.no_grad():
        out2 = model(out.logits[:, -1].argmax(-1, keepdim=True), past_key_values=cache, use_cache=True)
    assert cache.get_seq_length() == 21
    assert torch.allclose(out2.logits[:, -1], ref2, atol=1e-4)
    toks, _ = patched.greedy_generate(model, ids, 4)
    assert toks.shape == (1, 4) and toks[0, 0] == ref_next[0, 0]


def test_unsupported_methods_fail_loudly(patched):
    with pytest.raises(NotImplementedError):
        patched.replace_llama("h2o")
    with pytest.raises(NotImplementedError):
        patched.set_model(_tiny_llama(), _args(method="streamingllm"))


def test_mistral_patch_installs(patched):
    from transformers.models.mistral import modeling_mistral as mm
    orig = mm.MistralModel.forward
    patched.replace_mistral("fastkv")
    assert mm.MistralModel.forward is not orig
    patched.replace_mistral("fullkv")
    assert mm.MistralModel.forward is orig


def test_compact_cache_update_contract():
    from fastkv_b200.cache import CompactKVCache
    c = CompactKVCache(max_new_tokens=4)
    k, v = torch.randn(2, 2, 10, 64), torch.randn(2, 2, 10, 64)
    ko, vo = c.update(k, v, 0)
    assert ko.shape == (2, 2, 10, 64) and torch.equal(ko, k) and c.get_seq_length() == 10
    k1, v1 = torch.randn(2, 2, 1, 64), torch.randn(2, 2, 1, 64)
    ko, vo = c.update(k1, v1, 0)
    assert ko.shape[2] == 11 and torch.equal(ko[:, :, 10], k1[:, :, 0]) and torch.equal(vo[:, :, :10], v)
    for _ in range(3):
        c.update(k1, v1, 0)
    with pytest.raises(RuntimeError):
        c.update(k1, v1, 0)


def test_gemfilter_patch_plumbing():
    """baselines/gemfilter: attention attributes, set_topk via set_model, layer truncation for the selection pass,
    and the stock path when select_mode is off (no kernels involved on CPU)."""
    from fastkv_b200 import patch
    patch.replace_llama("gemfilter")
    try:
        model = _tiny_llama(3)
        attn = model.model.layers[0].self_attn
        assert (attn.topk, attn.select_layer_idx, attn.select_mode, attn.indecies) == (1024, 13, False, None)   # llama_model.py:88-98
        patch.set_model(model, _args(method="gemfilter", max_capacity_prompts=40, retain_rate=0.3, eviction_mode="proportional"))
        assert all(l.self_attn.topk == 40 and l.self_attn.retain_rate == 0.3 and l.self_attn.eviction_mode == "proportional"
                   for l in model.model.layers)                                                                # utils.py:78-89
        assert patch.set_select_layer(model, 1) == 1 and all(l.self_attn.select_layer_idx == 1 for l in model.model.layers)
        m2, orig = patch.reduce_layer(model, 1)
        assert len(m2.model.layers) == 2
        patch.recover_layer(model, orig)
        assert len(model.model.layers) == 3
        # select mode off: identical to the stock model
        ids = torch.randint(0, 97, (1, 20), generator=torch.Generator().manual_seed(1))
        with torch.no_grad():
            got = model(ids, use_cache=True).logits
        patch.replace_llama("fullkv")
        with torch.no_grad():
            ref = model(ids, use_cache=True).logits
        assert torch.allclose(got, ref, atol=1e-5)
        # the selection itself needs the CUDA library: no CPU fallback
        patch.replace_llama("gemfilter")
        patch.set_select_mode(model, True)
        patch.set_select_layer(model, 0)
        with pytest.raises(RuntimeError):
            with torch.no_grad():
                model(ids, use_cache=False)
    finally:
        patch.replace_llama("fullkv")


def test_bench_reference_arm_contract():
    """bench.py --impl reference (the oracle port on host cores) prints the contract's JSON line without a GPU."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert key in line, key
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("port", "reference")
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["config"]["workload"]


class _KeepFirstRows:
    """Test double for the per-layer compressor (CPU tests cannot run the CUDA kernels): keeps the first C-W rows and
    the W window rows, so the cache is COMPACT and the position rule of the generate() path can be checked."""
    tsp_layer = False

    def __init__(self, capacity=16, window=4):
        self.c, self.w = capacity, window

    def update_kv(self, k, q, v, attention_mask, groups, layer_idx=None):
        if k.shape[2] < self.c:
            return k, v, None
        sel = lambda x: torch.cat([x[:, :, :self.c - self.w], x[:, :, -self.w:]], dim=2).contiguous()
        return sel(k), sel(v), None


@pytest.mark.parametrize("with_mask,own_cache", [(True, False), (False, True), (True, True)])
def test_generate_path_positions_follow_the_true_length(patched, with_mask, own_cache):
    """baselines/monkeypatch.py:280-288: decode position ids continue from the true prompt length although the cache
    holds only the compacted rows — from the attention mask when there is one, else from CompactKVCache.true_length."""
    from fastkv_b200.cache import CompactKVCache
    model = _tiny_llama()
    model.generation_config.pad_token_id = 0
    patched.set_model(model, _args())
    for layer in model.model.layers:
        layer.self_attn.kv_cluster = _KeepFirstRows()
    n, new = 40, 5
    ids = torch.randint(1, 97, (1, n), generator=torch.Generator().manual_seed(2))
    seen, inner = [], model.model.forward

    def spy(*a, **kw):
        pos = kw.get("position_ids")
        seen.append(None if pos is None else pos.flatten().tolist())
        return inner(*a, **kw)

    model.model.forward = spy
    kw = {}
    if with_mask:
        kw["attention_mask"] = torch.ones_like(ids)
    cache = CompactKVCache(model.config, max_new_tokens=8) if own_cache else None
    if own_cache:
        kw["past_key_values"] = cache
    if with_mask:
        out = model.generate(ids, max_new_tokens=new, do_sample=False, **kw)
        assert out.shape == (1, n + new)
    else:
        # no mask: drive prepare_inputs_for_generation by hand, like a custom loop would
        with torch.no_grad():
            cur = ids
            for _ in range(new):
                inputs = model.prepare_inputs_for_generation(cur, past_key_values=cache, use_cache=True)
                logits = model(**inputs).logits
                cur = torch.cat([cur, logits[:, -1].argmax(-1, keepdim=True)], dim=1)
    assert seen[0] is None or seen[0] == list(range(n))
    assert [p[-1] for p in seen[1:]] == list(range(n, n + new - 1)), seen[1:]
    if own_cache:
        assert cache.get_seq_length() == 16 + new - 1 and cache.true_length == n + new - 1


class _KeepEveryThird:
    """Compressor double with a TSP selection (CPU): keeps the first C-W rows + window in the cache and propagates every
    third token plus the window — enough to drive the glue of the TSP layer without a GPU."""

    def __init__(self, tsp_layer, capacity=16, window=4):
        self.tsp_layer, self.c, self.w = tsp_layer, capacity, window

    def update_kv(self, k, q, v, attention_mask, groups, layer_idx=None):
        n = k.shape[2]
        if n < self.c:
            return k, v, None
        sel = lambda x: torch.cat([x[:, :, :self.c - self.w], x[:, :, -self.w:]], dim=2).contiguous()
        tsp = None
        if self.tsp_layer:
            keep = torch.cat([torch.arange(0, n - self.w, 3), torch.arange(n - self.w, n)])
            tsp = keep.unsqueeze(0).expand(k.shape[0], -1).contiguous()
        return sel(k), sel(v), tsp


@pytest.mark.parametrize("impl,window", [("sdpa", None), ("eager", None), ("eager", 20)])
def test_tsp_layer_row_pruning_matches_the_reference_order_on_cpu(patched, monkeypatch, impl, window):
    """The TSP layer computes attention / MLP rows for the kept tokens only (default) or for all tokens with the gather
    after the MLP (``fastkv_sparse_tsp_layer = False``, the reference's order, llama_model.py:252-259): same hidden states,
    same propagated positions.  ``ops.gather_rows`` is the CUDA kernel: a torch double stands in for it here."""
    from fastkv_b200 import ops, patch as P

    def gather_rows_double(hidden, idx, position_ids=None):
        out = hidden.gather(1, idx[..., None].expand(-1, -1, hidden.shape[-1]))
        if position_ids is None:
            return out
        return out, position_ids.expand(hidden.shape[0], -1).gather(1, idx)

    monkeypatch.setattr(ops, "gather_rows", gather_rows_double)
    if window is not None:
        from transformers import MistralConfig, MistralForCausalLM
        P.replace_mistral("fastkv")
        cfg = MistralConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, num_key_value_heads=1,
                            head_dim=64, vocab_size=97, max_position_embeddings=512, sliding_window=window, attn_implementation=impl)
        make = lambda: MistralForCausalLM(cfg).eval()
    else:
        from transformers import LlamaConfig, LlamaForCausalLM
        cfg = LlamaConfig(hidden_size=128, intermediate_size=256, num_hidden_layers=3, num_attention_heads=2, num_key_value_heads=1,
                          head_dim=64, vocab_size=97, max_position_embeddings=512, attn_implementation=impl)
        make = lambda: LlamaForCausalLM(cfg).eval()
    ids = torch.randint(1, 97, (2, 50), generator=torch.Generator().manual_seed(5))
    outs = []
    try:
        for sparse in (False, True):
            torch.manual_seed(0)
            model = make()
            model.config.fastkv_sparse_tsp_layer = sparse
            patched.set_model(model, _args())
            for i, layer in enumerate(model.model.layers):
                layer.self_attn.kv_cluster = _KeepEveryThird(tsp_layer=(i == 1))
            with torch.no_grad():
                out = model.model(ids, use_cache=True)
            assert (model.model.layers[1].self_attn.rows_kept is not None) == sparse
            assert model.model.layers[1].new_position_ids.shape == (2, 16 + 4)
            outs.append((out.last_hidden_state, model.model.layers[1].new_position_ids.clone(),
                         out.past_key_values.get_seq_length(2)))
    finally:
        if window is not None:
            P.replace_mistral("fullkv")
    (h0, p0, n0), (h1, p1, n1) = outs
    assert torch.equal(p0, p1) and n0 == n1
    assert torch.allclose(h0, h1, atol=2e-5, rtol=1e-4), (h0 - h1).abs().max()
