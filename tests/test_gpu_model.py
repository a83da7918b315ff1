"""GPU tests of the patched model forwards against the CPU glue oracle."""
import types

import pytest
import torch

from oracle import fastkv_oracle as O
from oracle import glue_oracle as G

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _model(dtype, layers=3, seed=0):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=layers, num_attention_heads=8,
                      num_key_value_heads=2, head_dim=64, vocab_size=211, max_position_embeddings=4096,
                      attn_implementation="sdpa")
    torch.manual_seed(seed)
    return LlamaForCausalLM(cfg).eval().to(dtype)


def _args(**kw):
    base = dict(method="fastkv", max_capacity_prompts=128, window_size=8, kernel_size=7, pooling="avgpool",
                retain_rate=0.1, eviction_mode="constant", tsp_len=256, tsp_rate=0.2, tsp_idx=1, merge=None)
    base.update(kw)
    return types.SimpleNamespace(**base)


def _check_cache_rows(idx, ref_idx, k_got, v_got, k_ref, v_ref, w):
    """Compact cache == the oracle's at every position where the score-descending orders agree (only near-tie swaps
    may differ), window rows always."""
    same = idx == ref_idx
    assert same.float().mean().item() > 0.97, same.float().mean().item()
    nsel = idx.shape[-1]
    assert torch.allclose(k_got[:, :, :nsel][same], k_ref[:, :, :nsel][same], atol=2e-4, rtol=1e-4)
    assert torch.allclose(v_got[:, :, :nsel][same], v_ref[:, :, :nsel][same], atol=2e-4, rtol=1e-4)
    assert torch.allclose(k_got[:, :, nsel:], k_ref[:, :, nsel:], atol=2e-4, rtol=1e-4)
    assert torch.allclose(v_got[:, :, nsel:], v_ref[:, :, nsel:], atol=2e-4, rtol=1e-4)


@pytest.fixture
def patched():
    from fastkv_b200 import patch
    patch.replace_llama("fastkv")
    yield patch
    patch.replace_llama("fullkv")


def test_fp32_prefill_matches_glue_oracle(patched):
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    model = _model(torch.float32)
    ids = torch.randint(0, 211, (2, 600), generator=torch.Generator().manual_seed(3))
    args = _args()
    knobs = [dict(window_size=8, max_capacity_prompt=128, kernel_size=7, pooling="avgpool", tsp_layer=(i == 1), tsp_length=256)
             for i in range(3)]
    patched.replace_llama("fullkv")
    ref_hidden, ref_caches, traces = G.fastkv_prefill(model, ids, knobs, apply_rotary_pos_emb)
    patched.replace_llama("fastkv")
    gm = _model(torch.float32).to(DEV)
    patched.set_model(gm, args)
    for layer in gm.model.layers:
        layer.self_attn.kv_cluster.keep_aux = True
    with torch.no_grad():
        out = gm.model(ids.to(DEV), use_cache=True)
    torch.cuda.synchronize()
    cache = out.past_key_values
    # TSP layer: same token set (near-tie aware against the oracle's own importance)
    tsp = gm.model.layers[1].self_attn.tsp_idx
    assert tsp is not None and tsp.shape == (2, 256)
    for b in range(2):
        assert O.near_tie_set_mismatch(traces[1].importance[b], ref_caches[1][2][b, :-8], tsp[b, :-8].cpu()) == 0
    # later layers ran on T tokens: q_len (256) >= C (128) -> compacted again to C rows
    for i in range(3):
        k_ref, v_ref, _ = ref_caches[i]
        kbuf, vbuf, n, _ = cache.buffers(i)
        assert n == k_ref.shape[2] == 128
        _check_cache_rows(gm.model.layers[i].self_attn.kv_cluster.last_aux["idx"].cpu(), traces[i].idx, kbuf[:, :, :n].cpu(),
                          vbuf[:, :, :n].cpu(), k_ref, v_ref, 8)
    # the propagated token set is the oracle's, so the last hidden state must be too (fp32 model)
    assert torch.equal(tsp.cpu(), ref_caches[1][2])
    assert torch.allclose(out.last_hidden_state.cpu(), ref_hidden, atol=2e-3, rtol=1e-3)
    assert out.last_hidden_state.shape == (2, 1, 512)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 6e-2)])
def test_tsp_layer_computes_only_the_kept_rows(patched, dtype, tol):
    """at the TSP layer attention / o_proj / MLP run for the T kept tokens only (default);
    ``config.fastkv_sparse_tsp_layer = False`` is the reference's order (all L rows, gather after the MLP).  Same
    selection, same caches up to the TSP layer, same hidden states after it."""
    ids = torch.randint(0, 211, (2, 900), generator=torch.Generator().manual_seed(11)).to(DEV)
    outs = []
    for sparse in (False, True):
        gm = _model(dtype).to(DEV)
        gm.config.fastkv_sparse_tsp_layer = sparse
        for layer in gm.model.layers:
            layer.self_attn.config.fastkv_sparse_tsp_layer = sparse
        patched.set_model(gm, _args())
        with torch.no_grad():
            out = gm.model(ids, use_cache=True)
        torch.cuda.synchronize()
        assert (gm.model.layers[1].self_attn.rows_kept is not None) == sparse
        assert gm.model.layers[0].self_attn.rows_kept is None and gm.model.layers[2].self_attn.rows_kept is None
        outs.append((out.last_hidden_state.float().cpu(), gm.model.layers[1].self_attn.tsp_idx.cpu(),
                     [tuple(t[:, :, :out.past_key_values.buffers(i)[2]].float().cpu() for t in out.past_key_values.buffers(i)[:2])
                      for i in range(3)]))
    (h0, t0, c0), (h1, t1, c1) = outs
    assert torch.equal(t0, t1)
    for i in range(2):                                  # caches of layers 0 and 1 do not depend on the TSP layer's attention
        assert torch.equal(c0[i][0], c1[i][0]) and torch.equal(c0[i][1], c1[i][1])
    assert torch.allclose(h0, h1, atol=tol, rtol=tol)


def test_bf16_prefill_and_native_decode(patched):
    """Decode over the compact cache with the K5 kernel agrees with HF attention over the same cache."""
    from fastkv_b200 import ops
    model = _model(torch.bfloat16).to(DEV)
    patched.set_model(model, _args())
    ids = torch.randint(0, 211, (1, 1024), generator=torch.Generator().manual_seed(4)).to(DEV)
    toks, cache = patched.greedy_generate(model, ids, 6)
    assert toks.shape == (1, 6)
    assert cache.get_seq_length(0) == 128 + 5 and cache.get_seq_length(2) == 128 + 5
    # re-run the last decode step's attention of layer 0 with torch on the cached rows
    kbuf, vbuf, n, ndev = cache.buffers(0)
    assert int(ndev.item()) == n
    q = torch.randn(1, 8, 64, device=DEV, dtype=torch.bfloat16)
    got = ops.decode_attn(q, kbuf, vbuf, n)
    exp = O.decode_attention(q.cpu(), kbuf.cpu(), vbuf.cpu(), n)
    assert torch.allclose(got.cpu().float(), exp, atol=2e-2, rtol=2e-2)


def test_snapkv_method_through_the_same_kernels(patched):
    patched.replace_llama("snapkv")
    model = _model(torch.bfloat16, layers=2).to(DEV)
    patched.set_model(model, _args(method="snapkv", window_size=16, kernel_size=5, max_capacity_prompts=96))
    ids = torch.randint(0, 211, (1, 512), generator=torch.Generator().manual_seed(5)).to(DEV)
    with torch.no_grad():
        out = model(ids, use_cache=True)
    cache = out.past_key_values
    assert cache.get_seq_length(0) == 96 and cache.k[0].shape[1] == 8      # one cache head per query head
    assert torch.isfinite(out.logits).all()


def test_graphed_decode_matches_eager_decode(patched):
    """CUDA-graph replay of the decode step (device-side lengths / positions) == eager loop."""
    model = _model(torch.bfloat16).to(DEV)
    patched.set_model(model, _args())
    ids = torch.randint(0, 211, (2, 700), generator=torch.Generator().manual_seed(6)).to(DEV)
    eager, cache_e = patched.greedy_generate(model, ids, 9)
    model.config.fastkv_max_new_tokens = 16
    model.model.config.fastkv_max_new_tokens = 16
    with torch.no_grad():
        out = model(ids, use_cache=True)
    first = out.logits[:, -1, :].argmax(dim=-1, keepdim=True)
    gd = patched.GraphedDecode(model, out.past_key_values, first, warmup=2)
    rest = gd.generate(6)
    got = torch.cat([first, rest], dim=1)
    assert got.shape == (2, 9)
    assert torch.equal(got, eager)
    assert out.past_key_values.get_seq_length(0) == cache_e.get_seq_length(0) == 128 + 8


def test_gemfilter_two_pass_driver_selects_like_the_oracle():
    """GemFilter (SURVEY.md §8 a13 + §8f rank 4): pass 1 over layers [0, select layer] records the indices with the
    tensor-core GemFilter kernel; they must be the oracle's top-k of the same layer's q/k, pass 2 prefills on exactly
    those tokens in their original order, and decoding continues from that cache."""
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    from fastkv_b200 import patch
    from parity_utils import check_selection
    sel_layer, topk, n = 1, 96, 700
    ids = torch.randint(0, 211, (1, n), generator=torch.Generator().manual_seed(5))
    patch.replace_llama("gemfilter")
    try:
        model = _model(torch.bfloat16).to(DEV)
        patch.set_model(model, _args(method="gemfilter", max_capacity_prompts=topk))
        pred, cache = patch.gemfilter_generate_selection_prefill(ids.to(DEV), None, model, None, select_layer_idx=sel_layer)
        torch.cuda.synchronize()
        attn = model.model.layers[sel_layer].self_attn
        got = attn.indecies
        assert got.shape == (1, 1, topk) and all(layer.self_attn.select_mode is False for layer in model.model.layers)
        assert len(model.model.layers) == 3 and cache.get_seq_length() == topk and pred.shape == (1, 1)
        # oracle on the q/k the selection layer saw (hidden state entering it, its own projections + RoPE)
        with torch.no_grad():
            hs = model.model(ids.to(DEV), output_hidden_states=True, use_cache=False).hidden_states[sel_layer]
            x = model.model.layers[sel_layer].input_layernorm(hs)
            q = attn.q_proj(x).view(1, n, -1, 64).transpose(1, 2)
            k = attn.k_proj(x).view(1, n, -1, 64).transpose(1, 2)
            pos = torch.arange(n, device=DEV)[None]
            cos, sin = model.model.rotary_emb(x, pos)
            q, k = apply_rotary_pos_emb(q, k, cos, sin)
        ref_scores, ref_idx, _ = O.gemfilter_select(q.float().cpu(), k.float().cpu(), q.shape[1] // k.shape[1], topk)
        check_selection(ref_scores[0], ref_idx[0], got[0].cpu(), "gemfilter driver")
        # pass 2 ran on the selected tokens in original order
        new_ids = patch.get_layer_context(model, None, ids[0].to(DEV), sel_layer)
        assert torch.equal(new_ids[0].cpu(), ids[0][got[0, 0].sort().values.cpu()])
        out_ids = patch.gemfilter_generate(model, None, pred, cache, max_gen_len=4)
        assert len(out_ids) == 5 and cache.get_seq_length() == topk + 4
    finally:
        patch.replace_llama("fullkv")


def test_proportional_preset_end_to_end(patched):
    """The reference's canonical preset (scripts/eval_e2e.sh: eviction_mode=proportional, retain_rate, tsp_rate;
    SURVEY.md §8f rank 3) through set_model + a patched prefill: every layer keeps int(q_len * rate) rows with
    the per-layer rate of compress_fastkv (utils.py:43-46, :86-87, :123-124), the TSP layer keeps
    int(L * tsp_rate) tokens, and the selected sets are the glue oracle's."""
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    n, retain, tsp_rate, tsp_idx = 640, 0.1, 0.25, 1
    model = _model(torch.float32)
    ids = torch.randint(0, 211, (1, n), generator=torch.Generator().manual_seed(9))
    knobs = [dict(window_size=8, max_capacity_prompt=128, kernel_size=7, pooling="avgpool", tsp_layer=(i == tsp_idx),
                  tsp_length=256, tsp_rate=tsp_rate, eviction_mode="proportional",
                  retain_rate=retain if i <= tsp_idx else retain / tsp_rate) for i in range(3)]
    patched.replace_llama("fullkv")
    _, ref_caches, traces = G.fastkv_prefill(model, ids, knobs, apply_rotary_pos_emb)
    patched.replace_llama("fastkv")
    gm = _model(torch.float32).to(DEV)
    patched.set_model(gm, _args(eviction_mode="proportional", retain_rate=retain, tsp_rate=tsp_rate, tsp_idx=tsp_idx))
    for layer in gm.model.layers:
        layer.self_attn.kv_cluster.keep_aux = True
    with torch.no_grad():
        out = gm.model(ids.to(DEV), use_cache=True)
    torch.cuda.synchronize()
    cache = out.past_key_values
    t = int(n * tsp_rate)                                           # tokens propagated past the TSP layer
    expect = [int(n * retain), int(n * retain), int(t * (retain / tsp_rate))]
    for i in range(3):
        assert cache.buffers(i)[2] == expect[i] == ref_caches[i][0].shape[2]
        cl = gm.model.layers[i].self_attn.kv_cluster
        assert cl.max_capacity_prompt == expect[i]                  # the attribute is rewritten like the reference's
    tsp = gm.model.layers[tsp_idx].self_attn.tsp_idx
    assert tsp.shape == (1, t) and gm.model.layers[tsp_idx].self_attn.kv_cluster.tsp_length == t
    assert O.near_tie_set_mismatch(traces[tsp_idx].importance[0], ref_caches[tsp_idx][2][0, :-8], tsp[0, :-8].cpu()) == 0


# ---------------------------------------------------------------------------------------------
# Mistral glue (baselines/fastkv/mistral_model.py:100-107, :217-224, :328-330, :351)
# ---------------------------------------------------------------------------------------------
def _mistral(dtype, sliding_window, impl="sdpa", layers=3, seed=0):
    from transformers import MistralConfig, MistralForCausalLM
    cfg = MistralConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=layers, num_attention_heads=8,
                        num_key_value_heads=2, head_dim=64, vocab_size=211, max_position_embeddings=65536,
                        sliding_window=sliding_window, attn_implementation=impl)
    torch.manual_seed(seed)
    return MistralForCausalLM(cfg).eval().to(dtype)


@pytest.mark.parametrize("sliding_window,impl", [(32768, "sdpa"), (200, "sdpa"), (200, "eager")])
def test_mistral_fp32_prefill_matches_glue_oracle(sliding_window, impl):
    """Mistral twin of the Llama test: Ministral's sliding_window=32768 (a no-op at this length, as at 32k) and a window
    shorter than the prompt (index-banded causal attention, flash-attn's ``sliding_window`` semantics)."""
    from transformers.models.mistral.modeling_mistral import apply_rotary_pos_emb
    from fastkv_b200 import patch
    ids = torch.randint(0, 211, (2, 600), generator=torch.Generator().manual_seed(13))
    knobs = [dict(window_size=8, max_capacity_prompt=128, kernel_size=7, pooling="avgpool", tsp_layer=(i == 1), tsp_length=256)
             for i in range(3)]
    patch.replace_mistral("fullkv")
    model = _mistral(torch.float32, sliding_window, impl)
    ref_hidden, ref_caches, traces = G.fastkv_prefill(model, ids, knobs, apply_rotary_pos_emb, sliding_window=sliding_window)
    patch.replace_mistral("fastkv")
    try:
        gm = _mistral(torch.float32, sliding_window, impl).to(DEV)
        patch.set_model(gm, _args())
        for layer in gm.model.layers:
            layer.self_attn.kv_cluster.keep_aux = True
        with torch.no_grad():
            out = gm.model(ids.to(DEV), use_cache=True)
        torch.cuda.synchronize()
        cache = out.past_key_values
        tsp = gm.model.layers[1].self_attn.tsp_idx
        assert tsp is not None and tsp.shape == (2, 256)
        assert torch.equal(gm.model.layers[1].new_position_ids.cpu(), tsp.cpu())       # mistral_model.py:220-222
        for b in range(2):
            assert O.near_tie_set_mismatch(traces[1].importance[b], ref_caches[1][2][b, :-8], tsp[b, :-8].cpu()) == 0
        for i in range(3):
            k_ref, v_ref, _ = ref_caches[i]
            kbuf, vbuf, n, _ = cache.buffers(i)
            assert n == k_ref.shape[2] == 128
            _check_cache_rows(gm.model.layers[i].self_attn.kv_cluster.last_aux["idx"].cpu(), traces[i].idx,
                              kbuf[:, :, :n].cpu(), vbuf[:, :, :n].cpu(), k_ref, v_ref, 8)
        assert torch.equal(tsp.cpu(), ref_caches[1][2])
        assert torch.allclose(out.last_hidden_state.cpu(), ref_hidden, atol=2e-3, rtol=1e-3)
    finally:
        patch.replace_mistral("fullkv")


def test_eager_attention_prefill_is_causal(patched):
    """``_attn_implementation='eager'`` applies no mask unless it is given one: the patched forward must hand it an
    explicit causal mask, i.e. produce the same hidden state and cache as the sdpa path."""
    from transformers import LlamaConfig, LlamaForCausalLM
    outs = []
    for impl in ("sdpa", "eager"):
        cfg = LlamaConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=8,
                          num_key_value_heads=2, head_dim=64, vocab_size=211, max_position_embeddings=4096, attn_implementation=impl)
        torch.manual_seed(0)
        model = LlamaForCausalLM(cfg).eval().to(DEV)
        patched.set_model(model, _args(tsp_idx=0))
        ids = torch.randint(0, 211, (1, 500), generator=torch.Generator().manual_seed(8)).to(DEV)
        with torch.no_grad():
            out = model.model(ids, use_cache=True)
        outs.append((out.last_hidden_state.cpu(), out.past_key_values.buffers(1)[0][:, :, :128].cpu()))
    assert torch.allclose(outs[0][0], outs[1][0], atol=2e-3, rtol=1e-3)
    assert torch.allclose(outs[0][1], outs[1][1], atol=2e-3, rtol=1e-3)


# ---------------------------------------------------------------------------------------------
# SnapKV at model level (baselines/snapkv/llama_model.py:385-404): prefill cache + greedy decode
# ---------------------------------------------------------------------------------------------
def test_snapkv_model_prefill_and_decode_match_glue_oracle(patched):
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    n, c, w = 512, 96, 16
    ids = torch.randint(0, 211, (1, n), generator=torch.Generator().manual_seed(5))
    knobs = [dict(window_size=w, max_capacity_prompt=c, kernel_size=5, pooling="avgpool") for _ in range(2)]
    patched.replace_llama("fullkv")
    model = _model(torch.float32, layers=2)
    ref_hidden, ref_caches, traces = G.snapkv_prefill(model, ids, knobs, apply_rotary_pos_emb)
    first = model.lm_head(ref_hidden)[:, -1].argmax(-1, keepdim=True)
    # the reference's benchmark loop passes no position ids: decode continues from the compact length (e2e.py:174)
    ref_toks, _ = G.decode_steps(model, ref_caches, first, c, 5, apply_rotary_pos_emb, expand_new_kv=True)
    patched.replace_llama("snapkv")
    gm = _model(torch.float32, layers=2).to(DEV)
    patched.set_model(gm, _args(method="snapkv", window_size=w, kernel_size=5, max_capacity_prompts=c))
    for layer in gm.model.layers:
        layer.self_attn.kv_cluster.keep_aux = True
    toks, cache = patched.greedy_generate(gm, ids.to(DEV), 6)
    torch.cuda.synchronize()
    for i in range(2):
        kbuf, vbuf, nrows, _ = cache.buffers(i)
        assert kbuf.shape[1] == 8 and nrows == c + 5                      # one cache head per QUERY head, every row written
        _check_cache_rows(gm.model.layers[i].self_attn.kv_cluster.last_aux["idx"].cpu(), traces[i].idx,
                          kbuf[:, :, :c].cpu(), vbuf[:, :, :c].cpu(), ref_caches[i][0], ref_caches[i][1], w)
        # decode rows: all 8 heads hold the new token's K (head h = kv head h // 4), none is left uninitialised
        new_rows = kbuf[:, :, c:nrows].cpu()
        assert torch.isfinite(new_rows).all()
        assert torch.equal(new_rows[:, 0], new_rows[:, 3]) and torch.equal(new_rows[:, 4], new_rows[:, 7])
        assert not torch.equal(new_rows[:, 0], new_rows[:, 4])
    assert toks[0, 0].item() == first.item()
    assert torch.equal(toks[:, 1:].cpu(), ref_toks)


# ---------------------------------------------------------------------------------------------
# generate() path: positions continue from the TRUE prompt length (baselines/monkeypatch.py:280-288)
# ---------------------------------------------------------------------------------------------
def test_generate_positions_continue_from_the_true_prompt_length(patched):
    from transformers.models.llama.modeling_llama import apply_rotary_pos_emb
    n, new = 600, 8
    ids = torch.randint(0, 211, (1, n), generator=torch.Generator().manual_seed(21))
    knobs = [dict(window_size=8, max_capacity_prompt=128, kernel_size=7, pooling="avgpool", tsp_layer=(i == 1), tsp_length=256)
             for i in range(3)]
    patched.replace_llama("fullkv")
    model = _model(torch.float32)
    ref_hidden, ref_caches, _ = G.fastkv_prefill(model, ids, knobs, apply_rotary_pos_emb)
    first = model.lm_head(ref_hidden)[:, -1].argmax(-1, keepdim=True)
    ref_toks, ref_pos = G.decode_steps(model, [c[:2] for c in ref_caches], first, n, new - 1, apply_rotary_pos_emb)
    assert ref_pos == list(range(n, n + new - 1))
    patched.replace_llama("fastkv")
    gm = _model(torch.float32).to(DEV)
    gm.generation_config.pad_token_id = 0
    patched.set_model(gm, _args())
    seen = []
    inner = gm.model.forward

    def spy(*a, **kw):
        seen.append(kw["position_ids"].flatten().tolist() if kw.get("position_ids") is not None else None)
        return inner(*a, **kw)

    gm.model.forward = spy
    out = gm.generate(ids.to(DEV), attention_mask=torch.ones_like(ids).to(DEV), max_new_tokens=new, do_sample=False)
    assert seen[0] == list(range(n))
    assert [p[0] for p in seen[1:]] == list(range(n, n + new - 1))          # NOT 128 + t (the compact length)
    assert torch.equal(out[0, n:].cpu(), torch.cat([first[0], ref_toks[0]]))
