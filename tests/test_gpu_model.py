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
        idx = gm.model.layers[i].self_attn.kv_cluster.last_aux["idx"].cpu()
        if torch.equal(idx, traces[i].idx):
            assert torch.allclose(kbuf[:, :, :n].cpu(), k_ref, atol=2e-4, rtol=1e-4)
            assert torch.allclose(vbuf[:, :, :n].cpu(), v_ref, atol=2e-4, rtol=1e-4)
    if torch.equal(tsp.cpu(), ref_caches[1][2]):
        assert torch.allclose(out.last_hidden_state.cpu(), ref_hidden, atol=2e-3, rtol=1e-3)
    assert out.last_hidden_state.shape == (2, 1, 512)


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
