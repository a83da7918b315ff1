"""Pins the CPU oracle (oracle/fastkv_oracle.py) to outputs of the unmodified reference.

The fixtures under tests/golden/ were produced by tests/golden/make_golden.py, which imports the
reference's FastKVCluster / SnapKVCluster / find_context from /root/reference and runs them on CPU.
"""
import pytest
import torch

from oracle import fastkv_oracle as O
from conftest import golden_names, load_golden, full_query


@pytest.fixture(autouse=True)
def _single_thread():
    n = torch.get_num_threads()
    torch.set_num_threads(1)          # same reduction order as the fixture generator
    yield
    torch.set_num_threads(n)


@pytest.mark.parametrize("name", golden_names("fastkv"))
def test_fastkv_restatement_matches_reference(name):
    fx = load_golden(name)
    l = fx["shape"]["L"]
    dt = torch.float32 if fx["run_dtype"] == "fp32" else torch.bfloat16
    g = fx["shape"]["Hq"] // fx["shape"]["Hkv"]
    k, v, q = fx["k"].to(dt), fx["v"].to(dt), full_query(fx["q_window"], l).to(dt)
    tr = O.fastkv_update_kv(k, q, v, g, **fx["knobs"])
    assert torch.equal(tr.key.to(torch.bfloat16), fx["key_out"])
    assert torch.equal(tr.value.to(torch.bfloat16), fx["value_out"])
    if fx["tsp_idx"] is None:
        assert tr.tsp_idx is None
    else:
        assert torch.equal(tr.tsp_idx, fx["tsp_idx"])
    assert tr.capacity == fx["capacity_after"]
    if not tr.skipped:
        # reference leaves self.tsp_length rewritten only on TSP layers in proportional mode
        assert tr.tsp_length == fx["tsp_length_after"]


@pytest.mark.parametrize("name", golden_names("fastkv"))
def test_fastkv_deterministic_ties_selects_same_scores(name):
    """(score desc, index asc) order picks the same score multiset as the reference's topk."""
    fx = load_golden(name)
    if fx["run_dtype"] != "fp32":
        pytest.skip("parity protocol is fp32")
    l = fx["shape"]["L"]
    g = fx["shape"]["Hq"] // fx["shape"]["Hkv"]
    k, v, q = fx["k"].float(), fx["v"].float(), full_query(fx["q_window"], l).float()
    a = O.fastkv_update_kv(k, q, v, g, **fx["knobs"])
    b = O.fastkv_update_kv(k, q, v, g, deterministic_ties=True, **fx["knobs"])
    if a.skipped:
        assert b.skipped
        return
    sa = torch.gather(a.score, -1, a.idx)
    sb = torch.gather(b.score, -1, b.idx)
    assert torch.equal(sa, sb)                      # both are score-descending -> same sequences
    if a.tsp_idx is not None:
        ia = torch.gather(a.importance, -1, a.tsp_idx[:, : -fx["knobs"]["window_size"]]).sort(-1).values
        ib = torch.gather(b.importance, -1, b.tsp_idx[:, : -fx["knobs"]["window_size"]]).sort(-1).values
        assert torch.equal(ia, ib)
        # max-pooling makes neighbouring keys share one value, so exact ties are the norm there
        if fx["dist"] != "ties" and fx["knobs"]["pooling"] == "avgpool":
            assert torch.equal(a.tsp_idx, b.tsp_idx)
            assert torch.equal(a.idx, b.idx)


@pytest.mark.parametrize("name", golden_names("snapkv"))
def test_snapkv_restatement_matches_reference(name):
    fx = load_golden(name)
    l = fx["shape"]["L"]
    g = fx["shape"]["Hq"] // fx["shape"]["Hkv"]
    k = O.expand_kv_heads(fx["k"].float(), g)
    v = O.expand_kv_heads(fx["v"].float(), g)
    tr = O.snapkv_update_kv(k, full_query(fx["q_window"], l).float(), v, **fx["knobs"])
    assert torch.equal(tr.key.to(torch.bfloat16), fx["key_out"])
    assert torch.equal(tr.value.to(torch.bfloat16), fx["value_out"])


@pytest.mark.parametrize("name", golden_names("gemfilter"))
def test_gemfilter_restatement_matches_reference(name):
    fx = load_golden(name)
    n = fx["shape"]["L"]
    g = fx["shape"]["Hq"] // fx["shape"]["Hkv"]
    _, idx, srt = O.gemfilter_select(full_query(fx["q_window"], n).float(), fx["k"].float(), g, fx["topk"])
    assert torch.equal(idx, fx["indices"])
    assert torch.equal(srt, fx["sorted_indices"])


def test_window_bias_matches_reference_mask():
    # baselines/fastkv/utils.py:95-97: mask[r, c] = 0 iff c < r + 1
    w = 8
    ref = torch.full((w, w), torch.finfo(torch.float32).min)
    cond = torch.arange(w)
    ref.masked_fill_(cond < (cond + 1).view(w, 1), 0)
    assert torch.equal(O.window_causal_bias(w, torch.float32), ref)


def test_tsp_propagate_and_decode_attention_shapes():
    h = torch.randn(2, 10, 16)
    pos = torch.arange(10).repeat(2, 1)
    idx = torch.tensor([[0, 3, 9], [1, 2, 8]])
    nh, npos = O.tsp_propagate(h, pos, idx)
    assert torch.equal(nh[1, 2], h[1, 8]) and torch.equal(npos, idx)
    q = torch.randn(2, 8, 32)
    kc, vc = torch.randn(2, 2, 20, 32), torch.randn(2, 2, 20, 32)
    out = O.decode_attention(q, kc, vc, 13)
    ref = torch.nn.functional.scaled_dot_product_attention(
        q[:, :, None, :], O.expand_kv_heads(kc[:, :, :13], 4), O.expand_kv_heads(vc[:, :, :13], 4))[:, :, 0]
    assert torch.allclose(out, ref, atol=1e-5)


def test_near_tie_rule():
    s = torch.tensor([5.0, 4.0, 3.0, 3.0 * (1 + 5e-6), 1.0])
    ref = torch.tensor([0, 1, 3])
    assert O.near_tie_set_mismatch(s, ref, torch.tensor([0, 1, 2])) == 0      # swap inside tolerance
    assert O.near_tie_set_mismatch(s, ref, torch.tensor([0, 1, 4])) > 0       # real miss
    assert O.near_tie_set_mismatch(s, ref, torch.tensor([0, 0, 1])) > 0       # duplicates
