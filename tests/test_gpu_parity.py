"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle.

Protocol (SURVEY.md §8c / BASELINE.json north_star): inputs hold bf16-representable values; the
oracle sees them as fp32, the kernels as bf16 (tensor path) or fp32 (simple path).
  * selected index sets equal the oracle's except where the oracle's own score gap to its k-th
    score is < 1e-5 relative (near ties);
  * compacted K/V and propagated hidden rows are bit-exact given OUR indices;
  * scores within rtol 2e-5 of the oracle's fp32 scores (fp32 accumulation-order noise).
"""
import math

import pytest
import torch

from oracle import fastkv_oracle as O
from conftest import golden_names, load_golden, full_query
from parity_utils import (make_inputs, check_selection, check_scores_close, check_compaction_exact,
                          check_topk_consistent, check_tsp_consistent)

pytestmark = pytest.mark.gpu
SCORE_RTOL = 2e-5
DEV = "cuda"


def _lib():
    from fastkv_b200 import _lib as L
    return L


def _run_cluster(fx_knobs, k, q, v, groups, impl):
    from fastkv_b200 import FastKVCluster
    cl = FastKVCluster(**fx_knobs)
    cl.keep_aux = True
    cl.score_impl = impl
    out = cl.update_kv(k, q, v, None, groups, 0)
    torch.cuda.synchronize()
    return cl, out


def _impls_for(dtype):
    L = _lib()
    lib = L.load()
    impls = [L.IMPL_SIMPLE]
    # the tensor-core path exists for 16-bit inputs only; exercise it whenever AUTO would pick it
    if dtype != torch.float32:
        impls.append(L.IMPL_AUTO)
    return impls


# ---------------------------------------------------------------------------------------------
# golden fixtures (outputs of the unmodified reference)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("name", [n for n in golden_names("fastkv") if n != "fastkv_bf16_run"])
def test_fastkv_matches_reference_golden(name, dtype):
    fx = load_golden(name)
    sh, knobs = fx["shape"], fx["knobs"]
    l, g, w = sh["L"], sh["Hq"] // sh["Hkv"], knobs["window_size"]
    k32, v32, q32 = fx["k"].float(), fx["v"].float(), full_query(fx["q_window"], l).float()
    ref = O.fastkv_update_kv(k32, q32, v32, g, **knobs)                  # == reference (pinned on CPU)
    for impl in _impls_for(dtype):
        cl, (ko, vo, tsp) = _run_cluster(knobs, k32.to(DEV, dtype), q32.to(DEV, dtype), v32.to(DEV, dtype), g, impl)
        assert cl.max_capacity_prompt == fx["capacity_after"]
        if ref.skipped:
            assert tsp is None and ko.shape == k32.shape
            continue
        aux = cl.last_aux
        c = ref.capacity
        assert ko.shape == (sh["B"], sh["Hkv"], c, sh["D"]) and ko.dtype == dtype
        check_scores_close(ref.score, aux["score"], SCORE_RTOL, f"{name}/{impl}")
        if fx["dist"] == "ties" or knobs["pooling"] == "maxpool":
            # exact ties: tie order is unspecified upstream -> compare the selected score multisets
            ours = torch.gather(ref.score, -1, aux["idx"].cpu()).sort(-1).values
            theirs = torch.gather(ref.score, -1, ref.idx).sort(-1).values
            assert torch.allclose(ours, theirs, rtol=1e-5, atol=0)
        else:
            flat = ref.score.reshape(-1, ref.score.shape[-1])
            check_selection(flat, ref.idx.reshape(flat.shape[0], -1), aux["idx"].reshape(flat.shape[0], -1), name)
            # golden K/V: the reference's rows at every position where the score-descending orders agree (near-tie
            # swaps are the only tolerated difference), and the window rows always
            same = aux["idx"].cpu() == ref.idx
            assert same.float().mean().item() > 0.97, same.float().mean().item()
            nsel = ref.idx.shape[-1]
            assert torch.equal(ko.cpu().to(torch.bfloat16)[:, :, :nsel][same], fx["key_out"][:, :, :nsel][same])
            assert torch.equal(vo.cpu().to(torch.bfloat16)[:, :, :nsel][same], fx["value_out"][:, :, :nsel][same])
            assert torch.equal(ko.cpu().to(torch.bfloat16)[:, :, nsel:], fx["key_out"][:, :, nsel:])
            assert torch.equal(vo.cpu().to(torch.bfloat16)[:, :, nsel:], fx["value_out"][:, :, nsel:])
        check_topk_consistent(aux["score"], aux["idx"], name)
        check_compaction_exact(k32.to(dtype), v32.to(dtype), aux["idx"], ko, vo, w)
        if ref.tsp_idx is None:
            assert tsp is None
        else:
            assert tsp.shape == ref.tsp_idx.shape and tsp.dtype == torch.int64
            check_tsp_consistent(aux["importance"], tsp, w, name)
            if fx["dist"] == "ties" or knobs["pooling"] == "maxpool":
                ours = torch.gather(ref.importance, -1, tsp.cpu()[:, :-w]).sort(-1).values
                theirs = torch.gather(ref.importance, -1, ref.tsp_idx[:, :-w]).sort(-1).values
                assert torch.allclose(ours, theirs, rtol=1e-5, atol=0)
            else:
                check_selection(ref.importance, ref.tsp_idx[:, :-w], tsp[:, :-w], name + "/tsp")
                assert torch.equal(ref.tsp_idx, fx["tsp_idx"])               # the oracle is the pinned reference
                for bi in range(tsp.shape[0]):                               # only near-tie members may differ
                    assert len(set(tsp[bi].tolist()) ^ set(fx["tsp_idx"][bi].tolist())) <= max(2, tsp.shape[1] // 100)
                assert torch.equal(tsp.cpu()[:, -w:], fx["tsp_idx"][:, -w:])


def test_bf16_run_of_the_reference_is_within_bf16_tolerance():
    """Informational leg of the protocol: the reference run in bf16 rounds probabilities to bf16
    (utils.py:103); our fp32 scores must agree with it to bf16 precision on the kept mass."""
    fx = load_golden("fastkv_bf16_run")
    sh, knobs = fx["shape"], fx["knobs"]
    l, g = sh["L"], sh["Hq"] // sh["Hkv"]
    k, v, q = fx["k"], fx["v"], full_query(fx["q_window"], l)
    ref = O.fastkv_update_kv(k, q, v, g, **knobs)                         # bf16 arithmetic like the reference
    cl, _ = _run_cluster(knobs, k.to(DEV), q.to(DEV), v.to(DEV), g, _lib().IMPL_AUTO)
    got = cl.last_aux["score"].cpu()
    rel = ((got - ref.score.float()).abs() / ref.score.float().abs().clamp_min(1e-6)).median().item()
    assert rel < 2e-2, rel                                                # bf16 has 8 bits of mantissa
    overlap = len(set(cl.last_aux["idx"][0, 0].tolist()) & set(ref.idx[0, 0].tolist())) / ref.idx.shape[-1]
    assert overlap > 0.8, overlap


@pytest.mark.parametrize("name", golden_names("snapkv"))
@pytest.mark.parametrize("expanded", [True, False])
def test_snapkv_matches_reference_golden(name, expanded):
    from fastkv_b200 import SnapKVCluster
    fx = load_golden(name)
    sh, knobs = fx["shape"], fx["knobs"]
    l, g, w = sh["L"], sh["Hq"] // sh["Hkv"], knobs["window_size"]
    k32, v32, q32 = fx["k"].float(), fx["v"].float(), full_query(fx["q_window"], l).float()
    ref = O.snapkv_update_kv(O.expand_kv_heads(k32, g), q32, O.expand_kv_heads(v32, g), **knobs)
    kin, vin = (O.expand_kv_heads(k32, g), O.expand_kv_heads(v32, g)) if expanded else (k32, v32)
    cl = SnapKVCluster(**knobs)
    cl.keep_aux = True
    ko, vo = cl.update_kv(kin.to(DEV, torch.bfloat16), q32.to(DEV, torch.bfloat16), vin.to(DEV, torch.bfloat16), None, g)
    torch.cuda.synchronize()
    aux = cl.last_aux
    assert ko.shape == (sh["B"], sh["Hq"], ref.capacity, sh["D"])
    check_scores_close(ref.score, aux["score"], SCORE_RTOL, name)
    check_topk_consistent(aux["score"], aux["idx"], name)
    check_compaction_exact(k32.bfloat16(), v32.bfloat16(), aux["idx"], ko, vo, w, head_div=g)
    if knobs["pooling"] == "avgpool":
        flat = ref.score.reshape(-1, ref.score.shape[-1])
        check_selection(flat, ref.idx.reshape(flat.shape[0], -1), aux["idx"].reshape(flat.shape[0], -1), name)
        same = aux["idx"].cpu() == ref.idx
        assert same.float().mean().item() > 0.97, same.float().mean().item()
        nsel = ref.idx.shape[-1]
        assert torch.equal(ko.cpu()[:, :, :nsel][same], fx["key_out"][:, :, :nsel][same])
        assert torch.equal(vo.cpu()[:, :, :nsel][same], fx["value_out"][:, :, :nsel][same])
        assert torch.equal(ko.cpu()[:, :, nsel:], fx["key_out"][:, :, nsel:]) and torch.equal(vo.cpu()[:, :, nsel:], fx["value_out"][:, :, nsel:])


@pytest.mark.parametrize("name", golden_names("gemfilter"))
def test_gemfilter_matches_reference_golden(name):
    import types
    from fastkv_b200 import find_context
    fx = load_golden(name)
    sh = fx["shape"]
    n, g = sh["L"], sh["Hq"] // sh["Hkv"]
    q32, k32 = full_query(fx["q_window"], n).float(), fx["k"].float()
    ref_scores, ref_idx, ref_sorted = O.gemfilter_select(q32, k32, g, fx["topk"])
    holder = types.SimpleNamespace(eviction_mode="constant", retain_rate=0.1, topk=fx["topk"], indecies=None,
                                   layer_idx=13, select_layer_idx=13, num_key_value_groups=g)
    find_context(holder, q32.to(DEV, torch.bfloat16), k32.to(DEV, torch.bfloat16))
    torch.cuda.synchronize()
    got = holder.indecies
    assert got.shape == ref_idx.shape
    check_selection(ref_scores[0], ref_idx[0], got[0], name)
    assert torch.equal(ref_sorted, fx["sorted_indices"])                     # the oracle is the pinned reference
    assert len(set(got[0, 0].tolist()) ^ set(fx["sorted_indices"].tolist())) <= max(2, fx["topk"] // 100)


@pytest.mark.parametrize("b,hq,hkv,l,d,ks", [(1, 32, 8, 5000, 128, 5), (2, 16, 4, 1234, 128, 5), (1, 8, 8, 700, 64, 5),
                                             (1, 32, 8, 32768, 128, 5), (3, 16, 2, 300, 128, 7)])
def test_gemfilter_tensor_path_matches_fp32_path(b, hq, hkv, l, d, ks):
    """tcgen05 GemFilter scores (sum over kv heads inside the TMEM accumulator) vs the fp32 CUDA-core path."""
    from fastkv_b200 import ops
    L = _lib()
    q, k, _ = [x.to(DEV) for x in make_inputs(l + hq, b, hq, hkv, l, d)]
    ref = ops.score(q, k, window_size=1, kernel_size=ks, mode=L.MODE_GEMFILTER, impl=L.IMPL_SIMPLE).double()
    got = ops.score(q, k, window_size=1, kernel_size=ks, mode=L.MODE_GEMFILTER, impl=L.IMPL_TENSOR).double()
    torch.cuda.synchronize()
    assert got.shape == ref.shape == (b, 1, l)
    # head-summed dot products cancel: compare against the magnitude of the terms, not of the (small) sum
    scale = ref.abs().amax(-1, keepdim=True)
    assert ((ref - got).abs() / scale).max().item() < 2e-5


# ---------------------------------------------------------------------------------------------
# seeded inputs against the oracle at sizes it finishes in seconds
# ---------------------------------------------------------------------------------------------
CASES = [
    # b, hq, hkv, L, D, W, ks, pooling, C, T, dist
    (1, 32, 8, 8192, 128, 8, 7, "avgpool", 512, 2048, "gauss"),        # BASELINE config 1 shape
    (1, 32, 8, 4099, 128, 8, 7, "avgpool", 256, 1024, "peaky"),        # ragged length
    (2, 16, 4, 2048, 128, 16, 5, "avgpool", 128, 512, "gauss"),        # batch 2, W=16
    (1, 32, 8, 3000, 128, 32, 13, "avgpool", 300, 600, "peaky"),       # W=32, kernel 13
    (1, 8, 8, 1500, 64, 8, 7, "avgpool", 100, 200, "gauss"),           # MHA (G=1), D=64
    (1, 16, 2, 1024, 128, 8, 7, "maxpool", 128, 256, "gauss"),         # G=8, maxpool
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
@pytest.mark.parametrize("case", CASES, ids=[f"L{c[3]}_W{c[5]}_k{c[6]}_{c[7]}_C{c[8]}" for c in CASES])
def test_fastkv_seeded_vs_oracle(case, dtype):
    b, hq, hkv, l, d, w, ks, pooling, c, t, dist = case
    g = hq // hkv
    q, k, v = make_inputs(100 + l, b, hq, hkv, l, d, dtype=torch.float32, dist=dist)
    if dtype == torch.float16:                       # keep values fp16-representable too
        q, k, v = [x.clamp(-8, 8) for x in (q, k, v)]
    knobs = dict(window_size=w, max_capacity_prompt=c, kernel_size=ks, pooling=pooling, tsp_layer=True, tsp_length=t)
    ref = O.fastkv_update_kv(k, q, v, g, **knobs)
    for impl in _impls_for(dtype):
        cl, (ko, vo, tsp) = _run_cluster(knobs, k.to(DEV, dtype), q.to(DEV, dtype), v.to(DEV, dtype), g, impl)
        aux = cl.last_aux
        err = check_scores_close(ref.score, aux["score"], SCORE_RTOL, f"impl {impl}")
        check_topk_consistent(aux["score"], aux["idx"])
        check_tsp_consistent(aux["importance"], tsp, w)
        check_compaction_exact(k.to(dtype), v.to(dtype), aux["idx"], ko, vo, w)
        assert not ko.data_ptr() == k.data_ptr()
        if pooling == "avgpool":
            flat = ref.score.reshape(-1, ref.score.shape[-1])
            check_selection(flat, ref.idx.reshape(flat.shape[0], -1), aux["idx"].reshape(flat.shape[0], -1), f"impl {impl}")
            check_selection(ref.importance, ref.tsp_idx[:, :-w], tsp[:, :-w], "tsp")
        else:
            ours = torch.gather(ref.score, -1, aux["idx"].cpu()).sort(-1).values
            theirs = torch.gather(ref.score, -1, ref.idx).sort(-1).values
            assert torch.allclose(ours, theirs, rtol=1e-5, atol=0)


SHORT_CASES = [
    # b, hq, hkv, L, D, ks, pooling, C, dist   (window 8, no TSP at this layer, L <= 3904: the one-kernel cluster path)
    (1, 32, 8, 2048, 128, 7, "avgpool", 2048, "gauss"),       # the post-TSP layer of BASELINE configs[1]: L == C, full re-order
    (1, 32, 8, 2048, 128, 7, "avgpool", 512, "peaky"),
    (2, 16, 4, 3900, 128, 7, "avgpool", 256, "gauss"),        # 32 tiles: 4 per CTA, batch 2
    (1, 8, 8, 300, 64, 5, "avgpool", 64, "gauss"),            # MHA (G=1), D=64, fewer tiles than CTAs in the cluster
    (3, 8, 4, 1000, 128, 13, "maxpool", 128, "gauss"),        # G=2, maxpool, kernel 13
    (1, 32, 8, 130, 128, 7, "avgpool", 40, "gauss"),          # two tiles only
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("case", SHORT_CASES)
def test_short_context_fused_kernel_vs_oracle(case, dtype, monkeypatch):
    """Opt-in path (FKV_SMALL_FUSE=1): non-TSP layers with a short context run score + top-k + compaction as
    ONE cluster kernel (small_compress_kernel); same checks as the three-kernel path."""
    monkeypatch.setenv("FKV_SMALL_FUSE", "1")
    before = _lib().load().fkv_launch_count()
    b, hq, hkv, l, d, ks, pooling, c, dist = case
    g, w = hq // hkv, 8
    q, k, v = make_inputs(500 + l + c, b, hq, hkv, l, d, dtype=torch.float32, dist=dist)
    if dtype == torch.float16:
        q, k, v = [x.clamp(-8, 8) for x in (q, k, v)]
    knobs = dict(window_size=w, max_capacity_prompt=c, kernel_size=ks, pooling=pooling, tsp_layer=False, tsp_length=2048)
    ref = O.fastkv_update_kv(k, q, v, g, **knobs)
    cl, (ko, vo, tsp) = _run_cluster(knobs, k.to(DEV, dtype), q.to(DEV, dtype), v.to(DEV, dtype), g, _lib().IMPL_AUTO)
    aux = cl.last_aux
    assert _lib().load().fkv_launch_count() - before == 1, "the fused kernel did not run"
    assert tsp is None and ko.shape == (b, hkv, c, d)
    check_scores_close(ref.score, aux["score"], SCORE_RTOL, "fused")
    check_topk_consistent(aux["score"], aux["idx"])
    check_compaction_exact(k.to(dtype), v.to(dtype), aux["idx"], ko, vo, w)
    if pooling == "avgpool":
        flat = ref.score.reshape(-1, ref.score.shape[-1])
        check_selection(flat, ref.idx.reshape(flat.shape[0], -1), aux["idx"].reshape(flat.shape[0], -1), "fused")
    else:
        ours = torch.gather(ref.score, -1, aux["idx"].cpu()).sort(-1).values
        theirs = torch.gather(ref.score, -1, ref.idx).sort(-1).values
        assert torch.allclose(ours, theirs, rtol=1e-5, atol=0)
    # and the three-kernel path on the same inputs selects the same rows wherever the gaps are clear
    cs, (ks_, vs_, _) = _run_cluster(knobs, k.to(DEV, dtype), q.to(DEV, dtype), v.to(DEV, dtype), g, _lib().IMPL_SIMPLE)
    sa = cs.last_aux["score"].reshape(-1, cs.last_aux["score"].shape[-1])
    if pooling == "avgpool":
        check_selection(sa, cs.last_aux["idx"].reshape(sa.shape[0], -1), aux["idx"].reshape(sa.shape[0], -1), "fused vs simple")


def test_scratch_users_leave_the_scoring_header_alone():
    """The scoring kernel keeps its departure counter + launch epoch in the first 256 bytes of the shared
    per-stream workspace; top-k (generic path, global sort scratch) and decode attention use the same buffer
    as plain scratch and must not clobber it (a clobbered epoch makes stale partial records look valid)."""
    from fastkv_b200 import ops
    L = _lib()
    q, k, v = [x.to(DEV, torch.bfloat16) for x in make_inputs(77, 1, 32, 8, 9000, 128)]
    run = lambda: ops.compress(q, k, v, window_size=8, kernel_size=7, pooling="avgpool", capacity=512, tsp_len=1024, return_aux=True)
    k0, v0, t0, aux0 = run()
    big = torch.rand(2, 60000, device=DEV)
    ops.topk(big, 30000)                                            # k > 2048: generic path, sorts in the workspace
    qd = torch.randn(2, 32, 128, device=DEV, dtype=torch.bfloat16)
    kc = torch.randn(2, 8, 4096, 128, device=DEV, dtype=torch.bfloat16)
    ops.decode_attn(qd, kc, kc.clone(), 4000)                       # split-K partials in the workspace
    for _ in range(3):
        k1, v1, t1, aux1 = run()
        assert torch.equal(k0, k1) and torch.equal(v0, v1) and torch.equal(t0, t1)
        assert torch.equal(aux0["score"], aux1["score"])


_UNINIT_WS = r"""
import sys, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from fastkv_b200 import ops, _lib as L
from parity_utils import make_inputs
buf = torch.full((1 << 24,), 0xAB, dtype=torch.uint8, device="cuda")      # never passed through fkv_workspace_init
L.workspace = lambda device, nbytes: buf
q, k, v = [x.cuda().bfloat16() for x in make_inputs(1, 1, 8, 2, 2000, 128)]
try:
    ops.score(q, k, window_size=8, kernel_size=7, impl=L.IMPL_TENSOR)
    torch.cuda.synchronize()
    print("NO ERROR")
except Exception as e:
    print("FAILED LOUDLY:", type(e).__name__)
"""


def test_uninitialised_workspace_fails_loudly():
    """fkv_score on a workspace that skipped fkv_workspace_init must not compute on garbage header state."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _UNINIT_WS, root], capture_output=True, text=True, timeout=170)
    assert "FAILED LOUDLY" in r.stdout and "NO ERROR" not in r.stdout, r.stdout[-1500:] + r.stderr[-1500:]


@pytest.mark.parametrize("l,c,t", [(9000, 512, 1024), (2048, 2048, 0), (33000, 2048, 2048), (5000, 256, 0)])
def test_overlapped_select_equals_the_serial_call(l, c, t):
    """ops.SelectOverlap: scoring on the caller's stream, select + compact on a side stream (two histogram slots, SM
    partition) — eagerly and captured in a CUDA graph — must give exactly the serial results."""
    from fastkv_b200 import ops
    q, k, v = [x.to(DEV) for x in make_inputs(l, 1, 32, 8, l, 128)]
    kw = dict(window_size=8, kernel_size=7, pooling="avgpool", capacity=c, tsp_len=t, return_aux=True)
    ref = ops.compress(q, k, v, **kw)

    def same_as_serial(got):
        assert torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1]) and torch.equal(ref[3]["idx"], got[3]["idx"])
        assert (ref[2] is None and got[2] is None) or torch.equal(ref[2], got[2])

    def consistent(got):
        # a capped scoring grid splits the keys over fewer CTAs: the fp32 partial sums merge in another order, so the
        # scores agree to rounding (not bit for bit) and everything downstream must be exact GIVEN those scores
        check_scores_close(ref[3]["score"], got[3]["score"], 1e-5, "partitioned")
        check_topk_consistent(got[3]["score"], got[3]["idx"], "partitioned")
        check_compaction_exact(k, v, got[3]["idx"], got[0], got[1], 8)
        if t:
            check_tsp_consistent(got[3]["importance"], got[2], 8, "partitioned")

    for partition in (False, True):
        ov = ops.SelectOverlap(partition=partition)
        outs = [ops.compress(q, k, v, overlap=ov, **kw) for _ in range(3)]        # slots 1, 2, 1: the third call waits for the first
        ov.join()
        torch.cuda.synchronize()
        for got in outs:
            (consistent if partition else same_as_serial)(got)
        assert all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][3]["idx"], o[3]["idx"]) for o in outs[1:])   # deterministic
    ov = ops.SelectOverlap(partition=False)
    ko, vo = torch.empty_like(ref[0]), torch.empty_like(ref[1])
    ops.compress(q, k, v, overlap=ov, k_out=ko, v_out=vo, **kw); ov.join(); torch.cuda.synchronize()      # warm-up
    ko.zero_(); vo.zero_()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(2):
            ops.compress(q, k, v, overlap=ov, k_out=ko, v_out=vo, **kw)
        ov.join()
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(ko, ref[0]) and torch.equal(vo, ref[1])
    from fastkv_b200 import _lib as LL
    assert LL.workspace_status(torch.device(DEV, torch.cuda.current_device())) == 0      # no barrier time-out was flagged


def test_select_compact_entry_point_builds_its_own_histogram():
    """fkv_select_compact with have_hist = 0 (what the fp32 scoring path gets): the CTAs build the row histograms from
    their slices first.  Long rows, bf16 K/V, scores from the tensor-core kernel WITHOUT a histogram slot."""
    import ctypes as C
    from fastkv_b200 import ops
    L = _lib()
    lib = L.load()
    b, hq, hkv, l, d, w, cap = 2, 16, 4, 12000, 128, 8, 520
    q, k, v = [x.to(DEV) for x in make_inputs(5, b, hq, hkv, l, d)]
    score = ops.score(q, k, window_size=w, kernel_size=7, pooling="avgpool")             # [B, Hkv, P], no histogram
    n_sel, p = cap - w, l - w
    idx = torch.empty(b, hkv, n_sel, dtype=torch.int64, device=DEV)
    ko = torch.empty(b, hkv, cap, d, dtype=k.dtype, device=DEV)
    vo = torch.empty_like(ko)
    ws = L.workspace(q.device, lib.fkv_workspace_bytes(L.MODE_FASTKV, b, hq, hkv, l, w, cap, 0))
    sc = L.SelectCompactParams()
    sc.struct_size = C.sizeof(L.SelectCompactParams)
    sc.hist_slot, sc.have_hist = 2, 0
    sc.score, sc.idx_out = score.data_ptr(), idx.data_ptr()
    g = sc.gather
    g.struct_size = C.sizeof(L.GatherKvParams)
    g.dtype, g.B, g.L, g.D, g.window_size, g.n_sel, g.h0, g.h1, g.head_div = L.dtype_code(k.dtype), b, l, d, w, n_sel, 0, hkv, 1
    g.k, g.v, g.k_out, g.v_out = L.view4(k), L.view4(v), L.view4(ko), L.view4(vo)
    sc.workspace, sc.workspace_bytes = ws.data_ptr(), ws.numel()
    for _ in range(2):                                                                   # twice: the tables must come back clean
        L.check(lib.fkv_select_compact(C.byref(sc), L.stream_ptr(q.device)))
    torch.cuda.synchronize()
    check_topk_consistent(score, idx, "select_compact, own histogram")
    check_compaction_exact(k, v, idx, ko, vo, w)
    assert L.workspace_status(q.device) == 0


def test_strided_blhd_views_equal_contiguous_inputs():
    """update_kv gets [B,H,L,D] views of [B,L,H,D] memory; results must not depend on the layout."""
    knobs = dict(window_size=8, max_capacity_prompt=128, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=256)
    q, k, v = [x.to(DEV) for x in make_inputs(7, 2, 32, 8, 1536, 128, layout="blhd")]
    assert not k.is_contiguous()
    _, a = _run_cluster(knobs, k, q, v, 4, _lib().IMPL_AUTO)
    _, b = _run_cluster(knobs, k.contiguous(), q.contiguous(), v.contiguous(), 4, _lib().IMPL_AUTO)
    for x, y in zip(a, b):
        assert torch.equal(x, y)


def test_simple_and_auto_paths_select_identically_on_clear_gaps():
    L = _lib()
    knobs = dict(window_size=8, max_capacity_prompt=256, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=512)
    q, k, v = [x.to(DEV) for x in make_inputs(11, 1, 32, 8, 4096, 128, dist="peaky")]
    ca, _ = _run_cluster(knobs, k, q, v, 4, L.IMPL_SIMPLE)
    cb, _ = _run_cluster(knobs, k, q, v, 4, L.IMPL_AUTO)
    check_scores_close(ca.last_aux["score"], cb.last_aux["score"], SCORE_RTOL)
    sa = ca.last_aux["score"].reshape(-1, ca.last_aux["score"].shape[-1])
    check_selection(sa, ca.last_aux["idx"].reshape(sa.shape[0], -1), cb.last_aux["idx"].reshape(sa.shape[0], -1))


_PHASE_CHECK = r"""
import sys, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from fastkv_b200 import ops, _lib as L
from parity_utils import make_inputs
for (b, hq, hkv, l, w, ks, pool) in [(2, 16, 4, 1500, 8, 7, "avgpool"), (3, 8, 8, 700, 8, 5, "maxpool"), (1, 32, 8, 4099, 8, 7, "avgpool"),
                                     (2, 16, 4, 2048, 16, 13, "avgpool"), (1, 16, 4, 1000, 32, 7, "avgpool")]:
    q, k, v = [x.cuda() for x in make_inputs(l + b, b, hq, hkv, l, 128)]
    ref = ops.score(q, k, window_size=w, kernel_size=ks, pooling=pool, impl=L.IMPL_SIMPLE).double()
    got = ops.score(q, k, window_size=w, kernel_size=ks, pooling=pool, impl=L.IMPL_TENSOR).double()
    torch.cuda.synchronize()
    scale = ref.abs().amax(-1, keepdim=True)
    err = ((ref - got).abs() / (ref.abs() + 1e-3 * scale)).max().item()
    assert err < 2e-5 and not torch.isnan(got).any(), (b, hq, hkv, l, w, ks, pool, err)
print("phases ok")
"""


@pytest.mark.parametrize("ppp", ["1", "2", "3"])
def test_scoring_kernel_phase_loop_at_small_sizes(ppp):
    """FKV_TC_PPP caps the (b, h) pairs per phase, so the multi-phase path of the tcgen05 kernel (TMEM
    blocks recycled between phases, per-phase partial exchange) runs at sizes the fp32 path checks in
    milliseconds.  The knob is read once per process, hence the subprocess."""
    import os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, FKV_TC_PPP=ppp)
    r = subprocess.run([sys.executable, "-c", _PHASE_CHECK, root], env=env, capture_output=True, text=True, timeout=170)
    assert r.returncode == 0 and "phases ok" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


# ---------------------------------------------------------------------------------------------
# the contract sizes (BASELINE.json configs[1..3]) against the pinned CPU oracle, full protocol
# ---------------------------------------------------------------------------------------------
def _contract_inputs(seed, b, l, hq, hkv, d, dist):
    """bf16 q/k/v generated on the GPU ([B,H,L,D] views of [B,L,H,D] memory, the call-site layout)."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    mk = lambda h: torch.randn(b, l, h, d, generator=g, device=DEV, dtype=torch.bfloat16)
    q, k, v = mk(hq), mk(hkv), mk(hkv)
    if dist == "peaky":                              # 1 % heavy-hitter tokens
        hot = torch.rand(l, generator=g, device=DEV) < 0.01
        k[:, hot] = (k[:, hot].float() * 4).to(torch.bfloat16)
    elif dist == "ties":                             # K quantised to 4 distinct rows: exact score ties
        rows = torch.randn(4, d, generator=g, device=DEV, dtype=torch.bfloat16)
        k = rows[torch.randint(0, 4, (b, l, hkv), generator=g, device=DEV)]
    return q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2)


CONTRACT = [
    # L, C, T, B, dist
    (32768, 2048, 2048, 1, "gauss"),                 # configs[1]: Llama-3.1-8B, 32k prefill, C = T = 2048
    (32768, 2048, 2048, 1, "peaky"),
    (32768, 2048, 2048, 1, "ties"),
    (131072, 2048, 2048, 1, "gauss"),                # configs[2]: 128k prefill (4-phase TMEM-recycling path)
    (32768, 512, 2048, 8, "gauss"),                  # configs[3]: Ministral-8B geometry, batch 8 x 32k, C = 512 (8 phases)
]


@pytest.mark.parametrize("l,c,t,b,dist", CONTRACT, ids=[f"L{x[0]}_C{x[1]}_B{x[3]}_{x[4]}" for x in CONTRACT])
def test_contract_sizes_vs_oracle(l, c, t, b, dist):
    """BASELINE configs[1..3] at their REAL sizes against the pinned oracle (fp32, CPU), full protocol: scores
    <= 2e-5, per-head and TSP index sets equal up to the 1e-5 near-tie rule, K/V bit-exact given our indices and
    equal to the oracle's rows wherever the indices agree."""
    hq, hkv, d, w = 32, 8, 128, 8
    q, k, v = _contract_inputs(l + c + b, b, l, hq, hkv, d, dist)
    knobs = dict(window_size=w, max_capacity_prompt=c, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=t)
    cl, (ko, vo, tsp) = _run_cluster(knobs, k, q, v, hq // hkv, _lib().IMPL_AUTO)
    aux = cl.last_aux
    assert _lib().load().fkv_score_impl_name is not None
    assert ko.shape == (b, hkv, c, d) and tsp.shape == (b, t)
    check_topk_consistent(aux["score"], aux["idx"], "contract")
    check_tsp_consistent(aux["importance"], tsp, w, "contract")
    check_compaction_exact(k, v, aux["idx"], ko, vo, w)
    score, idx, imp = aux["score"].cpu(), aux["idx"].cpu(), aux["importance"].cpu()
    ko_c, vo_c, tsp_c = ko.cpu(), vo.cpu(), tsp.cpu()
    for bi in range(b):                               # one prompt at a time keeps the fp32 oracle's footprint small
        ref = O.fastkv_update_kv(k[bi:bi + 1].cpu().float(), q[bi:bi + 1].cpu().float(), v[bi:bi + 1].cpu().float(),
                                 hq // hkv, **knobs)
        check_scores_close(ref.score, score[bi:bi + 1], SCORE_RTOL, f"contract b{bi}")
        check_scores_close(ref.importance, imp[bi:bi + 1], SCORE_RTOL, f"contract importance b{bi}")
        if dist == "ties":
            # exact ties: the reference's tie order is unspecified -> the selected score multisets must agree
            ours = torch.gather(ref.score[0], -1, idx[bi]).sort(-1).values
            theirs = torch.gather(ref.score[0], -1, ref.idx[0]).sort(-1).values
            assert torch.allclose(ours, theirs, rtol=1e-5, atol=0)
            ours = torch.gather(ref.importance[0], -1, tsp_c[bi, :-w]).sort(-1).values
            theirs = torch.gather(ref.importance[0], -1, ref.tsp_idx[0, :-w]).sort(-1).values
            assert torch.allclose(ours, theirs, rtol=1e-5, atol=0)
            continue
        check_selection(ref.score[0], ref.idx[0], idx[bi], f"contract b{bi}")
        check_selection(ref.importance, ref.tsp_idx[:, :-w], tsp_c[bi:bi + 1, :-w], f"contract tsp b{bi}")
        assert torch.equal(tsp_c[bi, -w:], ref.tsp_idx[0, -w:])
        # cache layout: same rows as the reference at every position where the (score-descending) orders agree;
        # continuous scores at these sizes leave at most a handful of near-tie swaps
        same = idx[bi] == ref.idx[0]                                             # [Hkv, C-W]
        assert same.float().mean().item() > 0.99, same.float().mean().item()
        assert torch.equal(ko_c[bi, :, :c - w][same], ref.key[0, :, :c - w].to(torch.bfloat16)[same])
        assert torch.equal(vo_c[bi, :, :c - w][same], ref.value[0, :, :c - w].to(torch.bfloat16)[same])
        assert torch.equal(ko_c[bi, :, c - w:], ref.key[0, :, c - w:].to(torch.bfloat16))
        assert torch.equal(vo_c[bi, :, c - w:], ref.value[0, :, c - w:].to(torch.bfloat16))
    # propagation at the real size: hidden rows / position ids bit-exact for the selected set
    from fastkv_b200 import ops
    hidden = torch.randn(b, l, 4096, generator=torch.Generator(device=DEV).manual_seed(l), device=DEV, dtype=torch.bfloat16)
    pos = torch.arange(l, device=DEV).unsqueeze(0)
    nh, npos = ops.gather_rows(hidden, tsp, pos)
    assert torch.equal(npos, tsp)
    assert torch.equal(nh, torch.gather(hidden, 1, tsp.unsqueeze(-1).expand(-1, -1, 4096)))


def test_len_equal_capacity_reorders_everything():
    """q_len == C still runs the selection over all L-W prefix tokens (strict '<', utils.py:89)."""
    knobs = dict(window_size=8, max_capacity_prompt=2048, kernel_size=7, pooling="avgpool", tsp_layer=False)
    q, k, v = [x.to(DEV) for x in make_inputs(5, 1, 32, 8, 2048, 128)]
    cl, (ko, vo, tsp) = _run_cluster(knobs, k, q, v, 4, _lib().IMPL_AUTO)
    assert tsp is None and ko.shape == (1, 8, 2048, 128)
    idx = cl.last_aux["idx"]
    assert torch.equal(idx.sort(-1).values, torch.arange(2040, device=DEV).expand(1, 8, -1))
    check_compaction_exact(k, v, idx, ko, vo, 8)


# ---------------------------------------------------------------------------------------------
# individual kernels
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,k", [(1, 1), (7, 7), (1000, 1), (1000, 999), (4096, 512), (32760, 2040), (131064, 2040),
                                 (131064, 13099), (50000, 20000)])
def test_topk_kernel_is_exact(n, k):
    from fastkv_b200 import ops
    L = _lib()
    g = torch.Generator().manual_seed(n * 7 + k)
    rows = torch.stack([
        torch.rand(n, generator=g),                                    # positive, continuous
        torch.randn(n, generator=g),                                   # signed
        torch.randint(0, 5, (n,), generator=g).float(),                # heavy exact ties
        torch.zeros(n),                                                # all equal
        torch.rand(n, generator=g).pow(8) * 1e-6,                      # tiny positive (softmax-like)
        1.0 + 0.01 * torch.rand(n, generator=g),                       # narrow: one crowded 12-bit bin (refinement pass)
        1.0 + 1e-5 * torch.rand(n, generator=g),                       # very narrow: crowded after refinement too
        (3e-5 * (1.0 + 0.05 * torch.randn(n, generator=g))).abs(),     # pooled-score-like: tight bell around a tiny mean
    ])
    rows[1, ::97] = -0.0
    got = ops.topk(rows.to(DEV), k, order=L.ORDER_VALUE_DESC).cpu()
    assert torch.equal(got, O.stable_topk_desc(rows, k))
    got = ops.topk(rows.to(DEV), k, order=L.ORDER_INDEX_ASC, tail=3).cpu()
    exp = torch.cat([O.stable_topk_desc(rows, k).sort(-1).values, torch.arange(n, n + 3).expand(rows.shape[0], -1)], -1)
    assert torch.equal(got, exp)


def test_topk_head_sum_is_fixed_order():
    from fastkv_b200 import ops
    L = _lib()
    x = torch.rand(3, 8, 5000, generator=torch.Generator().manual_seed(3))
    idx, imp = ops.topk(x.to(DEV), 300, order=L.ORDER_INDEX_ASC, tail=8, nsum=8, return_importance=True)
    ref = x[:, 0].clone()
    for h in range(1, 8):
        ref += x[:, h]
    assert torch.equal(imp.cpu(), ref)
    assert torch.equal(idx.cpu()[:, :300], O.stable_topk_desc(ref, 300).sort(-1).values)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_gather_rows_bit_exact(dtype):
    from fastkv_b200 import ops
    g = torch.Generator().manual_seed(1)
    hidden = torch.randn(2, 777, 4096, generator=g).to(dtype)
    idx = torch.stack([torch.randperm(777, generator=g)[:100].sort().values for _ in range(2)])
    pos = torch.arange(777).unsqueeze(0) + 5
    nh, npos = ops.gather_rows(hidden.to(DEV), idx.to(DEV), pos.to(DEV))
    eh, ep = O.tsp_propagate(hidden, pos.expand(2, -1), idx)
    assert torch.equal(nh.cpu(), eh) and torch.equal(npos.cpu(), ep)


@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 2e-2), (torch.float16, 2e-3), (torch.float32, 1e-5)])
@pytest.mark.parametrize("b,hq,hkv,d,s,smax", [(1, 32, 8, 128, 2048, 2304), (8, 32, 8, 128, 520, 768), (2, 8, 8, 64, 77, 128),
                                               (1, 16, 2, 128, 1, 64)])
def test_decode_attention_vs_oracle(dtype, tol, b, hq, hkv, d, s, smax):
    from fastkv_b200 import ops
    g = torch.Generator().manual_seed(s)
    q = torch.randn(b, hq, d, generator=g).to(dtype)
    kc = torch.randn(b, hkv, smax, d, generator=g).to(dtype)
    vc = torch.randn(b, hkv, smax, d, generator=g).to(dtype)
    ref = O.decode_attention(q, kc, vc, s)
    out = ops.decode_attn(q.to(DEV), kc.to(DEV), vc.to(DEV), s)
    assert torch.allclose(out.cpu().float(), ref, atol=tol, rtol=tol)
    # device-resident length (CUDA-graph friendly) gives the same answer
    sl = torch.tensor([s], dtype=torch.int32, device=DEV)
    out2 = ops.decode_attn(q.to(DEV), kc.to(DEV), vc.to(DEV), 0, seq_len_dev=sl)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("b,hq,hkv,d,s,smax", [(1, 32, 8, 128, 2304, 2304), (8, 32, 8, 128, 640, 768), (2, 24, 8, 128, 333, 512),
                                               (1, 32, 32, 64, 100, 128)])
def test_decode_attention_vs_library_attention_on_the_gpu(b, hq, hkv, d, s, smax):
    """the reference's decode attention is flash-attn's kernel (third party, unpinned).  Cross-check K5 on
    the GPU box against torch SDPA (always) and the installed flash_attn (when it runs on this GPU), bf16 tolerance."""
    import torch.nn.functional as F
    from fastkv_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(s)
    q = torch.randn(b, hq, d, generator=g, device=DEV, dtype=torch.bfloat16)
    kc = torch.randn(b, hkv, smax, d, generator=g, device=DEV, dtype=torch.bfloat16)
    vc = torch.randn(b, hkv, smax, d, generator=g, device=DEV, dtype=torch.bfloat16)
    got = ops.decode_attn(q, kc, vc, s).float()
    ke = kc[:, :, :s].repeat_interleave(hq // hkv, dim=1)
    ve = vc[:, :, :s].repeat_interleave(hq // hkv, dim=1)
    sdpa = F.scaled_dot_product_attention(q[:, :, None, :], ke, ve, is_causal=False)[:, :, 0].float()
    assert torch.allclose(got, sdpa, atol=2e-2, rtol=2e-2), (got - sdpa).abs().max().item()
    try:
        from flash_attn import flash_attn_func
        fa = flash_attn_func(q[:, None], kc[:, :, :s].transpose(1, 2), vc[:, :, :s].transpose(1, 2), causal=True)[:, 0].float()
    except Exception as e:  # noqa: BLE001  (flash-attn 2.x has no sm_100 kernels on some builds)
        pytest.skip(f"SDPA leg passed; flash_attn not runnable here: {type(e).__name__}: {str(e)[:80]}")
    assert torch.allclose(got, fa, atol=2e-2, rtol=2e-2), (got - fa).abs().max().item()


def test_full_cache_drops_the_row_and_keeps_the_length():
    """Device-resident lengths (CUDA-graph decode): appending to a full cache must not advance the length, and the
    attention kernel never reads past max_seq_len."""
    from fastkv_b200 import ops
    b, hkv, d, smax = 1, 2, 128, 16
    kc = torch.randn(b, hkv, smax, d, device=DEV, dtype=torch.bfloat16)
    vc = torch.randn(b, hkv, smax, d, device=DEV, dtype=torch.bfloat16)
    k0, v0 = kc.clone(), vc.clone()
    sl = torch.tensor([smax], dtype=torch.int32, device=DEV)
    kn = torch.randn(b, hkv, 1, d, device=DEV, dtype=torch.bfloat16)
    ops.cache_append(kn, kn, kc, vc, 0, seq_len_dev=sl, bump=True)
    assert sl.item() == smax and torch.equal(kc, k0) and torch.equal(vc, v0)
    q = torch.randn(b, 8, d, device=DEV, dtype=torch.bfloat16)
    sl.fill_(smax + 5)                                  # a corrupted length is clamped, not followed
    got = ops.decode_attn(q, kc, vc, 0, seq_len_dev=sl)
    exp = O.decode_attention(q.cpu(), kc.cpu(), vc.cpu(), smax)
    assert torch.allclose(got.cpu().float(), exp, atol=2e-2, rtol=2e-2)


def test_cache_append_and_decode_loop():
    from fastkv_b200 import ops
    g = torch.Generator().manual_seed(9)
    b, hkv, d, s0, smax = 2, 8, 128, 40, 64
    kc = torch.zeros(b, hkv, smax, d, dtype=torch.bfloat16, device=DEV)
    vc = torch.zeros_like(kc)
    base_k = torch.randn(b, hkv, s0, d, generator=g).bfloat16()
    kc[:, :, :s0] = base_k.to(DEV)
    sl = torch.tensor([s0], dtype=torch.int32, device=DEV)
    news = []
    for _ in range(5):
        kn = torch.randn(b, 1, hkv, d, generator=g).bfloat16().to(DEV).transpose(1, 2)      # [B,Hkv,1,D] strided view
        vn = torch.randn(b, 1, hkv, d, generator=g).bfloat16().to(DEV).transpose(1, 2)
        ops.cache_append(kn, vn, kc, vc, 0, seq_len_dev=sl, bump=True)
        news.append((kn, vn))
    assert sl.item() == s0 + 5
    for i, (kn, vn) in enumerate(news):
        assert torch.equal(kc[:, :, s0 + i], kn[:, :, 0]) and torch.equal(vc[:, :, s0 + i], vn[:, :, 0])
    assert torch.equal(kc[:, :, :s0].cpu(), base_k)


def test_host_buffer_path_matches_device_path():
    """update_kv on pinned host tensors (the e2e bench path) stages through the GPU kernels."""
    from fastkv_b200 import FastKVCluster
    knobs = dict(window_size=8, max_capacity_prompt=128, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=256)
    q, k, v = make_inputs(21, 1, 32, 8, 2048, 128, layout="bhld")
    a = FastKVCluster(**knobs).update_kv(k.pin_memory(), q.pin_memory(), v.pin_memory(), None, 4, 0)
    b = FastKVCluster(**knobs).update_kv(k.to(DEV), q.to(DEV), v.to(DEV), None, 4, 0)
    for x, y in zip(a, b):
        assert not x.is_cuda and torch.equal(x, y.cpu())


def test_pipelined_host_calls_deliver_the_blocking_results():
    """FastKVCluster.host_pipeline: calls on pinned host tensors return without synchronising (round-robin streams);
    after join() every call's results equal the blocking call's."""
    from fastkv_b200 import FastKVCluster, ops
    knobs = dict(window_size=8, max_capacity_prompt=256, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=512)
    inputs = [[t.pin_memory() for t in make_inputs(30 + i, 1, 32, 8, 4096, 128, layout="bhld")] for i in range(5)]
    want = [FastKVCluster(**knobs).update_kv(k, q, v, None, 4, 0) for (q, k, v) in inputs]
    pipe = ops.HostPipeline(DEV, depth=2)
    cl = FastKVCluster(**knobs)
    cl.host_pipeline = pipe
    got = [cl.update_kv(k, q, v, None, 4, 0) for (q, k, v) in inputs]
    pipe.join()
    for a, b in zip(got, want):
        for x, y in zip(a, b):
            assert not x.is_cuda and x.is_pinned() and torch.equal(x, y)
    # results delivered into the caller's own pinned host cache slots (out_provider asked with device = cpu)
    slots = [tuple(torch.empty((1, 8, 256, 128), dtype=torch.bfloat16, pin_memory=True) for _ in range(2)) for _ in inputs]
    got = []
    for i, (q, k, v) in enumerate(inputs):
        cl.out_provider = lambda *a, _s=slots[i]: _s
        got.append(cl.update_kv(k, q, v, None, 4, 0))
    pipe.join()
    for i, (a, b) in enumerate(zip(got, want)):
        assert a[0] is slots[i][0] and a[1] is slots[i][1]
        for x, y in zip(a, b):
            assert torch.equal(x, y)


def test_gather_kv_reads_pinned_host_sources_zero_copy():
    """K/V gather sources may be page-locked host tensors (only the selected rows cross PCIe); outputs are on the GPU."""
    from fastkv_b200 import ops
    g = torch.Generator().manual_seed(2)
    k = torch.randn(2, 700, 4, 128, generator=g).bfloat16().pin_memory().transpose(1, 2)
    v = torch.randn(2, 700, 4, 128, generator=g).bfloat16().pin_memory().transpose(1, 2)
    idx = torch.stack([torch.stack([torch.randperm(692, generator=g)[:100] for _ in range(4)]) for _ in range(2)]).to(DEV)
    ko, vo = ops.gather_kv(k.to(DEV), v, idx, window_size=8)
    kr, vr = ops.gather_kv(k.to(DEV), v.to(DEV), idx, window_size=8)
    torch.cuda.synchronize()
    assert ko.is_cuda and vo.is_cuda and torch.equal(ko, kr) and torch.equal(vo, vr)


def test_bad_pooling_raises_like_the_reference():
    from fastkv_b200 import FastKVCluster
    q, k, v = [x.to(DEV) for x in make_inputs(1, 1, 8, 2, 256, 128)]
    cl = FastKVCluster(window_size=8, max_capacity_prompt=64, pooling="minpool")
    with pytest.raises(ValueError, match="Pooling method not supported"):
        cl.update_kv(k, q, v, None, 4, 0)
