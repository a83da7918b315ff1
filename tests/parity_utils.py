"""Shared checks for the GPU parity tests (oracle = checker, CUDA path = thing under test)."""
import torch

from oracle import fastkv_oracle as O

RTOL_TIE = 1e-5          # north-star near-tie tolerance on the reference's own score gaps


def make_inputs(seed, b, hq, hkv, l, d, dtype=torch.bfloat16, dist="gauss", layout="blhd"):
    """Seeded q/k/v with bf16-representable values.  ``layout='blhd'`` returns [B,H,L,D] *views* of
    [B,L,H,D] memory, the layout HF attention layers hand to update_kv."""
    g = torch.Generator().manual_seed(seed)
    def mk(h):
        x = torch.randn(b, l, h, d, generator=g).to(torch.bfloat16)
        return x
    q, k, v = mk(hq), mk(hkv), mk(hkv)
    if dist == "peaky":
        hot = torch.rand(l, generator=g) < 0.01
        k[:, hot] = (k[:, hot].float() * 4).to(torch.bfloat16)
    elif dist == "ties":
        rows = torch.randn(4, d, generator=g).to(torch.bfloat16)
        k = rows[torch.randint(0, 4, (b, l, hkv), generator=g)]
    outs = []
    for x in (q, k, v):
        x = x.to(dtype)
        outs.append(x.transpose(1, 2) if layout == "blhd" else x.transpose(1, 2).contiguous())
    return outs


def check_selection(ref_scores, ref_idx, got_idx, what=""):
    """ref_scores [R,n] fp32, ref_idx/got_idx [R,k]: sets equal up to the near-tie rule."""
    assert ref_idx.shape == got_idx.shape, (what, ref_idx.shape, got_idx.shape)
    ref_scores, ref_idx, got_idx = ref_scores.cpu(), ref_idx.cpu(), got_idx.cpu()
    for r in range(ref_idx.shape[0]):
        bad = O.near_tie_set_mismatch(ref_scores[r], ref_idx[r], got_idx[r], RTOL_TIE)
        assert bad == 0, f"{what}: row {r}: {bad} selections differ beyond the {RTOL_TIE} near-tie tolerance"


def check_scores_close(ref, got, rtol, what=""):
    ref, got = ref.cpu().double(), got.cpu().double()
    scale = ref.abs().amax(dim=-1, keepdim=True).clamp_min(1e-30)
    err = ((ref - got).abs() / (ref.abs() + 1e-3 * scale)).max().item()
    assert err <= rtol, f"{what}: max relative score error {err:.3e} > {rtol:.1e}"
    return err


def check_compaction_exact(k, v, idx, k_out, v_out, window, head_div=1):
    """Gathered K/V must be bit-identical to torch.gather with OUR indices + the window rows."""
    k, v, idx, k_out, v_out = k.cpu(), v.cpu(), idx.cpu(), k_out.cpu(), v_out.cpu()
    if head_div > 1:
        k, v = O.expand_kv_heads(k, head_div), O.expand_kv_heads(v, head_div)
    d = k.shape[-1]
    g = idx.unsqueeze(-1).expand(-1, -1, -1, d)
    exp_k = torch.cat([k[:, :, :-window].gather(2, g), k[:, :, -window:]], dim=2)
    exp_v = torch.cat([v[:, :, :-window].gather(2, g), v[:, :, -window:]], dim=2)
    assert torch.equal(k_out, exp_k), "compacted K differs from gather(K, idx) ++ window rows"
    assert torch.equal(v_out, exp_v), "compacted V differs from gather(V, idx) ++ window rows"
    assert k_out.is_contiguous() and v_out.is_contiguous()


def check_topk_consistent(scores, idx, what=""):
    """Size-independent property: idx is a valid top-k of OUR scores in (score desc, index asc) order."""
    scores, idx = scores.cpu(), idx.cpu()
    n, k = scores.shape[-1], idx.shape[-1]
    flat_s, flat_i = scores.reshape(-1, n), idx.reshape(-1, k)
    exp = O.stable_topk_desc(flat_s, k)
    assert torch.equal(flat_i, exp), f"{what}: indices are not the (score desc, index asc) top-k of the scores"


def check_tsp_consistent(importance, tsp, window, what=""):
    importance, tsp = importance.cpu(), tsp.cpu()
    b, p = importance.shape
    t = tsp.shape[1]
    exp_sel = O.stable_topk_desc(importance, t - window).sort(-1).values
    tail = torch.arange(p, p + window).unsqueeze(0).expand(b, -1)
    assert torch.equal(tsp, torch.cat([exp_sel, tail], dim=-1)), f"{what}: TSP indices inconsistent with importance"
