"""Generate golden vectors by running the UNMODIFIED reference on CPU.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py

Imports ``FastKVCluster`` (baselines/fastkv/utils.py:49), ``SnapKVCluster``
(baselines/snapkv/utils.py:26) and ``standard_dis_index`` / ``find_context``
(baselines/gemfilter/utils.py:25,41) straight from the reference tree, feeds them seeded inputs
whose values are exactly representable in bf16 (so the same numbers can be handed to the CUDA
kernels as bf16 and to the reference as fp32 — the parity protocol of SURVEY.md §8c), and stores
inputs + outputs as small ``.pt`` fixtures next to this script.

Only the last W query rows are stored: the reference reads nothing else from ``query_states``
(baselines/fastkv/utils.py:94), the loader re-creates a full-length Q with zeros in front.
"""
import os
import sys
import types

import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _import_reference():
    sys.path.insert(0, REF)
    from baselines.fastkv.utils import FastKVCluster            # noqa: E402
    from baselines.snapkv.utils import SnapKVCluster            # noqa: E402
    from baselines.gemfilter import utils as gem                # noqa: E402
    return FastKVCluster, SnapKVCluster, gem


def bf16_exact(*shape, gen, scale=1.0):
    return (torch.randn(*shape, generator=gen) * scale).to(torch.bfloat16)


def make_inputs(seed, b, hq, hkv, l, d, w, dist="gauss"):
    g = torch.Generator().manual_seed(seed)
    k = bf16_exact(b, hkv, l, d, gen=g)
    v = bf16_exact(b, hkv, l, d, gen=g)
    qw = bf16_exact(b, hq, w, d, gen=g)
    if dist == "peaky":                       # 1 % heavy-hitter keys, x4
        hot = torch.rand(l, generator=g) < 0.01
        k[:, :, hot, :] = (k[:, :, hot, :].float() * 4).to(torch.bfloat16)
    elif dist == "ties":                      # keys drawn from 4 distinct rows -> exact score ties
        rows = bf16_exact(4, d, gen=g)
        pick = torch.randint(0, 4, (b, hkv, l), generator=g)
        k = rows[pick]
    return k, v, qw


def full_q(qw, l):
    b, hq, w, d = qw.shape
    q = torch.zeros(b, hq, l, d, dtype=qw.dtype)
    q[:, :, l - w:, :] = qw
    return q


FASTKV_CASES = [
    # name, seed, B, Hq, Hkv, L, D, knobs, dist, run_dtype
    ("fastkv_avg_small", 1, 1, 8, 2, 384, 128,
     dict(window_size=8, max_capacity_prompt=64, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=128), "gauss", "fp32"),
    ("fastkv_max_b2_ragged", 2, 2, 8, 2, 301, 128,
     dict(window_size=16, max_capacity_prompt=96, kernel_size=5, pooling="maxpool", tsp_layer=True, tsp_length=100), "gauss", "fp32"),
    ("fastkv_llama_shape", 3, 1, 32, 8, 512, 128,
     dict(window_size=8, max_capacity_prompt=128, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=256), "peaky", "fp32"),
    ("fastkv_no_tsp_layer", 4, 1, 8, 2, 257, 128,
     dict(window_size=8, max_capacity_prompt=40, kernel_size=13, pooling="avgpool", tsp_layer=False, tsp_length=64), "gauss", "fp32"),
    ("fastkv_early_return", 5, 1, 8, 2, 64, 128,
     dict(window_size=8, max_capacity_prompt=128, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=32), "gauss", "fp32"),
    ("fastkv_len_equals_cap", 6, 1, 8, 2, 128, 128,
     dict(window_size=8, max_capacity_prompt=128, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=128), "gauss", "fp32"),
    ("fastkv_proportional", 7, 1, 8, 2, 400, 128,
     dict(window_size=8, max_capacity_prompt=512, kernel_size=7, pooling="maxpool", tsp_layer=True, tsp_length=2048,
          tsp_rate=0.2, retain_rate=0.1, eviction_mode="proportional"), "peaky", "fp32"),
    ("fastkv_ties", 8, 1, 8, 2, 320, 128,
     dict(window_size=8, max_capacity_prompt=64, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=96), "ties", "fp32"),
    ("fastkv_bf16_run", 9, 1, 8, 2, 384, 128,
     dict(window_size=8, max_capacity_prompt=64, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=128), "gauss", "bf16"),
    ("fastkv_d64", 10, 1, 8, 4, 200, 64,
     dict(window_size=8, max_capacity_prompt=48, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=80), "gauss", "fp32"),
]

SNAPKV_CASES = [
    ("snapkv_avg", 21, 1, 8, 2, 300, 128,
     dict(window_size=16, max_capacity_prompt=80, kernel_size=5, pooling="avgpool"), "gauss"),
    ("snapkv_max_b2", 22, 2, 8, 2, 260, 128,
     dict(window_size=8, max_capacity_prompt=72, kernel_size=7, pooling="maxpool"), "peaky"),
]

GEMFILTER_CASES = [
    ("gemfilter_basic", 31, 8, 2, 333, 128, 50),
    ("gemfilter_topk_exceeds_n", 32, 8, 2, 40, 128, 64),
]


def main():
    FastKVCluster, SnapKVCluster, gem = _import_reference()
    torch.manual_seed(0)
    torch.set_num_threads(1)          # fixed reduction order for reproducible fixtures

    for name, seed, b, hq, hkv, l, d, knobs, dist, run_dtype in FASTKV_CASES:
        k, v, qw = make_inputs(seed, b, hq, hkv, l, d, knobs["window_size"], dist)
        dt = torch.float32 if run_dtype == "fp32" else torch.bfloat16
        cl = FastKVCluster(**knobs)
        ko, vo, tsp = cl.update_kv(k.to(dt), full_q(qw, l).to(dt), v.to(dt), None, hq // hkv, 0)
        fx = dict(kind="fastkv", name=name, shape=dict(B=b, Hq=hq, Hkv=hkv, L=l, D=d), knobs=knobs, dist=dist,
                  run_dtype=run_dtype, k=k, v=v, q_window=qw, key_out=ko.to(torch.bfloat16), value_out=vo.to(torch.bfloat16),
                  tsp_idx=tsp, capacity_after=cl.max_capacity_prompt, tsp_length_after=cl.tsp_length)
        torch.save(fx, os.path.join(HERE, name + ".pt"))
        print(name, tuple(ko.shape), None if tsp is None else tuple(tsp.shape))

    for name, seed, b, hq, hkv, l, d, knobs, dist in SNAPKV_CASES:
        k, v, qw = make_inputs(seed, b, hq, hkv, l, d, knobs["window_size"], dist)
        g = hq // hkv
        # caller-side expansion, baselines/snapkv/llama_model.py:385-386
        from baselines.snapkv.utils import repeat_kv
        cl = SnapKVCluster(**knobs)
        ko, vo = cl.update_kv(repeat_kv(k.float(), g), full_q(qw, l).float(), repeat_kv(v.float(), g), None, g)
        fx = dict(kind="snapkv", name=name, shape=dict(B=b, Hq=hq, Hkv=hkv, L=l, D=d), knobs=knobs, dist=dist,
                  k=k, v=v, q_window=qw, key_out=ko.to(torch.bfloat16), value_out=vo.to(torch.bfloat16))
        torch.save(fx, os.path.join(HERE, name + ".pt"))
        print(name, tuple(ko.shape))

    for name, seed, hq, hkv, n, d, topk in GEMFILTER_CASES:
        k, v, qw = make_inputs(seed, 1, hq, hkv, n, d, 1, "gauss")
        holder = types.SimpleNamespace(eviction_mode="constant", retain_rate=0.1, topk=topk, indecies=None,
                                       layer_idx=13, select_layer_idx=13, num_key_value_groups=hq // hkv)
        gem.find_context(holder, full_q(qw, n).float(), k.float())
        idx = holder.indecies
        fx = dict(kind="gemfilter", name=name, shape=dict(B=1, Hq=hq, Hkv=hkv, L=n, D=d), topk=topk,
                  k=k, q_window=qw, indices=idx, sorted_indices=torch.sort(idx[0, 0, :]).values)
        torch.save(fx, os.path.join(HERE, name + ".pt"))
        print(name, tuple(idx.shape))


if __name__ == "__main__":
    main()
