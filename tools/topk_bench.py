"""Times fkv_topk (per-head rows + TSP row) on L2-resident score rows: python tools/topk_bench.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastkv_b200 import ops, _lib as L

def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

for n in (2040, 32760, 131064):
    x = torch.rand(1, 8, n, device="cuda").pow(6) * 1e-3
    k = min(2040, n)
    g = torch.cuda.CUDAGraph()
    rows = x.view(8, n)
    ops.topk(rows, k); ops.topk(x, k, order=L.ORDER_INDEX_ASC, tail=8, nsum=8)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(10):
            a = ops.topk(rows, k)
            b = ops.topk(x, k, order=L.ORDER_INDEX_ASC, tail=8, nsum=8)
    us = timeit(g.replay, 20) / 10
    ga = torch.cuda.CUDAGraph()
    with torch.cuda.graph(ga):
        for _ in range(10): a = ops.topk(rows, k)
    usa = timeit(ga.replay, 20) / 10
    ref = torch.sort(rows, dim=-1, descending=True, stable=True).indices[:, :k]
    ok = torch.equal(a, ref)
    imp = x[:, 0].clone()
    for h in range(1, 8): imp += x[:, h]
    refb = torch.sort(imp, dim=-1, descending=True, stable=True).indices[:, :k].sort(-1).values
    okb = torch.equal(b[:, :k], refb)
    print(f"n={n} k={k}: per-head rows {usa:.2f} us, per-head + TSP {us:.2f} us   exact: {ok} {okb}", flush=True)
