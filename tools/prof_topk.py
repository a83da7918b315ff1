"""Tiny driver for ncu captures of the per-head top-k + fused KV copy: python tools/prof_topk.py <ctx> <reps>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastkv_b200 import ops
ctx, reps = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda h: torch.randn(1, ctx, h, 128, generator=g, device="cuda", dtype=torch.bfloat16).transpose(1, 2)
bufs = [(mk(32), mk(8), mk(8)) for _ in range(3)]
for i in range(reps):
    q, k, v = bufs[i % 3]
    out = ops.compress(q, k, v, window_size=8, kernel_size=7, pooling="avgpool", capacity=2048)
torch.cuda.synchronize()
print("done", out[0].shape)
