"""Tiny driver for ncu captures of the scoring stage: python tools/prof_score.py <ctx> <reps> [impl]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastkv_b200 import ops, _lib as L
ctx, reps = int(sys.argv[1]), int(sys.argv[2])
impl = {"auto": L.IMPL_AUTO, "simple": L.IMPL_SIMPLE, "tensor": L.IMPL_TENSOR}[sys.argv[3] if len(sys.argv) > 3 else "auto"]
b = int(sys.argv[4]) if len(sys.argv) > 4 else 1
g = torch.Generator(device="cuda").manual_seed(1)
bufs = []
for i in range(3):
    q = torch.randn(b, ctx, 32, 128, generator=g, device="cuda", dtype=torch.bfloat16).transpose(1, 2)
    k = torch.randn(b, ctx, 8, 128, generator=g, device="cuda", dtype=torch.bfloat16).transpose(1, 2)
    bufs.append((q, k))
for i in range(reps):
    q, k = bufs[i % 3]
    s = ops.score(q, k, window_size=8, kernel_size=7, pooling="avgpool", impl=impl)
torch.cuda.synchronize()
print("done", s.shape, float(s.sum()))
