"""FKV_TOPK_TRACE=1 python tools/topk_fused_trace.py [L] — phase times of the per-head top-k with the fused KV copy."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastkv_b200 import ops
l = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda h: torch.randn(1, l, h, 128, generator=g, device="cuda", dtype=torch.bfloat16).transpose(1, 2)
bufs = [(mk(32), mk(8), mk(8)) for _ in range(3)]
for i in range(6):
    q, k, v = bufs[i % 3]
    ops.compress(q, k, v, window_size=8, kernel_size=7, pooling="avgpool", capacity=2048)
torch.cuda.synchronize()
