"""Driver for ncu captures / timing of the decode-attention kernel: python tools/prof_decode.py <B> <S> <Smax> [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastkv_b200 import ops
b, s, smax = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 6
g = torch.Generator(device="cuda").manual_seed(1)
nbuf = max(4, int(300e6 // (2 * b * 8 * smax * 128 * 2)) + 1)
caches = [(torch.randn(b, 8, smax, 128, generator=g, device="cuda", dtype=torch.bfloat16),
           torch.randn(b, 8, smax, 128, generator=g, device="cuda", dtype=torch.bfloat16)) for _ in range(nbuf)]
q = torch.randn(b, 32, 128, generator=g, device="cuda", dtype=torch.bfloat16)
out = torch.empty_like(q)
for i in range(reps):
    kc, vc = caches[i % nbuf]
    ops.decode_attn(q, kc, vc, s, out=out)
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for kc, vc in caches:
        ops.decode_attn(q, kc, vc, s, out=out)
gr.replay(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    gr.replay()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (10 * nbuf)
nb = 2 * b * 8 * s * 128 * 2
print(f"B={b} S={s}: {us:.2f} us/call, {nb / us / 1e3:.0f} GB/s", flush=True)
