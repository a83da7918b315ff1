"""Graph-replay timing of the scoring kernel (no Python launch overhead, inputs rotated so K comes from HBM).

    python tools/score_time.py <ctx> [batch] [nbuf]
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastkv_b200 import ops, _lib as L
ctx = int(sys.argv[1]); b = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kbytes = b * ctx * 8 * 128 * 2
nbuf = int(sys.argv[3]) if len(sys.argv) > 3 else max(3, int(400e6 // kbytes) + 1)
g = torch.Generator(device="cuda").manual_seed(1)
mk = lambda h: torch.randn(b, ctx, h, 128, generator=g, device="cuda", dtype=torch.bfloat16).transpose(1, 2)
bufs = [(mk(32), mk(8)) for _ in range(nbuf)]
run = lambda: [ops.score(q, k, window_size=8, kernel_size=7, pooling="avgpool", impl=L.IMPL_TENSOR) for (q, k) in bufs]
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    run(); run()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, stream=st):
        run()
    for _ in range(3): gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record(st)
    for _ in range(reps): gr.replay()
    e1.record(st); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (reps * nbuf)
print(f"ctx {ctx} batch {b} nbuf {nbuf} ppp={os.environ.get('FKV_TC_PPP','auto')}: {us:.2f} us/call  {kbytes / us / 1e3:.0f} GB/s (K bytes)", flush=True)
