"""Do the scoring kernel and the select + compact kernel really run concurrently?  Times N scoring calls on one stream,
N full compress calls (different buffers) on another, and both together."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastkv_b200 import ops  # noqa: E402

l, c = 32768, 2048
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda h: torch.randn(1, l, h, 128, generator=g, device="cuda", dtype=torch.bfloat16).transpose(1, 2)
A = [(mk(32), mk(8), mk(8)) for _ in range(3)]
B = [(mk(32), mk(8), mk(8)) for _ in range(3)]
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
N = 30


def score_loop():
    with torch.cuda.stream(s0):
        for i in range(N):
            q, k, v = A[i % 3]
            ops.score(q, k, window_size=8, kernel_size=7, pooling="avgpool")


def compress_loop():
    with torch.cuda.stream(s1):
        for i in range(N):
            q, k, v = B[i % 3]
            ops.compress(q, k, v, window_size=8, kernel_size=7, pooling="avgpool", capacity=c)


def timed(fns):
    for f in fns:
        f()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cur = torch.cuda.current_stream()
        s0.wait_stream(cur); s1.wait_stream(cur)
        for f in fns:
            f()
        cur.wait_stream(s0); cur.wait_stream(s1)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * N)


a = timed([score_loop]); b = timed([compress_loop]); ab = timed([score_loop, compress_loop])
print(f"score alone {a:.1f} us/call, compress alone {b:.1f} us/call, both streams together {ab:.1f} us/pair (sum {a + b:.1f})")
