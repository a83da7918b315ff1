"""FKV_TC_TRACE=1 python tools/trace_layout.py <ctx> <layout blhd|bhld>  — phase timeline of the scoring kernel.
Needs a library built with FKV_BUILD_TRACE=1 python fastkv_b200/_build.py (tracing is compiled out by default)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastkv_b200 import ops, _lib as L
ctx, layout = int(sys.argv[1]), sys.argv[2]
g = torch.Generator(device="cuda").manual_seed(1)
def mk(h):
    x = torch.randn(1, ctx, h, 128, generator=g, device="cuda", dtype=torch.bfloat16).transpose(1, 2)
    return x.contiguous() if layout == "bhld" else x
nbuf = int(sys.argv[3]) if len(sys.argv) > 3 else 3
bufs = [(mk(32), mk(8)) for _ in range(nbuf)]
for i in range(4):
    q, k = bufs[i % nbuf]
    if i == 3: print(f"--- ctx {ctx} layout {layout}", file=sys.stderr, flush=True)
    s = ops.score(q, k, window_size=8, kernel_size=7, pooling="avgpool", impl=L.IMPL_TENSOR)
torch.cuda.synchronize()
