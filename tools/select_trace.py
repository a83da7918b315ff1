"""FKV_SELECT_TRACE=1 python tools/select_trace.py [L C B]: device-side timeline of the select + compact kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastkv_b200 import ops  # noqa: E402

l, c, b = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (32768, 2048, 1)
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda h: torch.randn(b, l, h, 128, generator=g, device="cuda", dtype=torch.bfloat16).transpose(1, 2)
q, k, v = mk(32), mk(8), mk(8)
for _ in range(4):
    ops.compress(q, k, v, window_size=8, kernel_size=7, pooling="avgpool", capacity=c)
torch.cuda.synchronize()
