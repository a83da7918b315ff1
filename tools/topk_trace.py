import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastkv_b200 import ops, _lib as L
for n in (2040, 32760, 131064):
    x = torch.rand(1, 8, n, device="cuda").pow(6) * 1e-3
    for _ in range(3):
        ops.topk(x.view(8, n), min(2040, n))
        ops.topk(x, min(2040, n), order=L.ORDER_INDEX_ASC, tail=8, nsum=8)
torch.cuda.synchronize()
