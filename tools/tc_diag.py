import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from fastkv_b200 import ops, _lib as L
from parity_utils import make_inputs
b, l = int(sys.argv[1]), int(sys.argv[2])
q, k, v = [x.cuda() for x in make_inputs(l, b, 32, 8, l, 128)]
ref = ops.score(q, k, window_size=8, kernel_size=7, pooling='avgpool', impl=L.IMPL_SIMPLE).double()
got = ops.score(q, k, window_size=8, kernel_size=7, pooling='avgpool', impl=L.IMPL_TENSOR).double()
torch.cuda.synchronize()
scale = ref.abs().amax(-1, keepdim=True)
err = ((ref - got).abs() / (ref.abs() + 1e-3 * scale))
S = 122
for bb in range(b):
    for h in range(8):
        e = err[bb, h]
        bad = (e > 1e-3).nonzero().flatten()
        if bad.numel():
            tiles = sorted(set((bad // S).tolist()))
            print(f"b{bb} h{h}: {bad.numel()} bad, j range {bad.min().item()}..{bad.max().item()}, tiles {tiles[:6]}..{tiles[-3:]} ratio got/ref {(got[bb,h,bad[0]]/ref[bb,h,bad[0]]).item():.4f}")
print("max err", err.max().item())
