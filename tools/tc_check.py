"""Quick GPU check: tensor-core scoring vs the simple fp32 path on a few shapes, plus timings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from fastkv_b200 import ops, _lib as L
from parity_utils import make_inputs

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps

for (b, hq, hkv, l, w, ks, pool) in [(1, 32, 8, 2048, 8, 7, 'avgpool'), (1, 32, 8, 4099, 8, 7, 'avgpool'), (2, 16, 4, 2048, 16, 5, 'maxpool'),
                                     (1, 32, 8, 32768, 8, 7, 'avgpool'), (1, 8, 8, 1500, 8, 7, 'avgpool'), (1, 32, 8, 3000, 16, 13, 'avgpool'), (2, 32, 8, 5000, 32, 7, 'avgpool'), (1, 64, 8, 4000, 8, 7, 'maxpool'),
                                     (1, 32, 8, 32768, 16, 7, 'avgpool'), (1, 32, 8, 32768, 32, 7, 'avgpool'),
                                     (8, 32, 8, 32768, 8, 7, 'avgpool'), (1, 32, 8, 131072, 8, 7, 'avgpool')]:
    q, k, v = [x.cuda() for x in make_inputs(l, b, hq, hkv, l, 128)]
    s_simple = ops.score(q, k, window_size=w, kernel_size=ks, pooling=pool, impl=L.IMPL_SIMPLE)
    s_tc = ops.score(q, k, window_size=w, kernel_size=ks, pooling=pool, impl=L.IMPL_TENSOR)
    torch.cuda.synchronize()
    ref, got = s_simple.double(), s_tc.double()
    scale = ref.abs().amax(-1, keepdim=True)
    err = ((ref - got).abs() / (ref.abs() + 1e-3 * scale)).max().item()
    us = timeit(lambda: ops.score(q, k, window_size=w, kernel_size=ks, pooling=pool, impl=L.IMPL_TENSOR))
    kb = b * hkv * l * 128 * 2
    print((b, hq, hkv, l, w, ks, pool), f'max rel err {err:.2e} nan {int(torch.isnan(s_tc).sum())}  {us:.1f} us  {kb / us / 1e3:.0f} GB/s (K bytes, L2-warm)', flush=True)
