"""Randomised parity fuzz: tcgen05 scoring kernels (softmax phases, head sub-groups, GemFilter) against the fp32
CUDA-core path over random shapes / strides / knobs.  python tools/fuzz_score.py [n_cases] [seed]"""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastkv_b200 import ops, _lib as L

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
g = torch.Generator(device="cuda").manual_seed(rng.randrange(1 << 30))
worst = 0.0
for case in range(n_cases):
    mode = rng.choice(["fastkv", "fastkv", "snapkv", "gemfilter"])
    hkv = rng.choice([1, 2, 4, 8]); grp = rng.choice([1, 2, 4, 8]); hq = hkv * grp
    d = rng.choice([64, 128]); w = rng.choice([8, 8, 16, 32]); ks = rng.choice([1, 3, 5, 7, 13, 21])
    pool = rng.choice(["avgpool", "maxpool"]); b = rng.choice([1, 1, 2, 3, 5])
    l = rng.choice([w + 1, w + 2, 64, 121, 122, 123, 500, 1000, 2047, 4100, 9000, 20000]) if rng.random() < 0.8 else rng.randrange(w + 1, 30000)
    if b * l * hkv > 400000:
        b = 1
    dt = rng.choice([torch.bfloat16, torch.float16])
    blhd = rng.random() < 0.6                                   # [B,L,H,D] memory viewed as [B,H,L,D] (the real call site) or contiguous
    def mk(h):
        x = torch.randn(b, l, h, d, generator=g, device="cuda", dtype=torch.float32).mul_(rng.choice([0.5, 1.0, 2.0])).to(dt).transpose(1, 2)
        return x if blhd else x.contiguous()
    q, k = mk(hq), mk(hkv)
    code = {"fastkv": L.MODE_FASTKV, "snapkv": L.MODE_SNAPKV, "gemfilter": L.MODE_GEMFILTER}[mode]
    kw = dict(window_size=1 if mode == "gemfilter" else w, kernel_size=ks, pooling="avgpool" if mode == "gemfilter" else pool, mode=code)
    desc = (case, mode, b, hq, hkv, l, d, w, ks, pool, str(dt), blhd)
    try:
        ref = ops.score(q, k, impl=L.IMPL_SIMPLE, **kw).double()
        got = ops.score(q, k, impl=L.IMPL_TENSOR, **kw).double()
    except ValueError as e:                                   # shape outside the tensor-core path
        print("skip", desc, str(e)[:60]); continue
    torch.cuda.synchronize()
    if mode == "gemfilter":
        err = ((ref - got).abs() / ref.abs().amax(-1, keepdim=True).clamp_min(1e-30)).max().item()
    else:
        scale = ref.abs().amax(-1, keepdim=True)
        err = ((ref - got).abs() / (ref.abs() + 1e-3 * scale)).max().item() if ref.numel() else 0.0
    worst = max(worst, err)
    bad = (not torch.isfinite(got).all()) or err > (5e-5 if dt == torch.bfloat16 else 5e-5)
    print("FAIL" if bad else "ok  ", desc, f"{err:.2e}", flush=True)
    if bad:
        sys.exit(1)
print("fuzz done, worst rel err", worst)
