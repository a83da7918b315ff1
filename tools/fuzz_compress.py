"""Randomised self-consistency fuzz of the whole update_kv chain (fkv_compress): the per-head indices must be the
stable top-k (value desc, index asc) of the scores the call itself returned, the TSP set the stable top-k of its own
importance vector plus the window, and the compacted K/V bit-equal to torch.gather with those indices.
python tools/fuzz_compress.py [n_cases] [seed]"""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastkv_b200 import ops, _lib as L

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
g = torch.Generator(device="cuda").manual_seed(rng.randrange(1 << 30))
for case in range(n_cases):
    mode = rng.choice(["fastkv", "fastkv", "fastkv", "snapkv"])
    hkv = rng.choice([1, 2, 4, 8]); grp = rng.choice([1, 2, 4]); hq = hkv * grp
    d = rng.choice([64, 128]); w = rng.choice([8, 8, 16, 32]); ks = rng.choice([3, 5, 7, 13])
    pool = rng.choice(["avgpool", "maxpool"]); b = rng.choice([1, 1, 2, 3])
    l = rng.choice([100, 500, 1000, 2048, 4100, 9000, 20000, 40000])
    if b * l * hq > 1500000:
        b = 1
    c = rng.choice([w + 1, 64, 128, 512, 2048, l])
    c = max(w + 1, min(c, l))
    t = 0
    if mode == "fastkv" and rng.random() < 0.5 and l > w + 2:
        t = rng.randrange(w + 1, l)
    dt = rng.choice([torch.bfloat16, torch.float16])
    os.environ["FKV_SMALL_FUSE"] = rng.choice(["0", "1"])
    mk = lambda h: torch.randn(b, l, h, d, generator=g, device="cuda", dtype=torch.float32).to(dt).transpose(1, 2)
    q, k, v = mk(hq), mk(hkv), mk(hkv)
    code = L.MODE_FASTKV if mode == "fastkv" else L.MODE_SNAPKV
    desc = (case, mode, b, hq, hkv, l, d, w, ks, pool, c, t, str(dt), os.environ["FKV_SMALL_FUSE"])
    ko, vo, tsp, aux = ops.compress(q, k, v, window_size=w, kernel_size=ks, pooling=pool, capacity=c, tsp_len=t, mode=code, return_aux=True)
    torch.cuda.synchronize()
    score, idx = aux["score"], aux["idx"]
    n_sel = c - w
    exp = torch.sort(score, dim=-1, descending=True, stable=True).indices[..., :n_sel]
    checks = {"idx": torch.equal(idx, exp), "finite": bool(torch.isfinite(score).all())}
    hs = score.shape[1]
    div = hq // hkv if mode == "snapkv" else 1
    src = torch.arange(hs, device="cuda") // div
    full_idx = torch.cat([idx, torch.arange(l - w, l, device="cuda").expand(b, hs, w)], -1)
    kg = torch.gather(k[:, src], 2, full_idx[..., None].expand(-1, -1, -1, d))
    vg = torch.gather(v[:, src], 2, full_idx[..., None].expand(-1, -1, -1, d))
    checks["k"] = torch.equal(ko, kg); checks["v"] = torch.equal(vo, vg)
    if t:
        imp = aux["importance"] if aux["importance"] is not None else None
        ref = score[:, 0].clone()
        for h in range(1, hs):
            ref += score[:, h]
        checks["imp"] = (imp is None or torch.equal(imp, ref))
        top = torch.sort(ref, dim=-1, descending=True, stable=True).indices[:, :t - w].sort(-1).values
        exp_t = torch.cat([top, torch.arange(l - w, l, device="cuda").expand(b, w)], -1)
        checks["tsp"] = torch.equal(tsp, exp_t)
        if not checks["tsp"]:
            bad = (tsp != exp_t).nonzero()
            print("   tsp mismatches:", bad.shape[0], "first:", bad[:3].tolist(), tsp[bad[0, 0], bad[0, 1]].item(), exp_t[bad[0, 0], bad[0, 1]].item())
    ok = all(checks.values())
    if not checks["idx"]:
        bad = (idx != exp).nonzero()
        r = bad[0]
        print("   idx mismatches:", bad.shape[0], "first:", r.tolist(), idx[r[0], r[1], r[2]].item(), exp[r[0], r[1], r[2]].item(),
              score[r[0], r[1], idx[r[0], r[1], r[2]]].item(), score[r[0], r[1], exp[r[0], r[1], r[2]]].item())
    print("ok  " if ok else "FAIL " + str(checks), desc, flush=True)
    if not ok:
        sys.exit(1)
print("fuzz done")
