"""SURVEY.md §8(d) config-5 sweep: pipeline time of one update_kv call (score -> pool -> top-k -> compaction)
over context length, KV budget, window, pooling kernel, pooling type and selection mode, timed by
CUDA-graph replay of the C-ABI launches with rotating input buffers (K comes from HBM, not L2).

    python tools/sweep.py [--quick] [--out gpurun_out/sweep.md]

Timing only (no oracle here: parity over the same grid lives in tests/test_gpu_parity.py).
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from fastkv_b200 import ops, _lib as L  # noqa: E402

D, HQ, HKV = 128, 32, 8


def alg_bytes(mode, b, l, c, w, e=2):
    """§8(d) algorithmic bytes of one call (K read once, Q window, score round trip, indices, compaction)."""
    hsel = HKV if mode == "fastkv" else HQ
    if mode == "gemfilter":
        return b * HKV * l * D * e + b * HQ * D * e + 2 * b * l * 4 + b * c * 8
    p = l - w
    return (b * HKV * l * D * e + b * HQ * w * D * e + 2 * b * hsel * p * 4 + b * hsel * (c - w) * 8
            + 2 * 2 * b * hsel * c * D * e)


def make(b, l, h, g, dtype=torch.bfloat16):
    return torch.randn(b, l, h, D, generator=g, device="cuda", dtype=torch.bfloat16).transpose(1, 2).to(dtype)


def run_one(mode, b, l, c, w, ks, pool, reps=10, dtype="bf16"):
    dt = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}[dtype]
    g = torch.Generator(device="cuda").manual_seed(42)
    kbytes = b * l * HKV * D * 2
    nbuf = max(2, min(6, int(300e6 // kbytes) + 1))
    bufs = [(make(b, l, HQ, g, dt), make(b, l, HKV, g, dt), make(b, l, HKV, g, dt)) for _ in range(nbuf)]
    code = {"fastkv": L.MODE_FASTKV, "snapkv": L.MODE_SNAPKV, "gemfilter": L.MODE_GEMFILTER}[mode]

    def call(q, k, v):
        if mode == "gemfilter":
            s = ops.score(q, k, window_size=1, kernel_size=5, pooling="avgpool", mode=code)
            return ops.topk(s.view(b, -1), min(c, l), order=L.ORDER_INDEX_ASC)
        return ops.compress(q, k, v, window_size=w, kernel_size=ks, pooling=pool, capacity=c, mode=code)

    def score_only(q, k, v):
        if mode == "gemfilter":
            return ops.score(q, k, window_size=1, kernel_size=5, pooling="avgpool", mode=code)
        return ops.score(q, k, window_size=w, kernel_size=ks, pooling=pool, mode=code)

    out = []
    for fn in (call, score_only):
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for x in bufs:
                fn(*x)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                for x in bufs:
                    fn(*x)
            for _ in range(2):
                gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(reps):
                gr.replay()
            e1.record(st)
            torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1) * 1e3 / (reps * nbuf))
        del gr
    del bufs
    torch.cuda.empty_cache()
    return out


def grid(quick=False):
    """The config-5 grid (shared with ``bench.py --sweep``, which adds the reference-torch-on-this-GPU column)."""
    base = dict(mode="fastkv", b=1, l=32768, c=2048, w=8, ks=7, pool="avgpool")
    cfgs = []
    add = lambda **kw: cfgs.append({**base, **kw})  # noqa: E731
    for l in ([8192, 32768] if quick else [4096, 8192, 16384, 32768, 65536, 131072]):
        add(l=l)
    if not quick:
        for c in [128, 256, 512, 1024]:
            add(c=c)
        for w in [16, 32]:
            add(w=w)
        for ks in [5, 13]:
            add(ks=ks)
        add(pool="maxpool")
        add(l=131072, c=128, w=32, ks=13, pool="maxpool")
        add(b=8, c=512)                                  # config 4 shape
        add(l=8192, c=512, dtype="f32")                  # BASELINE configs[0] on the GPU: fp32 inputs, CUDA-core scoring path
        for mode in ["snapkv", "gemfilter"]:
            for l in [32768, 131072]:
                add(mode=mode, l=l)
    return cfgs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default="gpurun_out/sweep.md")
    ap.add_argument("--peak", type=float, default=6546.2, help="HBM GB/s the fractions are quoted against")
    a = ap.parse_args()
    cfgs = grid(a.quick)
    rows = ["| mode | B | L | C | W | kernel | pool | score impl | pipeline µs | score stage µs | algorithmic MB | pipeline GB/s | frac of %.0f GB/s |" % a.peak,
            "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for c in cfgs:
        tot, sc = run_one(**c)
        nb = alg_bytes(c["mode"], c["b"], c["l"], c["c"], c["w"], e=4 if c.get("dtype") == "f32" else 2)
        impl = "tcgen05"
        if impl == "tcgen05" and c["w"] * (HQ // HKV) > 32:
            impl += " x%d head groups" % (c["w"] * (HQ // HKV) // 32)
        gbs = nb / tot / 1e3
        if c.get("dtype") == "f32":
            impl = "fp32 CUDA cores"
        rows.append(f"| {c['mode']} | {c['b']} | {c['l']} | {c['c']} | {c['w']} | {c['ks']} | {c['pool']} | {impl} | {tot:.1f} | {sc:.1f} | "
                    f"{nb / 1e6:.1f} | {gbs:.0f} | {gbs / a.peak:.2f} |")
        print(rows[-1], flush=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as f:
        f.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
