"""Per-call timing of the update_kv pipeline (score -> select + compact) and of its scoring stage, CUDA-graph replay over
rotating input sets (K comes from HBM).  usage: python tools/pipe_time.py [ctx:cap:batch ...]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastkv_b200 import ops  # noqa: E402

HQ, HKV, D, W, KS = 32, 8, 128, 8, 7
PEAK = 6546.2
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass


def layer_bytes(b, l, c):
    p = l - W
    return b * HKV * l * D * 2 + b * HQ * W * D * 2 + 2 * b * HKV * p * 4 + b * HKV * (c - W) * 8 + 4 * b * HKV * c * D * 2


def score_bytes(b, l):
    return b * HKV * l * D * 2 + b * HQ * W * D * 2 + b * HKV * (l - W) * 4


def timed(fn, bufs, reps=20):
    for x in bufs:
        fn(*x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for x in bufs:
            fn(*x)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (reps * len(bufs))


def main():
    cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(32768, 2048, 1), (131072, 2048, 1), (2048, 2048, 1), (32768, 512, 8)]
    gen = torch.Generator(device="cuda").manual_seed(0)
    for (l, c, b) in cases:
        nb = max(2, min(6, int(400e6 // (b * l * HKV * D * 2)) + 1))
        mk = lambda h: torch.randn(b, l, h, D, generator=gen, device="cuda", dtype=torch.bfloat16).transpose(1, 2)
        bufs = [(mk(HQ), mk(HKV), mk(HKV)) for _ in range(nb)]
        pipe = timed(lambda q, k, v: ops.compress(q, k, v, window_size=W, kernel_size=KS, pooling="avgpool", capacity=c), bufs)
        sc = timed(lambda q, k, v: ops.score(q, k, window_size=W, kernel_size=KS, pooling="avgpool"), bufs)
        pb, sb = layer_bytes(b, l, c), score_bytes(b, l)
        print(f"L={l} C={c} B={b}: pipeline {pipe:7.2f} us ({pb / pipe / 1e3 / PEAK:.3f} of {PEAK:.0f} GB/s)  score {sc:7.2f} us "
              f"({sb / sc / 1e3 / PEAK:.3f})  select+compact ~{pipe - sc:6.2f} us", flush=True)
        del bufs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
