#!/usr/bin/env python
"""Benchmark of the TSP + KV-compression hot path (BASELINE.json metric / configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--ctx 32768]

One "step" = the hot path of ONE Llama-3.1-8B-shaped prefill: 32 ``update_kv`` calls
(layers 0..tsp_idx at full context, the TSP layer additionally selecting tsp_len tokens and
gathering the hidden states, the later layers at tsp_len tokens) on synthetic bf16 tensors.

  value    algorithmic GB/s over the whole step with all inputs resident in HBM (CUDA-graph replay
           of the C-ABI calls, per-layer distinct inputs so each step streams > L2 bytes)
  e2e      same metric through the reference-facing ``FastKVCluster.update_kv`` with pinned HOST
           buffers: host->device copies of K / V / Q-window and device->host reads of the results
           inside the timed region
  roofline the dominant kernel stage (observation-window scoring, K1+K2) timed alone with CUDA
           events against MEASURED_PEAKS.json
  cpu_baseline  the oracle port (torch CPU, same ATen ops as the reference) on this box's cores

N > 1 (torchrun): weak scaling — N prompts, every rank owns Hkv/N KV heads of all of them, and
the TSP layer does the path's one exchange before the global top-k: the per-head score rows go
straight into the peers' HBM (fkv_exchange_push / _wait over NVLink peer memory, inside the step's
CUDA graph); NCCL all-gather when no peer memory can be mapped.  ``--impl reference`` times the oracle port only (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "tsp_kv_compress_algorithmic_hbm_throughput"
UNIT = "GB/s"

# Llama-3.1-8B-Instruct shape and BASELINE configs[1] knobs
HQ, HKV, D, HIDDEN, LAYERS, TSP_IDX = 32, 8, 128, 4096, 32, 15
W, KS, POOL = 8, 7, "avgpool"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ctx", type=int, default=32768)
    ap.add_argument("--capacity", type=int, default=2048)
    ap.add_argument("--tsp-len", type=int, default=2048)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "gather", "reduce"],
                    help="N > 1: the TSP layer's exchange: p2p = the library's own stores into peer memory (auto: when peer memory "
                         "can be mapped, else gather), gather = NCCL all-gather of the per-head scores (both bit-identical selection "
                         "at every N), reduce = the north star's single fp32 all-reduce of the partial importance")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = N prompts, Hkv/N heads of each per rank (default); strong = ONE prompt, Hkv/N heads per rank")
    ap.add_argument("--no-overlap", action="store_true", help="run the select + compact stage on the same stream as the scoring")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the 32k / 128k single-call extra")
    ap.add_argument("--no-model", action="store_true")
    ap.add_argument("--score-impl", default="auto", choices=["auto", "simple", "tensor"])
    ap.add_argument("--sweep", default="", metavar="FILE.md",
                    help="instead of the bench line: run the config-5 kernel sweep (ours vs the reference torch path on this GPU) and write a table")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
# algorithmic bytes (stated in DESIGN.md)
# ---------------------------------------------------------------------------------------------
def layer_bytes(b, hkv, hq, l, c, tsp_t, e=2):
    """Bytes one update_kv call must move: K read + Q-window read + score round trip + index write
    + KV compaction (read + write of K and V rows); TSP layer adds the importance round trip,
    the index set and the hidden-state gather."""
    p = l - W
    n = b * hkv * l * D * e                     # K read once
    n += b * hq * W * D * e                     # Q window
    n += 2 * b * hkv * p * 4                    # fp32 score written, read by top-k
    n += b * hkv * (c - W) * 8                  # int64 indices
    n += 2 * 2 * b * hkv * c * D * e            # K and V: C rows read + written
    if tsp_t:
        n += 2 * b * p * 4 + b * tsp_t * 8 + 2 * b * tsp_t * HIDDEN * e
    return n


def score_stage_bytes(b, hkv, hq, l, e=2):
    return b * hkv * l * D * e + b * hq * W * D * e + b * hkv * (l - W) * 4


def step_plan(ctx, cap, tsp_len):
    """[(L, is_tsp_layer)] for the 32 layers of one prefill."""
    plan = []
    for i in range(LAYERS):
        if i < TSP_IDX:
            plan.append((ctx, False))
        elif i == TSP_IDX:
            plan.append((ctx, True))
        else:
            plan.append((tsp_len if ctx > tsp_len else ctx, False))
    return plan


# ---------------------------------------------------------------------------------------------
# clocks sampling
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.idx), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_step_sample(ctx, cap, tsp_len, reps, warmup=1):
    """Times a bounded sample of the step on the CPU with the oracle port (same ATen op sequence as
    baselines/fastkv/utils.py:80-134, fp32): one ordinary full-context layer, the TSP layer with the
    hidden-state gather, and one post-TSP layer.  Returns (GB/s, seconds per sample, description)."""
    from oracle import fastkv_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(42)
    q = torch.randn(1, HQ, ctx, D, generator=g)
    k = torch.randn(1, HKV, ctx, D, generator=g)
    v = torch.randn(1, HKV, ctx, D, generator=g)
    hidden = torch.randn(1, ctx, HIDDEN, generator=g)
    pos = torch.arange(ctx).unsqueeze(0)
    short = min(tsp_len, ctx)
    qs, ks_, vs = q[:, :, :short].contiguous(), k[:, :, :short].contiguous(), v[:, :, :short].contiguous()
    knobs = dict(window_size=W, max_capacity_prompt=cap, kernel_size=KS, pooling=POOL, tsp_length=tsp_len)

    def sample():
        O.fastkv_update_kv(k, q, v, HQ // HKV, tsp_layer=False, **knobs)
        tr = O.fastkv_update_kv(k, q, v, HQ // HKV, tsp_layer=True, **knobs)
        if tr.tsp_idx is not None:
            O.tsp_propagate(hidden, pos, tr.tsp_idx)
        O.fastkv_update_kv(ks_, qs, vs, HQ // HKV, tsp_layer=False, **knobs)

    nbytes = (layer_bytes(1, HKV, HQ, ctx, cap, 0) + layer_bytes(1, HKV, HQ, ctx, cap, tsp_len if ctx > tsp_len else 0)
              + layer_bytes(1, HKV, HQ, short, min(cap, short), 0))
    for _ in range(max(1, warmup)):                     # warm-up (allocator, thread pool): untimed full samples
        sample()
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        sample()
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    desc = (f"3 of the 32 layer calls (1 full-context layer, the TSP layer + hidden gather, 1 post-TSP layer) at "
            f"ctx={ctx}, fp32, median of {reps}")
    return nbytes / med / 1e9, med, desc, cores


def run_reference(args):
    """--impl reference: the reference's CPU path (oracle port, kind='port') on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    reps = max(1, args.steps)
    gbs, sec, desc, cores = cpu_step_sample(args.ctx, args.capacity, args.tsp_len, reps, warmup=max(1, args.warmup))
    # a "step" of this arm is the bounded 3-layer sample (its bytes / its time give the same GB/s metric); the time a
    # full 32-layer step would take at that rate is reported separately
    full_bytes = sum(layer_bytes(1, HKV, HQ, l, args.capacity, args.tsp_len if (t and l > args.tsp_len) else 0)
                     for (l, t) in step_plan(args.ctx, args.capacity, args.tsp_len))
    line = {
        "impl": "reference", "metric": METRIC, "value": round(gbs, 4), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": round(sec * 1e3, 3),
        "ms_per_sample": round(sec * 1e3, 3), "ms_per_full_step_extrapolated": round(full_bytes / (gbs * 1e9) * 1e3, 1),
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": round(gbs, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": round(gbs, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def gpu_torch_baseline(ctx_list, cap, tsp_len, dev, peak):
    """The reference's torch path (oracle port = the same ATen op chain as baselines/fastkv/utils.py:80-134) on CUDA
    tensors of THIS box, bf16 and fp32: the like-for-like "reference on the same GPU" row (a reported baseline)."""
    from oracle import fastkv_oracle as O
    out = []
    gen = torch.Generator(device=dev).manual_seed(7)
    for ctx in ctx_list:
        for dt in (torch.bfloat16, torch.float32):
            try:
                mk = lambda h: torch.randn(1, ctx, h, D, generator=gen, device=dev, dtype=torch.bfloat16).transpose(1, 2).to(dt)
                q, k, v = mk(HQ), mk(HKV), mk(HKV)
                knobs = dict(window_size=W, max_capacity_prompt=cap, kernel_size=KS, pooling=POOL, tsp_length=tsp_len, tsp_layer=False)
                for _ in range(3):
                    O.fastkv_update_kv(k, q, v, HQ // HKV, **knobs)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10
                e0.record()
                for _ in range(reps):
                    O.fastkv_update_kv(k, q, v, HQ // HKV, **knobs)
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / reps
                nb = layer_bytes(1, HKV, HQ, ctx, cap, 0)
                out.append({"ctx": ctx, "dtype": str(dt).replace("torch.", ""), "us_per_call": round(us, 1),
                            "algorithmic_gbs": round(nb / us / 1e3, 1), "frac_of_peak": round(nb / us / 1e3 / peak, 4)})
                del q, k, v
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                out.append({"ctx": ctx, "dtype": str(dt), "error": f"{type(e).__name__}: {e}"})
    return {"what": "oracle port of FastKVCluster.update_kv (same ATen ops as the reference) on CUDA tensors, one non-TSP layer call, "
                    "eager, CUDA events, 10 reps after 3 warm-ups", "rows": out}


def decode_attention_numbers(dev, peak):
    """K5 (split-K decode attention over the compact cache) at the two configured shapes: us per call and HBM GB/s
    (algorithmic bytes 2 * B * Hkv * S * D * 2), CUDA-graph replay over rotating caches larger than L2."""
    from fastkv_b200 import ops
    rows = []
    gen = torch.Generator(device=dev).manual_seed(11)
    for (b, s, smax, what) in ((1, 2304, 2304, "configs[2]: C=2048 + 256 generated"), (8, 640, 768, "configs[3]: B=8, C=512 + 128 generated")):
        nbuf = max(4, int(300e6 // (2 * b * HKV * smax * D * 2)) + 1)
        caches = [(torch.randn(b, HKV, smax, D, generator=gen, device=dev, dtype=torch.bfloat16),
                   torch.randn(b, HKV, smax, D, generator=gen, device=dev, dtype=torch.bfloat16)) for _ in range(nbuf)]
        q = torch.randn(b, HQ, D, generator=gen, device=dev, dtype=torch.bfloat16)
        out = torch.empty_like(q)
        for kc, vc in caches[:2]:
            ops.decode_attn(q, kc, vc, s, out=out)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for kc, vc in caches:
                ops.decode_attn(q, kc, vc, s, out=out)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (5 * nbuf)
        nb = 2 * b * HKV * s * D * 2
        rows.append({"batch": b, "seq_len": s, "config": what, "us_per_call": round(us, 2), "algorithmic_bytes": nb,
                     "gbs": round(nb / us / 1e3, 1), "frac_of_peak": round(nb / us / 1e3 / peak, 4), "caches_rotated": nbuf})
        del caches
        torch.cuda.empty_cache()
    return rows


def _try_full_graph(step):
    """Capture the whole multi-GPU step (NCCL all-gather included) as one CUDA graph; None when that fails."""
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        g.replay()
        torch.cuda.synchronize()
        _try_full_graph.graph = g
        return g
    except Exception as e:  # noqa: BLE001
        print(f"[bench] capturing the NCCL exchange failed ({type(e).__name__}: {e}); using two graphs around an eager TSP layer", file=sys.stderr)
        torch.cuda.synchronize()
        return None


def workload_config(args, n):
    return {
        "workload": (f"BASELINE configs[1]: Llama-3.1-8B-shaped layer tensors (32 q / 8 kv heads, d=128, hidden 4096), "
                     f"{args.ctx}-token prefill bf16, TSP at layer {TSP_IDX} with tsp_len={args.tsp_len}, KV budget "
                     f"{args.capacity}, window {W}, kernel {KS} {POOL}; step = the 32 update_kv calls of one prefill "
                     f"(+ hidden-state gather at the TSP layer)"),
        "ctx": args.ctx, "capacity": args.capacity, "tsp_len": args.tsp_len, "window": W, "kernel_size": KS,
        "pooling": POOL, "prompts": n, "exchange": args.exchange if int(os.environ.get('WORLD_SIZE', '1')) > 1 else None, "sharding": f"kv-heads/{int(os.environ.get('WORLD_SIZE', '1'))}" if int(os.environ.get('WORLD_SIZE', '1')) > 1 else "none",
        "l2": "per-layer distinct K/V/Q buffers: each step streams > 2 GB, far above the 126 MB L2",
    }


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from fastkv_b200 import FastKVCluster, ops, _lib as L
    from fastkv_b200.dist import HeadShard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.load()
    impl = {"auto": L.IMPL_AUTO, "simple": L.IMPL_SIMPLE, "tensor": L.IMPL_TENSOR}[args.score_impl]

    strong = args.scaling == "strong" and world > 1
    n_prompts = 1 if strong else world                  # weak scaling: one prompt per GPU in the batch
    hkv_local, hq_local = HKV // world, HQ // world
    ctx, cap, tsp_len = args.ctx, args.capacity, args.tsp_len
    plan = step_plan(ctx, cap, tsp_len)
    shard = None
    if world > 1:
        if args.exchange in ("auto", "p2p"):
            from fastkv_b200.peer import PeerRegion
            try:                                        # collective: either every rank gets a region or every rank raises
                region = PeerRegion(PeerRegion.bytes_for(n_prompts, HKV, ctx - W))
                shard = HeadShard(exchange="p2p", region=region)
                args.exchange = f"p2p ({region.backend})"
            except RuntimeError as e:
                if args.exchange == "p2p":
                    raise
                print(f"[bench] no peer memory ({e}); falling back to the NCCL all-gather", file=sys.stderr)
                args.exchange = "gather"
        if shard is None:
            shard = HeadShard(exchange=args.exchange)

    gen = torch.Generator(device=dev).manual_seed(42 + rank)

    def rand(b, l, h):
        return torch.randn(b, l, h, D, generator=gen, device=dev, dtype=torch.bfloat16).transpose(1, 2)

    layers = []
    for (l, is_tsp) in plan:
        q, k, v = rand(n_prompts, l, hq_local), rand(n_prompts, l, hkv_local), rand(n_prompts, l, hkv_local)
        cl = FastKVCluster(window_size=W, max_capacity_prompt=cap, kernel_size=KS, pooling=POOL, tsp_layer=is_tsp,
                           tsp_length=tsp_len)
        cl.score_impl = impl
        cl.shard = shard
        layers.append((cl, q, k, v))
    # hidden states: every prompt's rows are propagated by exactly ONE rank (weak: rank r owns prompt r; strong: the
    # single prompt's hidden columns are split over the ranks), so no propagated byte is moved or counted twice
    own_prompt = rank if (world > 1 and not strong) else 0
    hid_cols = HIDDEN // world if strong else HIDDEN
    hidden = torch.randn(1, ctx, hid_cols, generator=gen, device=dev, dtype=torch.bfloat16)
    pos = torch.arange(ctx, device=dev).unsqueeze(0)

    # the select + compact stage of layer i runs on a side stream, overlapping the scoring of layer i + 1 (in a model:
    # the layer's own attention and MLP); everything is joined at the end of the step, inside the timed region
    overlap = None if args.no_overlap else ops.SelectOverlap(dev, partition=True)
    for (cl, q, k, v) in layers:
        cl.overlap = overlap

    def run_layers(sel):
        out = None
        for (cl, q, k, v) in sel:
            ko, vo, tsp = cl.update_kv(k, q, v, None, HQ // HKV, 0)
            if tsp is not None:
                out = ops.gather_rows(hidden, tsp[own_prompt:own_prompt + 1], pos)
        if overlap is not None:
            overlap.join()
        return out

    def one_step():
        return run_layers(layers)

    # algorithmic bytes of the whole job = those of n_prompts single-GPU prefills (nothing replicated is counted)
    total_bytes_step = 0
    for (l, is_tsp) in plan:
        total_bytes_step += layer_bytes(n_prompts, HKV, HQ, l, cap, tsp_len if (is_tsp and l > tsp_len) else 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident inputs, graph replay of the whole step --------------------------
    args.warmup = max(args.warmup, 3)                   # the timing rules need >= 3 warm-up steps; the line reports what ran
    for _ in range(args.warmup):
        one_step()
    torch.cuda.synchronize()
    c0 = lib.fkv_launch_count()
    one_step()
    launches_per_step = lib.fkv_launch_count() - c0
    # CUDA-graph replay so that host launch latency is not what gets timed.  N == 1: the whole step is
    # one graph.  N > 1: the layers without a collective are captured as two graphs around the TSP
    # layer, whose NCCL exchange runs eagerly (capturing NCCL hung at teardown in testing).
    graph = None
    timed_with = "eager C-ABI launches"
    run = one_step
    if not args.no_graph:
        try:
            if world == 1 or args.exchange.startswith("p2p"):
                # (N > 1 with the p2p exchange: nothing in the step leaves the device, so the whole step is one graph too;
                # the ranks meet inside it, at the exchange's flags)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    one_step()
                graph.replay()
                torch.cuda.synchronize()
                run = graph.replay
                timed_with = "cuda-graph replay of the C-ABI launches" + (" (peer-memory score exchange inside the graph)" if world > 1 else "")
            elif os.environ.get("FKV_BENCH_NCCL_IN_GRAPH", "0") == "1" and _try_full_graph(one_step) is not None:
                graph = _try_full_graph.graph
                run = graph.replay
                timed_with = "cuda-graph replay of the whole step, the TSP layer's NCCL exchange captured in the graph"
            else:
                tsp_pos = [i for i, (l, t) in enumerate(plan) if t][0]
                pre, mid, post = layers[:tsp_pos], layers[tsp_pos:tsp_pos + 1], layers[tsp_pos + 1:]
                g_pre, g_post = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_pre):
                    run_layers(pre)
                with torch.cuda.graph(g_post):
                    run_layers(post)
                graph = (g_pre, g_post)

                def run():
                    g_pre.replay()
                    run_layers(mid)
                    g_post.replay()
                run()
                torch.cuda.synchronize()
                timed_with = "cuda-graph replay (2 graphs) + eager TSP layer with the NCCL exchange"
        except Exception as e:  # noqa: BLE001
            print(f"[bench] graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
            graph = None
            run = one_step
            timed_with = "eager C-ABI launches"
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        run()
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = t.item() / args.steps
    value = total_bytes_step / (ms_step * 1e-3) / 1e9

    # ---- roofline: scoring stage (K1+K2) timed alone, full-context layers, rotating inputs -------
    full_layers = [(cl, q, k, v) for (cl, q, k, v), (l, _) in zip(layers, plan) if l == ctx]
    for (cl, q, k, v) in full_layers[:3]:
        ops.score(q, k, window_size=W, kernel_size=KS, pooling=POOL, impl=impl)
    torch.cuda.synchronize()
    c0 = lib.fkv_launch_count()
    ops.score(full_layers[0][1], full_layers[0][2], window_size=W, kernel_size=KS, pooling=POOL, impl=impl)
    score_kernels = lib.fkv_launch_count() - c0
    torch.cuda.synchronize()
    # one pass over all full-context layers (distinct K per layer: > L2 bytes between re-reads),
    # replayed as a CUDA graph so that host launch latency is not what gets timed
    def score_pass():
        for (cl, q, k, v) in full_layers:
            ops.score(q, k, window_size=W, kernel_size=KS, pooling=POOL, impl=impl)
    sgraph = None
    if not args.no_graph:
        try:
            sgraph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(sgraph):
                score_pass()
            sgraph.replay()
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"[bench] score graph capture failed ({type(e).__name__}: {e})", file=sys.stderr)
            sgraph = None
            torch.cuda.synchronize()
    srun = sgraph.replay if sgraph is not None else score_pass
    sev0, sev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sreps = max(3, args.steps // 2)
    sev0.record()
    for _ in range(sreps):
        srun()
    sev1.record()
    torch.cuda.synchronize()
    reps = sreps * len(full_layers)
    score_us = sev0.elapsed_time(sev1) * 1e3 / reps
    clocks = sampler.stop() if rank == 0 else None

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    sbytes = score_stage_bytes(n_prompts, hkv_local, hq_local, ctx)
    achieved = sbytes / (score_us * 1e-6) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "score_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get(str(ctx))
        except Exception:  # noqa: BLE001
            traffic = None
    roofline = {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "kernel": "fkv_score stage (K1+K2)",
                "kernels_per_call": int(score_kernels), "us_per_call": round(score_us, 2),
                "algorithmic_bytes_per_call": sbytes, "peak_source": peak_src}

    # ---- extra: one update_kv call (score -> top-k -> compaction) and its scoring stage at 32k and 128k,
    # the two ends of the north star's range; graph replay over rotating inputs, rank 0 at N = 1 only
    context_sweep = None
    if world == 1 and not args.no_sweep:
        context_sweep = []
        for sl in (32768, 131072):
            try:
                nb_ = 3 if sl <= 32768 else 2
                bufs_ = [(rand(1, sl, HQ), rand(1, sl, HKV), rand(1, sl, HKV)) for _ in range(nb_)]
                rec = {"ctx": sl}
                for name, fn in (("pipeline_us", lambda q, k, v: ops.compress(q, k, v, window_size=W, kernel_size=KS, pooling=POOL, capacity=args.capacity)),
                                 ("score_us", lambda q, k, v: ops.score(q, k, window_size=W, kernel_size=KS, pooling=POOL))):
                    for x in bufs_:
                        fn(*x)
                    torch.cuda.synchronize()
                    g_ = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g_):
                        for x in bufs_:
                            fn(*x)
                    g_.replay()
                    torch.cuda.synchronize()
                    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a0.record()
                    for _ in range(10):
                        g_.replay()
                    a1.record()
                    torch.cuda.synchronize()
                    rec[name] = round(a0.elapsed_time(a1) * 1e3 / (10 * nb_), 2)
                    del g_
                pb_, sb_ = layer_bytes(1, HKV, HQ, sl, args.capacity, 0), score_stage_bytes(1, HKV, HQ, sl)
                rec.update({"pipeline_algorithmic_bytes": pb_, "pipeline_frac": round(pb_ / (rec["pipeline_us"] * 1e-6) / 1e9 / peak, 4),
                            "score_algorithmic_bytes": sb_, "score_frac": round(sb_ / (rec["score_us"] * 1e-6) / 1e9 / peak, 4)})
                context_sweep.append(rec)
                del bufs_
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                context_sweep.append({"ctx": sl, "error": f"{type(e).__name__}: {e}"})

    # ---- e2e: FastKVCluster.update_kv on pinned HOST buffers -----------------------------------
    e2e = None
    if not args.no_e2e:
        host_layers = []
        q_host = {}
        for (cl, q, k, v), (l, is_tsp) in zip(layers, plan):
            if l not in q_host:                         # only the window rows of Q are ever read
                q_host[l] = q.contiguous().cpu().pin_memory()
            host_layers.append((cl, q_host[l], k.contiguous().cpu().pin_memory(), v.contiguous().cpu().pin_memory()))
        hidden_host = hidden.cpu().pin_memory()
        h2d = d2h = 0
        es = 2
        for (l, is_tsp) in plan:
            # K in full, the Q window, and only the C gathered rows of V (read in place from pinned memory)
            h2d += n_prompts * (hkv_local * l * D + hq_local * W * D + hkv_local * min(cap, l) * D) * es
            d2h += 2 * n_prompts * hkv_local * min(cap, l) * D * es
            if is_tsp and l > tsp_len:
                h2d += n_prompts * tsp_len * 8 + tsp_len * hid_cols * es       # index set + the T gathered hidden rows of the own prompt / column shard
                d2h += n_prompts * tsp_len * 8 + tsp_len * (8 + hid_cols * es)

        # calls on host tensors are pipelined over two streams (FastKVCluster.host_pipeline: non-blocking results, joined at
        # the end of the step, all inside the timed region): the H2D copy of layer i + 1 overlaps the kernels and the
        # D2H copy of layer i.  The blocking form of the same API is timed next to it.
        pipe = ops.HostPipeline(dev, depth=int(os.environ.get("FKV_E2E_DEPTH", "4")))
        # the caller's host-side KV cache: one pinned (k, v) slot per layer, allocated once (FastKVCluster.out_provider)
        host_cache = [tuple(torch.empty((n_prompts, hkv_local, min(cap, l), D), dtype=torch.bfloat16, pin_memory=True) for _ in range(2))
                      for (l, _t) in plan]
        for i, (cl, _q, _k, _v) in enumerate(host_layers):
            cl.out_provider = (lambda *a, _slot=host_cache[i]: _slot)

        def host_step(pipelined=True):
            res = None
            for (cl, q, k, v) in host_layers:
                cl.host_pipeline = pipe if pipelined else None
                ko, vo, tsp = cl.update_kv(k, q, v, None, HQ // HKV, 0)
                if tsp is not None:
                    # hidden states stay in pinned host memory: only the T selected rows cross the bus.  Pipelined: the
                    # gather is enqueued on the stream of the call that produces the index set (stream order, no host wait)
                    with torch.cuda.stream(pipe.last_stream() if pipelined else torch.cuda.current_stream()):
                        nh, npos = ops.gather_rows(hidden_host, tsp[own_prompt:own_prompt + 1].to(dev, non_blocking=True), pos)
                        if pipelined:
                            res = tuple(torch.empty(t.shape, dtype=t.dtype, pin_memory=True).copy_(t, non_blocking=True) for t in (nh, npos))
                        else:
                            res = (nh.cpu(), npos.cpu())
            if pipelined:
                pipe.join()                             # every layer's compact K/V is on the host
            return res

        def time_host(pipelined):
            for _ in range(int(os.environ.get("FKV_E2E_WARMUP", "2"))):
                host_step(pipelined)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                t1 = time.perf_counter()
                host_step(pipelined)
                if os.environ.get("FKV_E2E_VERBOSE") == "1":
                    torch.cuda.synchronize()
                    print(f"[bench] e2e step pipelined={pipelined}: {(time.perf_counter() - t1) * 1e3:.2f} ms", file=sys.stderr)
            barrier()
            dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return dt.item() / e_steps

        e_steps = max(2, min(args.steps, 5))
        sec_blocking = time_host(False)
        sec = time_host(True)
        for (cl, _q, _k, _v) in host_layers:
            cl.host_pipeline = cl.out_provider = None
        e_val = total_bytes_step / sec / 1e9
        e2e = {"value": round(e_val, 2), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": round(sec * 1e3, 3), "steps": e_steps,
               "pcie_gbs": {"h2d": round(h2d / sec / 1e9, 1), "d2h": round(d2h / sec / 1e9, 1)},
               "blocking_calls": {"value": round(total_bytes_step / sec_blocking / 1e9, 2), "ms_per_step": round(sec_blocking * 1e3, 3)},
               "api": f"FastKVCluster.update_kv(key_states, query_states, value_states, ...) on pinned host tensors, compact K/V delivered into "
                      f"the caller's pinned per-layer host cache (cluster.out_provider); calls pipelined over {len(pipe.streams)} streams "
                      f"(cluster.host_pipeline, results complete at the join that ends the step); blocking_calls = the same API with "
                      f"every call synchronising before it returns"}

    # ---- the reference's torch path on this GPU + K5 numbers (rank 0, N == 1 only) ------------------
    gpu_torch = k5 = None
    if rank == 0 and world == 1 and not args.no_sweep:
        try:
            gpu_torch = gpu_torch_baseline((32768, 131072), cap, tsp_len, dev, peak)
        except Exception as e:  # noqa: BLE001
            gpu_torch = {"error": f"{type(e).__name__}: {e}"}
        try:
            k5 = decode_attention_numbers(dev, peak)
        except Exception as e:  # noqa: BLE001
            k5 = {"error": f"{type(e).__name__}: {e}"}

    # ---- cpu baseline (rank 0, N == 1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gbs, sec, desc, cores = cpu_step_sample(ctx, cap, tsp_len, 5)
        cpu = {"value": round(gbs, 4), "unit": UNIT, "cores": cores, "kind": "port", "sample": desc,
               "seconds_per_sample": round(sec, 4)}

    # ---- model-level extras (N == 1): prefill TTFT and decode tok/s of a random-init Llama-3.1-8B-shaped
    # model through the monkeypatch (the other half of BASELINE.json's metric), reference protocol
    model_extras = None
    if rank == 0 and world == 1 and not args.no_model and os.environ.get("FKV_BENCH_MODEL", "1") != "0":
        try:
            del layers, hidden
            if not args.no_e2e:
                del host_layers, hidden_host, q_host
            graph = sgraph = None
            torch.cuda.empty_cache()
            import types as _types
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import model_bench
            margs = _types.SimpleNamespace(model="llama", method="fastkv", ctx=ctx, batch=1, genlen=64, capacity=cap,
                                           tsp_len=tsp_len, tsp_idx=TSP_IDX, layers=0, runs=2, graph=True)
            import io, contextlib
            with contextlib.redirect_stdout(io.StringIO()):
                model_extras = model_bench.run(margs)
        except Exception as e:  # noqa: BLE001
            model_extras = {"error": f"{type(e).__name__}: {e}"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 4), "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args, n_prompts),
            "algorithmic_bytes_per_step": int(total_bytes_step),
            "timed_with": timed_with + ("" if args.no_overlap else "; select + compact stage of layer i on a side stream (joined at step end)"),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "roofline": roofline, "cpu_baseline": cpu, "gpu_torch_baseline": gpu_torch, "decode_attention": k5,
            "context_sweep": context_sweep, "model": model_extras,
        }
        print(json.dumps(line))
    if world > 1:
        graph = sgraph = None                           # graphs holding NCCL work go before the communicator
        if hasattr(_try_full_graph, "graph"):
            del _try_full_graph.graph
        torch.cuda.synchronize()
        if shard is not None and shard.region is not None:
            shard.region.close()                        # collective (barrier inside): nobody unmaps while a peer may still store
        dist.barrier()
        dist.destroy_process_group()


def run_sweep(args):
    """BASELINE configs[4]: kernel sweep over context x budget x window / kernel / pooling x selection mode, every row next
    to the reference's torch path (oracle port, same ATen ops) timed on THIS GPU in bf16."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sweep as S
    from oracle import fastkv_oracle as O
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0

    def torch_us(c):
        g = torch.Generator(device=dev).manual_seed(1)
        dt = torch.float32 if c.get("dtype") == "f32" else torch.bfloat16
        mk = lambda h: torch.randn(c["b"], c["l"], h, D, generator=g, device=dev, dtype=torch.bfloat16).transpose(1, 2).to(dt)
        q, k, v = mk(HQ), mk(HKV), mk(HKV)
        if c["mode"] == "fastkv":
            fn = lambda: O.fastkv_update_kv(k, q, v, HQ // HKV, window_size=c["w"], max_capacity_prompt=c["c"], kernel_size=c["ks"],
                                            pooling=c["pool"], tsp_layer=False)
        elif c["mode"] == "snapkv":
            fn = lambda: O.snapkv_update_kv(O.expand_kv_heads(k, HQ // HKV), q, O.expand_kv_heads(v, HQ // HKV), window_size=c["w"],
                                            max_capacity_prompt=c["c"], kernel_size=c["ks"], pooling=c["pool"])
        else:
            fn = lambda: O.gemfilter_select(q[:1], k[:1], HQ // HKV, c["c"])
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / 5

    rows = [f"| mode | B | L | C | W | kernel | pool | ours: pipeline µs | ours: score stage µs | reference torch on this GPU µs (same dtype, eager) | speed-up | "
            f"algorithmic MB | ours GB/s | frac of {peak:.0f} GB/s |", "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for c in S.grid(quick=False):
        tot, sc = S.run_one(**c)
        try:
            ref = torch_us(c)
        except Exception as e:  # noqa: BLE001
            ref = float("nan")
            print(f"[sweep] torch path failed for {c}: {type(e).__name__}: {e}", file=sys.stderr)
        torch.cuda.empty_cache()
        nb = S.alg_bytes(c["mode"], c["b"], c["l"], c["c"], c["w"], e=4 if c.get("dtype") == "f32" else 2)
        gbs = nb / tot / 1e3
        rows.append(f"| {c['mode']}{' fp32' if c.get('dtype') == 'f32' else ''} | {c['b']} | {c['l']} | {c['c']} | {c['w']} | {c['ks']} | {c['pool']} | {tot:.1f} | {sc:.1f} | {ref:.0f} | "
                    f"{ref / tot:.1f}x | {nb / 1e6:.1f} | {gbs:.0f} | {gbs / peak:.2f} |")
        print(rows[-1], flush=True)
    os.makedirs(os.path.dirname(args.sweep) or ".", exist_ok=True)
    with open(args.sweep, "w") as f:
        f.write("\n".join(rows) + "\n")


def main():
    args = parse()
    if args.sweep:
        run_sweep(args)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
