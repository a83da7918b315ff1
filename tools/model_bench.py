"""End-to-end TTFT / decode tok/s of a random-init Llama-3.1-8B-shaped (or Ministral-8B-shaped) model
through the fastkv_b200 monkeypatch, mirroring the reference protocol (benchmark/e2e.py:55-56,
:162-183: random-init weights here, all-ones token ids, batch 1, greedy hand-rolled decode loop,
CUDA events, 1 warm-up).

    python tools/model_bench.py --ctx 32768 --genlen 64 [--method fastkv|fullkv] [--model llama|ministral]

Prints one JSON line.  Used by bench.py (extras) and for the numbers quoted in profiles/.
"""
import argparse
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def build_model(kind, layers=None, dtype=torch.bfloat16, device="cuda", tp=1):
    """``tp`` > 1: this rank's tensor-parallel SHARD of the model (heads and MLP columns divided by tp; the all-reduces
    behind o_proj / down_proj are added by ``add_tp_allreduce``).  The KV heads a rank owns are exactly the heads its
    FastKV compressor scores, selects and compacts."""
    from transformers import LlamaConfig, LlamaForCausalLM, MistralConfig, MistralForCausalLM
    if kind == "llama":       # Llama-3.1-8B-Instruct public model-card shape
        cfg = LlamaConfig(hidden_size=4096, intermediate_size=14336 // tp, num_hidden_layers=layers or 32, num_attention_heads=32 // tp,
                          num_key_value_heads=8 // tp, head_dim=128, vocab_size=128256, max_position_embeddings=131072,
                          rope_theta=500000.0, attn_implementation="sdpa")
        cls = LlamaForCausalLM
    else:                     # Ministral-8B-Instruct-2410 shape (sliding_window 32768: a no-op up to 32k prompts)
        cfg = MistralConfig(hidden_size=4096, intermediate_size=12288 // tp, num_hidden_layers=layers or 36, num_attention_heads=32 // tp,
                            num_key_value_heads=8 // tp, head_dim=128, vocab_size=131072, max_position_embeddings=131072,
                            sliding_window=32768, attn_implementation="sdpa")
        cls = MistralForCausalLM
    torch.manual_seed(0)
    with torch.device(device):
        model = cls(cfg).to(dtype)
    for p in model.parameters():                      # N(0, 0.02) random init (no checkpoints offline)
        p.data.normal_(0.0, 0.02)
    return model.eval()


def add_tp_allreduce(model, peer=None):
    """Row-parallel projections: sum the partial outputs of o_proj and down_proj over the ranks (2 all-reduces per
    layer; outside the hot path).  Embedding, norms and lm_head are replicated.  Prefill-sized tensors go through library
    NCCL; with ``peer`` (a PeerRegion) the small decode-step sums use the library's own peer-memory all-reduce
    (fkv_allreduce_sum), which can be captured in the decode graph."""
    import torch.distributed as dist
    cap = peer.allreduce_capacity() if peer is not None else 0

    def hook(_mod, _inp, out):
        if peer is not None and out.numel() * out.element_size() <= cap and out.is_contiguous() and (out.numel() * out.element_size()) % 16 == 0:
            peer.all_reduce_(out)
        else:
            dist.all_reduce(out)
        return out

    for layer in model.model.layers:
        layer.self_attn.o_proj.register_forward_hook(hook)
        layer.mlp.down_proj.register_forward_hook(hook)


def run(args):
    from fastkv_b200 import patch
    tsp_idx = args.tsp_idx if args.tsp_idx >= 0 else (15 if args.model == "llama" else 17)
    if args.model == "llama":
        patch.replace_llama(args.method)
    else:
        patch.replace_mistral(args.method)
    tp = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if tp > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    model = build_model(args.model, layers=args.layers, tp=tp)
    peer_ar = None
    if tp > 1:
        if os.environ.get("FKV_PEER_ALLREDUCE", "1") != "0":
            from fastkv_b200.peer import PeerRegion
            try:
                peer_ar = PeerRegion(PeerRegion.allreduce_bytes_for(tp, 256 << 10))
            except RuntimeError as e:
                print(f"[model_bench] no peer memory ({e}); decode all-reduces stay on NCCL (eager)", file=sys.stderr)
        add_tp_allreduce(model, peer_ar)
    a = types.SimpleNamespace(method=args.method, max_capacity_prompts=args.capacity, window_size=8, kernel_size=7,
                              pooling="avgpool", retain_rate=0.1, eviction_mode="constant", tsp_len=args.tsp_len,
                              tsp_rate=0.2, tsp_idx=tsp_idx, merge=None)
    patch.set_model(model, a)
    if tp > 1 and args.method == "fastkv":
        from fastkv_b200.dist import HeadShard
        shard = HeadShard(exchange=os.environ.get("FKV_EXCHANGE", "gather"))   # the path's one exchange: per-head scores at the TSP layer
        for layer in model.model.layers:
            layer.self_attn.kv_cluster.shard = shard
    ids = torch.ones((args.batch, args.ctx), dtype=torch.int64, device="cuda")      # benchmark/e2e.py:55
    if getattr(args, "dense_tsp_layer", False):
        model.config.fastkv_sparse_tsp_layer = False
    model.config.fastkv_max_new_tokens = args.genlen
    model.model.config.fastkv_max_new_tokens = args.genlen
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    results = []
    with torch.no_grad():
        for it in range(args.runs + 1):               # first iteration is the warm-up
            s, e = ev(), ev()
            s.record()
            out = model(ids, use_cache=True, logits_to_keep=1)
            e.record()
            torch.cuda.synchronize()
            ttft = s.elapsed_time(e)
            cache = out.past_key_values
            tok = out.logits[:, -1, :].argmax(dim=-1, keepdim=True)
            s2, e2 = ev(), ev()
            if args.graph and args.method != "fullkv" and (tp == 1 or peer_ar is not None):   # (NCCL inside a captured graph hangs on this stack; the peer-memory all-reduce is capturable)
                gd = patch.GraphedDecode(model, cache, tok)          # 2 eager steps + capture (untimed)
                n_timed = args.genlen - 1 - 2
                s2.record()
                gd.generate(n_timed)
                e2.record()
                torch.cuda.synchronize()
                dec = s2.elapsed_time(e2) * (args.genlen - 1) / n_timed   # per-token time scaled to genlen-1 steps
            else:
                s2.record()
                for _ in range(args.genlen - 1):
                    out = model(input_ids=tok, past_key_values=cache, use_cache=True)
                    tok = out.logits[:, -1, :].argmax(dim=-1, keepdim=True)
                e2.record()
                torch.cuda.synchronize()
                dec = s2.elapsed_time(e2)
            if it > 0:
                results.append((ttft, dec))
            del out, cache
    ttft = sum(r[0] for r in results) / len(results)
    dec = sum(r[1] for r in results) / len(results)
    if tp > 1:                                        # max over ranks, like bench.py
        import torch.distributed as dist
        t = torch.tensor([ttft, dec], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ttft, dec = t[0].item(), t[1].item()
    line = {
        "n_gpus": tp, "parallelism": f"tp{tp} (kv-head-sharded compressor, 2 hidden all-reduces per layer)" if tp > 1 else "none",
        "model": args.model, "layers": args.layers or (32 if args.model == "llama" else 36), "method": args.method,
        "ctx": args.ctx, "batch": args.batch, "capacity": args.capacity, "tsp_len": args.tsp_len, "tsp_idx": tsp_idx,
        "genlen": args.genlen, "tsp_layer_rows": "all" if getattr(args, "dense_tsp_layer", False) else "kept only", "decode_cuda_graph": bool(args.graph and args.method != "fullkv" and (tp == 1 or peer_ar is not None)),
        "decode_allreduce": ("peer memory (fkv_allreduce_sum)" if peer_ar is not None else "nccl") if tp > 1 else None, "ttft_ms": round(ttft, 2), "decode_ms_per_token": round(dec / max(args.genlen - 1, 1), 3),
        "decode_tok_s": round(args.batch * (args.genlen - 1) / (dec / 1e3), 1),
        "reference_throughput_tok_s": round((args.genlen - 1) / ((ttft + dec) / 1e3), 2),   # benchmark/e2e.py:183 definition
        "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2), "data": "random-init weights, all-ones ids",
    }
    if rank == 0:
        print(json.dumps(line))
    if tp > 1:
        import torch.distributed as dist
        torch.cuda.synchronize()
        if peer_ar is not None:
            peer_ar.close()
        dist.barrier()
        dist.destroy_process_group()
    return line


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="llama", choices=["llama", "ministral"])
    ap.add_argument("--method", default="fastkv", choices=["fastkv", "fullkv", "snapkv"])
    ap.add_argument("--ctx", type=int, default=32768)
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--genlen", type=int, default=64)
    ap.add_argument("--capacity", type=int, default=2048)
    ap.add_argument("--tsp-len", type=int, default=2048)
    ap.add_argument("--tsp-idx", type=int, default=-1)
    ap.add_argument("--layers", type=int, default=0)
    ap.add_argument("--runs", type=int, default=2)
    ap.add_argument("--graph", action="store_true", help="CUDA-graph the decode step")
    ap.add_argument("--dense-tsp-layer", action="store_true",
                    help="the reference's order at the TSP layer: attention / MLP for all L rows, gather afterwards (default: kept rows only)")
    run(ap.parse_args())
