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


def build_model(kind, layers=None, dtype=torch.bfloat16, device="cuda"):
    from transformers import LlamaConfig, LlamaForCausalLM, MistralConfig, MistralForCausalLM
    if kind == "llama":       # Llama-3.1-8B-Instruct public model-card shape
        cfg = LlamaConfig(hidden_size=4096, intermediate_size=14336, num_hidden_layers=layers or 32, num_attention_heads=32,
                          num_key_value_heads=8, head_dim=128, vocab_size=128256, max_position_embeddings=131072,
                          rope_theta=500000.0, attn_implementation="sdpa")
        cls = LlamaForCausalLM
    else:                     # Ministral-8B-Instruct-2410 shape
        cfg = MistralConfig(hidden_size=4096, intermediate_size=12288, num_hidden_layers=layers or 36, num_attention_heads=32,
                            num_key_value_heads=8, head_dim=128, vocab_size=131072, max_position_embeddings=131072,
                            sliding_window=None, attn_implementation="sdpa")
        cls = MistralForCausalLM
    torch.manual_seed(0)
    with torch.device(device):
        model = cls(cfg).to(dtype)
    for p in model.parameters():                      # N(0, 0.02) random init (no checkpoints offline)
        p.data.normal_(0.0, 0.02)
    return model.eval()


def run(args):
    from fastkv_b200 import patch
    tsp_idx = args.tsp_idx if args.tsp_idx >= 0 else (15 if args.model == "llama" else 17)
    if args.model == "llama":
        patch.replace_llama(args.method)
    else:
        patch.replace_mistral(args.method)
    model = build_model(args.model, layers=args.layers)
    a = types.SimpleNamespace(method=args.method, max_capacity_prompts=args.capacity, window_size=8, kernel_size=7,
                              pooling="avgpool", retain_rate=0.1, eviction_mode="constant", tsp_len=args.tsp_len,
                              tsp_rate=0.2, tsp_idx=tsp_idx, merge=None)
    patch.set_model(model, a)
    ids = torch.ones((args.batch, args.ctx), dtype=torch.int64, device="cuda")      # benchmark/e2e.py:55
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
            if args.graph and args.method != "fullkv":
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
    line = {
        "model": args.model, "layers": args.layers or (32 if args.model == "llama" else 36), "method": args.method,
        "ctx": args.ctx, "batch": args.batch, "capacity": args.capacity, "tsp_len": args.tsp_len, "tsp_idx": tsp_idx,
        "genlen": args.genlen, "decode_cuda_graph": bool(args.graph and args.method != "fullkv"), "ttft_ms": round(ttft, 2), "decode_ms_per_token": round(dec / max(args.genlen - 1, 1), 3),
        "decode_tok_s": round(args.batch * (args.genlen - 1) / (dec / 1e3), 1),
        "reference_throughput_tok_s": round((args.genlen - 1) / ((ttft + dec) / 1e3), 2),   # benchmark/e2e.py:183 definition
        "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2), "data": "random-init weights, all-ones ids",
    }
    print(json.dumps(line))
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
    run(ap.parse_args())
