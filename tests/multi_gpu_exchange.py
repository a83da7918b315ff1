"""Two (or more) processes, one per GPU: the peer-memory score exchange against the NCCL all-gather, eagerly, with a
straggling rank, inside a CUDA graph, and through the KV-head-sharded FastKVCluster.  Launched by torchrun
(tests/test_gpu_exchange.py::test_two_process_exchange, tools/r2_gpu_job4.sh); prints EXCHANGE OK on rank 0."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    from fastkv_b200 import FastKVCluster, ops, _lib as L
    from fastkv_b200.dist import HeadShard
    from fastkv_b200.peer import PeerRegion

    b, hl, p = 1, 8 // world, 32760
    backends = [os.environ["FKV_PEER_BACKEND"]] if os.environ.get("FKV_PEER_BACKEND") else ["symm", "ipc"]
    report = {}
    for backend in backends:
        try:
            reg = PeerRegion(PeerRegion.bytes_for(b, hl * world, p), backend=backend)
        except RuntimeError as e:
            report[backend] = f"unavailable ({e})"
            continue
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        for it in range(6):
            s = torch.rand(b, hl, p, generator=g, device=dev)
            flat = torch.empty(world * b, hl, p, device=dev)
            dist.all_gather_into_tensor(flat, s)
            want = flat.view(world, b, hl, p).permute(1, 0, 2, 3).reshape(b, world * hl, p)
            if it % 2 == 1 and rank == world - 1:
                time.sleep(0.2)                          # a straggler: the others spin in the handshake / the wait
            reg.push(s)
            got = reg.wait()
            torch.cuda.synchronize()
            assert torch.equal(got, want), (backend, it)
            assert reg.status() == 0
        # latency: exchange alone, ranks in step (eager launches; graph replay below)
        s = torch.rand(b, hl, p, generator=g, device=dev)
        flat = torch.empty(world * b, hl, p, device=dev)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for _ in range(5):
            reg.push(s); reg.wait()
            dist.all_gather_into_tensor(flat, s)
        graph = torch.cuda.CUDAGraph()
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            with torch.cuda.graph(graph, stream=st):
                for _ in range(20):
                    reg.push(s); reg.wait()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            graph.replay()
            ev[0].record(st)
            graph.replay()
            ev[1].record(st)
            torch.cuda.synchronize()
        dist.barrier(); torch.cuda.synchronize()
        ev[2].record()
        for _ in range(20):
            dist.all_gather_into_tensor(flat, s)
        ev[3].record()
        torch.cuda.synchronize()
        assert reg.status() == 0
        report[backend] = {"p2p_us_per_exchange_graph": round(ev[0].elapsed_time(ev[1]) * 1e3 / 20, 2),
                           "nccl_all_gather_us_eager": round(ev[2].elapsed_time(ev[3]) * 1e3 / 20, 2),
                           "bytes_per_rank": b * hl * p * 4}
        del graph
        reg.close()
    if not any(isinstance(v, dict) for v in report.values()):
        raise SystemExit(f"no peer-memory backend works here: {report}")

    # the sharded compressor: p2p vs NCCL gather give the same TSP indices, eagerly and from a replayed graph
    lq, cap, tsp_len, w = 8192, 512, 1024, 8
    gq = torch.Generator(device=dev).manual_seed(5)     # same full tensors on every rank, each takes its head slice
    q = torch.randn(1, lq, 32, 128, generator=gq, device=dev, dtype=torch.bfloat16).transpose(1, 2)
    k = torch.randn(1, lq, 8, 128, generator=gq, device=dev, dtype=torch.bfloat16).transpose(1, 2)
    v = torch.randn(1, lq, 8, 128, generator=gq, device=dev, dtype=torch.bfloat16).transpose(1, 2)
    h0, h1 = rank * hl, (rank + 1) * hl
    ql, kl, vl = q[:, 4 * h0:4 * h1], k[:, h0:h1], v[:, h0:h1]
    res = {}
    for mode in ("gather", "p2p"):
        cl = FastKVCluster(window_size=w, max_capacity_prompt=cap, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=tsp_len)
        cl.shard = HeadShard(exchange=mode)
        ko, vo, tsp = cl.update_kv(kl, ql, vl, None, 4, 0)
        torch.cuda.synchronize()
        res[mode] = (ko.clone(), vo.clone(), tsp.clone())
        if mode == "p2p":
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                cl.update_kv(kl, ql, vl, None, 4, 0)
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=st):
                    out = cl.update_kv(kl, ql, vl, None, 4, 0)
                for _ in range(3):
                    graph.replay()
                torch.cuda.synchronize()
            assert torch.equal(out[2], res["gather"][2]) and torch.equal(out[0], res["gather"][0])
            assert cl.shard.region.status() == 0
            report["sharded_cluster_backend"] = cl.shard.region.backend
            del graph
    assert torch.equal(res["p2p"][2], res["gather"][2])
    assert torch.equal(res["p2p"][0], res["gather"][0]) and torch.equal(res["p2p"][1], res["gather"][1])
    single = FastKVCluster(window_size=w, max_capacity_prompt=cap, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=tsp_len)
    _, _, tsp1 = single.update_kv(k, q, v, None, 4, 0)
    torch.cuda.synchronize()
    assert torch.equal(tsp1, res["p2p"][2])             # identical to the unsharded selection
    # the small all-reduce (tensor-parallel decode): == fp32 rank-order sum, same bits on every rank, graph replay, latency
    n = 4096
    ar = PeerRegion(PeerRegion.allreduce_bytes_for(world, 64 << 10))
    gen = torch.Generator(device=dev).manual_seed(7)     # same stream of numbers on every rank -> every rank knows all partials
    for it in range(4):
        parts = [torch.randn(n, generator=gen, device=dev).bfloat16() for _ in range(world)]
        want = torch.zeros(n, device=dev)
        for p_ in parts:
            want += p_.float()
        mine = parts[rank].clone()
        if it == 2 and rank == 0:
            time.sleep(0.1)
        ar.all_reduce_(mine)
        torch.cuda.synchronize()
        assert torch.equal(mine, want.bfloat16()), ("allreduce", it)
    x = torch.randn(n, device=dev).bfloat16()
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        ar.all_reduce_(x)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for _ in range(64):
                ar.all_reduce_(x)
        dist.barrier(); torch.cuda.synchronize()
        graph.replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        graph.replay()
        e1.record(st)
        torch.cuda.synchronize()
    y = torch.randn(n, device=dev).bfloat16()
    for _ in range(5):
        dist.all_reduce(y)
    torch.cuda.synchronize()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(64):
        dist.all_reduce(y)
    e3.record()
    torch.cuda.synchronize()
    assert ar.status() == 0
    report["allreduce_8KiB"] = {"peer_us_graph": round(e0.elapsed_time(e1) * 1e3 / 64, 2), "nccl_us_eager": round(e2.elapsed_time(e3) * 1e3 / 64, 2)}
    del graph
    ar.close()
    torch.cuda.synchronize()
    dist.barrier()
    if rank == 0:
        print("EXCHANGE OK", report, flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
