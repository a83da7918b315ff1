"""Timeline of one peer-memory score exchange between real GPUs (%globaltimer stamps written by the kernels of an
FKV_BUILD_XTRACE=1 build into the region header).

    FKV_BUILD_XTRACE=1 python fastkv_b200/_build.py
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/exchange_trace.py [--heads 4 --p 32760]

Stamps (CTA 0 of the push kernel, then the wait kernel): 0 push enters | 1 handshake done | 2 stores issued |
3 fence done, DATA counters bumped | 6 wait kernel enters | 7 every peer's CTAs counted.  (4, 5: unused.)
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--heads", type=int, default=0)
    ap.add_argument("--p", type=int, default=32760)
    ap.add_argument("--b", type=int, default=1)
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", torch.cuda.current_device())
    dist.init_process_group("nccl", device_id=dev)
    from fastkv_b200.peer import PeerRegion
    hl = a.heads or 8 // world
    reg = PeerRegion(PeerRegion.bytes_for(a.b, hl * world, a.p))
    s = torch.rand(a.b, hl, a.p, device=dev)
    st = torch.cuda.Stream()
    filler = torch.empty(64 << 20, device=dev, dtype=torch.uint8)
    rows = []
    with torch.cuda.stream(st):
        for _ in range(3):
            reg.push(s)
            reg.wait()
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            filler.zero_()                                # something in front, like the scoring kernel in the real step
            reg.push(s)
            reg.wait()
        for it in range(6):
            dist.barrier()
            torch.cuda.synchronize()
            graph.replay()
            torch.cuda.synchronize()
            t = reg.local[128:128 + 64].view(torch.int64).cpu().tolist()
            rows.append([(x - t[0]) / 1e3 for x in t])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(g2, stream=st):
            for _ in range(20):
                reg.push(s)
                reg.wait()
        dist.barrier()
        torch.cuda.synchronize()
        g2.replay()
        e0.record(st)
        g2.replay()
        e1.record(st)
        torch.cuda.synchronize()
    out = [None] * world
    dist.all_gather_object(out, {"rank": rank, "us_per_exchange": e0.elapsed_time(e1) * 1e3 / 20, "timeline_us": rows[-3:]})
    if rank == 0:
        print(f"world {world}, {a.b} x {hl} x {a.p} fp32 per rank ({a.b * hl * a.p * 4} B), backend {reg.backend}")
        for o in out:
            print(f"rank {o['rank']}: {o['us_per_exchange']:.2f} us per back-to-back exchange")
            for r in o["timeline_us"]:
                print("   " + "  ".join(f"{i}:{x:6.2f}" for i, x in enumerate(r) if i in (0, 1, 2, 3, 6, 7)))
    assert reg.status() == 0
    del graph, g2
    reg.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
