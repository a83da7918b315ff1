"""The TSP layer's score exchange over peer memory (fkv_exchange_push / fkv_exchange_wait, csrc/exchange.cu).

One GPU is enough for the protocol: ``PeerRegion.local_group`` puts N "ranks" (regions + streams) on one device, and
the kernels do not care whether a peer pointer is remote.  The two-process version over real NVLink peers lives in
tests/multi_gpu_exchange.py (run under torchrun by test_two_process_exchange when the box has >= 2 GPUs)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _exchange_once(regions, streams, scores):
    views = [None] * len(regions)
    for r, (reg, st) in enumerate(zip(regions, streams)):
        with torch.cuda.stream(st):
            reg.push(scores[r], timeout_ms=5000)
    for r, (reg, st) in enumerate(zip(regions, streams)):
        with torch.cuda.stream(st):
            views[r] = reg.wait()
    return views


@pytest.mark.parametrize("world,b,hl,p", [(2, 1, 4, 32760), (4, 1, 2, 8184), (8, 1, 1, 4096), (2, 2, 4, 1021), (2, 8, 4, 4088)])
def test_every_rank_ends_up_with_all_rows_in_head_order(world, b, hl, p):
    from fastkv_b200.peer import PeerRegion
    dev = torch.device("cuda", 0)
    regions = PeerRegion.local_group(world, PeerRegion.bytes_for(b, hl * world, p), dev)
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    g = torch.Generator(device=dev).manual_seed(7)
    for it in range(4):                                 # several exchanges through the same regions (epochs 1..4)
        scores = [torch.rand(b, hl, p, generator=g, device=dev) + it for _ in range(world)]
        torch.cuda.synchronize()
        views = _exchange_once(regions, streams, scores)
        torch.cuda.synchronize()
        want = torch.cat(scores, dim=1)
        for r in range(world):
            assert torch.equal(views[r], want), (it, r)
            assert regions[r].status() == 0


def test_strided_source_rows_and_single_rank():
    from fastkv_b200.peer import PeerRegion
    dev = torch.device("cuda", 0)
    big = torch.rand(2, 6, 4096, device=dev)
    src = big[:, 1:5, :4088]                            # non-contiguous b / h strides
    (reg,) = PeerRegion.local_group(1, PeerRegion.bytes_for(2, 4, 4088), dev)
    reg.push(src)
    out = reg.wait()
    torch.cuda.synchronize()
    assert torch.equal(out, src)


def test_exchange_replays_inside_a_cuda_graph():
    """No host-side state: the epoch lives in the region, so a captured push / wait pair can be replayed."""
    from fastkv_b200.peer import PeerRegion
    from fastkv_b200 import ops, _lib as L
    dev = torch.device("cuda", 0)
    world, b, hl, p, k = 2, 1, 4, 32760, 2040
    regions = PeerRegion.local_group(world, PeerRegion.bytes_for(b, hl * world, p), dev)
    side = torch.cuda.Stream(dev)
    main = torch.cuda.Stream(dev)
    scores = [torch.rand(b, hl, p, device=dev) for _ in range(world)]
    outs = [None, None]

    def step():                                         # rank 0 on the capturing stream, rank 1 forked onto a side stream
        cur = torch.cuda.current_stream(dev)
        fork = torch.cuda.Event()
        fork.record(cur)
        side.wait_event(fork)
        regions[0].push(scores[0], timeout_ms=5000)
        with torch.cuda.stream(side):
            regions[1].push(scores[1], timeout_ms=5000)
            full1 = regions[1].wait()
            outs[1] = ops.topk(full1, k, order=L.ORDER_INDEX_ASC, tail=8, nsum=hl * world)
            join = torch.cuda.Event()
            join.record(side)
        full0 = regions[0].wait()
        outs[0] = ops.topk(full0, k, order=L.ORDER_INDEX_ASC, tail=8, nsum=hl * world)
        cur.wait_event(join)

    with torch.cuda.stream(main):
        step()                                          # eager warm-up (allocations)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=main):
            step()
        for it in range(3):
            for s in scores:
                s.copy_(torch.rand_like(s))
            torch.cuda.synchronize()
            graph.replay()
            torch.cuda.synchronize()
            ref = ops.topk(torch.cat(scores, dim=1), k, order=L.ORDER_INDEX_ASC, tail=8, nsum=hl * world)
            torch.cuda.synchronize()
            assert torch.equal(outs[0], ref) and torch.equal(outs[1], ref), it
    assert regions[0].status() == 0 and regions[1].status() == 0


def test_missing_peer_times_out_instead_of_hanging():
    from fastkv_b200.peer import PeerRegion
    from fastkv_b200 import _lib as L
    dev = torch.device("cuda", 0)
    regions = PeerRegion.local_group(2, PeerRegion.bytes_for(1, 2, 1024), dev)
    regions[0].push(torch.rand(1, 1, 1024, device=dev), timeout_ms=50)     # rank 1 never shows up
    regions[0].wait()
    torch.cuda.synchronize()
    assert regions[0].status() == L.ETIMEOUT


@pytest.mark.parametrize("world,n,dtype", [(2, 4096, torch.bfloat16), (4, 4096, torch.bfloat16), (8, 8 * 4096, torch.bfloat16),
                                           (2, 1000 * 8, torch.float16), (4, 40000, torch.float32)])
def test_peer_allreduce_sums_in_rank_order_on_every_rank(world, n, dtype):
    """fkv_allreduce_sum with N "ranks" on one device: fp32 adds in rank order, identical bits on all ranks, graph replay."""
    from fastkv_b200.peer import PeerRegion
    dev = torch.device("cuda", 0)
    regions = PeerRegion.local_group(world, PeerRegion.allreduce_bytes_for(world, n * 4), dev)
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    g = torch.Generator(device=dev).manual_seed(9)
    bufs = [torch.empty(n, dtype=dtype, device=dev) for _ in range(world)]

    def fill():
        parts = [torch.randn(n, generator=g, device=dev).to(dtype) for _ in range(world)]
        for b_, p_ in zip(bufs, parts):
            b_.copy_(p_)
        acc = torch.zeros(n, dtype=torch.float32, device=dev)
        for p_ in parts:                                 # rank order, fp32
            acc += p_.float()
        return acc.to(dtype)

    def launch_all():
        for r, st in enumerate(streams):
            with torch.cuda.stream(st):
                regions[r].all_reduce_(bufs[r], timeout_ms=5000)

    for it in range(3):                                  # both slots, several epochs
        want = fill()
        torch.cuda.synchronize()
        launch_all()
        torch.cuda.synchronize()
        for r in range(world):
            assert torch.equal(bufs[r], want), (it, r)
    # one graph holding all ranks' kernels (fork / join), replayed with fresh inputs: no host-side state
    main = torch.cuda.Stream(dev)
    with torch.cuda.stream(main):
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=main):
            fork = torch.cuda.Event()
            fork.record(main)
            for r, st in enumerate(streams):
                st.wait_event(fork)
                with torch.cuda.stream(st):
                    regions[r].all_reduce_(bufs[r], timeout_ms=5000)
                    ev = torch.cuda.Event()
                    ev.record(st)
                main.wait_event(ev)
        for it in range(3):
            want = fill()
            torch.cuda.synchronize()
            graph.replay()
            torch.cuda.synchronize()
            for r in range(world):
                assert torch.equal(bufs[r], want), ("graph", it, r)
    assert all(reg.status() == 0 for reg in regions)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_process_exchange():
    env = dict(os.environ, PYTHONPATH=ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "multi_gpu_exchange.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=420)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-4000:]
    assert "EXCHANGE OK" in out.stdout
