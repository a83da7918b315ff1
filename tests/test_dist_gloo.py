"""world_size-2 gloo tests (CPU) of the KV-head sharding logic and the one exchange of the path.

The kernels are CUDA-only, so the per-rank scores here come from the oracle; what is tested is the
host logic of fastkv_b200.dist: head ranges, the all-gather / all-reduce exchange, and that the
importance vector every rank ends up with equals the single-process one (bit-identical in 'gather'
mode, within fp32 reassociation in 'reduce' mode)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fastkv_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, exchange, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from fastkv_b200.dist import HeadShard, head_range, combine_on_host
        torch.manual_seed(0)
        b, hq, hkv, l, d, w = 2, 8, 4, 300, 64, 8
        g = torch.Generator().manual_seed(5)
        q = torch.randn(b, hq, l, d, generator=g)
        k = torch.randn(b, hkv, l, d, generator=g)
        v = torch.randn(b, hkv, l, d, generator=g)
        knobs = dict(window_size=w, max_capacity_prompt=64, kernel_size=7, pooling="avgpool", tsp_layer=True, tsp_length=128)
        full = O.fastkv_update_kv(k, q, v, hq // hkv, deterministic_ties=True, **knobs)
        shard = HeadShard(exchange=exchange)
        assert shard.world_size == world and shard.rank == rank
        h0, h1 = shard.local_heads(hkv)
        assert (h0, h1) == head_range(hkv, rank, world)
        gq = hq // hkv
        # each rank scores only its own kv heads (and their query heads)
        local = O.fastkv_update_kv(k[:, h0:h1], q[:, h0 * gq:h1 * gq], v[:, h0:h1], gq, deterministic_ties=True, **knobs)
        assert torch.equal(local.score, full.score[:, h0:h1])             # heads are independent units
        assert torch.equal(local.key, full.key[:, h0:h1])
        part = local.score[:, 0].clone()
        for h in range(1, local.score.shape[1]):
            part += local.score[:, h]
        ex = shard.exchange_scores(local.score.contiguous(), partial=part)
        if exchange == "gather":
            assert ex.shape == full.score.shape and torch.equal(ex, full.score)
            imp = combine_on_host([full.score[:, r * (hkv // world):(r + 1) * (hkv // world)] for r in range(world)], 0, w, "gather")
            seq = full.score[:, 0].clone()
            for h in range(1, hkv):
                seq += full.score[:, h]
            assert torch.equal(imp, seq)
        else:
            assert ex.shape == (b, 1, l - w)
            assert torch.allclose(ex[:, 0], full.importance, rtol=1e-6, atol=0)
        # identical TSP selection on every rank
        imp_rank = ex.sum(1) if exchange == "reduce" else None
        if exchange == "gather":
            acc = ex[:, 0].clone()
            for h in range(1, ex.shape[1]):
                acc += ex[:, h]
            imp_rank = acc
        tsp = torch.cat([O.stable_topk_desc(imp_rank, 128 - w).sort(-1).values,
                         torch.arange(l - w, l).expand(b, -1)], -1)
        gathered = [torch.empty_like(tsp) for _ in range(world)]
        dist.all_gather(gathered, tsp)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        if exchange == "gather":
            assert torch.equal(tsp, full.tsp_idx)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["gather", "reduce"])
def test_head_sharded_exchange_world2(tmp_path, exchange):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, exchange, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_head_range_rejects_uneven_split():
    from fastkv_b200.dist import head_range
    assert head_range(8, 3, 4) == (6, 8)
    with pytest.raises(ValueError):
        head_range(8, 0, 3)


def test_exchange_modes_and_peer_region_sizes():
    """Host-side pieces of the peer-memory exchange that need no GPU: mode validation, region sizing, and that a
    single-rank shard never touches peer memory."""
    from fastkv_b200.dist import HeadShard
    from fastkv_b200.peer import PeerRegion
    from fastkv_b200 import _lib as L
    for mode in ("p2p", "gather", "reduce"):
        assert HeadShard(exchange=mode).world_size == 1
    with pytest.raises(ValueError):
        HeadShard(exchange="broadcast")
    assert PeerRegion.bytes_for(1, 8, 32760) == L.EXCHANGE_HEADER + 4 * 8 * 32760
    assert PeerRegion.bytes_for(8, 8, 32760) == L.EXCHANGE_HEADER + 8 * 4 * 8 * 32760
    shard = HeadShard(exchange="p2p")
    s = torch.rand(1, 2, 64)
    shard.begin_exchange(s)                              # world size 1: nothing to exchange, no region is created
    assert shard.region is None and shard.exchange_scores(s) is s
    with pytest.raises(RuntimeError):
        PeerRegion(1024)                                 # needs an initialised process group (one rank per GPU)
