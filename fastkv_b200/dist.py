"""KV-head sharding across the GPUs of one box and the single exchange the path needs.

The reference has no distributed code; this is the multi-GPU design of the
north star: every rank owns ``Hkv / world_size`` KV heads (and their query heads), scores, pools,
selects and compacts only those, and the TSP layer exchanges the per-head score vectors once so
that every rank runs the *same* global top-k (baselines/fastkv/utils.py:127 needs the sum over
ALL heads).

Three exchange modes:
  * ``"p2p"`` — the library's own exchange over peer memory (fkv_exchange_push / fkv_exchange_wait): every rank
    stores its rows straight into the other ranks' HBM over NVLink / NVSwitch, flags instead of a collective,
    capturable in a CUDA graph.  Same gathered layout and therefore the same bit-identical selection as "gather".
  * ``"gather"`` (default when no peer memory is available) — all-gather the fp32 per-KV-head scores ([B, Hkv/N, P] per rank, 4 MiB
    total at 128k) and sum them in fixed head order inside the top-k kernel: bit-identical
    selection at every GPU count, identical to the single-GPU result.
  * ``"reduce"`` — one fp32 sum all-reduce of the rank-local partial importance [B, P]
    (512 KiB at 128k): the north star's minimal exchange; summation order depends on N, so
    selections may differ from N=1 on near-ties (< 1e-5 relative).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib as L
from . import ops


def head_range(num_heads: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block of heads owned by ``rank`` (heads must divide evenly)."""
    if num_heads % world_size:
        raise ValueError(f"{num_heads} heads do not shard over {world_size} ranks")
    per = num_heads // world_size
    return rank * per, (rank + 1) * per


class HeadShard:
    """Process-group handle attached to ``FastKVCluster.shard`` on KV-head-sharded runs."""

    def __init__(self, group: Optional[dist.ProcessGroup] = None, exchange: str = "gather", region=None):
        if exchange not in ("gather", "reduce", "p2p"):
            raise ValueError("exchange must be 'p2p', 'gather' or 'reduce'")
        self.group = group
        self.exchange = exchange
        self.region = region                      # PeerRegion of the p2p exchange (allocated on first use when None)
        self.pending = None
        if region is not None:
            self.world_size, self.rank = region.world_size, region.rank
        else:
            self.world_size = dist.get_world_size(group) if dist.is_initialized() else 1
            self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def local_heads(self, num_heads: int) -> Tuple[int, int]:
        return head_range(num_heads, self.rank, self.world_size)

    # -- the exchange -----------------------------------------------------------------------------
    def begin_exchange(self, score_local: torch.Tensor) -> None:
        """p2p only: enqueue the stores of this rank's rows right behind the scoring kernel, so that they travel while
        the rank-local selection and compaction run; ``exchange_scores`` of the same tensor then only waits."""
        if self.exchange != "p2p" or self.world_size == 1:
            return
        from .peer import PeerRegion
        b, hl, p = score_local.shape
        need = PeerRegion.bytes_for(b, hl * self.world_size, p)
        if self.region is None or self.region.nbytes < need:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("fastkv_b200: the peer region must be allocated before CUDA-graph capture (run one eager step first)")
            if self.region is not None:
                self.region.close()
            self.region = PeerRegion(max(need, 4 << 20), self.group)     # collective: every rank reaches this with the same sizes
        self.region.push(score_local)
        self.pending = score_local

    def exchange_scores(self, score_local: torch.Tensor, partial: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[B, Hl, P] rank-local scores -> what the global top-k consumes.

        gather: [B, Hl * N, P] in global head order;  reduce: [B, 1, P] all-reduced partial sums
        (``partial`` = this rank's fixed-order head sum, computed on device when not given)."""
        n = self.world_size
        if n == 1:
            return score_local
        b, hl, p = score_local.shape
        if self.exchange == "p2p":
            if self.pending is not score_local:
                self.begin_exchange(score_local)
            self.pending = None
            return self.region.wait()
        if self.exchange == "gather":
            flat = torch.empty((n * b, hl, p), dtype=score_local.dtype, device=score_local.device)
            dist.all_gather_into_tensor(flat, score_local.contiguous(), group=self.group)   # rank-major concat on dim 0
            gathered = flat.view(n, b, hl, p)
            if b == 1:
                return gathered.view(1, n * hl, p)
            return gathered.permute(1, 0, 2, 3).reshape(b, n * hl, p)
        if partial is None:
            partial = local_importance(score_local)
        partial = partial.clone()
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=self.group)
        return partial.view(b, 1, p)

    def global_tsp(self, score_local: torch.Tensor, k: int, window: int):
        """Identical TSP index set on every rank: (tsp_idx [B, k + W] ascending, importance [B, P])."""
        full = self.exchange_scores(score_local)
        nsum = full.shape[1]
        if nsum == 1:
            imp = full[:, 0]
            return ops.topk(imp, k, order=L.ORDER_INDEX_ASC, tail=window), imp
        return ops.topk(full, k, order=L.ORDER_INDEX_ASC, tail=window, nsum=nsum, return_importance=True)


def local_importance(score_local: torch.Tensor) -> torch.Tensor:
    """Rank-local partial of the TSP importance: sum over the local heads in head order, on device.

    Uses the top-k kernel's fixed-order head-sum (k = 1 selection is discarded) so that no torch
    reduction with an unspecified order enters the path."""
    b, hl, p = score_local.shape
    if hl == 1:
        return score_local[:, 0].contiguous()
    _, imp = ops.topk(score_local, 1, order=L.ORDER_INDEX_ASC, nsum=hl, return_importance=True)
    return imp


def combine_on_host(per_rank_scores, k: int, window: int, exchange: str = "gather"):
    """Host-side model of the exchange for CPU tests of the sharding logic (gloo / fake ranks).

    ``per_rank_scores``: list over ranks of [B, Hl, P] CPU tensors in rank order.  Returns the
    importance vector [B, P] exactly as the device path forms it: fixed head order for 'gather',
    per-rank partials summed in rank order for 'reduce'."""
    if exchange == "gather":
        full = torch.cat(per_rank_scores, dim=1)
        imp = full[:, 0].clone()
        for h in range(1, full.shape[1]):
            imp += full[:, h]
        return imp
    parts = []
    for s in per_rank_scores:
        acc = s[:, 0].clone()
        for h in range(1, s.shape[1]):
            acc += s[:, h]
        parts.append(acc)
    imp = parts[0].clone()
    for part in parts[1:]:
        imp += part
    return imp
