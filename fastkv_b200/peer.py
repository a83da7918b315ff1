"""Peer-mapped device memory for the TSP layer's score exchange (include/fastkv_b200.h: fkv_exchange_*).

One ``PeerRegion`` per rank: a device buffer that every rank of the box can store into over NVLink / NVSwitch.
Two ways to get one (set-up plumbing only; the exchange itself is the library's own kernels):

* ``"symm"``  a torch symmetric-memory allocation (``torch.distributed._symmetric_memory``), rendezvoused over the
              process group;
* ``"ipc"``   ``fkv_peer_alloc`` / ``fkv_peer_open`` (cudaMalloc + CUDA IPC handles, exchanged with
              ``all_gather_object``).

``backend=None`` tries them in that order; the choice is made collectively (every rank must succeed).
``FKV_PEER_BACKEND`` overrides.  A single process may also build several "ranks" on one device
(``PeerRegion.local_group``) — that is how the protocol is tested on one GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib as L


class _RawCudaBuffer:
    """__cuda_array_interface__ shim: lets torch wrap a raw device pointer without copying."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


class PeerRegion:
    """``ptrs[r]`` = rank r's region as mapped in this process; ``local`` = uint8 tensor over this rank's own region."""

    def __init__(self, nbytes: int, group: Optional[dist.ProcessGroup] = None, backend: Optional[str] = None):
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerRegion needs an initialised process group (one rank per GPU)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        if self.world_size > L.MAX_PEERS:
            raise ValueError(f"at most {L.MAX_PEERS} ranks per box")
        self.nbytes = (int(nbytes) + 255) // 256 * 256
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._ipc_opened: List[int] = []
        self._ipc_owned = None
        self._symm = None
        backend = backend or os.environ.get("FKV_PEER_BACKEND") or None
        order = [backend] if backend else ["symm", "ipc"]
        errors = []
        self.backend = None
        for b in order:
            err = None
            try:
                getattr(self, "_setup_" + b)()
            except Exception as e:  # noqa: BLE001 — any failure here just means "this backend is unavailable"
                err = f"{type(e).__name__}: {e}"
            ok = torch.tensor([0 if err else 1], device=self.device, dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if ok.item() == 1:
                self.backend = b
                break
            errors.append(f"{b}: {err or 'failed on another rank'}")
            self._teardown()
        if self.backend is None:
            raise RuntimeError("no peer-memory backend available: " + "; ".join(errors))
        torch.cuda.synchronize(self.device)
        dist.barrier(group=group)                     # every region is zeroed and mapped before anyone pushes

    # -- backends ---------------------------------------------------------------------------------
    def _setup_symm(self):
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        hdl = symm.rendezvous(t, self.group if self.group is not None else dist.group.WORLD)
        t.zero_()
        self.local = t
        self.ptrs = [int(p) for p in hdl.buffer_ptrs]
        self._symm = hdl

    def _setup_ipc(self):
        lib = L.load()
        ptr = C.c_void_p()
        handle = C.create_string_buffer(64)
        L.check(lib.fkv_peer_alloc(self.nbytes, C.byref(ptr), handle))
        self._ipc_owned = ptr.value
        handles = [None] * self.world_size
        dist.all_gather_object(handles, bytes(handle.raw), group=self.group)
        self.ptrs = []
        for r, h in enumerate(handles):
            if r == self.rank:
                self.ptrs.append(ptr.value)
                continue
            mapped = C.c_void_p()
            L.check(lib.fkv_peer_open(C.create_string_buffer(h, 64), C.byref(mapped)))
            self._ipc_opened.append(mapped.value)
            self.ptrs.append(mapped.value)
        self.local = torch.as_tensor(_RawCudaBuffer(ptr.value, self.nbytes), device=self.device)

    def _teardown(self):
        lib = L.load()
        for m in self._ipc_opened:
            lib.fkv_peer_close(m)
        self._ipc_opened = []
        if self._ipc_owned is not None:
            lib.fkv_peer_free(self._ipc_owned)
            self._ipc_owned = None
        self._symm = None
        self.local = None
        self.ptrs = []

    def close(self):
        """Collective: nobody may unmap while a peer can still store into the region."""
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)
        self._teardown()

    # -- single-process groups (tests, one GPU) -----------------------------------------------------
    @classmethod
    def local_group(cls, world_size: int, nbytes: int, device) -> List["PeerRegion"]:
        """``world_size`` regions on ONE device, each seeing all of them: the protocol without a second GPU."""
        nbytes = (int(nbytes) + 255) // 256 * 256
        bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=device) for _ in range(world_size)]
        out = []
        for r in range(world_size):
            self = cls.__new__(cls)
            self.group, self.rank, self.world_size, self.nbytes = None, r, world_size, nbytes
            self.device = torch.device(device)
            self.local, self.ptrs, self.backend = bufs[r], [b.data_ptr() for b in bufs], "local"
            self._ipc_opened, self._ipc_owned, self._symm = [], None, None
            self._keep = bufs
            out.append(self)
        return out

    # -- the exchange ---------------------------------------------------------------------------------
    def _params(self, score_local: torch.Tensor, max_ctas: int, timeout_ms: int) -> "L.ExchangeParams":
        b, hl, p = score_local.shape
        if score_local.dtype != torch.float32 or score_local.stride(2) != 1:
            raise ValueError("fastkv_b200: the exchanged scores must be fp32 with a contiguous last dim")
        x = L.ExchangeParams()
        x.struct_size = C.sizeof(L.ExchangeParams)
        x.rank, x.world, x.B, x.heads_local, x.P = self.rank, self.world_size, b, hl, p
        x.score, x.score_stride_b, x.score_stride_h = score_local.data_ptr(), score_local.stride(0), score_local.stride(1)
        for r, ptr in enumerate(self.ptrs):
            x.region[r] = ptr
        x.region_bytes = self.nbytes
        x.max_ctas, x.timeout_ms = max_ctas, timeout_ms
        return x

    def push(self, score_local: torch.Tensor, *, max_ctas: int = 0, timeout_ms: int = 0):
        """Enqueue the handshake + stores of this rank's [B, Hl, P] rows on the current stream."""
        L.require_cuda(score_local)
        self._shape = tuple(score_local.shape)
        self._x = self._params(score_local, max_ctas, timeout_ms)
        L.check(L.load().fkv_exchange_push(C.byref(self._x), L.stream_ptr(self.device)))

    def wait(self) -> torch.Tensor:
        """Enqueue the wait for every peer's rows; returns the gathered [B, world * Hl, P] view of the local region
        (valid for kernels enqueued on the current stream after this call, until the next ``push``)."""
        L.check(L.load().fkv_exchange_wait(C.byref(self._x), L.stream_ptr(self.device)))
        b, hl, p = self._shape
        n = b * hl * self.world_size * p
        return self.local[L.EXCHANGE_HEADER:L.EXCHANGE_HEADER + 4 * n].view(torch.float32).view(b, hl * self.world_size, p)

    def status(self) -> int:
        """fkv_exchange_status (synchronises the current stream): 0 or ETIMEOUT."""
        return L.load().fkv_exchange_status(self.ptrs[self.rank], L.stream_ptr(self.device))

    @staticmethod
    def bytes_for(b: int, heads_total: int, p: int) -> int:
        return L.EXCHANGE_HEADER + 4 * b * heads_total * p

    # -- small all-reduce (a region used for this is used for nothing else) ---------------------------
    @staticmethod
    def allreduce_bytes_for(world_size: int, max_bytes: int) -> int:
        """Region size whose slots take tensors of up to ``max_bytes`` (a tensor may use half a slot)."""
        return L.ALLREDUCE_HEADER + 2 * world_size * 2 * ((int(max_bytes) + 255) // 256 * 256)

    def allreduce_capacity(self) -> int:
        """Largest tensor (bytes) one ``all_reduce_`` call can sum through this region."""
        return (((self.nbytes - L.ALLREDUCE_HEADER) // (2 * self.world_size)) // 256 * 256) // 2

    def all_reduce_(self, t: torch.Tensor, *, timeout_ms: int = 0) -> torch.Tensor:
        """In-place sum of ``t`` over the ranks (fkv_allreduce_sum: one kernel on the current stream, fp32 adds in rank
        order — the same bits on every rank; CUDA-graph capturable)."""
        L.require_cuda(t)
        if not t.is_contiguous():
            raise ValueError("fastkv_b200: all_reduce_ needs a contiguous tensor")
        a = L.AllreduceParams()
        a.struct_size = C.sizeof(L.AllreduceParams)
        a.rank, a.world, a.dtype, a.n = self.rank, self.world_size, L.dtype_code(t.dtype), t.numel()
        a.data = t.data_ptr()
        for r, ptr in enumerate(self.ptrs):
            a.region[r] = ptr
        a.region_bytes, a.timeout_ms = self.nbytes, timeout_ms
        L.check(L.load().fkv_allreduce_sum(C.byref(a), L.stream_ptr(self.device)))
        return t
