"""Peer-memory plumbing for the kernels that exchange over NVLink (csrc/comm.cu, K1's finalize).

Every rank owns staging / gather buffers and an int32 flag array, allocated by the library (cudaMalloc) and exported
through CUDA IPC; the 64-byte handles travel over torch.distributed and every rank maps its peers' buffers into its own
device context (cudaIpcOpenMemHandle with lazy peer access).  torch only wraps the local buffer as a tensor.
The mapping outcome is COLLECTIVE: either every rank maps every peer, or every rank raises (callers fall back to NCCL).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch
import torch.distributed as dist

from . import _cabi


class _CudaArray:
    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def _alloc(nbytes: int):
    ptr, handle = C.c_void_p(), C.create_string_buffer(64)
    _cabi.check(_cabi.pfpn_peer_alloc(nbytes, C.byref(ptr), handle))
    return ptr.value, handle.raw


def _share(local: Sequence, device: torch.device, group) -> List[List[int]]:
    """local = [(ptr, handle), ...] of this rank; returns ptrs[k][r] = rank r's k-th buffer mapped into this process."""
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world > 8:
        raise ValueError("the peer-memory kernels cover one NVSwitch node (<= 8 ranks)")
    gathered: List = [None] * world
    dist.all_gather_object(gathered, tuple(h for _, h in local), group=group)
    ptrs = [[0] * world for _ in local]
    err = None
    for r, handles in enumerate(gathered):
        for k, h in enumerate(handles):
            if r == rank:
                ptrs[k][r] = local[k][0]
                continue
            p = C.c_void_p()
            try:  # (a peer on another host / without P2P: cudaIpcOpenMemHandle fails on THIS rank only)
                _cabi.check(_cabi.pfpn_peer_open(h, C.byref(p)))
            except Exception as e:  # noqa: BLE001 -- reported collectively below
                err = e
            ptrs[k][r] = p.value or 0
    ok = torch.tensor([0 if err is not None else 1], device=device, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
    if int(ok.item()) == 0:
        raise RuntimeError(f"peer-memory mapping failed on at least one rank ({err!r}): CUDA IPC + P2P need all ranks on "
                           "one NVLink/NVSwitch node")
    dist.barrier(group=group)
    return ptrs


class PeerBuckets:
    """PULL protocol of the fused all-reduce + Adam (pfpn_peer_allreduce_adam[_rs]): every rank publishes its bucket in
    its own double-buffered staging area; the kernels read the peers' staging areas."""

    def __init__(self, n_total: int, device: torch.device, group=None, with_reduced: bool = False):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n_total, self.dev, self.group = n_total, device, group
        with torch.cuda.device(device):
            stage = _alloc(2 * n_total * 4)
            flag = _alloc(64 * 4)
            # averaged-slice buffer of the two-phase all-reduce (pfpn_peer_allreduce_adam_rs); tiny dummy otherwise
            red = _alloc((n_total if with_reduced else 4) * 4)
            self._stage_ptr, self._flag_ptr, self._red_ptr = stage[0], flag[0], red[0]
            self.stage = torch.as_tensor(_CudaArray(self._stage_ptr, 2 * n_total, "<f4"), device=device).view(2, n_total)
            stage_ptrs, flag_ptrs, red_ptrs = _share([stage, flag, red], device, group)
        self._flag_ptrs = (C.c_void_p * self.world)(*flag_ptrs)
        self.calls = 0  # exchange calls issued on these buffers (flag value / buffer parity)
        self.reduced_ptrs = (C.c_void_p * self.world)(*red_ptrs)
        self._bucket_ptrs = [(C.c_void_p * self.world)(*[p + par * n_total * 4 for p in stage_ptrs]) for par in (0, 1)]

    def ptrs(self, parity: int):
        return self._bucket_ptrs[parity], self._flag_ptrs


class PeerSum:
    """The sharded head's [2, A, P] exchange, pull form (`pfpn_peer_allreduce_sum`): the caller writes dloc / dlogstd into
    ``slot()`` and then calls ``reduce``.  Kept as the comparison point of ``PeerGather``."""

    def __init__(self, n: int, device: torch.device, group=None):
        if n % 4:
            raise ValueError("n must be a multiple of 4 floats")
        self.pb = PeerBuckets(n, device, group)
        self.n, self.calls = n, 0

    def slot(self) -> torch.Tensor:
        """Staging buffer [n] the NEXT ``reduce`` call will publish."""
        return self.pb.stage[(self.calls + 1) & 1]

    def reduce(self, out: torch.Tensor, scale: float = 1.0, stream_ptr: int = 0):
        self.calls += 1
        buckets, flags = self.pb.ptrs(self.calls & 1)
        _cabi.check(_cabi.pfpn_peer_allreduce_sum(buckets, flags, self.pb.rank, self.pb.world, self.calls, self.n,
                                                  out.data_ptr(), scale, stream_ptr))
        return out


class PeerGather:
    """PUSH protocol for the sharded head's [2, A, P] exchange (`pfpn_head_logprob_push` + `pfpn_peer_gather_sum`): K1's
    finalize kernel stores this rank's dloc / dlogstd into row `rank` of EVERY rank's gather buffer and raises the flags;
    the consumer waits for the N flags and sums the N local rows in rank order.  No exchange kernel sits behind K1.

    Producer and consumer are decoupled: ``push_args()`` numbers the exchanges 1, 2, 3, ... and ``reduce()`` consumes the
    OLDEST one not yet consumed.  A caller that consumes exchange v only after it has produced v+1 (``lag = 1``: the sum is
    needed by the optimizer, not by the next head launch) never waits for the slowest rank inside a step -- the flags of
    v arrived a whole step ago -- and everything stays on ONE stream.  Measured: a consumer on a second stream (events in
    both directions every step) was SLOWER than the in-step consumer (0.1733 vs 0.1692 ms per step on 2 GPUs).
    Buffers rotate over NBUF = 4 exchange slots: rank r's push of v+3 follows its own reduce of v+1, which needs rank p's
    push of v+1, which follows p's reduce of v-1 -- so slot (v-1) mod 4 is consumed everywhere before v+3 overwrites it
    (holds for lag <= 1).

    Per rank: gather [NBUF][world][n] floats, flags int32[64] (word r = last exchange rank r pushed), one local ticket."""

    NBUF = 4
    MAX_LAG = 1

    def __init__(self, n: int, device: torch.device, group=None):
        if n % 4:
            raise ValueError("n must be a multiple of 4 floats")
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n, self.dev = n, device
        self.pushed = 0    # exchanges produced (push_args handed out)
        self.consumed = 0  # exchanges consumed by reduce()
        with torch.cuda.device(device):
            gather = _alloc(self.NBUF * self.world * n * 4)
            flag = _alloc(64 * 4)
            self._gather_ptr, self._flag_ptr = gather[0], flag[0]
            self._gather_ptrs, self._flag_ptrs = _share([gather, flag], device, group)
        self.gather = torch.as_tensor(_CudaArray(self._gather_ptr, self.NBUF * self.world * n, "<f4"),
                                      device=device).view(self.NBUF, self.world, n)
        self.ticket = torch.zeros(1, dtype=torch.int32, device=device)
        self._push = [self._make_push(slot) for slot in range(self.NBUF)]

    @property
    def calls(self) -> int:
        return self.pushed

    def _make_push(self, slot: int):
        p = _cabi.HeadPush()
        for r in range(self.world):
            p.out[r] = self._gather_ptrs[r] + ((slot * self.world + self.rank) * self.n) * 4
            p.flags[r] = self._flag_ptrs[r] + 4 * self.rank
        p.ticket = self.ticket.data_ptr()
        p.nranks = self.world
        return p

    def slot_of(self, call: int) -> int:
        return call % self.NBUF

    def push_args(self):
        """The `pfpn_head_push` of the next exchange: pass it to pfpn_head_logprob_push (head_call(push=self))."""
        if self.pushed - self.consumed > self.MAX_LAG:
            raise RuntimeError("consume the pending exchanges first (reduce()): at most one exchange may lag")
        self.pushed += 1
        p = self._push[self.slot_of(self.pushed)]
        p.value = self.pushed
        return p

    @property
    def pending(self) -> int:
        return self.pushed - self.consumed

    def reduce(self, out: torch.Tensor, scale: float = 1.0, stream_ptr: int = 0):
        """out[n] = scale * sum over ranks (rank order) of the OLDEST unconsumed exchange, on the producer's stream."""
        if self.consumed >= self.pushed:
            raise RuntimeError("nothing to consume: no exchange was pushed")
        self.consumed += 1
        slot = self.slot_of(self.consumed)
        _cabi.check(_cabi.pfpn_peer_gather_sum(self._gather_ptr + slot * self.world * self.n * 4, self._flag_ptr, self.world,
                                               self.consumed, self.n, out.data_ptr(), scale, stream_ptr))
        return out
