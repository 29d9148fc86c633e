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
    """PUSH protocol for the sharded head's [2, A, P] exchange (`pfpn_head_logprob_push`, protocol 1 of `pfpn_head_push`):
    K1's finalize kernel stores this rank's dloc / dlogstd into row `rank` of EVERY rank's gather buffer as 8-byte packets
    {value, exchange number}; an aligned 8-byte store is single-copy atomic, so the receiver validates every element by
    its own sequence word -- no fence, ticket or flag round trip follows the data.  No exchange kernel sits behind K1.

    Producer and consumer are decoupled: ``push_args()`` numbers the exchanges 1, 2, 3, ...; the sum over ranks (rank
    order, bit-identical on every rank) of the OLDEST unconsumed exchange is produced either

    * by the NEXT head launch itself (``push_args(consume_into=out)``: the finalize thread that owns element i also sums
      element i of the previous exchange -- its packets arrived a step ago, so nothing waits and the exchange costs no
      kernel of its own; the sum feeds the optimizer, not the next head launch, so it may lag one exchange), or
    * by ``reduce(out)`` (`pfpn_peer_gather_sum_packets`): the last exchange of a run, or a caller that needs it at once.

    Measured on 8 B200 (bench.py exchange_breakdown): K1 alone 0.1585 ms/step; rows + flags + a consumer kernel per step
    0.1743; a consumer on a second stream was slower still (events in both directions every step).
    Buffers rotate over NBUF = 4 exchange slots: rank r's push of v+3 follows its own consumption of v+1, which needs rank
    p's push of v+1, which follows p's consumption of v-1 -- so slot (v-1) mod 4 is consumed everywhere before v+3
    overwrites it (holds while at most one exchange lags).

    Per rank: gather [NBUF][world][n] packets (8 bytes each), zero-initialised (sequence 0 is never valid)."""

    NBUF = 4
    MAX_LAG = 1

    def __init__(self, n: int, device: torch.device, group=None):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.n, self.dev = n, device
        self.pushed = 0    # exchanges produced (push_args handed out)
        self.consumed = 0  # exchanges whose sum was produced
        with torch.cuda.device(device):
            gather = _alloc(self.NBUF * self.world * n * 8)
            self._gather_ptr = gather[0]
            (self._gather_ptrs,) = _share([gather], device, group)
        # [NBUF][world][n][2] float32 view: [..., 0] = value, [..., 1] = sequence word (as float bits)
        self.gather = torch.as_tensor(_CudaArray(self._gather_ptr, self.NBUF * self.world * n * 2, "<f4"),
                                      device=device).view(self.NBUF, self.world, n, 2)
        self._push = [self._make_push(slot) for slot in range(self.NBUF)]

    @property
    def calls(self) -> int:
        return self.pushed

    def _rows_ptr(self, slot: int) -> int:
        return self._gather_ptr + slot * self.world * self.n * 8

    def _make_push(self, slot: int):
        p = _cabi.HeadPush()
        for r in range(self.world):
            p.out[r] = self._gather_ptrs[r] + ((slot * self.world + self.rank) * self.n) * 8
        p.nranks = self.world
        p.protocol = 1
        return p

    def slot_of(self, call: int) -> int:
        return call % self.NBUF

    def row(self, call: int, rank: int) -> torch.Tensor:
        """Values rank `rank` pushed in exchange `call` (this rank's local copy)."""
        return self.gather[self.slot_of(call), rank, :, 0]

    @property
    def pending(self) -> int:
        return self.pushed - self.consumed

    def push_args(self, consume_into: torch.Tensor = None, scale: float = 1.0):
        """The `pfpn_head_push` of the next exchange (head_call(push=...)).  With ``consume_into`` [n] the same launch also
        writes scale * (sum over ranks of the oldest unconsumed exchange) there, if one is pending."""
        if self.pending > self.MAX_LAG:
            raise RuntimeError("consume the pending exchange first (reduce()): at most one exchange may lag")
        self.pushed += 1
        p = self._push[self.slot_of(self.pushed)]
        p.value = self.pushed
        p.consume_value = 0
        if consume_into is not None and self.pending > 1:
            if consume_into.numel() != self.n or consume_into.dtype != torch.float32 or not consume_into.is_contiguous():
                raise ValueError("consume_into must be a contiguous float32 tensor of n elements")
            self.consumed += 1
            p.consume_value = self.consumed
            p.consume_rows = self._rows_ptr(self.slot_of(self.consumed))
            p.consume_out = consume_into.data_ptr()
            p.consume_scale = scale
        return p

    def reduce(self, out: torch.Tensor, scale: float = 1.0, stream_ptr: int = 0):
        """out[n] = scale * sum over ranks (rank order) of the OLDEST unconsumed exchange, on the producer's stream."""
        if self.consumed >= self.pushed:
            raise RuntimeError("nothing to consume: no exchange was pushed")
        self.consumed += 1
        _cabi.check(_cabi.pfpn_peer_gather_sum_packets(self._rows_ptr(self.slot_of(self.consumed)), self.world, self.consumed,
                                                       self.n, out.data_ptr(), scale, stream_ptr))
        return out
