"""Peer-memory plumbing for the fused all-reduce + Adam kernel (csrc/comm.cu).

Every rank owns a double-buffered staging area for its [gradient | statistics] bucket and an int32
flag array, allocated by the library (cudaMalloc) and exported through CUDA IPC; the 64-byte handles
travel over torch.distributed and every rank maps its peers' buffers into its own device context
(cudaIpcOpenMemHandle with lazy peer access).  torch only wraps the local buffer as a tensor.
"""
from __future__ import annotations

import ctypes as C
from typing import List

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


class PeerBuckets:
    def __init__(self, n_total: int, device: torch.device, group=None, with_reduced: bool = False):
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 8:
            raise ValueError("peer all-reduce covers one NVSwitch node (<= 8 ranks)")
        self.n_total, self.dev, self.group = n_total, device, group
        with torch.cuda.device(device):
            self._stage_ptr, h_stage = _alloc(2 * n_total * 4)
            self._flag_ptr, h_flag = _alloc(64 * 4)
            # averaged-slice buffer of the two-phase all-reduce (pfpn_peer_allreduce_adam_rs); tiny dummy otherwise
            self._red_ptr, h_red = _alloc((n_total if with_reduced else 4) * 4)
            self.stage = torch.as_tensor(_CudaArray(self._stage_ptr, 2 * n_total, "<f4"), device=device).view(2, n_total)
            gathered: List = [None] * self.world
            dist.all_gather_object(gathered, (h_stage, h_flag, h_red), group=group)
            stage_ptrs, flag_ptrs, red_ptrs = [], [], []
            err = None
            for r, (hs, hf, hr) in enumerate(gathered):
                if r == self.rank:
                    stage_ptrs.append(self._stage_ptr)
                    flag_ptrs.append(self._flag_ptr)
                    red_ptrs.append(self._red_ptr)
                    continue
                ps, pf, pr = C.c_void_p(), C.c_void_p(), C.c_void_p()
                try:  # (a peer on another host / without P2P: cudaIpcOpenMemHandle fails on THIS rank only)
                    _cabi.check(_cabi.pfpn_peer_open(hs, C.byref(ps)))
                    _cabi.check(_cabi.pfpn_peer_open(hf, C.byref(pf)))
                    _cabi.check(_cabi.pfpn_peer_open(hr, C.byref(pr)))
                except Exception as e:  # noqa: BLE001 -- reported collectively below
                    err = e
                stage_ptrs.append(ps.value or 0)
                flag_ptrs.append(pf.value or 0)
                red_ptrs.append(pr.value or 0)
            # the outcome must be COLLECTIVE: either every rank maps every peer or every rank raises (and the caller
            # falls back to the NCCL path on all ranks together)
            ok = torch.tensor([0 if err is not None else 1], device=device, dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                raise RuntimeError(f"peer-memory mapping failed on at least one rank ({err!r}): CUDA IPC + P2P need all "
                                   "ranks on one NVLink/NVSwitch node")
        dist.barrier(group=group)
        self._flag_ptrs = (C.c_void_p * self.world)(*flag_ptrs)
        self.calls = 0  # exchange calls issued on these buffers (flag value / buffer parity / CTA-counter epoch)
        self.reduced_ptrs = (C.c_void_p * self.world)(*red_ptrs)
        self._bucket_ptrs = [(C.c_void_p * self.world)(*[p + par * n_total * 4 for p in stage_ptrs]) for par in (0, 1)]

    def ptrs(self, parity: int):
        return self._bucket_ptrs[parity], self._flag_ptrs


class PeerSum:
    """The sharded head's [2, A, P] exchange over peer memory (`pfpn_peer_allreduce_sum`): the caller lets
    K1's finalize kernel write dloc / dlogstd straight into ``slot(parity)`` and then calls ``reduce``."""

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
