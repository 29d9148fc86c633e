"""Host-buffer entry point of the head: the call a user makes when the minibatch
lives in (pinned) host memory, as the reference's ``net.train(sess, ...)`` feed
does (/root/reference/networks/actor_critic/ppo.py:64-72).

The batch is cut into chunks that are pushed through ``n_streams`` CUDA streams so
that the host->device copy of chunk i+1, the kernel of chunk i and the device->host
copy of chunk i-1 overlap (PCIe is full duplex).  Everything that touches the data
is either a cudaMemcpyAsync or one of our kernels.
"""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from . import _cabi
from . import head as _head


def bind_host_to_gpu(device_index: int) -> Optional[str]:
    """Pin this process to the CPUs next to GPU `device_index` (NVML's ideal CPU set, cut to what the process may use),
    so that the pinned host buffers allocated afterwards land on the GPU's own NUMA node.  With one process per GPU on a
    multi-socket host this decides whether the host<->device copies cross the socket interconnect.  Returns the CPU list
    as a string, or None when NVML / the affinity call is unavailable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(device_index)
        try:
            bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:  # noqa: BLE001 -- older torch: fall back to the NVML index
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        ncpu = os.cpu_count() or 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus = sorted(ideal & os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"{cpus[0]}-{cpus[-1]} ({len(cpus)} cpus)"
    except Exception:  # noqa: BLE001 -- binding is an optimisation, never a requirement
        return None


class HostHeadPipeline:
    """PPO-fused head fwd+bwd over host-resident minibatches of a fixed shape."""

    def __init__(self, B: int, A: int, P: int, device: torch.device, chunk: int = 8192, n_streams: int = 3,
                 eps_clip: float = 0.2):
        self.B, self.A, self.P, self.dev = B, A, P, device
        self.chunk = min(chunk, B)
        self.bounds = self._schedule(B, self.chunk)
        self.nchunks = len(self.bounds)
        self.eps_clip = eps_clip
        self.streams = [torch.cuda.Stream(device) for _ in range(n_streams)]
        f = lambda *s: torch.empty(*s, dtype=torch.float32, device=device)
        self.slots = []
        for _ in range(n_streams):
            self.slots.append(dict(logits=f(self.chunk, A, P), value=f(self.chunk, A), loss=f(1),
                                   ws=torch.empty(_head.head_workspace_bytes(A, P), dtype=torch.uint8, device=device),
                                   done=torch.cuda.Event()))
        # per-state vectors live whole on the device: one copy each way per step instead of one per chunk (a small copy
        # costs the copy engine ~5 us whatever its size)
        self.adv_d, self.lp_old_d, self.lp_d, self.ent_d = f(B), f(B), f(B), f(B)
        self.stats = f(2)
        self.loc_d, self.logstd_d = f(A, P), f(A, P)
        self.partials = f(self.nchunks, 3, A * P)  # per chunk: dloc, dlogstd, (loss in [2,0])
        self.stats_ready = torch.cuda.Event()
        # pinned result buffers
        p = lambda *s: torch.empty(*s, dtype=torch.float32).pin_memory()
        self.out = dict(lp=p(B), ent=p(B), dlogits=p(B, A, P), dloc=p(A, P), dlogstd=p(A, P), loss=p(1))
        self.h2d_bytes = 4 * (B * A * P + B * A + 2 * B + 2 * A * P)
        self.d2h_bytes = 4 * (B * A * P + 2 * B + 2 * A * P + 1)
        self.launches_per_step = 1 + 2 * self.nchunks + 1

    @staticmethod
    def _schedule(B: int, chunk: int):
        """Chunk boundaries.  The first chunk's host->device copy and the last chunk's device->host copy overlap with
        nothing (pipeline fill / drain: 1/8 of the step at 8 equal chunks, measured 41.6 of the 46 GB/s each way the box
        sustains in both directions at once), so the schedule ramps up from chunk/8 and back down to it."""
        ramp = [max(256, chunk >> sft) for sft in (3, 2, 1)]
        body = B - 2 * sum(ramp)
        if body < chunk:
            sizes = [chunk] * (B // chunk) + ([B % chunk] if B % chunk else [])
        else:
            sizes = ramp + [chunk] * (body // chunk) + ([body % chunk] if body % chunk else []) + ramp[::-1]
        out, lo = [], 0
        for n in sizes:
            out.append((lo, lo + n))
            lo += n
        assert lo == B
        return out

    def run(self, logits, loc, logstd, value, adv, lp_old) -> Dict[str, torch.Tensor]:
        """All arguments are pinned fp32 host tensors; returns pinned host tensors
        (valid after the call: it synchronises the device at the end)."""
        B, A, P, ch = self.B, self.A, self.P, self.chunk
        main = torch.cuda.current_stream(self.dev)
        with torch.cuda.stream(main):
            self.adv_d.copy_(adv, non_blocking=True)
            self.lp_old_d.copy_(lp_old, non_blocking=True)
            self.loc_d.copy_(loc, non_blocking=True)
            self.logstd_d.copy_(logstd, non_blocking=True)
            _head.adv_stats(self.adv_d, out=self.stats)
            self.stats_ready.record(main)
        for ci in range(self.nchunks):
            s = self.streams[ci % len(self.streams)]
            sl = self.slots[ci % len(self.streams)]
            lo, hi = self.bounds[ci]
            n = hi - lo
            with torch.cuda.stream(s):
                sl["logits"][:n].copy_(logits[lo:hi], non_blocking=True)
                sl["value"][:n].copy_(value[lo:hi], non_blocking=True)
                s.wait_event(self.stats_ready)  # (only the kernel needs the whole-batch statistics, not the copies)
                a = _cabi.HeadArgs()
                a.logits = a.dlogits = sl["logits"].data_ptr()  # gradient written in place
                a.loc, a.logstd, a.value = self.loc_d.data_ptr(), self.logstd_d.data_ptr(), sl["value"].data_ptr()
                a.adv, a.lp_old, a.adv_stats = self.adv_d[lo:hi].data_ptr(), self.lp_old_d[lo:hi].data_ptr(), self.stats.data_ptr()
                a.eps_clip, a.loss_scale = self.eps_clip, 1.0 / B
                a.lp, a.ent = self.lp_d[lo:hi].data_ptr(), self.ent_d[lo:hi].data_ptr()
                part = self.partials[ci]
                a.dloc, a.dlogstd, a.loss = part[0].data_ptr(), part[1].data_ptr(), part[2].data_ptr()
                a.B, a.A, a.P, a.mode, a.flags = n, A, P, _cabi.HEAD_PPO, 0
                _cabi.check(_cabi.pfpn_head_logprob(a, sl["ws"].data_ptr(), sl["ws"].numel(), s.cuda_stream))
                self.out["dlogits"][lo:hi].copy_(sl["logits"][:n], non_blocking=True)
                sl["done"].record(s)
        with torch.cuda.stream(main):
            for sl in self.slots:
                main.wait_event(sl["done"])
            self.out["lp"].copy_(self.lp_d, non_blocking=True)
            self.out["ent"].copy_(self.ent_d, non_blocking=True)
            tot = self.partials.sum(0)  # [3, AP] -- tiny plumbing reduction over chunks
            self.out["dloc"].copy_(tot[0].view(A, P), non_blocking=True)
            self.out["dlogstd"].copy_(tot[1].view(A, P), non_blocking=True)
            self.out["loss"].copy_(tot[2, :1], non_blocking=True)
        main.synchronize()
        return self.out
