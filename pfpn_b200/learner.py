"""Data-parallel learner update: the B200 replacement of the reference's synchronous
parameter-server step (/root/reference/models/sync_model.py:60-101 +
models/workers/base_worker.py:25-120).

Reference order of operations ([graph]): per worker tf.gradients -> LOCAL
clip_by_global_norm -> accumulator MEAN over workers of the clipped gradients and of the pushed
statistics (state mean/std, max/sum_active: averaged, not max-reduced) -> assign statistics ->
Adam -> train_flag tick / resample.  Here the weights are replicated on every GPU, one process per
GPU, and the accumulator is ONE NCCL all-reduce over NVLink of the flat bucket
[gradients | statistics]; Adam then runs identically on every rank, so no broadcast is needed.

The collective plumbing below is device-agnostic torch.distributed code (covered by world-size-2
gloo tests on CPU); the arithmetic is CUDA kernels behind the C ABI.
"""
from __future__ import annotations

from typing import Optional, Tuple

import os

import torch
import torch.distributed as dist

from . import _cabi
from .head import _stream_ptr


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n minibatch rows (B/N rows per rank)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mean_(bucket: torch.Tensor, group=None) -> float:
    """Sum all-reduce of the flat bucket in place; returns the 1/N factor the caller still has to
    apply (the Adam kernel folds it into its gradient read)."""
    _, n = world()
    if n > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / n


def assert_replicas_identical(params: torch.Tensor, group=None, what: str = "parameters"):
    """There is no parameter broadcast on this path (every rank applies the identical Adam step to identical weights), so
    the replicas must START identical: raise if the ranks were built with different seeds / checkpoints.  Compares a 64-bit
    checksum of the raw fp32 bits; a no-op for a single process."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) < 2:
        return
    bits = params.detach().contiguous().view(torch.int32).to(torch.int64)
    w = torch.arange(1, bits.numel() + 1, device=bits.device, dtype=torch.int64)
    h = torch.stack([bits.sum(), (bits * (w % 65521)).sum()])
    hs = [torch.empty_like(h) for _ in range(dist.get_world_size(group))]
    dist.all_gather(hs, h, group=group)
    if any(not torch.equal(x, hs[0]) for x in hs):
        raise RuntimeError(f"data-parallel replicas hold different {what}: construct every rank's network with the same "
                           "`seed` (the per-rank sampling stream is derived from the rank automatically) or load the same checkpoint")


class SyncReplicasAdam:
    """`SyncReplicasOptimizer(AdamOptimizer(lr), replicas_to_aggregate=N)` + clip_grads, for one
    network whose parameters / gradients are flat buffers (pfpn_b200.network)."""

    def __init__(self, lr: float = 1e-4, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
                 norm_clip: Optional[float] = 1.0, group=None, fused_peer: Optional[bool] = None):
        self.lr, self.beta1, self.beta2, self.eps = lr, beta1, beta2, eps
        self.norm_clip = float(norm_clip) if norm_clip else 0.0
        self.group = group
        # N > 1: sum the buckets straight out of NVLink peer memory inside the Adam kernel (csrc/comm.cu)
        # instead of NCCL all-reduce + Adam; PFPN_FUSED_ALLREDUCE=0 selects the NCCL path.
        import os
        self.fused_peer = (os.environ.get("PFPN_FUSED_ALLREDUCE", "1") != "0") if fused_peer is None else fused_peer
        self._peers = None
        self.step = 0
        self.m = self.v = None
        self.norm_scale = None
        self._scratch = None

    def _lazy(self, net):
        if self.m is None:
            self.m = torch.zeros_like(net.params)
            self.v = torch.zeros_like(net.params)
            assert_replicas_identical(net.params, self.group)
        if self.norm_scale is None:
            self.norm_scale = torch.zeros(2, dtype=torch.float32, device=net.params.device)
            self._scratch = torch.empty(296 * 8, dtype=torch.uint8, device=net.params.device)

    def pack_stats(self, net):
        """Statistics pushed through the same accumulators as the gradients (sync_model.py:37-45)."""
        n, S, AP = net.n_params, net.S, net.A * net.P
        tail = net.bucket[n:]
        if net.normalize_state:
            tail[0:S].copy_(net._new_mean)
            tail[S:2 * S].copy_(net._new_std)
        tail[2 * S:2 * S + AP].copy_(net.max_active.reshape(-1))
        tail[2 * S + AP:2 * S + 2 * AP].copy_(net.sum_active.reshape(-1))

    def unpack_stats(self, net, inv_n: float):
        n, S, AP = net.n_params, net.S, net.A * net.P
        tail = net.bucket[n:]
        if inv_n != 1.0:
            tail.mul_(inv_n)
        if net.normalize_state:
            net.state_mean.copy_(tail[0:S])
            net.state_std.copy_(tail[S:2 * S])
        net.max_active.copy_(tail[2 * S:2 * S + AP].view_as(net.max_active))
        net.sum_active.copy_(tail[2 * S + AP:2 * S + 2 * AP].view_as(net.sum_active))

    def apply_gradients(self, net):
        self._lazy(net)
        st = _stream_ptr()
        # 1. local clip (before aggregation: optimizer/clip_by_global_norm/mul_* feed the accumulators)
        _cabi.check(_cabi.pfpn_clip_by_global_norm(net.grads.data_ptr(), net.n_params, self.norm_clip,
                                                   self.norm_scale.data_ptr(), self._scratch.data_ptr(),
                                                   self._scratch.numel(), st))
        self.pack_stats(net)
        if self.fused_peer and world()[1] > 1 and net.params.is_cuda:
            return self._apply_fused_peer(net, st)
        return self._apply_nccl(net, st)

    def _apply_nccl(self, net, st):
        # 2-4. one all-reduce of [clipped gradients | statistics], mean, assign statistics
        inv_n = allreduce_mean_(net.bucket, self.group)
        self.unpack_stats(net, inv_n)
        # 5. Adam (identical on every rank -> replicas stay bit-identical)
        self.step += 1
        _cabi.check(_cabi.pfpn_adam_step(net.params.data_ptr(), net.grads.data_ptr(), self.m.data_ptr(), self.v.data_ptr(),
                                         net.n_params, self.lr, self.beta1, self.beta2, self.eps, self.step, inv_n, st))
        net.global_step += 1
        # 6. train_ops chained after the optimizer step (sync_model.py:79-81): resample tick
        for op in net.train_ops:
            op()

    def _apply_fused_peer(self, net, st):
        """Steps 2-5 in one kernel over peer memory: sum in rank order, mean, Adam, averaged statistics."""
        from .peer import PeerBuckets
        if self._peers is None:
            try:
                self._peers = PeerBuckets(net.bucket.numel(), net.params.device, self.group, with_reduced=True)
            except (RuntimeError, ValueError) as e:  # collective outcome (peer.py): every rank lands here together
                import warnings
                warnings.warn(f"fused peer-memory all-reduce unavailable ({e}); using NCCL all-reduce + Adam")
                self.fused_peer = False
                return self._apply_nccl(net, st)
        pb = self._peers
        self.step += 1
        # the exchange protocol counts ITS OWN calls (1, 2, 3, ... since the buffers were created); the Adam
        # step may start anywhere (checkpoint resume)
        pb.calls += 1
        val, parity = pb.calls, pb.calls & 1
        pb.stage[parity].copy_(net.bucket)  # publish this rank's clipped bucket (device-to-device, 8.4 MB)
        buckets, flags = pb.ptrs(parity)
        _cabi.check(_cabi.pfpn_peer_signal(buckets, flags, pb.rank, pb.world, val, st))
        if pb.world >= int(os.environ.get("PFPN_PEER_TWO_PHASE_MIN", "6")):
            # reduce-scatter + all-gather inside the kernel: 2(N-1)/N bucket volumes per GPU instead of N-1
            _cabi.check(_cabi.pfpn_peer_allreduce_adam_rs(buckets, pb.reduced_ptrs, flags, pb.rank, pb.world, val,
                                                          net.n_params, net.bucket.numel(), net.params.data_ptr(),
                                                          self.m.data_ptr(), self.v.data_ptr(), net.bucket.data_ptr(), self.lr,
                                                          self.beta1, self.beta2, self.eps, self.step, st))
        else:
            _cabi.check(_cabi.pfpn_peer_allreduce_adam(buckets, flags, pb.rank, pb.world, val, net.n_params,
                                                       net.bucket.numel(), net.params.data_ptr(), self.m.data_ptr(),
                                                       self.v.data_ptr(), net.bucket.data_ptr(), self.lr, self.beta1, self.beta2,
                                                       self.eps, self.step, st))
        self.unpack_stats(net, 1.0)  # the kernel wrote the averaged bucket back
        net.global_step += 1
        for op in net.train_ops:
            op()

    def state_dict(self):
        return dict(step=self.step, m=None if self.m is None else self.m.clone(), v=None if self.v is None else self.v.clone())

    def load_state_dict(self, sd):
        self.step = int(sd["step"])
        if sd["m"] is not None:
            self.m, self.v = sd["m"].clone(), sd["v"].clone()


# ---- a20: the on-policy training loop over one rollout (models/distributed_model.py:320-345) ---------------------------
def minibatch_indices(n: int, batch_size, opt_epochs: int, rng):
    """Index arrays of the minibatches `flat_train` feeds, in order: per epoch ONE `shuffle` of arange(n) (re-shuffling the
    previous permutation, as the reference does), then contiguous slices of `batch_size` (the last one may be short);
    `batch_size` falsy -> the whole permutation as one batch.  `rng` is a numpy RandomState (the reference uses the
    process-global one, seeded per worker in distributed_model.py:568)."""
    import numpy as np
    ids = np.arange(n)
    for _ in range(opt_epochs):
        rng.shuffle(ids)
        if batch_size:
            for s in range(0, n, batch_size):
                yield ids[s:s + batch_size].copy()
        else:
            yield ids.copy()


def flat_train(net, optimizer, exp: dict, batch_size, opt_epochs: int, rng):
    """`AbstractDistributedWorker.flat_train`, on-policy branch, on device-resident rollout tensors: exp holds
    state [n,S], action [n,A], value [n], log_prob [n], advantage [n] (the PPO worker's `train_args`, workers/ppo.py:43-73).
    Every rank runs this over ITS OWN rollout, exactly like a reference worker; the optimizer aggregates.
    Returns the list of (loss, entropy, policy_loss, value_loss) device scalars, one per minibatch."""
    keys = ("state", "action", "value", "log_prob", "advantage")
    dev = net.device
    data = {k: torch.as_tensor(exp[k], dtype=torch.float32).to(dev) for k in keys}
    n = data["state"].shape[0]
    out = []
    for ids in minibatch_indices(n, batch_size, opt_epochs, rng):
        sel = torch.as_tensor(ids, dtype=torch.long, device=dev)
        out.append(net.compute_gradients(*(data[k].index_select(0, sel) for k in keys)))
        optimizer.apply_gradients(net)
    return out
