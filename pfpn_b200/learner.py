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
    bits = params.detach().contiguous().reshape(-1).view(torch.int32).to(torch.int64)
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
            self.launches_last_step = None

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
        """One optimizer step on the gradients `net.compute_gradients` left in the bucket, then the network's train_ops."""
        self._launch_step(net)
        self._after_step(net)

    # The step is split in two so that the device work can be captured in a CUDA graph (GraphedUpdate below):
    # `_launch_step` only enqueues kernels whose arguments never change; `_after_step` is the host bookkeeping.
    def _launch_step(self, net):
        self._lazy(net)
        st = _stream_ptr()
        n_world = world()[1] if net.params.is_cuda else 1
        self._mode = "sync_step"
        if os.environ.get("PFPN_SYNC_STEP", "1") == "0" or not hasattr(net, "dev_counters"):
            self._mode = "legacy"
        elif n_world > 1:
            if not self.fused_peer:
                self._mode = "legacy"
            elif self._peers is None:
                from .peer import PeerBuckets
                try:
                    self._peers = PeerBuckets(net.bucket.numel(), net.params.device, self.group, with_reduced=True)
                except (RuntimeError, ValueError) as e:  # collective outcome (peer.py): every rank lands here together
                    import warnings
                    warnings.warn(f"fused peer-memory all-reduce unavailable ({e}); using NCCL all-reduce + Adam")
                    self.fused_peer = False
                    self._mode = "legacy"
        if self._mode == "legacy":
            return self._launch_legacy(net, st)
        # ---- three launches: sum of squares (+ counters), clip -> stage -> flags, rank-ordered mean -> Adam -> statistics
        a = self._sync_args(net, n_world)
        net._ensure_counters(self._peers.calls if self._peers is not None else None, self.step)
        _cabi.check(_cabi.pfpn_sync_step(a, st))
        self.launches_last_step = 3

    def _sync_args(self, net, n_world):
        key = (id(net), n_world, int(os.environ.get("PFPN_PEER_TWO_PHASE_MIN", "6")))
        if getattr(self, "_sync_key", None) == key:
            return self._sync_a
        import ctypes as C
        a = _cabi.SyncArgs()
        a.grads, a.n_params, a.n_total, a.clip = net.bucket.data_ptr(), net.n_params, net.bucket.numel(), self.norm_clip
        if net.normalize_state:
            a.new_mean, a.new_std = net._new_mean.data_ptr(), net._new_std.data_ptr()
            a.state_mean, a.state_std, a.S = net.state_mean.data_ptr(), net.state_std.data_ptr(), net.S
        a.max_active, a.sum_active, a.AP = net.max_active.data_ptr(), net.sum_active.data_ptr(), net.A * net.P
        a.params, a.m, a.v = net.params.data_ptr(), self.m.data_ptr(), self.v.data_ptr()
        if getattr(net, "use_presplit", False) and getattr(net, "use_tensor_cores", False):
            a.params_lo = net.params_lo.data_ptr()  # the step keeps the GEMMs' pre-split weight halves current
        a.lr, a.beta1, a.beta2, a.eps = self.lr, self.beta1, self.beta2, self.eps
        a.counters, a.norm_scale = net.dev_counters.data_ptr(), self.norm_scale.data_ptr()
        a.scratch, a.scratch_bytes = self._scratch.data_ptr(), self._scratch.numel()
        a.rank, a.nranks, a.two_phase = 0, 1, 0
        if n_world > 1:
            pb = self._peers
            stage, flags = pb.ptrs(0)
            self._keep = (stage, flags, pb.reduced_ptrs)
            a.stage = C.cast(stage, C.c_void_p)
            a.flags = C.cast(flags, C.c_void_p)
            a.reduced = C.cast(pb.reduced_ptrs, C.c_void_p)
            a.rank, a.nranks = pb.rank, pb.world
            a.two_phase = 1 if pb.world >= key[2] else 0
        self._sync_key, self._sync_a = key, a
        return a

    def _after_step(self, net):
        if self._mode == "sync_step":
            self.step += 1
            if self._peers is not None:
                self._peers.calls += 1
            net.global_step += 1
            net._dev_shadow = [net._dev_shadow[0] + 1, net._dev_shadow[1] + 1, net._dev_shadow[2] + 1]
        # train_ops chained after the optimizer step (sync_model.py:79-81): resample tick
        for op in net.train_ops:
            op()

    def _launch_legacy(self, net, st):
        if hasattr(net, "invalidate_lo"):
            net.invalidate_lo()  # these kernels update the parameters through raw pointers without the low halves
        return self._launch_legacy_impl(net, st)

    def _launch_legacy_impl(self, net, st):
        """Round-1 chain, kept as the comparison point (PFPN_SYNC_STEP=0) and as the NCCL fallback: clip, pack, exchange
        (peer kernel or NCCL all-reduce), Adam, unpack -- host-side step numbers in the arguments."""
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

    def state_dict(self):
        return dict(step=self.step, m=None if self.m is None else self.m.clone(), v=None if self.v is None else self.v.clone())

    def load_state_dict(self, sd):
        self.step = int(sd["step"])
        if sd["m"] is not None:
            if self.m is not None and self.m.shape == sd["m"].shape:
                self.m.copy_(sd["m"])  # in place: kernels / captured graphs hold these pointers
                self.v.copy_(sd["v"])
            else:
                self.m, self.v = sd["m"].clone(), sd["v"].clone()
                self._sync_key = None


class GraphedUpdate:
    """One DPPO minibatch update -- `net.compute_gradients(...)` + `optimizer.apply_gradients(net)` -- captured ONCE in a
    CUDA graph and replayed (at 8192 states per GPU the update is a chain of ~45 launch-latency-sized kernels).
    Possible because nothing in the captured kernels' arguments changes between steps: the Adam step, the exchange call
    number / staging parity and the normaliser's step are read from device memory (csrc/syncstep.cu).  The minibatch is
    copied into fixed input buffers; the resample tick (host-side interval logic, a2c.py:370-383) runs eagerly after
    the replay.  Every rank must call `run` the same number of times (the replayed step contains the exchange)."""

    KEYS = ("state", "action", "value", "log_prob", "advantage")

    def __init__(self, net, optimizer, batch: int, warmup: int = 2):
        dev = net.device
        self.net, self.opt, self.B = net, optimizer, int(batch)
        f = lambda *s: torch.zeros(*s, dtype=torch.float32, device=dev)
        self.inputs = dict(state=f(batch, net.S), action=f(batch, net.A), value=f(batch), log_prob=f(batch), advantage=f(batch))
        self.graph = None
        self.losses = None
        self._warm = int(warmup)
        self.replays = 0

    def _set(self, state, action, value, log_prob, advantage):
        for k, v in zip(self.KEYS, (state, action, value, log_prob, advantage)):
            if v is not None and v is not self.inputs[k]:
                self.inputs[k].copy_(torch.as_tensor(v, dtype=torch.float32).reshape(self.inputs[k].shape), non_blocking=True)

    def run(self, state=None, action=None, value=None, log_prob=None, advantage=None):
        """Arguments left None keep what is already in ``self.inputs`` (zero-copy use: write the minibatch there)."""
        self._set(state, action, value, log_prob, advantage)
        net, opt = self.net, self.opt
        args = tuple(self.inputs[k] for k in self.KEYS)
        # Replay pays where the update is a chain of launch-latency-sized kernels (<= ~16k states per GPU: 0.69 vs 0.71 ms at
        # 8192, 1.18 vs 1.22 at 16384); at 65536 states the eager multi-stream issue order overlaps the branches better
        # (4.43 vs 4.73 ms measured), so large shards stay eager.
        too_big = self.B > int(os.environ.get("PFPN_GRAPH_MAX_BATCH", "24576"))
        if self._warm > 0 or too_big or os.environ.get("PFPN_GRAPH", "1") == "0":  # eager steps (also: allocate buffers, map peers)
            self._warm -= 1
            self.losses = net.compute_gradients(*args)
            opt.apply_gradients(net)
            return self.losses
        if self.graph is None:
            net._ensure_counters(opt._peers.calls if opt._peers is not None else None, opt.step)
            torch.cuda.synchronize(net.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                self.losses = net.compute_gradients(*args)
                opt._launch_step(net)
            if opt._mode != "sync_step":
                raise RuntimeError("the graph-captured update needs the device-counter step (PFPN_SYNC_STEP=1, peer memory available)")
        self.graph.replay()
        self.replays += 1
        opt._after_step(net)
        return self.losses


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
