"""World-size-2 gloo test (CPU) of the data-parallel host logic: shard bounds, the flat-bucket
all-reduce-mean, and equality with the single-process N-shard emulation of the reference's
SyncReplicasOptimizer step (oracle/network.py::sync_update)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _tiny_params(S, H1, H2, A, P, g):
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64) * 0.1
    return {"global_net/actor/fc1/weight": r(S, H1), "global_net/actor/fc1/bias": r(H1),
            "global_net/actor/fc2/weight": r(H1, H2), "global_net/actor/fc2/bias": r(H2),
            "global_net/actor/samples": torch.linspace(-1, 1, P, dtype=torch.float64).repeat(A, 1),
            "global_net/actor/samples_std": torch.full((A, P), -1.0, dtype=torch.float64),
            "global_net/actor/fc_policy/weight": r(H2, A * P), "global_net/actor/fc_policy/bias": r(A * P),
            "global_net/critic/fc1/weight": r(S, H1), "global_net/critic/fc1/bias": r(H1),
            "global_net/critic/fc2/weight": r(H1, H2), "global_net/critic/fc2/bias": r(H2),
            "global_net/critic/fc3/weight": r(H2, 1), "global_net/critic/fc3/bias": r(1)}


def _batch(B, S, A, g):
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    return dict(state=r(B, S), action=torch.rand(B, A, generator=g, dtype=torch.float64) * 2 - 1, value=r(B),
                log_prob=r(B) * 0.1 + 4.0, advantage=r(B))


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import network as on
    from pfpn_b200.learner import allreduce_mean_, shard_bounds, world as get_world
    S, H1, H2, A, P, B = 5, 8, 6, 2, 4, 22
    g = torch.Generator().manual_seed(0)
    p = _tiny_params(S, H1, H2, A, P, g)
    batch = _batch(B, S, A, g)
    mean, std = torch.zeros(S, dtype=torch.float64), torch.ones(S, dtype=torch.float64)
    assert get_world() == (rank, world)
    lo, hi = shard_bounds(B, rank, world)
    grads, _ = on.gradients(p, *(batch[k][lo:hi] for k in ("state", "action", "value", "log_prob", "advantage")), mean, std, A, P)
    grads, _ = on.clip_by_global_norm(grads, 1.0)
    nm, ns = on.normalizer_update(mean, std, batch["state"][lo:hi], 0)
    keys = sorted(grads)
    bucket = torch.cat([grads[k].reshape(-1) for k in keys] + [nm, ns])  # [gradients | statistics]
    inv = allreduce_mean_(bucket)
    bucket *= inv
    if rank == 0:
        q.put((keys, bucket.clone(), lo, hi))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bucket_allreduce_equals_sync_replicas_emulation():
    sys.path.insert(0, ROOT)
    from oracle import network as on
    from pfpn_b200.learner import shard_bounds
    assert [shard_bounds(22, r, 2) for r in range(2)] == [(0, 11), (11, 22)]
    assert [shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    keys, bucket, lo, hi = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    S, H1, H2, A, P, B = 5, 8, 6, 2, 4, 22
    g = torch.Generator().manual_seed(0)
    p = _tiny_params(S, H1, H2, A, P, g)
    batch = _batch(B, S, A, g)
    mean, std = torch.zeros(S, dtype=torch.float64), torch.ones(S, dtype=torch.float64)
    shards = [{k: v[a:b] for k, v in batch.items()} for a, b in (shard_bounds(B, r, 2) for r in range(2))]
    m = {k: torch.zeros_like(v) for k, v in p.items()}
    v = {k: torch.zeros_like(t) for k, t in p.items()}
    acc, nm, ns, _ = on.sync_update({k: t.clone() for k, t in p.items()}, m, v, 1, shards, mean, std, A, P)
    ref = torch.cat([acc[k].reshape(-1) for k in keys] + [nm, ns])
    assert torch.allclose(bucket, ref, rtol=1e-12, atol=1e-14)


def test_minibatch_indices_equal_the_reference_flat_train_loop():
    """a20: same shuffles, same slices, same order as models/distributed_model.py:320-345 for the same RandomState."""
    import numpy as np
    from pfpn_b200.learner import minibatch_indices

    def reference_loop(n, batch_size, opt_epochs, rng):  # literal restatement of the on-policy branch
        ids = np.arange(n)
        out, epoch = [], 0
        while epoch < opt_epochs:
            rng.shuffle(ids)
            if batch_size:
                for s in range(0, n, batch_size):
                    out.append(ids[s:s + batch_size].copy())
            else:
                out.append(ids[ids >= 0].copy())
            epoch += 1
        return out

    for n, bs, ep in ((368, 32, 5), (100, 32, 3), (10, None, 2), (7, 7, 1), (5, 8, 2)):
        a = list(minibatch_indices(n, bs, ep, np.random.RandomState(28949)))
        b = reference_loop(n, bs, ep, np.random.RandomState(28949))
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
        assert sum(len(x) for x in a) == n * ep


# ---- SyncReplicasAdam.apply_gradients itself (pack_stats / exchange / unpack_stats / Adam / train_ops) under gloo -------------
class _HostCabi:
    """Stands in for the two CUDA entry points apply_gradients calls, on HOST pointers (numpy over ctypes), so the
    collective host logic of the PRODUCT class runs under gloo without a GPU.  Same semantics as csrc/optim.cu."""

    @staticmethod
    def _arr(ptr, n):
        import ctypes as C
        import numpy as np
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(n,))

    def check(self, status):
        assert status == 0

    def pfpn_clip_by_global_norm(self, grads, n, clip, norm_scale, scratch, scratch_bytes, st):
        import numpy as np
        g, ns = self._arr(grads, n), self._arr(norm_scale, 2)
        norm = float(np.sqrt(np.sum(g.astype(np.float64) ** 2)))
        scale = clip * min(1.0 / norm, 1.0 / clip)
        g *= np.float32(scale)
        ns[0], ns[1] = norm, scale
        return 0

    def pfpn_adam_step(self, params, grads, m, v, n, lr, b1, b2, eps, step, grad_scale, st):
        import math
        import numpy as np
        p, g, mm, vv = (self._arr(x, n) for x in (params, grads, m, v))
        gs = g.astype(np.float64) * grad_scale
        lr_t = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
        m64 = b1 * mm.astype(np.float64) + (1 - b1) * gs
        v64 = b2 * vv.astype(np.float64) + (1 - b2) * gs * gs
        mm[:], vv[:] = m64, v64
        p[:] = p.astype(np.float64) - lr_t * m64 / (np.sqrt(v64) + eps)
        return 0


class _HostNet:
    """The attributes SyncReplicasAdam touches, with host tensors."""

    def __init__(self, n_params, S, A, P, rank):
        g = torch.Generator().manual_seed(3)
        self.n_params, self.S, self.A, self.P, self.normalize_state = n_params, S, A, P, True
        self.params = torch.randn(n_params, generator=g)
        self.bucket = torch.zeros(n_params + 2 * S + 2 * A * P)
        self.grads = self.bucket[:n_params]
        gr = torch.Generator().manual_seed(100 + rank)
        self.grads.copy_(torch.randn(n_params, generator=gr) * 3)
        self._new_mean, self._new_std = torch.randn(S, generator=gr), torch.rand(S, generator=gr) + 0.5
        self.state_mean, self.state_std = torch.zeros(S), torch.ones(S)
        self.max_active, self.sum_active = torch.rand(A, P, generator=gr), torch.rand(A, P, generator=gr) * 9
        self.global_step, self.ticks = 0, 0
        self.train_ops = [self._tick]

    def _tick(self):
        self.ticks += 1


def _adam_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pfpn_b200 import learner
    learner._cabi = _HostCabi()
    learner._stream_ptr = lambda: 0
    net = _HostNet(64, 5, 2, 4, rank)
    local = dict(g=net.grads.clone(), mean=net._new_mean.clone(), std=net._new_std.clone(), mx=net.max_active.clone(),
                 sm=net.sum_active.clone(), p0=net.params.clone())
    opt = learner.SyncReplicasAdam(lr=1e-3, norm_clip=1.0, fused_peer=False)
    opt.apply_gradients(net)
    q.put((rank, local, dict(params=net.params.clone(), mean=net.state_mean.clone(), std=net.state_std.clone(),
                             mx=net.max_active.clone(), sm=net.sum_active.clone(), step=opt.step, gstep=net.global_step,
                             ticks=net.ticks, norm=float(opt.norm_scale[0]))))
    # replicas that start different must be refused (no parameter broadcast on this path)
    bad = _HostNet(64, 5, 2, 4, rank)
    bad.params[3] += float(rank)
    try:
        learner.SyncReplicasAdam(fused_peer=False).apply_gradients(bad)
        q.put((rank, "no error", None))
    except RuntimeError as e:
        q.put((rank, "refused" if "replicas hold different" in str(e) else repr(e), None))
    dist.barrier()
    dist.destroy_process_group()


def test_sync_replicas_adam_apply_gradients_two_ranks_gloo():
    """The product's SyncReplicasAdam (NCCL-path code, here over gloo): local clip BEFORE aggregation, mean of the clipped
    gradients and of the four pushed statistics (max_active averaged, not max-reduced: sync_model.py:92-96), identical Adam on
    both ranks, train_ops chained after the step."""
    import math
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_adam_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    got = [q.get(timeout=120) for _ in range(4)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    res = {r: (loc, out) for r, loc, out in got if isinstance(loc, dict)}
    refusals = [loc for _, loc, out in got if not isinstance(loc, dict)]
    assert refusals == ["refused", "refused"]
    clipped = []
    for r in range(2):
        g = res[r][0]["g"].double()
        norm = float(g.norm())
        assert abs(res[r][1]["norm"] - norm) < 1e-4 * norm
        clipped.append(g * min(1.0 / norm, 1.0))
    gm = (clipped[0] + clipped[1]) / 2
    lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
    m, v = 0.1 * gm, 0.001 * gm * gm
    p_ref = res[0][0]["p0"].double() - lr_t * m / (v.sqrt() + 1e-8)
    for r in range(2):
        out = res[r][1]
        assert torch.allclose(out["params"].double(), p_ref, rtol=1e-5, atol=1e-6)
        for key, loc_key in (("mean", "mean"), ("std", "std"), ("mx", "mx"), ("sm", "sm")):
            want = (res[0][0][loc_key] + res[1][0][loc_key]) / 2
            assert torch.allclose(out[key], want, rtol=1e-6, atol=1e-7), key
        assert out["step"] == 1 and out["gstep"] == 1 and out["ticks"] == 1
    assert torch.equal(res[0][1]["params"], res[1][1]["params"])  # replicas bit-identical


def test_peer_gather_bookkeeping_without_a_gpu(monkeypatch):
    """PeerGather's host-side protocol (exchange numbering, rotating slots, one-exchange lag, who consumes what) with the
    device calls stubbed: push v may carry the sum of v-1 (consume_into), never more than one exchange lags, reduce()
    consumes the oldest, and slot(v) = v mod 4 so that v+3 never lands on an unconsumed slot."""
    import ctypes as C
    from pfpn_b200 import _cabi, peer

    calls = []
    monkeypatch.setattr(_cabi, "pfpn_peer_gather_sum_packets",
                        lambda rows, nr, value, n, out, scale, st: calls.append((rows, nr, value, n)) or 0)
    pg = object.__new__(peer.PeerGather)
    pg.rank, pg.world, pg.n, pg.dev = 1, 4, 2520, torch.device("cpu")
    pg.pushed = pg.consumed = 0
    pg._gather_ptr = 1 << 20
    pg._gather_ptrs = [(r + 1) << 24 for r in range(4)]
    pg._push = [pg._make_push(s) for s in range(pg.NBUF)]
    out = torch.zeros(2520)
    p1 = pg.push_args(consume_into=out)            # nothing to consume yet
    assert (p1.value, p1.consume_value, p1.protocol, p1.nranks) == (1, 0, 1, 4) and pg.pending == 1
    assert p1.out[2] == pg._gather_ptrs[2] + ((1 * 4 + 1) * 2520) * 8   # slot 1, my row (rank 1), 8-byte packets
    p2 = pg.push_args(consume_into=out, scale=0.25)
    assert (p2.value, p2.consume_value) == (2, 1) and pg.pending == 1
    assert p2.consume_rows == pg._gather_ptr + 1 * 4 * 2520 * 8 and p2.consume_out == out.data_ptr()
    assert abs(p2.consume_scale - 0.25) < 1e-9
    p3 = pg.push_args()                             # producer only: exchange 2 stays pending
    assert (p3.value, p3.consume_value) == (3, 0) and pg.pending == 2
    with pytest.raises(RuntimeError):
        pg.push_args()                              # a third unconsumed exchange would overwrite a live slot
    pg.reduce(out)
    pg.reduce(out)
    assert [c[2] for c in calls] == [2, 3] and pg.pending == 0
    assert calls[0][0] == pg._gather_ptr + 2 * 4 * 2520 * 8 and calls[1][0] == pg._gather_ptr + 3 * 4 * 2520 * 8
    with pytest.raises(RuntimeError):
        pg.reduce(out)
    for v in range(4, 12):                          # steady state: slots rotate 0,1,2,3,...
        p = pg.push_args(consume_into=out)
        assert p.value == v and pg.slot_of(v) == v % 4 and (p.consume_value == v - 1 or v == 4)
