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
    port = 29500 + os.getpid() % 2000
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
