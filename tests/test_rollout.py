"""Rollout post-processing: GAE / value target.  CPU: the oracle against golden vectors
produced by executing the reference's own functions; GPU: the kernel against the oracle, bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import rollout as orr

GOLD = os.path.join(os.path.dirname(__file__), "golden", "gae_golden.npz")


def cases():
    g = np.load(GOLD)
    i = 0
    while f"c{i}_reward" in g:
        yield {k: g[f"c{i}_{k}"] for k in ("reward", "value", "gamma", "gae_gamma", "adv", "target")}
        i += 1


def test_oracle_matches_executed_reference_bit_exact():
    n = 0
    for c in cases():
        adv = orr.generalized_advantage_estimate(c["reward"], c["value"], float(c["gamma"]), float(c["gae_gamma"]))
        assert adv.dtype == np.float32
        assert np.array_equal(adv.astype(np.float64), c["adv"])
        assert np.array_equal(orr.value_target_estimate(c["value"][:-1], adv).astype(np.float64), c["target"])
        n += 1
    assert n == 5


def test_float64_scan_of_the_numpy1_era_is_within_tolerance():
    for c in cases():
        r, v = c["reward"].astype(np.float32), c["value"].astype(np.float32)
        td = (r + np.float32(c["gamma"]) * v[1:] - v[:-1]).astype(np.float32).astype(np.float64)
        run, ref = 0.0, np.empty(len(td))
        for t in range(len(td) - 1, -1, -1):
            run = td[t] + float(c["gae_gamma"]) * run
            ref[t] = run
        assert np.abs(ref - c["adv"]).max() <= 1e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("E,T,gamma,lambd", [(1, 1, 0.99, 0.95), (5, 368, 0.95, 0.95), (300, 64, 0.99, None), (4096, 32, 0.9, 1.0)])
def test_gae_kernel_bit_exact(cuda_dev, E, T, gamma, lambd):
    from pfpn_b200.network import ParticleFilteringClipPPONetwork
    net = ParticleFilteringClipPPONetwork(True, [8], [2], action_lower_bound=[-1.] * 2, action_upper_bound=[1.] * 2, particles=5,
                                          resample=-1, actor_net_shape=[16], critic_net_shape=[16], gamma=gamma, lambd=lambd,
                                          device=cuda_dev, seed=1).init()
    rng = np.random.RandomState(E + T)
    reward = rng.randn(E, T).astype(np.float32)
    value = (rng.randn(E, T + 1) * 3).astype(np.float32)
    adv, tgt = net.advantage_and_value_target(reward, value)
    gg = 0.0 if lambd is None else gamma * lambd
    for e in range(min(E, 40)):
        ref = orr.generalized_advantage_estimate(reward[e], value[e], gamma, gg)
        assert np.array_equal(adv[e].cpu().numpy(), ref)
        assert np.array_equal(tgt[e].cpu().numpy(), orr.value_target_estimate(value[e][:-1], ref))
    a1 = net.generalized_advantage_estimate(reward[0], value[0])
    assert a1.shape == (T,) and np.array_equal(a1.cpu().numpy(), adv[0].cpu().numpy())


@pytest.mark.gpu
def test_gae_kernel_on_the_golden_vectors(cuda_dev):
    from pfpn_b200 import _cabi
    from pfpn_b200.head import _stream_ptr
    for c in cases():
        r = torch.tensor(c["reward"], device=cuda_dev)
        v = torch.tensor(c["value"], device=cuda_dev)
        adv, tgt = torch.empty_like(r), torch.empty_like(r)
        _cabi.check(_cabi.pfpn_gae(r.data_ptr(), v.data_ptr(), adv.data_ptr(), tgt.data_ptr(), 1, r.numel(), float(c["gamma"]),
                                   float(c["gae_gamma"]), _stream_ptr()))
        assert np.array_equal(adv.cpu().numpy().astype(np.float64), c["adv"])
        assert np.array_equal(tgt.cpu().numpy().astype(np.float64), c["target"])


def test_host_pipeline_chunk_schedule_covers_the_batch_and_ramps():
    """HostHeadPipeline._schedule (pure host logic): contiguous cover of [0, B), no chunk above the slot size, small first
    and last chunks when the batch is large enough (pipeline fill / drain), plain equal chunks otherwise."""
    from pfpn_b200.host import HostHeadPipeline
    for B, chunk in [(65536, 8192), (65537, 8192), (30000, 8192), (20000, 8192), (8192, 8192), (1000, 1000), (5, 5), (100000, 4096)]:
        b = HostHeadPipeline._schedule(B, chunk)
        assert b[0][0] == 0 and b[-1][1] == B
        assert all(lo < hi and hi - lo <= chunk for lo, hi in b)
        assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
    big = [hi - lo for lo, hi in HostHeadPipeline._schedule(65536, 8192)]
    assert big[0] == 1024 and big[-1] == 1024 and max(big) == 8192
    assert [hi - lo for lo, hi in HostHeadPipeline._schedule(20000, 8192)] == [8192, 8192, 3616]
