"""SAC-PFPN learner step against the fp64 oracle (oracle/sac.py) on the same minibatch, the same
Gumbel uniforms / location normals: losses, both gradient lists after the joint clip, both Adam updates, the learned
temperature and the soft target update."""
import numpy as np
import pytest
import torch

from oracle import head as oh
from oracle import sac as osac

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def make(cuda_dev, B, seed=3):
    from pfpn_b200.sac import ParticleFilteringSACNetwork
    S, A, P = 197, 36, 35
    net = ParticleFilteringSACNetwork(True, [S], [A], action_lower_bound=[-1.] * A, action_upper_bound=[1.] * A, particles=P,
                                      resample=-1, resample_interval=12000, normalize_state=True, clip_state=5.0,
                                      device=cuda_dev, seed=seed).init()
    g = torch.Generator().manual_seed(seed)
    # trained-looking weights: the 0.01 initialisation makes q almost constant, which would hide errors
    for name, (p, _) in net.named_parameters().items():
        if name.endswith("/weight") and "target_net" not in name:
            p.copy_((torch.randn(p.shape, generator=g) * (2.0 / p.shape[0]) ** 0.5).to(cuda_dev))
        elif name.endswith("/bias") and "target_net" not in name:
            p.copy_((torch.randn(p.shape, generator=g) * 0.1).to(cuda_dev))
    net.log_alpha.fill_(-0.3)
    net.target_params.copy_(net.params[:net.n_critic] * 0.9)
    net.state_mean.copy_(torch.randn(S, generator=g) * 0.1)
    net.state_std.copy_(torch.rand(S, generator=g) + 0.5)
    batch = dict(state=torch.randn(B, S, generator=g), action=torch.rand(B, A, generator=g) * 1.9 - 0.95,
                 reward=torch.randn(B, generator=g), not_terminal=(torch.rand(B, generator=g) > 0.1).float(),
                 state_=torch.randn(B, S, generator=g))
    draws = [torch.rand(B, A, P, generator=g).clamp_min(oh.F32_TINY), torch.randn(B, A, P, generator=g),
             torch.rand(B, A, P, generator=g).clamp_min(oh.F32_TINY), torch.randn(B, A, P, generator=g)]
    return net, batch, draws, (S, A, P)


def oracle_params(net):
    return {k: p.detach().double().cpu().clone() if p.dim() else p.detach().double().cpu().reshape(()).clone()
            for k, (p, _) in net.named_parameters().items()}


@pytest.mark.parametrize("B", [64, 256, 640])
def test_sac_train_step_matches_oracle(cuda_dev, B):
    from pfpn_b200.sac import SACOptimizer
    net, batch, draws, (S, A, P) = make(cuda_dev, B)
    p = oracle_params(net)
    p["global_net/alpha/log_alpha"] = p["global_net/alpha/log_alpha"].reshape(())
    mean, std = net.state_mean.double().cpu(), net.state_std.double().cpu()
    ob = tuple(batch[k].double() for k in ("state", "action", "reward", "not_terminal", "state_"))
    od = tuple(d.double() for d in draws)
    cg, ag, (l_ref, pl_ref, vl_ref), aux = osac.gradients(p, *ob, mean, std, A, P, od, gamma=0.95, value_loss_coef=0.5)
    cu = lambda t: t.to(cuda_dev)
    loss, _, pl, vl = net.compute_gradients(cu(batch["state"]), cu(batch["action"]), cu(batch["reward"]), cu(batch["not_terminal"]),
                                            cu(batch["state_"]), draws=[cu(d) for d in draws])
    assert abs(float(vl) - float(vl_ref)) < TOL * max(1.0, abs(float(vl_ref)))
    assert abs(float(pl) - float(pl_ref)) < TOL * max(1.0, abs(float(pl_ref)))
    named = net.named_parameters()
    for k, g_ref in list(cg.items()) + list(ag.items()):
        assert rel(named[k][1], g_ref) < 2e-5, k
    # the policy loss must not leak into the critic variables nor the value loss into the actor's
    assert set(cg) == {k for k in named if "/critic/q" in k and "target_net" not in k}
    # ---- optimizer: joint clip, critic Adam, actor Adam (+ log_alpha), soft target sync
    before = {k: v[0].detach().clone() for k, v in named.items()}
    opt = SACOptimizer(lr_critic=1e-4, lr_actor=1e-4, norm_clip=1.0)
    opt.apply_gradients(net)
    slots = {key: ({k: torch.zeros_like(v) for k, v in p.items()}, {k: torch.zeros_like(v) for k, v in p.items()}) for key in "ca"}
    ls, cgc, agc, norm, _ = osac.train_step(p, slots, 1, ob, mean, std, A, P, od, gamma=0.95, value_loss_coef=0.5, tau=0.005)
    assert abs(float(opt.norm_scale[0]) - norm) < 1e-5 * norm
    for k, g_ref in list(cgc.items()) + list(agc.items()):
        assert rel(named[k][1], g_ref) < 2e-5, k
    for k in named:
        if "target_net" in k:
            assert rel(named[k][0], p[k]) < 1e-6, k
        else:
            d_ref = p[k] - before[k].double().cpu().reshape(p[k].shape)
            d = named[k][0].detach().double().cpu().reshape(p[k].shape) - before[k].double().cpu().reshape(p[k].shape)
            assert float((d - d_ref).abs().max()) < 2e-6, k        # steps are <= lr = 1e-4
            cos = float((d * d_ref).sum() / (d.norm() * d_ref.norm()).clamp_min(1e-30))
            assert cos > 0.999, k
    assert net.global_step == 1 and net.train_flag == 1


def test_sac_reference_facing_calls(cuda_dev):
    from pfpn_b200.sac import SACOptimizer
    net, batch, _, (S, A, P) = make(cuda_dev, 32)
    act = net.run(None, np.zeros(S, dtype=np.float32))
    assert len(act) == 1 and act[0].shape == (A,) and np.all(np.abs(act[0]) <= 1.0)
    opt = SACOptimizer()
    t0 = net.target_params.clone()
    (loss, ent, pl, vl), extra = net.train(None, opt, None, *(batch[k].numpy() for k in ("state", "action", "reward", "not_terminal", "state_")))
    assert ent is None and np.isfinite([loss, pl, vl]).all() and abs(loss - (pl + vl)) < 1e-5 * max(1.0, abs(loss)) and extra == []
    assert not torch.equal(t0, net.target_params)  # soft update ran
    # two Philox-driven steps from identical state are identical bit for bit (same counters -> same draws; no atomics
    # anywhere on the path)
    n1, b1, _, _ = make(cuda_dev, 32)
    n2, _, _, _ = make(cuda_dev, 32)
    for n in (n1, n2):
        n.train(None, SACOptimizer(), None, *(b1[k].numpy() for k in ("state", "action", "reward", "not_terminal", "state_")))
    assert torch.equal(n1.params, n2.params) and torch.equal(n1.target_params, n2.target_params)


def test_sac_step_with_the_fused_head_kernel_equals_the_autograd_composition(cuda_dev):
    """P = 100 (BASELINE c5 shape): compute_gradients through K3f (one pass: rsample backward + tanh log_prob fwd/bwd,
    draws regenerated from the forward's Philox counters) vs the three-launch autograd composition."""
    import os
    from pfpn_b200.sac import ParticleFilteringSACNetwork
    S, A, P, B = 197, 36, 100, 300
    grads = {}
    for mode in ("1", "0"):
        os.environ["PFPN_SAC_FUSED"] = mode
        try:
            net = ParticleFilteringSACNetwork(True, [S], [A], action_lower_bound=[-1.] * A, action_upper_bound=[1.] * A,
                                              particles=P, resample=-1, resample_interval=12000, normalize_state=True,
                                              clip_state=5.0, device=cuda_dev, seed=3).init()
            g = torch.Generator().manual_seed(9)
            net.params.add_((0.05 * torch.randn(net.params.shape, generator=g)).to(cuda_dev))
            batch = (torch.randn(B, S, generator=g), torch.rand(B, A, generator=g) * 1.9 - 0.95, torch.randn(B, generator=g),
                     (torch.rand(B, generator=g) > 0.1).float(), torch.randn(B, S, generator=g))
            losses = net.compute_gradients(*batch)
            torch.cuda.synchronize()
            grads[mode] = (net.grads.clone(), [float(x) for x in losses if x is not None])
        finally:
            os.environ.pop("PFPN_SAC_FUSED", None)
    assert rel(grads["1"][0], grads["0"][0]) < 2 * TOL
    assert np.allclose(grads["1"][1], grads["0"][1], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("P,interval", [(35, 12000), (100, 12000), (35, 3)])
def test_sac_step_replayed_as_a_cuda_graph_equals_the_eager_steps(cuda_dev, P, interval):
    """GraphedSACUpdate: the whole learner step (two actor forwards, six critic evaluations, head forward / backward,
    losses, joint clip, two Adams, target sync) captured once and replayed must leave the network bit-identical to the same
    steps issued eagerly -- including the Philox offsets of the training draws, the Adam step number and the normaliser's
    step, which the kernels read from device memory.  P = 35: autograd composition; P = 100: fused K3f; interval = 3: the
    particle-resampling tick (host-side interval logic, eager, rewires fc_policy in place) falls between replays."""
    from pfpn_b200.sac import GraphedSACUpdate, ParticleFilteringSACNetwork, SACOptimizer
    S, A, B = 197, 36, 256
    mk = lambda: ParticleFilteringSACNetwork(True, [S], [A], action_lower_bound=[-1.0] * A, action_upper_bound=[1.0] * A,
                                             particles=P, resample=-1, resample_interval=interval, normalize_state=True,
                                             clip_state=5.0, device=cuda_dev, seed=5).init()
    g = torch.Generator().manual_seed(11)
    roll = torch.randn(512, S, generator=g).to(cuda_dev)
    batches = [(torch.randn(B, S, generator=g), torch.rand(B, A, generator=g) * 1.8 - 0.9, torch.randn(B, generator=g),
                (torch.rand(B, generator=g) > 0.1).float(), torch.randn(B, S, generator=g)) for _ in range(6)]
    batches = [tuple(t.to(cuda_dev) for t in b) for b in batches]
    net_e, opt_e = mk(), SACOptimizer()
    net_e.run_batch(roll)  # activity statistics for the resampler
    losses_e = []
    for b in batches:
        out = net_e.compute_gradients(*b)
        opt_e.apply_gradients(net_e)
        losses_e.append(float(out[0]))
    net_g, opt_g = mk(), SACOptimizer()
    net_g.run_batch(roll)
    gu = GraphedSACUpdate(net_g, opt_g, B, warmup=2)
    losses_g = []
    for b in batches:
        out = gu.run(*b)
        losses_g.append(float(out[0]))
    torch.cuda.synchronize()
    assert gu.replays == 4
    assert losses_e == losses_g
    assert torch.equal(net_e.params, net_g.params)
    assert torch.equal(net_e.target_params, net_g.target_params)
    assert torch.equal(net_e.state_mean, net_g.state_mean) and torch.equal(net_e.state_std, net_g.state_std)
    assert torch.equal(opt_e.m, opt_g.m) and torch.equal(opt_e.v, opt_g.v)
    assert net_e.global_step == net_g.global_step == 6 and int(net_g._gstep_dev.item()) == 6 and int(opt_g._step_dev.item()) == 6
    assert int(net_g._train_rng.item()) == 24
    # different steps draw different variates (the device word advances inside the graph)
    assert len(set(losses_g)) == len(losses_g)
