"""K6 / K7 and the DPPO train step through the reference-facing network object, against the fp64
oracle (oracle/network.py).  Tolerance: 1e-5 norm-wise per tensor."""
import numpy as np
import pytest
import torch

from oracle import network as on
from pfpn_b200 import _cabi
from pfpn_b200.head import _stream_ptr
from pfpn_b200.learner import SyncReplicasAdam
from pfpn_b200.network import ParticleFilteringClipPPONetwork

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("M,K,N", [(300, 200, 1024), (1000, 1024, 512), (517, 512, 1260), (129, 512, 1), (4, 8, 4)])
def test_linear_layers_match_fp64(cuda_dev, M, K, N):
    g = torch.Generator().manual_seed(M + N)
    X = torch.randn(M, K, generator=g).clamp(-3, 6.5)
    W = torch.randn(K, N, generator=g) * 0.05
    b = torch.randn(N, generator=g) * 0.1
    dY = torch.randn(M, N, generator=g)
    cu = lambda t: t.to(cuda_dev).contiguous()
    Xd, Wd, bd, dYd = cu(X), cu(W), cu(b), cu(dY)
    st = _stream_ptr()
    for relu6 in ((0, 1) if N > 1 else (0,)):
        Y = torch.empty(M, N, device=cuda_dev) if N > 1 else torch.empty(M, device=cuda_dev)
        _cabi.check(_cabi.pfpn_mlp_linear_fwd(Xd.data_ptr(), K, Wd.data_ptr(), bd.data_ptr(), Y.data_ptr(), N, M, K, N, relu6, st))
        ref = X.double() @ W.double() + b.double()
        ref = ref.clamp(0, 6) if relu6 else ref
        assert rel(Y.reshape(M, N), ref) < TOL
    dX = torch.empty(M, K, device=cuda_dev)
    _cabi.check(_cabi.pfpn_mlp_linear_bwd_input(dYd.data_ptr(), N, Wd.data_ptr(), Xd.data_ptr(), dX.data_ptr(), K, M, K, N, st))
    mask = ((X > 0) & (X < 6)).double()
    assert rel(dX, (dY.double() @ W.double().T) * mask) < TOL
    import ctypes as C
    n = C.c_size_t(0)
    _cabi.check(_cabi.pfpn_mlp_wgrad_workspace_bytes(M, K, N, C.byref(n)))
    ws = torch.empty(n.value, dtype=torch.uint8, device=cuda_dev)
    dW, db = torch.empty(K, N, device=cuda_dev), torch.empty(N, device=cuda_dev)
    _cabi.check(_cabi.pfpn_mlp_linear_bwd_weight(Xd.data_ptr(), K, dYd.data_ptr(), N, dW.data_ptr(), db.data_ptr(), M, K, N,
                                                 ws.data_ptr(), ws.numel(), st))
    assert rel(dW, X.double().T @ dY.double()) < TOL and rel(db, dY.double().sum(0)) < TOL


def make_batch(B, S, A, seed):
    g = torch.Generator().manual_seed(seed)
    return dict(state=torch.randn(B, S, generator=g) * 1.5 + 0.3, action=torch.rand(B, A, generator=g) * 2 - 1,
                value=torch.randn(B, generator=g), log_prob=torch.randn(B, generator=g) * 0.3 + 60.0,
                advantage=torch.randn(B, generator=g))


def build(cuda_dev, S=197, A=36, P=35, **kw):
    net = ParticleFilteringClipPPONetwork(True, [S], [A], action_lower_bound=[-1.0] * A, action_upper_bound=[1.0] * A,
                                          particles=P, resample=-1, resample_interval=368, normalize_state=True,
                                          clip_state=5.0, normalize_advantage=True, value_loss_coef=0.5, device=cuda_dev,
                                          seed=28949, **kw).init()
    # perturb so every parameter matters (bias 0 / equal particles would hide indexing bugs)
    g = torch.Generator().manual_seed(1)
    net.params.add_(0.02 * torch.randn(net.params.shape, generator=g).to(cuda_dev))
    for l in net.actor + net.critic[:1]:
        pass
    net.actor[0].W[net.S:].zero_(); net.critic[0].W[net.S:].zero_()
    net.state_mean.copy_(torch.randn(S, generator=g) * 0.2)
    net.state_std.copy_(torch.rand(S, generator=g) + 0.5)
    return net


def oracle_params(net):
    return {k: p.detach().double().cpu().clone() for k, (p, _) in net.named_parameters().items()}


@pytest.mark.parametrize("B", [512, 33])
def test_train_step_gradients_and_adam_match_oracle(cuda_dev, B):
    net = build(cuda_dev)
    batch = make_batch(B, 197, 36, seed=B)
    p0 = oracle_params(net)
    mean, std = net.state_mean.double().cpu(), net.state_std.double().cpu()
    b64 = {k: v.double() for k, v in batch.items()}
    # make lp_old realistic: behaviour log-prob near the current one
    logits, _ = on.forward(p0, b64["state"], mean, std)
    from oracle import head as oh
    lp = oh.MixtureGaussianOracle(logits.reshape(B, 36, 35), p0["global_net/actor/samples"],
                                  p0["global_net/actor/samples_std"].exp(), False).log_prob(b64["action"])
    batch["log_prob"] = (lp + 0.1 * torch.randn(B, dtype=torch.float64)).float()
    b64["log_prob"] = batch["log_prob"].double()
    g_ref, l_ref = on.gradients(p0, b64["state"], b64["action"], b64["value"], b64["log_prob"], b64["advantage"],
                                mean, std, 36, 35)
    opt = SyncReplicasAdam(lr=1e-4, norm_clip=1.0)
    losses = net.compute_gradients(batch["state"], batch["action"], batch["value"], batch["log_prob"], batch["advantage"])
    for k, (_, g) in net.named_parameters().items():
        assert rel(g, g_ref[k]) < TOL, k
    assert abs(float(losses[3]) - float(l_ref[3])) < TOL * float(l_ref[3])
    assert abs(float(losses[0]) - float(l_ref[0])) < 1e-4 * max(1.0, abs(float(l_ref[0])))
    # clip + Adam + statistics
    m = {k: torch.zeros_like(v) for k, v in p0.items()}
    v = {k: torch.zeros_like(t) for k, t in p0.items()}
    p1 = {k: t.clone() for k, t in p0.items()}
    _, norm = on.clip_by_global_norm(g_ref, 1.0)
    # Adam normalises by |g|: near-zero gradients amplify any gradient noise to O(1) in the update, so
    # the optimizer kernels are checked on the kernel's own (fp32) gradient
    g_k = {k: g.detach().double().cpu().clone() for k, (_, g) in net.named_parameters().items()}
    gc, norm_k = on.clip_by_global_norm(g_k, 1.0)
    assert abs(norm_k - norm) < TOL * norm
    on.adam_step(p1, gc, m, v, 1)
    nm, ns = on.normalizer_update(mean, std, b64["state"], 0)
    opt.apply_gradients(net)
    assert abs(float(opt.norm_scale[0]) - norm) < TOL * norm
    for k, (p, _) in net.named_parameters().items():
        # fp32 parameter storage bounds the agreement: |dp| = 1e-4 on weights up to ~1
        assert rel(p, p1[k]) < 1e-6, k
        du, dr = (p.double().cpu() - p0[k]).reshape(-1), (p1[k] - p0[k]).reshape(-1)
        assert float(torch.dot(du, dr) / (du.norm() * dr.norm())) > 0.999, k
    assert rel(net.state_mean, nm) < TOL and rel(net.state_std, ns) < TOL
    assert net.global_step == 1 and net.train_flag == 1


def test_reference_facing_train_and_run_signatures(cuda_dev):
    net = build(cuda_dev)
    opt = SyncReplicasAdam()
    batch = make_batch(64, 197, 36, seed=3)
    np_batch = {k: v.numpy() for k, v in batch.items()}
    (loss, ent, pl, vl), extra = net.train(None, opt, None, np_batch["state"], np_batch["action"], np_batch["value"],
                                           np_batch["log_prob"], np_batch["advantage"])
    assert ent is None and np.isfinite([loss, pl, vl]).all() and extra == []
    out = net.run(None, np_batch["state"][0])
    assert out[0].shape == (36,) and isinstance(out[1], float) and isinstance(out[2], float)
    assert isinstance(net.evaluate(None, np_batch["state"][0]), float)
    assert float(net.sum_active.sum()) > 0  # running_update_ops executed on run()
    assert len(net.local_update_variables) == 4 and len(net.train_ops) == 1


def test_resample_tick_fires_at_interval_and_is_replica_deterministic(cuda_dev):
    nets = [build(cuda_dev) for _ in range(2)]
    for net in nets:
        net.resample_interval = 3
        net.max_active.fill_(0.5); net.sum_active.fill_(1.0)
        net.max_active[:, ::5] = 1e-6
        fired = []
        for i in range(3):
            net.global_step = i
            fired.append(net.update())
        assert fired == [False, False, True] and net.train_flag == 0 and not net.max_active.any()
    assert torch.equal(nets[0].params, nets[1].params)


def test_flat_train_runs_the_reference_minibatch_schedule(cuda_dev):
    """a20: flat_train == looping compute_gradients/apply_gradients over the reference's shuffled slices."""
    import numpy as np
    from pfpn_b200.learner import SyncReplicasAdam, flat_train, minibatch_indices
    from pfpn_b200.network import ParticleFilteringClipPPONetwork
    S, A, P, n = 197, 36, 35, 96

    def make():
        return ParticleFilteringClipPPONetwork(True, [S], [A], action_lower_bound=[-1.] * A, action_upper_bound=[1.] * A, particles=P,
                                               resample=-1, resample_interval=368, normalize_state=True, clip_state=5.0,
                                               normalize_advantage=True, device=cuda_dev, seed=3).init()
    g = torch.Generator().manual_seed(0)
    exp = dict(state=torch.randn(n, S, generator=g), action=torch.rand(n, A, generator=g) * 2 - 1, value=torch.randn(n, generator=g),
               log_prob=torch.randn(n, generator=g) * 0.1 - 30, advantage=torch.randn(n, generator=g))
    n1, n2 = make(), make()
    losses = flat_train(n1, SyncReplicasAdam(lr=1e-4, norm_clip=1.0), exp, 32, 2, np.random.RandomState(5))
    assert len(losses) == 6 and all(np.isfinite(float(l[0])) for l in losses) and n1.global_step == 6
    opt = SyncReplicasAdam(lr=1e-4, norm_clip=1.0)
    for ids in minibatch_indices(n, 32, 2, np.random.RandomState(5)):
        n2.compute_gradients(*(exp[k][ids] for k in ("state", "action", "value", "log_prob", "advantage")))
        opt.apply_gradients(n2)
    assert torch.equal(n1.params, n2.params) and torch.equal(n1.state_mean, n2.state_mean)


# ---------------------------------------------------------------------------------------------------------------------
# f3: vectorised rollout inference `run_batch` (ppo.py:56-62 -> actor_critic.py:368-380 -> a2c.py:225-243, 346-365)
# against the oracle with EXTERNAL draws: trunk + value vs oracle/network.forward, particle index bit-exact and
# action / log_prob vs MixtureGaussianOracle.sample / log_prob, activity statistics vs softmax max / sum over the batch.
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tanh", [False, True])
def test_run_batch_matches_oracle_with_external_draws(cuda_dev, tanh):
    from oracle import head as oh
    B, S, A, P = 300, 197, 36, 35
    net = build(cuda_dev, normalize_policy_output=tanh)
    g = torch.Generator().manual_seed(11)
    state = torch.randn(B, S, generator=g) * 1.5 + 0.3
    normal = torch.randn(B, A, P, generator=g)
    if tanh:
        tiny = float(np.finfo(np.float32).tiny)
        uniform = torch.rand(B, A, P, generator=g).clamp_min(tiny)
    else:
        uniform = torch.rand(B, A, generator=g, dtype=torch.float64)
    p0 = oracle_params(net)
    mean, std = net.state_mean.double().cpu(), net.state_std.double().cpu()
    max0 = torch.rand(A, P, generator=g) * 0.05
    sum0 = torch.rand(A, P, generator=g)
    net.max_active.copy_(max0)
    net.sum_active.copy_(sum0)
    action, lp, value = net.run_batch(state, ext_uniform=uniform.to(cuda_dev), ext_normal=normal.to(cuda_dev))
    # 1. trunk: logits and value against the fp64 forward
    logits_ref, v_ref = on.forward(p0, state.double(), mean, std)
    logits_gpu = net._act["logits"].detach().cpu().reshape(B, A, P)
    assert rel(logits_gpu, logits_ref.reshape(B, A, P)) < TOL
    assert rel(value, v_ref) < TOL
    # 2. sampling + log_prob on the logits the kernels saw (the fp32 logits decide the particle index)
    loc, scale = p0["global_net/actor/samples"], p0["global_net/actor/samples_std"].exp()
    dist = oh.MixtureGaussianOracle(logits_gpu.double(), loc, scale, tanh)
    if tanh:
        smp, s_ = dist.sample(1, uniform=uniform.double(), normal=normal.double())
        act_ref, lp_ref = smp[0], dist.log_prob((smp[0], s_[0]))
    else:
        act_ref = dist.sample(1, uniform=uniform.numpy(), normal=normal.double())[0]
        lp_ref = dist.log_prob(act_ref)
    assert rel(action, act_ref) < TOL
    assert rel(lp, lp_ref) < TOL
    # 3. running_update_ops (a2c.py:346-365): max_active = max(max_active, max_b probs), sum_active += sum_b probs
    probs = oh.tf_softmax(logits_gpu.double())
    assert rel(net.max_active, torch.maximum(max0.double(), probs.amax(0))) < TOL
    assert rel(net.sum_active, sum0.double() + probs.sum(0)) < TOL


def test_run_batch_particle_index_is_bit_exact(cuda_dev):
    """dis_action of the plain branch: TF Multinomial CPU semantics on the network's own fp32 logits."""
    from oracle import head as oh
    from pfpn_b200 import sampling
    B, S, A, P = 257, 197, 36, 35
    net = build(cuda_dev)
    g = torch.Generator().manual_seed(5)
    state = torch.randn(B, S, generator=g)
    u = torch.rand(B, A, generator=g, dtype=torch.float64)
    nrm = torch.randn(B, A, P, generator=g)
    net.run_batch(state, ext_uniform=u.to(cuda_dev), ext_normal=nrm.to(cuda_dev))
    logits = net._act["logits"].view(B, A, P)
    _, idx = sampling.sample_plain(logits, net.loc, net.logstd, ext_uniform=u.to(cuda_dev), ext_normal=nrm.to(cuda_dev))
    ref = oh.tf_multinomial_cpu(logits.cpu().numpy().reshape(B * A, P), u.numpy().reshape(B * A, 1)).reshape(B, A)
    assert np.array_equal(idx.cpu().numpy(), ref)


def test_ppo_train_step_with_tanh_squashed_actions_matches_oracle(cuda_dev):
    """normalize_policy_output=True under PPO: the stored action is tanh'd and log_prob applies atanh (utils.py:120-126)."""
    B = 256
    net = build(cuda_dev, normalize_policy_output=True)
    batch = make_batch(B, 197, 36, seed=77)
    batch["action"] = batch["action"] * 0.97  # strictly inside (-1, 1): atanh finite, as after a tanh squash
    p0 = oracle_params(net)
    mean, std = net.state_mean.double().cpu(), net.state_std.double().cpu()
    b64 = {k: v.double() for k, v in batch.items()}
    from oracle import head as oh
    logits, _ = on.forward(p0, b64["state"], mean, std)
    lp = oh.MixtureGaussianOracle(logits.reshape(B, 36, 35), p0["global_net/actor/samples"],
                                  p0["global_net/actor/samples_std"].exp(), True).log_prob(b64["action"])
    batch["log_prob"] = (lp + 0.1 * torch.randn(B, dtype=torch.float64)).float()
    b64["log_prob"] = batch["log_prob"].double()
    g_ref, l_ref = on.gradients(p0, b64["state"], b64["action"], b64["value"], b64["log_prob"], b64["advantage"],
                                mean, std, 36, 35, tanh=True)
    losses = net.compute_gradients(batch["state"], batch["action"], batch["value"], batch["log_prob"], batch["advantage"])
    for k, (_, g) in net.named_parameters().items():
        assert rel(g, g_ref[k]) < 2 * TOL, k  # (atanh of an fp32 action adds its own rounding on top of the head's)
    assert abs(float(losses[3]) - float(l_ref[3])) < TOL * float(l_ref[3])


# ---------------------------------------------------------------------------------------------------------------------
# a17-a19: the 3-launch optimizer step with device-resident counters (csrc/syncstep.cu) and its CUDA-graph replay
# ---------------------------------------------------------------------------------------------------------------------
def _run_steps(cuda_dev, n_steps, mode, B=192):
    import os
    from pfpn_b200.learner import GraphedUpdate
    net = build(cuda_dev)
    opt = SyncReplicasAdam(lr=1e-4, norm_clip=1.0)
    old = os.environ.get("PFPN_SYNC_STEP")
    os.environ["PFPN_SYNC_STEP"] = "0" if mode == "legacy" else "1"
    try:
        gu = GraphedUpdate(net, opt, B, warmup=2) if mode == "graph" else None
        for i in range(n_steps):
            b = make_batch(B, 197, 36, seed=100 + i)
            b["log_prob"] = b["log_prob"] - 60.0 - 30.0  # near the mixture's log-density so the ratio stays finite
            args = (b["state"].to(cuda_dev), b["action"].to(cuda_dev), b["value"].to(cuda_dev), b["log_prob"].to(cuda_dev),
                    b["advantage"].to(cuda_dev))
            if gu is not None:
                gu.run(*args)
            else:
                net.compute_gradients(*args)
                opt.apply_gradients(net)
        torch.cuda.synchronize()
    finally:
        if old is None:
            os.environ.pop("PFPN_SYNC_STEP", None)
        else:
            os.environ["PFPN_SYNC_STEP"] = old
    return net, opt, gu


def test_sync_step_matches_the_legacy_optimizer_chain(cuda_dev):
    """clip -> stage -> mean -> Adam -> statistics in 3 launches vs the round-1 chain (clip, pack, Adam, unpack) over 4
    steps: same parameters / slots / statistics up to the rounding of the norm (different but fixed summation orders)."""
    n1, o1, _ = _run_steps(cuda_dev, 4, "legacy")
    n2, o2, _ = _run_steps(cuda_dev, 4, "sync")
    assert o1._mode == "legacy" and o2._mode == "sync_step" and o2.launches_last_step == 3
    assert o1.step == o2.step == 4 and n1.global_step == n2.global_step == 4 and n1.train_flag == n2.train_flag == 4
    assert rel(o2.norm_scale, o1.norm_scale) < 1e-6
    assert rel(n2.params, n1.params) < 1e-6 and rel(o2.m, o1.m) < 1e-5 and rel(o2.v, o1.v) < 1e-5
    upd1, upd2 = n1.params - build(cuda_dev).params, n2.params - build(cuda_dev).params
    assert rel(upd2, upd1) < 1e-3  # the update itself (|dp| ~ 4e-4), not just the parameters it is added to
    assert rel(n2.state_mean, n1.state_mean) < 1e-6 and rel(n2.state_std, n1.state_std) < 1e-6
    assert torch.equal(n2.max_active, n1.max_active) and torch.equal(n2.sum_active, n1.sum_active)
    assert n2.dev_counters[:3].tolist() == [4, 4, 4]


def test_graphed_update_is_bit_identical_to_eager_steps(cuda_dev):
    """The whole update captured once and replayed (2 eager warm-up steps + 4 replays, a different minibatch each) equals
    6 eager steps bit for bit: the step number, the normaliser's decay and the exchange parity come from device memory."""
    n1, o1, _ = _run_steps(cuda_dev, 6, "sync")
    n2, o2, gu = _run_steps(cuda_dev, 6, "graph")
    assert gu.graph is not None and gu.replays == 4
    assert torch.equal(n1.params, n2.params) and torch.equal(o1.m, o2.m) and torch.equal(o1.v, o2.v)
    assert torch.equal(n1.state_mean, n2.state_mean) and torch.equal(n1.state_std, n2.state_std)
    assert o2.step == 6 and n2.global_step == 6 and n2.dev_counters[:3].tolist() == [6, 6, 6]
    assert np.isfinite(float(gu.losses[0]))
