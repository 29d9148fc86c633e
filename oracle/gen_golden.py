"""Generates tests/golden/reference_golden.npz by EXECUTING THE REFERENCE'S OWN SOURCE
(/root/reference/networks/...) on the eager TF-1 shim (oracle/tf_shim.py), in fp64.

    python -m oracle.gen_golden            # needs /root/reference (this container only)

The committed .npz is what the CPU and GPU test-suites load; /root/reference is never read at
test time.  Every case stores the inputs, the injected random draws and the reference outputs.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import tf_shim as S  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
G: dict = {}


def put(case, **arrs):
    for k, v in arrs.items():
        if isinstance(v, S.T):
            v = v.t
        if isinstance(v, torch.Tensor):
            v = v.detach().numpy()
        G[f"{case}/{k}"] = np.asarray(v)


def var(a):
    return S.T(torch.tensor(np.asarray(a), dtype=torch.float64, requires_grad=True))


def mixture_inputs(rng, B, A, P, tanh):
    logits = rng.normal(0, 2, (B, A, P))
    if tanh:
        c = -1 + (2 / P) * (np.arange(P) + 0.5)
        loc = np.arctanh(c)[None].repeat(A, 0) + rng.normal(0, .02, (A, P))
        logstd = np.log(np.gradient(np.arctanh(c)))[None].repeat(A, 0) + rng.normal(0, .1, (A, P))
    else:
        loc = np.linspace(-1, 1, P)[None].repeat(A, 0) + rng.normal(0, .02, (A, P))
        logstd = np.log(2 / (P - 1)) + rng.normal(0, .1, (A, P))
    k = rng.integers(0, P, (B, A))
    value = np.take_along_axis(loc[None].repeat(B, 0), k[..., None], 2)[..., 0] + \
        np.exp(np.take_along_axis(logstd[None].repeat(B, 0), k[..., None], 2)[..., 0]) * rng.normal(0, 1, (B, A))
    return logits, loc, logstd, value


def case_logprob_entropy(mods, rng, name, tanh, guard=False):
    B, A, P = 8, 4, 7
    logits, loc, logstd, value = mixture_inputs(rng, B, A, P, tanh)
    if guard:
        value[2, 1] = 80.0
    g_lp, g_ent = rng.normal(0, 1, B), rng.normal(0, .1, (B, A))
    lg, lc, ls, v = var(logits), var(loc), var(logstd), var(value)
    dist = mods["utils"].MixtureGaussianDistribution(lg, lc, S.exp(ls), tanh)
    lp = dist.log_prob((S.tanh(v), v) if tanh else v)
    ent = dist.entropy()
    pr = dist.prob((S.tanh(v), v) if tanh else v)
    L = torch.sum(torch.as_tensor(g_lp) * lp.t) + torch.sum(torch.as_tensor(g_ent) * ent.t)
    L.backward()
    put(name, logits=logits, loc=loc, logstd=logstd, value=value, g_lp=g_lp, g_ent=g_ent, tanh=int(tanh),
        lp=lp, ent=ent, prob=pr, probs=dist.dis_dist.probs, dlogits=lg.t.grad, dloc=lc.t.grad, dlogstd=ls.t.grad,
        dvalue=v.t.grad)


def make_net(cls, **attrs):
    net = object.__new__(cls)
    net.valid_data_mask = lambda x: x
    for k, v in attrs.items():
        setattr(net, k, v)
    return net


def case_ppo(mods, rng):
    B, A, P = 16, 4, 7
    logits, loc, logstd, value = mixture_inputs(rng, B, A, P, False)
    adv = rng.normal(0, 1, B)
    lg, lc, ls = var(logits), var(loc), var(logstd)
    dist = mods["utils"].MixtureGaussianDistribution(lg, lc, S.exp(ls), False)
    lp0 = dist.log_prob(S.T(torch.tensor(value))).t.detach().numpy()
    lp_old = lp0 + rng.normal(0, .3, B)  # wide enough to hit both clip sides
    # advantage normalisation exactly as actor_critic.py:151-155 builds it
    a = S.T(torch.tensor(adv))
    mean, v = S.moments(a, axes=[0])
    adv_n = S.stop_gradient((a - mean) / (S.sqrt(v) + 1e-8))
    net = make_net(mods["ppo"].ParticleFilteringClipPPONetwork, epsilon=0.2, normalize_policy_output=False,
                   normalize_policy_output_=False, running_log_prob=S.T(torch.tensor(lp_old)))
    loss = net.build_policy_loss(dist, S.T(torch.tensor(value)), adv_n)
    loss.t.backward()
    put("ppo", logits=logits, loc=loc, logstd=logstd, value=value, adv=adv, lp_old=lp_old, adv_n=adv_n, loss=loss,
        dlogits=lg.t.grad, dloc=lc.t.grad, dlogstd=ls.t.grad)
    # value loss of ppo.py:39-42 for completeness of the learner-update oracle
    val, vold = rng.normal(0, 1, B), rng.normal(0, 1, B)
    net.advantage, net.running_value = a, S.T(torch.tensor(vold))
    vt = net.setup_value_target_tensor()
    vl = net.build_value_loss(S.T(torch.tensor(val)), vt)
    put("ppo_value", value_pred=val, value_old=vold, adv=adv, value_loss=vl)


def case_sample_plain(mods, rng):
    B, A, P = 32, 4, 7
    logits, loc, logstd, _ = mixture_inputs(rng, B, A, P, False)
    logits[0, 0, :2] = -np.inf
    u, eps = rng.random(B * A), rng.normal(0, 1, (B, A, P))
    S.DRAWS.update(cat_uniform=[u], normal=[eps])
    dist = mods["utils"].MixtureGaussianDistribution(S.T(torch.tensor(logits)), S.T(torch.tensor(loc)),
                                                     S.T(torch.tensor(np.exp(logstd))), False)
    smp = dist.sample(1)
    put("sample_plain", logits=logits, loc=loc, logstd=logstd, uniform=u.reshape(B, A), normal=eps, sample=smp,
        dis_action=dist.dis_action)
    # the glue of a2c.py:225-243 / ppo.py:27-29: (action, log_prob) as the rollout sees them
    S.DRAWS.update(cat_uniform=[u], normal=[eps])
    net = make_net(mods["ppo"].ParticleFilteringClipPPONetwork, normalize_policy_output=False, normalize_policy_output_=False)
    act = net.build_action_sampler(True, dist)
    put("sample_plain", action=act, action_log_prob=net.action_log_prob)


def case_rsample_sac(mods, rng):
    B, A, P = 16, 4, 10
    logits, loc, logstd, _ = mixture_inputs(rng, B, A, P, True)
    U = np.maximum(rng.random((1, B, A, P)), np.finfo(np.float32).tiny)
    eps = rng.normal(0, 1, (B, A, P))
    g_a, g_u = rng.normal(0, 1, (B, A)), rng.normal(0, .3, (B, A))
    lg, lc, ls = var(logits), var(loc), var(logstd)
    S.DRAWS.update(gumbel_uniform=[U], normal=[eps])
    dist = mods["utils"].MixtureGaussianDistribution(lg, lc, S.exp(ls), True)
    smp, s_ = dist.sample(1)
    (torch.sum(torch.as_tensor(g_a) * smp.t[0]) + torch.sum(torch.as_tensor(g_u) * s_.t[0])).backward()
    put("rsample", logits=logits, loc=loc, logstd=logstd, uniform=U[0], normal=eps, g_sample=g_a, g_s_pre=g_u,
        sample=smp, s_pre=s_, dis_action=dist.dis_action, dlogits=lg.t.grad, dloc=lc.t.grad, dlogstd=ls.t.grad)
    # SAC policy loss through the reference glue: a2c.py:225-243 + sac.py:128-130,166-173
    lg, lc, ls = var(logits), var(loc), var(logstd)
    S.DRAWS.update(gumbel_uniform=[U], normal=[eps])
    dist = mods["utils"].MixtureGaussianDistribution(lg, lc, S.exp(ls), True)
    qw1, qw2 = rng.normal(0, 1, A), rng.normal(0, 1, A)
    log_alpha = var(np.float64(-0.3))
    net = make_net(mods["sac"].ParticleFilteringSACNetwork, normalize_policy_output=False, normalize_policy_output_=True,
                   log_alpha=log_alpha, alpha=S.stop_gradient(S.exp(log_alpha)))
    action = net.build_action_sampler(True, dist)  # sets net.target_log_prob
    net.q1_a_target = S.reduce_sum(action * S.T(torch.tensor(qw1)), axis=1)
    net.q2_a_target = S.reduce_sum(action * S.T(torch.tensor(qw2)), axis=1)
    loss = net.build_policy_loss(dist, S.T(torch.zeros(B, A, dtype=torch.float64)), None)
    loss.t.backward()
    put("sac", qw1=qw1, qw2=qw2, log_alpha=-0.3, action=action, target_log_prob=net.target_log_prob, loss=loss,
        dlogits=lg.t.grad, dloc=lc.t.grad, dlogstd=ls.t.grad, dlog_alpha=log_alpha.t.grad)


def case_mean(mods, rng):
    B, A, P = 8, 4, 7
    for tanh in (False, True):
        logits, loc, logstd, _ = mixture_inputs(rng, B, A, P, tanh)
        dist = mods["utils"].MixtureGaussianDistribution(S.T(torch.tensor(logits)), S.T(torch.tensor(loc)),
                                                         S.T(torch.tensor(np.exp(logstd))), tanh)
        put(f"mean_tanh{int(tanh)}", logits=logits, loc=loc, logstd=logstd, mean=dist.mean())


def case_build_policy(mods, rng):
    """a2c.py:476-559: particle grid, `samples` / `samples_std` variables, fc_policy, reshape."""
    for tanh, A, P, H in ((False, 3, 35, 6), (True, 2, 10, 5)):
        S.VARIABLES.clear()
        w0 = rng.normal(0, .01, (H, A * P))
        S.DRAWS.update(weight_init=[w0])
        h = rng.normal(0, 1, (4, H))
        net = make_net(mods["a2c"].ParticleFilteringA2CNetwork, identical_n_particles=True, dis_action_shape=[P] * A,
                       action_upper_bound=[3.0] * A, action_lower_bound=[-2.0] * A, normalize_policy_output=False,
                       normalize_policy_output_=tanh, init_sigma=None, trainable=True, fixed_sigma=False,
                       weight_initializer=lambda: S.truncated_normal_initializer(0.0, 0.01))
        dist = net.build_policy(True, S.T(torch.tensor(h)), [A])
        names = [v.name for v in S.VARIABLES]
        assert names == ["samples:0", "samples_std:0", "fc_policy/weight:0", "fc_policy/bias:0"], names
        put(f"build_policy_tanh{int(tanh)}", h=h, weight=w0, loc=dist.loc, scale=dist.scale, logstd=net.logstd,
            logits=dist.logits, bias=net.policy_bias, normalize_output=int(dist.normalize_output))


def case_resample(mods, rng):
    A, P, H = 4, 10, 8
    cases = {
        "none": dict(dead=[], mode=-1, tanh=False),
        "few": dict(dead=[(0, 2), (0, 7), (2, 0), (3, 9)], mode=-1, tanh=False),
        "many_tanh": dict(dead=[(a, j) for a in range(A) for j in range(P) if (a + j) % 2], mode=-1, tanh=True),
        "all_but_one": dict(dead=[(a, j) for a in range(A) for j in range(P) if j != a + 1], mode=-1, tanh=False),
        "topk": dict(dead=[(0, 1), (1, 3), (1, 4), (3, 8)], mode=3, tanh=False),
        "dead_source": dict(dead=[(1, 2), (1, 6)], mode=-1, tanh=False, force_row=(1, 2)),
    }
    for name, c in cases.items():
        lgt = rng.normal(0, 3, (64, A, P))
        pr = np.exp(lgt - lgt.max(-1, keepdims=True))
        pr /= pr.sum(-1, keepdims=True)
        max_active, sum_active = pr.max(0), pr.sum(0)
        for (a, j) in c["dead"]:
            max_active[a, j] = 1e-6
            sum_active[a, j] *= 1e-3
        if "force_row" in c:  # all categorical mass on one (dead) particle
            a, j = c["force_row"]
            sum_active[a] = 1e-12
            sum_active[a, j] = 1.0
        # the reference graph is fp32: store fp32-representable inputs so oracle/GPU see the same bits
        f32 = lambda x: np.asarray(x, np.float32).astype(np.float64)
        max_active, sum_active = f32(max_active), f32(sum_active)
        loc = f32(np.linspace(-1, 1, P)[None].repeat(A, 0) * (0.9 if c["tanh"] else 1) + rng.normal(0, .02, (A, P)))
        logstd = f32(np.log(2 / (P - 1)) + rng.normal(0, .1, (A, P)))
        bias, W = f32(rng.normal(0, 1, A * P)), f32(rng.normal(0, .01, (H, A * P)))
        cat_u, noise_u = rng.random((A, P)), f32(rng.random(A * P) * 2 - 1)
        choice = rng.integers(0, 3, A * P).astype(np.int32)
        S.set_dtype(torch.float32)  # run the resampler in the reference's own precision
        try:
            t32 = lambda x: S.T(torch.tensor(np.asarray(x), dtype=torch.float32))
            vloc, vlogstd, vbias, vW = t32(loc), t32(logstd), t32(bias), t32(W)
            S.DRAWS.update(resample_cat_uniform=[cat_u], float_uniform=[noise_u], int_uniform=[choice])
            S.TRACE.clear()
            net = make_net(mods["a2c"].ParticleFilteringA2CNetwork, identical_n_particles=True, dis_action_shape=[P] * A,
                           resample=c["mode"], resample_threshold=None, max_active=t32(max_active),
                           sum_active=t32(sum_active), policy_weight=vW, fixed_sigma=False,
                           normalize_policy_output=False, normalize_policy_output_=c["tanh"], init_sigma=None)
            policy = types.SimpleNamespace(loc=vloc, scale=S.exp(vlogstd))
            net.build_resample_ops(policy, vlogstd, vbias, vW)
        finally:
            S.set_dtype(torch.float64)
        tr = S.TRACE
        inv = tr["where"].numpy().astype(np.int32)
        cand = (tr["categorical"] if c["mode"] < 0 else tr["top_k"]).numpy().astype(np.int32)
        put(f"resample_{name}", max_active=max_active, sum_active=sum_active, loc=loc, logstd=logstd, bias=bias, weight=W,
            cat_u=cat_u, noise_u=noise_u, choice=choice, mode=c["mode"], tanh=int(c["tanh"]),
            out_loc=vloc, out_logstd=vlogstd, out_bias=vbias, out_weight=vW, invalid=inv, cand=cand,
            tcol=tr["unique_in"].numpy().astype(np.int32), uniq=tr["unique_out"][0].numpy().astype(np.int32),
            idx=tr["unique_out"][1].numpy(), count=tr["unique_out"][2].numpy(),
            delta=tr["map_fn"].numpy().astype(np.int32) if inv.shape[0] else np.zeros(0, np.int32))


def main():
    mods = S.import_reference()
    rng = np.random.default_rng(34114)
    case_logprob_entropy(mods, rng, "logprob_plain", False)
    case_logprob_entropy(mods, rng, "logprob_tanh", True)
    case_logprob_entropy(mods, rng, "logprob_guard", False, guard=True)
    case_ppo(mods, rng)
    case_sample_plain(mods, rng)
    case_rsample_sac(mods, rng)
    case_mean(mods, rng)
    case_build_policy(mods, rng)
    case_resample(mods, rng)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **G)
    print(f"wrote {OUT}: {len(G)} arrays, {os.path.getsize(OUT) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
