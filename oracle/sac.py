"""ORACLE (test infrastructure, not product code): one SAC-PFPN learner step.

Restates, in torch-CPU with autograd (run it in float64 as the arbiter),
  * AbstractSACNetwork            /root/reference/networks/actor_critic/sac.py:12-173
        build_q :107-126, build_action_sampler :128-130, setup_value_target_tensor :132-139,
        build_value_loss :160-164, build_policy_loss :166-173, target sync :58-73
  * the loss assembly around it   networks/actor_critic/actor_critic.py:74-184 (value_loss *= value_loss_coef)
  * the two-optimizer update      models/workers/base_worker.py:25-120 with `separate_optimizer = True`
                                  (models/workers/ddpg.py:39-42): c_grads = d value_loss / d {vars without "actor"},
                                  a_grads = d policy_loss / d {vars without "critic"}; lr_actor == lr_critic ==> ONE joint
                                  clip_by_global_norm over both lists, then two AdamOptimizers (critic first).
Variable names follow the reference scopes: global_net/actor/*, global_net/critic/q{1,2}/fc{1,2,3}/*,
global_net/alpha/log_alpha, global_net/target_net/critic/q{1,2}/*.  The target network shares the online actor
(`build_actor_net_op = lambda: self.actor_net`, a tf.make_template) and the state normaliser (:141-143).
Parity unpinned by reference tests (none exist); pinned only through the shared, pinned head oracle (oracle/head.py).
"""
from __future__ import annotations

import math
from typing import Dict

import torch

from . import head as oh
from .network import relu6

ACTOR = "global_net/actor"
Q = "global_net/critic/q{}"
QT = "global_net/target_net/critic/q{}"
LOG_ALPHA = "global_net/alpha/log_alpha"


def normalize(state, mean, std, clip=5.0):
    x = (state - mean) / std
    if clip:
        x = torch.clamp(x, -clip, clip)
    return x.detach()  # tf.stop_gradient (actor_critic.py:78)


def actor_logits(p, x):
    h = relu6(x @ p[f"{ACTOR}/fc1/weight"] + p[f"{ACTOR}/fc1/bias"])
    h = relu6(h @ p[f"{ACTOR}/fc2/weight"] + p[f"{ACTOR}/fc2/bias"])
    return h @ p[f"{ACTOR}/fc_policy/weight"] + p[f"{ACTOR}/fc_policy/bias"]


def q_value(p, prefix, x, a):
    """build_q -> build_value on concat([x, a]) (sac.py:107-113): relu6 MLP, linear scalar output, squeezed."""
    h = torch.cat([x, a], dim=-1)
    h = relu6(h @ p[f"{prefix}/fc1/weight"] + p[f"{prefix}/fc1/bias"])
    h = relu6(h @ p[f"{prefix}/fc2/weight"] + p[f"{prefix}/fc2/bias"])
    return (h @ p[f"{prefix}/fc3/weight"] + p[f"{prefix}/fc3/bias"]).squeeze(1)


def policy(p, x, A, P, uniform, normal):
    B = x.shape[0]
    dist = oh.MixtureGaussianOracle(actor_logits(p, x).reshape(B, A, P), p[f"{ACTOR}/samples"],
                                    torch.exp(p[f"{ACTOR}/samples_std"]), True)
    a, s_ = dist.sample(1, uniform=uniform, normal=normal)
    a, s_ = a[0], s_[0]
    return a, dist.log_prob((a, s_)), dist


def losses(p, state, action_hist, reward, not_terminal, state_, mean, std, A, P, draws, *, gamma=0.95, value_loss_coef=0.5,
           clip_state=5.0):
    """draws = (U, EPS, U_next, EPS_next): the Gumbel uniforms / location normals of the two policy evaluations."""
    x, x2 = normalize(state, mean, std, clip_state), normalize(state_, mean, std, clip_state)
    log_alpha = p[LOG_ALPHA]
    alpha = torch.exp(log_alpha).detach()
    a, logp, _ = policy(p, x, A, P, draws[0], draws[1])
    q1a, q2a = q_value(p, Q.format(1), x, a), q_value(p, Q.format(2), x, a)
    q1r, q2r = q_value(p, Q.format(1), x, action_hist), q_value(p, Q.format(2), x, action_hist)
    a2, logp2, _ = policy(p, x2, A, P, draws[2], draws[3])           # the ONLINE actor evaluated at s' (shared template)
    vf = torch.minimum(q_value(p, QT.format(1), x2, a2), q_value(p, QT.format(2), x2, a2)) - alpha * logp2
    q_target = (reward + gamma * not_terminal * vf).detach()
    value_loss = value_loss_coef * torch.mean(torch.square(q_target - q1r) + torch.square(q_target - q2r))
    target_entropy = -float(A)
    policy_loss = torch.mean(alpha * logp - torch.minimum(q1a, q2a) - log_alpha * (logp + target_entropy).detach())
    return policy_loss + value_loss, policy_loss, value_loss, dict(action=a, logp=logp, q_target=q_target, alpha=alpha)


def split_vars(p):
    train = [k for k in p if "/target_net/" not in k]
    critic_vars = [k for k in train if "actor" not in k.split("/")]   # base_worker.py:51-55
    actor_vars = [k for k in train if "critic" not in k.split("/")]
    return critic_vars, actor_vars


def gradients(p, *args, **kw):
    q = {k: t.detach().clone().requires_grad_("/target_net/" not in k) for k, t in p.items()}
    loss, pl, vl, aux = losses(q, *args, **kw)
    cv, av = split_vars(q)
    cg = torch.autograd.grad(vl, [q[k] for k in cv], retain_graph=True, allow_unused=True)
    ag = torch.autograd.grad(pl, [q[k] for k in av], allow_unused=True)
    c_grads = {k: g for k, g in zip(cv, cg) if g is not None}     # log_alpha: None under value_loss
    a_grads = {k: g for k, g in zip(av, ag) if g is not None}
    return c_grads, a_grads, (loss.detach(), pl.detach(), vl.detach()), aux


def clip_joint(c_grads, a_grads, clip=1.0):
    """clip_grads, `elif self.lr_critic == self.lr_actor` branch (base_worker.py:97-102): one global norm over both lists."""
    norm = math.sqrt(sum(float(torch.sum(g.double() ** 2)) for g in list(c_grads.values()) + list(a_grads.values())))
    scale = clip * min(1.0 / norm, 1.0 / clip) if math.isfinite(norm) else float("nan")
    return {k: g * scale for k, g in c_grads.items()}, {k: g * scale for k, g in a_grads.items()}, norm


def train_step(p, slots, step, batch, mean, std, A, P, draws, *, lr_critic=1e-4, lr_actor=1e-4, clip=1.0, tau=0.005, **kw):
    """One full update: gradients, joint clip, critic Adam then actor Adam, soft target sync (sac.py:67-73).
    slots = {"c": (m, v), "a": (m, v)} dicts keyed by variable name; returns (losses, clipped grads, norm)."""
    from .network import adam_step
    cg, ag, ls, aux = gradients(p, *batch, mean, std, A, P, draws, **kw)
    cg, ag, norm = clip_joint(cg, ag, clip)
    for key, grads, lr in (("c", cg, lr_critic), ("a", ag, lr_actor)):
        m, v = slots[key]
        sub = {k: p[k] for k in grads}
        adam_step(sub, grads, m, v, step, lr=lr)
        p.update(sub)
    for i in (1, 2):
        for leaf in ("fc1/weight", "fc1/bias", "fc2/weight", "fc2/bias", "fc3/weight", "fc3/bias"):
            k, kt = f"{Q.format(i)}/{leaf}", f"{QT.format(i)}/{leaf}"
            p[kt] = (1 - tau) * p[kt] + tau * p[k]
    return ls, cg, ag, norm, aux
