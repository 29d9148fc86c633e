"""Restatement of the DPPO learner update around the PFPN head.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/networks/actor_critic/actor_critic.py:74-184,223-244 (state normaliser,
loss assembly), networks/ops.py:82-118 + networks/utils.py:17-43,60-68 (trunk, moving-average
normaliser), ppo.py:39-54 (losses), models/workers/base_worker.py:25-120 (gradients, clip,
Adam) and models/sync_model.py:60-101 (aggregation order), in torch-CPU with autograd.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch

from . import head as oh


def relu6(x):
    return torch.clamp(x, 0.0, 6.0)


def forward(p: Dict[str, torch.Tensor], state, mean, std, clip=5.0, normalize=True):
    """p: reference variable names -> tensors.  Returns (logits [B,A*P], value [B])."""
    x = state
    if normalize:
        x = (x - mean) / std
    if clip:
        x = torch.clamp(x, -clip, clip)
    x = x.detach()  # tf.stop_gradient (actor_critic.py:78)
    h = relu6(x @ p["global_net/actor/fc1/weight"] + p["global_net/actor/fc1/bias"])
    h = relu6(h @ p["global_net/actor/fc2/weight"] + p["global_net/actor/fc2/bias"])
    logits = h @ p["global_net/actor/fc_policy/weight"] + p["global_net/actor/fc_policy/bias"]
    c = relu6(x @ p["global_net/critic/fc1/weight"] + p["global_net/critic/fc1/bias"])
    c = relu6(c @ p["global_net/critic/fc2/weight"] + p["global_net/critic/fc2/bias"])
    v = (c @ p["global_net/critic/fc3/weight"] + p["global_net/critic/fc3/bias"]).squeeze(1)
    return logits, v


def ppo_losses(p, state, action, value_old, lp_old, adv, mean, std, A, P, *, eps=0.2, normalize_adv=True,
               value_loss_coef=0.5, entropy_beta=None, tanh=False, clip_state=5.0, normalize_state=True):
    logits, v = forward(p, state, mean, std, clip_state, normalize_state)
    B = state.shape[0]
    dist = oh.MixtureGaussianOracle(logits.reshape(B, A, P), p["global_net/actor/samples"],
                                    torch.exp(p["global_net/actor/samples_std"]), tanh)
    # `action` is the stored action_hist; with normalize_output the reference's log_prob receives it as a plain
    # tensor and applies atanh itself (utils.py:120-126)
    lp = dist.log_prob(action)
    adv_n = oh.normalize_advantage(adv).detach() if normalize_adv else adv
    policy_loss = oh.ppo_policy_loss(lp, lp_old, adv_n, eps)
    entropy = None
    if entropy_beta:
        entropy = torch.mean(torch.sum(dist.entropy(), dim=1))
        policy_loss = policy_loss - entropy_beta * entropy
    value_loss = torch.mean(torch.square(v - (adv + value_old).detach()))  # ppo.py:31-42
    value_loss = value_loss_coef * value_loss  # actor_critic.py:131-133: self.value_loss *= value_loss_coef (what train() returns)
    loss = policy_loss + value_loss
    return loss, entropy, policy_loss, value_loss


def gradients(p, *args, **kw):
    q = {k: t.detach().clone().requires_grad_(True) for k, t in p.items()}
    losses = ppo_losses(q, *args, **kw)
    losses[0].backward()
    return {k: t.grad for k, t in q.items()}, tuple(None if x is None else x.detach() for x in losses)


def clip_by_global_norm(grads: Dict[str, torch.Tensor], clip: float):
    """TF: scale = clip * min(1/norm, 1/clip) ([graph] optimizer/clip_by_global_norm)."""
    norm = torch.sqrt(sum(torch.sum(g.double() ** 2) for g in grads.values()))
    scale = clip * min(1.0 / float(norm), 1.0 / clip) if math.isfinite(float(norm)) else float("nan")
    return {k: g * scale for k, g in grads.items()}, float(norm)


def adam_step(p, g, m, v, step, lr=1e-4, b1=0.9, b2=0.999, eps=1e-8):
    """TF AdamOptimizer (epsilon outside the bias correction)."""
    lr_t = lr * math.sqrt(1 - b2 ** step) / (1 - b1 ** step)
    for k in p:
        m[k] = b1 * m[k] + (1 - b1) * g[k]
        v[k] = b2 * v[k] + (1 - b2) * g[k] * g[k]
        p[k] = p[k] - lr_t * m[k] / (torch.sqrt(v[k]) + eps)


def normalizer_update(mean, std, X, step):
    """networks/utils.py:60-68."""
    decay = min(0.9999, (1 + step) / (10 + step))
    m = X.mean(0)
    var = ((X - m) ** 2).mean(0)
    return decay * mean + (1 - decay) * m, torch.clamp(decay * std + (1 - decay) * torch.sqrt(var), min=1e-6)


def sync_update(p, m, v, step, shards: List[dict], mean, std, A, P, global_step=0, clip=1.0, **kw):
    """One SyncReplicasOptimizer step over len(shards) workers (sync_model.py:60-101): local
    gradients, local clip, mean over workers (gradients and pushed state statistics), Adam."""
    acc = {k: torch.zeros_like(t) for k, t in p.items()}
    new_mean, new_std, losses = torch.zeros_like(mean), torch.zeros_like(std), []
    for s in shards:
        g, l = gradients(p, s["state"], s["action"], s["value"], s["log_prob"], s["advantage"], mean, std, A, P, **kw)
        g, _ = clip_by_global_norm(g, clip)
        for k in acc:
            acc[k] += g[k] / len(shards)
        nm, ns = normalizer_update(mean, std, s["state"], global_step)
        new_mean += nm / len(shards)
        new_std += ns / len(shards)
        losses.append(l)
    adam_step(p, acc, m, v, step)
    return acc, new_mean, new_std, losses
