"""CPU checks of the SAC oracle (oracle/sac.py): the closed forms the device kernel `pfpn_sac_losses` implements are the
gradients autograd derives from the restated losses, the two variable lists are the reference's name-based partition, and
the full step behaves (Adam on both lists, soft target update)."""
import math

import numpy as np
import torch

from oracle import head as oh
from oracle import sac as osac

S, A, P, B = 9, 3, 5, 12


def make(seed=0):
    g = torch.Generator().manual_seed(seed)
    lin = lambda k, n: (torch.randn(k, n, dtype=torch.float64, generator=g) * 0.4, torch.randn(n, dtype=torch.float64, generator=g) * 0.1)
    p = {}
    for nm, (k, n) in {"global_net/actor/fc1": (S, 8), "global_net/actor/fc2": (8, 6), "global_net/actor/fc_policy": (6, A * P)}.items():
        p[nm + "/weight"], p[nm + "/bias"] = lin(k, n)
    loc, ls = oh.init_particles(A, P, tanh=True)
    p["global_net/actor/samples"] = torch.tensor(loc, dtype=torch.float64)
    p["global_net/actor/samples_std"] = torch.tensor(ls, dtype=torch.float64)
    for pre in ("global_net/critic", "global_net/target_net/critic"):
        for i in (1, 2):
            for nm, (k, n) in {"fc1": (S + A, 8), "fc2": (8, 6), "fc3": (6, 1)}.items():
                p[f"{pre}/q{i}/{nm}/weight"], p[f"{pre}/q{i}/{nm}/bias"] = lin(k, n)
    p["global_net/alpha/log_alpha"] = torch.tensor(-0.2, dtype=torch.float64)
    batch = (torch.randn(B, S, dtype=torch.float64, generator=g), torch.rand(B, A, dtype=torch.float64, generator=g) * 1.8 - 0.9,
             torch.randn(B, dtype=torch.float64, generator=g), (torch.rand(B, generator=g) > 0.2).double(),
             torch.randn(B, S, dtype=torch.float64, generator=g))
    draws = tuple(t for _ in range(2) for t in (torch.rand(B, A, P, dtype=torch.float64, generator=g).clamp_min(1e-30),
                                                 torch.randn(B, A, P, dtype=torch.float64, generator=g)))
    return p, batch, draws, torch.zeros(S, dtype=torch.float64), torch.ones(S, dtype=torch.float64)


def test_variable_partition_follows_the_reference_name_rule():
    p, *_ = make()
    cv, av = osac.split_vars(p)
    assert all("/critic/q" in k or k.endswith("log_alpha") for k in cv) and all("target_net" not in k for k in cv + av)
    assert all("/actor/" in k or k.endswith("log_alpha") for k in av)
    assert "global_net/alpha/log_alpha" in cv and "global_net/alpha/log_alpha" in av  # in both lists; None grad under value_loss


def test_closed_form_scalar_gradients_equal_autograd():
    """What pfpn_sac_losses writes: d value_loss/d q_r, d policy_loss/d {q_a, logp, log_alpha}."""
    p, batch, draws, mean, std = make(1)
    q = {k: v.clone().requires_grad_("target_net" not in k) for k, v in p.items()}
    state, a_hist, r, nt, state_ = batch
    x, x2 = osac.normalize(state, mean, std), osac.normalize(state_, mean, std)
    a, logp, _ = osac.policy(q, x, A, P, draws[0], draws[1])
    a2, logp2, _ = osac.policy(q, x2, A, P, draws[2], draws[3])
    qa = [osac.q_value(q, osac.Q.format(i), x, a.detach()).detach().requires_grad_(True) for i in (1, 2)]
    qr = [osac.q_value(q, osac.Q.format(i), x, a_hist).detach().requires_grad_(True) for i in (1, 2)]
    qt = [osac.q_value(q, osac.QT.format(i), x2, a2).detach() for i in (1, 2)]
    lp = logp.detach().requires_grad_(True)
    la = q[osac.LOG_ALPHA].detach().clone().requires_grad_(True)
    alpha, gamma, coef, te = torch.exp(la).detach(), 0.95, 0.5, -float(A)
    q_target = r + gamma * nt * (torch.minimum(qt[0], qt[1]) - alpha * logp2.detach())
    vl = coef * torch.mean((q_target - qr[0]) ** 2 + (q_target - qr[1]) ** 2)
    pl = torch.mean(alpha * lp - torch.minimum(qa[0], qa[1]) - la * (lp.detach() + te))
    vl.backward()
    pl.backward()
    first = (qa[0] <= qa[1]).double()
    assert torch.allclose(qr[0].grad, -2 * coef * (q_target - qr[0].detach()) / B) and torch.allclose(qr[1].grad, -2 * coef * (q_target - qr[1].detach()) / B)
    assert torch.allclose(qa[0].grad, -first / B) and torch.allclose(qa[1].grad, -(1 - first) / B)
    assert torch.allclose(lp.grad, torch.full((B,), float(alpha) / B, dtype=torch.float64))
    assert math.isclose(float(la.grad), -float(torch.mean(lp.detach() + te)), rel_tol=1e-12)
    # and they agree with the assembled losses of the oracle
    _, pl_ref, vl_ref, _ = osac.losses(p, *batch, mean, std, A, P, draws)
    assert math.isclose(float(pl.detach()), float(pl_ref), rel_tol=1e-12) and math.isclose(float(vl.detach()), float(vl_ref), rel_tol=1e-12)


def test_full_step_updates_both_lists_and_the_target():
    p, batch, draws, mean, std = make(2)
    before = {k: v.clone() for k, v in p.items()}
    slots = {key: ({k: torch.zeros_like(v) for k, v in p.items()}, {k: torch.zeros_like(v) for k, v in p.items()}) for key in "ca"}
    (loss, pl, vl), cg, ag, norm, _ = osac.train_step(p, slots, 1, batch, mean, std, A, P, draws, tau=0.005)
    assert norm > 0 and math.isclose(float(loss), float(pl) + float(vl), rel_tol=1e-12)
    clipped = math.sqrt(sum(float((g ** 2).sum()) for g in list(cg.values()) + list(ag.values())))
    assert clipped <= 1.0 + 1e-9  # one joint clip over both gradient lists
    for k in p:
        if "target_net" in k:
            online = k.replace("target_net/", "")
            assert torch.allclose(p[k], 0.995 * before[k] + 0.005 * p[online])
        else:
            assert not torch.equal(p[k], before[k]), k   # every trainable variable moved (first Adam step: |delta| ~ lr)
            assert float((p[k] - before[k]).abs().max()) <= 1.0001e-4
