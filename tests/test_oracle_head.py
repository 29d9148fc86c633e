"""CPU checks that pin the oracle's head restatement with the closed-form invariants
of SURVEY.md section 8c(3) / Appendix A: finite differences, quadrature, bounds, guard."""
import math

import numpy as np
import torch

from oracle import head as oh
from pfpn_b200 import synth


def _inputs(B=6, A=3, P=7, seed=0):
    d = synth.head_inputs(B, A, P, seed=seed, far_frac=0.0)
    return {k: (v.double() if torch.is_tensor(v) else v) for k, v in d.items()}


def test_entropy_bounds_and_uniform_case():
    d = _inputs()
    dist = oh.MixtureGaussianOracle(d["logits"], d["loc"], d["logstd"].exp(), False)
    H = dist.entropy()
    assert H.shape == (6, 3) and (H >= 0).all() and (H <= math.log(7) + 1e-12).all()
    uni = oh.MixtureGaussianOracle(torch.zeros(2, 3, 7, dtype=torch.float64), d["loc"], d["logstd"].exp(), False)
    assert torch.allclose(uni.entropy(), torch.full((2, 3), math.log(7), dtype=torch.float64))


def test_log_prob_integrates_to_one_per_dimension():
    d = _inputs(B=1, A=2, P=9)
    dist = oh.MixtureGaussianOracle(d["logits"], d["loc"], d["logstd"].exp(), False)
    xs = torch.linspace(-4, 4, 20001, dtype=torch.float64)
    for a in range(2):
        sub = oh.MixtureGaussianOracle(d["logits"][:, a:a + 1].expand(len(xs), 1, 9), d["loc"][a:a + 1],
                                       d["logstd"][a:a + 1].exp(), False)
        pdf = sub.prob(xs[:, None])
        assert abs(float(torch.trapezoid(pdf, xs)) - 1.0) < 1e-6
    # tanh variant: density of t = tanh(u) on (-1, 1)
    ts = torch.linspace(-1 + 1e-6, 1 - 1e-6, 200001, dtype=torch.float64)
    sub = oh.MixtureGaussianOracle(d["logits"][:, :1].expand(len(ts), 1, 9), d["loc"][:1] * 0.5,
                                   d["logstd"][:1].exp(), True)
    pdf = sub.prob(ts[:, None])
    assert abs(float(torch.trapezoid(pdf, ts)) - 1.0) < 1e-3


def test_closed_form_gradients_match_autograd():
    """Appendix A1: dL/dl_k = g (r_k - pi_k); dmu = sum g r z / sigma; dlogstd = sum g r (z^2 - 1)."""
    d = _inputs(B=5, A=4, P=6, seed=3)
    g = torch.randn(5, dtype=torch.float64, generator=torch.Generator().manual_seed(1))
    ref = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], g, want_dvalue=True)
    pi = torch.softmax(d["logits"], -1)
    sig = d["logstd"].exp()
    z = (d["value"][..., None] - d["loc"]) / sig
    n = torch.exp(-0.5 * z ** 2 - (oh.HALF_LOG_2PI + d["logstd"]))
    p = (pi * n).sum(-1, keepdim=True)
    r = pi * n / p
    gb = g[:, None, None]
    assert torch.allclose(ref["dlogits"], gb * (r - pi), atol=1e-13)
    assert torch.allclose(ref["dloc"], (gb * r * z / sig).sum(0), atol=1e-12)
    assert torch.allclose(ref["dlogstd"], (gb * r * (z ** 2 - 1)).sum(0), atol=1e-12)
    assert torch.allclose(ref["dvalue"], -(gb * r * z / sig).sum(-1), atol=1e-12)


def test_finite_differences():
    d = _inputs(B=3, A=2, P=5, seed=4)
    g = torch.tensor([0.3, -1.2, 0.7], dtype=torch.float64)
    for tanh in (False, True):
        ref = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], g, torch.full((3,), 0.2),
                              tanh=tanh, want_dvalue=True)

        def L(lg, lc, ls, v):
            dist = oh.MixtureGaussianOracle(lg, lc, ls.exp(), tanh)
            lp = dist.log_prob((torch.tanh(v), v) if tanh else v)
            return float((g * lp).sum() + 0.2 * dist.entropy().sum())

        eps = 1e-6
        for name, key, idx in (("logits", "dlogits", (1, 1, 2)), ("loc", "dloc", (1, 3)),
                               ("logstd", "dlogstd", (0, 4)), ("value", "dvalue", (2, 1))):
            args = {k: d[k].clone() for k in ("logits", "loc", "logstd", "value")}
            args[name][idx] += eps
            up = L(args["logits"], args["loc"], args["logstd"], args["value"])
            args[name][idx] -= 2 * eps
            dn = L(args["logits"], args["loc"], args["logstd"], args["value"])
            assert abs((up - dn) / (2 * eps) - float(ref[key][idx])) < 1e-6, (name, tanh)


def test_guard_zeroes_gradient_when_p_underflows():
    d = _inputs(B=4, A=2, P=5, seed=5)
    d["value"][1, 0] = 60.0
    for dt in (torch.float32, torch.float64):
        ref = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], torch.ones(4), dtype=dt)
        assert torch.isinf(ref["lp"][1]) and torch.isfinite(ref["lp"][[0, 2, 3]]).all()
        assert torch.count_nonzero(ref["dlogits"][1, 0]) == 0 and torch.isfinite(ref["dlogits"]).all()
        assert torch.count_nonzero(ref["dlogits"][1, 1]) > 0  # the other dimension still learns


def test_ppo_gradient_matches_closed_form_with_tie_rule():
    d = _inputs(B=64, A=3, P=7, seed=6)
    ref0 = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], torch.zeros(64))
    lp_old = ref0["lp"] + d["lp_noise"] * 6
    out = oh.ppo_head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], d["adv"], lp_old)
    an = oh.normalize_advantage(d["adv"])
    ratio = torch.exp(ref0["lp"] - lp_old)
    surr, clipped = ratio * an, ratio.clamp(0.8, 1.2) * an
    gb = torch.where(surr <= clipped, -ratio * an / 64, torch.zeros_like(an))
    chk = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], gb)
    assert (surr > clipped).any() and (surr <= clipped).any()
    assert torch.allclose(out["dlogits"], chk["dlogits"], atol=1e-14)
    assert torch.allclose(out["dloc"], chk["dloc"], atol=1e-13)


def test_fp32_op_order_run_is_within_tolerance_of_fp64():
    """norm-wise 1e-5 is attainable by an fp32 replay of the reference op order."""
    d = synth.head_inputs(512, 36, 35, seed=34114, far_frac=0.0)
    g = torch.randn(512, generator=torch.Generator().manual_seed(0))
    r64 = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], g, dtype=torch.float64)
    r32 = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], g, dtype=torch.float32)
    for k in ("lp", "dlogits", "dloc", "dlogstd"):
        err = float((r32[k].double() - r64[k]).abs().max() / r64[k].abs().max())
        assert err < 1e-5, (k, err)


def test_particle_grid_init():
    loc, logstd = oh.init_particles(3, 35, tanh=False)
    assert np.allclose(loc[0], np.linspace(-1, 1, 35)) and np.allclose(logstd, math.log(2 / 34))
    loc, logstd = oh.init_particles(2, 10, tanh=True)
    c = -1 + (2 / 10) * (np.arange(10) + 0.5)
    assert np.allclose(np.tanh(loc[0]), c)
    mu = np.arctanh(c)
    assert np.isclose(np.exp(logstd[0, 0]), mu[1] - mu[0]) and np.isclose(np.exp(logstd[0, 4]), max(mu[4] - mu[3], mu[5] - mu[4]))
