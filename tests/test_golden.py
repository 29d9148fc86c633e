"""Pins the oracle (CPU) and the CUDA kernels (GPU) to tests/golden/reference_golden.npz, which was
produced by executing the reference's OWN source files on the eager TF-1 shim
(oracle/gen_golden.py; /root/reference is not needed at test time)."""
import os

import numpy as np
import pytest
import torch

from oracle import head as oh
from oracle import resample as orr

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden.npz"))


def case(name):
    pre = name + "/"
    return {k[len(pre):]: GOLD[k] for k in GOLD.files if k.startswith(pre)}


def t64(x):
    return torch.tensor(np.asarray(x), dtype=torch.float64)


def close(a, b, tol=1e-10):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    if fin.any():
        assert np.abs(a[fin] - b[fin]).max() <= tol * max(1.0, np.abs(b[fin]).max()), np.abs(a[fin] - b[fin]).max()


RESAMPLE_CASES = ["none", "few", "many_tanh", "all_but_one", "topk", "dead_source"]

# ------------------------------------------------------------------------------- CPU: oracle ----


@pytest.mark.parametrize("name", ["logprob_plain", "logprob_tanh", "logprob_guard"])
def test_oracle_logprob_entropy_matches_reference_source(name):
    c = case(name)
    tanh = bool(c["tanh"])
    lg, lc, ls, v = (t64(c[k]).requires_grad_(True) for k in ("logits", "loc", "logstd", "value"))
    dist = oh.MixtureGaussianOracle(lg, lc, ls.exp(), tanh)
    lp = dist.log_prob((torch.tanh(v), v) if tanh else v)
    ent = dist.entropy()
    (torch.sum(t64(c["g_lp"]) * lp) + torch.sum(t64(c["g_ent"]) * ent)).backward()
    close(lp.detach(), c["lp"]); close(ent.detach(), c["ent"]); close(dist.probs.detach(), c["probs"])
    close(lg.grad, c["dlogits"]); close(lc.grad, c["dloc"]); close(ls.grad, c["dlogstd"]); close(v.grad, c["dvalue"])
    if name == "logprob_guard":
        # lp = -inf; the guarded row only carries the entropy gradient (which bypasses `foo`)
        assert np.isinf(c["lp"][2]) and np.isfinite(c["dlogits"]).all() and np.isfinite(c["dvalue"]).all()
        assert c["dvalue"][2, 1] == 0.0


def test_oracle_ppo_loss_matches_reference_source():
    c = case("ppo")
    out = oh.ppo_head_fwd_bwd(t64(c["logits"]), t64(c["loc"]), t64(c["logstd"]), t64(c["value"]), t64(c["adv"]),
                              t64(c["lp_old"]), eps=0.2, normalize_adv=True)
    close(oh.normalize_advantage(t64(c["adv"])), c["adv_n"])
    close(out["loss"], c["loss"]); close(out["dlogits"], c["dlogits"]); close(out["dloc"], c["dloc"])
    close(out["dlogstd"], c["dlogstd"])
    v = case("ppo_value")
    vl = torch.mean(torch.square(t64(v["value_pred"]) - (t64(v["adv"]) + t64(v["value_old"]))))
    close(vl, v["value_loss"])


def test_oracle_sampling_matches_reference_source():
    c = case("sample_plain")
    dist = oh.MixtureGaussianOracle(t64(c["logits"]), t64(c["loc"]), t64(c["logstd"]).exp(), False)
    s = dist.sample(1, uniform=c["uniform"], normal=t64(c["normal"]))
    assert np.array_equal(dist.dis_action.numpy(), c["dis_action"])
    close(s, c["sample"]); close(s[0], c["action"]); close(dist.log_prob(s[0]), c["action_log_prob"])
    c = case("rsample")
    lg, lc, ls = (t64(c[k]).requires_grad_(True) for k in ("logits", "loc", "logstd"))
    dist = oh.MixtureGaussianOracle(lg, lc, ls.exp(), True)
    smp, s_ = dist.sample(1, uniform=t64(c["uniform"]), normal=t64(c["normal"]))
    (torch.sum(t64(c["g_sample"]) * smp[0]) + torch.sum(t64(c["g_s_pre"]) * s_[0])).backward()
    assert np.array_equal(dist.dis_action.numpy(), c["dis_action"])
    close(smp.detach(), c["sample"]); close(s_.detach(), c["s_pre"])
    close(lg.grad, c["dlogits"]); close(lc.grad, c["dloc"]); close(ls.grad, c["dlogstd"])
    for tanh in (0, 1):
        m = case(f"mean_tanh{tanh}")
        close(oh.MixtureGaussianOracle(t64(m["logits"]), t64(m["loc"]), t64(m["logstd"]).exp(), bool(tanh)).mean(), m["mean"])


def test_oracle_sac_policy_loss_matches_reference_source():
    r, c = case("rsample"), case("sac")
    lg, lc, ls = (t64(r[k]).requires_grad_(True) for k in ("logits", "loc", "logstd"))
    log_alpha = torch.tensor(float(c["log_alpha"]), dtype=torch.float64, requires_grad=True)
    dist = oh.MixtureGaussianOracle(lg, lc, ls.exp(), True)
    smp, s_ = dist.sample(1, uniform=t64(r["uniform"]), normal=t64(r["normal"]))
    logp = dist.log_prob((smp[0], s_[0]))
    loss = oh.sac_policy_loss(logp, torch.minimum((smp[0] * t64(c["qw1"])).sum(1), (smp[0] * t64(c["qw2"])).sum(1)),
                              log_alpha, A=r["logits"].shape[1])
    loss.backward()
    close(smp[0].detach(), c["action"]); close(logp.detach(), c["target_log_prob"]); close(loss.detach(), c["loss"])
    close(lg.grad, c["dlogits"]); close(lc.grad, c["dloc"]); close(ls.grad, c["dlogstd"]); close(log_alpha.grad, c["dlog_alpha"])


def test_oracle_particle_grid_matches_reference_source():
    for tanh in (0, 1):
        c = case(f"build_policy_tanh{tanh}")
        A, P = c["loc"].shape
        loc, logstd = oh.init_particles(A, P, tanh=bool(tanh))
        close(loc, c["loc"]); close(logstd, c["logstd"]); close(np.exp(logstd), c["scale"])
        close((c["h"] @ c["weight"] + c["bias"]).reshape(-1, A, P), c["logits"])
        assert not c["bias"].any() and int(c["normalize_output"]) == tanh


@pytest.mark.parametrize("name", RESAMPLE_CASES)
def test_oracle_resampler_matches_reference_source(name):
    c = case("resample_" + name)
    out, ints = orr.resample(c["max_active"], c["sum_active"], c["loc"], c["logstd"], c["bias"], c["weight"],
                             resample=int(c["mode"]), tanh=bool(c["tanh"]), cat_u=c["cat_u"], choice=c["choice"],
                             noise_u=c["noise_u"])
    assert ints["M"] == c["invalid"].shape[0]
    for k in ("invalid", "cand", "tcol", "uniq", "idx", "count", "delta"):
        assert np.array_equal(ints[k], c[k].astype(ints[k].dtype).reshape(ints[k].shape)), k
    for k in ("loc", "logstd", "bias", "weight"):
        assert np.allclose(out[k], c["out_" + k], rtol=2e-6, atol=1e-7), k


# ------------------------------------------------------------------------------- GPU: kernels ---
def rel(a, b):
    a, b = np.asarray(torch.as_tensor(a).detach().cpu(), np.float64), np.asarray(b, np.float64)
    fin = np.isfinite(b)
    assert np.array_equal(np.isfinite(a), fin)
    return float(np.abs(a[fin] - b[fin]).max() / max(np.abs(b[fin]).max(), 1e-30)) if fin.any() else 0.0


def f32(x, dev):
    return torch.tensor(np.asarray(x), dtype=torch.float32, device=dev)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["logprob_plain", "logprob_tanh", "logprob_guard"])
def test_kernel_logprob_entropy_matches_reference_source(cuda_dev, name):
    from pfpn_b200 import _cabi, head
    c = case(name)
    out = head.head_call(_cabi.HEAD_GRAD, f32(c["logits"], cuda_dev), f32(c["loc"], cuda_dev), f32(c["logstd"], cuda_dev),
                         f32(c["value"], cuda_dev), tanh=bool(c["tanh"]), g_lp=f32(c["g_lp"], cuda_dev),
                         g_ent_ba=f32(c["g_ent"], cuda_dev), want_dvalue=True, want_ent_ba=True)
    for k, g in (("lp", "lp"), ("ent_ba", "ent"), ("dlogits", "dlogits"), ("dloc", "dloc"), ("dlogstd", "dlogstd"),
                 ("dvalue", "dvalue")):
        assert rel(out[k], c[g]) < 1e-5, k


@pytest.mark.gpu
def test_kernel_ppo_matches_reference_source(cuda_dev):
    from pfpn_b200 import _cabi, head
    c = case("ppo")
    adv = f32(c["adv"], cuda_dev)
    out = head.head_call(_cabi.HEAD_PPO, f32(c["logits"], cuda_dev), f32(c["loc"], cuda_dev), f32(c["logstd"], cuda_dev),
                         f32(c["value"], cuda_dev), adv=adv, lp_old=f32(c["lp_old"], cuda_dev),
                         adv_stats_t=head.adv_stats(adv), eps_clip=0.2)
    for k in ("dlogits", "dloc", "dlogstd"):
        assert rel(out[k], c[k]) < 1e-5, k
    assert abs(float(out["loss"].cpu()) - float(c["loss"])) < 1e-5 * float(np.abs(c["adv_n"]).mean())


@pytest.mark.gpu
def test_kernel_sampling_matches_reference_source(cuda_dev):
    from pfpn_b200 import sampling
    from pfpn_b200.distribution import MixtureGaussianDistribution
    c = case("sample_plain")
    act, idx = sampling.sample_plain(f32(c["logits"], cuda_dev), f32(c["loc"], cuda_dev), f32(c["logstd"], cuda_dev),
                                     ext_uniform=torch.tensor(c["uniform"], device=cuda_dev), ext_normal=f32(c["normal"], cuda_dev))
    assert np.array_equal(idx.cpu().numpy(), c["dis_action"]) and rel(act, c["action"]) < 1e-6
    c, s = case("rsample"), case("sac")
    args = [f32(c[k], cuda_dev).requires_grad_(True) for k in ("logits", "loc", "logstd")]
    dist = MixtureGaussianDistribution(args[0], args[1], args[2].exp(), True, logstd=args[2])
    smp, s_ = dist.sample(1, ext_uniform=f32(c["uniform"], cuda_dev), ext_normal=f32(c["normal"], cuda_dev))
    assert np.array_equal(dist.dis_action.cpu().numpy(), c["dis_action"])
    assert rel(smp, c["sample"]) < 1e-5 and rel(s_, c["s_pre"]) < 1e-5
    # SAC policy loss (sac.py:166-173) through the drop-in distribution object
    logp = dist.log_prob((smp[0], s_[0]))
    alpha = float(np.exp(s["log_alpha"]))
    q = torch.minimum((smp[0] * f32(s["qw1"], cuda_dev)).sum(1), (smp[0] * f32(s["qw2"], cuda_dev)).sum(1))
    loss = (alpha * logp - q).mean()
    loss.backward()
    assert rel(logp, s["target_log_prob"]) < 1e-5
    for x, k in zip(args, ("dlogits", "dloc", "dlogstd")):
        assert rel(x.grad, s[k]) < 1e-5, k
    for tanh in (0, 1):
        m = case(f"mean_tanh{tanh}")
        a, _ = sampling.mean_action(f32(m["logits"], cuda_dev), f32(m["loc"], cuda_dev), tanh=bool(tanh))
        assert rel(a, m["mean"]) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("name", RESAMPLE_CASES)
def test_kernel_resampler_matches_reference_source(cuda_dev, name):
    from pfpn_b200 import resampling
    c = case("resample_" + name)
    g = {k: f32(c[k], cuda_dev).contiguous() for k in ("max_active", "sum_active", "loc", "logstd", "bias", "weight")}
    out = resampling.resample_(g["max_active"], g["sum_active"], g["loc"], g["logstd"], g["bias"], g["weight"],
                               resample=int(c["mode"]), tanh=bool(c["tanh"]),
                               ext_cat_u=torch.tensor(c["cat_u"], device=cuda_dev),
                               ext_choice=torch.tensor(c["choice"], dtype=torch.int32, device=cuda_dev),
                               ext_noise_u=f32(c["noise_u"], cuda_dev), verify=True)
    M = int(out["M"].item())
    assert M == c["invalid"].shape[0]
    assert np.array_equal(out["cand"].cpu().numpy(), c["cand"])
    assert np.array_equal(out["invalid"].cpu().numpy()[:M], c["invalid"])
    assert np.array_equal(out["tcol"].cpu().numpy()[:M], c["tcol"])
    assert np.array_equal(out["idx"].cpu().numpy()[:M], c["idx"])
    nu = int(out["nuniq"].item())
    for k in ("uniq", "count", "delta"):
        assert np.array_equal(out[k].cpu().numpy()[:nu], c[k].astype(np.int32)), k
    for k in ("loc", "logstd", "bias", "weight"):
        assert np.allclose(g[k].cpu().numpy(), c["out_" + k], rtol=1e-5, atol=1e-6), k
    assert not g["max_active"].any() and not g["sum_active"].any()


def test_product_particle_grid_matches_reference_source():
    """a8: the PRODUCT's initialiser (pfpn_b200.network.initial_particles, what ParticleFilteringClipPPONetwork.init
    writes into `samples` / `samples_std`) against ParticleFilteringA2CNetwork.build_policy executed from source
    (a2c.py:476-535), both grids."""
    from pfpn_b200.network import initial_particles
    for tanh in (0, 1):
        c = case(f"build_policy_tanh{tanh}")
        A, P = c["loc"].shape
        loc, logstd = initial_particles(A, P, bool(tanh))
        assert loc.dtype == torch.float32 and tuple(loc.shape) == (A, P) and tuple(logstd.shape) == (A, P)
        assert np.abs(loc.numpy().astype(np.float64) - c["loc"]).max() <= 2e-7 * max(1.0, np.abs(c["loc"]).max())
        assert np.abs(logstd.numpy().astype(np.float64) - c["logstd"]).max() <= 2e-7 * np.abs(c["logstd"]).max()
        assert np.abs(np.exp(logstd.numpy().astype(np.float64)) - c["scale"]).max() <= 1e-6 * np.abs(c["scale"]).max()
