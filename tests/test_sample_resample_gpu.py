"""K2 / K3 / K4 / K5 parity through the C ABI against the oracle, with externally supplied draws.
Integers (particle indices, dead-particle bookkeeping) must be bit-exact; floats within 1e-5."""
import numpy as np
import pytest
import torch

from oracle import head as oh
from oracle import resample as orr
from pfpn_b200 import resampling, sampling, synth
from pfpn_b200.distribution import MixtureGaussianDistribution
from tests.test_oracle_resample import synth as rs_synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("B,A,P", [(2048, 36, 35), (300, 36, 100), (257, 36, 10), (5, 3, 7)])
def test_plain_sample_indices_bit_exact(cuda_dev, B, A, P):
    d = synth.head_inputs(B, A, P, seed=34114, far_frac=0.0)
    g = torch.Generator().manual_seed(1)
    u = torch.rand(B, A, dtype=torch.float64, generator=g)
    eps = torch.randn(B, A, P, generator=g)
    d["logits"][0, 0, :3] = float("-inf")  # non-finite logits carry no mass in TF's kernel
    dist = oh.MixtureGaussianOracle(d["logits"], d["loc"], d["logstd"].exp(), False)
    ref = dist.sample(1, uniform=u, normal=eps)[0]
    cu = lambda t: t.to(cuda_dev)
    act, idx = sampling.sample_plain(cu(d["logits"]), cu(d["loc"]), cu(d["logstd"]), ext_uniform=cu(u), ext_normal=cu(eps))
    assert torch.equal(idx.cpu().long(), dist.dis_action)
    assert rel(act, ref) < 1e-6
    # the reference-facing object
    gd = MixtureGaussianDistribution(cu(d["logits"]), cu(d["loc"]), cu(d["logstd"]).exp(), False, logstd=cu(d["logstd"]))
    s = gd.sample(1, ext_uniform=cu(u), ext_normal=cu(eps))
    assert s.shape == (1, B, A) and torch.equal(s[0], act) and torch.equal(gd.dis_action, idx)


def test_plain_sample_uniforms_on_cdf_boundaries_take_the_exact_fp64_path(cuda_dev):
    """K2 finds the CDF interval in fp32 and accepts it only with a safety margin; uniforms placed ON the fp64
    interval ends (1e-9 .. 1e-6 of the total away from them) must fall back to the literal fp64 algorithm and agree with it."""
    import numpy as np
    rows, P = 6000, 35
    rng = np.random.RandomState(7)
    logits = (rng.randn(rows, P) * 2).astype(np.float32)
    logits[5, :4] = -np.inf
    logits[6, 10] = -120.0  # underflows in fp32, not in fp64
    mx = logits.max(1, keepdims=True).astype(np.float64)
    e = np.where(np.isfinite(logits), np.exp(logits.astype(np.float64) - mx), 0.0)
    cdf = np.cumsum(e, 1)
    k = rng.randint(0, P - 1, size=rows)
    # 1e-9 .. 1e-6 away from an interval end, relative to the total: far inside the fp32 uncertainty zone (forces the
    # fallback), far outside the last-ulp differences between libm implementations of exp(double)
    off = 10.0 ** rng.uniform(-9, -6, size=rows) * rng.choice([-1.0, 1.0], size=rows)
    u = np.clip(cdf[np.arange(rows), k] / cdf[:, -1] + off, 0.0, np.nextafter(1.0, 0.0))
    ref = oh.tf_multinomial_cpu(logits, u[:, None])[:, 0]
    lg = torch.tensor(logits, device=cuda_dev).view(rows, 1, P)
    loc = torch.zeros(1, P, device=cuda_dev)
    ls = torch.zeros(1, P, device=cuda_dev)
    _, idx = sampling.sample_plain(lg, loc, ls, ext_uniform=torch.tensor(u, device=cuda_dev).view(rows, 1),
                                   ext_normal=torch.zeros(rows, 1, P, device=cuda_dev))
    assert np.array_equal(idx.cpu().numpy().reshape(-1), ref)


def test_plain_sample_philox_is_distributed_like_the_mixture(cuda_dev):
    B, A, P = 200000, 2, 5
    logits = torch.tensor([[0.0, 1.0, 2.0, -1.0, 0.5], [3.0, 0.0, 0.0, 0.0, -2.0]]).repeat(B, 1, 1).contiguous()
    loc, logstd = synth.particle_grid(A, P, torch.Generator().manual_seed(0))
    cu = lambda t: t.to(cuda_dev)
    act, idx = sampling.sample_plain(cu(logits), cu(loc), cu(logstd), seed=123, offset=7)
    act2, idx2 = sampling.sample_plain(cu(logits), cu(loc), cu(logstd), seed=123, offset=7)
    assert torch.equal(idx, idx2) and torch.equal(act, act2)  # counter-based: reproducible
    _, idx3 = sampling.sample_plain(cu(logits), cu(loc), cu(logstd), seed=123, offset=8)
    assert not torch.equal(idx, idx3)
    freq = torch.stack([torch.bincount(idx[:, a].long().cpu(), minlength=P) for a in range(A)]).double() / B
    assert (freq - torch.softmax(logits[0].double(), -1)).abs().max() < 5e-3
    z = (act.cpu() - loc.expand(B, A, P).gather(2, idx.cpu().long()[..., None])[..., 0]) / \
        logstd.exp().expand(B, A, P).gather(2, idx.cpu().long()[..., None])[..., 0]
    assert abs(float(z.mean())) < 1e-2 and abs(float(z.std()) - 1) < 1e-2


@pytest.mark.parametrize("B,A,P", [(1024, 36, 100), (300, 36, 35), (7, 4, 10)])
def test_rsample_forward_backward(cuda_dev, B, A, P):
    g = torch.Generator().manual_seed(12831)
    logits = torch.randn(B, A, P, generator=g) * 2
    loc_np, logstd_np = oh.init_particles(A, P, tanh=True)
    loc = torch.tensor(loc_np, dtype=torch.float32)
    logstd = torch.tensor(logstd_np, dtype=torch.float32)
    U = torch.rand(B, A, P, generator=g).clamp_min(oh.F32_TINY)
    eps = torch.randn(B, A, P, generator=g)
    g_a = torch.randn(B, A, generator=g)
    g_u = torch.randn(B, A, generator=g) * 0.3
    lg = logits.double().requires_grad_(True)
    lc = loc.double().requires_grad_(True)
    ls = logstd.double().requires_grad_(True)
    dist = oh.MixtureGaussianOracle(lg, lc, ls.exp(), True)
    sample, s_ = dist.sample(1, uniform=U.double(), normal=eps.double())
    (sample[0] * g_a.double()).sum().backward(retain_graph=True)
    (s_[0] * g_u.double()).sum().backward()
    cu = lambda t: t.to(cuda_dev)
    ext = dict(ext_uniform=cu(U), ext_normal=cu(eps))
    smp, spre, idx = sampling.rsample_fwd(cu(logits), cu(loc), cu(logstd), **ext)
    assert torch.equal(idx.cpu().long(), dist.dis_action)
    assert rel(smp, sample[0]) < TOL and rel(spre, s_[0]) < TOL
    dl, dc, ds = sampling.rsample_bwd(cu(logits), cu(loc), cu(logstd), cu(g_a), cu(g_u), **ext)
    assert rel(dl, lg.grad) < TOL and rel(dc, lc.grad) < TOL and rel(ds, ls.grad) < TOL
    # no atomics: the particle gradients are reproducible bit for bit
    dl2, dc2, ds2 = sampling.rsample_bwd(cu(logits), cu(loc), cu(logstd), cu(g_a), cu(g_u), **ext)
    assert torch.equal(dc, dc2) and torch.equal(ds, ds2) and torch.equal(dl, dl2)


@pytest.mark.parametrize("P", [35, 100, 200])
def test_rsample_philox_mode_distribution_and_fwd_bwd_consistency(cuda_dev, P):
    """Production mode (Philox draws, MUFU arithmetic): the selected particle follows softmax(logits), the
    location draw is N(0,1), and the backward regenerates exactly the draws the forward used."""
    B, A = 20000, 6
    g = torch.Generator().manual_seed(77 + P)
    row = torch.randn(A, P, generator=g) * 1.5
    logits = row.expand(B, A, P).contiguous()
    loc_np, logstd_np = oh.init_particles(A, P, tanh=True)
    loc = torch.tensor(loc_np, dtype=torch.float32)
    logstd = torch.tensor(logstd_np, dtype=torch.float32) + 0.3 * torch.randn(A, P, generator=g)
    cu = lambda t: t.to(cuda_dev)
    kw = dict(seed=991, offset=5)
    smp, spre, idx = sampling.rsample_fwd(cu(logits), cu(loc), cu(logstd), **kw)
    smp2, spre2, idx2 = sampling.rsample_fwd(cu(logits), cu(loc), cu(logstd), **kw)
    assert torch.equal(idx, idx2) and torch.equal(spre, spre2) and torch.equal(smp, smp2)
    _, _, idx3 = sampling.rsample_fwd(cu(logits), cu(loc), cu(logstd), seed=991, offset=6)
    assert not torch.equal(idx, idx3)
    idx_l = idx.cpu().long()
    assert int(idx_l.min()) >= 0 and int(idx_l.max()) < P
    pi = torch.softmax(row.double(), -1)
    for a in range(A):
        freq = torch.bincount(idx_l[:, a], minlength=P).double() / B
        sd = (pi[a] * (1 - pi[a]) / B).sqrt()
        assert float(((freq - pi[a]).abs() / (sd + 1e-4)).max()) < 5.0
    a_ix = torch.arange(A).expand(B, A)
    z = (spre.cpu().double() - loc.double()[a_ix, idx_l]) / logstd.double().exp()[a_ix, idx_l]
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1) < 0.02
    assert abs(float((z ** 4).mean()) - 3.0) < 0.15 and float(z.abs().max()) < 6.5
    assert rel(smp, torch.tanh(spre.cpu().double())) < TOL
    # backward on the same (seed, offset): d loc[a,k] = sum over rows that selected k of sech^2(s_) g_a + g_u
    g_a = torch.randn(B, A, generator=g)
    g_u = torch.randn(B, A, generator=g) * 0.3
    dl, dc, ds = sampling.rsample_bwd(cu(logits), cu(loc), cu(logstd), cu(g_a), cu(g_u), **kw)
    s64 = spre.cpu().double()
    gp = (1 - torch.tanh(s64) ** 2) * g_a.double() + g_u.double()
    dc_ref = torch.zeros(A, P, dtype=torch.float64).index_put_((a_ix.reshape(-1), idx_l.reshape(-1)), gp.reshape(-1),
                                                              accumulate=True)
    ds_ref = torch.zeros(A, P, dtype=torch.float64).index_put_(
        (a_ix.reshape(-1), idx_l.reshape(-1)), (gp * (s64 - loc.double()[a_ix, idx_l])).reshape(-1), accumulate=True)
    assert rel(dc, dc_ref) < 1e-4 and rel(ds, ds_ref) < 1e-4
    # softmax Jacobian rows sum to zero -- up to fp32 rounding of the row's own terms (the straight-through coefficient
    # g_u / max(1e-6, 1 - t^2) can be huge for a saturated tanh, so the bound is relative to the row's magnitude)
    assert bool((dl.sum(-1).abs() <= 4e-6 * dl.abs().sum(-1) + 1e-6).all())
    assert torch.isfinite(dl).all()


def test_sac_policy_gradient_through_the_distribution_object(cuda_dev):
    """sample -> log_prob((sample, s_)) -> alpha*logp - q(sample): the composition of sac.py:128-130,166-173."""
    B, A, P = 256, 36, 35
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(B, A, P, generator=g) * 2
    loc_np, logstd_np = oh.init_particles(A, P, tanh=True)
    loc, logstd = torch.tensor(loc_np, dtype=torch.float32), torch.tensor(logstd_np, dtype=torch.float32)
    U = torch.rand(B, A, P, generator=g).clamp_min(oh.F32_TINY)
    eps = torch.randn(B, A, P, generator=g)
    qw = torch.randn(A, generator=g)

    def run(mod, lg, lc, ls, to):
        if mod is oh.MixtureGaussianOracle:
            dist = mod(lg, lc, ls.exp(), True)
            smp, s_ = dist.sample(1, uniform=to(U), normal=to(eps))
        else:
            dist = mod(lg, lc, ls.exp(), True, logstd=ls)
            smp, s_ = dist.sample(1, ext_uniform=to(U), ext_normal=to(eps))
        smp, s_ = smp[0], s_[0]
        logp = dist.log_prob((smp, s_))
        loss = (0.2 * logp - (smp * to(qw)).sum(1)).mean()
        loss.backward()
        return logp

    a64 = [t.double().requires_grad_(True) for t in (logits, loc, logstd)]
    lp_ref = run(oh.MixtureGaussianOracle, *a64, lambda t: t.double())
    agpu = [t.to(cuda_dev).requires_grad_(True) for t in (logits, loc, logstd)]
    lp = run(MixtureGaussianDistribution, *agpu, lambda t: t.to(cuda_dev))
    assert rel(lp, lp_ref) < TOL
    for x, r in zip(agpu, a64):
        assert rel(x.grad, r.grad) < TOL


def test_mean_action(cuda_dev):
    d = synth.head_inputs(500, 36, 35, seed=3, far_frac=0.0)
    cu = lambda t: t.to(cuda_dev)
    for tanh in (False, True):
        ref = oh.MixtureGaussianOracle(d["logits"], d["loc"], d["logstd"].exp(), tanh).mean()
        act, idx = sampling.mean_action(cu(d["logits"]), cu(d["loc"]), tanh=tanh)
        assert torch.equal(idx.cpu().long(), d["logits"].argmax(-1))
        assert rel(act, ref) < 1e-6


def test_activity_statistics(cuda_dev):
    d = synth.head_inputs(4096, 36, 35, seed=9, far_frac=0.0)
    probs = torch.softmax(d["logits"].double(), -1)
    mx0 = torch.rand(36, 35) * 0.05
    sm0 = torch.rand(36, 35)
    ref_max, ref_sum = orr.activity_stats(probs.numpy(), mx0.numpy(), sm0.numpy())
    mx, sm = mx0.to(cuda_dev), sm0.to(cuda_dev)
    p = sampling.stats_update(d["logits"].to(cuda_dev), mx, sm, want_probs=True)
    assert rel(p, probs) < 1e-6
    assert rel(mx, torch.maximum(mx0.double(), probs.max(0).values)) < 1e-6
    assert rel(sm, sm0.double() + probs.sum(0)) < 1e-5
    assert rel(mx, ref_max) < 1e-6 and rel(sm, ref_sum) < 1e-4
    gd = MixtureGaussianDistribution(d["logits"].to(cuda_dev), d["loc"].to(cuda_dev), d["logstd"].exp().to(cuda_dev), False)
    assert rel(gd.dis_dist.probs, probs) < 1e-6
    # deterministic (no float atomics): a second pass from the same state reproduces the statistics bit for bit
    mx2, sm2 = mx0.to(cuda_dev), sm0.to(cuda_dev)
    sampling.stats_update(d["logits"].to(cuda_dev), mx2, sm2)
    assert torch.equal(mx, mx2) and torch.equal(sm, sm2)


@pytest.mark.parametrize("B,A,P", [(1, 36, 35), (7, 3, 100), (1001, 36, 10), (33, 5, 200)])
def test_activity_statistics_ragged_shapes(cuda_dev, B, A, P):
    g = torch.Generator().manual_seed(B + A + P)
    logits = torch.randn(B, A, P, generator=g) * 3
    probs = torch.softmax(logits.double(), -1)
    mx, sm = torch.zeros(A, P, device=cuda_dev), torch.zeros(A, P, device=cuda_dev)
    sampling.stats_update(logits.to(cuda_dev), mx, sm)
    sampling.stats_update(logits.to(cuda_dev), mx, sm)  # running update: max stays, sum doubles
    assert rel(mx, probs.max(0).values) < 1e-6 and rel(sm, 2 * probs.sum(0)) < 1e-5


INT_KEYS = ("invalid", "src", "col", "tcol", "idx")


def _run_resample(cuda_dev, t, d, **kw):
    ref, ints = orr.resample(**t, **d, **kw)
    g = {k: torch.tensor(v).to(cuda_dev) for k, v in t.items()}
    out = resampling.resample_(g["max_active"], g["sum_active"], g["loc"], g["logstd"], g["bias"], g["weight"],
                               ext_cat_u=torch.tensor(d["cat_u"]).to(cuda_dev),
                               ext_choice=torch.tensor(d["choice"]).to(cuda_dev),
                               ext_noise_u=torch.tensor(d["noise_u"]).to(cuda_dev), verify=True, **kw)
    M = int(out["M"].item())
    assert M == ints["M"]
    assert np.array_equal(out["cand"].cpu().numpy(), ints["cand"])
    for k in INT_KEYS:
        assert np.array_equal(out[k].cpu().numpy()[:M], ints[k]), k
    nu = int(out["nuniq"].item())
    assert nu == len(ints["uniq"])
    for k in ("uniq", "count", "delta"):
        assert np.array_equal(out[k].cpu().numpy()[:nu], ints[k]), k
    for k in ("loc", "logstd", "bias", "weight", "max_active", "sum_active"):
        assert np.allclose(g[k].cpu().numpy(), ref[k], rtol=1e-5, atol=1e-6), k  # floats: tolerance; ints above: exact
    return M


@pytest.mark.parametrize("P", [10, 35, 100])
@pytest.mark.parametrize("dead_frac,all_but_one", [(0.0, False), (0.03, False), (0.1, False), (0.5, False), (1.0, True)])
def test_resample_sweep_bit_exact(cuda_dev, P, dead_frac, all_but_one):
    """BASELINE config c3: P = 10/35/100, A = 36, dead fraction sweep."""
    t, d, dead = rs_synth(A=36, P=P, H=512, dead_frac=dead_frac, all_but_one=all_but_one, seed=33406 + P)
    M = _run_resample(cuda_dev, t, d, resample=-1)
    assert M == int(dead.sum())


def test_resample_tanh_and_topk(cuda_dev):
    t, d, _ = rs_synth(A=36, P=35, H=64, dead_frac=0.2)
    _run_resample(cuda_dev, t, d, resample=-1, tanh=True)
    _run_resample(cuda_dev, t, d, resample=3)


def test_resample_dead_source_case(cuda_dev):
    t, d, _ = rs_synth(A=2, P=6, H=4, dead_frac=0.0)
    t["max_active"][0, 1] = 1e-6
    t["max_active"][0, 4] = 1e-6
    t["sum_active"][0, :] = np.array([0, 1, 0, 0, 0, 0], np.float32) + 1e-12
    assert _run_resample(cuda_dev, t, d, resample=-1) == 2


def test_resample_philox_mode_is_replica_deterministic(cuda_dev):
    """every rank runs the resampler redundantly from the same seed -> identical results."""
    t, d, _ = rs_synth(A=36, P=35, H=512, dead_frac=0.2)
    outs = []
    for _ in range(2):
        g = {k: torch.tensor(v).to(cuda_dev) for k, v in t.items()}
        o = resampling.resample_(g["max_active"], g["sum_active"], g["loc"], g["logstd"], g["bias"], g["weight"],
                                 seed=33407, offset=368, verify=True)
        outs.append((g, o))
    for k in ("loc", "logstd", "bias", "weight"):
        assert torch.equal(outs[0][0][k], outs[1][0][k])
    assert torch.equal(outs[0][1]["src"], outs[1][1]["src"])


# ---------------------------------------------------------------------------------------------------------------------
# K3f: the SAC head in one pass (pfpn_sac_head_fwd_bwd) -- rsample forward + tanh log_prob forward + backward of both
# ---------------------------------------------------------------------------------------------------------------------
def _sac_inputs(B, A, P, seed):
    from pfpn_b200.network import initial_particles
    g = torch.Generator().manual_seed(seed)
    logits = torch.randn(B, A, P, generator=g) * 2.0
    loc, logstd = initial_particles(A, P, True)
    loc = loc + 0.02 * torch.randn(A, P, generator=g)
    logstd = logstd + 0.1 * torch.randn(A, P, generator=g)
    g_sample = torch.randn(B, A, generator=g)
    g_lp = torch.randn(B, generator=g) * 0.3
    tiny = float(np.finfo(np.float32).tiny)
    U = torch.rand(B, A, P, generator=g).clamp_min(tiny)
    EPS = torch.randn(B, A, P, generator=g)
    return logits, loc, logstd, g_sample, g_lp, U, EPS


@pytest.mark.parametrize("B", [37, 256])
def test_fused_sac_head_matches_oracle_with_external_draws(cuda_dev, B):
    """Verification mode against the fp64 oracle fed the same uniforms / normals (utils.py:108-144,156-186 with the three
    custom gradients): argmax particle bit-exact, everything else 1e-5 norm-wise.  B = 37: ragged last tile."""
    from oracle import head as oh
    from pfpn_b200 import sampling
    A, P = 36, 100
    logits, loc, logstd, g_sample, g_lp, U, EPS = _sac_inputs(B, A, P, seed=B)
    lg, lc, ls = (t.double().requires_grad_(True) for t in (logits, loc, logstd))
    dist = oh.MixtureGaussianOracle(lg, lc, ls.exp(), True)
    smp, s_ = dist.sample(1, uniform=U.double(), normal=EPS.double())
    logp = dist.log_prob((smp[0], s_[0]))
    (torch.sum(g_sample.double() * smp[0]) + torch.sum(g_lp.double() * logp)).backward()
    cu = lambda t: t.to(cuda_dev)
    out = sampling.sac_head_fused(cu(logits), cu(loc), cu(logstd), cu(g_sample), cu(g_lp), ext_uniform=cu(U), ext_normal=cu(EPS))
    torch.cuda.synchronize()
    assert np.array_equal(out["idx"].cpu().numpy(), dist.dis_action.numpy())
    assert rel(out["s_pre"], s_[0].detach()) < TOL and rel(out["sample"], smp[0].detach()) < TOL
    assert rel(out["logp"], logp.detach()) < TOL
    assert rel(out["dlogits"], lg.grad) < TOL
    assert rel(out["dloc"], lc.grad) < TOL and rel(out["dlogstd"], ls.grad) < TOL


def test_fused_sac_head_equals_the_three_launch_form_in_production_mode(cuda_dev):
    """Philox mode: with the same (seed, offset) the fused kernel draws the same variates as pfpn_head_rsample_fwd / _bwd, so
    it must reproduce the boundary-faithful three-launch composition (rsample fwd, tanh log_prob fwd+bwd with dvalue,
    rsample bwd) at a size no external-draw test reaches -- and be bit-reproducible."""
    from pfpn_b200 import _cabi, head, sampling
    B, A, P = 4099, 36, 100
    logits, loc, logstd, g_sample, g_lp, _, _ = _sac_inputs(B, A, P, seed=5)
    cu = lambda t: t.to(cuda_dev)
    logits, loc, logstd, g_sample, g_lp = map(cu, (logits, loc, logstd, g_sample, g_lp))
    smp, s_pre, idx = sampling.rsample_fwd(logits, loc, logstd, seed=11, offset=6)
    o = head.head_call(_cabi.HEAD_GRAD, logits, loc, logstd, s_pre, tanh=True, g_lp=g_lp, want_dvalue=True)
    dl2, dloc2, dls2 = sampling.rsample_bwd(logits, loc, logstd, g_sample, o["dvalue"], seed=11, offset=6)
    ref = dict(dlogits=o["dlogits"] + dl2, dloc=o["dloc"] + dloc2, dlogstd=o["dlogstd"] + dls2)
    f1 = sampling.sac_head_fused(logits, loc, logstd, g_sample, g_lp, seed=11, offset=6)
    f2 = sampling.sac_head_fused(logits, loc, logstd, g_sample, g_lp, seed=11, offset=6)
    torch.cuda.synchronize()
    assert torch.equal(f1["idx"], idx)
    assert rel(f1["s_pre"], s_pre) < 1e-6 and rel(f1["sample"], smp) < 1e-6
    assert rel(f1["logp"], o["lp"]) < TOL
    for k in ("dlogits", "dloc", "dlogstd"):
        assert rel(f1[k], ref[k]) < 2 * TOL, k
        assert torch.equal(f1[k], f2[k]), k  # no atomics, fixed summation order
    f3 = sampling.sac_head_fused(logits, loc, logstd, g_sample, g_lp, seed=11, offset=8)
    assert not torch.equal(f3["idx"], idx)  # another call counter -> other draws


# ---------------------------------------------------------------------------------------------------------------------
# K2f: the rollout side in one pass (pfpn_head_rollout) -- sample + log_prob + entropy + activity statistics
# ---------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B", [2050, 7])
def test_fused_rollout_matches_oracle_with_external_draws(cuda_dev, B):
    A, P = 36, 35
    d = synth.head_inputs(B, A, P, seed=34114 + B, far_frac=0.0)
    g = torch.Generator().manual_seed(B)
    u = torch.rand(B, A, dtype=torch.float64, generator=g)
    eps = torch.randn(B, A, P, generator=g)
    d["logits"][0, 0, :3] = float("-inf")  # non-finite logits carry no mass in TF's Multinomial kernel
    dist = oh.MixtureGaussianOracle(d["logits"].double(), d["loc"].double(), d["logstd"].double().exp(), False)
    act_ref = dist.sample(1, uniform=u, normal=eps.double())[0]
    cu = lambda t: t.to(cuda_dev)
    max0, sum0 = torch.rand(A, P, generator=g) * 0.05, torch.rand(A, P, generator=g)
    mx, sm = cu(max0), cu(sum0)
    out = sampling.rollout_fused(cu(d["logits"]), cu(d["loc"]), cu(d["logstd"]), ext_uniform=cu(u), ext_normal=cu(eps),
                                 max_active=mx, sum_active=sm, want_ent=True)
    torch.cuda.synchronize()
    assert torch.equal(out["idx"].cpu().long(), dist.dis_action)          # bit-exact particle selection
    assert rel(out["action"], act_ref) < 1e-6
    # log_prob / entropy / statistics on finite rows (row (0,0) holds -inf logits: softmax of it is the same in both)
    lp_ref = dist.log_prob(act_ref)
    assert rel(out["lp"], lp_ref) < TOL
    lg = d["logits"].double().clone()
    probs = torch.softmax(lg, -1)
    ent_ref = -(torch.where(probs > 0, probs * probs.clamp_min(1e-300).log(), torch.zeros_like(probs))).sum(-1).sum(-1)
    assert rel(out["ent"], ent_ref) < TOL
    assert rel(mx, torch.maximum(max0.double(), probs.amax(0))) < TOL
    assert rel(sm, sum0.double() + probs.sum(0)) < TOL


def test_fused_rollout_uniforms_on_cdf_boundaries_take_the_exact_fp64_path(cuda_dev):
    """Same adversarial uniforms as the K2 test: 1e-9 .. 1e-6 of the total away from an fp64 CDF interval end."""
    import numpy as np
    B, A, P = 170, 36, 35
    rows = B * A
    rng = np.random.RandomState(7)
    logits = (rng.randn(rows, P) * 2).astype(np.float32)
    logits[5, :4] = -np.inf
    logits[6, 10] = -120.0  # underflows in fp32, not in fp64
    mxv = logits.max(1, keepdims=True).astype(np.float64)
    e = np.where(np.isfinite(logits), np.exp(logits.astype(np.float64) - mxv), 0.0)
    cdf = np.cumsum(e, 1)
    k = rng.randint(0, P - 1, size=rows)
    off = 10.0 ** rng.uniform(-9, -6, size=rows) * rng.choice([-1.0, 1.0], size=rows)
    u = np.clip(cdf[np.arange(rows), k] / cdf[:, -1] + off, 0.0, np.nextafter(1.0, 0.0))
    ref = oh.tf_multinomial_cpu(logits, u[:, None])[:, 0]
    lg = torch.tensor(logits, device=cuda_dev).view(B, A, P)
    loc = torch.zeros(A, P, device=cuda_dev)
    ls = torch.zeros(A, P, device=cuda_dev)
    out = sampling.rollout_fused(lg, loc, ls, ext_uniform=torch.tensor(u, device=cuda_dev).view(B, A),
                                 ext_normal=torch.zeros(B, A, P, device=cuda_dev))
    assert np.array_equal(out["idx"].cpu().numpy().reshape(-1), ref)


def test_fused_rollout_equals_the_three_kernel_form_in_production_mode(cuda_dev):
    """Philox mode at a size the oracle does not reach: same streams as K2, so index and action are bit-identical to
    pfpn_head_sample, log_prob / entropy to K1's forward, the statistics to K4 -- from ONE pass over the logits."""
    from pfpn_b200 import _cabi, head
    B, A, P = 16387, 36, 35
    d = {k: (v.to(cuda_dev) if torch.is_tensor(v) else v) for k, v in synth.head_inputs(B, A, P, seed=3).items()}
    act, idx = sampling.sample_plain(d["logits"], d["loc"], d["logstd"], seed=77, offset=4)
    ref = head.head_call(_cabi.HEAD_FWD, d["logits"], d["loc"], d["logstd"], act)
    mx0, sm0 = torch.zeros(A, P, device=cuda_dev), torch.zeros(A, P, device=cuda_dev)
    sampling.stats_update(d["logits"], mx0, sm0)
    mx1, sm1 = torch.zeros(A, P, device=cuda_dev), torch.zeros(A, P, device=cuda_dev)
    out = sampling.rollout_fused(d["logits"], d["loc"], d["logstd"], seed=77, offset=4, max_active=mx1, sum_active=sm1,
                                 want_ent=True)
    torch.cuda.synchronize()
    assert torch.equal(out["idx"], idx) and torch.equal(out["action"], act)
    assert rel(out["lp"], ref["lp"]) < TOL and rel(out["ent"], ref["ent"]) < TOL
    assert rel(mx1, mx0) < 1e-6 and rel(sm1, sm0) < 1e-5
    out2 = sampling.rollout_fused(d["logits"], d["loc"], d["logstd"], seed=77, offset=4, want_ent=True)
    assert torch.equal(out2["lp"], out["lp"]) and torch.equal(out2["idx"], out["idx"])  # reproducible; statistics optional
