"""K1 parity: CUDA head (through the C ABI) vs the fp64 oracle on the same seeded
inputs.  Tolerance: ||delta||_inf / ||ref||_inf <= 1e-5 per tensor."""
import math

import pytest
import torch

from oracle import head as oh
from pfpn_b200 import _cabi, head, synth
from pfpn_b200.distribution import MixtureGaussianDistribution

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    fin = torch.isfinite(b)
    assert torch.equal(torch.isfinite(a), fin), "non-finite pattern differs"
    if fin.sum() == 0:
        return 0.0
    return float((a[fin] - b[fin]).abs().max() / b[fin].abs().max().clamp_min(1e-30))


def make(B, A, P, seed=34114, far=0.0):
    d = synth.head_inputs(B, A, P, seed=seed, far_frac=far)
    ref0 = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], torch.zeros(B))
    lp0 = torch.where(torch.isfinite(ref0["lp"]), ref0["lp"], torch.zeros_like(ref0["lp"]))
    d["lp_old"] = (lp0 + d["lp_noise"].double()).float()
    return d


@pytest.mark.parametrize("B,A,P", [(4096, 36, 35), (512, 36, 10), (384, 36, 100), (257, 6, 35),
                                   (33, 1, 7), (130, 17, 50), (64, 3, 200)])
def test_ppo_fused_matches_oracle(cuda_dev, B, A, P):
    d = make(B, A, P)
    ref = oh.ppo_head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], d["adv"], d["lp_old"])
    cu = lambda t: t.to(cuda_dev)
    stats = head.adv_stats(cu(d["adv"]))
    out = head.head_call(_cabi.HEAD_PPO, cu(d["logits"]), cu(d["loc"]), cu(d["logstd"]), cu(d["value"]),
                         adv=cu(d["adv"]), lp_old=cu(d["lp_old"]), adv_stats_t=stats, eps_clip=0.2)
    assert rel(out["lp"], ref["lp"]) < TOL
    assert rel(out["ent"], ref["ent"].sum(1)) < TOL
    assert rel(out["dlogits"], ref["dlogits"]) < TOL
    assert rel(out["dloc"], ref["dloc"]) < TOL
    assert rel(out["dlogstd"], ref["dlogstd"]) < TOL
    # the loss is a mean of B signed terms: bound the error by the mean |term| (norm-wise)
    an = oh.normalize_advantage(d["adv"].double())
    ratio = torch.exp(ref["lp"] - d["lp_old"].double())
    scale = torch.minimum(ratio * an, ratio.clamp(0.8, 1.2) * an).abs().mean()
    assert abs(float(out["loss"].cpu()) - float(ref["loss"])) < TOL * float(scale)


@pytest.mark.parametrize("tanh", [False, True])
@pytest.mark.parametrize("B,A,P", [(1000, 36, 35), (200, 36, 100), (5, 4, 10)])
def test_grad_mode_with_entropy_and_dvalue(cuda_dev, B, A, P, tanh):
    d = make(B, A, P, seed=33406)
    g = torch.Generator().manual_seed(7)
    g_lp = torch.randn(B, generator=g)
    g_ent = torch.randn(B, A, generator=g) * 0.1
    # with tanh=True `value` is the pre-tanh u (utils.py:120-126)
    ref = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], g_lp, None, tanh=tanh,
                          want_dvalue=True)
    # entropy upstream given per (b,a): oracle takes per-b weight, so fold manually
    lg = d["logits"].double().requires_grad_(True)
    ent = oh.MixtureGaussianOracle(lg, d["loc"].double(), d["logstd"].double().exp(), tanh).entropy()
    (ent * g_ent.double()).sum().backward()
    ref_dlogits = ref["dlogits"] + lg.grad
    cu = lambda t: t.to(cuda_dev)
    out = head.head_call(_cabi.HEAD_GRAD, cu(d["logits"]), cu(d["loc"]), cu(d["logstd"]), cu(d["value"]),
                         tanh=tanh, g_lp=cu(g_lp), g_ent_ba=cu(g_ent), want_dvalue=True, want_ent_ba=True)
    assert rel(out["lp"], ref["lp"]) < TOL
    assert rel(out["ent_ba"], ent) < TOL
    assert rel(out["dlogits"], ref_dlogits) < TOL
    assert rel(out["dloc"], ref["dloc"]) < TOL
    assert rel(out["dlogstd"], ref["dlogstd"]) < TOL
    assert rel(out["dvalue"], ref["dvalue"]) < TOL


@pytest.mark.parametrize("B", [1, 2, 3, 4, 5, 7, 8, 9, 4097])
def test_ragged_batches_forward(cuda_dev, B):
    d = make(B, 36, 35, seed=28949)
    ref = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], torch.zeros(B))
    cu = lambda t: t.to(cuda_dev)
    out = head.head_call(_cabi.HEAD_FWD, cu(d["logits"]), cu(d["loc"]), cu(d["logstd"]), cu(d["value"]))
    assert rel(out["lp"], ref["lp"]) < TOL
    assert rel(out["ent"], ref["ent"].sum(1)) < TOL


def test_empty_batch_is_a_noop(cuda_dev):
    z = torch.zeros(0, 36, 35, device=cuda_dev)
    loc, logstd = synth.particle_grid(36, 35, torch.Generator().manual_seed(0))
    out = head.head_call(_cabi.HEAD_FWD, z, loc.to(cuda_dev), logstd.to(cuda_dev),
                         torch.zeros(0, 36, device=cuda_dev))
    assert out["lp"].numel() == 0


def test_underflow_guard_zeroes_gradient(cuda_dev):
    """utils.py:109-117: p == 0 => lp = -inf and the row's gradient is exactly 0."""
    B, A, P = 64, 36, 35
    d = make(B, A, P, seed=12831)
    d["value"][:8, 5] = 9.0  # > 100 sigma from every particle: exp underflows in fp32 and fp64
    g_lp = torch.ones(B)
    ref = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], g_lp)
    cu = lambda t: t.to(cuda_dev)
    out = head.head_call(_cabi.HEAD_GRAD, cu(d["logits"]), cu(d["loc"]), cu(d["logstd"]), cu(d["value"]),
                         g_lp=cu(g_lp))
    lp = out["lp"].cpu()
    assert torch.isinf(lp[:8]).all() and (lp[:8] < 0).all() and torch.isfinite(lp[8:]).all()
    assert torch.isinf(ref["lp"][:8]).all()
    dl = out["dlogits"].cpu()
    assert torch.count_nonzero(dl[:8, 5]) == 0
    assert torch.isfinite(dl).all()
    assert rel(dl, ref["dlogits"]) < TOL
    assert rel(out["dloc"], ref["dloc"]) < TOL and rel(out["dlogstd"], ref["dlogstd"]) < TOL


def test_inplace_gradient_aliases_logits(cuda_dev):
    d = make(777, 36, 35, seed=39907)
    cu = lambda t: t.to(cuda_dev)
    g_lp = cu(torch.randn(777, generator=torch.Generator().manual_seed(1)))
    lg = cu(d["logits"])
    a = head.head_call(_cabi.HEAD_GRAD, lg, cu(d["loc"]), cu(d["logstd"]), cu(d["value"]), g_lp=g_lp)
    lg2 = lg.clone()
    b = head.head_call(_cabi.HEAD_GRAD, lg2, cu(d["loc"]), cu(d["logstd"]), cu(d["value"]), g_lp=g_lp,
                       dlogits_out=lg2)
    assert torch.equal(a["dlogits"], b["dlogits"]) and b["dlogits"].data_ptr() == lg2.data_ptr()
    assert torch.equal(a["dloc"], b["dloc"])


def test_deterministic_across_launches(cuda_dev):
    d = make(5000, 36, 35, seed=3)
    cu = lambda t: t.to(cuda_dev)
    args = (cu(d["logits"]), cu(d["loc"]), cu(d["logstd"]), cu(d["value"]))
    g_lp = cu(torch.randn(5000, generator=torch.Generator().manual_seed(2)))
    a = head.head_call(_cabi.HEAD_GRAD, *args, g_lp=g_lp)
    b = head.head_call(_cabi.HEAD_GRAD, *args, g_lp=g_lp)
    for k in ("lp", "dlogits", "dloc", "dlogstd"):
        assert torch.equal(a[k], b[k]), k


def test_autograd_wrapper_matches_oracle(cuda_dev):
    """The reference-facing object: MixtureGaussianDistribution.log_prob / entropy + backward."""
    B, A, P = 300, 36, 35
    d = make(B, A, P, seed=5)
    w = torch.randn(B, generator=torch.Generator().manual_seed(9))
    ref = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], w, torch.full((B,), -0.01))
    lg = d["logits"].to(cuda_dev).requires_grad_(True)
    loc = d["loc"].to(cuda_dev).requires_grad_(True)
    ls = d["logstd"].to(cuda_dev).requires_grad_(True)
    dist = MixtureGaussianDistribution(lg, loc, torch.exp(ls), False, logstd=ls)
    lp = dist.log_prob(d["value"].to(cuda_dev))
    ent = dist.entropy()
    assert ent.shape == (B, A)
    loss = (lp * w.to(cuda_dev)).sum() - 0.01 * ent.sum()
    loss.backward()
    assert rel(lp, ref["lp"]) < TOL and rel(ent, ref["ent"]) < TOL
    assert rel(lg.grad, ref["dlogits"]) < TOL
    assert rel(loc.grad, ref["dloc"]) < TOL and rel(ls.grad, ref["dlogstd"]) < TOL
    assert rel(dist.prob(d["value"].to(cuda_dev)), ref["lp"].exp()) < 1e-4


def test_full_size_shard_linearity(cuda_dev):
    """BASELINE c4 size (B=65536): per-state outputs of the whole batch equal those of
    its two halves bit-for-bit; dloc/dlogstd of the whole equal the sum of the halves."""
    B, A, P = 65536, 36, 35
    g = torch.Generator(device="cuda").manual_seed(1)
    logits = torch.randn(B, A, P, device=cuda_dev, generator=g) * 2
    loc, logstd = synth.particle_grid(A, P, torch.Generator().manual_seed(0))
    loc, logstd = loc.to(cuda_dev), logstd.to(cuda_dev)
    value = torch.rand(B, A, device=cuda_dev, generator=g) * 2 - 1
    g_lp = torch.randn(B, device=cuda_dev, generator=g) / B
    full = head.head_call(_cabi.HEAD_GRAD, logits, loc, logstd, value, g_lp=g_lp)
    h = B // 2
    lo = head.head_call(_cabi.HEAD_GRAD, logits[:h], loc, logstd, value[:h], g_lp=g_lp[:h])
    hi = head.head_call(_cabi.HEAD_GRAD, logits[h:], loc, logstd, value[h:], g_lp=g_lp[h:])
    assert torch.equal(full["lp"], torch.cat([lo["lp"], hi["lp"]]))
    assert torch.equal(full["dlogits"], torch.cat([lo["dlogits"], hi["dlogits"]]))
    assert rel(full["dloc"], (lo["dloc"].double() + hi["dloc"].double())) < 1e-5
    assert rel(full["dlogstd"], (lo["dlogstd"].double() + hi["dlogstd"].double())) < 1e-5
    # softmax-shift invariance: adding a per-row constant to the logits changes nothing but rounding
    shifted = head.head_call(_cabi.HEAD_FWD, logits + 3.0, loc, logstd, value)
    assert rel(shifted["lp"], full["lp"]) < 1e-5
    assert math.isfinite(float(full["dloc"].abs().sum()))


def test_push_form_single_rank_equals_plain_call(cuda_dev):
    """pfpn_head_logprob_push (the data-parallel form of K1's [2,A,P] output) on a world of one rank, both protocols:
    packets {value, sequence} consumed at once (pfpn_peer_gather_sum_packets), consumed ONE exchange late by the next
    launch's finalize kernel (consume_into) across all rotating slots, and rows + ticket + flags with
    pfpn_peer_gather_sum.  Each must reproduce the plain call's dloc / dlogstd bit for bit."""
    import ctypes as C
    import socket
    import torch.distributed as dist
    from pfpn_b200.head import _stream_ptr, _ws, head_workspace_bytes
    from pfpn_b200.peer import PeerGather
    created = False
    if not dist.is_initialized():
        with socket.socket() as sk:
            sk.bind(("127.0.0.1", 0))
            port = sk.getsockname()[1]
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=cuda_dev)
        created = True
    try:
        B, A, P = 1000, 36, 35
        n = 2 * A * P
        d = {k: (v.to(cuda_dev) if torch.is_tensor(v) else v) for k, v in synth.head_inputs(B, A, P, seed=9).items()}
        cat = lambda r: torch.cat([r["dloc"].reshape(-1), r["dlogstd"].reshape(-1)])
        pg = PeerGather(n, cuda_dev)
        out = torch.empty(n, device=cuda_dev)
        for it in range(5):  # packets, consumed at once
            g_lp = torch.randn(B, device=cuda_dev) / B
            ref = head.head_call(_cabi.HEAD_GRAD, d["logits"], d["loc"], d["logstd"], d["value"], g_lp=g_lp)
            o = head.head_call(_cabi.HEAD_GRAD, d["logits"], d["loc"], d["logstd"], d["value"], g_lp=g_lp, push=pg)
            pg.reduce(out, 1.0, _stream_ptr())
            torch.cuda.synchronize()
            assert torch.equal(o["dloc"], ref["dloc"]) and torch.equal(o["dlogstd"], ref["dlogstd"])
            assert torch.equal(out, cat(ref))
            assert torch.equal(pg.row(pg.pushed, 0), cat(ref))
        refs, outs = [], [torch.full((n,), float("nan"), device=cuda_dev) for _ in range(9)]
        for it in range(9):  # packets, each exchange summed by the NEXT launch (lag 1), scale 0.5
            g_lp = torch.randn(B, device=cuda_dev) / B
            refs.append(cat(head.head_call(_cabi.HEAD_GRAD, d["logits"], d["loc"], d["logstd"], d["value"], g_lp=g_lp)))
            if it == 0:
                head.head_call(_cabi.HEAD_GRAD, d["logits"], d["loc"], d["logstd"], d["value"], g_lp=g_lp, push=pg)
            else:
                with torch.cuda.device(cuda_dev):
                    o = head.head_call(_cabi.HEAD_GRAD, d["logits"], d["loc"], d["logstd"], d["value"], g_lp=g_lp, push=pg,
                                       consume_into=outs[it - 1])
            assert pg.pending == 1
        pg.reduce(outs[8], 1.0, _stream_ptr())
        torch.cuda.synchronize()
        assert pg.pending == 0
        for it in range(9):
            assert torch.equal(outs[it], refs[it]), it
        with pytest.raises(RuntimeError):
            pg.reduce(out)  # nothing pending

        # protocol 0: rows + CTA ticket + flags, raw structure over local buffers
        rows = torch.zeros(2, n, device=cuda_dev)
        flags = torch.zeros(64, dtype=torch.int32, device=cuda_dev)
        ticket = torch.zeros(1, dtype=torch.int32, device=cuda_dev)
        for it in range(1, 5):
            g_lp = torch.randn(B, device=cuda_dev) / B
            ref = head.head_call(_cabi.HEAD_GRAD, d["logits"], d["loc"], d["logstd"], d["value"], g_lp=g_lp)
            hp = _cabi.HeadPush()
            hp.out[0], hp.flags[0], hp.ticket = rows[it & 1].data_ptr(), flags.data_ptr(), ticket.data_ptr()
            hp.nranks, hp.value = 1, it

            class _Raw:
                def push_args(self, consume_into=None):
                    return hp
            o = head.head_call(_cabi.HEAD_GRAD, d["logits"], d["loc"], d["logstd"], d["value"], g_lp=g_lp, push=_Raw())
            _cabi.check(_cabi.pfpn_peer_gather_sum(rows[it & 1].data_ptr(), flags.data_ptr(), 1, it, n, out.data_ptr(), 1.0,
                                                   _stream_ptr()))
            torch.cuda.synchronize()
            assert torch.equal(out, cat(ref)) and torch.equal(cat(o), cat(ref))
            assert int(ticket.item()) == 0 and int(flags[0].item()) == it  # self-resetting CTA counter; monotonic flag
    finally:
        if created:
            dist.destroy_process_group()


def test_autograd_wrapper_entropy_before_log_prob_and_entropy_only(cuda_dev):
    """`.entropy()` before `.log_prob()` takes its own node; after it the two share ONE forward and ONE backward."""
    B, A, P = 130, 36, 35
    d = make(B, A, P, seed=6)
    w = torch.randn(B, generator=torch.Generator().manual_seed(3))
    ref = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], w, torch.full((B,), 0.02))
    for order in ("entropy_first", "entropy_only"):
        lg = d["logits"].to(cuda_dev).requires_grad_(True)
        loc = d["loc"].to(cuda_dev).requires_grad_(True)
        ls = d["logstd"].to(cuda_dev).requires_grad_(True)
        dist = MixtureGaussianDistribution(lg, loc, torch.exp(ls), False, logstd=ls)
        ent = dist.entropy()
        assert rel(ent, ref["ent"]) < TOL
        if order == "entropy_first":
            lp = dist.log_prob(d["value"].to(cuda_dev))
            ((lp * w.to(cuda_dev)).sum() + 0.02 * ent.sum()).backward()
            assert rel(lg.grad, ref["dlogits"]) < TOL and rel(loc.grad, ref["dloc"]) < TOL and rel(ls.grad, ref["dlogstd"]) < TOL
        else:
            (0.02 * ent.sum()).backward()
            only = oh.head_fwd_bwd(d["logits"], d["loc"], d["logstd"], d["value"], torch.zeros(B), torch.full((B,), 0.02))
            assert rel(lg.grad, only["dlogits"]) < TOL


@pytest.mark.parametrize("B", [5000, 40000])
def test_host_pipeline_equals_resident_call(cuda_dev, B):
    """pfpn_b200.host.HostHeadPipeline (pinned host buffers, ramped chunk schedule over three streams) against one resident
    K1 call on the same minibatch: per-state outputs and the logits gradient bit for bit (the advantage statistics are
    whole-batch in both), the [A,P] sums and the loss up to the chunked summation order."""
    from pfpn_b200.host import HostHeadPipeline
    A, P = 36, 35
    d = synth.head_inputs(B, A, P, seed=21)
    dd = {k: (v.to(cuda_dev) if torch.is_tensor(v) else v) for k, v in d.items()}
    fwd = head.head_call(_cabi.HEAD_FWD, dd["logits"], dd["loc"], dd["logstd"], dd["value"])
    lp_old = (fwd["lp"] + dd["lp_noise"]).contiguous()
    stats = head.adv_stats(dd["adv"])
    ref = head.head_call(_cabi.HEAD_PPO, dd["logits"], dd["loc"], dd["logstd"], dd["value"], adv=dd["adv"], lp_old=lp_old,
                         adv_stats_t=stats, loss_scale=1.0 / B)
    pin = lambda t: t.detach().cpu().contiguous().pin_memory()
    pipe = HostHeadPipeline(B, A, P, cuda_dev, chunk=8192)
    assert sum(hi - lo for lo, hi in pipe.bounds) == B and pipe.bounds[0][0] == 0
    out = pipe.run(pin(d["logits"]), pin(d["loc"]), pin(d["logstd"]), pin(d["value"]), pin(d["adv"]), pin(lp_old))
    assert torch.equal(out["lp"], ref["lp"].cpu())
    assert torch.equal(out["ent"], ref["ent"].cpu())
    assert torch.equal(out["dlogits"], ref["dlogits"].cpu())
    assert rel(out["dloc"], ref["dloc"]) < 1e-5 and rel(out["dlogstd"], ref["dlogstd"]) < 1e-5
    assert abs(float(out["loss"]) - float(ref["loss"])) <= 1e-5 * max(1.0, abs(float(ref["loss"])))
