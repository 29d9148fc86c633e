"""Secondary measurements for the other BASELINE.json configs (the headline line is bench.py):
  c2  head fwd+bwd at B=4096 (L2-resident: a latency case), fwd-only, sample
  c3  resampler sweep P = 10/35/100 (A=36, H=512): microseconds per call
  c5  SAC head sweep: rsample fwd + bwd and tanh log_prob fwd+bwd at A=36, P=100, B up to 1M,
      against the HBM roofline of the split (two-launch, boundary-faithful) variant.
Writes one JSON object to stdout (and to the path given as argv[1])."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pfpn_b200 import _cabi, head, resampling, sampling, synth

dev = torch.device("cuda:0")
PEAK = 6552.0
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass

def timeit(fn, n=50, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        fn(); ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    return ts[n // 2]

out = {"peak_hbm_gbs": PEAK}
# ---------------------------------------------------------------- c2
B, A, P = 4096, 36, 35
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synth.head_inputs(B, A, P).items()}
o = head.head_call(_cabi.HEAD_FWD, d["logits"], d["loc"], d["logstd"], d["value"])
lp_old = o["lp"] + d["lp_noise"]; stats = head.adv_stats(d["adv"]); buf = {}
ms = timeit(lambda: head.head_call(_cabi.HEAD_PPO, d["logits"], d["loc"], d["logstd"], d["value"], adv=d["adv"], lp_old=lp_old, adv_stats_t=stats, out=buf))
ms_f = timeit(lambda: head.head_call(_cabi.HEAD_FWD, d["logits"], d["loc"], d["logstd"], d["value"], out=buf))
ms_s = timeit(lambda: sampling.sample_plain(d["logits"], d["loc"], d["logstd"], seed=1, offset=2))
out["c2_head_B4096"] = {"note": "41.9 MB working set is L2-resident: latency case, not a roofline case", "ppo_fwd_bwd_us": ms * 1e3,
                        "Mstates_s": B / ms / 1e3, "fwd_only_us": ms_f * 1e3, "plain_sample_us": ms_s * 1e3}
# ---------------------------------------------------------------- c3
res = {}
for P3 in (10, 35, 100):
    rng = np.random.default_rng(33406 + P3); A3, H = 36, 512
    lg = rng.normal(0, 3, (512, A3, P3)); pr = np.exp(lg - lg.max(-1, keepdims=True)); pr /= pr.sum(-1, keepdims=True)
    mx, sm = pr.max(0).astype(np.float32), pr.sum(0).astype(np.float32)
    dead = rng.random((A3, P3)) < 0.1; mx[dead] = 1e-6
    t = lambda a: torch.tensor(a, dtype=torch.float32, device=dev)
    base = dict(mx=t(mx), sm=t(sm), loc=t(np.linspace(-1, 1, P3)[None].repeat(A3, 0)), ls=t(np.full((A3, P3), np.log(2 / (P3 - 1)))),
                b=t(rng.normal(0, 1, A3 * P3)), W=t(rng.normal(0, .01, (H, A3 * P3))))
    work = {k: v.clone() for k, v in base.items()}
    def call():
        for k in work: work[k].copy_(base[k])
        resampling.resample_(work["mx"], work["sm"], work["loc"], work["ls"], work["b"], work["W"], seed=3, offset=5)
    def copies():
        for k in work: work[k].copy_(base[k])
    us = (timeit(call, 30) - timeit(copies, 30)) * 1e3
    res[f"P{P3}"] = {"dead": int(dead.sum()), "us_per_resample": us, "bytes_touched_max": 4 * (2 * H * A3 * P3 + 8 * A3 * P3)}
out["c3_resample_A36_H512"] = dict(res, note="latency-bound (one plan CTA + two column-copy launches); roofline fraction not meaningful")
# ---------------------------------------------------------------- c5
A5, P5 = 36, 100
from pfpn_b200.network import initial_particles
loc5, ls5 = (t.to(dev) for t in initial_particles(A5, P5, True))
sweep = {}
for B5 in (65536, 262144, 1000000):
    g = torch.Generator(device="cuda"); g.manual_seed(12831)
    logits = torch.randn(B5, A5, P5, device=dev, generator=g) * 2
    g_a = torch.randn(B5, A5, device=dev, generator=g); g_lp = torch.full((B5,), 1.0 / B5, device=dev)
    smp, spre, _ = sampling.rsample_fwd(logits, loc5, ls5, seed=12831, offset=1)
    t_f = timeit(lambda: sampling.rsample_fwd(logits, loc5, ls5, seed=12831, offset=1), 10, 2)
    t_b = timeit(lambda: sampling.rsample_bwd(logits, loc5, ls5, g_a, None, seed=12831, offset=1), 10, 2)
    hb = {}
    t_l = timeit(lambda: head.head_call(_cabi.HEAD_GRAD, logits, loc5, ls5, spre, tanh=True, g_lp=g_lp, want_dvalue=True, out=hb), 10, 2)
    AP = A5 * P5
    bytes_f, bytes_b, bytes_l = 4 * AP + 12 * A5, 8 * AP + 8 * A5, 8 * AP + 8 * A5 + 12
    sweep[f"B{B5}"] = {"rsample_fwd_ms": t_f, "rsample_bwd_ms": t_b, "logprob_tanh_fwd_bwd_ms": t_l,
                       "rsample_fwd_frac_hbm": bytes_f * B5 / t_f / 1e6 / PEAK, "rsample_bwd_frac_hbm": bytes_b * B5 / t_b / 1e6 / PEAK,
                       "logprob_frac_hbm": bytes_l * B5 / t_l / 1e6 / PEAK,
                       "sac_head_Mstates_s": B5 / (t_f + t_b + t_l) / 1e3}
    del logits, g_a, smp, spre, hb
    torch.cuda.empty_cache()
out["c5_sac_head_A36_P100"] = dict(sweep, note="boundary-faithful split variant: rsample fwd, then log_prob fwd+bwd (K1, tanh, dvalue), then rsample bwd; "
                                   "Philox draws regenerated in bwd; algorithmic bytes per state: fwd 4AP+12A, bwd 8AP+8A, log_prob 8AP+8A+12")
js = json.dumps(out, indent=1)
print(js)
if len(sys.argv) > 1:
    open(sys.argv[1], "w").write(js)
