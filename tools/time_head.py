"""Tuning aid: time K1 (PPO fwd+bwd and FWD-only) for the current PFPN_HEAD_VARIANT."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfpn_b200 import _cabi, head, synth

B = int(os.environ.get("B", 65536)); A = int(os.environ.get("A", 36)); P = int(os.environ.get("P", 35))
dev = torch.device("cuda:0")
g = torch.Generator(device="cuda"); g.manual_seed(1)
logits = torch.randn(B, A, P, device=dev, generator=g) * 2
loc, logstd = synth.particle_grid(A, P, torch.Generator().manual_seed(0))
loc, logstd = loc.to(dev), logstd.to(dev)
value = torch.rand(B, A, device=dev, generator=g) * 2 - 1
adv = torch.randn(B, device=dev, generator=g)
out = head.head_call(_cabi.HEAD_FWD, logits, loc, logstd, value)
lp_old = out["lp"] + 0.05 * torch.randn(B, device=dev, generator=g)
stats = head.adv_stats(adv)
res = {"variant": os.environ.get("PFPN_HEAD_VARIANT", "0"), "B": B, "A": A, "P": P}
g_lp = torch.randn(B, device=dev, generator=g)
for name, mode in (("ppo", _cabi.HEAD_PPO), ("fwd", _cabi.HEAD_FWD), ("grad", _cabi.HEAD_GRAD), ("sac", _cabi.HEAD_GRAD)):
    kw = dict(adv=adv, lp_old=lp_old, adv_stats_t=stats) if mode == _cabi.HEAD_PPO else {}
    if name == "grad":
        kw = dict(g_lp=g_lp)
    if name == "sac":  # tanh-squashed value (the pre-tanh sample), gradient w.r.t. the value as well
        kw = dict(g_lp=g_lp, tanh=True, want_dvalue=True)
    o = {}
    for _ in range(5):
        head.head_call(mode, logits, loc, logstd, value, out=o, **kw)
    torch.cuda.synchronize()
    n = 50
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        head.head_call(mode, logits, loc, logstd, value, out=o, **kw)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(n))
    med = ts[n // 2]
    bytes_ = ((8 if mode else 4) * A * P + 4 * A + 16) * B
    res[name] = {"ms_med": round(med, 4), "ms_min": round(ts[0], 4), "GBps": round(bytes_ / med / 1e6, 1),
                 "Mstates_s": round(B / med / 1e3, 1)}
res["info"] = head.launch_info(A, P, _cabi.HEAD_PPO)
print(json.dumps(res))
