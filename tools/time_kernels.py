"""Time K3f, K2f and the DPPO update (optionally under one setting of PFPN_WAIT_NS, the producers' back-off); one JSON line."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfpn_b200 import sampling, synth
from pfpn_b200.network import initial_particles, ParticleFilteringClipPPONetwork
from pfpn_b200.learner import SyncReplicasAdam
dev = torch.device("cuda:0"); B, A = 65536, 36
g = torch.Generator(device="cuda"); g.manual_seed(1)

def timeit(fn, n=20, warm=5):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

res = {"wait_ns": os.environ.get("PFPN_WAIT_NS", "default")}
P = 100
lg = [torch.randn(B, A, P, device=dev, generator=g) * 2 for _ in range(3)]
loc, ls = (x.to(dev) for x in initial_particles(A, P, True))
gs = torch.randn(B, A, device=dev, generator=g); glp = torch.full((B,), 1.0 / B, device=dev)
out = {}; i = [0]
def k3f():
    i[0] += 1
    sampling.sac_head_fused(lg[i[0] % 3], loc, ls, gs, glp, seed=7, offset=i[0], out=out)
res["k3f_ms"] = round(timeit(k3f), 4)
del lg
P = 35
lg = [torch.randn(B, A, P, device=dev, generator=g) * 2 for _ in range(6)]
loc, ls = (x.to(dev) for x in synth.particle_grid(A, P, torch.Generator().manual_seed(0)))
mx, sm = torch.zeros(A, P, device=dev), torch.zeros(A, P, device=dev)
def k2f():
    i[0] += 1
    sampling.rollout_fused(lg[i[0] % 6], loc, ls, seed=1, offset=i[0], max_active=mx, sum_active=sm)
res["k2f_ms"] = round(timeit(k2f), 4)
S = 197
net = ParticleFilteringClipPPONetwork(True, [S], [A], action_lower_bound=[-1.] * A, action_upper_bound=[1.] * A, particles=P,
                                      resample=-1, resample_interval=368, normalize_state=True, clip_state=5.0,
                                      normalize_advantage=True, device=dev, seed=28949).init()
opt = SyncReplicasAdam(lr=1e-4, norm_clip=1.0)
state = torch.randn(B, S, device=dev, generator=g); action = torch.rand(B, A, device=dev, generator=g) * 2 - 1
value = torch.randn(B, device=dev, generator=g); adv = torch.randn(B, device=dev, generator=g)
_, lp, _ = net.run_batch(state); lp_old = lp + 0.05 * torch.randn(B, device=dev, generator=g)
def upd():
    net.compute_gradients(state, action, value, lp_old, adv)
    opt.apply_gradients(net)
res["dppo_ms"] = round(timeit(upd, n=10, warm=3), 4)
res["run_batch_ms"] = round(timeit(lambda: net.run_batch(state), n=10, warm=3), 4)
print(json.dumps(res))
