"""Time the SAC-PFPN learner step (replay sample -> compute_gradients -> joint clip + two Adams -> target sync)."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfpn_b200.sac import ParticleFilteringSACNetwork, SACOptimizer, ReplayRing, GraphedSACUpdate
dev = torch.device("cuda:0")
S, A = 197, 36
res = []
for P, B in ((35, 256), (100, 256), (100, 4096), (100, 65536)):
    net = ParticleFilteringSACNetwork(True, [S], [A], action_lower_bound=[-1.] * A, action_upper_bound=[1.] * A, particles=P,
                                      resample=-1, resample_interval=12000, normalize_state=True, clip_state=5.0, device=dev, seed=1).init()
    opt = SACOptimizer()
    ring = ReplayRing(1_000_000 if B <= 4096 else 200_000, S, A, device=dev, seed=2)
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    n0 = 100_000
    ring.append(torch.randn(n0, S, device=dev, generator=g), torch.rand(n0, A, device=dev, generator=g) * 2 - 1,
                torch.randn(n0, device=dev, generator=g), torch.ones(n0, device=dev), torch.randn(n0, S, device=dev, generator=g))
    def step():
        net.compute_gradients(*ring.sample(B))
        opt.apply_gradients(net)
    for _ in range(3): step()
    torch.cuda.synchronize()
    n = 20 if B <= 4096 else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter(); e0.record()
    for _ in range(n): step()
    e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
    row = {"P": P, "batch": B, "ms_per_step_device": round(e0.elapsed_time(e1) / n, 3), "ms_per_step_wall": round((w1 - w0) * 1e3 / n, 3),
           "samples_per_s": round(B / (e0.elapsed_time(e1) / n) * 1e3)}
    if B <= 4096:  # the same step replayed as one CUDA graph (replay sample eager: a gather into the graph's input buffers)
        gu = GraphedSACUpdate(net, opt, B, warmup=1)
        def gstep():
            gu.run(*ring.sample(B))
        for _ in range(4): gstep()
        torch.cuda.synchronize()
        w0 = time.perf_counter(); e0.record()
        for _ in range(n): gstep()
        e1.record(); torch.cuda.synchronize(); w1 = time.perf_counter()
        row.update(graph_ms_per_step_device=round(e0.elapsed_time(e1) / n, 3), graph_ms_per_step_wall=round((w1 - w0) * 1e3 / n, 3),
                   graph_samples_per_s=round(B / (e0.elapsed_time(e1) / n) * 1e3), graph_replays=gu.replays)
    res.append(row)
    del net, opt, ring
    torch.cuda.empty_cache()
print(json.dumps({"sac_learner_step": res, "note": "B=256 is the reference's batch_size (deepmimic_sac_base.py:8): launch/host-bound"}))
