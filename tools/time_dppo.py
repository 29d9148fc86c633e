"""Time the DPPO minibatch update (BASELINE c4): trunk + PFPN head + clip/all-reduce/Adam, B_total sharded over ranks."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pfpn_b200.learner import SyncReplicasAdam, shard_bounds
from pfpn_b200.network import ParticleFilteringClipPPONetwork

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B_total = int(os.environ.get("B", 65536)); S, A, P = 197, 36, 35
lo, hi = shard_bounds(B_total, rank, world); B = hi - lo
net = ParticleFilteringClipPPONetwork(True, [S], [A], action_lower_bound=[-1.] * A, action_upper_bound=[1.] * A, particles=P,
                                      resample=-1, resample_interval=368, normalize_state=True, clip_state=5.0,
                                      normalize_advantage=True, device=dev, seed=28949).init()
opt = SyncReplicasAdam(lr=1e-4, norm_clip=1.0)
g = torch.Generator(device="cuda"); g.manual_seed(28949 + rank)
state = torch.randn(B, S, device=dev, generator=g); action = torch.rand(B, A, device=dev, generator=g) * 2 - 1
value = torch.randn(B, device=dev, generator=g); adv = torch.randn(B, device=dev, generator=g)
_, lp, _ = net.run_batch(state); lp_old = lp + 0.05 * torch.randn(B, device=dev, generator=g)
def step():
    net.compute_gradients(state, action, value, lp_old, adv)
    opt.apply_gradients(net)
for _ in range(3): step()
torch.cuda.synchronize()
if world > 1: dist.barrier()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
e0.record(); w0 = time.perf_counter()
for _ in range(n): step()
w1 = time.perf_counter()  # host time to ISSUE the n steps (no sync): launch-bound if close to the device time
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
cpu_issue_ms = (w1 - w0) * 1e3 / n
t = torch.tensor([ms], device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
# ---- the same update captured once in a CUDA graph and replayed ----
from pfpn_b200.learner import GraphedUpdate
gu = GraphedUpdate(net, opt, B, warmup=0)
gu._set(state, action, value, lp_old, adv)
for _ in range(3): gu.run()
torch.cuda.synchronize()
if world > 1: dist.barrier()
e0.record(); w0 = time.perf_counter()
for _ in range(n): gu.run()
w1 = time.perf_counter()
e1.record(); torch.cuda.synchronize()
tg = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
if world > 1: dist.all_reduce(tg, op=dist.ReduceOp.MAX)
graph_issue_ms = (w1 - w0) * 1e3 / n
if rank == 0:
    print(json.dumps({"graph_ms_per_update": round(float(tg), 3), "graph_host_issue_ms": round(graph_issue_ms, 3),
                      "critic_stream": os.environ.get("PFPN_CRITIC_STREAM", "1")}))
    flops = 12.6e6 * B_total
    print(json.dumps({"world": world, "B_total": B_total, "ms_per_update": round(float(t), 3), "samples_per_s": round(B_total / float(t) * 1e3),
                      "trunk_TFLOPs": round(flops / float(t) / 1e9, 1), "host_issue_ms_per_update": round(cpu_issue_ms, 3), "params_equal_hash": float(net.params.double().sum())}))
if world > 1: dist.destroy_process_group()
