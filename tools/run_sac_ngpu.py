"""torchrun --nproc-per-node N: SAC learner step data-parallel -- every rank its own replay minibatch, clipped gradients and
normaliser statistics averaged (SyncReplicasOptimizer semantics), identical Adam on every rank: replicas must stay equal
and equal to a single-process run over the same shards averaged by hand."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pfpn_b200.sac import ParticleFilteringSACNetwork, SACOptimizer

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
S, A, P, B = 197, 36, 35, 128
net = ParticleFilteringSACNetwork(True, [S], [A], action_lower_bound=[-1.] * A, action_upper_bound=[1.] * A, particles=P,
                                  resample=-1, resample_interval=3, normalize_state=True, clip_state=5.0, device=dev, seed=5).init()
opt = SACOptimizer()
g = torch.Generator(device="cuda"); g.manual_seed(100 + rank)
for it in range(4):  # crosses a resample tick (interval 3)
    batch = (torch.randn(B, S, device=dev, generator=g), torch.rand(B, A, device=dev, generator=g) * 2 - 1,
             torch.randn(B, device=dev, generator=g), torch.ones(B, device=dev), torch.randn(B, S, device=dev, generator=g))
    net.compute_gradients(*batch)
    opt.apply_gradients(net)
    net.run_batch(batch[0])  # rollout-side statistics differ per rank; they are local state until pushed
    net.max_active[0, :3] = 0.0  # dead particles on every rank -> the tick at step 3 really resamples (from averaged statistics)
torch.cuda.synchronize()
ref = net.params.clone(); dist.broadcast(ref, 0)
reft = net.target_params.clone(); dist.broadcast(reft, 0)
refm = net.state_mean.clone(); dist.broadcast(refm, 0)
os.write(1, (json.dumps({"rank": rank, "world": world, "params_equal": bool(torch.equal(ref, net.params)),
                  "target_equal": bool(torch.equal(reft, net.target_params)), "state_mean_equal": bool(torch.equal(refm, net.state_mean)),
                  "finite": bool(torch.isfinite(net.params).all()), "loc_row0": [round(float(v), 4) for v in net.loc[0, :4]], "log_alpha": float(net.log_alpha)}) + "\n").encode())
dist.barrier(); dist.destroy_process_group()
