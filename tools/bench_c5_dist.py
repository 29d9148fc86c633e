"""BASELINE config c5 across GPUs: the SAC-PFPN head (rsample fwd -> tanh log_prob fwd+bwd -> rsample bwd) over
B_total = 1M states, A = 36, P = 100, states sharded over the ranks of one node; the only exchange is the [2,A,P]
particle-gradient sum (one peer-memory kernel).  Launch: torchrun --nproc-per-node N tools/bench_c5_dist.py [out.json]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pfpn_b200 import _cabi, head, sampling
from pfpn_b200.head import _stream_ptr
from pfpn_b200.learner import shard_bounds
from pfpn_b200.network import initial_particles

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B_total, A, P = int(os.environ.get("B", 1000000)), 36, 100
lo, hi = shard_bounds(B_total, rank, world); B = hi - lo
loc, ls = (t.to(dev) for t in initial_particles(A, P, True))
g = torch.Generator(device="cuda"); g.manual_seed(12831 + rank)
logits = torch.randn(B, A, P, device=dev, generator=g) * 2
g_a = torch.randn(B, A, device=dev, generator=g); g_lp = torch.full((B,), 1.0 / B_total, device=dev)
psum = None
if world > 1:
    from pfpn_b200.peer import PeerSum
    psum = PeerSum(2 * A * P, dev)
tot = torch.empty(2, A, P, device=dev)
hb = {}

def step(it):
    smp, spre, _ = sampling.rsample_fwd(logits, loc, ls, seed=12831, offset=2 * it)
    o = head.head_call(_cabi.HEAD_GRAD, logits, loc, ls, spre, tanh=True, g_lp=g_lp, want_dvalue=True, out=hb)
    _, dc, ds = sampling.rsample_bwd(logits, loc, ls, g_a, o["dvalue"], seed=12831, offset=2 * it)
    if psum is not None:  # dloc / dlogstd of both paths, summed over ranks in one kernel
        slot = psum.slot().view(2, A, P)
        torch.add(dc, o["dloc"], out=slot[0]); torch.add(ds, o["dlogstd"], out=slot[1])
        psum.reduce(tot.view(-1), 1.0, _stream_ptr())

for it in range(3): step(it)
torch.cuda.synchronize()
if world > 1: dist.barrier()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for it in range(n): step(3 + it)
e1.record(); torch.cuda.synchronize()
t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    ms = float(t)
    bytes_state = (4 * A * P + 12 * A) + (8 * A * P + 8 * A) + (8 * A * P + 8 * A + 12)
    res = {"config": "c5 SAC-PFPN head, split (boundary-faithful) variant", "n_gpus": world, "B_total": B_total, "A": A, "P": P,
           "ms_per_step": round(ms, 3), "Mstates_s": round(B_total / ms / 1e3, 2),
           "frac_of_hbm_roofline": round(bytes_state * B_total / world / ms / 1e6 / 6552.0, 3)}
    print(json.dumps(res))
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(json.dumps(res))
if world > 1: dist.destroy_process_group()
