"""torchrun --nproc-per-node N: fused peer-memory all-reduce + Adam (csrc/comm.cu) vs the NCCL path."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pfpn_b200.learner import SyncReplicasAdam
from pfpn_b200.network import ParticleFilteringClipPPONetwork

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
S, A, P, B = 197, 36, 35, 2048

def make():
    net = ParticleFilteringClipPPONetwork(True, [S], [A], action_lower_bound=[-1.] * A, action_upper_bound=[1.] * A, particles=P,
                                          resample=-1, resample_interval=3, normalize_state=True, clip_state=5.0,
                                          normalize_advantage=True, device=dev, seed=28949).init()
    return net

g = torch.Generator(device="cuda"); g.manual_seed(100 + rank)
state = torch.randn(B, S, device=dev, generator=g); action = torch.rand(B, A, device=dev, generator=g) * 2 - 1
value = torch.randn(B, device=dev, generator=g); adv = torch.randn(B, device=dev, generator=g)
res = {}
for name, fused in (("nccl", False), ("fused", True)):
    net, opt = make(), SyncReplicasAdam(lr=1e-4, norm_clip=1.0, fused_peer=fused)
    _, lp, _ = net.run_batch(state); lp_old = lp.clone()
    for it in range(4):  # crosses a resample tick (interval 3): replicas must stay identical
        net.compute_gradients(state, action, value, lp_old, adv)
        opt.apply_gradients(net)
    torch.cuda.synchronize()
    snap = dict(params=net.params.clone(), mean=net.state_mean.clone(), maxa=net.max_active.clone())
    # timing of the exchange + Adam alone (repeated applies on a stale bucket: timing only)
    net.compute_gradients(state, action, value, lp_old, adv)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): opt.apply_gradients(net)
    e1.record(); torch.cuda.synchronize()
    res[name] = dict(ms=e0.elapsed_time(e1) / 20, **snap)
# replicas identical across ranks?
for name in res:
    ref = res[name]["params"].clone(); dist.broadcast(ref, 0)
    res[name]["replica_equal"] = bool(torch.equal(ref, res[name]["params"]))
d = (res["fused"]["params"] - res["nccl"]["params"]).abs().max().item()
out = dict(rank=rank, world=world, max_abs_param_diff_fused_vs_nccl=d, replica_equal=(res["nccl"]["replica_equal"], res["fused"]["replica_equal"]),
           stats_equal=bool(torch.allclose(res["fused"]["mean"], res["nccl"]["mean"], rtol=1e-6, atol=1e-7)),
           ms_apply_nccl=round(res["nccl"]["ms"], 4), ms_apply_fused=round(res["fused"]["ms"], 4))
# ---- the whole update replayed as a CUDA graph on every rank (device-resident counters, exchange inside the graph) ----
from pfpn_b200.learner import GraphedUpdate
snaps = {}
for name, graphed in (("eager", False), ("graph", True)):
    net, opt = make(), SyncReplicasAdam(lr=1e-4, norm_clip=1.0, fused_peer=True)
    _, lp, _ = net.run_batch(state); lp_old = lp.clone()
    gu = GraphedUpdate(net, opt, B, warmup=2 if graphed else 10 ** 9)
    for it in range(6):  # crosses resample ticks (interval 3), both staging parities
        gu.run(state, action, value, lp_old, adv * (1.0 + 0.1 * it))
    torch.cuda.synchronize()
    snaps[name] = (net.params.clone(), net.state_mean.clone(), gu.replays)
refp = snaps["graph"][0].clone(); dist.broadcast(refp, 0)
out.update(graph_replays=snaps["graph"][2], graph_equals_eager=bool(torch.equal(snaps["graph"][0], snaps["eager"][0]) and torch.equal(snaps["graph"][1], snaps["eager"][1])),
           graph_replicas_equal=bool(torch.equal(refp, snaps["graph"][0])))
# ---- the sharded head's [2, A, P] exchange in one kernel (pfpn_peer_allreduce_sum) vs NCCL ----------------
from pfpn_b200.peer import PeerSum
from pfpn_b200.head import _stream_ptr
n = 2 * A * P
ps = PeerSum(n, dev)
ok_sum, t_peer, t_nccl = True, 0.0, 0.0
outp = torch.empty(n, device=dev)
for it in range(6):
    src = torch.randn(n, device=dev, generator=g)
    ps.slot().copy_(src)
    ps.reduce(outp, 1.0, _stream_ptr())
    ref = src.clone(); dist.all_reduce(ref)
    ok_sum = ok_sum and bool(torch.allclose(outp, ref, rtol=1e-6, atol=1e-6))
    chk = outp.clone(); dist.broadcast(chk, 0)
    ok_sum = ok_sum and bool(torch.equal(chk, outp))  # rank-ordered sum: bit-identical on every rank
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): ps.reduce(outp, 1.0, _stream_ptr())
e1.record(); torch.cuda.synchronize(); t_peer = e0.elapsed_time(e1) / 50
buf = torch.randn(n, device=dev)
dist.barrier(); e0.record()
for _ in range(50): dist.all_reduce(buf)
e1.record(); torch.cuda.synchronize(); t_nccl = e0.elapsed_time(e1) / 50
out.update(small_sum_ok=ok_sum, ms_small_sum_peer=round(t_peer, 4), ms_small_sum_nccl=round(t_nccl, 4))
# ---- the PUSH form: K1's finalize kernel stores dloc / dlogstd into every rank's gather row (pfpn_head_logprob_push),
#      pfpn_peer_gather_sum sums the local rows in rank order; vs NCCL on the same per-rank contributions ----------------
from pfpn_b200 import _cabi, head
from pfpn_b200.peer import PeerGather
Bh = 4096 + 64 * rank  # ragged shards
lg = torch.randn(Bh, A, P, device=dev, generator=g) * 2
locp, lsp = make().loc.clone(), make().logstd.clone()
val = torch.rand(Bh, A, device=dev, generator=g) * 2 - 1
glp = torch.randn(Bh, device=dev, generator=g) / Bh
pg = PeerGather(n, dev)
ok_push, outg = True, torch.empty(n, device=dev)
for it in range(5):
    glp = glp * (1.0 + 0.1 * it)
    o = head.head_call(_cabi.HEAD_GRAD, lg, locp, lsp, val, g_lp=glp, push=pg)
    pg.reduce(outg, 1.0, _stream_ptr())
    mine = torch.cat([o["dloc"].reshape(-1), o["dlogstd"].reshape(-1)])
    allc = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allc, mine)
    ordered = allc[0].clone()
    for r in range(1, world):
        ordered += allc[r]
    ok_push = ok_push and bool(torch.equal(outg, ordered))  # bit-equal to the rank-ordered sum, on every rank
# the same, each exchange summed ONE exchange late by the next launch's finalize kernel (consume_into), 4 rotating slots
ok_async, outs = True, [torch.empty(n, device=dev) for _ in range(8)]
exp = []
for it in range(8):
    glp2 = glp * (1.0 + 0.05 * it)
    o2 = head.head_call(_cabi.HEAD_GRAD, lg, locp, lsp, val, g_lp=glp2, push=pg, consume_into=outs[it - 1] if it else None)
    exp.append(torch.cat([o2["dloc"].reshape(-1), o2["dlogstd"].reshape(-1)]))
pg.reduce(outs[7], 1.0, _stream_ptr())
torch.cuda.synchronize()
for it in range(8):
    allc = [torch.empty_like(exp[it]) for _ in range(world)]
    dist.all_gather(allc, exp[it])
    ordered = allc[0].clone()
    for r in range(1, world):
        ordered += allc[r]
    ok_async = ok_async and bool(torch.equal(outs[it], ordered))
out.update(push_async_ok=ok_async)
torch.cuda.synchronize(); dist.barrier()
e0.record()
for _ in range(30):
    head.head_call(_cabi.HEAD_GRAD, lg, locp, lsp, val, g_lp=glp, push=pg, out=o, consume_into=outg)
pg.reduce(outg, 1.0, _stream_ptr())
e1.record(); torch.cuda.synchronize(); t_push = e0.elapsed_time(e1) / 30
e0.record()
for _ in range(30):
    head.head_call(_cabi.HEAD_GRAD, lg, locp, lsp, val, g_lp=glp, out=o)
e1.record(); torch.cuda.synchronize(); t_nopush = e0.elapsed_time(e1) / 30
out.update(push_sum_ok=ok_push, ms_head_with_push_exchange=round(t_push, 4), ms_head_alone=round(t_nopush, 4))
os.write(1, (json.dumps(out) + "\n").encode())  # one write: the ranks' lines never interleave
dist.barrier(); dist.destroy_process_group()
