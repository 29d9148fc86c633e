"""Single GPU: K1 step (adv_stats + head + finalize) with no push / packet push to self / push + lagged consume."""
import os, sys, json, socket
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from pfpn_b200 import _cabi, head, synth
from pfpn_b200.peer import PeerGather
dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
with socket.socket() as sk:
    sk.bind(("127.0.0.1", 0)); port = sk.getsockname()[1]
dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=dev)
B, A, P = 65536, 36, 35
g = torch.Generator(device="cuda"); g.manual_seed(1)
lg = torch.randn(B, A, P, device=dev, generator=g)
loc, ls = (x.to(dev) for x in synth.particle_grid(A, P, torch.Generator().manual_seed(0)))
val = torch.rand(B, A, device=dev, generator=g) * 2 - 1
adv = torch.randn(B, device=dev, generator=g); lpo = torch.randn(B, device=dev, generator=g) * 0.1 - 30
stats = torch.zeros(2, device=dev)
pg = PeerGather(2 * A * P, dev); outg = torch.empty(2 * A * P, device=dev)
o = {}
def step(mode):
    _cabi.check(_cabi.pfpn_adv_stats(adv.data_ptr(), B, stats.data_ptr(), head._stream_ptr()))
    kw = {}
    if mode == "push": kw = dict(push=pg)
    if mode == "full": kw = dict(push=pg, consume_into=outg)
    head.head_call(_cabi.HEAD_PPO, lg, loc, ls, val, adv=adv, lp_old=lpo, adv_stats_t=stats, loss_scale=1.0 / B, out=o, **kw)
    if mode == "push": pg.consumed = pg.pushed
res = {}
for rep in range(2):
    for mode in ("none", "push", "full"):
        for _ in range(10): step(mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200): step(mode)
        e1.record(); torch.cuda.synchronize()
        res[f"{mode}_{rep}"] = round(e0.elapsed_time(e1) / 200, 5)
        while pg.pending: pg.reduce(outg)
print(json.dumps(res))
dist.destroy_process_group()
