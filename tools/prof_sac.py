"""ncu target: a few launches of K3f (fused SAC head) at B=65536, A=36, P=100."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfpn_b200 import sampling
from pfpn_b200.network import initial_particles
dev = torch.device("cuda:0"); B, A, P = int(os.environ.get("B", 65536)), 36, 100
g = torch.Generator(device="cuda"); g.manual_seed(1)
logits = torch.randn(B, A, P, device=dev, generator=g) * 2
loc, ls = (x.to(dev) for x in initial_particles(A, P, True))
gs = torch.randn(B, A, device=dev, generator=g); glp = torch.full((B,), 1.0 / B, device=dev)
out = {}
for _ in range(4):
    sampling.sac_head_fused(logits, loc, ls, gs, glp, seed=7, offset=0, out=out)
torch.cuda.synchronize(); print("ok", float(out["logp"].sum()))
