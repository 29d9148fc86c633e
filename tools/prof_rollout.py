"""ncu target: a few launches of K2f (fused rollout head) at B=65536, A=36, P=35."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfpn_b200 import sampling, synth
dev = torch.device("cuda:0"); B, A, P = int(os.environ.get("B", 65536)), 36, 35
g = torch.Generator(device="cuda"); g.manual_seed(1)
logits = torch.randn(B, A, P, device=dev, generator=g) * 2
loc, ls = (x.to(dev) for x in synth.particle_grid(A, P, torch.Generator().manual_seed(0)))
mx, sm = torch.zeros(A, P, device=dev), torch.zeros(A, P, device=dev)
for _ in range(4):
    out = sampling.rollout_fused(logits, loc, ls, seed=1, offset=2, max_active=mx, sum_active=sm)
torch.cuda.synchronize(); print("ok", float(out["lp"].sum()))
