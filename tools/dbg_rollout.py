import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pfpn_b200 import sampling, synth
B = int(os.environ.get("B", 8)); A, P = 36, 35
dev = torch.device("cuda:0")
d = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in synth.head_inputs(B, A, P, seed=3).items()}
stats = os.environ.get("STATS", "1") == "1"
mx, sm = torch.zeros(A, P, device=dev), torch.zeros(A, P, device=dev)
out = sampling.rollout_fused(d["logits"], d["loc"], d["logstd"], seed=1, offset=2, max_active=mx if stats else None,
                             sum_active=sm if stats else None, want_ent=True)
torch.cuda.synchronize()
print("ok", B, float(out["lp"].sum()), int(out["idx"].sum()), float(sm.sum()))
