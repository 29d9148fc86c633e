cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
B=300 timeout 40 python tools/dbg_rollout.py 2>&1 | tail -1
B=300 timeout 40 python tools/dbg_rollout.py > /dev/null 2>&1 || { echo "rollout hangs"; exit 1; }
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
grep -q "pytest rc=0" gpurun_out/pytest_gpu.log || exit 1
timeout 100 python tools/prof_rollout.py
python - <<'PY'
import torch, time, sys, os
sys.path.insert(0, os.getcwd())
from pfpn_b200 import sampling, synth
dev = torch.device("cuda:0"); B, A, P = 65536, 36, 35
g = torch.Generator(device="cuda"); g.manual_seed(1)
logits = torch.randn(B, A, P, device=dev, generator=g) * 2
loc, ls = (x.to(dev) for x in synth.particle_grid(A, P, torch.Generator().manual_seed(0)))
mx, sm = torch.zeros(A, P, device=dev), torch.zeros(A, P, device=dev)
for _ in range(3): sampling.rollout_fused(logits, loc, ls, seed=1, offset=2, max_active=mx, sum_active=sm)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): sampling.rollout_fused(logits, loc, ls, seed=1, offset=2, max_active=mx, sum_active=sm)
e1.record(); torch.cuda.synchronize()
print("K2f us:", e0.elapsed_time(e1) / 20 * 1e3)
PY
bash tools/gpu_profiles.sh
