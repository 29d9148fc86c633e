cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
PFPN_TC_PAIR=1 timeout 120 python tools/dbg_pair.py > gpurun_out/pair1.log 2>&1; echo "pair=1 rc=$?"; tail -8 gpurun_out/pair1.log
PFPN_TC_PAIR=0 timeout 120 python tools/dbg_pair.py > gpurun_out/pair0.log 2>&1; echo "pair=0 rc=$?"; tail -3 gpurun_out/pair0.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
