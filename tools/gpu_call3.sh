set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
rm -f gpurun_out/time_dppo_sweep.jsonl
for B in 8192 16384 65536; do
  for cfg in "1 1" "0 1" "1 0"; do
    set -- $cfg
    echo "{\"B\": $B, \"PRESPLIT\": $1, \"CRITIC_STREAM\": $2}" >> gpurun_out/time_dppo_sweep.jsonl
    B=$B PFPN_PRESPLIT=$1 PFPN_CRITIC_STREAM=$2 timeout 200 python tools/time_dppo.py >> gpurun_out/time_dppo_sweep.jsonl 2>&1
  done
done
cat gpurun_out/time_dppo_sweep.jsonl
