set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
rm -f gpurun_out/time_dppo_sweep2.jsonl
for B in 8192 65536; do
  for w in 1 0; do
    echo "{\"B\": $B, \"WGRAD_STREAM\": $w}" >> gpurun_out/time_dppo_sweep2.jsonl
    B=$B PFPN_WGRAD_STREAM=$w PFPN_GRAPH_MAX_BATCH=1000000 timeout 200 python tools/time_dppo.py >> gpurun_out/time_dppo_sweep2.jsonl 2>&1
  done
done
cat gpurun_out/time_dppo_sweep2.jsonl
