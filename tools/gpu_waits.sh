cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for ns in 0 32 64 128 256 512 1000; do
  PFPN_WAIT_NS=$ns timeout 200 python tools/time_waits.py 2>&1 | tail -1
done | tee gpurun_out/time_waits.log
timeout 200 python tools/time_waits.py 2>&1 | tail -1 | tee -a gpurun_out/time_waits.log
