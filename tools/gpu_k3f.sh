cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_sample_resample_gpu.py tests/test_sac_gpu.py -m gpu -x -q -k "fused_sac or fused" > gpurun_out/pytest_k3f.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_k3f.log
timeout 200 python tools/time_waits.py 2>&1 | tail -1
