cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sac_gpu.py -m gpu -x -q > gpurun_out/pytest_sac.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_sac.log
timeout 600 python tools/time_sac.py 2>&1 | tail -1 | tee gpurun_out/sac_step.json
