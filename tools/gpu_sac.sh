cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sac_gpu.py tests/test_sample_resample_gpu.py -m gpu -x -q > gpurun_out/pytest_sac.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_sac.log
