# compute-sanitizer memcheck over small parity tests of the round-2 kernels (slow: keep the selection small)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TOOL=${TOOL:-memcheck}
CS="compute-sanitizer --tool $TOOL --error-exitcode 9 --print-limit 5"
timeout 240 $CS python -m pytest tests/test_sample_resample_gpu.py -m gpu -x -q -k "fused_rollout_matches_oracle or fused_sac_head_matches_oracle" > gpurun_out/san_fused.log 2>&1; echo "fused rc=$?"; tail -4 gpurun_out/san_fused.log
timeout 240 $CS python -m pytest tests/test_head_gpu.py -m gpu -x -q -k "push_form or host_pipeline" > gpurun_out/san_head.log 2>&1; echo "head rc=$?"; tail -4 gpurun_out/san_head.log
timeout 240 $CS python -m pytest tests/test_tc_gemm_gpu.py -m gpu -x -q -k "presplit" > gpurun_out/san_gemm.log 2>&1; echo "gemm rc=$?"; tail -4 gpurun_out/san_gemm.log
