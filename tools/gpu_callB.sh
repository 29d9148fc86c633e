cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_tc_gemm_gpu.py -m gpu -x -q 2>&1 | tail -4
[ ${PIPESTATUS[0]} -eq 0 ] || { echo "tc tests failed/hung"; exit 1; }
timeout 200 python -m pytest tests/test_learner_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== persistent"; timeout 120 python tools/time_tc_gemm.py 2>&1 | cut -c1-200
echo "== classic"; PFPN_TC_PERSISTENT=0 timeout 120 python tools/time_tc_gemm.py 2>&1 | cut -c1-200
for B in 8192 65536; do
  for pz in 1 0; do echo "== dppo B=$B persistent=$pz"; B=$B PFPN_TC_PERSISTENT=$pz timeout 200 python tools/time_dppo.py 2>&1 | tail -2 | cut -c1-250; done
done
for v in 0 2; do echo "== K1 variant $v"; PFPN_HEAD_VARIANT=$v timeout 100 python tools/time_head.py 2>&1 | cut -c1-330; done
PFPN_HEAD_VARIANT=2 timeout 200 python -m pytest tests/test_head_gpu.py -m gpu -x -q 2>&1 | tail -2
