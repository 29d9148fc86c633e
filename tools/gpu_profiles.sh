# ncu evidence of round 2 (one GPU).  Numbers printed by runs under ncu are never bench values.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --dppo-steps 2 > gpurun_out/b_ncu.log 2>&1; echo "launch list rc=$?"
timeout 300 $NCU --set full --import-source on -k regex:head_kernel -s 6 -c 1 -f -o gpurun_out/head_full python tools/time_head.py > gpurun_out/ncu_head.log 2>&1; echo "head rc=$?"
timeout 300 $NCU --set full --import-source on -k regex:sac_head_kernel -s 2 -c 1 -f -o gpurun_out/sac_full python tools/prof_sac.py > gpurun_out/ncu_sac.log 2>&1; echo "sac rc=$?"
timeout 300 $NCU --set full --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/rollout_full python tools/prof_rollout.py > gpurun_out/ncu_rollout.log 2>&1; echo "rollout rc=$?"
timeout 300 $NCU --set full --import-source on -k regex:tc_gemm_pair_kernel -s 16 -c 1 -f -o gpurun_out/tc_full python tools/time_tc_gemm.py > gpurun_out/ncu_tc.log 2>&1; echo "tc rc=$?"
ls -la gpurun_out/*.ncu-rep
