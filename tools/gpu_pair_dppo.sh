cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for rep in 1 2; do
for pr in 1 0; do
  echo "PAIR=$pr"; PFPN_TC_PAIR=$pr timeout 200 python tools/time_dppo.py 2>&1 | tail -2
done; done | tee gpurun_out/pair_dppo.log
