cd $GRAFT_REPO_ROOT
for B in 12 16 32 64; do echo "== B=$B"; B=$B timeout 60 python tools/dbg_rollout.py 2>&1 | grep -v "^$" | sort | uniq -c | sort -rn | head -8; done
