cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
grep -q "pytest rc=0" gpurun_out/pytest_gpu.log || exit 1
timeout 500 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['traffic_source'])
print(json.dumps(d['extra']['rollout_head_B65536_per_gpu']))
print(json.dumps(d['extra']['c5_sac_head']))
print(d['dppo_update']['ms_per_update'], d['dppo_update']['ms_per_update_eager'])
print(d['e2e']['value'])
PY
tail -3 gpurun_out/bench_n1.err
