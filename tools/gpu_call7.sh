cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== dbg B=300"; B=300 timeout 40 python tools/dbg_rollout.py 2>&1 | tail -2 || exit 1
B=300 timeout 40 python tools/dbg_rollout.py > /dev/null 2>&1 || { echo "rollout still hangs"; exit 1; }
timeout 200 python -m pytest tests/test_sample_resample_gpu.py -m gpu -x -q -k "fused_rollout" 2>&1 | tail -5
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
grep -q "pytest rc=0" gpurun_out/pytest_gpu.log || exit 1
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c7.json 2> gpurun_out/bench_c7.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c7.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'])
print(json.dumps(d['extra']['rollout_head_B65536_per_gpu'], indent=1))
print(json.dumps(d['extra']['c2_head_B4096_per_gpu']))
print(d['dppo_update']['ms_per_update'], d['dppo_update']['ms_per_update_eager'])
print(d['cpu_baseline'])
PY
tail -3 gpurun_out/bench_c7.err
