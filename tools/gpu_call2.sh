set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
rm -f gpurun_out/time_head_p100.jsonl
timeout 120 python tools/time_head.py > gpurun_out/time_head_default.json 2>&1; cat gpurun_out/time_head_default.json
for v in 0 1 2 3; do
  P=100 PFPN_HEAD_VARIANT=$v timeout 120 python tools/time_head.py >> gpurun_out/time_head_p100.jsonl 2>&1
done
P=100 timeout 120 python tools/time_head.py >> gpurun_out/time_head_p100.jsonl 2>&1
cat gpurun_out/time_head_p100.jsonl
P=100 timeout 300 python -m pytest tests/test_head_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['roofline']['kernel_ms_avg'])
print(d.get('dppo_update'))
PY
tail -5 gpurun_out/bench_c2.err
PFPN_CRITIC_STREAM=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no critic stream:', d.get('dppo_update'))"
