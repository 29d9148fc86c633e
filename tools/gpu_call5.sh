set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dppo > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c5.json').read().strip().splitlines()[-1])
print(json.dumps(d['extra']['c5_sac_head'], indent=1))
PY
tail -3 gpurun_out/bench_c5.err
