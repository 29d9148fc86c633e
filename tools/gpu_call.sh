# One-GPU validation: the GPU test suite, smoke(), the bench line.  usage: gpurun -- 'bash tools/gpu_call.sh'
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
grep -q "pytest rc=0" gpurun_out/pytest_gpu.log || exit 1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref.json | cut -c1-400
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_n1.json').read().strip().splitlines() if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['traffic_source'])
print(json.dumps(d['extra']['rollout_head_B65536_per_gpu'])[:300])
print(json.dumps(d['extra']['c5_sac_head'])[-260:])
print(json.dumps(d['extra']['k6_trunk_gemm'])[:330])
print(d['dppo_update']['ms_per_update'], d['dppo_update']['ms_per_update_eager'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'])
PY
tail -3 gpurun_out/bench_n1.err
