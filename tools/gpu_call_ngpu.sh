# usage: bash tools/gpu_call_ngpu.sh N
set -x
N=${1:-2}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m pytest tests/test_head_gpu.py -m gpu -x -q -k push > gpurun_out/pytest_push_$N.log 2>&1; rc=$?; tail -5 gpurun_out/pytest_push_$N.log
if [ $rc -ne 0 ]; then echo "push test failed rc=$rc"; exit 1; fi
timeout 900 python -m pytest tests/test_peer_multi_gpu.py -m gpu -x -q > gpurun_out/pytest_ngpu_$N.log 2>&1; rc=$?; echo "pytest rc=$rc" >> gpurun_out/pytest_ngpu_$N.log
tail -25 gpurun_out/pytest_ngpu_$N.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/bench_n$N.json').read().strip().splitlines() if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['roofline']['frac'])
    print('xcheck', json.dumps(d.get('xcheck'), indent=1))
    print('dppo', d['dppo_update']['ms_per_update'], d['dppo_update']['ms_per_update_eager'], d['dppo_update']['exchange'])
    print('rank_ms', d.get('rank_ms_per_step'), d.get('exchange_breakdown'))
    print('e2e', d['e2e']['value'])
    print('c5', d['extra']['c5_sac_head'].get('fused_Mstates_s_all_gpus'), d['extra']['c5_sac_head'].get('fused_frac_of_8d_roofline'))
except Exception as e:
    print('parse failed', e)
PY
tail -15 gpurun_out/bench_n$N.err
