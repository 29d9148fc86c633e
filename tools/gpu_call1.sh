set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for v in 0 2 3 4; do
  PFPN_HEAD_VARIANT=$v timeout 300 python -m pytest tests/test_head_gpu.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -1
  PFPN_HEAD_VARIANT=$v timeout 120 python tools/time_head.py >> gpurun_out/time_head_variants.jsonl 2>&1
done
cat gpurun_out/time_head_variants.jsonl
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "bench rc=$?"
tail -c 6000 gpurun_out/bench_c1.json; tail -5 gpurun_out/bench_c1.err
