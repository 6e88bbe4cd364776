set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_prove.py tests/test_gpu_multirank.py tests/test_gpu_verify.py -m gpu -x -q 2>&1 | tail -12
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1y.json 2> gpurun_out/bench_r1y.err
tail -3 gpurun_out/bench_r1y.err
cat gpurun_out/bench_r1y.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d.get('verify'), d['phases_ms_per_step'], d.get('standalone'))"
