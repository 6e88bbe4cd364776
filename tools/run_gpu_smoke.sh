mkdir -p gpurun_out
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1z2.json 2> gpurun_out/bench_r1z2.err
tail -2 gpurun_out/bench_r1z2.err
grep '^{' gpurun_out/bench_r1z2.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['proof_sha256'], d['verify'])"
timeout 100 python -m pytest tests/test_gpu_prove.py tests/test_gpu_multirank.py tests/test_cpp_mirror.py -m gpu -x -q 2>&1 | tail -2
