set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_primitives.py tests/test_gpu_prove.py tests/test_gpu_msm_affine.py tests/test_gpu_multirank.py -m gpu -x -q 2>&1 | tail -6
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1z.json 2> gpurun_out/bench_r1z.err
tail -3 gpurun_out/bench_r1z.err
cat gpurun_out/bench_r1z.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d.get('verify'), d['phases_ms_per_step'], d.get('standalone'))"
timeout 300 python -m typlonk_b200.sweep --msm 16,18,22 --ntt 16 --reps 3 2>&1 | tail -6
