set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1j.json 2> gpurun_out/bench_r1j.err; cat gpurun_out/bench_r1j.json; tail -3 gpurun_out/bench_r1j.err
