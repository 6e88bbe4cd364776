set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -12 | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1r.json 2> gpurun_out/bench_r1r.err; cat gpurun_out/bench_r1r.json; tail -3 gpurun_out/bench_r1r.err
ls -la gpurun_out | tail -3
