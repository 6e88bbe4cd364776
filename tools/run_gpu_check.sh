set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 ) 2>&1 | tail -16
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1z.json 2> gpurun_out/bench_r1z.err
tail -3 gpurun_out/bench_r1z.err
cat gpurun_out/bench_r1z.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d.get('verify'), d['phases_ms_per_step'], d.get('standalone'))"
