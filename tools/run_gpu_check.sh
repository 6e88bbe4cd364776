set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1x.json 2> gpurun_out/bench_r1x.err
cat gpurun_out/bench_r1x.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['verify'], d['phases_ms_per_step'])"
export TYPLONK_B200_LIB=$PWD/typlonk_b200/lib/libtyplonk_b200_polycall.so
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('polycall', d['value'], d['phases_ms_per_step'])"
