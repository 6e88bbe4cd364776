set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2a_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
