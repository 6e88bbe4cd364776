mkdir -p gpurun_out
timeout 600 python bench.py --log-n 22 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_r1z_2p22.json 2> gpurun_out/bench_r1z_2p22.err
tail -3 gpurun_out/bench_r1z_2p22.err
grep '^{' gpurun_out/bench_r1z_2p22.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['verify'], d['phases_ms_per_step'], d['proof_sha256'], d['standalone'], d['config'])"
