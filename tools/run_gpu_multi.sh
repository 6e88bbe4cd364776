set -x
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1z_n8.json 2> gpurun_out/bench_r1z_n8.err
tail -3 gpurun_out/bench_r1z_n8.err
grep '^{' gpurun_out/bench_r1z_n8.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'], d['phases_ms_per_step'], d['proof_sha256'])"
