set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
for N in 8 2; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r1s_n$N.json 2> gpurun_out/bench_r1s_n$N.err; cat gpurun_out/bench_r1s_n$N.json; tail -5 gpurun_out/bench_r1s_n$N.err
done
TP_NO_COSET_SHARD=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_r1s_n8_nocoset.json 2> gpurun_out/bench_r1s_n8_nocoset.err; cat gpurun_out/bench_r1s_n8_nocoset.json; tail -5 gpurun_out/bench_r1s_n8_nocoset.err
