set -x
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 3 --warmup 3 --log-n 22 --no-cpu-baseline --dump-proof gpurun_out/proof_2p22_n8.bin > gpurun_out/bench_r1v_2p22_n8.json 2> gpurun_out/bench_r1v_2p22_n8.err; cat gpurun_out/bench_r1v_2p22_n8.json; tail -3 gpurun_out/bench_r1v_2p22_n8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1v_n8.json 2> gpurun_out/bench_r1v_n8.err; cat gpurun_out/bench_r1v_n8.json; tail -3 gpurun_out/bench_r1v_n8.err
