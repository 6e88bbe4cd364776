set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m typlonk_b200.sweep --msm "" --ntt 20,22,24 > gpurun_out/sweep_ntt_r1d.jsonl 2> gpurun_out/sweep_r1d.err; cat gpurun_out/sweep_ntt_r1d.jsonl
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1h.json 2> gpurun_out/bench_r1h.err; cat gpurun_out/bench_r1h.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_r1h_2gpu.json 2> gpurun_out/bench_r1h_2gpu.err; cat gpurun_out/bench_r1h_2gpu.json
