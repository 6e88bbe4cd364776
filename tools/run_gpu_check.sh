set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6 | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1u.json 2> gpurun_out/bench_r1u.err; cat gpurun_out/bench_r1u.json; tail -3 gpurun_out/bench_r1u.err
timeout 600 python -m typlonk_b200.sweep --msm 16,18,20,22,24,26 --ntt 16,18,20,22,24 > gpurun_out/sweep_r1u.jsonl 2>gpurun_out/sweep_r1u.err; cut -c1-330 gpurun_out/sweep_r1u.jsonl; tail -3 gpurun_out/sweep_r1u.err
