set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python -m typlonk_b200.sweep --msm "" > gpurun_out/sweep_ntt_r1c.jsonl 2> gpurun_out/sweep_r1c.err; cat gpurun_out/sweep_ntt_r1c.jsonl
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; cat gpurun_out/bench_r1g.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_pass -s 7 -c 2 -o gpurun_out/prof_ntt_r1a -f python -m typlonk_b200.sweep --msm "" --ntt 22 --reps 1 > gpurun_out/ncu_ntt.log 2>&1
