set -x
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -6 | cut -c1-300
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1w.json 2> gpurun_out/bench_r1w.err; cat gpurun_out/bench_r1w.json; tail -3 gpurun_out/bench_r1w.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1w_ref.json 2> gpurun_out/bench_r1w_ref.err; cat gpurun_out/bench_r1w_ref.json
timeout 600 python bench.py --steps 3 --warmup 2 --log-n 22 --no-cpu-baseline > gpurun_out/bench_r1w_2p22.json 2> gpurun_out/bench_r1w_2p22.err; cat gpurun_out/bench_r1w_2p22.json; tail -3 gpurun_out/bench_r1w_2p22.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1w.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/ncu_summary.py gpurun_out/launches_r1w.csv --last 183 | head -40
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_msm_accumulate -s 5 -c 1 -o gpurun_out/prof_acc_r1w -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_acc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_r8 -s 60 -c 3 -o gpurun_out/prof_ntt_r1w -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_ntt.log 2>&1
timeout 600 python -m typlonk_b200.sweep --msm 16,18,20,22,24,26 --ntt 16,18,20,22,24 > gpurun_out/sweep_r1w.jsonl 2>gpurun_out/sweep_r1w.err; cut -c1-200 gpurun_out/sweep_r1w.jsonl
ls -la gpurun_out | tail -8
