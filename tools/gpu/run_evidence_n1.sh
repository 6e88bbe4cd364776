mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2v_pytest.log 2>&1; tail -2 gpurun_out/r2v_pytest.log
python bench.py > gpurun_out/r2v_bench_n1.json 2> gpurun_out/r2v_bench_n1.err; tail -c 300 gpurun_out/r2v_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2v_launches_prove_2p20.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-sweep --no-north-star > gpurun_out/r2v_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ntt_r8 -s 3 -c 6 -o gpurun_out/r2v_ntt -f \
  python -m typlonk_b200.sweep --msm "" --ntt 20 --reps 1 > gpurun_out/r2v_ncu_ntt.log 2>&1
ls -la gpurun_out/ | tail -8
python - <<'P'
import json
d=json.loads(open("gpurun_out/r2v_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["parity"]["digest_ok"], d["phases_ms_per_step"]); print(d["standalone"]); print(d["north_star"]["prove_ms"], d["north_star"]["e2e_ms"], d["north_star"]["parity"]["digest_ok"]); print(d["cpu_baseline"])
P
