# final single-GPU evidence of round 2: launch list, full sets of the kernels that changed, the driver-style bench and reference arm
set -x
mkdir -p gpurun_out /tmp/rep
B="python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu-baseline --no-north-star"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2l_launches_prove_2p20.csv $B > /tmp/rep/ncu_bench.log 2>&1
full() {  # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o /tmp/rep/r2l_$1 $B > /tmp/rep/ncu_$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i /tmp/rep/r2l_$1.ncu-rep --page raw --csv > gpurun_out/r2l_$1_raw.csv 2>/dev/null
}
full reduce 'k_msm_wsum_level|k_msm_treesum|k_msm_pair_fixup' 12 12
full poly 'k_perm_chunks|k_perm_finish|k_quotient_numerator|k_quotient_combine|k_lincomb|k_gate_check|k_linrec' 0 24
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2l_bench_reference.json 2> gpurun_out/r2l_bench_reference.err; echo "ref rc=$?"
cut -c1-900 gpurun_out/r2l_bench_reference.json
python - <<PY
import json
d = json.loads(open("gpurun_out/r2l_bench_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["parity"]["digest_ok"], d["gpu_launches"], json.dumps(d["phases_ms_per_step"]))
print(json.dumps(d["cpu_baseline"]), json.dumps(d["b0_as_written"])[:300])
print(json.dumps(d["roofline"])[:400]); print(json.dumps(d["roofline_int"])[:500])
PY
du -sh gpurun_out
