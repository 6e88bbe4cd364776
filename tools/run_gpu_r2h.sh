set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-sweep --no-cpu-baseline --no-north-star"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2h_launches_prove_2p20.csv $B > /tmp/ncu_bench.log 2>&1
python tools/ncu_summary.py gpurun_out/r2h_launches_prove_2p20.csv --seq | grep -E "rowcol|bitsums|wsum_level|pair_fixup|combine" | tail -24
for m in 1 2; do
TP_MSM_REDUCE_L1=$m timeout 300 python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline --no-north-star > gpurun_out/r2h_l1mode$m.json 2>> gpurun_out/r2h.err
done
python - <<PY
import json
for m in (1, 2):
    d = json.loads(open("gpurun_out/r2h_l1mode%d.json" % m).read().strip().splitlines()[-1])
    print("l1 mode", m, round(d["value"], 3), d["parity"]["digest_ok"], json.dumps(d["phases_ms_per_step"]))
PY
