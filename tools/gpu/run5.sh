mkdir -p gpurun_out
python -m pytest tests/test_gpu_primitives.py -x -q -k "ntt" > gpurun_out/r2s_pytest.log 2>&1; tail -2 gpurun_out/r2s_pytest.log
python -m typlonk_b200.sweep --msm "" --ntt 16,18,20,22,24 --reps 5 > gpurun_out/r2s_ntt_b3.jsonl 2>/dev/null
TP_NTT_B=2 python -m typlonk_b200.sweep --msm "" --ntt 16,18,20,22,24 --reps 5 > gpurun_out/r2s_ntt_b2.jsonl 2>/dev/null
python bench.py --no-north-star --no-cpu-baseline --no-sweep --steps 10 --warmup 3 --ab ntt_radix_log=2 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python - <<'P'
import json
a=[json.loads(l) for l in open("gpurun_out/r2s_ntt_b3.jsonl")]
b=[json.loads(l) for l in open("gpurun_out/r2s_ntt_b2.jsonl")]
for x,y in zip(a,b): print(x["sweep"], x["log_n"], x["ms"], y["ms"], x["wide_mul_frac"], y["wide_mul_frac"])
d=json.loads(open("gpurun_out/r2s_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["parity"]["digest_ok"], d["phases_ms_per_step"]); print(d["ab"])
P
