mkdir -p gpurun_out
python -m typlonk_b200.sweep --msm "" --ntt 16,18,20,22,24 --reps 5 > gpurun_out/r2t_ntt_b2_3blk.jsonl 2>/dev/null
python bench.py --no-north-star --no-cpu-baseline --no-sweep --steps 10 --warmup 3 --ab ntt_radix_log=3 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
export TYPLONK_B200_LIB=$PWD/typlonk_b200/lib/libtyplonk_b200_ntt4.so
python -m pytest tests/test_gpu_primitives.py -x -q -k "ntt" > gpurun_out/r2t_pytest_4blk.log 2>&1; tail -2 gpurun_out/r2t_pytest_4blk.log
python -m typlonk_b200.sweep --msm "" --ntt 16,18,20,22,24 --reps 5 > gpurun_out/r2t_ntt_b2_4blk.jsonl 2>/dev/null
python bench.py --no-north-star --no-cpu-baseline --no-sweep --steps 10 --warmup 3 > gpurun_out/r2t_bench_4blk.json 2> gpurun_out/r2t_bench_4blk.err
python - <<'P'
import json
a=[json.loads(l) for l in open("gpurun_out/r2t_ntt_b2_3blk.jsonl")]
b=[json.loads(l) for l in open("gpurun_out/r2t_ntt_b2_4blk.jsonl")]
for x,y in zip(a,b): print(x["sweep"], x["log_n"], x["ms"], y["ms"], x["wide_mul_frac"], y["wide_mul_frac"])
for f in ["r2t_bench","r2t_bench_4blk"]:
    d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
    print(f, d["value"], d["e2e"]["value"], d["parity"]["digest_ok"], d["phases_ms_per_step"]); print(d["ab"]); print(d["standalone"])
P
