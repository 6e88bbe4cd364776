mkdir -p gpurun_out
for v in "" w4 w5 a4; do
  if [ -n "$v" ]; then export TYPLONK_B200_LIB=$PWD/typlonk_b200/lib/libtyplonk_b200_$v.so; fi
  python bench.py --no-north-star --no-cpu-baseline --no-sweep --steps 10 --warmup 3 > gpurun_out/r2u_bench_$v.json 2> gpurun_out/r2u_bench_$v.err
done
python - <<'P'
import json
for v in ["","w4","w5","a4"]:
    try:
        d=json.loads(open("gpurun_out/r2u_bench_%s.json"%v).read().strip().splitlines()[-1])
        print(v or "default", d["value"], d["parity"]["digest_ok"], d["phases_ms_per_step"])
    except Exception as e: print(v, "ERR", e)
P
