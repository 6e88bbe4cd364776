set -x
mkdir -p gpurun_out
timeout 600 python bench.py --log-n 16 --no-north-star --no-sweep --steps 20 --warmup 5 > gpurun_out/r2m_bench_2p16_n1.json 2> gpurun_out/r2m.err; echo "rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/r2m_bench_ns_n1.json 2>> gpurun_out/r2m.err; echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2m_bench_2p16_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["parity"], json.dumps(d["phases_ms_per_step"]), json.dumps(d["cpu_baseline"]))
d = json.loads(open("gpurun_out/r2m_bench_ns_n1.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], json.dumps(d["north_star"])[:500])
PY
tail -3 gpurun_out/r2m.err
