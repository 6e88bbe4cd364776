set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2c_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-sweep --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
TP_NO_SIDE_STREAM=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-sweep --no-cpu-baseline --no-north-star > gpurun_out/r2c_bench_noside.json 2>> gpurun_out/r2c_bench.err; echo "bench rc=$?"
python - <<PY
import json
for f in ("r2c_bench", "r2c_bench_noside"):
    d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
    print(f, d["value"], d["e2e"]["value"], d["parity"]["digest_ok"], d["gpu_launches"], json.dumps(d["phases_ms_per_step"]))
    if d.get("north_star"): print("  ns", json.dumps(d["north_star"])[:600])
PY
tail -5 gpurun_out/r2c_bench.err
