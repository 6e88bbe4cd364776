set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2g_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2g_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["parity"]["digest_ok"], d["gpu_launches"], json.dumps(d["phases_ms_per_step"]))
print("  ns", json.dumps(d["north_star"])[:600])
for r in d["sweeps"]:
    if r["sweep"] == "msm": print(r)
PY
tail -5 gpurun_out/r2g_bench.err
