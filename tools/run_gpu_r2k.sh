set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2k_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-sweep > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2k_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["e2e"]["value"], d["parity"]["digest_ok"], d["gpu_launches"], json.dumps(d["phases_ms_per_step"]))
print("  ns", json.dumps(d["north_star"])[:300])
PY
tail -5 gpurun_out/r2k_bench.err
